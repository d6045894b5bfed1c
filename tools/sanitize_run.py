"""Small end-to-end workload for compute-sanitizer memcheck: single frame, 3-frame ragged batch (device + host paths),
unaligned device input, noise frame, top-2 match (LOP3+POPC and tcgen05), key-frame pair association, stereo match,
optical flow with points outside the frame."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rumi_slam_b200 import ORBextractor, ORBmatcher
from rumi_slam_b200.synth import synthetic_batch, stereo_pair
fr = synthetic_batch(3, 640, 480, seed0=5)
ex = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=2)
m0, k0, d0 = ex(fr[0])
ex.extract_batch(fr)
dev = torch.from_numpy(fr).cuda()
ex.extract_batch_device(dev)
pad = torch.zeros((3, 480, 643), dtype=torch.uint8, device="cuda"); pad[:, :, :640] = dev
ex.extract_batch_device(pad[:, :, :640])
rng = np.random.default_rng(0)
ex(rng.integers(0, 256, (480, 640), dtype=np.uint8))
l, r = stereo_pair(1)
exl, exr = ORBextractor(1200, 1.2, 8, 20, 7), ORBextractor(1200, 1.2, 8, 20, 7)
_, lk, ld = exl(l); _, rk, rd = exr(r)
m = ORBmatcher()
m.top2(ld, rd)
u, d, n = m.stereo_match(exl, exr, lk.copy(), ld.copy(), rk.copy(), rd.copy(), 47.9, 47.9 / 435.2)
# tcgen05 top-2 (forced: the problem is small), key-frame pair association, optical flow on an odd-sized frame
os.environ["RUMI_MATCH"] = "umma"
mu = ORBmatcher()
os.environ.pop("RUMI_MATCH")
Q = rng.integers(0, 256, (700, 32), dtype=np.uint8); T = rng.integers(0, 256, (1333, 32), dtype=np.uint8)
iu = mu.top2(Q, T); ip = m.top2(Q, T)
assert all(np.array_equal(a, b) for a, b in zip(iu, ip)) and mu.last_path() == "umma"
m.top2_pairs([ld, rd[:100], ld[:0]], [rd, ld, rd[:7]])
from rumi_slam_b200 import SparsePyrLK
from rumi_slam_b200.synth import motion_sequence
seq = motion_sequence(3, 333, 257, seed=2)
pts = np.stack([rng.uniform(-5, 338, 300), rng.uniform(-5, 262, 300)], 1).astype(np.float32)
flow = SparsePyrLK()
flow.set_prev(seq[0])
for i in (1, 2):
    pts, st, _ = flow.track_next(seq[i], pts, advance=True)
print("sanitize workload ok", len(k0), n, int(st.sum()))
