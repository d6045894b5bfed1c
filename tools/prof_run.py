"""Short profiling workload: one 64-frame chunk of synthetic 640x480 frames through the device-resident batch
call (2 warm-up + 2 profiled passes), 8192 x 40000 top-2 on the LOP3+POPC kernel, 16384 x 40000 on the tcgen05 kernel, and 4 optical-flow steps of 1000 points.  Used under ncu (see profiles/README.md)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rumi_slam_b200 import ORBextractor, ORBmatcher
from rumi_slam_b200.synth import synthetic_batch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
frames = torch.from_numpy(synthetic_batch(n, 640, 480, seed0=0, unique=16)).cuda()
ex = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=n)
out = None
LIGHT = os.environ.get("PROF_LIGHT") == "1"            # under `ncu --set full`: one warm-up + one profiled pass of everything
for _ in range(2 if LIGHT else 4):
    out = ex.extract_batch_device(frames, out=out, sync=True)
desc = out[1][:, :1000].reshape(-1, 32)
Q, T = desc[:8192].contiguous(), desc[:40000].contiguous()
def matcher(mode):
    os.environ["RUMI_MATCH"] = mode
    mm = ORBmatcher()
    os.environ.pop("RUMI_MATCH")
    return mm
m = matcher("popc")                                  # K8: LOP3 + POPC
for _ in range(1 if LIGHT else 2):
    m.top2_device(Q, T)
Q2 = desc[:16384].contiguous()
for mode in ("umma",):                               # K8-U (tcgen05 + TMEM), 16384 x 40000
    mm = matcher(mode)
    for _ in range(1 if LIGHT else 2):
        mm.top2_device(Q2, T)
# optical flow of the key-frame sampler: 1000 keypoints tracked over 4 frames of a translating sequence
import numpy as np
from rumi_slam_b200 import SparsePyrLK
from rumi_slam_b200.synth import motion_sequence
seq = motion_sequence(5, 640, 480, seed=500)
gx, gy = np.meshgrid(np.linspace(20, 620, 40), np.linspace(20, 460, 25))      # (no extra extraction launches here)
pts = np.stack([gx.ravel(), gy.ravel()], 1).astype(np.float32)
lk = SparsePyrLK()
lk.set_prev(seq[0])
for i in range(1, 3 if LIGHT else 5):
    pts, st, _ = lk.track_next(seq[i], pts, advance=True)
print("ok", int(out[2].sum()), int(st.sum()))
