"""Quick A/B timing of the all-pairs top-2 (cfg 5a 40000^2 and a 262144 x 1000000 slice of cfg 5b) on real-ish descriptors
(extracted descriptors tiled with 10 % of the bits flipped).  usage: match_bench.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rumi_slam_b200 import ORBextractor, ORBmatcher
from rumi_slam_b200.synth import synthetic_batch
frames = torch.from_numpy(synthetic_batch(80, 640, 480, seed0=0, unique=16)).cuda()
ex = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=40)
_, desc, cnt, _ = ex.extract_batch_device(frames)
rows = torch.cat([desc[i, :min(int(c), 1000)] for i, c in enumerate(cnt.tolist())])
gen = torch.Generator(device="cuda"); gen.manual_seed(7)
def tiled(n):
    out = rows.repeat(-(-n // rows.shape[0]), 1)[:n].clone()
    mask = torch.zeros_like(out)
    for bit in range(8):
        mask |= (torch.rand(out.shape, device="cuda", generator=gen) < 0.1).to(torch.uint8) << bit
    return (out ^ mask).contiguous()
m = ORBmatcher()
for nq, nt, steps in ((40000, 40000, 20), (262144, 1000000, 2)):
    Q, T = (rows[:nq].contiguous(), rows[nq:nq + nt].flip(0).contiguous()) if nq + nt <= rows.shape[0] else (tiled(nq), tiled(nt))
    out = None
    for _ in range(2):
        out = m.top2_device(Q, T, out=out, sync=True)
    m.timer_start()
    for _ in range(steps):
        m.top2_device(Q, T, out=out, sync=False)
    ms = m.timer_stop() / steps
    print("%d x %d: %.3f ms  %.3e pairs/s  (%s, %.0f TOP/s)" % (nq, nt, ms, nq * float(nt) / ms * 1e3, m.last_path(), nq * float(nt) * 512 / ms * 1e3 / 1e12))
