"""Latency of the per-frame matcher calls with HOST pointers (the sizes a SLAM front-end issues every frame)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rumi_slam_b200 import ORBmatcher
rng = np.random.default_rng(0)
m = ORBmatcher()
def lat(fn, n=200):
    for _ in range(20): fn()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return np.median(ts) * 1e3
for nq, nt in ((1000, 1000), (1200, 1200), (2000, 2000), (5000, 5000)):
    Q = rng.integers(0, 256, (nq, 32), dtype=np.uint8); T = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
    print("top2 %5d x %5d host pointers: %.3f ms" % (nq, nt, lat(lambda: m.top2(Q, T))))
# candidate lists: 1000 queries x 20 candidates
Q = rng.integers(0, 256, (1000, 32), dtype=np.uint8); T = rng.integers(0, 256, (1200, 32), dtype=np.uint8)
off = (np.arange(1001) * 20).astype(np.int32); idx = rng.integers(0, 1200, 20000).astype(np.int32)
print("candidates 1000 x 20: %.3f ms" % lat(lambda: m.candidates(Q, T, off, idx)))
print("candidates 1000 x 20 + top2: %.3f ms" % lat(lambda: m.candidates(Q, T, off, idx, top2=True)))
