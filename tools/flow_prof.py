"""Short profiling workload of the optical-flow path: a translating 640x480 sequence, 1000 extracted keypoints tracked
frame to frame (KFDSample::Step's steady state).  Used under ncu; prints host-side latency otherwise.
usage: flow_prof.py [frames] [points]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rumi_slam_b200 import ORBextractor, SparsePyrLK
from rumi_slam_b200.synth import motion_sequence

n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
npts = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
seq = motion_sequence(n, 640, 480, seed=500)
ex = ORBextractor(max(npts, 1000), 1.2, 8, 20, 7)
_, kp, _ = ex(seq[0])
pts = np.stack([kp["x"], kp["y"]], 1).astype(np.float32)[:npts]
lk = SparsePyrLK()
for rep in range(2):
    lk.set_prev(seq[0])
    cur, ts, tracked = pts, [], 0
    for i in range(1, n):
        t0 = time.perf_counter()
        cur, st, _ = lk.track_next(seq[i], cur, advance=True)
        ts.append(time.perf_counter() - t0)
        tracked += int(st.sum())
print("points %d  tracked %.3f  median call %.3f ms  min %.3f ms" % (len(pts), tracked / ((n - 1) * len(pts)),
      np.median(ts) * 1e3, np.min(ts) * 1e3))
