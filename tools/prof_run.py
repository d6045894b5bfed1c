"""Short profiling workload: one 64-frame chunk of synthetic 640x480 frames through the device-resident batch
call (2 warm-up + 2 profiled passes) and one 8192 x 40000 top-2 match.  Used under ncu (see profiles/README.md)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rumi_slam_b200 import ORBextractor, ORBmatcher
from rumi_slam_b200.synth import synthetic_batch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
frames = torch.from_numpy(synthetic_batch(n, 640, 480, seed0=0, unique=16)).cuda()
ex = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=n)
out = None
for _ in range(4):
    out = ex.extract_batch_device(frames, out=out, sync=True)
desc = out[1][:, :1000].reshape(-1, 32)
Q, T = desc[:8192].contiguous(), desc[:40000].contiguous()
m = ORBmatcher()
for _ in range(2):
    m.top2_device(Q, T)
print("ok", int(out[2].sum()))
