"""Exclusive per-stage times of one 64-frame chunk (one stream, CUDA events of the library's profile mode) and the
resident two-stream rate of a 1024-frame batch.  usage: stage_times.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rumi_slam_b200 import ORBextractor
from rumi_slam_b200.synth import synthetic_batch
dev = torch.from_numpy(synthetic_batch(1024, 640, 480, seed0=0, unique=16)).cuda()
ex = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=64)
od = None
for _ in range(3): od = ex.extract_batch_device(dev, out=od)
ex.timer_start()
for _ in range(10): ex.extract_batch_device(dev, out=od, sync=False)
print("resident: %.0f frames/s" % (1024 * 10e3 / ex.timer_stop()))
ex.set_streams(1); ex.profile(True)
ex.extract_batch_device(dev, out=od); ex.profile_read()
for _ in range(3): ex.extract_batch_device(dev, out=od)
r = ex.profile_read()
print("  ".join("%s %.1f us" % (k, v[0] * 1e3 / max(v[1], 1)) for k, v in r.items() if v[1]))
