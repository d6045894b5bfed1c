"""Resident-input throughput of the extraction only (no CPU baseline, no matching): quick A/B runs on the GPU box.
usage: quick_bench.py [frames] [chunk] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rumi_slam_b200 import ORBextractor
from rumi_slam_b200.synth import synthetic_batch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 64
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
frames = torch.from_numpy(synthetic_batch(n, 640, 480, seed0=0, unique=32)).cuda()
ex = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=chunk)
out = None
for _ in range(3):
    out = ex.extract_batch_device(frames, out=out, sync=True)
ex.timer_start()
for _ in range(steps):
    ex.extract_batch_device(frames, out=out, sync=False)
ms = ex.timer_stop()
print("frames/s %.0f  ms/step %.3f  (chunk %d, env %s)" % (n * steps / ms * 1e3, ms / steps, chunk,
      {k: v for k, v in os.environ.items() if k.startswith("RUMI_")}))
