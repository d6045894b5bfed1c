"""Profiling workload for the extraction kernels: 64-frame chunks of synthetic 640x480 frames through the
device-resident batch call on ONE stream (so that every kernel runs alone).  usage: prof_extract.py [frames] [passes]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rumi_slam_b200 import ORBextractor
from rumi_slam_b200.synth import synthetic_batch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 4
frames = torch.from_numpy(synthetic_batch(n, 640, 480, seed0=0, unique=16)).cuda()
ex = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=n)
ex.set_streams(1)
out = None
for _ in range(passes):
    out = ex.extract_batch_device(frames, out=out, sync=True)
print("ok", int(out[2].sum()))
