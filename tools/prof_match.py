"""Profiling workload for the top-2 kernels: 16384 x 40000 real-ish descriptors, kernel chosen by RUMI_MATCH."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rumi_slam_b200 import ORBmatcher
g = torch.Generator(device="cuda"); g.manual_seed(1)
Q = torch.randint(0, 256, (16384, 32), dtype=torch.uint8, device="cuda", generator=g)
T = torch.randint(0, 256, (40000, 32), dtype=torch.uint8, device="cuda", generator=g)
m = ORBmatcher()
for _ in range(3):
    m.top2_device(Q, T)
print("ok")
