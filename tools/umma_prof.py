"""ncu workload for the tcgen05 top-2 kernel: 40000 x 40000 descriptors, a few launches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
os.environ["RUMI_MATCH"] = "umma"
from rumi_slam_b200 import ORBmatcher
from rumi_slam_b200.synth import perturbed_descriptors
rng = np.random.default_rng(1)
T = torch.from_numpy(rng.integers(0, 256, (40000, 32), dtype=np.uint8)).cuda()
Q = torch.from_numpy(perturbed_descriptors(T.cpu().numpy(), 40000, seed=3, flip_p=0.08)).cuda()
m = ORBmatcher()
for _ in range(3):
    m.top2_device(Q, T)
print("ok", m.last_path())
