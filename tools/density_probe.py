"""Stage times of one frame vs FAST candidate density (synthetic frame + noise of growing amplitude): where the quad-tree
leaves its shared-memory sort path (2048 keys per level) and what that costs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rumi_slam_b200 import ORBextractor
from rumi_slam_b200.synth import synthetic_frame
base = synthetic_frame(3, 640, 480).astype(np.int32)
rng = np.random.default_rng(0)
ex = ORBextractor(1000, 1.2, 8, 20, 7)
for amp in (0, 6, 10, 14, 20, 30, 60):
    img = np.clip(base + rng.integers(-amp, amp + 1, base.shape), 0, 255).astype(np.uint8)
    for _ in range(5): ex(img)
    n0 = len(ex.debug_candidates(0)); n1 = len(ex.debug_candidates(1))
    ex.profile(True); ex.profile_read(True)
    for _ in range(30): ex(img)
    st = ex.profile_read(True)
    ex.profile(False)
    print("noise +-%2d: candidates level0 %6d level1 %6d | us: %s" % (amp, n0, n1, " ".join("%s %.0f" % (k, v[0] / 30 * 1e3) for k, v in st.items() if v[0] > 0)))

# the same noise on a resident 256-frame batch: what dense frames cost the throughput path
import torch
from rumi_slam_b200.synth import synthetic_batch
bx = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=64)
base_b = synthetic_batch(256, 640, 480, seed0=0, unique=16).astype(np.int16)
for amp in (0, 10, 20):
    noisy = np.clip(base_b + rng.integers(-amp, amp + 1, base_b.shape, dtype=np.int16), 0, 255).astype(np.uint8)
    dev = torch.from_numpy(noisy).cuda()
    od = None
    for _ in range(3): od = bx.extract_batch_device(dev, out=od)
    bx.timer_start()
    for _ in range(10): bx.extract_batch_device(dev, out=od, sync=False)
    print("noise +-%2d: resident batch %.0f frames/s" % (amp, 256 * 10e3 / bx.timer_stop()))
