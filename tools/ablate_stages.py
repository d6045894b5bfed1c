"""Marginal cost of every stage inside the pipelined batch: time the 1024-frame resident batch with one stage not
launched (stale data of the previous full run keeps the other stages' work realistic).  Timing experiment only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rumi_slam_b200 import ORBextractor
from rumi_slam_b200.synth import synthetic_batch
frames = torch.from_numpy(synthetic_batch(1024, 640, 480, seed0=0, unique=32)).cuda()
ex = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=64)
out = None
def run(mask, steps=10):
    global out
    ex._L.rumi_orb_debug_skip_stages(ex._h, 0)
    for _ in range(2):
        out = ex.extract_batch_device(frames, out=out, sync=True)
    ex._L.rumi_orb_debug_skip_stages(ex._h, mask)
    ex.extract_batch_device(frames, out=out, sync=True)
    ex.timer_start()
    for _ in range(steps):
        ex.extract_batch_device(frames, out=out, sync=False)
    return ex.timer_stop() / steps
base = run(0)
print("all stages: %.3f ms/step" % base)
for name, bit in (("pyramid", 1), ("fast", 2), ("octree", 4), ("blur", 16), ("describe", 32)):
    t = run(bit)
    print("without %-8s: %.3f ms/step  (marginal %.3f ms = %.0f us per 64-frame chunk)" % (name, t, base - t, (base - t) / 16 * 1e3))
