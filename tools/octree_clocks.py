"""Per-phase cycle counters of the quad-tree kernel (frame 0, every level) for one synthetic chunk."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rumi_slam_b200 import ORBextractor
from rumi_slam_b200.synth import synthetic_batch
frames = torch.from_numpy(synthetic_batch(8, 640, 480, seed0=0)).cuda()
ex = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=8)
ex.extract_batch_device(frames)
ex._L.rumi_orb_debug_octree_clocks(ex._h, None, 0)
ex.extract_batch_device(frames)
out = np.zeros(16 * 16, np.int64)
ex._L.rumi_orb_debug_octree_clocks(ex._h, out.ctypes.data, out.size)
names = ["keys", "sortK", "hist", "groups", "sortL", "pendinit", "R1", "serial", "R4", "final", "gather", "best"]
print("level " + " ".join("%8s" % n for n in names) + "        M     nout")
for l in range(8):
    r = out[16 * l:16 * l + 16]
    print("%5d " % l + " ".join("%8d" % v for v in r[:12]) + " %8d %8d" % (r[12], r[13]))
