"""Per-phase cycle counters of the quad-tree kernel on ONE frame of growing FAST-candidate density (see density_probe.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rumi_slam_b200 import ORBextractor
from rumi_slam_b200.synth import synthetic_frame
base = synthetic_frame(3, 640, 480).astype(np.int32)
rng = np.random.default_rng(0)
names = ["keys", "sortK", "hist", "groups", "sortL", "pendinit", "R1", "serial", "R4", "final", "gather", "best"]
for amp in (0, 10, 20, 30):
    img = np.clip(base + rng.integers(-amp, amp + 1, base.shape), 0, 255).astype(np.uint8)
    ex = ORBextractor(1000, 1.2, 8, 20, 7)
    ex(img)
    ex._L.rumi_orb_debug_octree_clocks(ex._h, None, 0)
    ex(img)
    out = np.zeros(16 * 16, np.int64)
    ex._L.rumi_orb_debug_octree_clocks(ex._h, out.ctypes.data, out.size)
    print("noise +-%d" % amp)
    print("level " + " ".join("%8s" % n for n in names) + "        M     nout")
    for l in range(3):
        r = out[16 * l:16 * l + 16]
        print("%5d " % l + " ".join("%8d" % v for v in r[:12]) + " %8d %8d" % (r[12], r[13]))
    ex.close()
