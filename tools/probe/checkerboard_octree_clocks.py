import sys
sys.path.insert(0, "/root/repo")
import numpy as np
from rumi_slam_b200 import ORBextractor
h, w = 480, 640
rng = np.random.default_rng(1)
noise = rng.integers(0, 16, (h, w)).astype(np.uint8)
yy, xx = np.mgrid[0:h, 0:w]
img = (np.where(((xx // 16 + yy // 16) % 2) == 1, 160, 60) + noise).astype(np.uint8)
names = ["keys", "sortK", "hist", "groups", "sortL", "pendinit", "R1", "serial", "R4", "final", "gather", "best"]
ex = ORBextractor(1000, 1.2, 8, 20, 7)
ex(img)
ex._L.rumi_orb_debug_octree_clocks(ex._h, None, 0)
ex(img)
out = np.zeros(16 * 16, np.int64)
ex._L.rumi_orb_debug_octree_clocks(ex._h, out.ctypes.data, out.size)
print("level " + " ".join("%8s" % n for n in names) + "        M     nout  replay")
for l in range(8):
    r = out[16 * l:16 * l + 16]
    print("%5d " % l + " ".join("%8d" % v for v in r[:12]) + " %8d %8d %8d" % (r[12], r[13], r[14]))
