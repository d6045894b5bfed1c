import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
mode = sys.argv[1] if len(sys.argv) > 1 else "all"
if mode == "all":
    for m in ("notma", "tma"):
        r = subprocess.run([sys.executable, __file__, m], capture_output=True, text=True, timeout=300)
        print("=====", m, "rc", r.returncode); print(r.stdout[-3000:]); print(r.stderr[-2000:])
    sys.exit(0)
os.environ["RUMI_NO_TMA"] = "1" if mode == "notma" else "0"
import numpy as np
from rumi_slam_b200 import ORBextractor
from rumi_slam_b200.synth import synthetic_frame
from oracle import orb_oracle as O
img = synthetic_frame(11)
ex = ORBextractor(1000, 1.2, 8, 20, 7)
try:
    mono, kps, desc = ex(img)
except Exception as e:
    print("extract failed:", e); sys.exit(1)
ref = O.pyramid(img); got = ex.mvImagePyramid
for l in range(8):
    print("pyr", l, got[l].shape, np.array_equal(got[l], ref[l]), int((got[l].astype(int)-ref[l]).__abs__().max()))
bl = ex.blurred_pyramid()
for l in range(8): print("blur", l, np.array_equal(bl[l], O.blur(ref[l])))
tb = O.tables()
for l in range(8):
    cand, nfb = O.grid_fast(ref[l]); g = ex.debug_candidates(l)
    print("fast", l, len(cand), len(g), np.array_equal(g, cand.astype(np.int32)))
    lh, lw = ref[l].shape
    sel = O.octree(cand, 16, lw-16, 16, lh-16, int(tb["quota"][l])); gs = ex.debug_candidates(l, True)
    print("oct", l, len(sel), len(gs), np.array_equal(gs, cand[sel].astype(np.int32)))
rk, rd, rm = O.extract(img)
print("n", len(kps), len(rk), mono, rm)
if len(kps)==len(rk):
    for f in rk.dtype.names: print(f, np.array_equal(kps[f], rk[f]))
    print("desc", np.array_equal(desc, rd), (desc!=rd).any(1).sum())
