import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["RUMI_NO_TMA"] = "1"
import numpy as np, ctypes as C
from rumi_slam_b200 import ORBextractor
from rumi_slam_b200.synth import synthetic_frame
from oracle import orb_oracle as O
img = synthetic_frame(11)
ex = ORBextractor(1000, 1.2, 8, 20, 7)
L = ex._L
L.rumi_orb_debug_fast_tile(ex._h, 0, None, 0, None)
ex(img)
buf = np.zeros(1 << 16, np.uint8); dims = np.zeros(5, np.int32)
nb = L.rumi_orb_debug_fast_tile(ex._h, 0, buf.ctypes.data, buf.size, dims.ctypes.data)
tp, tr, sp, sr, mw = map(int, dims)
print("dims", tp, tr, sp, sr, mw, nb)
tile = buf[:tp * tr].reshape(tr, tp); score = buf[tp * tr: tp * tr + sp * sr].reshape(sr, sp)
# cell 0 of level 0: iniX=iniY=16, wCell=36,hCell=38 -> sub-image 42x44
sub = img[16:16 + 44, 16:16 + 42]
print("tile equal:", np.array_equal(tile[:44, :42], sub))
if not np.array_equal(tile[:44, :42], sub):
    d = np.argwhere(tile[:44, :42] != sub); print(" first diffs", d[:10], len(d))
sm = O.fast_score_map(img)[19:19 + 38, 19:19 + 36]
is7 = sm >= 7
got = score[1:39, 1:37].astype(int)
print("score nonzero gpu", (got > 0).sum(), "oracle>=7", is7.sum())
print("score equal where oracle>=7:", np.array_equal(got[is7], sm[is7]), "gpu nonzero only where oracle>=7:", ((got > 0) & ~is7).sum())
bad = np.argwhere((got != np.where(is7, sm, 0)))
print("mismatch count", len(bad), bad[:10])
for (y, x) in bad[:5]: print("  at", y, x, "gpu", got[y, x], "oracle", sm[y, x])
print("score frame zero:", score[0, :38].sum(), score[:, 0].sum(), score[39, :38].sum(), score[:40, 37].sum())
