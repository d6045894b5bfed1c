import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["RUMI_NO_TMA"] = "1"
import numpy as np
from rumi_slam_b200 import ORBextractor
from rumi_slam_b200.synth import synthetic_frame
from oracle import orb_oracle as O
img = synthetic_frame(11)
ex = ORBextractor(1000, 1.2, 8, 20, 7)
ex(img)
ref = O.pyramid(img)
for l in (0, 7):
    cand, nfb = O.grid_fast(ref[l]); g = ex.debug_candidates(l)
    c7, _ = O.grid_fast(ref[l], 7, 7)
    S = set(map(tuple, cand.astype(int))); G = set(map(tuple, g)); S7 = set(map(tuple, c7.astype(int)))
    print("level", l, "oracle", len(S), "gpu", len(g), "gpu unique", len(G), "oracle th7", len(S7))
    print(" gpu&oracle", len(G & S), "gpu&th7", len(G & S7), "gpu resp<20:", int((g[:, 2] < 20).sum()), "oracle resp<20", int((cand[:,2]<20).sum()))
    only = sorted(G - S7)[:10]; print(" gpu not in th7:", only)
    miss = sorted(S - G)[:10]; print(" oracle not in gpu:", miss)
    sm = O.fast_score_map(ref[l])
    for (x, y, r) in only[:5]:
        print("   score map at", x + 16, y + 16, "=", sm[y + 16, x + 16], "gpu resp", r)
