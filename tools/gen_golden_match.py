"""Generates tests/golden/match_functions.npz: what the UNMODIFIED reference matcher functions (oracle/_ref/librefframe.so, cut
from ORBmatcher.cc / Frame.cc at build time by `make -C oracle refframe`) return on seeded scenes -- the scenes of
tests/test_ref_frame_pin.py, rebuilt by the same helper functions.  tests/test_golden_match.py compares the oracle with these
frozen outputs, so the pin also holds where the reference sources are absent (the GPU box).  Run in the dev container."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from oracle import orb_oracle as oracle, ref_frame_lib as rf, bow_oracle as B
import golden_match_cases as G

oracle.build(); B.build()
out = {}
for name, fn in G.CASES.items():
    res = fn(oracle, B, rf, reference=True)
    for k, v in res.items():
        out["%s/%s" % (name, k)] = np.asarray(v)
path = os.path.join(ROOT, "tests", "golden", "match_functions.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")
