"""Generates tests/golden/bow_kats.npz from the UNMODIFIED reference DBoW2 (oracle/_ref/librefbow.so, built from
/root/reference by `make -C oracle refbow`): a small synthetic vocabulary, features, and what the reference's own
loadFromTextFile + transform return for them.  Run in the dev container (the reference is absent on the GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import bow_oracle as B
from rumi_slam_b200.synth import synthetic_vocabulary, descriptors_near_vocabulary

k, L = 6, 4
par, leaf, desc, w = synthetic_vocabulary(k, L, seed=42, stop_every=11)
R = B.ReferenceVocabulary(k, L, par, leaf, desc, w, 0, 0)
rng = np.random.default_rng(43)
f = np.concatenate([descriptors_near_vocabulary(desc, leaf, 300, 44), rng.integers(0, 256, (100, 32), dtype=np.uint8)])
out = dict(k=k, L=L, parent=par, is_leaf=leaf, desc=desc, weight=w, features=f)
for lu in (0, 2, 4):
    word, weight, node = R.transform(f, lu)
    out["word_%d" % lu], out["weight_%d" % lu], out["node_%d" % lu] = word, weight, node
bow, fv = R.vectors(f, 2)
out["bow_ids"] = np.array(sorted(bow), np.int32)
out["bow_vals"] = np.array([bow[i] for i in sorted(bow)])
out["fv_nodes"] = np.array(sorted(fv), np.int32)
out["fv_idx"] = np.concatenate([np.array(fv[n], np.int32) for n in sorted(fv)])
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "bow_kats.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes;", len(par), "nodes,", len(f), "features")
