#!/usr/bin/env python3
"""Generates tests/golden/*.npz -- golden vectors for the hot path, produced in the DEV CONTAINER from
(a) the OpenCV primitives the reference calls, through cv2 (oracle/cv2_oracle.py), and
(b) the UNMODIFIED reference ORBextractor.cc compiled over oracle/cvstub (oracle/_ref/liborb_ref.so).
The reference ships no golden vectors of its own for this path; these pin the oracle and travel
to the GPU box, where neither cv2's presence nor /root/reference is assumed.
"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cv2_oracle as A, orb_oracle as O, ref_lib as R
from rumi_slam_b200.synth import synthetic_frame

assert A.HAVE_CV2 and R.available(), "needs cv2 and /root/reference (dev container)"
out = os.path.join(ROOT, "tests", "golden")
os.makedirs(out, exist_ok=True)
rng = np.random.default_rng(2026)

# 1. cv2 primitives on small inputs
img = synthetic_frame(77, 200, 152)
ws, hs = O.level_sizes(200, 152, 1.2, 4)
pyr = A.pyramid(img, ws, hs)
noise = rng.integers(0, 256, (60, 72), dtype=np.uint8)
patch = np.ascontiguousarray(synthetic_frame(78, 96, 80))
yx = rng.integers(-300000, 300000, (4000, 2)).astype(np.int32)
yx[:8] = [[0, 0], [0, 5], [5, 0], [-5, 0], [0, -5], [7, 7], [-7, 7], [7, -7]]
at = np.array([A.fast_atan2(y, x) for y, x in yx], np.float32)
np.savez_compressed(os.path.join(out, "cv2_primitives.npz"),
                    img=img, ws=ws, hs=hs, pyr1=pyr[1], pyr2=pyr[2], pyr3=pyr[3],
                    noise=noise, noise_fast20=A.fast(noise, 20), noise_fast7=A.fast(noise, 7),
                    patch=patch, patch_fast20=A.fast(patch, 20), patch_fast7=A.fast(patch, 7),
                    patch_blur=A.blur(patch), noise_blur=A.blur(noise),
                    grid_img=pyr[1], grid_cand=A.grid_fast(pyr[1])[0],
                    atan_yx=yx, atan_deg=at)

# 2. unmodified reference extractor on two small frames (full operator() output) + one octree problem
cases = {}
for name, (w, h, nf, nl, lap, seed) in {"a": (320, 240, 500, 4, (0, 0), 5), "b": (400, 300, 300, 5, (0, 1000), 6),
                                        "c": (360, 200, 400, 3, (100, 250), 7)}.items():
    im = synthetic_frame(seed, w, h)
    k, d, m = R.extract(im, nfeatures=nf, nlevels=nl, lapping=lap)
    cases.update({"img_" + name: im, "kps_" + name: k.view(np.uint8).reshape(-1, 28), "desc_" + name: d,
                  "meta_" + name: np.array([nf, nl, lap[0], lap[1], m], np.int32)})
cand, _ = O.grid_fast(synthetic_frame(9, 640, 480))
sel = R.octree(cand, 16, 640 - 16, 16, 480 - 16, 217)
np.savez_compressed(os.path.join(out, "reference_extract.npz"), oct_cand=cand, oct_sel=sel, **cases)

# 3. matcher known-answer vectors (semantics of ORBmatcher.cc:236-261 / BFMatcher.knnMatch k=2)
zeros, ones = np.zeros((1, 32), np.uint8), np.full((1, 32), 255, np.uint8)
one = zeros.copy(); one[0, 7] = 0x10
desc = cases["desc_a"]
Q, T = desc[:200].copy(), desc[100:500].copy()
T[250] = T[3]
bi, bd1, bd2 = A.knn2(Q, T)
np.savez_compressed(os.path.join(out, "matcher_kats.npz"), Q=Q, T=T, bf_d1=bd1, bf_d2=bd2, bf_idx=bi,
                    kat_T=np.concatenate([ones, one, one, zeros]), kat_expect=np.array([3, 0, 1], np.int32))
print("golden written:", sorted(os.listdir(out)), sum(os.path.getsize(os.path.join(out, f)) for f in os.listdir(out)), "bytes")
