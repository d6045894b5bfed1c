"""Generates tests/golden/flow_kats.npz: what cv2 (OpenCV 4.13.0, the same OpenCV code the reference links, through
another binding) returns for cv::calcOpticalFlowPyrLK as KFDSample::Step calls it (winSize 31, maxLevel 2,
TermCriteria(COUNT+EPS, 20, 0.03)), for cv::pyrDown and for the Scharr derivatives, on small synthetic frame pairs.
Run in the dev container (needs cv2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2
import numpy as np
from rumi_slam_b200.synth import synthetic_batch

cv2.setNumThreads(1)
crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 20, 0.03)
out = {}
rng = np.random.default_rng(11)
cases = [("a", 320, 240, 300, [[1.01, 0.012, 2.3], [-0.009, 0.995, -1.6]]),      # 3 levels
         ("b", 160, 120, 120, [[1.0, 0.0, 1.4], [0.0, 1.0, 0.7]]),               # pyramid stops at level 1 (40x30 <= 31)
         ("c", 211, 173, 150, [[0.99, -0.02, -3.1], [0.015, 1.005, 4.2]])]       # odd sizes, larger motion
for name, w, h, n, M in cases:
    prev = synthetic_batch(1, w, h, seed0=70 + len(out))[0]
    nxt = cv2.warpAffine(prev, np.array(M, np.float32), (w, h), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101)
    nxt = np.clip(nxt.astype(int) + rng.integers(-2, 3, nxt.shape), 0, 255).astype(np.uint8)
    pts = np.stack([rng.uniform(0, w - 1, n), rng.uniform(0, h - 1, n)], 1).astype(np.float32)
    pts[:8] = [[0, 0], [w - 1, h - 1], [0.5, 0.5], [w - 1.5, 3], [5, h - 1], [w - 1, 0], [15.25, 15.75], [w / 2, h / 2]]
    p1, st, err = cv2.calcOpticalFlowPyrLK(prev, nxt, pts.reshape(-1, 1, 2), None, winSize=(31, 31), maxLevel=2, criteria=crit)
    err = err.ravel().copy()
    err[st.ravel() == 0] = 0                        # cv2 leaves err unset where the flow was not found
    out.update({name + "_prev": prev, name + "_next": nxt, name + "_pts": pts, name + "_cv_next": p1.reshape(-1, 2),
                name + "_cv_status": st.ravel(), name + "_cv_err": err})
    l1 = cv2.pyrDown(prev)
    out[name + "_pyr1"] = l1
    out[name + "_pyr2"] = cv2.pyrDown(l1)
    if name == "b":                                 # (kept small: the derivative image of one case)
        out[name + "_scharr"] = np.stack([cv2.Scharr(prev, cv2.CV_16S, 1, 0, borderType=cv2.BORDER_REFLECT_101),
                                          cv2.Scharr(prev, cv2.CV_16S, 0, 1, borderType=cv2.BORDER_REFLECT_101)], -1)
    print(name, w, h, "tracked", int(st.sum()), "of", n)
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "flow_kats.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes")
