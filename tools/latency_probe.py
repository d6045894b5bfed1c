"""Single-frame latency of ORBextractor::operator() (host image in, keypoints + descriptors out), BASELINE configs[0]."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rumi_slam_b200 import ORBextractor
from rumi_slam_b200.synth import synthetic_frame
img = synthetic_frame(0, 640, 480)
ex = ORBextractor(1000, 1.2, 8, 20, 7)
for _ in range(20):
    ex(img)
ts = []
for _ in range(200):
    t0 = time.perf_counter(); ex(img); ts.append(time.perf_counter() - t0)
ts = np.array(ts) * 1e3
print("operator() latency ms: median %.3f p10 %.3f p90 %.3f  (%.0f frames/s single stream of single frames)" % (
    np.median(ts), np.percentile(ts, 10), np.percentile(ts, 90), 1e3 / np.median(ts)))
ex.profile(True); ex.profile_read(True)
for _ in range(50):
    ex(img)
st = ex.profile_read(True)
print({k: round(v[0] / 50 * 1e3, 1) for k, v in st.items()}, "us per frame per stage (device)")
