import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
from rumi_slam_b200 import ORBextractor
h, w = 480, 640
n = h * w
# LCG sequence of tools/probe/adapter_latency.cc
s = np.empty(n, np.uint64); v = 12345
a, c = 1664525, 1013904223
vals = []
for i in range(n):
    v = (v * a + c) & 0xFFFFFFFF
    vals.append(v >> 28)
noise = np.array(vals, np.uint8).reshape(h, w)
yy, xx = np.mgrid[0:h, 0:w]
img = (np.where(((xx // 16 + yy // 16) % 2) == 1, 160, 60) + noise).astype(np.uint8)
ex = ORBextractor(1000, 1.2, 8, 20, 7)
for _ in range(20): ex(img)
ts = []
for _ in range(200):
    t0 = time.perf_counter(); m, k, d = ex(img); ts.append(time.perf_counter() - t0)
print("python mirror on the same checkerboard image: median %.3f ms, %d keypoints" % (np.median(ts) * 1e3, len(k)))
ex.profile(True); ex.profile_read(True)
for _ in range(50): ex(img)
st = ex.profile_read(True)
print({k: round(v[0] / 50 * 1e3, 1) for k, v in st.items()})
