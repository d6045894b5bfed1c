"""Host batch path on PAGEABLE memory: frames/s vs the number of staging threads (RUMI_COPY_THREADS) and vs the driver's own
staging (RUMI_NO_HOST_STAGING=1)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rumi_slam_b200 import ORBextractor, KP_DTYPE
from rumi_slam_b200.synth import synthetic_batch
n = 1024
host = synthetic_batch(n, 640, 480, seed0=0, unique=16)
ex = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=64)
cap = ex.frame_capacity(640, 480)
out = (np.zeros((n, cap), KP_DTYPE), np.zeros((n, cap, 32), np.uint8), np.zeros(n, np.int32), np.zeros(n, np.int32))
for _ in range(2): ex.extract_batch(host, out=out)
t0 = time.perf_counter()
for _ in range(5): ex.extract_batch(host, out=out)
dt = (time.perf_counter() - t0) / 5
print("threads=%s staging=%s: %.2f ms per 1024 frames = %.0f frames/s" % (os.environ.get("RUMI_COPY_THREADS", "default"),
      "driver" if os.environ.get("RUMI_NO_HOST_STAGING") else "library", dt * 1e3, n / dt))
