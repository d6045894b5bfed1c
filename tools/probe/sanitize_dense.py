"""Small dense-frame workload for compute-sanitizer (memcheck / racecheck): single frames (1024-thread quad-tree CTAs, loop-form
bucket sort), then a chunk of 4 + 4 twice (global-memory sort first, second pass afterwards).  Checks batch == single."""
import sys
sys.path.insert(0, "/root/repo")
import numpy as np
from rumi_slam_b200 import ORBextractor
from rumi_slam_b200.synth import synthetic_frame
w, h = 640, 480
rng = np.random.default_rng(5)
base = synthetic_frame(3, w, h).astype(np.int32)
frames = np.stack([np.clip(base + rng.integers(-a, a + 1, base.shape), 0, 255).astype(np.uint8) for a in (0, 14, 30, 60, 8, 22, 0, 40)])
single = ORBextractor(1000, 1.2, 8, 20, 7)
ref = [single(f) for f in frames]
ex = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=8)
for rep in range(2):
    kps, desc, nkp, nmono = ex.extract_batch(frames)
    for i, (m, k, d) in enumerate(ref):
        assert nkp[i] == len(k) and np.array_equal(desc[i, :nkp[i]], d) and np.array_equal(kps[i, :nkp[i]], k), (rep, i)
print("ok", [int(n) for n in nkp])
