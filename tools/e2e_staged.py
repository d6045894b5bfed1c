"""A/B of the host-batch pipelines: staged (copy streams + 3 input buffers + 2 workspaces, default) vs RUMI_STAGED=0
(H2D / kernels / D2H of a chunk on its own workspace stream, 4 workspaces)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rumi_slam_b200 import ORBextractor, KP_DTYPE
from rumi_slam_b200.synth import synthetic_batch
n = 1024
host = synthetic_batch(n, 640, 480, seed0=0, unique=16)
pinned = torch.from_numpy(host).pin_memory(); hp = pinned.numpy()
ref = None
for staged in ("1", "0", "1", "0"):
    os.environ["RUMI_STAGED"] = staged
    ex = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=64)
    cap = ex.frame_capacity(640, 480)
    keep = [torch.zeros((n, cap, 28), dtype=torch.uint8).pin_memory(), torch.zeros((n, cap, 32), dtype=torch.uint8).pin_memory(),
            torch.zeros(n, dtype=torch.int32).pin_memory(), torch.zeros(n, dtype=torch.int32).pin_memory()]
    out = (keep[0].numpy().view(KP_DTYPE).reshape(n, cap), keep[1].numpy(), keep[2].numpy(), keep[3].numpy())
    for _ in range(3): ex.extract_batch(hp, out=out)
    t0 = time.perf_counter()
    for _ in range(8): ex.extract_batch(hp, out=out)
    dt = time.perf_counter() - t0
    sig = (out[2].copy(), out[1][:, :1000].copy())
    if ref is None: ref = sig
    same = np.array_equal(sig[0], ref[0]) and np.array_equal(sig[1], ref[1])
    print("staged", staged, "e2e %.3f ms/step  %.0f frames/s  identical %s" % (dt / 8 * 1e3, n * 8 / dt, same))
    ex.close()
