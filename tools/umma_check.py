"""Parity + rate of the tcgen05 top-2 kernel (RUMI_MATCH=umma) against the LOP3+POPC kernel on the same data."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rumi_slam_b200 import ORBmatcher
from rumi_slam_b200.synth import perturbed_descriptors

def matcher(mode):
    os.environ["RUMI_MATCH"] = mode
    m = ORBmatcher()
    os.environ.pop("RUMI_MATCH")
    return m
rng = np.random.default_rng(1)
pop, um, im = matcher("popc"), matcher("umma"), matcher("imma")
for nq, nt in [(128, 128), (100, 77), (300, 1000), (1000, 5000), (4096, 40000)]:
    T = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
    Q = perturbed_descriptors(T, nq, seed=nq, flip_p=0.08)
    a = pop.top2(Q, T); b = um.top2(Q, T)
    ok = all(np.array_equal(x, y) for x, y in zip(a, b))
    print(nq, nt, "umma == popc:", ok, um.last_path())
    if not ok:
        bad = np.flatnonzero((a[0] != b[0]) | (a[1] != b[1]) | (a[2] != b[2]))
        print("  mismatches", len(bad), "first", bad[:5], [x[bad[:5]] for x in a], [x[bad[:5]] for x in b])
nq = nt = 40000
T = torch.from_numpy(rng.integers(0, 256, (nt, 32), dtype=np.uint8)).cuda()
Q = torch.from_numpy(perturbed_descriptors(T.cpu().numpy(), nq, seed=3, flip_p=0.08)).cuda()
for name, m in (("popc", pop), ("imma", im), ("umma", um)):
    out = m.top2_device(Q, T)
    for _ in range(3):
        m.top2_device(Q, T, out=out)
    m.timer_start()
    for _ in range(10):
        m.top2_device(Q, T, out=out, sync=False)
    ms = m.timer_stop() / 10
    print("%s: %.3f ms  %.3g pairs/s" % (name, ms, nq * nt / ms * 1e3))
