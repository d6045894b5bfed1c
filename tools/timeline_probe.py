"""Stage timeline of ONE host-path batch (RUMI_TIMELINE dump of the profile events): prints per chunk the H2D, kernel and
D2H spans so that pipeline bubbles are visible."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
path = "/tmp/rumi_timeline.txt"
os.environ["RUMI_TIMELINE"] = path
import numpy as np, torch
from rumi_slam_b200 import ORBextractor, KP_DTYPE
from rumi_slam_b200.synth import synthetic_batch
n = 1024
host = synthetic_batch(n, 640, 480, seed0=0, unique=16)
hp = torch.from_numpy(host).pin_memory().numpy()
ex = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=64)
cap = ex.frame_capacity(640, 480)
keep = [torch.zeros((n, cap, 28), dtype=torch.uint8).pin_memory(), torch.zeros((n, cap, 32), dtype=torch.uint8).pin_memory(),
        torch.zeros(n, dtype=torch.int32).pin_memory(), torch.zeros(n, dtype=torch.int32).pin_memory()]
out = (keep[0].numpy().view(KP_DTYPE).reshape(n, cap), keep[1].numpy(), keep[2].numpy(), keep[3].numpy())
for _ in range(3): ex.extract_batch(hp, out=out)
ex.profile(True)
ex.extract_batch(hp, out=out)
if os.path.exists(path): os.remove(path)
ex.extract_batch(hp, out=out)
ex.profile_read()
names = {0: "pyr", 1: "fast", 2: "oct", 3: "slot", 4: "blur", 5: "desc", 6: "H2D", 7: "D2H"}
rows = [l.split() for l in open(path) if not l.startswith("-1")]
rows = [(int(a), float(b), float(c)) for a, b, c in rows]
chunk, cur = [], {}
for st, a, b in rows:
    if st == 6 and cur: chunk.append(cur); cur = {}
    cur[st] = (a, b)
chunk.append(cur)
print("chunk  H2D[start-end]      kernels[start-end] (pyr fast oct slot blur desc durations)      D2H[start-end]")
for i, c in enumerate(chunk):
    k0 = c[0][0]; k1 = c[5][1]
    print("%2d  %7.3f-%7.3f   %7.3f-%7.3f  (%s)   %7.3f-%7.3f" % (i, c[6][0], c[6][1], k0, k1, " ".join("%.3f" % (c[s][1] - c[s][0]) for s in range(6)), c[7][0], c[7][1]))
