"""Does PCIe traffic slow the extraction kernels down?  Resident batches (rumi_orb_extract_batch_device) timed alone and
while another stream keeps copying unrelated pinned buffers H2D / D2H at the e2e path's volume (315 MB in, 63 MB out per
1024-frame step).  If the resident rate holds under copy load, the e2e gap is scheduling; if it drops, it is contention."""
import os, sys, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rumi_slam_b200 import ORBextractor
from rumi_slam_b200.synth import synthetic_batch
n = 1024
dev = torch.from_numpy(synthetic_batch(n, 640, 480, seed0=0, unique=16)).cuda()
ex = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=64)
od = None
for _ in range(3): od = ex.extract_batch_device(dev, out=od)
def run(steps=8):
    ex.timer_start()
    for _ in range(steps): ex.extract_batch_device(dev, out=od, sync=False)
    return n * steps * 1e3 / ex.timer_stop()
print("alone: %.0f frames/s" % run())
hin = torch.empty(315 << 20, dtype=torch.uint8).pin_memory(); din = torch.empty_like(hin, device="cuda")
hout = torch.empty(63 << 20, dtype=torch.uint8).pin_memory(); dout = torch.empty(63 << 20, dtype=torch.uint8, device="cuda")
for mode in ("h2d", "d2h", "both"):
    stop = False
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def pump():
        while not stop:
            if mode in ("h2d", "both"):
                with torch.cuda.stream(s1): din.copy_(hin, non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(s2): hout.copy_(dout, non_blocking=True)
            s1.synchronize(); s2.synchronize()
    t = threading.Thread(target=pump); t.start()
    import time; time.sleep(0.05)
    r = run(16)
    stop = True; t.join()
    print("with background %s copies: %.0f frames/s" % (mode, r))
