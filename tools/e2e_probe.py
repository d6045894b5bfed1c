import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rumi_slam_b200 import ORBextractor, KP_DTYPE
from rumi_slam_b200.synth import synthetic_batch
n = 1024
host = synthetic_batch(n, 640, 480, seed0=0, unique=16)
pinned = torch.from_numpy(host).pin_memory(); hp = pinned.numpy(); dev = pinned.cuda()
for streams in [int(x) for x in os.environ.get('PROBE_STREAMS', '2,3,4').split(',')]:
  for chunk in [int(x) for x in os.environ.get('PROBE_CHUNKS', '32,64').split(',')]:
    os.environ["RUMI_STREAMS"] = str(streams)
    ex = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=chunk)
    cap = ex.frame_capacity(640, 480)
    keep = [torch.zeros((n, cap, 28), dtype=torch.uint8).pin_memory(), torch.zeros((n, cap, 32), dtype=torch.uint8).pin_memory(),
            torch.zeros(n, dtype=torch.int32).pin_memory(), torch.zeros(n, dtype=torch.int32).pin_memory()]
    out = (keep[0].numpy().view(KP_DTYPE).reshape(n, cap), keep[1].numpy(), keep[2].numpy(), keep[3].numpy())
    for _ in range(2): ex.extract_batch(hp, out=out)
    ex.timer_start()
    for _ in range(5): ex.extract_batch(hp, out=out)
    ms = ex.timer_stop()
    od = None
    for _ in range(2): od = ex.extract_batch_device(dev, out=od)
    ex.timer_start()
    for _ in range(5): ex.extract_batch_device(dev, out=od, sync=False)
    msr = ex.timer_stop()
    print("streams", streams, "chunk", chunk, "e2e %.2f ms/step (%.0f fps)  resident %.2f ms/step (%.0f fps)" % (ms / 5, n * 5e3 / ms, msr / 5, n * 5e3 / msr))
    ex.close()
