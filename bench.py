#!/usr/bin/env python3
"""bench.py -- ORB frames/s (640x480, 1000 kp) + Hamming matches/s on B200s, beside the reference CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One step = one pass of the hot path over one batch of synthetic frames (BASELINE.json configs[1]: 1024 synthetic
640x480 frames, TUM ORB settings 1000/1.2/8/20/7).  Frames shard by frame across ranks with no data-path
collective, each rank owning its own 1024-frame batch ("scaling": "weak").

  value       frames/s over all ranks, inputs already resident in HBM, device-timed on the launching streams
  e2e         same metric through ORBextractor.extract_batch (HOST pinned buffers; H2D + kernels + D2H timed);
              e2e.pageable = the same call on ordinary (pageable) host memory
  parity      OUTSIDE the timed regions: sampled frames of both result sets, every query of cfg 5a, sampled queries of
              cfg 5b and one stereo pair per shape are compared with the oracle; sharded == unsharded at N > 1
  roofline    dominant kernel: algorithmic bytes per launch / its CUDA-event duration vs the measured HBM peak
  strong      literal configs[1]: ONE 1024-frame batch split over the N ranks (strong scaling)
  matching    cfg 5a (40 front + 40 back keyframes x 1000 descriptors, all-pairs top-2): matches/s and pairs/s
  matching_5b cfg 5b: 10^6 x 10^6, train set sharded over the ranks, candidate all-gather (NCCL) inside the C ABI
  stereo_euroc / stereo_kitti   configs[2] / [3]: left + right extraction + Frame::ComputeStereoMatches, pairs/s
  submap_merge  descriptor-based key-point association of 40 + 40 key frames (describe + top-2 + acceptance)
  cpu_baseline  the oracle (oracle/_ref = the unmodified reference ORBextractor.cc over a cv stub, when built;
                else the C++ port) on the host cores, bounded sample

`--impl reference` times that CPU implementation alone on the same config / metric.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W_IMG, H_IMG, NFEAT = 640, 480, 1000
ORB = dict(nfeatures=NFEAT, scale=1.2, nlevels=8, ini=20, mn=7)
METRIC = "ORB frames/sec (640x480, 1000 kp)"
WORKLOAD = "ORB extraction, synthetic 640x480 frames (BASELINE configs[1]), 1000 features / 8 levels / 1.2 / FAST 20-7"
NCU_PROFILE = os.path.join(ROOT, "profiles", "r2_ncu_kernels.json")


def level_sizes(w, h, nlevels=8, scale=1.2):
    s = np.float32(1.0)
    out = []
    for _ in range(nlevels):
        inv = np.float32(1.0) / s
        out.append((int(np.rint(np.float32(w) * inv)), int(np.rint(np.float32(h) * inv))))
        s = np.float32(s * np.float32(scale))
    return out


def algorithmic_bytes(w, h, nkp):
    """Per-frame ALGORITHMIC bytes of each stage (DESIGN.md 'bytes per unit')."""
    px = [a * b for a, b in level_sizes(w, h)]
    return {
        "pyramid": sum(px[:-1]) + sum(px[1:]),          # read level l-1, write level l, l = 1..7
        "fast": sum(px),                                 # read every level once
        "octree": 0,                                     # candidate lists only (KBs): not an HBM-bound stage
        "slots": 0,
        "blur": 2 * sum(px),                             # read + write every level
        "describe": nkp * (749 + 512 + 60),              # circular patch + 512 samples + 60 B out per keypoint
        "frame": sum(px) + sum(px[1:]) + 60 * nkp,       # whole path: 1 653 864 B for 640x480 / 1000 kp
    }


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while a timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, windows):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for t, line in self.lines:
            if not any(a <= t <= b for a, b in windows):
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation on the host cores
# --------------------------------------------------------------------------------------------------------------------
class CpuExtractor:
    """The CPU reference path on a PERSISTENT pool of host threads (ctypes releases the GIL): oracle/_ref = the unmodified
    reference ORBextractor.cc compiled over oracle/cvstub when it was built, else the C++ port."""

    def __init__(self, threads, orb=None):
        from concurrent.futures import ThreadPoolExecutor
        from oracle import orb_oracle, ref_lib
        self.orb = dict(ORB if orb is None else orb)
        self.kind = "reference" if ref_lib.available() else "port"
        self.fn = ref_lib.extract if self.kind == "reference" else orb_oracle.extract
        orb_oracle.build()
        self.threads = threads
        self.pool = ThreadPoolExecutor(max_workers=threads)

    def run(self, frames):
        """Extracts every frame once; returns seconds."""
        t0 = time.perf_counter()
        list(self.pool.map(lambda im: self.fn(im, **self.orb), frames, chunksize=1))
        return time.perf_counter() - t0

    def close(self):
        self.pool.shutdown()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from rumi_slam_b200.synth import synthetic_batch
    cores = len(os.sched_getaffinity(0))
    per_step = 40 * cores                      # >= 40 frames per thread per step: the pool runs at its steady rate
    frames = synthetic_batch(per_step, W_IMG, H_IMG, seed0=1000, unique=min(per_step, 32))
    cpu = CpuExtractor(cores)
    cpu.run(frames[:2 * cores])                # load + first touch
    for _ in range(args.warmup):
        cpu.run(frames[:4 * cores])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu.run(frames)
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_step": per_step,
                       "sample": "bounded sample of the 1024-frame batch: %d frames per step" % per_step},
            "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": cpu.kind,
                             "sample": "%d frames per step (40 per thread) on a persistent pool of %d host threads; "
                                       "oracle/_ref = unmodified reference ORBextractor.cc compiled over oracle/cvstub "
                                       "(cv2-pinned primitives)" % (per_step, cores)},
            "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    cpu.close()
    print(json.dumps(line))


def match_roofline(path, pairs_per_s, world, sm_mhz):
    """The top-2 kernels are bound by an execution pipe, not by HBM (operands are reused from shared memory):
    popc: 5 POPC per pair at the measured 16 POPC/clk/SM (tools/probe/pipe_probe.cu)."""
    if path == "umma":
        # peak: the nominal dense int8 rate, 4.5 POP/s per GPU.  ncu's own counter agrees with it: a capture at 2.69e12
        # pairs/s (1377 TOP/s) read sm__ops_path_tensor_op_utcimma_src_int8 = 30.4 % of peak
        # (profiles/r1_ncu_umma_v3_excerpt.txt).  Twice the measured bf16 rate of MEASURED_PEAKS.json is given beside it.
        peak_tops = 4500.0
        extra = None
        try:
            extra = 2 * float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"])
        except Exception:
            pass
        peak = world * peak_tops * 1e12 / 512.0           # 256 MACs = 512 ops per pair
        return {"bound": "tensor", "kernel": "tcgen05.mma kind::i8 (UTCIMMA), TMEM accumulators", "ops_per_pair": 512,
                "achieved": round(pairs_per_s * 512 / 1e12, 1), "peak": world * peak_tops, "unit": "TOP/s",
                "peak_pairs_per_s": peak, "frac": pairs_per_s / peak,
                "peak_source": "nominal dense int8 4.5 POP/s per GPU (consistent with ncu's utcimma pct_of_peak)",
                "frac_of_2x_measured_bf16": (pairs_per_s * 512 / 1e12) / (world * extra) if extra else None,
                "note": "query operand in TMEM, train tiles from a shared-memory ring; the pop(t) term is a ninth K step "
                        "(288 / 256 of the MMA work is algorithmic); frac_of_2x_measured_bf16 compares with twice the "
                        "burst cuBLAS bf16 rate measured on this pool (profiles/r2_summary.md)"}
    peak = world * 148 * 16 * sm_mhz * 1e6 / 5
    return {"bound": "popc-pipe", "popc_per_pair": 5, "popc_per_clk_per_sm": 16, "peak_pairs_per_s": peak,
            "frac": pairs_per_s / peak,
            "note": "POPC rate measured 15.8/clk/SM (tools/probe/pipe_probe.cu); 5 POPC + 14 LOP3 per pair"}


def bind_to_gpu_numa_node(torch, local):
    """With several ranks on one host the pinned staging buffers should live on the NUMA node whose PCIe root the GPU
    hangs off, and the rank should run there: sysfs gives the node's CPU list for the GPU's PCI address."""
    try:
        p = torch.cuda.get_device_properties(local)
        addr = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        base = "/sys/bus/pci/devices/" + addr
        node = int(open(base + "/numa_node").read())
        cpus = set()
        for part in open(base + "/local_cpulist").read().strip().split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return {"pci": addr, "numa_node": node, "cpus": len(cpus)}
    except Exception as e:                                   # no sysfs / no permission: run unbound
        return {"error": str(e)[:80]}


def host_topology():
    """NUMA / PCIe layout of the box (context for the e2e scaling curve: every rank's H2D shares the host links)."""
    out = {}
    try:
        r = subprocess.run(["numactl", "-H"], capture_output=True, text=True, timeout=5)
        out["numactl_H"] = [l for l in r.stdout.splitlines() if l.startswith(("available", "node")) and "size" not in l][:6]
    except Exception:
        try:
            nodes = sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node"))
            out["numa_nodes"] = {n: open("/sys/devices/system/node/%s/cpulist" % n).read().strip() for n in nodes}
        except Exception:
            pass
    try:
        r = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=10)
        out["nvidia_smi_topo"] = [l for l in r.stdout.splitlines() if l.startswith("GPU")][:8]
    except Exception:
        pass
    return out


def kp_bytes_equal(dev_rows, ref):
    """dev_rows: [cap,7] float32 view of 28-byte records; ref: KP_DTYPE array."""
    raw = np.ascontiguousarray(dev_rows).view(np.uint8).reshape(-1, 28)[:len(ref)]
    return np.array_equal(raw.reshape(-1).view(ref.dtype), ref)


def run_b200(args):
    import torch
    import torch.distributed as dist
    from rumi_slam_b200 import KP_DTYPE, ORBextractor, ORBmatcher
    from rumi_slam_b200.sharding import frame_shard, init_matcher_comm, query_shard, sharded_top2, train_shard
    from rumi_slam_b200.synth import stereo_pair, synthetic_batch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    nframes, chunk = args.frames, args.chunk
    cores = len(os.sched_getaffinity(0))
    check_oracle = not args.no_check
    if check_oracle:
        from oracle import orb_oracle
        orb_oracle.build()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_ranks_true(ok):
        if world == 1:
            return bool(ok)
        t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    # ---- host placement: run (and first-touch the pinned buffers) on the CPUs next to this rank's GPU ----
    numa = bind_to_gpu_numa_node(torch, local) if world > 1 else None
    threads_here = max(1, len(os.sched_getaffinity(0)) // world)       # host threads this rank may use for the checks

    # ---- synthetic inputs: 1024 frames per rank, distinct seeds per rank ----
    host = synthetic_batch(nframes, W_IMG, H_IMG, seed0=10000 * rank, unique=args.unique)
    pinned = torch.from_numpy(host).pin_memory()
    resident = pinned.to(dev)                                   # 315 MB > 126 MB of L2: no flush needed
    ex = ORBextractor(NFEAT, 1.2, 8, 20, 7, device=local, max_batch=chunk)
    cap = ex.frame_capacity(W_IMG, H_IMG)
    out_dev = (torch.zeros((nframes, cap, 7), dtype=torch.float32, device=dev),
               torch.zeros((nframes, cap, 32), dtype=torch.uint8, device=dev),
               torch.zeros(nframes, dtype=torch.int32, device=dev), torch.zeros(nframes, dtype=torch.int32, device=dev))
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    windows = []

    # ---- value: inputs resident in HBM ----
    for _ in range(args.warmup):
        ex.extract_batch_device(resident, out=out_dev, sync=True)
    if rank == 0:
        sampler.start()
    ex.launch_count(reset=True)
    barrier()
    t_a = time.perf_counter()
    ex.timer_start()
    for _ in range(args.steps):
        ex.extract_batch_device(resident, out=out_dev, sync=False)
    ms = ex.timer_stop()
    barrier()
    windows.append((t_a, time.perf_counter()))
    ms = max_over_ranks(ms)
    launches = ex.launch_count(reset=True)
    # per-stage durations: a second pass of the SAME work on ONE stream, so that the CUDA-event span of a stage is
    # exclusive (in the timed region above chunks alternate between streams and stages of neighbouring chunks overlap)
    psteps = max(2, args.steps // 2)
    ex.set_streams(1)
    ex.extract_batch_device(resident, out=out_dev, sync=True)
    ex.profile(True)
    ex.profile_read(reset=True)
    t_a = time.perf_counter()
    ex.timer_start()
    for _ in range(psteps):
        ex.extract_batch_device(resident, out=out_dev, sync=False)
    ms_serial = ex.timer_stop()
    windows.append((t_a, time.perf_counter()))
    stages = ex.profile_read(reset=True)
    ex.profile(False)
    ex.set_streams(0)
    # leave the results of the PIPELINED configuration (what `value` timed) in out_dev for the parity check
    for t in out_dev:
        t.zero_()
    ex.extract_batch_device(resident, out=out_dev, sync=True)
    ex.launch_count(reset=True)
    value = world * nframes * args.steps / (ms / 1e3)
    nkp_mean = float(out_dev[2].float().mean().item())

    # ---- e2e: host pinned buffers through the public batch call (H2D + kernels + D2H inside the timed region) ----
    out_host = (torch.zeros((nframes, cap, 28), dtype=torch.uint8).pin_memory().numpy().view(KP_DTYPE).reshape(nframes, cap),
                torch.zeros((nframes, cap, 32), dtype=torch.uint8).pin_memory().numpy(),
                torch.zeros(nframes, dtype=torch.int32).pin_memory().numpy(),
                torch.zeros(nframes, dtype=torch.int32).pin_memory().numpy())
    host_pinned = pinned.numpy()
    for _ in range(max(1, args.warmup - 1)):
        ex.extract_batch(host_pinned, out=out_host)
    barrier()
    t_a = time.perf_counter()
    ex.timer_start()
    for _ in range(args.steps):
        ex.extract_batch(host_pinned, out=out_host)
    ms_e2e = ex.timer_stop()
    barrier()
    windows.append((t_a, time.perf_counter()))
    ms_e2e = max_over_ranks(ms_e2e)
    launches_e2e = ex.launch_count(reset=True)
    e2e_value = world * nframes * args.steps / (ms_e2e / 1e3)
    d2h = int(ex.last_d2h_bytes()) if hasattr(ex, "last_d2h_bytes") else nframes * cap * (28 + 32) + 8 * nframes
    # the same call on ordinary pageable host memory (what a caller holding cv::Mat images has): the library stages the
    # chunks through its own pinned blocks with several host threads (the driver's single-threaded staging: 25 k frames/s)
    out_page = (np.zeros((nframes, cap), KP_DTYPE), np.zeros((nframes, cap, 32), np.uint8), np.zeros(nframes, np.int32),
                np.zeros(nframes, np.int32))
    ex.extract_batch(host, out=out_page)
    barrier()
    pg_steps = max(2, args.steps // 4)
    t_a = time.perf_counter()
    for _ in range(pg_steps):
        ex.extract_batch(host, out=out_page)
    torch.cuda.synchronize()
    dt_page = max_over_ranks(time.perf_counter() - t_a)
    barrier()
    windows.append((t_a, time.perf_counter()))
    e2e_pageable = world * nframes * pg_steps / dt_page
    ex.launch_count(reset=True)
    # raw host->device bandwidth of the same pinned buffer, all ranks copying at once (context for e2e: 307 200 B per
    # frame must cross the host links; this is the ceiling of the e2e metric on this box)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    scratch = torch.empty_like(resident)
    scratch.copy_(pinned, non_blocking=True)
    barrier()
    ev0.record()
    for _ in range(3):
        scratch.copy_(pinned, non_blocking=True)
    ev1.record()
    torch.cuda.synchronize()
    h2d_gbps = 3 * pinned.numel() / (max_over_ranks(ev0.elapsed_time(ev1)) / 1e3) / 1e9
    del scratch
    h2d = nframes * W_IMG * H_IMG
    ceiling = world * h2d_gbps * 1e9 / (W_IMG * H_IMG)

    # ---- parity of what was just timed (outside every timed region): frames spread over all chunks / workspaces ----
    parity = {"checked": False}
    if check_oracle:
        nsample = min(nframes, 128 if world == 1 else 64)
        pick = np.unique(np.linspace(0, nframes - 1, nsample).astype(int))
        ref = orb_oracle.extract_mt([host[i] for i in pick], threads=threads_here)
        dk = out_dev[0][torch.from_numpy(pick).to(dev)].cpu().numpy()
        dd = out_dev[1][torch.from_numpy(pick).to(dev)].cpu().numpy()
        dn, dm = out_dev[2].cpu().numpy(), out_dev[3].cpu().numpy()
        bad = []
        for j, i in enumerate(pick):
            rk, rd, rm = ref[j]
            ok_dev = dn[i] == len(rk) and dm[i] == rm and kp_bytes_equal(dk[j], rk) and np.array_equal(dd[j, :len(rk)], rd)
            ok_host = (out_host[2][i] == len(rk) and out_host[3][i] == rm and np.array_equal(out_host[0][i, :len(rk)], rk)
                       and np.array_equal(out_host[1][i, :len(rk)], rd))
            ok_page = (out_page[2][i] == len(rk) and np.array_equal(out_page[0][i, :len(rk)], rk)
                       and np.array_equal(out_page[1][i, :len(rk)], rd))
            if not (ok_dev and ok_host and ok_page):
                bad.append((int(i), bool(ok_dev), bool(ok_host), bool(ok_page)))
        parity = {"checked": True, "frames": int(len(pick)) * world, "chunks_covered": int(len(set(pick // chunk))),
                  "device_path_ok": all_ranks_true(not any(not b[1] for b in bad)),
                  "host_path_ok": all_ranks_true(not any(not b[2] for b in bad)),
                  "pageable_path_ok": all_ranks_true(not any(not b[3] for b in bad)),
                  "first_bad": bad[:3],
                  "against": "oracle/orb_oracle.cpp (pinned to the unmodified reference ORBextractor.cc), keypoints incl. "
                             "order, angles and descriptors bit for bit"}
    del out_page

    # ---- strong scaling, the literal configs[1]: ONE batch of 1024 frames split over the ranks ----
    strong = None
    if world > 1:
        b, e = frame_shard(nframes, rank, world)
        sub_dev, sub_host = resident[b:e], host_pinned[b:e]
        sub_out = tuple(t[b:e] for t in out_dev)
        sub_out_host = tuple(t[b:e] for t in out_host)
        for _ in range(3):
            ex.extract_batch_device(sub_dev, out=sub_out, sync=True)
        barrier()
        t_a = time.perf_counter()
        ex.timer_start()
        for _ in range(args.steps):
            ex.extract_batch_device(sub_dev, out=sub_out, sync=False)
        ms_s = max_over_ranks(ex.timer_stop())
        barrier()
        ex.extract_batch(sub_host, out=sub_out_host)
        barrier()
        ex.timer_start()
        for _ in range(args.steps):
            ex.extract_batch(sub_host, out=sub_out_host)
        ms_se = max_over_ranks(ex.timer_stop())
        barrier()
        windows.append((t_a, time.perf_counter()))
        strong = {"workload": "ONE batch of %d frames frame-sharded over %d GPUs (%d per GPU)" % (nframes, world, e - b),
                  "scaling": "strong", "value": nframes * args.steps / (ms_s / 1e3), "e2e": nframes * args.steps / (ms_se / 1e3),
                  "unit": "frames/s", "ms_per_step": ms_s / args.steps, "e2e_ms_per_step": ms_se / args.steps}
        ex.launch_count(reset=True)

    # ---- single-frame latency of operator() (BASELINE configs[0]): host image in, keypoints + descriptors out ----
    latency = None
    if rank == 0:
        ex1 = ORBextractor(NFEAT, 1.2, 8, 20, 7, device=local)
        for _ in range(20):
            ex1(host[0])
        ts = []
        for i in range(100):
            t0 = time.perf_counter()
            ex1(host[i % args.unique])
            ts.append(time.perf_counter() - t0)
        ts = np.array(ts) * 1e3
        latency = {"workload": "ORBextractor::operator() on ONE 640x480 host frame, synchronous (configs[0])",
                   "median_ms": float(np.median(ts)), "p90_ms": float(np.percentile(ts, 90)), "calls": len(ts)}
        ex1.close()

    # ---- textured frames (context, rank 0 only): the same frames + uniform noise of +-20 grey levels put ~4x the FAST
    # candidates on every level (real camera images sit between the two), which moves the quad-tree onto its dense paths ----
    textured = None
    if rank == 0 and not args.no_textured:
        try:
            nd = 256
            rngd = np.random.default_rng(20)
            noisy = np.clip(host[:nd].astype(np.int16) + rngd.integers(-20, 21, host[:nd].shape, dtype=np.int16),
                            0, 255).astype(np.uint8)
            exd = ORBextractor(NFEAT, 1.2, 8, 20, 7, device=local, max_batch=chunk)
            ddev = torch.from_numpy(noisy).to(dev)
            od = None
            for _ in range(3):
                od = exd.extract_batch_device(ddev, out=od)
            exd.launch_count(reset=True)
            exd.timer_start()
            for _ in range(5):
                exd.extract_batch_device(ddev, out=od, sync=False)
            ms_d = exd.timer_stop()
            nl_d = exd.launch_count(reset=True)
            ex1 = ORBextractor(NFEAT, 1.2, 8, 20, 7, device=local)
            for _ in range(10):
                ex1(noisy[0])
            ts = []
            for i in range(50):
                t0 = time.perf_counter()
                ex1(noisy[i % nd])
                ts.append(time.perf_counter() - t0)
            cand0 = len(ex1.debug_candidates(0))
            ex1.close()
            textured = {"workload": "%d of the same frames + uniform noise of +-20 grey levels, resident in HBM" % nd,
                        "frames_per_s": nd * 5 / (ms_d / 1e3), "gpu_launches": nl_d,
                        "level0_fast_candidates_last_frame": cand0,
                        "single_frame_median_ms": float(np.median(ts) * 1e3)}
            if check_oracle:
                pick = [0, 85, 170, 255]
                refd = orb_oracle.extract_mt([noisy[i] for i in pick], threads=threads_here)
                dk = od[0][pick].cpu().numpy()
                dd = od[1][pick].cpu().numpy()
                dn = od[2].cpu().numpy()
                textured["parity_ok"] = bool(all(dn[i] == len(rk) and kp_bytes_equal(dk[j], rk) and
                                                 np.array_equal(dd[j, :len(rk)], rd)
                                                 for j, (i, (rk, rd, rm)) in enumerate(zip(pick, refd))))
            exd.close()
        except Exception as e:                               # context row only: never fails the bench
            textured = {"error": str(e)[:200]}

    # ---- matching: cfg 5a, 40 front + 40 back keyframes x 1000 descriptors, all pairs; train sharded over ranks ----
    nq = nt = 40 * 1000
    desc = out_dev[1][:80]                                      # 80 extracted frames of this rank
    cnt = out_dev[2][:80].tolist()
    rows = torch.cat([desc[i, :min(c, 1000)] for i, c in enumerate(cnt)])
    reps = -(-(nq + nt) // rows.shape[0])
    rows = rows.repeat(reps, 1)
    Q, T = rows[:nq].contiguous(), rows[nq:nq + nt].flip(0).contiguous()
    if world > 1:                                              # every rank must see the same Q / T
        dist.broadcast(Q, 0); dist.broadcast(T, 0)
    m = ORBmatcher(device=local)
    if world > 1:
        init_matcher_comm(m)                                   # the matcher's own NCCL communicator (C ABI)
    b, e = train_shard(nt, rank, world)
    T_local = T[b:e].contiguous()
    res5a = (torch.empty(nq, dtype=torch.int32, device=dev), torch.empty(nq, dtype=torch.int16, device=dev),
             torch.empty(nq, dtype=torch.int16, device=dev))
    torch.cuda.synchronize()

    def match_once():
        if world == 1:
            return m.top2_device(Q, T_local, t_base=b, out=res5a, sync=False)
        return sharded_top2(m, Q, T_local, b, out=res5a)

    for _ in range(3):
        match_once()
    barrier()
    m.launch_count(reset=True)
    msteps = max(args.steps, 10)
    t_a = time.perf_counter()
    m.timer_start()
    for _ in range(msteps):
        match_once()
    ms_match = m.timer_stop()
    barrier()
    windows.append((t_a, time.perf_counter()))
    ms_match = max_over_ranks(ms_match)
    match_launches = m.launch_count(reset=True)
    match_path = m.last_path()
    matches_per_s = nq * msteps / (ms_match / 1e3)
    pairs_per_s = matches_per_s * nt
    matching = {"workload": "cfg 5a: 40000 x 40000 descriptors all-pairs top-2, train set sharded over %d GPU(s)%s"
                            % (world, "" if world == 1 else ", NCCL candidate all-gather + fold inside rumi_hamming_top2_sharded"),
                "matches_per_s": matches_per_s, "pairs_per_s": pairs_per_s, "ms_per_step": ms_match / msteps,
                "gpu_launches": match_launches, "steps": msteps, "kernel": match_path}
    # small problems can shard the QUERIES instead (no exchange, the result stays sharded by query)
    if world > 1:
        qb, qe = query_shard(nq, rank, world)
        Qs = Q[qb:qe].contiguous()
        outq = tuple(t[qb:qe] for t in (torch.empty_like(res5a[0]), torch.empty_like(res5a[1]), torch.empty_like(res5a[2])))
        for _ in range(3):
            m.top2_device(Qs, T, out=outq, sync=False)
        barrier()
        m.timer_start()
        for _ in range(msteps):
            m.top2_device(Qs, T, out=outq, sync=False)
        ms_q = max_over_ranks(m.timer_stop())
        barrier()
        matching["query_sharded"] = {"pairs_per_s": float(nq) * nt * msteps / (ms_q / 1e3), "ms_per_step": ms_q / msteps,
                                     "note": "each rank scans its own %d queries against all train rows; no exchange, "
                                             "result sharded by query" % (qe - qb)}
        m.launch_count(reset=True)
    # parity + CPU baseline of the matching metric (rank 0): the reference's bit-hack DescriptorDistance scan
    # (ORBmatcher.cc:1830-1844 inside :253-261) on 1 core (bounded sample) and on all cores (ALL queries = the parity check)
    if check_oracle:
        ok5a, cpu5a = True, None
        if rank == 0:
            q_np, t_np = Q.cpu().numpy(), T.cpu().numpy()
            t0 = time.perf_counter()
            orb_oracle.hamming_top2(q_np[:1500], t_np)
            dt1 = time.perf_counter() - t0
            t0 = time.perf_counter()
            ri, rd1, rd2 = orb_oracle.hamming_top2_mt(q_np, t_np, threads=cores)
            dta = time.perf_counter() - t0
            gi, g1, g2 = (x.cpu().numpy() for x in res5a)
            ok5a = bool(np.array_equal(gi, ri) and np.array_equal(g1.astype(np.uint16), rd1) and np.array_equal(g2.astype(np.uint16), rd2))
            if world > 1:                                      # sharded == unsharded on this GPU
                ui, u1, u2 = m.top2_device(Q, T, sync=True)
                ok5a = ok5a and bool(torch.equal(ui, res5a[0]) and torch.equal(u1, res5a[1]) and torch.equal(u2, res5a[2]))
            cpu5a = {"value": float(nq) * nt / dta, "unit": "pairs/s", "cores": cores, "kind": "port",
                     "matches_per_s": nq / dta, "single_thread_pairs_per_s": 1500.0 * nt / dt1,
                     "sample": "all 40000 x 40000 pairs on %d host threads (%.1f s); 1 thread: 1500 x 40000 pairs (%.1f s); "
                               "oracle/orb_oracle.cpp orc_hamming_top2 = the reference's SWAR DescriptorDistance in the "
                               "best/second-best scan, -O2" % (cores, dta, dt1)}
        matching["parity_ok"] = all_ranks_true(ok5a)
        matching["parity"] = "all 40000 queries == oracle scan%s" % ("" if world == 1 else "; sharded == unsharded on rank 0")
        matching["cpu_baseline"] = cpu5a

    # ---- cfg 5b: 10^6 x 10^6 all-pairs top-2 (descriptors = extracted ones tiled with ~10 % of the bits flipped, so
    #      true neighbours exist), train set sharded over the ranks, candidate all-gather inside the C ABI ----
    match_big = None
    if args.match_big > 0:
        nbig = args.match_big
        gen = torch.Generator(device=dev)
        base = rows[:min(rows.shape[0], 80000)]

        def tiled(n, seed):
            gen.manual_seed(seed)
            out = base.repeat(-(-n // base.shape[0]), 1)[:n].clone()
            for c0 in range(0, n, 1 << 18):                        # flip ~10 % of the bits, chunked to bound memory
                c1 = min(c0 + (1 << 18), n)
                mask = torch.zeros((c1 - c0, 32), dtype=torch.uint8, device=dev)
                for bit in range(8):
                    mask |= (torch.rand((c1 - c0, 32), device=dev, generator=gen) < 0.1).to(torch.uint8) << bit
                out[c0:c1] ^= mask
            return out.contiguous()
        Qb, Tb = tiled(nbig, 7), tiled(nbig, 8)
        if world > 1:
            dist.broadcast(Qb, 0); dist.broadcast(Tb, 0)
        bb, eb = train_shard(nbig, rank, world)
        Tb_local = Tb[bb:eb].contiguous()
        resb = (torch.empty(nbig, dtype=torch.int32, device=dev), torch.empty(nbig, dtype=torch.int16, device=dev),
                torch.empty(nbig, dtype=torch.int16, device=dev))
        torch.cuda.synchronize()

        def big_once():
            if world == 1:
                return m.top2_device(Qb, Tb_local, t_base=bb, out=resb, sync=False)
            return sharded_top2(m, Qb, Tb_local, bb, out=resb)
        big_once()
        barrier()
        bsteps = 3
        t_a = time.perf_counter()
        m.timer_start()
        for _ in range(bsteps):
            big_once()
        ms_big = m.timer_stop()
        barrier()
        windows.append((t_a, time.perf_counter()))
        ms_big = max_over_ranks(ms_big)
        big_pairs = float(nbig) * nbig * bsteps / (ms_big / 1e3)
        match_big = {"workload": "cfg 5b: %d x %d descriptors all-pairs top-2, train set sharded over %d GPU(s)" % (nbig, nbig, world),
                     "pairs_per_s": big_pairs, "matches_per_s": nbig * bsteps / (ms_big / 1e3),
                     "ms_per_step": ms_big / bsteps, "steps": bsteps, "kernel": m.last_path(),
                     "roofline": match_roofline(m.last_path(), big_pairs, world, 1965.0)}
        if check_oracle:
            okb = True
            if rank == 0:
                pick = np.random.default_rng(5).choice(nbig, 2000, replace=False)
                pk = torch.from_numpy(pick).to(dev)
                ri, rd1, rd2 = orb_oracle.hamming_top2_mt(Qb[pk].cpu().numpy(), Tb.cpu().numpy(), threads=cores)
                okb = bool(np.array_equal(resb[0][pk].cpu().numpy(), ri)
                           and np.array_equal(resb[1][pk].cpu().numpy().astype(np.uint16), rd1)
                           and np.array_equal(resb[2][pk].cpu().numpy().astype(np.uint16), rd2))
                if world > 1:
                    ui, u1, u2 = m.top2_device(Qb, Tb, sync=True)
                    okb = okb and bool(torch.equal(ui, resb[0]) and torch.equal(u1, resb[1]) and torch.equal(u2, resb[2]))
            match_big["parity_ok"] = all_ranks_true(okb)
            match_big["parity"] = "2000 sampled queries == brute-force oracle scan of all %d train rows%s" % (
                nbig, "" if world == 1 else "; sharded == unsharded (all queries) on rank 0")
        del Qb, Tb, Tb_local, resb
        m.launch_count(reset=True)

    # ---- configs[2] / [3]: stereo frames.  Per pair, as Frame::Frame does (Frame.cc:116-119, 828-985): the left and the
    #      right extractor run on two handles with both frames in flight, then ComputeStereoMatches on the device-resident
    #      pyramids.  Host images in, mvKeys / descriptors / mvuRight / mvDepth out: an end-to-end figure. ----
    def stereo_row(name, w, h, nf, fx, bf, npairs, unique):
        pairs = [stereo_pair(100 * rank + s, w, h) for s in range(unique)]
        exl, exr = ORBextractor(nf, 1.2, 8, 20, 7, device=local), ORBextractor(nf, 1.2, 8, 20, 7, device=local)
        ms_ = ORBmatcher(device=local)

        def one(i):
            # both frames are put in flight from this thread (rumi_orb_extract_begin x 2), then collected: the GPU-side
            # equivalent of the two std::threads of Frame::Frame
            left, right = pairs[i % unique]
            exl.begin(left); exr.begin(right)
            (_, lk, ld), (_, rk, rd) = exl.end(), exr.end()
            u, d, n = ms_.stereo_match(exl, exr, lk, ld, rk, rd, bf, bf / fx)
            return lk, ld, rk, rd, u, d, n
        for i in range(8):
            one(i)
        l0 = exl.launch_count(reset=True) + exr.launch_count(reset=True) + ms_.launch_count(reset=True)
        barrier()
        t0 = time.perf_counter()
        nm = 0
        for i in range(npairs):
            nm += one(i)[6]
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        barrier()
        windows.append((t0, time.perf_counter()))
        row = {"workload": "%s: %d synthetic %dx%d stereo pairs per GPU, %d features per image, left + right extraction "
                           "(two handles, both frames in flight from one host thread) + Frame::ComputeStereoMatches; host images in, results out"
                           % (name, npairs, w, h, nf),
               "pairs_per_s": world * npairs / dt, "frames_per_s": 2 * world * npairs / dt, "ms_per_pair": 1e3 * dt / npairs,
               "mean_stereo_matches": nm / float(npairs),
               "gpu_launches": exl.launch_count() + exr.launch_count() + ms_.launch_count(),
               "timing": "host wall clock around synchronous calls (H2D, kernels, D2H inside), max over ranks"}
        if check_oracle:
            lk, ld, rk, rd, u, d, n = one(1)
            left, right = pairs[1 % unique]
            ok_, od_, om_ = orb_oracle.extract(left, nfeatures=nf)
            ru, rdp, rn = orb_oracle.stereo_match(left, right, lk, ld, rk, rd, bf, bf / fx)
            row["parity_ok"] = all_ranks_true(bool(np.array_equal(lk, ok_) and np.array_equal(ld, od_) and n == rn
                                                   and np.array_equal(u, ru) and np.array_equal(d, rdp)))
            if rank == 0 and world == 1 and not args.no_cpu:
                cpu = CpuExtractor(cores, dict(nfeatures=nf, scale=1.2, nlevels=8, ini=20, mn=7))
                nsamp = max(2 * cores, 8)

                def cpu_pair(i):
                    left, right = pairs[i % unique]
                    a, bq = cpu.fn(left, **cpu.orb), cpu.fn(right, **cpu.orb)
                    orb_oracle.stereo_match(left, right, a[0], a[1], bq[0], bq[1], bf, bf / fx)
                cpu_pair(0)
                t0 = time.perf_counter()
                list(cpu.pool.map(cpu_pair, range(nsamp)))
                dtc = time.perf_counter() - t0
                cpu.close()
                row["cpu_baseline"] = {"value": nsamp / dtc, "unit": "pairs/s", "cores": cores,
                                       "kind": cpu.kind + " extraction + port of ComputeStereoMatches",
                                       "sample": "%d of the same pairs on %d host threads (%.1f s)" % (nsamp, cores, dtc)}
        exl.close(); exr.close(); ms_.close()
        return row

    stereo = {}
    if not args.no_stereo:
        stereo["stereo_euroc"] = stereo_row("cfg 3 (EuRoC shape)", 752, 480, 1200, 435.2, 47.9, args.stereo_pairs, 8)
        stereo["stereo_kitti"] = stereo_row("cfg 4 (KITTI shape)", 1241, 376, 2000, 718.856, 386.1448, args.stereo_pairs, 8)

    # ---- submap merge: descriptor-based key-point association of 40 matched key-frame pairs.  The
    #      cloud key frames carry no descriptors in the reference (cloud_edge_main.cpp:937), which therefore associates by
    #      pixel distance (CloudMerging.cc:503-551): that loop is the CPU baseline printed beside it. ----
    submap = None
    if not args.no_submap:
        npairs = 40
        kf_a, kf_b = host[:npairs], np.ascontiguousarray(np.roll(host[:npairs], (1, -1), axis=(1, 2)))
        exs = ORBextractor(NFEAT, 1.2, 8, 20, 7, device=local, max_batch=npairs)
        ka, _, na, _ = exs.extract_batch(kf_a)
        kb, _, nb, _ = exs.extract_batch(kf_b)
        keys1 = [ka[i, :na[i]].copy() for i in range(npairs)]
        keys2 = [kb[i, :nb[i]].copy() for i in range(npairs)]
        rng = np.random.default_rng(11 + rank)
        valid1 = [rng.random(len(k)) < 0.9 for k in keys1]
        valid2 = [rng.random(len(k)) < 0.9 for k in keys2]
        ma = ORBmatcher(0.75, device=local)
        ma.AssociateSubmap(exs, kf_a, keys1, valid1, kf_b, keys2, valid2)
        l0 = exs.launch_count(reset=True) + ma.launch_count(reset=True)
        barrier()
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            m12, counts = ma.AssociateSubmap(exs, kf_a, keys1, valid1, kf_b, keys2, valid2)
        dt = max_over_ranks(time.perf_counter() - t0)
        barrier()
        windows.append((t0, time.perf_counter()))
        submap = {"workload": "submap merge: %d front + %d back key frames per GPU (640x480, ~1000 key points each, 90 %% with a "
                              "map point): CloudFrameComputeDescriptors of all %d key frames + all-pairs top-2 of the %d matched "
                              "pairs + acceptance; host images / key points in, associations out" % (npairs, npairs, 2 * npairs, npairs),
                  "keyframe_pairs_per_s": world * npairs * reps / dt, "ms_per_merge": 1e3 * dt / reps,
                  "descriptor_pairs_per_s": world * reps * float(sum(int(a.sum()) * int(b.sum()) for a, b in zip(valid1, valid2))) / dt,
                  "associations": int(sum(counts)), "gpu_launches": exs.launch_count() + ma.launch_count(),
                  "timing": "host wall clock around synchronous calls (H2D, kernels, D2H inside), max over ranks"}
        if check_oracle:
            p = 3
            d1 = orb_oracle.describe(kf_a[p], keys1[p])[1]
            d2 = orb_oracle.describe(kf_b[p], keys2[p])[1]
            s1, s2 = np.flatnonzero(valid1[p]), np.flatnonzero(valid2[p])
            i1, e1, e2 = orb_oracle.hamming_top2(d1[s1], d2[s2])
            ok = ma.accept_bow(e1, e2) & (i1 >= 0)
            want = np.full(len(keys1[p]), -1, np.int32)
            want[s1[ok]] = s2[i1[ok]]
            submap["parity_ok"] = all_ranks_true(bool(np.array_equal(m12[p], want)))
            if rank == 0 and world == 1 and not args.no_cpu:
                from oracle import match_oracle
                match_oracle.build()
                t0 = time.perf_counter()
                npx = 0
                for _ in range(20):
                    for q in range(npairs):
                        npx += match_oracle.associate_pixels(keys1[q], valid1[q], keys2[q], valid2[q], (0, 0, W_IMG, H_IMG), 3.0)[0]
                dtc = time.perf_counter() - t0
                submap["cpu_baseline"] = {"value": 20 * npairs / dtc, "unit": "keyframe pairs/s", "cores": 1, "kind": "port",
                                          "sample": "the reference's own association of the same %d pairs: nearest key point within "
                                                    "3 px through KeyFrame::GetFeaturesInArea (CloudMerging.cc:503-551, "
                                                    "oracle/match_oracle.cpp), 20 repetitions, %.2f s -- a different, much cheaper "
                                                    "criterion (no descriptors): the reference has no descriptor path here" % (npairs, dtc),
                                          "associations": npx // 20}
        exs.close(); ma.close()

    # ---- bag of words: Frame::ComputeBoW tree descent of this rank's descriptors on an ORBvoc-shaped vocabulary
    #      (k = 10, L = 6: 1 111 111 nodes, 10^6 words; synthetic because ORBvoc.txt is a missing blob) ----
    bow = None
    if not args.no_bow:
        from rumi_slam_b200 import ORBVocabulary
        from rumi_slam_b200.synth import synthetic_vocabulary
        par, leaf, vdesc, vw = synthetic_vocabulary(10, 6, seed=5)
        voc = ORBVocabulary(10, 6, par, leaf, vdesc, vw, device=local)
        feats = out_dev[1][:, :1000].reshape(-1, 32).contiguous()           # 1000 descriptors of each of the frames
        for _ in range(3):
            voc.transform_features_device(feats, 4, sync=True)
        voc.launch_count(reset=True)
        barrier()
        t_a = time.perf_counter()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(msteps):
            voc.transform_features_device(feats, 4, sync=False)              # torch's stream is ordered behind each call
        ev1.record()
        torch.cuda.synchronize()
        barrier()
        windows.append((t_a, time.perf_counter()))
        ms_bow = max_over_ranks(ev0.elapsed_time(ev1))
        bow = {"workload": "DBoW2 tree descent (levelsup 4) of %d descriptors per GPU per step, vocabulary k=10 L=6 "
                           "(1111111 nodes, synthetic)" % feats.shape[0],
               "features_per_s": world * feats.shape[0] * msteps / (ms_bow / 1e3), "ms_per_step": ms_bow / msteps,
               "distances_per_feature": 60, "gpu_launches": voc.launch_count(reset=True), "steps": msteps}
        if rank == 0 and world == 1 and not args.no_cpu:
            from oracle import bow_oracle
            cpu_voc = bow_oracle.Vocabulary(10, 6, par, leaf, vdesc, vw)
            sample = feats[:100000].cpu().numpy()
            t0 = time.perf_counter()
            cw, _, cn = cpu_voc.transform(sample, 4)
            dt = time.perf_counter() - t0
            gw, _, gn = voc.transform_features_device(feats[:100000].contiguous(), 4, sync=True)
            bow["parity_ok"] = bool(np.array_equal(gw.cpu().numpy(), cw) and np.array_equal(gn.cpu().numpy(), cn))
            bow["cpu_baseline"] = {"value": len(sample) / dt, "unit": "features/s", "cores": 1, "kind": "port",
                                   "sample": "%d of the same descriptors, oracle/bow_oracle.cpp (pinned against the "
                                             "reference DBoW2), 1 thread, %.2f s" % (len(sample), dt)}

    # ---- KFDSample::Step's optical flow: the frames of this rank tracked one after the other,
    #      1000 points, 31x31 window, 3 levels, <= 20 iterations.  Sequential by construction (the points of frame k
    #      come from frame k-1), so this is a latency figure: one image H2D + 2 pyrDown + tracker + D2H + Scharr per call ----
    flow = None
    if not args.no_flow:
        from rumi_slam_b200 import SparsePyrLK
        lk = SparsePyrLK(31, 2, 20, 0.03, device=local)
        from rumi_slam_b200.synth import motion_sequence
        nseq = 65
        seq = motion_sequence(nseq, W_IMG, H_IMG, seed=500 + rank)         # camera translating 1.3 / -0.7 px per frame
        exf = ORBextractor(NFEAT, 1.2, 8, 20, 7, device=local)
        _, kp0, _ = exf(seq[0])
        exf.close()
        pts0 = np.stack([kp0["x"], kp0["y"]], 1).astype(np.float32)[:1000]

        def flow_pass():
            lk.set_prev(seq[0])
            cur, tracked = pts0, 0
            for i in range(1, nseq):
                cur, st, _ = lk.track_next(seq[i], cur, advance=True)      # old = next (KFDSample.cc:166)
                tracked += int(st.sum())
            return tracked
        flow_pass()
        l0 = lk.launches()
        barrier()
        t_a = time.perf_counter()
        tracked = flow_pass()
        t_b = time.perf_counter()
        barrier()
        windows.append((t_a, t_b))
        dt_flow = max_over_ranks((t_b - t_a) * 1e3) / 1e3
        flow = {"workload": "KFDSample::Step flow: %d consecutive 640x480 host frames, %d points each, "
                            "calcOpticalFlowPyrLK(31x31, maxLevel 2, 20 it, eps 0.03), synchronous calls; replicas only "
                            "(sequential by construction)" % (nseq - 1, len(pts0)),
                "frames_per_s": world * (nseq - 1) / dt_flow, "points_per_s": world * (nseq - 1) * len(pts0) / dt_flow,
                "ms_per_call": dt_flow / (nseq - 1) * 1e3, "gpu_launches": lk.launches() - l0,
                "tracked_fraction": tracked / float((nseq - 1) * len(pts0)),
                "timing": "host wall clock around synchronous calls (H2D, kernels, D2H inside)"}
        if rank == 0 and world == 1 and not args.no_cpu:
            from oracle import flow_oracle
            t0 = time.perf_counter()
            for i in range(1, 9):
                flow_oracle.lk(seq[i - 1], seq[i], pts0)
            dt = time.perf_counter() - t0
            flow["cpu_baseline"] = {"value": 8 / dt, "unit": "frames/s", "cores": 1, "kind": "port",
                                    "sample": "8 of the same frame pairs, oracle/flow_oracle.cpp (pinned against cv2), "
                                              "1 thread, %.2f s" % dt}
            try:
                import cv2
                cv2.setNumThreads(1)
                crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 20, 0.03)
                t0 = time.perf_counter()
                for i in range(1, 33):
                    cv2.calcOpticalFlowPyrLK(seq[i - 1], seq[i], pts0.reshape(-1, 1, 2), None, winSize=(31, 31),
                                             maxLevel=2, criteria=crit)
                flow["cv2_frames_per_s_1_thread"] = 32 / (time.perf_counter() - t0)
            except Exception:
                pass

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    clocks = sampler.stop(windows)

    # ---- roofline of the dominant kernel (stage with the largest share of the device time) ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (burst copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    alg = algorithmic_bytes(W_IMG, H_IMG, nkp_mean)
    total_stage_ms = sum(v[0] for v in stages.values()) or 1.0
    stage_rows = {}
    stages = {k: v for k, v in stages.items() if k in alg}      # h2d / d2h spans only exist in the e2e pass
    for name, (sms, nl) in stages.items():
        by = alg[name] * nframes * psteps                      # bytes this stage moved algorithmically in the pass
        gbs = by / (sms / 1e3) / 1e9 if sms > 0 else 0.0
        stage_rows[name] = {"ms": round(sms, 3), "share": round(sms / total_stage_ms, 4), "launches": nl,
                            "alg_GBps": round(gbs, 1), "frac_hbm": round(gbs / hbm_peak, 4)}
    dom = max(stages, key=lambda k: stages[k][0])
    dms, dn = stages[dom]
    per_launch_bytes = alg[dom] * nframes * psteps / max(dn, 1)
    per_launch_s = dms / 1e3 / max(dn, 1)
    achieved = per_launch_bytes / per_launch_s / 1e9 if per_launch_s > 0 else 0.0
    # DRAM traffic of the dominant kernel per launch (one chunk) from the committed ncu --set full capture
    traffic, ncu_ctx = None, None
    try:
        ncu = json.load(open(NCU_PROFILE))
        k = ncu["kernels"].get(dom)
        if k and ncu.get("frames_per_launch") == chunk:
            traffic = int(k["dram_read_bytes"] + k["dram_write_bytes"])
            ncu_ctx = {kk: k[kk] for kk in ("issue_active_pct", "alu_pipe_pct", "fma_pipe_pct", "lsu_pipe_pct",
                                            "warps_active_pct", "warp_instructions") if kk in k}
            ncu_ctx["source"] = "%s (ncu --set full, %d-frame chunk)" % (os.path.relpath(NCU_PROFILE, ROOT), chunk)
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": round(achieved, 2), "peak": hbm_peak, "unit": "GB/s",
                "frac": round(achieved / hbm_peak, 5), "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": int(per_launch_bytes), "avg_launch_ms": round(per_launch_s * 1e3, 4),
                "note": "the path is integer-issue bound, not HBM bound (SURVEY.md 8d): the dominant kernel's binding "
                        "resource is the ALU pipe, see ncu",
                "ncu": ncu_ctx,
                "stage_pass": {"streams": 1, "steps": psteps, "ms_per_step": round(ms_serial / psteps, 3)},
                "whole_step": {"algorithmic_bytes_per_frame": int(alg["frame"]),
                               "achieved": round(value * alg["frame"] / 1e9, 2),
                               "frac": round(value * alg["frame"] / 1e9 / hbm_peak, 5)},
                "stages": stage_rows}

    # ---- CPU baseline on the host cores (bounded sample, rank 0, N == 1 only) ----
    cpu = None
    if world == 1 and not args.no_cpu:
        sample = max(40 * cores, 64)
        pool = CpuExtractor(cores)
        pool.run(host[:2 * cores])
        dt = pool.run(host[:sample])
        pool.close()
        one = CpuExtractor(1)
        dt1 = one.run(host[:48])
        one.close()
        cpu = {"value": sample / dt, "unit": "frames/s", "cores": cores, "kind": pool.kind,
               "sample": "%d of the same synthetic frames on %d host threads (%.1f s); 1 thread: %.1f frames/s"
                         % (sample, cores, dt, 48 / dt1),
               "single_thread": 48 / dt1}

    if matching is not None:
        matching["roofline"] = match_roofline(match_path, pairs_per_s, world, clocks.get("sm_mhz") or 1965.0)
    parity_ok = bool(parity.get("device_path_ok", False) and parity.get("host_path_ok", False)
                     and parity.get("pageable_path_ok", False)
                     and (matching or {}).get("parity_ok", True) and (match_big or {}).get("parity_ok", True)
                     and all(v.get("parity_ok", True) for v in stereo.values())
                     and (submap or {}).get("parity_ok", True)
                     and (textured or {}).get("parity_ok", True)) if check_oracle else None
    line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD + ", batch of %d frames per GPU, frame-sharded" % nframes,
                       "frames_per_gpu": nframes, "chunk_frames": chunk, "mean_keypoints_per_frame": nkp_mean,
                       "distinct_frames": "%d generated per GPU, the rest are 2-D circular shifts of them" % args.unique,
                       "l2": "inputs (315 MB per GPU) exceed the 126 MB L2, no flush between steps",
                       "host_binding": numa, "host_topology": host_topology() if world > 1 else None},
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps, "h2d_GBps_raw": round(h2d_gbps, 2),
                    "frames_per_s_at_raw_h2d": round(ceiling, 1),
                    "e2e_over_raw_h2d_ceiling": round(e2e_value / ceiling, 4),
                    "pageable": {"value": e2e_pageable, "unit": "frames/s", "steps": pg_steps,
                                 "note": "same call, ordinary pageable host arrays in and out"}},
            "parity_checked_frames": parity.get("frames", 0), "parity_ok": parity_ok, "parity": parity,
            "gpu_launches": launches,
            "gpu_launches_e2e": launches_e2e,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "strong": strong,
            "matching": matching,
            "matching_5b": match_big,
            "stereo_euroc": stereo.get("stereo_euroc"),
            "stereo_kitti": stereo.get("stereo_kitti"),
            "submap_merge": submap,
            "bow": bow,
            "flow": flow,
            "single_frame_latency": latency,
            "textured_frames": textured,
            "clocks": clocks}
    emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def quiet_stdout():
    """Libraries (NCCL's version banner) write to fd 1; the contract is ONE JSON line on stdout.  Everything else goes
    to stderr, the JSON line is written to the saved descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(text):
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        print(text, flush=True)
    else:
        os.write(_REAL_STDOUT, (text + "\n").encode())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=1024)
    ap.add_argument("--chunk", type=int, default=int(os.environ.get("RUMI_CHUNK", "64")))
    ap.add_argument("--unique", type=int, default=64, help="distinct synthetic frames generated per rank")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-bow", action="store_true")
    ap.add_argument("--no-flow", action="store_true")
    ap.add_argument("--no-stereo", action="store_true")
    ap.add_argument("--no-submap", action="store_true")
    ap.add_argument("--no-textured", action="store_true")
    ap.add_argument("--no-check", action="store_true", help="skip the oracle parity checks (they run outside the timed regions)")
    ap.add_argument("--stereo-pairs", type=int, default=256)
    ap.add_argument("--match-big", type=int, default=1000000, help="cfg 5b descriptors per side (0 = skip)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    if args.gpus > 1 and "RANK" not in os.environ:          # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    quiet_stdout()
    run_b200(args)


if __name__ == "__main__":
    main()
