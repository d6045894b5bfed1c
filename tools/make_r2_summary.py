"""Writes profiles/r2_summary.md from the round-2 evidence files in profiles/ (bench JSONs, ncu launch list, per-kernel ncu
counters).  usage: python tools/make_r2_summary.py"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = lambda n: os.path.join(ROOT, "profiles", n)
d = json.load(open(P("r2_bench_n1.json")))
ref = json.load(open(P("r2_bench_ref.json")))
k = json.load(open(P("r2_ncu_kernels.json")))["kernels"]
tab = open(P("r2_ncu_kernels.md")).read()
ls = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launch_summary.py"), P("r2_launches_bench.csv"), "--md"],
                    capture_output=True, text=True).stdout
st = d["roofline"]["stages"]
rows = []
for n in (1, 2, 4, 8):
    if not os.path.exists(P("r2_bench_n%d.json" % n)):
        continue
    b = json.load(open(P("r2_bench_n%d.json" % n)))
    m, big, e = b["matching"], b["matching_5b"], b["e2e"]
    qs = m.get("query_sharded", {}).get("ms_per_step")
    s = b.get("strong")
    rows.append("| %d | %.0f | %.0f | %.2f (%.1f GB/s) | %s | %.3f / %s | %.1f ms = %.2e |" % (
        n, b["value"], e["value"], e["e2e_over_raw_h2d_ceiling"], e["h2d_GBps_raw"],
        ("%.0f / %.0f" % (s["value"], s["e2e"])) if s else "-", m["ms_per_step"], ("%.3f" % qs) if qs else "-",
        big["ms_per_step"], big["pairs_per_s"]))
md = f"""# Round 2 -- final state (B200, one GPU unless stated)

Everything here comes from `tools/r2_profile_all.sh` on the final code (GPU tests 148 passed / 1 skipped, smoke ok, then bench,
reference arm and the ncu passes) and, for section 5, from `bench.py --gpus N` under torchrun.  Bench numbers are never taken
under ncu.  (This file is generated: `python tools/make_r2_summary.py`.)

## 1. Bench (`r2_bench_n1.json`, `python bench.py`, clocks {d['clocks']['sm_mhz']:.0f} MHz, throttle reasons: {d['clocks']['reasons'] or 'none'})

| row | round 1 | round 2 |
|---|---|---|
| `value`: ORB frames/s, 1024 x 640x480 resident in HBM | 148 500 | **{d['value']:.0f}** |
| `e2e`: pinned host buffers in and out | 127 900 | **{d['e2e']['value']:.0f}** ({d['e2e']['e2e_over_raw_h2d_ceiling']:.2f} of the raw H2D ceiling, {d['e2e']['h2d_GBps_raw']} GB/s = {d['e2e']['frames_per_s_at_raw_h2d']:.0f} frames/s) |
| pageable host arrays | - | {d['e2e']['pageable']['value']:.0f} |
| single-frame `operator()` latency (configs[0]) | 0.28 ms | **{d['single_frame_latency']['median_ms']:.3f} ms** |
| reference arm, 16 host threads (`cpu_baseline` / `--impl reference`) | 630 / ~300 | {d['cpu_baseline']['value']:.0f} / {ref['value']:.0f} |
| cfg 5a, 40 000^2 top-2 | 0.38 ms, 4.2e12 pairs/s | **{d['matching']['ms_per_step']:.3f} ms, {d['matching']['pairs_per_s']:.2e} pairs/s** ({d['matching']['roofline']['frac']:.2f} of nominal int8) |
| cfg 5b, 10^6 x 10^6 top-2 | 174 ms, 5.7e12 | **{d['matching_5b']['ms_per_step']:.0f} ms, {d['matching_5b']['pairs_per_s']:.2e}** ({d['matching_5b']['roofline']['frac']:.2f} of nominal int8, {d['matching_5b']['roofline']['frac_of_2x_measured_bf16']:.2f} of 2 x the measured cuBLAS bf16 burst rate) |
| cfg 3 (EuRoC stereo) / cfg 4 (KITTI stereo) | - | {d['stereo_euroc']['pairs_per_s']:.0f} / {d['stereo_kitti']['pairs_per_s']:.0f} pairs/s (CPU 16 threads: {d['stereo_euroc']['cpu_baseline']['value']:.0f} / {d['stereo_kitti']['cpu_baseline']['value']:.0f}) |
| submap merge, 40 + 40 key frames | - | {d['submap_merge']['ms_per_merge']:.1f} ms |
| BoW descent | 2.5e9 features/s | {d['bow']['features_per_s']:.2e} |
| optical-flow step (1000 points) | 0.115 ms | {d['flow']['ms_per_call']:.3f} ms |

`parity_ok`: {str(d['parity_ok']).lower()} (128 frames over all 16 chunks of the device, pinned-host and pageable runs; all 40 000 queries of
cfg 5a; 2 000 sampled queries of cfg 5b; every stereo pair and key-frame pair).  `gpu_launches`: {d['gpu_launches']} in the timed
region = 16 chunks x 5 kernels x 10 steps.

Exclusive stage times of the 1024-frame batch on one stream (`roofline.stages`, ms per 5 steps / share):
""" + "\n".join(f"* {n}: {v['ms']:.2f} ms, {v['share']*100:.1f} %" + (" (fused into the quad-tree kernel: no launch of its own)" if n == "slots" else "") for n, v in st.items()) + f"""

## 2. Launch list of bench.py (`r2_launches_bench.csv`, `ncu --metrics gpu__time_duration.sum`, first 900 launches)

{ls}
Shares agree with the stage shares above.  Exactly five launches per 64-frame chunk: pyramid (all levels), FAST (all levels),
quad-tree (+ slot assignment in its tail), blur (all levels), descriptors.

## 3. Per-kernel counters (`r2_ncu_kernels.json`: `ncu --set full --clock-control none --import-source on`, `PROF_LIGHT=1 tools/prof_run.py`)

{tab}
(One 64-frame chunk; `hamming_top2` = K8 on 8192 x 40000, `hamming_top2_umma` = K8-U on 16384 x 40000.  ncu flushes caches
between kernels, so DRAM bytes are cold-cache figures: in the pipeline levels 1-7 are still in L2 when FAST / blur read them.)

## 4. What changed in round 2 and what the counters say

| kernel | round 1 | round 2 | how |
|---|---|---|---|
| pyramid | 7 launches, 111 us / chunk | **1 launch, {k['pyramid']['duration_us']:.0f} us** | strip kernel: a CTA carries one strip of one frame through all levels, levels 2-7 read from shared memory; packed-pair vertical filter. {k['pyramid']['warp_instructions']/1e6:.0f} M warp instructions (was 51 M), issue {k['pyramid']['issue_active_pct']:.0f} %, DRAM {k['pyramid']['dram_read_bytes']/1e6:.1f} MB read (= level 0 once) / {k['pyramid']['dram_write_bytes']/1e6:.1f} MB written (the other levels stay in L2).  Algorithmic 1.57 MB per frame -> 1.8 TB/s = 27 % of the HBM burst figure; the limiter is issue (10 integer instructions per output pixel), not memory.  VERDICT r1 asked for <= 35 us: not reached |
| quad-tree (+ slots) | 110 + 9 us | **{k['octree']['duration_us']:.0f} us** | bucket sort on path-code digits instead of a 64-bit bitonic network, parallel rank sort of the leaf list, serial std::sort replay only when equal keys meet in the consumed suffix; the slot assignment runs in the tail of the last CTA of a frame (VERDICT r1: <= 60 us for the quad-tree alone: 55 us before the fusion) |
| descriptors | 72 us | {k['describe']['duration_us']:.0f} us | blurred 37x37 window staged in shared memory (needed after the pyramid / quad-tree changes raised concurrency: the gather version rose to 102 us) |
| FAST | 167 us | {k['fast']['duration_us']:.0f} us | pretest on 8 pixels per lane (two column groups per trip share the row loads), queue compaction from four ballots on the per-lane hit count.  ALU pipe {k['fast']['alu_pipe_pct']:.0f} %, issue {k['fast']['issue_active_pct']:.0f} % -- 36 % of the step and the kernel `roofline` reports |
| blur | 77 us | {k['blur']['duration_us']:.0f} us alone | strips of 96 rows for chunks: {k['blur']['warp_instructions']/1e6:.1f} M warp instructions instead of 54.6 M (less halo recomputation), slower alone (longer items) but +1 % on the pipelined batch; VERDICT r1 asked for <= 40 us: not reached, the 8 DP4A + 6 funnel shifts + 28 multiply-adds per 4 pixels are the floor of this formulation |
| K8-U top-2 (40 000^2) | 0.38 ms, tensor pipe 54 % | **{d['matching']['ms_per_step']:.3f} ms, tensor pipe 72.5 % of elapsed / 75.9 % of active** (`r2_ncu_umma_keys.txt`, `utcimma_src_int8` 72.5 % of its peak) | pop(t) as a ninth K step + max-tree epilogue on raw accumulators; persistent CTAs over flattened ranges; converged issuer warp with uniform-register operands; query operand in TMEM (see DESIGN.md "K8-U") |

K8-U in numbers: 148 persistent CTAs x 576 threads, 68 registers; issue {k['hamming_top2_umma']['issue_active_pct']:.0f} %, ALU pipe {k['hamming_top2_umma']['alu_pipe_pct']:.0f} %: the
epilogue is no longer the limiter.  Shared-memory operand reads by the tensor core (`l1tex__data_pipe_tc_wavefronts_mem_shared`)
are at 36 % of peak (they were ~90 % with both operands in shared memory).  On long scans (262 144 x 10^6) the kernel delivers
3.31 POP/s; cuBLAS bf16 8192^3 on the same pool measures 1.667 PFLOP/s burst (`MEASURED_PEAKS.json`), i.e. 3.33 POP/s int8-equivalent.

SASS evidence (`r2_sass_tcgen05.txt`): `hamming_top2_umma_kernel` has 108 `UTCIMMA` (A operand `tmem[...]`), `UTCBAR`
(tcgen05.commit), `LDTM.x32`, `STTM.x32`, `UBLKCP.S.G` (cp.async.bulk ring), `SYNCS` (mbarriers), `REDUX` / `ELECT` (uniform
issuer); `pyramid_level_kernel<true>` has `UTMALDG.3D` (TMA tensor load); blur / pyramid use `IDP.4A` / `IDP.2A`; FAST
`VABSDIFF4` / `VIMNMX3.S16x2`.

Dense (textured) frames -- `r2_bench_textured_row.json`, `tools/density_probe.py`, `tools/octree_clocks_dense.py`: the benchmark
frames put ~1 400 FAST candidates on level 0, real images 3-10 k.  Quad-tree stage of ONE frame vs level-0 candidates
(1 416 / 2 182 / 5 525 / 10 738): 47 / 96 / 241 / 569 us before, 43 / 49 / 74 / 91 us now -- second pass with a 16 384-key
shared-memory buffer (adaptive per handle: sparse workloads keep the single launch), loop-form bucket sort (442 k -> 112 k
cycles at 10.7 k keys), 1 024-thread CTAs where a CTA has its SM to itself.  256 frames with +-20 grey levels of noise (~5 k
level-0 candidates): 152 k frames/s resident, 0.196 ms per single frame, parity_ok against the oracle.

## 5. Multi-GPU (`r2_bench_n2.json`, `r2_bench_n4.json`, `r2_bench_n8.json`: `bench.py --gpus N` under torchrun on one 8-GPU box)

| N | `value` frames/s (weak, 1024 frames per GPU) | `e2e` frames/s | e2e / raw-H2D ceiling (raw H2D per GPU) | strong: ONE 1024-frame batch, value / e2e | cfg 5a ms: train-sharded + NCCL / query-sharded | cfg 5b (10^6 x 10^6, train-sharded + NCCL) |
|---|---|---|---|---|---|---|
""" + "\n".join(rows) + """

`parity_ok` true at every N (incl. sharded == unsharded on all queries of cfg 5a and of cfg 5b).  Extraction and both matcher
configurations scale linearly on the device; `e2e` does not: the box's host links deliver 55 GB/s to one GPU but ~23 GB/s per
GPU when eight copy at once, so the end-to-end rate stays at 0.77-0.87 of what the host links allow at every N.
`tests/test_gpu_match.py -k shard` (4 tests incl. the NCCL path through `rumi_hamming_top2_sharded`) passes at N = 2.
"""
open(P("r2_summary.md"), "w").write(md)
print("written", len(md))
