"""Writes profiles/r1_summary.md from profiles/r1_launches_final.csv, r1_ncu_kernels.json, r1_bench_final.json
(+ r1_bench_n2.json / r1_bench_n8.json when present)."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = lambda n: os.path.join(ROOT, "profiles", n)
k = json.load(open(P("r1_ncu_kernels.json")))
b = json.load(open(P("r1_bench_final.json")))
tab = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launch_summary.py"), P("r1_launches_final.csv"), "--md"],
                     capture_output=True, text=True).stdout
rows = []
for st, v in k["kernels"].items():
    rows.append("| %s | %d | %.1f | %.1f / %.1f | %.1f | %.0f | %.0f / %.0f / %.0f | %.0f | %d |" % (
        st, v["launches_per_chunk"], v["duration_us"], v["dram_read_bytes"] / 1e6, v["dram_write_bytes"] / 1e6,
        v["warp_instructions"] / 1e6, v["issue_active_pct"], v["alu_pipe_pct"], v["fma_pipe_pct"], v["lsu_pipe_pct"],
        v["warps_active_pct"], v["registers"]))
st = b["roofline"]["stages"]
fa = k["kernels"]["fast"]
scal = []
for name in ("r1_bench_n2.json", "r1_bench_n8.json"):
    if os.path.exists(P(name)):
        d = json.load(open(P(name)))
        big = d.get("matching_5b")
        scal.append("| %d | %.0f | %.0f | %.3g | %s | %s |" % (d["n_gpus"], d["value"], d["e2e"]["value"], d["matching"]["pairs_per_s"],
                                                             ("%.3g (%.0f ms)" % (big["pairs_per_s"], big["ms_per_step"])) if big else "-",
                                                             d["e2e"].get("h2d_GBps_raw")))
big1 = ""
if os.path.exists(P("r1_bench_5b_n1.json")):
    d1 = json.load(open(P("r1_bench_5b_n1.json")))
    big1 = "\n| 1 | %.0f | - | %.3g | %.3g (%.0f ms) | - |" % (d1["value"], d1["matching"]["pairs_per_s"], d1["matching_5b"]["pairs_per_s"],
                                                              d1["matching_5b"]["ms_per_step"])
md = f"""# Round 1 -- final state of the round (supersedes r1_first_path_summary.md, kept for history)

Workload of the captures: `tools/prof_run.py` = one 64-frame chunk of synthetic 640x480 frames (1000 features, 8 levels,
1.2, FAST 20/7) through `rumi_orb_extract_batch_device`, 2 warm-up + 2 profiled passes, then 8192 x 40000 Hamming top-2.

## 1. Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv`, `r1_launches_final.csv`)

{tab}
Shares agree with bench.py's own exclusive stage pass (CUDA events, one stream, `r1_bench_final.json`): pyramid {st['pyramid']['share']:.3f},
FAST {st['fast']['share']:.3f}, quad-tree {st['octree']['share']:.3f}, slots {st['slots']['share']:.3f}, blur {st['blur']['share']:.3f}, describe {st['describe']['share']:.3f} of the extraction step.

## 2. Per-kernel counters (`ncu --set full --clock-control none --import-source on`, `r1_ncu_kernels.json`)

| stage | launches/chunk | us/chunk | DRAM r / w (MB) | warp instr (M) | issue % | ALU / FMA / LSU pipe % | warps active % | regs |
|---|---|---|---|---|---|---|---|---|
""" + "\n".join(rows) + f"""

Reading: no kernel moves more DRAM bytes than its algorithmic bytes (FAST reads {fa['dram_read_bytes']/1e6:.1f} MB for 60.8 MB of pixels: levels
1-7 are still in L2; blur 63.7 + 18.5 MB for 121.7 MB), so nothing is re-read; every kernel is **instruction-issue
bound on the integer (ALU) pipe** or latency bound (quad-tree, the per-level pyramid launches), not HBM bound, exactly
as SURVEY.md 8d predicts.  FAST is the dominant kernel: {fa['warp_instructions']/1e6:.0f} M warp instructions per chunk = {fa['warp_instructions']*32/60.8e6:.0f} thread
instructions per pixel, ALU pipe {fa['alu_pipe_pct']:.0f} % of its peak, issue slots {fa['issue_active_pct']:.0f} % busy; its stalls are spread over fixed-latency
waits, "not selected" and math-pipe throttle, i.e. the schedulers are saturated for this instruction mix.  Where its
instructions go (`tools/sass_phases.py` on the source page): SWAR pretest 27 %, exact score 33 %, queue append 10 %,
NMS + emission 15 %, staging / set-up 11 %.

## 3. What changed during the round (per 64-frame chunk, ncu durations)

| kernel | first path | now | how |
|---|---|---|---|
| FAST | 993 us | {fa['duration_us']:.0f} us | SWAR 4-pixel pretest, ballot-compacted queue, both-polarity packed score (1 IMAD per ring pixel, 40 VIMNMX3.S16x2), unified tile/score pitch, branch-free NMS, in-place survivor list + rank emission, cell table |
| quad-tree | 276 us | {k['kernels']['octree']['duration_us']:.0f} us | closed-form level phase, block scans, register bitonic sort, parallel stable-rank replay of std::sort's insertion phase |
| blur | 268 us | {k['kernels']['blur']['duration_us']:.0f} us | register-marching warps, DP4A rows, 7-row register ring, ping-pong prefetch, no shared memory |
| pyramid | 112 us (TMA tiles) | {k['kernels']['pyramid']['duration_us']:.0f} us | marching warps, 3 word loads + 2 PRMT + 4 DP2A per source row (51 M vs 61 M warp instructions); TMA tiles kept for calls of < 8 frames (single frame: 32 vs 62 us) |
| describe | 70 us | {k['kernels']['describe']['duration_us']:.0f} us | unchanged |
| Hamming top-2, LOP3+POPC (8192 x 40000) | 637 us | {k['kernels']['hamming_top2']['duration_us']:.0f} us | LOP3 carry-save tree: 5 POPC per pair |
| Hamming top-2, tensor cores (16384 x 40000) | (1 170 us at the POPC kernel's rate) | {k['kernels']['hamming_top2_imma']['duration_us']:.0f} us | descriptors expanded once to 0/1 bytes, int8 IMMA m16n8k32 dot products, cp.async 3-stage tiles, ping-pong accumulators so the top-2 update overlaps the MMAs; IMMA pipe 51 % busy (ncu) |

Experiments that did NOT pay (kept as switches, documented in DESIGN.md): chaining all pyramid levels inside one launch with
completion flags (`RUMI_PYRAMID_SPLIT=1`: 127-377 us vs 102 us for 7 launches -- a dependent chain of latency-bound items),
more than 2 workspaces for resident input (L2 thrash), FAST at 48 / 56 registers (no change), 1 / 4 warps per FAST CTA.

## 4. bench.py on B200 (1965 MHz, no throttle reasons)

value **{b['value']:.0f} frames/s** (inputs resident), e2e **{b['e2e']['value']:.0f} frames/s** (pinned host buffers, H2D + D2H inside; raw H2D
{b['e2e']['h2d_GBps_raw']} GB/s would allow {b['e2e']['frames_per_s_at_raw_h2d']:.0f}), single-frame `operator()` latency {b['single_frame_latency']['median_ms']:.3f} ms, matching {b['matching']['pairs_per_s']:.3g} pairs/s
({b['matching']['kernel']} kernel, {b['matching']['roofline']['frac']:.0%} of the measured {b['matching']['roofline']['bound']} rate; the LOP3+POPC kernel gives 7.4e11), BoW descent {b['bow']['features_per_s']:.3g} features/s (CPU port {b['bow']['cpu_baseline']['value']:.3g} on one
thread), CPU reference arm {b['cpu_baseline']['value']:.0f} frames/s on {b['cpu_baseline']['cores']} host threads ({b['cpu_baseline']['single_thread']:.1f} on one).
"""
if scal:
    md += """
Multi-GPU (one process per GPU, frames sharded, train set sharded + NCCL all-gather of 8 B/query candidates):

| GPUs | value frames/s | e2e frames/s | matching 40k x 40k pairs/s | matching 10^6 x 10^6 pairs/s | raw H2D GB/s per GPU |
|---|---|---|---|---|---|
""" + "\n".join(scal) + big1 + """

`value` scales linearly (no data-path collective); cfg 5b (10^12 pairs) takes 1287 ms on one GPU and 161 ms on eight
(8.0x: the candidate all-gather is 8 MB per rank).  `e2e` at 8 GPUs is bound by the host links of the box: with 8 ranks
copying at once the raw pinned H2D bandwidth per GPU drops from 55 to 24 GB/s (single NUMA node VM, `nvidia-smi topo`),
i.e. 307 KB/frame caps each GPU at 78 k frames/s.
"""
open(P("r1_summary.md"), "w").write(md)
print(md[-1500:])
