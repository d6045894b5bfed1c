"""Condenses an `ncu --set full` report of tools/prof_run.py into profiles/<name>.json + a markdown table.
usage: ncu_to_json.py <report.ncu-rep> <out.json> <frames_per_launch> [<launch-list.csv>]
Per stage (= what bench.py calls a stage): duration, DRAM bytes, warp instructions, pipe utilisation, averaged over the
captured launches; the pyramid stage is the SUM of its per-level launches of one chunk."""
import collections
import csv
import json
import subprocess
import sys

rep, out_json, fpl = sys.argv[1], sys.argv[2], int(sys.argv[3])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
H, U = rows[0], rows[1]
col = {h: i for i, h in enumerate(H)}
STAGE = [("hamming_top2_umma", "hamming_top2_umma"), ("umma_pack_train", "umma_pack_train"),
         ("umma_fill_partial", "umma_fill_partial"), ("hamming_candidates", "hamming_candidates"),
         ("hamming_top2_segments", "hamming_top2_pairs"),
         ("flow_lk", "flow_lk"), ("flow_pyrdown", "flow_pyrdown"), ("flow_scharr", "flow_scharr"),
         ("fast_kernel", "fast"), ("octree_kernel", "octree"), ("blur_kernel", "blur"),
         ("describe_given", "describe_given"), ("describe_kernel", "describe"),
         ("assign_slots", "slots"), ("pyramid", "pyramid"), ("hamming_top2", "hamming_top2"),
         ("top2_merge_packed", "top2_merge_packed"), ("top2_merge", "top2_merge"), ("stereo", "stereo")]


def val(r, name, scale=1.0):
    if name not in col:
        return None
    try:
        v = float(r[col[name]].replace(",", ""))
    except ValueError:
        return None
    u = U[col[name]]
    mult = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ns": 1e-3, "us": 1.0, "ms": 1e3, "msecond": 1e3,
            "usecond": 1.0, "nsecond": 1e-3}.get(u, 1.0)
    return v * mult * scale


acc = collections.OrderedDict()
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    st = next((s for k, s in STAGE if k in name), None)
    if not st:
        continue
    acc.setdefault(st, []).append(r)
kernels = {}
for st, rs in acc.items():
    n = len(rs)
    per_chunk = 1
    if st == "pyramid":
        grids = [int(float(r[col["launch__grid_size"]])) for r in rs]
        per_chunk = max(1, grids.count(grids[0]) and n // grids.count(grids[0]))
    k = n / per_chunk                                          # chunks captured
    s = lambda name: sum(val(r, name) or 0.0 for r in rs)
    a = lambda name: sum(val(r, name) or 0.0 for r in rs) / n
    kernels[st] = {
        "launches_captured": n, "launches_per_chunk": per_chunk,
        "duration_us": round(s("gpu__time_duration.sum") / k, 2),
        "dram_read_bytes": int(s("dram__bytes_read.sum") / k), "dram_write_bytes": int(s("dram__bytes_write.sum") / k),
        "warp_instructions": int(s("smsp__inst_executed.sum") / k),
        "issue_active_pct": round(a("smsp__issue_active.avg.pct_of_peak_sustained_active"), 1),
        "alu_pipe_pct": round(a("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"), 1),
        "fma_pipe_pct": round(a("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"), 1),
        "lsu_pipe_pct": round(a("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"), 1),
        "warps_active_pct": round(a("sm__warps_active.avg.pct_of_peak_sustained_active"), 1),
        "registers": int(a("launch__registers_per_thread")),
        "dram_pct_of_peak": round(a("dram__throughput.avg.pct_of_peak_sustained_elapsed"), 1),
    }
doc = {"report": rep.split("/")[-1], "frames_per_launch": fpl,
       "command": "ncu --set full --clock-control none --import-source on -k regex:rumi python tools/prof_run.py",
       "kernels": kernels}
json.dump(doc, open(out_json, "w"), indent=1)
print("| stage | launches/chunk | us/chunk | DRAM r / w (MB) | warp instr (M) | issue % | ALU / FMA / LSU pipe % | warps active % | regs |")
print("|---|---|---|---|---|---|---|---|---|")
for st, k in kernels.items():
    print("| %s | %d | %.1f | %.1f / %.1f | %.1f | %.0f | %.0f / %.0f / %.0f | %.0f | %d |" % (
        st, k["launches_per_chunk"], k["duration_us"], k["dram_read_bytes"] / 1e6, k["dram_write_bytes"] / 1e6,
        k["warp_instructions"] / 1e6, k["issue_active_pct"], k["alu_pipe_pct"], k["fma_pipe_pct"], k["lsu_pipe_pct"],
        k["warps_active_pct"], k["registers"]))
