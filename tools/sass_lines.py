"""Joins an ncu SASS source-page CSV (per-instruction counters) with nvdisasm line info of the same kernel, and
prints instruction counts / stall samples aggregated per CUDA source line.
usage: sass_lines.py <ncu_source.csv> <cubin> <kernel-substring> [top]"""
import csv, re, subprocess, sys, collections
src_csv, cubin, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout
# walk nvdisasm output: track current function, current line
lines_for = []
cur_fn, cur_line, in_fn = None, None, False
for ln in dis.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", ln)
    if m:
        cur_fn = m.group(1); in_fn = kern in cur_fn; cur_line = None; continue
    if not in_fn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur_line = (m.group(1).split("/")[-1], int(m.group(2)))
        # inlined-at info follows sometimes; keep innermost
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        lines_for.append((cur_line, m.group(2).strip()))
rows = list(csv.reader(open(src_csv)))
hi = [i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r][0]
h = rows[hi]; si, ii, sm = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
insts = [(r[si].strip(), int(r[ii]), int(r[sm]) if r[sm].isdigit() else 0) for r in rows[hi + 1:] if len(r) > ii and r[ii].isdigit()]
print("sass instrs: ncu %d, nvdisasm %d" % (len(insts), len(lines_for)))
agg = collections.defaultdict(lambda: [0, 0, 0])
n = min(len(insts), len(lines_for))
tot = sum(i[1] for i in insts); tots = sum(i[2] for i in insts)
for k in range(n):
    a = agg[lines_for[k][0]]; a[0] += insts[k][1]; a[1] += insts[k][2]; a[2] += 1
print("total warp-instr %d samples %d" % (tot, tots))
for key, (c, s, ninst) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-22s inst %11d %5.1f%%  samples %6d %5.1f%%  sass %3d" % ("%s:%s" % key if key else "?", c, 100.0 * c / tot, s, 100.0 * s / max(tots, 1), ninst))
