"""Joins an ncu SASS source-page CSV with `nvdisasm -gi` line info (outermost inlined-at frame in the given source
file) and prints warp-instruction counts / stall samples per source line range ("phase") and per opcode.
usage: sass_phases.py <ncu_source.csv> <cubin> <kernel-substring> <file.cu> name:lo-hi [name:lo-hi ...]"""
import collections
import csv
import re
import subprocess
import sys

src_csv, cubin, kern, fname = sys.argv[1:5]
phases = []
for a in sys.argv[5:]:
    n, r = a.split(":")
    lo, hi = r.split("-")
    phases.append((n, int(lo), int(hi)))
dis = subprocess.run(["nvdisasm", "-gi", cubin], capture_output=True, text=True).stdout
lines_for = []
in_fn, cur = False, None
pending = []
for ln in dis.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", ln)
    if m:
        in_fn = kern in m.group(1)
        cur = None
        continue
    if not in_fn:
        continue
    if "//## File" in ln:
        for m in re.finditer(r'"([^"]+)", line (\d+)', ln):
            if m.group(1).endswith(fname):
                cur = int(m.group(2))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        lines_for.append((cur, m.group(2).strip()))
rows = list(csv.reader(open(src_csv)))
hi_ = [i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r][0]
h = rows[hi_]
si, ii, sm = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
insts = [(r[si].strip(), int(r[ii]), int(r[sm]) if r[sm].isdigit() else 0) for r in rows[hi_ + 1:]
         if len(r) > ii and r[ii].isdigit()]
assert len(insts) == len(lines_for), (len(insts), len(lines_for))
tot = sum(i[1] for i in insts)
tots = sum(i[2] for i in insts)
agg = collections.OrderedDict((p[0], [0, 0, 0]) for p in phases)
agg["other"] = [0, 0, 0]
ops = collections.defaultdict(lambda: collections.defaultdict(int))
for (line, _), (sass, c, s) in zip(lines_for, insts):
    name = "other"
    for n, lo, hi in phases:
        if line is not None and lo <= line <= hi:
            name = n
            break
    a = agg[name]
    a[0] += c
    a[1] += s
    a[2] += 1
    op = sass.split()[0] if not sass.startswith("@") else sass.split()[1]
    ops[name][op.split(".")[0]] += c
print("total warp-instr %d, samples %d" % (tot, tots))
for n, (c, s, k) in agg.items():
    top = sorted(ops[n].items(), key=lambda kv: -kv[1])[:8]
    print("%-12s inst %11d %5.1f%%  samples %6d %5.1f%%  sass %4d | %s" % (
        n, c, 100.0 * c / tot, s, 100.0 * s / max(tots, 1), k,
        " ".join("%s:%.1f%%" % (o, 100.0 * v / tot) for o, v in top)))
