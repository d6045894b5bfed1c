"""Summarises an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel. usage: launch_summary.py <csv> [--md]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    name = r[ki].split("(")[0]
    v = float(r[vi].replace(",", ""))
    v = v / 1000 if r[ui] == "ns" else v * 1000 if r[ui] == "ms" else v
    agg.setdefault(name, []).append(v)
own = {k: v for k, v in agg.items() if "rumi::" in k}
tot = sum(sum(v) for v in own.values())
md = "--md" in sys.argv
if md:
    print("| kernel | launches | total us | avg us | share of own kernels |\n|---|---|---|---|---|")
for k, v in sorted(own.items(), key=lambda kv: -sum(kv[1])):
    if md:
        print("| %s | %d | %.1f | %.1f | %.3f |" % (k.replace("void ", ""), len(v), sum(v), sum(v) / len(v), sum(v) / tot))
    else:
        print("%-50s n=%3d total=%9.1f us avg=%8.1f share=%.3f" % (k[:50], len(v), sum(v), sum(v) / len(v), sum(v) / tot))
