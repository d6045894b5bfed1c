"""Lists the Blackwell-specific / packed-integer SASS mnemonics of every kernel in librumi_orb.so (evidence that the hot
paths run tcgen05 / TMEM / bulk copies / packed SIMD).  usage: cuobjdump -sass <lib.so> | python tools/sass_evidence.py"""
import collections
import re
import subprocess
import sys

fn = None
cnt, first = collections.OrderedDict(), {}
pat = re.compile(r"\b(UTCIMMA|UTCBAR|LDTM|STTM|UBLKCP|UTMALDG|UTMASTG|SYNCS|UTCATOMSWS|REDUX|ELECT|IDP|VIMNMX3|VABSDIFF4)")
for l in sys.stdin:
    m = re.search(r"Function : (\S+)", l)
    if m:
        d = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        d = d.replace("(anonymous namespace)::", "").replace("rumi::", "")
        fn = d.split("(")[0].replace("void ", "")
        continue
    m = pat.search(l)
    if m and fn:
        k = (fn, m.group(1))
        cnt[k] = cnt.get(k, 0) + 1
        first.setdefault(k, re.sub(r"\s*/\*\s*0x[0-9a-f]*\s*\*?/?\s*$", "", l.rstrip()).strip())
print("# cuobjdump -sass rumi_slam_b200/librumi_orb.so (sm_100a): Blackwell-specific / packed-integer mnemonics per kernel")
print("# kernel | mnemonic | count | first occurrence")
for (f, mn), c in cnt.items():
    print("%s | %s | %d | %s" % (f, mn, c, first[(f, mn)][:120]))
