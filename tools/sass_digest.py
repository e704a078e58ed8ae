#!/usr/bin/env python
"""SASS digest of the shipped library: per kernel, the counts of the mnemonics that prove which hardware paths it
uses (tcgen05 MMA = UTCHMMA, TMEM loads = LDTM, TMA bulk copies = UBLKCP, tcgen05.commit = UTCBAR, mbarrier = SYNCS,
packed FP32 = FFMA2/FADD2/FMUL2), plus registers / spills from `cuobjdump -res-usage`.
    python tools/sass_digest.py [lib.so] > profiles/r2_sass_digest.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "lightpycl_b200", "_lib", "liblightpycl_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
usage = {}
for m in re.finditer(r"Function (\S+):\n\s*(.*)", res):
    usage[m.group(1)] = m.group(2).strip()
WATCH = ["UTCHMMA", "UTCBAR", "LDTM", "UBLKCP", "SYNCS", "ELECT", "VOTE", "FFMA2", "FADD2", "FMUL2", "FFMA", "LOP3", "R2UR",
         "ATOMG", "RED", "ATOMS", "LDL", "STL", "BAR", "NANOSLEEP"]
cur, counts, total = None, {}, {}
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur], total[cur] = collections.Counter(), 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        total[cur] += 1
        op = m.group(1)
        for w in WATCH:
            if op == w or op.startswith(w + "."):
                counts[cur][w] += 1
print("# SASS digest of %s (sm_100a), from cuobjdump -sass / -res-usage" % os.path.relpath(lib, ROOT))
for fn in sorted(counts, key=lambda f: -total[f]):
    name = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()[:110]
    print("\n%s\n  %d instructions; %s" % (name, total[fn], usage.get(fn, "")))
    print("  " + "  ".join("%s %d" % (w, counts[fn][w]) for w in WATCH if counts[fn][w]))
