#!/usr/bin/env python
"""Per-SASS-instruction execution counts of one kernel from an .ncu-rep captured with --import-source on.
usage: ncu_source.py report.ncu-rep kernel-substring [min_count]"""
import csv
import io
import subprocess
import sys

rep, sub = sys.argv[1], sys.argv[2]
minc = int(sys.argv[3]) if len(sys.argv) > 3 else 0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
k, out, hdr = None, {}, None
for r in rows:
    if r and r[0] == "Kernel Name":
        k = r[1]; out.setdefault(k, []); out[k].clear(); continue
    if r and r[0] == "Address":
        hdr = r; continue
    if k and len(r) > 5:
        out[k].append(r)
name = [n for n in out if sub in n][0]
v = out[name]
ie, ss = hdr.index("Instructions Executed"), hdr.index("# Samples")
tot = sum(int(r[ie]) for r in v)
print(name, len(v), "static instr; total executed", tot)
for i, r in enumerate(v):
    if int(r[ie]) >= minc:
        print(f"{i:4d} {int(r[ie]):9d} {int(r[ss]):6d}  {r[1].strip()[:100]}")
