#!/usr/bin/env python
"""Opcode histogram per kernel of libhkf_b200.so (cuobjdump -sass), for profiles/: which Blackwell-specific instructions
the built library actually contains (bulk async copies UBLKCP, mbarrier SYNCS, programmatic-dependent-launch ACQBULK /
PREEXIT, cooperative-grid machinery) next to the plain arithmetic / memory mix.
usage: sass_histogram.py [library] > profiles/rN_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "hybkinectfu_b200", "libhkf_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], stdout=subprocess.PIPE, text=True).stdout.strip()
kern, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = m.group(1); hist[kern] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", line)
    if m and kern:
        hist[kern][m.group(1).split(".")[0]] += 1
special = ("UBLKCP", "UTMALDG", "SYNCS", "LDGSTS", "ACQBULK", "PREEXIT", "UCGABAR", "CCTL", "MEMBAR", "ERRBAR", "ATOMG", "REDG", "ATOMS", "MUFU", "DADD", "F2F", "R2UR")
print(f"{lib}: {len(hist)} kernels\n")
for k, h in hist.items():
    tot = sum(h.values())
    name = demangle(k)
    name = re.sub(r"\(.*", "", name)
    print(f"== {name}   ({tot} SASS instructions)")
    print("   top: " + ", ".join(f"{o} {n}" for o, n in h.most_common(12)))
    sp = [f"{o} {h[o]}" for o in special if h.get(o)]
    if sp:
        print("   special: " + ", ".join(sp))
