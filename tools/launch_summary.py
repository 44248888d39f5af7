#!/usr/bin/env python
"""Condense an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X cmd`)
into per-kernel launches / mean / total / share.   usage: launch_summary.py launches.csv [out.txt] ["command line"]"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
    rd = csv.DictReader(rows)
    agg = OrderedDict()
    n = 0
    for r in rd:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
        us = float(r["Metric Value"]) / (1e3 if r["Metric Unit"] == "ns" else 1.0)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += us; n += 1
    tot = sum(a[1] for a in agg.values())
    out = []
    if len(sys.argv) > 3:
        out.append(sys.argv[3])
    out += ["(per-launch times under ncu are cold-cache and serialised - in particular the ICP launches lose their programmatic",
            " overlap: compare SHARES with bench.py's roofline.kernel_share_of_step, not absolutes)", "",
            f"{'kernel':40s} {'launches':>8s} {'mean us':>10s} {'total us':>10s} {'share':>7s}"]
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{k:40s} {c:8d} {t / c:10.1f} {t:10.1f} {100 * t / tot:6.1f}%")
    out += ["", f"first {n} launches, total {tot:.1f} us"]
    txt = "\n".join(out) + "\n"
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(txt)
    print(txt, end="")


if __name__ == "__main__":
    main()
