#!/usr/bin/env python
"""SASS instructions of one kernel ranked by one stall reason (ncu report with --import-source on), with a few lines of context:
    python tools/ncu_sass.py <report.ncu-rep> <kernel substring> [stall column = stall_no_inst] [top = 20]"""
import collections
import csv
import io
import subprocess
import sys


def main(rep, pat, col="stall_no_inst", top=20):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], capture_output=True, text=True).stdout
    hdr, on, agg = None, False, collections.OrderedDict()
    for r in csv.reader(io.StringIO(out)):
        if not r:
            continue
        if r[0] == "Kernel Name":
            on = pat in r[1]
        elif r[0] == "Address":
            hdr = r
        elif on and hdr and len(r) >= len(hdr) - 1:
            a = agg.setdefault(int(r[0], 16), [r[1].strip(), 0.0, 0.0, 0.0])
            a[1] += float(r[hdr.index(col)] or 0)
            a[2] += float(r[hdr.index("Instructions Executed")] or 0)
            a[3] += float(r[hdr.index("# Samples")] or 0)
    addrs = list(agg)
    base = addrs[0]
    tot = sum(a[1] for a in agg.values())
    print(f"{col}: {tot:.0f} samples over {len(addrs)} instructions")
    for i in sorted(range(len(addrs)), key=lambda i: -agg[addrs[i]][1])[:int(top)]:
        a = agg[addrs[i]]
        print(f"--- {100 * a[1] / tot:5.2f}%  +0x{addrs[i] - base:05x}  executed {a[2]:.3g}")
        for j in range(max(0, i - 3), min(len(addrs), i + 2)):
            b = agg[addrs[j]]
            print(f"      {'>>' if j == i else '  '} +0x{addrs[j] - base:05x}  {b[0][:80]:80s} {b[1]:7.0f} / {b[3]:7.0f}")


if __name__ == "__main__":
    main(*sys.argv[1:])
