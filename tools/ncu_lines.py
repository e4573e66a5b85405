#!/usr/bin/env python
"""Executed warp instructions and stall samples per source line of one kernel in an ncu report (needs -lineinfo builds and
--import-source on):   python tools/ncu_lines.py <report.ncu-rep> <kernel name substring> [top]"""
import collections
import csv
import io
import subprocess
import sys


def main(rep, pat, top=40):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
    cur_file = cur_fn = hdr = None
    lines = collections.defaultdict(lambda: [0.0, 0.0])
    launches = 0
    for r in csv.reader(io.StringIO(out)):
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif r[0] == "Function Name":
            cur_fn = r[1]
        elif r[0] == "Line No":
            hdr = r
        elif hdr and r[0].strip().isdigit():
            try:
                v = (float(r[hdr.index("# Samples")] or 0), float(r[hdr.index("Instructions Executed")] or 0))
            except Exception:
                continue
            if pat in cur_fn:
                a = lines[(cur_file, int(r[0]), r[1][:100])]
                a[0] += v[0]
                a[1] += v[1]
    tot, tots = sum(a[1] for a in lines.values()), sum(a[0] for a in lines.values())
    print("total warp instructions %.4g, stall samples %.4g (all captured launches of the kernel)" % (tot, tots))
    byfile = collections.defaultdict(lambda: [0, 0])
    for (f, l, t), a in lines.items():
        byfile[f][0] += a[0]
        byfile[f][1] += a[1]
    for f, a in sorted(byfile.items(), key=lambda x: -x[1][1]):
        print("%-28s inst %5.1f%% smp %5.1f%%" % (f, 100 * a[1] / tot, 100 * a[0] / tots))
    print()
    for (f, l, t), a in sorted(lines.items(), key=lambda x: -x[1][1])[:int(top)]:
        print("%5.1f%% inst %5.1f%% smp  %s:%d  %s" % (100 * a[1] / tot, 100 * a[0] / tots, f, l, t))


if __name__ == "__main__":
    main(*sys.argv[1:])
