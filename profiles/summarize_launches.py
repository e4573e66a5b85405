#!/usr/bin/env python
"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, total us, share)."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, start = r, i + 1
            break
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    mi = hdr.index("Metric Name") if "Metric Name" in hdr else None
    agg = collections.OrderedDict()
    for r in rows[start:]:
        if len(r) <= vi:
            continue
        if mi is not None and r[mi] != "gpu__time_duration.sum":  # launch lists may carry further metrics per launch
            continue
        name = r[ki].replace("<unnamed>::", "").replace("unnamed>::", "").replace("b3d::", "").replace("void ", "").split("(")[0]
        if name.startswith("compact_kernel<"):
            name = "compact_kernel<" + name[len("compact_kernel<"):].split(",")[0].strip() + ">"
        else:
            name = name.split("<")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    print(f"{'kernel':48s} {'launches':>8s} {'total_us':>12s} {'share':>7s}")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{k:48s} {a[0]:8d} {a[1] / 1e3:12.1f} {100 * a[1] / tot:6.1f}%")
    print(f"{len(rows) - start} launches, {tot / 1e6:.3f} ms of kernel time (cold-cache, serialised: compare SHARES, not absolutes)")


if __name__ == "__main__":
    main(sys.argv[1])
