#!/usr/bin/env python
"""Counts the Blackwell-specific SASS instructions per kernel of the built library and prints the staging / scan excerpt of the
ICP pass:   python tools/sass_excerpt.py > profiles/r02_sass_excerpt.txt   (cuobjdump -sass, no GPU needed)"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "3d_reconstruction_project_b200", "libb200recon.so")
PAT = re.compile(r"\b(UBLKCP|SYNCS|FFMA2|CREDUX|REDUX|FMNMX3|UTMALDG|UTMASTG)\b[.\w]*")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    names = {}
    counts = collections.defaultdict(collections.Counter)
    body = collections.defaultdict(list)
    fn = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            continue
        if fn is None or "/*" not in line or line.strip().startswith("/* 0x"):
            continue
        ins = re.sub(r"/\*[0-9a-fx ]+\*/", "", line).strip()
        body[fn].append(ins)
        m = PAT.search(ins)
        if m:
            counts[fn][m.group(0)] += 1
    dem = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
    for k, d in zip(counts, dem):
        names[k] = re.sub(r"b3d::\(anonymous namespace\)::", "", d).split("(")[0]
    print("# cuobjdump -sass of 3d_reconstruction_project_b200/libb200recon.so (sm_100a): Blackwell-era instructions per kernel")
    print("# UBLKCP = cp.async.bulk (global -> shared bulk copy), SYNCS.* = mbarrier (ARRIVE.TRANS64 = arrive / expect-tx,")
    print("# PHASECHK...TRYWAIT = try_wait.parity), FFMA2 = packed fma.rn.f32x2, CREDUX / REDUX = single-instruction warp reduction,")
    print("# FMNMX3 = three-input min / max")
    for k in sorted(counts, key=lambda k: names[k]):
        print(f"{names[k]:60s} " + "  ".join(f"{op} x{n}" for op, n in sorted(counts[k].items())))
    key = next((k for k in counts if "icp_pass2_kernelILi1" in k), None)
    if key:
        print("\n# excerpt of icp_pass2_kernel<1>: box agreement (CREDUX), the bulk copies of the probed cells on one mbarrier, the packed scan")
        keep = re.compile(r"UBLKCP|SYNCS|FFMA2|CREDUX|FMNMX3|LDS\.128|FMNMX ")
        shown = 0
        for ins in body[key]:
            if keep.search(ins) and "SYNCS.CCTL" not in ins:
                print("    " + ins)
                shown += 1
                if shown >= 90:
                    break


if __name__ == "__main__":
    main()
