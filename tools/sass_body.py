#!/usr/bin/env python
"""Static size and composition of a kernel's MAIN BODY (the code before its out-of-line callees):
    python tools/sass_body.py <cubin> <kernel substring> [top]"""
import collections
import glob
import re
import subprocess
import sys


def main(cubin, pat, top=30):
    txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.split("\n")
    start = next(i for i, l in enumerate(txt) if l.startswith(".text.") and pat in l)
    cur, recs = None, []
    for l in txt[start + 1:]:
        if l.startswith("//-----") or l.startswith("\t.section"):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            recs.append((int(m.group(1), 16), cur, m.group(2).strip()))
    rets = [a for a, c, t in recs if "RET" in t.split()[0] or (t.startswith("@") and "RET" in t)]
    first_ret = min(rets) if rets else 1 << 60
    end = max(a for a, c, t in recs if a < first_ret and "EXIT" in t)
    body = [r for r in recs if r[0] <= end]
    print("main body: %d instructions (%.1f KB); whole function with callees: %d" % (len(body), len(body) * 16 / 1024, len(recs)))
    cnt = collections.Counter(c for a, c, t in body)
    src = {}
    for c in cnt:
        if c and c[0] not in src:
            paths = [p for p in glob.glob("/root/repo/3d_reconstruction_project_b200/csrc/*") + glob.glob("/usr/local/cuda/include/*") +
                     glob.glob("/usr/local/cuda/include/crt/*") if p.endswith("/" + c[0])]
            src[c[0]] = open(paths[0], errors="ignore").read().split("\n") if paths else []
    for c, v in cnt.most_common(int(top)):
        lines = src.get(c[0], []) if c else []
        text = lines[c[1] - 1].strip()[:100] if c and c[1] - 1 < len(lines) else ""
        print("%4d  %s:%s  %s" % (v, c[0] if c else None, c[1] if c else None, text))


if __name__ == "__main__":
    main(*sys.argv[1:])
