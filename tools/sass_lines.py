#!/usr/bin/env python
"""Static SASS instruction counts per source line of one kernel (nvdisasm -g -c on a cubin extracted with cuobjdump -xelf):
    python tools/sass_lines.py <cubin> <kernel name substring> [top]
Used to see where a kernel's code size (and a chunk's instruction budget) goes without a GPU."""
import collections
import re
import subprocess
import sys


def main(cubin, pat, top=40):
    txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.split("\n")
    start = next(i for i, l in enumerate(txt) if l.startswith(".text.") and pat in l)
    cur, cnt, n = None, collections.Counter(), 0
    for l in txt[start + 1:]:
        if l.startswith("//-----") or l.startswith("\t.section"):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
            cnt[cur] += 1
            n += 1
    print("instructions", n)
    byfile = collections.Counter()
    for k, v in cnt.items():
        byfile[k[0] if k else None] += v
    print(byfile.most_common())
    for k, v in cnt.most_common(int(top)):
        print(v, k)


if __name__ == "__main__":
    main(*sys.argv[1:])
