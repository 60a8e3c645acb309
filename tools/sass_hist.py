#!/usr/bin/env python
"""Static SASS opcode histogram of one kernel of a cubin / shared library (no GPU needed):
    python tools/sass_hist.py claymore_b200/lib/libclaymore_b200.so 'g2p2g_kernelILi1E' [--top 30]
Counts are static (one per instruction in the binary), a proxy for straight-line regions; use tools/ncu_lines.py for executed counts."""
import collections
import re
import subprocess
import sys


def main():
    path, pat = sys.argv[1], sys.argv[2]
    top = int(sys.argv[4]) if len(sys.argv) > 4 and sys.argv[3] == "--top" else 30
    out = subprocess.run(["cuobjdump", "-sass", path], stdout=subprocess.PIPE, text=True).stdout
    cur, hist, total = None, collections.Counter(), 0
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur is None or pat not in cur:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            hist[m.group(2)] += 1
            total += 1
    print(f"{pat}: {total} SASS instructions")
    for op, n in hist.most_common(top):
        print(f"  {op:12s} {n:6d}  {100 * n / max(total, 1):5.1f}%")


if __name__ == "__main__":
    main()
