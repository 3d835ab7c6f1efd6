#!/usr/bin/env python
"""Opcode histogram per kernel from `cuobjdump -sass` (analysis tooling): python scripts/sass_hist.py obj [filter]"""
import collections
import re
import subprocess
import sys

obj = sys.argv[1]
flt = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
cur = None
hist = collections.defaultdict(collections.Counter)
regs = {}
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and cur:
        hist[cur][m.group(2)] += 1
for fn, h in hist.items():
    if flt and flt not in fn:
        continue
    d = subprocess.run(["cu++filt", fn], capture_output=True, text=True).stdout.strip()[:110]
    print(f"== {d}  ({sum(h.values())} instrs)")
    print("   " + "  ".join(f"{k}:{v}" for k, v in h.most_common(28)))
