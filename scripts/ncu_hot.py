#!/usr/bin/env python
"""Hot SASS instructions of an ncu `--page source --csv` dump: python scripts/ncu_hot.py file.csv [topN]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
i_src, i_s = hdr.index('Source'), hdr.index('# Samples')
i_ex = hdr.index('Instructions Executed')
stall = [i for i, h in enumerate(hdr) if h.startswith('stall_')]
data = rows[2:]
tot = sum(int(r[i_s] or 0) for r in data)
print("total samples", tot, "instructions", len(data))
order = sorted(range(len(data)), key=lambda k: -int(data[k][i_s] or 0))[:top]
for k in sorted(order):
    r = data[k]
    st = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall), reverse=True)[:2]
    print(f"{k:5d} {100*int(r[i_s])/tot:5.1f}% ex={r[i_ex]:>8} {r[i_src].strip()[:70]:70s} {st}")
