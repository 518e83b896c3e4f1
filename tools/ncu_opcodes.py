#!/usr/bin/env python3
"""Opcode histogram of one kernel from `ncu -i rep --page source --csv`, weighted by executed warp instructions.

  python tools/ncu_opcodes.py source.csv
Also prints thread-level efficiency (avg active threads) per opcode.
"""
import csv
import re
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[h]
si, ii, ti, pi = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("Predicated-On Thread Instructions Executed")
agg = defaultdict(lambda: [0, 0, 0])
tot = 0
for r in rows[h + 1:]:
    if len(r) <= pi or not r[ii].isdigit():
        continue
    s = r[si].strip()
    s = re.sub(r"^@!?U?P\d+\s+", "", s)
    op = s.split()[0].split(".")[0] if s else "?"
    full = ".".join(s.split()[0].split(".")[:2]) if op in ("LDS", "STS", "LDG", "STG", "SHFL", "BAR") else op
    a = agg[full]
    a[0] += int(r[ii]); a[1] += int(r[ti]); a[2] += int(r[pi])
    tot += int(r[ii])
print(f"total warp instructions {tot}")
for op, (n, t, p) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:45]:
    print(f"{op:14s} {100*n/tot:5.1f}%  avg threads {t/max(n,1):5.1f}  pred-on {p/max(n,1):5.1f}")
allt = sum(a[1] for a in agg.values()); allp = sum(a[2] for a in agg.values())
print(f"overall avg active threads {allt/tot:.1f}, predicated-on {allp/tot:.1f}")
