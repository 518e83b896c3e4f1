#!/usr/bin/env python3
"""Aggregate an ncu `--page source --csv` SASS dump by CUDA source line.

  ncu -i rep.ncu-rep --page source --csv --kernel-name k_emit > sass.csv
  cuobjdump -xelf all lib.so; nvdisasm -g -c mc_kernels.sm_100a.cubin > mc.sass
  python tools/ncu_by_line.py sass.csv mc.sass k_emit [source.cu]

nvdisasm -g prints a `//## File "...", line N` marker before the SASS it belongs to; the n-th
instruction of the function in that listing is the n-th row of the ncu table.
"""
import csv
import re
import sys


def sass_lines(path, func):
    out, cur, on = [], None, False
    for ln in open(path, errors="replace"):
        if ln.startswith(".text.") or re.match(r"\s*\.section\s+\.text\.", ln):
            on = func in ln
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4}\*/", ln):
            out.append(cur)
    return out


def main():
    table, sass, func = sys.argv[1:4]
    src = sys.argv[4] if len(sys.argv) > 4 else None
    rows = list(csv.reader(open(table)))
    h = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
    hdr = rows[h]
    ii, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
    body = [r for r in rows[h + 1:] if len(r) > ii]
    lines = sass_lines(sass, func)
    if len(lines) != len(body):
        print(f"warning: {len(lines)} SASS instructions in listing vs {len(body)} rows", file=sys.stderr)
    agg = {}
    for r, ln in zip(body, lines):
        a = agg.setdefault(ln, [0, 0])
        a[0] += int(r[ii])
        a[1] += int(r[si] or 0)
    tot = sum(a[0] for a in agg.values())
    smp = sum(a[1] for a in agg.values()) or 1
    text = {}
    if src:
        text = {i + 1: s.rstrip() for i, s in enumerate(open(src))}
    print(f"total warp instructions {tot}")
    for ln, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
        t = text.get(ln[1], "") if ln and src and ln[0] in src else ""
        print(f"{100 * n / tot:5.1f}% inst {100 * s / smp:5.1f}% samples  {ln}  {t.strip()[:100]}")


if __name__ == "__main__":
    main()
