#!/usr/bin/env python3
"""Per-source-line view of one kernel of an .ncu-rep, here (no GPU): warp instructions (optionally per unit of work),
average active threads and stall samples, every line, in source order.

  python tools/ncu_lines.py rep.ncu-rep MANGLED_SUBSTR [--cubin NAME] [--per N] [--min X] [--src FILE]

Joins `ncu --page source --csv` (one row per SASS instruction) with the `nvdisasm -g` line markers of the cubin
inside primitive3d_b200/libprim3d_b200.so, which must be the build that was profiled.  --per N divides the counts
by N (tiles, rows ...), --min hides lines below X instructions per unit.
"""
import argparse
import csv
import io
import os
import re
import subprocess
import tempfile
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sass_lines(func, lib, nrows=None):
    """((file, line), SASS text) per instruction of the function whose section name contains `func`.  A template has
    one section per instance: the one with exactly `nrows` instructions (the launch that was profiled) is taken."""
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, capture_output=True)
    sections = []
    for f in sorted(os.listdir(d)):
        if not f.endswith(".cubin"):
            continue
        text = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, f)], capture_output=True, text=True).stdout
        if func not in text:
            continue
        cur, on = None, False
        for ln in text.splitlines():
            if re.match(r"\s*\.section\s+\.text\.", ln) or ln.startswith(".text."):
                on = func in ln
                if on:
                    sections.append([])
                continue
            if not on:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur = (m.group(1).split("/")[-1], int(m.group(2)))
            elif re.match(r"\s+/\*[0-9a-f]{4,5}\*/", ln):
                sections[-1].append((cur, ln.strip()))
    sections = [sec for sec in sections if sec]
    exact = [sec for sec in sections if nrows is not None and len(sec) == nrows]
    return exact[0] if exact else (sections[0] if sections else [])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("func")
    ap.add_argument("--per", type=float, default=1.0)
    ap.add_argument("--min", type=float, default=0.0)
    ap.add_argument("--lib", default=os.path.join(ROOT, "primitive3d_b200", "libprim3d_b200.so"))
    ap.add_argument("--sass", action="store_true", help="one line per SASS instruction instead of per source line")
    ap.add_argument("--launch", type=int, default=0)
    args = ap.parse_args()
    text = subprocess.run(["ncu", "-i", args.rep, "--page", "source", "--csv", "--launch-skip", str(args.launch),
                           "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(text)))
    h = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
    hdr = rows[h]
    ii, ti, si = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    body = [r for r in rows[h + 1:] if len(r) > ti and r[0] != "Kernel Name"]
    lines = sass_lines(args.func, args.lib, len(body))
    print(f"# {len(body)} SASS rows in the report, {len(lines)} in the local cubin")
    src = {}
    for f in os.listdir(os.path.join(ROOT, "primitive3d_b200", "csrc")):
        src[f] = {i + 1: s.rstrip() for i, s in enumerate(open(os.path.join(ROOT, "primitive3d_b200", "csrc", f), errors="replace"))}
    tot = sum(int(r[ii]) for r in body)
    smp = sum(int(r[si] or 0) for r in body) or 1
    print(f"# warp instructions {tot} ({tot / args.per:.1f} per unit), samples {smp}")
    if args.sass:
        for (cur, s), r in zip(lines, body):
            n = int(r[ii])
            if n / args.per >= args.min:
                print(f"{n / args.per:9.2f} thr {int(r[ti]) / max(n, 1):5.1f} smp {int(r[si] or 0):5d} {cur[0][:16] if cur else '?':16s}:{cur[1] if cur else 0:<4d} {s[:100]}")
        return
    agg = defaultdict(lambda: [0, 0, 0])
    for (cur, _), r in zip(lines, body):
        a = agg[cur]
        a[0] += int(r[ii])
        a[1] += int(r[ti])
        a[2] += int(r[si] or 0)
    for cur, (n, t, s) in sorted(agg.items(), key=lambda kv: kv[0] or ("", 0)):
        if n / args.per >= args.min:
            code = src.get(cur[0], {}).get(cur[1], "") if cur else ""
            print(f"{n / args.per:9.2f} thr {t / max(n, 1):5.1f} smp {100 * s / smp:5.1f}% {cur[0][:16] if cur else '?':16s}:{cur[1] if cur else 0:<4d} {code.strip()[:100]}")


if __name__ == "__main__":
    main()
