#!/usr/bin/env python3
"""Summarise an .ncu-rep here (no GPU needed): per-launch key metrics, and for one launch the warp
instructions / stall samples per CUDA source line.

  python tools/ncu_report.py rep.ncu-rep                       # launch table
  python tools/ncu_report.py rep.ncu-rep LAUNCH_ID FUNC [N]    # per-line view of launch LAUNCH_ID
                                                               # (FUNC = substring of the mangled kernel name)
The per-line view joins `ncu --page source --csv` (SASS rows) with `nvdisasm -g` line markers of the
cubin inside primitive3d_b200/libprim3d_b200.so, which must be the build that was profiled.
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct"]
STALLS = ["stall_barrier", "stall_long_sb", "stall_short_sb", "stall_mio", "stall_lg", "stall_wait", "stall_math",
          "stall_not_selected", "stall_selected", "stall_branch_resolving", "stall_no_inst", "stall_dispatch",
          "stall_membar", "stall_sleeping", "stall_drain", "stall_tex", "stall_misc"]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def launch_table(rep):
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    idx = [(k, hdr.index(k)) for k in KEYS if k in hdr]
    name = hdr.index("Kernel Name")
    print("id  kernel" + " " * 26 + "  ".join(k.split(".")[0][-18:] for k, _ in idx))
    for n, r in enumerate(rows[2:]):
        print(f"{n:<3d} {r[name][:30]:30s} " + "  ".join(f"{r[i][:12]:>12s}" for _, i in idx))
    print("units: " + ", ".join(f"{k}={units[i]}" for k, i in idx))


def sass_lines(func, nrows=None):
    """(file, line) per SASS instruction of the function whose section name contains `func`.  A template has one
    section per instance: the one with exactly `nrows` instructions (the launch that was profiled) is taken."""
    so = os.path.join(ROOT, "primitive3d_b200", "libprim3d_b200.so")
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=d, capture_output=True)
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
                sections[-1].append(cur)
    sections = [sec for sec in sections if sec]
    exact = [sec for sec in sections if nrows is not None and len(sec) == nrows]
    return exact[0] if exact else (sections[0] if sections else [])


def per_line(rep, launch, func, top):
    text = ncu(["-i", rep, "--page", "source", "--csv", "--launch-skip", str(launch), "--launch-count", "1"])
    rows = list(csv.reader(io.StringIO(text)))
    h = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
    hdr = rows[h]
    ii, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
    st = [(s, hdr.index(s)) for s in STALLS if s in hdr]
    body = []
    for r in rows[h + 1:]:
        if r and r[0] == "Kernel Name":
            break
        if len(r) > ii:
            body.append(r)
    lines = sass_lines(func, len(body))
    print(f"{rows[0][1][:70]}: {len(body)} SASS rows, {len(lines)} in the local cubin")
    agg, stall_tot = {}, [0] * len(st)
    for r, ln in zip(body, lines):
        a = agg.setdefault(ln, [0, 0, [0] * len(st)])
        a[0] += int(r[ii])
        a[1] += int(r[si] or 0)
        for k, (_, i) in enumerate(st):
            v = int(r[i] or 0)
            a[2][k] += v
            stall_tot[k] += v
    tot = sum(a[0] for a in agg.values())
    smp = sum(a[1] for a in agg.values()) or 1
    print(f"warp instructions {tot}, samples {smp}")
    print("stall samples: " + ", ".join(f"{s[6:]}={100 * v / smp:.1f}%" for (s, _), v in sorted(zip(st, stall_tot), key=lambda t: -t[1]) if v * 100 > smp))
    src = {}
    for f in sorted(os.listdir(os.path.join(ROOT, "primitive3d_b200", "csrc"))):
        p = os.path.join(ROOT, "primitive3d_b200", "csrc", f)
        if os.path.exists(p):
            src[f] = {i + 1: s.rstrip() for i, s in enumerate(open(p))}
    for ln, (n, s, sv) in sorted(agg.items(), key=lambda kv: (kv[0] or ("", 0))):
        if 1000 * n > top * tot or 1000 * s > top * smp:
            t = src.get(ln[0], {}).get(ln[1], "") if ln else ""
            best = max(range(len(st)), key=lambda k: sv[k]) if s else 0
            why = f"{st[best][0][6:]}" if s else ""
            print(f"{100 * n / tot:5.1f}% inst {100 * s / smp:5.1f}% smp {why:12s} {ln[0][:14] if ln else '?':14s}:{ln[1] if ln else 0:<4d} {t.strip()[:95]}")


if __name__ == "__main__":
    if len(sys.argv) == 2:
        launch_table(sys.argv[1])
    else:
        per_line(sys.argv[1], int(sys.argv[2]), sys.argv[3], float(sys.argv[4]) if len(sys.argv) > 4 else 5.0)
