#!/usr/bin/env python3
"""Local side of the evidence under profiles/ (no GPU needed): reads the .ncu-rep files tools/collect_profiles.sh
left in gpurun_out/ and writes, per captured kernel,

  profiles/TAG_ncu_full_<kernel>[_<workload>].txt   key metrics, then warp instructions and stall samples per source line
  profiles/TAG_sass_opcodes_<kernel>[_<workload>].txt   executed-opcode histogram (UTMALDG / SYNCS prove TMA + mbarrier)
  profiles/ncu_traffic.json                          DRAM bytes per launch, keyed by kernel and slab planes (bench.py reads it)

  python tools/summarise_profiles.py TAG
"""
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"

# report file -> (workload label, planes for ncu_traffic.json or None, [(launch index, kernel substring, short name)])
REPORTS = {
    f"{tag}_mc1024.ncu-rep": ("gyroid 1024^3", 1024, [(0, "k_tile", "k_tile"), (1, "k_faces_rows", "k_faces_rows")]),
    f"{tag}_mc2048slab.ncu-rep": ("gyroid 2048^3, planes [0,257) (the shard of one of 8 GPUs)", 257,
                                  [(0, "k_tile", "k_tile"), (1, "k_faces", "k_faces")]),
    f"{tag}_small.ncu-rep": ("bunny 66^3, then a batch of 64 of them", None, [(0, "k_small", "k_small"), (1, "k_small", "k_small_batch64")]),
    f"{tag}_mtx.ncu-rep": ("marching tetrahedra, Kuhn 128^3, second call", None,
                           [(0, "k_mtx_pack", "k_mtx_pack"), (1, "k_mtx_classify", "k_mtx_classify"), (2, "k_mtx_sort", "k_mtx_sort"),
                            (3, "k_mtx_emit", "k_mtx_emit"), (4, "k_mtx_faces", "k_mtx_faces")]),
}


def run(args):
    return subprocess.run(args, capture_output=True, text=True).stdout


def raw_rows(rep):
    rows = list(csv.reader(io.StringIO(run(["ncu", "-i", rep, "--page", "raw", "--csv"]))))
    return rows[0], rows[2:]


traffic = {}
for fname, (workload, planes, kernels) in REPORTS.items():
    rep = os.path.join(OUT, fname)
    if not os.path.exists(rep):
        print("missing", fname)
        continue
    hdr, rows = raw_rows(rep)
    col = {k: hdr.index(k) for k in ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                     "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum") if k in hdr}
    units = list(csv.reader(io.StringIO(run(["ncu", "-i", rep, "--page", "raw", "--csv"]))))[1]
    suffix = "" if planes in (None, 1024) else f"_slab{planes}"
    for idx, sub, short in kernels:
        if idx >= len(rows) or sub not in rows[idx][col["Kernel Name"]]:
            print("unexpected launch order in", fname, idx, rows[idx][col["Kernel Name"]] if idx < len(rows) else None)
            continue
        table = run([sys.executable, os.path.join(ROOT, "tools", "ncu_report.py"), rep])
        lines = run([sys.executable, os.path.join(ROOT, "tools", "ncu_report.py"), rep, str(idx), sub, "8"])
        details = run(["ncu", "-i", rep, "--page", "details", "--launch-skip", str(idx), "--launch-count", "1"])
        keep = [ln for ln in details.splitlines() if any(k in ln for k in (
            "Duration", "Throughput", "Hit Rate", "Registers Per", "Shared Memory Config", "Dynamic Shared", "Static Shared", "Grid Size",
            "Block Size", "Achieved Occupancy", "Theoretical Occupancy", "Executed Ipc", "Issue Slots Busy", "Eligible Warps",
            "Mem Busy", "Max Bandwidth", "Bank Conflicts", "Waves Per SM"))]
        with open(os.path.join(PROF, f"{tag}_ncu_full_{short}{suffix}.txt"), "w") as f:
            f.write(f"# ncu --set full --import-source on --clock-control none, one launch of {sub}; workload: {workload}\n")
            f.write(f"# report: gpurun_out/{fname} (scratch), launch {idx}; summarised by tools/summarise_profiles.py\n\n")
            f.write("## launch table of the report\n" + table + "\n## this launch, details page (selected rows)\n" + "\n".join(keep))
            f.write("\n\n## warp instructions and stall samples per source line (lines with >= 0.8 % of either)\n" + lines)
        src_csv = os.path.join(OUT, f"{tag}_{short}{suffix}_source.csv")
        with open(src_csv, "w") as f:
            f.write(run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(idx), "--launch-count", "1"]))
        ops = run([sys.executable, os.path.join(ROOT, "tools", "ncu_opcodes.py"), src_csv])
        with open(os.path.join(PROF, f"{tag}_sass_opcodes_{short}{suffix}.txt"), "w") as f:
            f.write(f"# executed SASS opcodes of {sub} (ncu --page source, weighted by executed warp instructions); workload: {workload}\n")
            f.write(ops)
        if planes is not None:
            r = rows[idx]

            def val(k):
                i = col[k]
                v = float(r[i].replace(",", ""))
                u = units[i]
                return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1, "us": 1e-3, "ns": 1e-6}.get(u, 1)
            traffic.setdefault(short, []).append({
                "planes": planes, "workload": workload, "dram_bytes_read": val("dram__bytes_read.sum"),
                "dram_bytes_write": val("dram__bytes_write.sum"), "duration_ms_under_ncu": val("gpu__time_duration.sum"),
                "issue_active_pct": float(r[col["smsp__issue_active.avg.pct_of_peak_sustained_active"]]),
                "warp_instructions": float(r[col["smsp__inst_executed.sum"]].replace(",", "")),
                "source": f"profiles/{tag}_ncu_full_{short}{suffix}.txt (ncu --set full --clock-control none, one launch)"})
if traffic:
    json.dump(traffic, open(os.path.join(PROF, "ncu_traffic.json"), "w"), indent=1)
for f in (f"{tag}_launches_gyroid1024.csv", f"{tag}_launches_tets.csv", f"{tag}_kernel_times.json", f"{tag}_small_times.json"):
    if os.path.exists(os.path.join(OUT, f)):
        shutil.copy(os.path.join(OUT, f), os.path.join(PROF, f))
print(sorted(f for f in os.listdir(PROF) if f.startswith(tag) or f == "ncu_traffic.json"))
