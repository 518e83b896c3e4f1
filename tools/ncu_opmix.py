#!/usr/bin/env python3
"""Opcode mix of one kernel from an ncu `--page source --csv` dump: warp instructions per SASS opcode,
average active threads, and the share of each execution pipe (rough opcode -> pipe map).

  ncu -i rep.ncu-rep --page source --csv > src.csv ; python tools/ncu_opmix.py src.csv
"""
import csv
import re
import sys
from collections import defaultdict

PIPE = {
    "alu": "LOP3 SHF IADD3 IADD ISETP SEL PRMT FSETP FSEL FMNMX IMNMX LEA VOTE BMSK SGXT P2R R2P PLOP3 FSET VIADD VIMNMX MOV CS2R".split(),
    "fma": "IMAD FFMA FMUL FADD IMUL HFMA2 FADD2 FMUL2 FFMA2".split(),
    "xu": "POPC FLO BREV MUFU I2F F2I I2FP F2FP F2F I2I".split(),
    "lsu": "LDS STS LDG STG LD ST ATOM ATOMG ATOMS RED LDSM LDC SHFL MATCH REDUX REDG".split(),
    "cbu": "BRA BSSY BSYNC EXIT BAR WARPSYNC CALL RET BREAK NANOSLEEP YIELD".split(),
    "uni": "ULDC UMOV UIADD3 ULOP3 USHF UISETP USEL UIMAD ULEA UPRMT UFLO UPOPC R2UR S2UR UTMALDG SYNCS UBMSK UPLOP3 UP2UR UR2UP UMOV".split(),
}
OP2PIPE = {o: p for p, os in PIPE.items() for o in os}


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    h = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
    hdr = rows[h]
    si, ii, ti = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    ops = defaultdict(lambda: [0, 0])
    for r in rows[h + 1:]:
        if len(r) <= ti:
            continue
        s = r[si].strip()
        s = re.sub(r"^@!?U?P\d+\s+", "", s)
        op = s.split()[0].split(".")[0] if s else "?"
        ops[op][0] += int(r[ii])
        ops[op][1] += int(r[ti])
    tot = sum(v[0] for v in ops.values())
    pipes = defaultdict(int)
    for op, (n, t) in ops.items():
        pipes[OP2PIPE.get(op, "other")] += n
    print(f"warp instructions {tot}")
    print("pipes: " + ", ".join(f"{p}={100 * n / tot:.1f}%" for p, n in sorted(pipes.items(), key=lambda kv: -kv[1])))
    for op, (n, t) in sorted(ops.items(), key=lambda kv: -kv[1][0])[:45]:
        print(f"{100 * n / tot:5.1f}%  {n:>11d}  thr/inst {t / max(n, 1):5.1f}  {op}  [{OP2PIPE.get(op, 'other')}]")


if __name__ == "__main__":
    main()
