#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` dump by SASS opcode: shared-memory wavefronts
(actual / ideal / excess), warp instructions and stall samples.

    ncu -i capture.ncu-rep --page source --csv > source.csv
    python tools/ncu_opcode_table.py source.csv [--stalls]
"""
import csv
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    col = {h: i for i, h in enumerate(hdr)}
    agg = defaultdict(lambda: [0, 0, 0, 0, 0])
    tot_samples = 0
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        src = r[col["Source"]].strip()
        parts = src.split()
        op = parts[1] if parts and parts[0].startswith("@") else (parts[0] if parts else "?")
        a = agg[op]
        a[0] += int(r[col["L1 Wavefronts Shared"]] or 0)
        a[1] += int(r[col["L1 Wavefronts Shared Ideal"]] or 0)
        a[2] += int(r[col["L1 Wavefronts Shared Excessive"]] or 0)
        a[3] += int(r[col["Instructions Executed"]] or 0)
        a[4] += int(r[col["# Samples"]] or 0)
        tot_samples += int(r[col["# Samples"]] or 0)
    tot = sum(a[0] for a in agg.values())
    tot_inst = sum(a[3] for a in agg.values())
    print(f"{'opcode':28s}{'wavefronts':>14s}{'share':>8s}{'excess':>14s}{'ideal':>14s}{'warp instr':>14s}{'wf/instr':>9s}{'samples%':>9s}")
    for op, a in sorted(agg.items(), key=lambda kv: (-kv[1][0], -kv[1][3])):
        if a[0] == 0 and a[3] < tot_inst * 0.004:
            continue
        print(f"{op:28s}{a[0]:14d}{100.0 * a[0] / max(1, tot):7.1f}%{a[2]:14d}{a[1]:14d}{a[3]:14d}"
              f"{a[0] / max(1, a[3]):9.2f}{100.0 * a[4] / max(1, tot_samples):8.1f}%")
    print(f"{'total':28s}{tot:14d}{'':8s}{'':14s}{'':14s}{tot_inst:14d}")


if __name__ == "__main__":
    main()
