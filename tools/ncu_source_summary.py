#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` export (SASS view): per-opcode sample and instruction shares, the top stalled
instructions, and the stall-reason mix.  python tools/ncu_source_summary.py <source.csv> [top_n]"""
import csv
import re
import sys
from collections import Counter, defaultdict


def load(path):
    rows = list(csv.reader(open(path)))
    kernels, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            kernels.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None and r:
            cur["rows"].append(r)
    return kernels


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    for k in load(path):
        hdr = k["hdr"]
        ci = {h: i for i, h in enumerate(hdr)}
        S, I = ci["# Samples"], ci["Instructions Executed"]
        stall_cols = [(h, i) for h, i in ci.items() if h.startswith("stall_") and "Not Issued" not in h]
        tot_s = sum(int(r[S]) for r in k["rows"])
        tot_i = sum(int(r[I]) for r in k["rows"])
        print(f"## {re.sub(r'rmg::|<unnamed>::|void ', '', k['name'])[:110]}")
        print(f"instructions (SASS lines) {len(k['rows'])}, warp-instructions executed {tot_i:,}, samples {tot_s:,}")
        byop_s, byop_i = Counter(), Counter()
        for r in k["rows"]:
            op = r[ci["Source"]].split()
            op = [x for x in op if not x.startswith("@")][0].split(".")[0] if op else "?"
            byop_s[op] += int(r[S])
            byop_i[op] += int(r[I])
        print("\n| opcode | % samples | % warp-instructions |\n|---|---|---|")
        for op, s in byop_s.most_common(14):
            print(f"| {op} | {100 * s / max(tot_s, 1):.1f} | {100 * byop_i[op] / max(tot_i, 1):.1f} |")
        reasons = Counter()
        for r in k["rows"]:
            for h, i in stall_cols:
                reasons[h] += int(r[i])
        print("\n| stall reason | % samples |\n|---|---|")
        for h, s in reasons.most_common(8):
            print(f"| {h} | {100 * s / max(tot_s, 1):.1f} |")
        print(f"\ntop {top} instructions by samples (line index, samples %, executed, SASS, main stall):")
        order = sorted(range(len(k["rows"])), key=lambda j: -int(k["rows"][j][S]))[:top]
        for j in order:
            r = k["rows"][j]
            main_stall = max(stall_cols, key=lambda hi: int(r[hi[1]]))[0]
            print(f"  {j:5d} {100 * int(r[S]) / max(tot_s, 1):5.1f}% {int(r[I]):>10,}  {r[ci['Source']].strip()[:70]:70s} {main_stall}")
        print()


if __name__ == "__main__":
    main()
