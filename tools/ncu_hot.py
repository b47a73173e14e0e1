#!/usr/bin/env python
"""Developer tool: hottest SASS instructions of an `ncu --page source --csv` export.

    python tools/ncu_hot.py gpurun_out/prof_source.csv [top N] [context lines]
"""
import csv
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    ctx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[start]
    body = [r for r in rows[start + 1:] if len(r) == len(hdr)]
    ix = {h: i for i, h in enumerate(hdr)}
    samp = [int(r[ix["# Samples"]] or 0) for r in body]
    execd = [int(r[ix["Instructions Executed"]] or 0) for r in body]
    tot = sum(samp) or 1
    print(f"{len(body)} SASS instructions, {tot} samples, {sum(execd)} warp instructions executed")
    stall_cols = [h for h in hdr if h.startswith("stall_")]
    order = sorted(range(len(body)), key=lambda i: -samp[i])[:top]
    for i in sorted(order):
        lo, hi = max(0, i - ctx), min(len(body), i + ctx + 1)
        for j in range(lo, hi):
            r = body[j]
            st = sorted(((int(r[ix[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:2] if stall_cols else []
            mark = ">>" if j == i else "  "
            print(f"{mark} {j:5d} {100 * samp[j] / tot:5.1f}% exec={execd[j]:>10d} {r[ix['Source']].strip()[:90]:<90} {st}")
        if ctx:
            print()


if __name__ == "__main__":
    main()
