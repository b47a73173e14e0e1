#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the committed summaries under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches_r1.csv profiles/r1_launches_summary.md
    python tools/ncu_summary.py full gpurun_out/prof_r1b.ncu-rep profiles/r1_ncu_full_medium.md
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors.sum", "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_atom.sum",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers"]


def short(name):
    name = re.sub(r"\(.*", "", name.replace("void ", "").replace("hj3d::", ""))
    return re.sub(r"<unnamed>::|\(anonymous namespace\)::", "", name)[:70]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    tot = 0.0
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        ms = v / 1e6 if r[ui].startswith("ns") else (v / 1e3 if r[ui].startswith("us") else v)
        a = agg.setdefault(short(r[ki]), [0, 0.0])
        a[0] += 1; a[1] += ms; tot += ms
    ours = {k: v for k, v in agg.items() if k.startswith("k_")}
    tot_ours = sum(v[1] for v in ours.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list summary ({src})\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised launches: compare SHARES).\n\n")
        f.write(f"{len(rows) - 1} launches, {tot:.3f} ms total, {tot_ours:.3f} ms in hj3d kernels.\n\n")
        f.write("| kernel | launches | total ms | ms / launch | share of hj3d kernels |\n|---|---:|---:|---:|---:|\n")
        for k, (c, ms) in sorted(ours.items(), key=lambda x: -x[1][1]):
            f.write(f"| `{k}` | {c} | {ms:.3f} | {ms / c:.3f} | {100 * ms / tot_ours:.1f}% |\n")
        other = tot - tot_ours
        f.write(f"\nOther (torch data generation / verification kernels): {other:.3f} ms.\n")
    print(open(dst).read())


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    st = [h for h in hdr if "issue_stalled" in h and "pcsamp" in h and "not_issued" not in h]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary ({src})\n\n")
        for r in rows[2:]:
            f.write(f"## `{short(r[idx['Kernel Name']])}`\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in idx:
                    f.write(f"| {k} | {r[idx[k]]} | {units[idx[k]]} |\n")
            tot = sum(float(r[idx[h]] or 0) for h in st) or 1.0
            top = sorted(((float(r[idx[h]] or 0), h) for h in st), reverse=True)[:6]
            f.write("\nwarp stall samples: " + ", ".join(
                f"{h.replace('smsp__pcsamp_warps_issue_stalled_', '')} {100 * v / tot:.0f}%" for v, h in top) + "\n\n")
    print(open(dst).read()[:3000])


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
