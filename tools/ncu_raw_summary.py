#!/usr/bin/env python
"""Turn an `ncu --page raw --csv` export into the markdown summary committed under profiles/.

    python tools/ncu_raw_summary.py gpurun_out/prof_raw.csv profiles/r1c_ncu_full_config2.md "title"
"""
import csv
import re
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sectors.sum", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed_op_shared_atom.sum"]


def short(name):
    name = re.sub(r"\(.*", "", name.replace("void ", "").replace("hj3d::", ""))
    return re.sub(r"\(int\)|\(bool\)", "", name)[:80]


def main():
    src, dst, title = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
    rows = list(csv.reader(open(src)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    st = [h for h in hdr if "issue_stalled" in h and "pcsamp" in h and "not_issued" not in h]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary: {title}\n\nsource: `{src}` (`ncu --set full --clock-control none --import-source on`)\n\n")
        for r in rows[2:]:
            f.write(f"## `{short(r[ix['Kernel Name']])}`\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in ix:
                    f.write(f"| {k} | {r[ix[k]]} | {units[ix[k]]} |\n")
            try:
                rd, wr = float(r[ix["dram__bytes_read.sum"]]), float(r[ix["dram__bytes_write.sum"]])
                u = units[ix["dram__bytes_read.sum"]]
                ms = float(r[ix["gpu__time_duration.sum"]])
                tu = units[ix["gpu__time_duration.sum"]]
                if u == "Gbyte" and tu in ("ms", "msecond"):
                    f.write(f"| DRAM traffic (read + write) | {rd + wr:.3f} | Gbyte |\n| DRAM GB/s over the launch | {(rd + wr) / ms * 1e3:.0f} | GB/s |\n")
            except Exception:
                pass
            vals = sorted(((float(r[ix[h]].replace(",", "") or 0), h.split("issue_stalled_")[1].split("_per")[0].replace(".pct", "")) for h in st), reverse=True)
            tot = sum(v for v, _ in vals) or 1.0
            f.write("\nwarp stall samples: " + ", ".join(f"{n} {100 * v / tot:.0f}%" for v, n in vals[:6]) + "\n\n")
    print(open(dst).read())


if __name__ == "__main__":
    main()
