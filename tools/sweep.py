#!/usr/bin/env python
"""Developer tool: time the phases of one plan under several engine option settings (data generated once).

    python tools/sweep.py --plan Csr --log2-build 27 --log2-probe 30 --set 2=0 --set 3=4194304 --set 3=8388608
Each --set is one run with that HJ3D_OPT id=value applied on top of the defaults (comma separated lists allowed).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--plan", default="Csr")
    ap.add_argument("--log2-build", type=int, default=27)
    ap.add_argument("--log2-probe", type=int, default=30)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--set", action="append", default=[""])
    ap.add_argument("--no-write", action="store_true")
    ap.add_argument("--no-checksum", action="store_true")
    a = ap.parse_args()
    import torch
    import hj3d_loader
    pkg = hj3d_loader.load()
    from bench import PLANS
    dev = torch.device("cuda", 0)
    kind_name, build_rel, mode = PLANS[a.plan]
    nR, nS = 1 << a.log2_build, 1 << a.log2_probe
    g = torch.Generator(device=dev); g.manual_seed(1234)
    R = torch.zeros((nR, 3), dtype=torch.int32, device=dev)
    R[:, 0] = torch.randperm(nR, device=dev, generator=g, dtype=torch.int64).to(torch.int32)
    S = torch.zeros((nS, 3), dtype=torch.int32, device=dev)
    S[:, 0] = torch.arange(nS, device=dev, dtype=torch.int64).to(torch.int32)
    S[:, 1] = torch.randint(0, nR, (nS,), device=dev, generator=g, dtype=torch.int64).to(torch.int32)
    ksR, ksS = pkg.KeySpec(12, 0), pkg.KeySpec(12, 4)
    B, ksB, nB, P, ksP, nP = (R, ksR, nR, S, ksS, nS) if build_rel == "R" else (S, ksS, nS, R, ksR, nR)
    D = nR if build_rel == "R" else max(nR - int(nR * (1 - 1 / nR) ** nS), 1)
    out = None if a.no_write else torch.empty((nS, 2), dtype=torch.int32, device=dev)
    nest = torch.empty((nP, 2), dtype=torch.int32, device=dev) if mode == 3 else None
    flags = 0 if a.no_checksum else pkg.F_CHECKSUM
    for setting in a.set:
        ctx = pkg.Context(0, stream=torch.cuda.current_stream().cuda_stream)
        for kv in [x for x in setting.split(",") if x]:
            k, v = kv.split("=")
            ctx.set_option(int(k), int(v))
        table = ctx.table(pkg.CHAINING if kind_name == "chaining" else pkg.NESTED, D)
        rows = []
        for rep in range(a.reps):
            table.clear()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            table.build(B, nB, ksB)
            tb = ctx.timings()
            if mode <= 1:
                rc, c = table.probe_chaining(P, nP, ksP, unique=(mode == 1), flags=flags, out=out, out_cap=nS if out is not None else 0)
                tp = ctx.timings(); tu = {"unnest_ms": 0.0}; res = c
            else:
                rc, c, res = table.probe_nested_unnest(P, nP, ksP, flags=flags, out=out, out_cap=nS if out is not None else 0)
                tp = ctx.timings(); tu = {"unnest_ms": 0.0}
            e1.record(); torch.cuda.synchronize()
            rows.append({"total": e0.elapsed_time(e1), "b_part": tb["partition_ms"], "hist": tb["histogram_ms"], "scan": tb["scan_ms"],
                         "scatter": tb["scatter_ms"], "group": tb["group_ms"], "build": tb["total_ms"],
                         "p_part": tp["partition_ms"], "probe": tp["probe_ms"], "probe_call": tp["total_ms"],
                         "unnest": tu["unnest_ms"], "out": res["out_tuples"], "cmps": c["num_cmps"]})
        best = min(rows, key=lambda r: r["total"])
        print(json.dumps({"set": setting, "plan": a.plan, **{k: (round(v, 3) if isinstance(v, float) else v) for k, v in best.items()},
                          "Gtuples_s": round((nR + nS) / best["total"] / 1e6, 2)}), flush=True)
        table.destroy(); ctx.close()


if __name__ == "__main__":
    main()
