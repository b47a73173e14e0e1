#!/usr/bin/env python
"""bench.py -- headline benchmark: join input tuples/s (build + probe) of the key/foreign-key join of
main_experiment1 at the reference's largest shape (-R 27 -S 30: 2^27 build / 2^30 probe, uint32 keys,
12-byte {k,a,b} row-store tuples, b=1 => 2^27 buckets), plan Csr (chaining, IsBuildKeyUnique) with
materialised (probe row, build row) result pairs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--plan Csr|CsrUU|Nsr|Crs|Nrs]
                    [--log2-build 27 --log2-probe 30] [--zipf S] [--config 3|4|5]

One "step" = clear the table, build strand, probe strand (+ unnest for nested plans) over one batch of synthetic
input resident in HBM.  The relations are generated on the device by the engine's counter-based generators
(csrc/datagen.cu): every value is a pure function of the global row id, so a run on N GPUs joins EXACTLY the data a run
on one GPU joins (rank r holds rows [r*n/N, (r+1)*n/N) of both relations).

N > 1 (torchrun): the same total workload is sharded by bucket range (strong scaling).  Every rank partitions its
slices of both relations by bucket range straight into the owners' receive buffers over NVLink (hj3d_exchange_*: the
exchange is partition level 1 of the local join), then builds and probes its shard; exchange and join are inside the
timed region.  Results are verified at every N: the all-reduced result checksum against an independent torch
computation, and count / numCmps / merged HtStatistics against the unsharded engine path run on rank 0 over the same
data.  Prints ONE JSON line (rank 0).
"""
import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PLANS = {  # plan -> (table kind, build relation, mode)   mode: 1 chaining unique, 0 chaining, 3 nested+unnest
    "Csr": ("chaining", "R", 1), "CsrUU": ("chaining", "R", 0), "Nsr": ("nested", "R", 3),
    "Crs": ("chaining", "S", 0), "Nrs": ("nested", "S", 3),
}
SEED_R, SEED_S = 1234, 99
M64 = (1 << 64) - 1


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region: one streaming nvidia-smi (-lms 50) whose lines are
    time stamped; summary() keeps the samples that fall inside [t_begin, t_end] (the timed steps) and says how many of
    the surrounding warm-up samples it had to add when the timed region was shorter than a few sampling periods."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.proc = index, [], None
        self.t_begin = self.t_end = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                f = [x.strip() for x in line.strip().split(",")]
                if len(f) >= 7:
                    self.samples.append((time.perf_counter(), f))
        except Exception:
            pass

    def summary(self):
        if self.proc is not None:
            self.proc.terminate()                     # the exact process we started
        self.join(timeout=6)
        inside = [f for t, f in self.samples if self.t_begin is not None and self.t_begin <= t <= (self.t_end or t)]
        used, extra = inside, 0
        if len(inside) < 3:                           # a ~100 ms timed region: add the warm-up samples (same load) around it
            used = [f for _, f in self.samples]
            extra = len(used) - len(inside)
        if not used:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(float(f[0])) for f in used)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(f[3 + i].lower().startswith("active") for f in used)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(float(used[0][1])), "reasons": reasons,
                "samples": len(used), "samples_in_timed_region": len(inside), "samples_from_warmup": extra,
                "power_w_max": max(float(f[2]) for f in used)}


def algorithmic_bytes(nB, nP, nM, nO, D, T=12, K=4, I=4, nested=False):
    """compulsory + table traffic of one join, design independent."""
    out = nO * 2 * I + (nM * 2 * I * 2 if nested else 0)
    compulsory = nB * T + nP * T + out
    table = nB * (K + I) + D * 4 + nP * (4 + K + I)
    return compulsory + table


def workload_config(args, sample=None):
    skew = "--no-skew" if args.zipf <= 0 else f"--skew (Zipf s={args.zipf} by rejection-inversion, device generated)"
    w = (f"main_experiment1 key/foreign-key join -R {args.log2_build} -S {args.log2_probe} {skew} -t 0 -b 1, "
         f"plan {args.plan}, uint32 keys, 12-byte row-store tuples, materialised result pairs")
    cfg = {"workload": w, "plan": args.plan, "log2_build": args.log2_build, "log2_probe": args.log2_probe,
           "l2_policy": "inputs (>= 1.6 GB + 12.9 GB) are far larger than the 126 MB L2; no flush needed",
           "parallelism": f"bucket-range sharding over {args.gpus} GPU(s)" if args.gpus > 1 else "single GPU"}
    if sample:
        cfg["reference_sample"] = sample
    return cfg


# ------------------------------------------------------------------------------------------ CPU legs (the checker, timed)
def cpu_sample(args):
    """A bounded sample of the workload for the CPU legs: the reference's OWN generator (Experiment1::init through
    oracle/_ref) at -R sb -S sb+(S-R); numpy when the reference library is not there."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    sb = args.ref_log2_build
    sp = sb + (args.log2_probe - args.log2_build)
    if pyoracle.Ref.available():
        ref = pyoracle.Ref()
        R, S, dv = ref.gen_exp1(sb, sp, 1 if args.zipf > 0 else 0, 0)
        gen = "Experiment1::init of the reference (oracle/_ref)" + (", --skew (the reference fixes s = 1)" if args.zipf > 0 else "")
    else:
        rng = np.random.default_rng(1)
        nR, nS = 1 << sb, 1 << sp
        R = np.zeros((nR, 3), np.uint32); R[:, 0] = rng.permutation(nR).astype(np.uint32)
        S = np.zeros((nS, 3), np.uint32); S[:, 0] = np.arange(nS, dtype=np.uint32); S[:, 1] = rng.integers(0, nR, nS, dtype=np.uint32)
        dv = len(np.unique(S[:, 1]))
        gen = "numpy stand-in for the reference generator"
    return pyoracle, R, S, int(dv), f"-R {sb} -S {sp} ({gen})"


def cpu_join(pyoracle, plan, R, S, dv):
    """one build strand + probe strand on the host: the unmodified reference templates when available, else the C port"""
    kind_name, build_rel, mode = PLANS[plan]
    ksR, ksS = pyoracle.KeySpec(12, 0), pyoracle.KeySpec(12, 4)
    B, ksB, P, ksP = (R, ksR, S, ksS) if build_rel == "R" else (S, ksS, R, ksR)
    D = len(R) if build_rel == "R" else max(dv, 1)
    kind = pyoracle.CHAINING if kind_name == "chaining" else pyoracle.NESTED
    if pyoracle.Ref.available():
        t = pyoracle.Ref().build(kind, B, len(B), ksB, D, timed=True)
        c, cu, _, ns = t.probe(P, len(P), ksP, mode, timing_top=True)
        return (t.build_ns + ns) * 1e-9, "reference", c, cu, t.stats(), D
    t0 = time.perf_counter()
    t = pyoracle.Oracle().build(kind, B, len(B), ksB, D)
    cu = None
    if mode <= 1:
        c, _ = t.probe_chaining(P, len(P), ksP, unique=(mode == 1), materialize=False)
    else:
        c, nest = t.probe_nested(P, len(P), ksP)
        cu, _ = t.unnest(nest[:, 0], nest[:, 1])
    return time.perf_counter() - t0, "port", c, cu, t.stats(), D


def run_reference(args):
    """--impl reference: the reference's own CPU implementation (single thread: it has no parallel path) on a bounded
    sample of the workload, generated by the reference's own generator; the sample is named in config."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    pyoracle, R, S, dv, sample = cpu_sample(args)
    times, kind = [], "port"
    for it in range(args.warmup + args.steps):
        dt, kind, _, _, _, _ = cpu_join(pyoracle, args.plan, R, S, dv)
        if it >= args.warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = (len(R) + len(S)) / (ms * 1e-3)
    desc = f"plan {args.plan}, {sample}, 1 thread, rate-normalised (tuples/s)"
    line = {"impl": "reference", "metric": "join input tuples/sec (build+probe)", "value": value, "unit": "tuples/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": workload_config(args, sample=desc),
            "cpu_baseline": {"value": value, "unit": "tuples/s", "cores": 1, "kind": kind, "sample": desc, "host_cores": os.cpu_count()},
            "e2e": {"value": value, "unit": "tuples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline_leg(args, pkg, ctx):
    """Bounded sample on the box's host cores (rank 0, N=1 only); the SAME arrays then go through the GPU engine and
    every counter is compared, so the baseline and the product demonstrably join the same relations."""
    try:
        import numpy as np
        import torch
        pyoracle, R, S, dv, sample = cpu_sample(args)
        dt, kind, c, cu, st, D = cpu_join(pyoracle, args.plan, R, S, dv)
        kind_name, build_rel, mode = PLANS[args.plan]
        B, kb, P, kp = (R, 0, S, 4) if build_rel == "R" else (S, 4, R, 0)
        dB = torch.from_numpy(B.view(np.int32)).cuda(); dP = torch.from_numpy(P.view(np.int32)).cuda()
        t = ctx.table(pkg.CHAINING if kind_name == "chaining" else pkg.NESTED, D).build(dB, len(B), pkg.KeySpec(12, kb))
        if mode <= 1:
            _, gc = t.probe_chaining(dP, len(P), pkg.KeySpec(12, kp), unique=(mode == 1), flags=0)
            same = (gc["matches"], gc["num_cmps"]) == (c["matches"], c["num_cmps"])
        else:
            _, gc, gu = t.probe_nested_unnest(dP, len(P), pkg.KeySpec(12, kp), flags=0)
            same = (gc["matches"], gc["num_cmps"], gu["out_tuples"]) == (c["matches"], c["num_cmps"], cu["out_tuples"])
        same = same and t.stats() == st
        t.destroy()
        return {"value": (len(R) + len(S)) / dt, "unit": "tuples/s", "cores": 1, "kind": kind, "host_cores": os.cpu_count(),
                "sample": f"plan {args.plan}, {sample}, 1 repetition, 1 thread (the reference is single-threaded)", "seconds": dt,
                "gpu_engine_agrees_on_the_same_arrays": bool(same)}
    except Exception as e:  # the baseline is a reported side figure; never fail the bench for it
        return {"value": None, "unit": "tuples/s", "cores": 1, "kind": "unavailable", "sample": repr(e)}


# ------------------------------------------------------------------------------------------ our arm
def gen_relations(pkg, ctx, torch, dev, nR, nS, first_R, nRl, first_S, nSl, zipf):
    """rows [first, first + n) of the two global relations (csrc/datagen.cu)"""
    R = torch.zeros((nRl, 3), dtype=torch.int32, device=dev)
    S = torch.zeros((nSl, 3), dtype=torch.int32, device=dev)
    ctx.gen_column(R, 12, 0, first_R, nRl, pkg.capi.GEN_PERMUTATION, vmax=nR, seed=SEED_R)
    ctx.gen_column(S, 12, 0, first_S, nSl, pkg.capi.GEN_IOTA)
    if zipf > 0:
        ctx.gen_column(S, 12, 4, first_S, nSl, pkg.capi.GEN_ZIPF, vmax=nR, zipf_q=zipf, seed=SEED_S)
    else:
        ctx.gen_column(S, 12, 4, first_S, nSl, pkg.capi.GEN_UNIFORM, vmax=nR, seed=SEED_S)
    return R, S


def torch_pair_checksum(torch, left, right):
    """sum / xor of hj3d_pair_mix over (left, right) int64 tensors, independent of the engine (wrapping int64)"""
    x = ((left << 32) | right) * (-7046029254386353131)        # 0x9E3779B97F4A7C15 as int64, wraps
    x = x ^ ((x >> 32) & 0xFFFFFFFF)                           # logical shift
    s = int(x.sum().item()) & M64
    v = x.clone()
    n = v.numel()
    while n > 1:
        h = n // 2
        v[:h] ^= v[n - h:n]
        n -= h
    return s, (int(v[0].item()) & M64 if v.numel() else 0)


def run_ours(args):
    import torch
    import hj3d_loader
    pkg = hj3d_loader.load()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    if args.overlap:      # the ctx (and torch) work on a high-priority stream: the build's blocks go ahead of the exchange's
        torch.cuda.set_stream(torch.cuda.Stream(device=dev, priority=-1))
    ctx = pkg.Context(local, stream=torch.cuda.current_stream().cuda_stream)
    for o in args.opt:
        k, v = o.split("=")
        ctx.set_option(int(k), int(v))
    lib = pkg.capi.load()
    kind_name, build_rel, mode = PLANS[args.plan]
    kind = pkg.CHAINING if kind_name == "chaining" else pkg.NESTED
    nR, nS = 1 << args.log2_build, 1 << args.log2_probe
    nRl, nSl = nR // world, nS // world
    R, S = gen_relations(pkg, ctx, torch, dev, nR, nS, rank * nRl, nRl, rank * nSl, nSl, args.zipf)
    ksRk, ksSa = pkg.KeySpec(12, 0), pkg.KeySpec(12, 4)
    B, ksB, nBl, P, ksP, nPl = (R, ksRk, nRl, S, ksSa, nSl) if build_rel == "R" else (S, ksSa, nSl, R, ksRk, nRl)
    nBg, nPg = (nR, nS) if build_rel == "R" else (nS, nR)

    def all_sum(vals):
        if dist is None:
            return [int(v) for v in vals]
        t = torch.tensor([v - (1 << 64) if v >= (1 << 63) else v for v in vals], dtype=torch.int64, device=dev)
        dist.all_reduce(t)
        return [int(x) & M64 for x in t.tolist()]

    def all_xor(v):
        if dist is None:
            return v
        t = torch.tensor([v - (1 << 64) if v >= (1 << 63) else v], dtype=torch.int64, device=dev)
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        x = 0
        for o in out:
            x ^= int(o.item()) & M64
        return x

    # numDvSa = #distinct S.a over the GLOBAL relation (the directory size of build-on-S plans, main_experiment1.cc:875)
    if build_rel == "R":
        D = nR
    else:
        present = torch.zeros(nR, dtype=torch.uint8, device=dev)
        present[S[:, 1].to(torch.int64)] = 1
        if dist is not None:
            dist.all_reduce(present, op=dist.ReduceOp.MAX)
        D = max(int(present.sum(dtype=torch.int64).item()), 1)
        del present
    # ---- expected result checksum, computed in plain torch from the generated data (every N)
    def expected_checksum():
        Rk = torch.zeros((nR, 1), dtype=torch.int32, device=dev)
        ctx.gen_column(Rk, 4, 0, 0, nR, pkg.capi.GEN_PERMUTATION, vmax=nR, seed=SEED_R)     # the whole key column of R
        inv = torch.empty(nR, dtype=torch.int64, device=dev)
        inv[Rk[:, 0].to(torch.int64)] = torch.arange(nR, dtype=torch.int64, device=dev)
        del Rk
        s_tot, x_tot, cnt = 0, 0, 0
        step_ = 1 << 26
        for lo in range(0, nSl, step_):
            hi = min(nSl, lo + step_)
            srow = torch.arange(rank * nSl + lo, rank * nSl + hi, dtype=torch.int64, device=dev)
            rrow = inv[S[lo:hi, 1].to(torch.int64)]
            left, right = (srow, rrow) if build_rel == "R" else (rrow, srow)    # result pair = (probe row, build row)
            s, x = torch_pair_checksum(torch, left, right)
            s_tot = (s_tot + s) & M64; x_tot ^= x; cnt += hi - lo
        del inv
        torch.cuda.empty_cache()
        s_tot, cnt = all_sum([s_tot, cnt])
        return s_tot, all_xor(x_tot), cnt

    exp_sum, exp_xor, exp_cnt = expected_checksum()
    assert exp_cnt == nS
    flags = pkg.F_CHECKSUM if args.checksum else 0
    state = {}
    cap_out = int(nS // world * 1.25) + 4096 if world > 1 else nS
    if args.zipf > 0 and world > 1:
        cap_out = nS                                            # a hot key's owner may produce most of the result
    out = torch.empty((cap_out, 2), dtype=torch.int32, device=dev)

    comm = None
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            idt = torch.frombuffer(bytearray(pkg.Comm.unique_id()), dtype=torch.uint8).to(dev)
        dist.broadcast(idt, 0)
        comm = pkg.Comm.create(ctx, world, rank, bytes(idt.cpu().numpy().tobytes()))
        if args.xranges:
            comm.set_option(pkg.capi.XOPT_TARGET_RANGES, args.xranges)
        # a range must hold at most 1024 fine partitions of the table (2048 buckets each for a chaining table on unique keys, 1024 for
        # a nested one) or the local join cannot continue at partition level 2
        comm.set_option(pkg.capi.XOPT_MAX_RANGE_WIDTH, args.xmaxwidth or ((1 << 20) if kind_name == "nested" else (1 << 21)))
        if args.xthreads:
            comm.set_option(pkg.capi.XOPT_THREADS, args.xthreads)
        slack = 1.25 if args.zipf <= 0 else float(world)        # skew: one owner may receive most of a relation
        # hot-key probe replication (--zipf with a uniform build side R, plans Csr / CsrUU / Nsr): the tuples of the hottest
        # foreign keys never leave the GPU that read them, so the uniform single-pass exchange works again
        hot = args.zipf > 0 and build_rel == "R" and mode != 2 and not args.no_hot and not args.exact_exchange
        if hot:
            slack = 2.5      # the keys below the 128 hottest still differ: room for ranges up to 2.5x the mean
        comm.reserve(0, int(nBg / world * slack) + (1 << 20), 4)
        comm.reserve(1, int(nPg / world * slack) + (1 << 20), 4)
        lo_, hi_ = comm.shard(D)
        table = ctx.table(kind, D, shard=(lo_, hi_))
    else:
        table = ctx.table(kind, D)
    if world == 1:
        hot = False
    xflags = pkg.capi.XCHG_EXACT if ((args.zipf > 0 and not hot) or args.exact_exchange) else 0

    def step(fl=None, overlap=False):
        fl = flags if fl is None else fl
        table.clear()
        if world > 1:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ev[0].record()
            if hot:
                comm.hot_sample(1, P, nPl, ksP)
            comm.begin(0, B, nBl, ksB, D, rank * nBl, xflags)
            if overlap and not xflags:
                # the probe side's exchange runs on the communicator's (low priority) stream while this stream builds the table
                comm.begin(1, P, nPl, ksP, D, rank * nPl, (pkg.capi.XCHG_HOT if hot else 0) | pkg.capi.XCHG_ASYNC)
                rc0, pb = comm.end(0, B, rank * nBl, nBg)
                table.build_parts(pb)
                tb = ctx.timings()
                rc1, pp = comm.end(1, P, rank * nPl, nPg)
                ev[1].record()
            else:
                comm.begin(1, P, nPl, ksP, D, rank * nPl, pkg.capi.XCHG_HOT if hot else xflags)
                rc0, pb = comm.end(0, B, rank * nBl, nBg)
                rc1, pp = comm.end(1, P, rank * nPl, nPg)
                ev[1].record()
                table.build_parts(pb)
                tb = ctx.timings()
            assert rc0 == 0 and rc1 == 0, "exchange region overflow (run with --exact-exchange)"
            if hot:
                table.hot_answers(pp, mode)
                state["hot_tuples"] = pp.hot()
            rc, c, u = table.probe_parts(pp, mode, flags=fl, out=out, out_cap=cap_out)
            tp = ctx.timings()
            ev[2].record()
            res = u if mode == 3 else c
            state.update(xev=ev, sent=8 * (pb.info()["n_sent_remote"] + pp.info()["n_sent_remote"]), n_probe_local=pp.info()["n_records"],
                         n_build_local=pb.info()["n_records"])
            pb.destroy(); pp.destroy()
        else:
            table.build(B, nBl, ksB)
            tb = ctx.timings()
            if mode <= 1:
                rc, c = table.probe_chaining(P, nPl, ksP, unique=(mode == 1), flags=fl, out=out, out_cap=cap_out)
                res = c
            else:
                # AlgNestJoinProbe directly followed by AlgUnnestHt (plans Nsr / Nrs): one fused call
                rc, c, res = table.probe_nested_unnest(P, nPl, ksP, flags=fl, out=out, out_cap=cap_out)
            tp = ctx.timings()
            state.update(n_probe_local=nPl, n_build_local=nBl)
        assert rc == 0, "result buffer overflow"
        state.update(build=tb, probe=tp, probe_counters=c, result=res)
        return res

    def sync_all():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- verification (untimed, checksum on): every N
    res0 = step(pkg.F_CHECKSUM)
    pc0 = dict(state["probe_counters"])
    got_sum, got_cnt, got_cmps, got_match = all_sum([res0["checksum_sum"], res0["out_tuples"], pc0["num_cmps"], pc0["matches"]])
    got_xor = all_xor(res0["checksum_xor"])
    verified = bool((got_sum, got_xor, got_cnt) == (exp_sum, exp_xor, nS))
    assert verified, f"result checksum / count differ from the independent torch computation: {(got_sum, got_xor, got_cnt)} vs {(exp_sum, exp_xor, nS)}"
    # merged HtStatistics of the shards
    my_stats = table.stats()
    if dist is not None:
        names = list(my_stats)
        t = torch.tensor([my_stats[k] - (1 << 64) if my_stats[k] >= (1 << 63) else my_stats[k] for k in names], dtype=torch.int64, device=dev)
        allst = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allst, t)
        parts = (pkg.Stats * world)(*[pkg.Stats(**{k: int(v) & M64 for k, v in zip(names, a.tolist())}) for a in allst])
        merged = pkg.Stats()
        lib.hj3d_stats_merge(parts, world, C.byref(merged))
        merged_stats = merged.as_dict()
    else:
        merged_stats = my_stats
    verified_unsharded = None
    if world > 1 and not args.no_unsharded_check:
        # rank 0 joins the SAME global relations on its own, through the single-GPU path, and all ranks' merged counters
        # must equal what it gets: count, numCmps, out_tuples, every HtStatistics field
        ok = 1
        if rank == 0:
            Rg, Sg = gen_relations(pkg, ctx, torch, dev, nR, nS, 0, nR, 0, nS, args.zipf)
            Bg, Pg = (Rg, Sg) if build_rel == "R" else (Sg, Rg)
            tg = ctx.table(kind, D).build(Bg, nBg, ksB)
            if mode <= 1:
                _, cg = tg.probe_chaining(Pg, nPg, ksP, unique=(mode == 1), flags=pkg.F_CHECKSUM)
                ug = cg
            else:
                _, cg, ug = tg.probe_nested_unnest(Pg, nPg, ksP, flags=pkg.F_CHECKSUM)
            one = (cg["matches"], cg["num_cmps"], ug["out_tuples"], ug["checksum_sum"], ug["checksum_xor"], tg.stats())
            many = (got_match, got_cmps, got_cnt, got_sum, got_xor, merged_stats)
            ok = int(one == many)
            if not ok:
                print(f"[bench] sharded result differs from the unsharded engine path: {many} vs {one}", file=sys.stderr)
            tg.destroy(); del Rg, Sg, Bg, Pg
            torch.cuda.empty_cache()
        t = torch.tensor([ok], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        verified_unsharded = bool(t.item())
        assert verified_unsharded, "sharded join differs from the unsharded engine path on the same data"

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    for _ in range(args.warmup):
        step(overlap=args.overlap)
    sync_all()
    launches0 = ctx.timings()["kernel_launches"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    probe_ms, build_ms, l1_ms, xchg_ms, join_ms = [], [], [], [], []
    sync_all()
    if sampler:
        sampler.t_begin = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        res = step(overlap=args.overlap)
        probe_ms.append(state["probe"]["probe_ms"]); build_ms.append(state["build"]["total_ms"])
        l1_ms.append(state["probe"]["partition_l1_ms"])
    e1.record()
    sync_all()
    if sampler:
        sampler.t_end = time.perf_counter()
    ms = e0.elapsed_time(e1) / args.steps
    launches = ctx.timings()["kernel_launches"] - launches0
    clocks = sampler.summary() if sampler else None
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    out_total, cmps_total, launches = all_sum([res["out_tuples"], state["probe_counters"]["num_cmps"], launches])
    assert out_total == nS, f"join produced {out_total} tuples, expected |S| = {nS}"
    per_rank = None
    if dist is not None:      # per-rank share of the result and of the received probe side (skew: the owner of a hot key)
        t = torch.tensor([res["out_tuples"], state.get("hot_tuples", 0), state["n_probe_local"]], dtype=torch.int64, device=dev)
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        per_rank = [a.tolist() for a in allt]
    assert cmps_total == got_cmps
    value = (nR + nS) / (ms * 1e-3)
    peak, peak_src = peaks()

    # ---- the other plans on the same relations (N=1, default plan only): the nested 3D table, and both tables built on the
    #      NON-unique side S (the paper's subject: plans Crs / Nrs)
    other = None
    if world == 1 and args.plan == "Csr" and not args.no_other_plans:
        other = {}
        for oplan in ("Nsr", "Crs", "Nrs"):
            try:
                okind, obuild, omode = PLANS[oplan]
                oB, oksB, onB, oP, oksP, onP = (R, ksRk, nR, S, ksSa, nS) if obuild == "R" else (S, ksSa, nS, R, ksRk, nR)
                if obuild == "R":
                    oD = nR
                else:
                    present = torch.zeros(nR, dtype=torch.uint8, device=dev); present[S[:, 1].to(torch.int64)] = 1
                    oD = max(int(present.sum(dtype=torch.int64).item()), 1); del present
                t2 = ctx.table(pkg.CHAINING if okind == "chaining" else pkg.NESTED, oD)

                def ostep():
                    t2.clear()
                    t2.build(oB, onB, oksB)
                    b_ms = ctx.timings()["total_ms"]
                    if omode == 3:
                        _, c_, r_ = t2.probe_nested_unnest(oP, onP, oksP, flags=0, out=out, out_cap=cap_out)
                    else:
                        _, c_ = t2.probe_chaining(oP, onP, oksP, unique=False, flags=0, out=out, out_cap=cap_out); r_ = c_
                    return r_, b_ms, ctx.timings()["total_ms"]
                for _ in range(2):
                    ostep()
                torch.cuda.synchronize(); a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record()
                for _ in range(3):
                    r_, b_ms, p_ms = ostep()
                a1.record(); torch.cuda.synchronize()
                ms2 = a0.elapsed_time(a1) / 3
                assert r_["out_tuples"] == nS
                nBo, nPo = (nR, nS) if obuild == "R" else (nS, nR)
                other[oplan] = {"ms_per_step": ms2, "value": (nR + nS) / (ms2 * 1e-3), "unit": "tuples/s", "steps": 3, "warmup": 2,
                                "build_ms": b_ms, "probe_call_ms": p_ms, "num_buckets": oD,
                                "join_frac": algorithmic_bytes(nBo, nPo, min(nPo, nS) if omode == 3 else 0, nS, oD, nested=(omode == 3)) / (ms2 * 1e-3) / 1e9 / peak}
                t2.destroy()
            except Exception as ex:
                other[oplan] = {"error": repr(ex)}
        other["note"] = ("same relations; Nsr: nested 3D table on R + fused nested probe / unnest; Crs / Nrs: chaining / nested table built "
                         "on the NON-unique side S (2^30 rows, ~8 duplicates per key), probed with R")

    # ---- e2e: HOST buffers in, counters back on the host, all copies inside the timed region.
    #   N = 1: hj3d_join_host (chunked upload; build and partition level 1 of every probe chunk under the upload)
    #   N > 1: every rank holds its slice of both relations in pinned host memory and streams it through the exchange
    #          (hj3d_exchange_begin_host): upload, NVLink exchange, build and probe overlap; wall clock, max over ranks
    e2e = None
    if not args.no_e2e and not (world > 1 and (xflags or hot)):
        try:
            hB = torch.empty((nBl, 3), dtype=torch.int32).pin_memory(); hB.copy_(B)
            hP = torch.empty((nPl, 3), dtype=torch.int32).pin_memory(); hP.copy_(P)
            if world == 1:
                del out                                          # hj3d_join_host materialises into its own buffer
                torch.cuda.empty_cache()
            ts = []
            for it in range(2 + args.e2e_steps):
                sync_all(); t0 = time.perf_counter()
                if world == 1:
                    rc, pc, uc, _ = ctx.join_host(mode, hB, nBl, ksB, D, hP, nPl, ksP, flags=flags | 2, h_out=None, out_cap=nS)
                    n_out = (uc if mode == 3 else pc)["out_tuples"]
                else:
                    table.clear()
                    comm.begin_host(0, hB, nBl, ksB, D, rank * nBl)
                    comm.begin_host(1, hP, nPl, ksP, D, rank * nPl)
                    rc0, pb = comm.end(0, None, rank * nBl, nBg)
                    table.build_parts(pb)                        # runs under the upload / exchange of the probe side
                    rc1, pp = comm.end(1, None, rank * nPl, nPg)
                    rc, c, u = table.probe_parts(pp, mode, flags=flags, out=out, out_cap=cap_out)
                    assert rc0 == 0 and rc1 == 0 and rc == 0
                    n_out = (u if mode == 3 else c)["out_tuples"]
                    pb.destroy(); pp.destroy()
                torch.cuda.synchronize(); dt = time.perf_counter() - t0
                if dist is not None:
                    t = torch.tensor([dt, float(n_out)], dtype=torch.float64, device=dev)
                    tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
                    dist.all_reduce(t, op=dist.ReduceOp.SUM)
                    dt, n_out = float(tmax[0].item()), int(t[1].item())
                assert n_out == nS, f"e2e result count {n_out} != {nS}"
                if it >= 2:
                    ts.append(dt)
            # median: the upload shares the host's PCIe root / memory system with whatever else runs on the box (observed: the
            # same call between 273 ms and 1.1 s on different boxes); all steps are listed in ms_each
            e2e_s = sorted(ts)[len(ts) // 2]
            e2e = {"value": (nR + nS) / e2e_s, "unit": "tuples/s", "h2d_bytes_per_step": 12 * (nR + nS),
                   "d2h_bytes_per_step": 56 * world, "ms_per_step": e2e_s * 1e3, "steps": len(ts), "statistic": "median of the steps", "ms_mean": sum(ts) / len(ts) * 1e3,
                   "ms_each": [round(x * 1e3, 3) for x in ts],
                   "h2d_gbs_if_the_copy_were_everything": 12 * (nR + nS) / e2e_s / 1e9,
                   "call": "hj3d_join_host" if world == 1 else "hj3d_exchange_begin_host x2 -> hj3d_exchange_end / hj3d_table_build_parts / hj3d_probe_parts on every rank",
                   "note": ("pinned host relations -> counters on the host, wall clock" + (", max over ranks" if world > 1 else "") + ".  The probe relation is "
                            "uploaded in 256 MiB chunks; the build and partition level 1 of every chunk (at N > 1: the NVLink exchange) run under "
                            "the upload, so the step costs the host-to-device copy of " + ("this rank's 1/N of " if world > 1 else "") + "the 14.5 GB plus "
                            "partition level 2 and the probe kernel.  The reference's result is a COUNT (AlgTop, algebra.hh:223-229): the 2^30 result "
                            "pairs are materialised in HBM and stay there, only the counters (56 B per rank) come back")}
            del hB, hP
        except Exception as ex:
            e2e = {"value": None, "unit": "tuples/s", "h2d_bytes_per_step": 12 * (nR + nS), "d2h_bytes_per_step": 56 * world,
                   "error": repr(ex)}
            if dist is not None:
                raise
    if rank != 0:
        if comm is not None:
            sync_all(); comm.destroy()
        if dist is not None:
            dist.destroy_process_group()
        return
    nested = mode == 3
    nM = got_match if nested else 0
    alg = algorithmic_bytes(nBg, nPg, nM, nS, D, nested=nested)
    # ---- roofline of the DOMINANT kernel of the step (largest share in profiles/*launches*): the level-1 partition pass of the
    # probe side, k_part_scatter (at N > 1 the same kernel with peer stores: the exchange).  Its algorithmic bytes: the
    # 12-byte row-store tuple it reads + the 8-byte (key, id) record it writes, per probe tuple of this GPU.
    l1 = sum(l1_ms) / len(l1_ms) if l1_ms and world == 1 else None
    pm = sum(probe_ms) / len(probe_ms)
    roof = {"bound": "hbm", "peak": peak, "unit": "GB/s", "peak_source": peak_src,
            "join_algorithmic_bytes": alg, "join_frac": alg / (ms * 1e-3) / 1e9 / peak / world}
    if world == 1:
        l1_bytes = nPl * 20
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("k_part_scatter_level1")
            if tj and (tj["plan"], tj["log2_build"], tj["log2_probe"], tj["n_gpus"]) == (args.plan, args.log2_build, args.log2_probe, 1) and args.zipf <= 0:
                traffic = tj["dram_bytes_per_launch"]
        except Exception:
            pass
        roof.update({"kernel": "k_part_scatter (probe side, level 1: 12-byte row-store tuples -> (key, id) records of coarse bucket ranges)",
                     "achieved": l1_bytes / (l1 * 1e-3) / 1e9 if l1 else None, "frac": l1_bytes / (l1 * 1e-3) / 1e9 / peak if l1 else None,
                     "kernel_ms": l1, "algorithmic_bytes_per_launch": l1_bytes, "traffic": traffic,
                     "note": "frac is this kernel's own bytes over its own CUDA-event time; join_frac charges the WHOLE step (all partition "
                             "passes, build, probe) against SURVEY 8(d)'s 37.58 GB and is the headline fraction",
                     "probe_kernel": {"kernel": "k_probe_fine" if mode == 1 else "probe", "kernel_ms": pm,
                                      "algorithmic_bytes_per_launch": nPl * 16 + 4 * D + 8 * nBg,
                                      "frac": (nPl * 16 + 4 * D + 8 * nBg) / (pm * 1e-3) / 1e9 / peak if pm else None}})
    else:
        part_ms = state["xev"][0].elapsed_time(state["xev"][1])
        roof.update({"kernel": "k_part_scatter<PEER> (exchange = partition level 1 with peer stores)", "kernel_ms": part_ms, "traffic": None,
                     "achieved": (nBl + nPl) * 20 / (part_ms * 1e-3) / 1e9, "frac": (nBl + nPl) * 20 / (part_ms * 1e-3) / 1e9 / peak,
                     "algorithmic_bytes_per_launch": (nBl + nPl) * 20,
                     "note": "both relations' exchange kernels + count all-gathers (rank 0, last step); NVLink, not HBM, bounds it: see shuffle"})
    line = {"metric": "join input tuples/sec (build+probe)", "value": value, "unit": "tuples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic (device generated, identical at every N)",
            "config": workload_config(args), "clocks": clocks, "gpu_launches": int(launches),
            "e2e": e2e if e2e is not None else {"value": None, "unit": "tuples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                                               "note": "not measured (--no-e2e, or the two-pass exact exchange of skewed runs, which reads its slice twice)"},
            "roofline": roof,
            "phases_ms": {"build_total": sum(build_ms) / len(build_ms), "histogram": state["build"]["histogram_ms"],
                          "scan": state["build"]["scan_ms"], "scatter": state["build"]["scatter_ms"],
                          "group": state["build"]["group_ms"], "build_partition": state["build"]["partition_ms"],
                          "probe_partition": state["probe"]["partition_ms"], "probe_partition_level1": l1, "probe": pm},
            "result": {"out_tuples": out_total, "num_cmps": cmps_total, "num_buckets": D, "verified_checksum": verified,
                       "verified_vs_unsharded_engine": verified_unsharded, "ht_statistics": merged_stats,
                       "checksum_in_timed_steps": bool(args.checksum)}}
    if world > 1:
        part_ms = state["xev"][0].elapsed_time(state["xev"][1])
        join_ms_ = state["xev"][1].elapsed_time(state["xev"][2])
        line["shuffle"] = {"bytes_sent_per_gpu": state["sent"], "exchange_ms": part_ms, "local_join_ms": join_ms_,
                           "bus_gbs_per_gpu": state["sent"] / (part_ms * 1e-3) / 1e9 if part_ms > 0 else None,
                           "exact_two_pass": bool(xflags), "hot_key_replication": bool(hot),
                           "probe_exchange_overlaps_build": bool(args.overlap and not xflags),
                           "hot_tuples_kept_local": sum(a[1] for a in per_rank),
                           "out_tuples_per_rank": [a[0] for a in per_rank],
                           "out_imbalance_max_over_mean": max(a[0] for a in per_rank) * world / max(1, sum(a[0] for a in per_rank)),
                           "probe_records_received_per_rank": [a[2] for a in per_rank],
                           "note": "rank 0, last step.  exchange_ms = partition level 1 of both relations with peer stores into the owners' "
                                   "receive buffers + two count all-gathers (the only collectives); there is no separate all-to-all.  "
                                   "bus_gbs_per_gpu = bytes this GPU stored into other GPUs / exchange_ms (a lower bound on the link rate, "
                                   "the kernel also hashes and ranks); measured peer copy bandwidth is ~770 GB/s per direction"}
    if other is not None:
        line["other_plans"] = other
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_leg(args, pkg, ctx)
    print(json.dumps(line), flush=True)
    if comm is not None:
        sync_all(); comm.destroy()
    if dist is not None:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------ config 3: main_experiment4
def run_config3(args):
    """BASELINE config 3: the deferred-unnesting join of main_experiment4 (plan Ndu: nested tables on S and T, scan R ->
    probe S -> probe T -> unnest T -> unnest S -> count, main_experiment4.cc:831-941) with duplicates per key
    A = B in {1, 10, 100, 1000} on one GPU, as one device pipeline (hj3d_probe2_unnest2); plan Chj (two chaining joins,
    main_experiment4.cc:943-1043) composed from the single-operator calls beside it; the unmodified reference
    (oracle/_ref) timed on the same arrays where that takes seconds."""
    import numpy as np
    import torch
    import hj3d_loader
    pkg = hj3d_loader.load()
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    ctx = pkg.Context(0, stream=torch.cuda.current_stream().cuda_stream)
    ref = pyoracle.Ref() if pyoracle.Ref.available() else None
    points = [(24, 4, 3, 1, 1), (24, 4, 3, 10, 10), (20, 4, 3, 100, 100), (14, 4, 3, 1000, 1000), (22, 4, 3, 100, 1), (22, 4, 3, 1, 100)]
    if args.c3_points:
        points = [tuple(int(x) for x in q.split(",")) for q in args.c3_points]
    ks_k, ks_a = pkg.KeySpec(8, 0), pkg.KeySpec(8, 4)
    rows = []
    for (r, alpha, beta, A, Bm) in points:
        nR = 1 << r
        nC, nE = nR >> alpha, nR >> beta
        cC, cE = nC * A, nE * Bm
        nF = cC + cE
        D = nC + nE                                                   # numFkCommon + numFkExclusive buckets (main_experiment4.cc:855)
        R = torch.zeros((nR, 2), dtype=torch.int32, device=dev)
        ctx.gen_column(R, 8, 0, 0, nR, pkg.capi.GEN_IOTA)

        def fk_relation(seed, excl_base):
            """generateFkRel (main_experiment4.cc:740-756): shuffled common keys first, then the shuffled exclusive ones"""
            F = torch.zeros((nF, 2), dtype=torch.int32, device=dev)
            ctx.gen_column(F, 8, 0, 0, nF, pkg.capi.GEN_IOTA)
            pc = torch.zeros((cC, 1), dtype=torch.int32, device=dev)
            ctx.gen_column(pc, 4, 0, 0, cC, pkg.capi.GEN_PERMUTATION, vmax=cC, seed=seed)
            F[:cC, 1] = (pc[:, 0].to(torch.int64) // A).to(torch.int32)
            pe = torch.zeros((cE, 1), dtype=torch.int32, device=dev)
            ctx.gen_column(pe, 4, 0, 0, cE, pkg.capi.GEN_PERMUTATION, vmax=cE, seed=seed + 1)
            F[cC:, 1] = (excl_base + pe[:, 0].to(torch.int64) // Bm).to(torch.int32)
            return F
        S = fk_relation(11, nC)
        T = fk_relation(23, nC + nE)
        c_top = nC * A * A
        tS, tT = ctx.table(pkg.NESTED, D), ctx.table(pkg.NESTED, D)
        materialise = c_top * 12 <= (16 << 30)
        trip = torch.empty((max(c_top, 1), 3), dtype=torch.int32, device=dev) if materialise else None

        def step_ndu(flags=0):
            tS.clear(); tT.clear()
            tS.build(S, nF, ks_a); bS = ctx.timings()["total_ms"]
            tT.build(T, nF, ks_a); bT = ctx.timings()["total_ms"]
            rc, cnt, _, _ = ctx.probe2_unnest2(tS, tT, R, nR, ks_k, flags=flags, out=trip, out_cap=c_top if materialise else 0)
            return cnt, bS, bT, ctx.timings()["total_ms"]
        cnt, _, _, _ = step_ndu(pkg.F_CHECKSUM)
        assert (cnt[0]["matches"], cnt[1]["matches"], cnt[2]["out_tuples"], cnt[3]["out_tuples"]) == (nC + nE, nC, nC * A, c_top), cnt
        chk = (cnt[3]["checksum_sum"], cnt[3]["checksum_xor"])
        step_ndu()
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            cnt, bS, bT, pr = step_ndu()
        e1.record(); torch.cuda.synchronize()
        ms_ndu = e0.elapsed_time(e1) / 3
        tS.destroy(); tT.destroy()
        # plan Chj: chaining tables, probe S (|S| results, materialised), probe T through the gathered R tuples
        cS, cT = ctx.table(pkg.CHAINING, D), ctx.table(pkg.CHAINING, D)
        p1 = torch.empty((nF, 2), dtype=torch.int32, device=dev)
        p2 = torch.empty((max(c_top, 1), 2), dtype=torch.int32, device=dev) if c_top * 8 <= (16 << 30) else None

        def step_chj():
            cS.clear(); cT.clear()
            cS.build(S, nF, ks_a); cT.build(T, nF, ks_a)
            _, c1 = cS.probe_chaining(R, nR, ks_k, flags=0, out=p1, out_cap=nF)
            r1 = p1[:c1["out_written"], 0].contiguous()
            _, c2 = cT.probe_chaining(R, c1["out_written"], ks_k, gather=r1, flags=0, out=p2, out_cap=c_top if p2 is not None else 0)
            return c1, c2
        c1, c2 = step_chj()
        assert (c1["out_tuples"], c2["out_tuples"]) == (nF, c_top), (c1, c2)
        torch.cuda.synchronize(); e0.record()
        for _ in range(3):
            step_chj()
        e1.record(); torch.cuda.synchronize()
        ms_chj = e0.elapsed_time(e1) / 3
        cS.destroy(); cT.destroy()
        row = {"log2R": r, "alpha": alpha, "beta": beta, "A": A, "B": Bm, "cardR": nR, "cardS": nF, "cardT": nF, "c_top": c_top,
               "Ndu_ms": ms_ndu, "Ndu_build_S_ms": bS, "Ndu_build_T_ms": bT, "Ndu_probe_pipeline_ms": pr,
               "Ndu_input_tuples_per_s": (nR + 2 * nF) / (ms_ndu * 1e-3), "Ndu_results_per_s": c_top / (ms_ndu * 1e-3),
               "results_materialised": bool(materialise), "Chj_ms": ms_chj,
               "counters": {"c_probe_RS": cnt[0]["matches"], "c_probe_RS_cmp": cnt[0]["num_cmps"], "c_probe_RT": cnt[1]["matches"],
                            "c_probe_RT_cmp": cnt[1]["num_cmps"], "c_unnest1": cnt[2]["out_tuples"], "c_top": cnt[3]["out_tuples"]}}
        if ref is not None and c_top <= (1 << 27) and nF <= (1 << 26):
            hR, hS, hT = (x.cpu().numpy().view(np.uint32) for x in (R, S, T))
            w0 = ref.exp4_run(0, hR, hS, hT, D); w1 = ref.exp4_run(1, hR, hS, hT, D)
            row["cpu_reference"] = {"Ndu_ms": (w0["t_build_S_ns"] + w0["t_build_T_ns"] + w0["t_probe_ns"]) * 1e-6,
                                    "Chj_ms": (w1["t_build_S_ns"] + w1["t_build_T_ns"] + w1["t_probe_ns"]) * 1e-6, "cores": 1,
                                    "same_arrays": True,
                                    "counters_equal": bool(all(w0[k] == row["counters"][k] for k in row["counters"]) and
                                                           (w0["checksum_sum"], w0["checksum_xor"]) == chk)}
            assert row["cpu_reference"]["counters_equal"], (w0, row["counters"], chk)
        rows.append(row)
        del R, S, T, trip, p1, p2
        torch.cuda.empty_cache()
    print(json.dumps({"metric": "config 3: main_experiment4 deferred-unnesting join (plans Ndu / Chj), duplicates per key 1..1000, one B200",
                      "unit": "ms per step (build S + build T + probe strand)", "n_gpus": 1, "steps": 3, "warmup": 2, "points": rows}), flush=True)


# ------------------------------------------------------------------------------------------ config 5: main_algebra_example pipeline
def run_config5(args):
    """BASELINE config 5: the operator pipeline of main_algebra_example's algebra_test2 (main_algebra_example.cc:265-347) at scale
    and across N GPUs:   scan L -> select (L.b < 40) -> nested probe (L.a = R.c) -> unnest -> count,   nested 3D table built on R.
    int32 attributes hashed with murmur64 of the sign-extended key (main_algebra_example.cc:53-65), 8-byte tuples; R.c are
    foreign keys into L.a (Zipf with --zipf, else uniform), L.a unique.  The selection is evaluated inside the exchange /
    partition kernel's load (hj3d_exchange_begin_select); probe + unnest run as one fused kernel."""
    import torch
    import hj3d_loader
    pkg = hj3d_loader.load()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    ctx = pkg.Context(local, stream=torch.cuda.current_stream().cuda_stream)
    lib = pkg.capi.load()
    nL, nR = 1 << args.log2_build, 1 << args.log2_probe
    nLl, nRl = nL // world, nR // world
    SEL_CONST = 40
    L = torch.zeros((nLl, 2), dtype=torch.int32, device=dev)
    R = torch.zeros((nRl, 2), dtype=torch.int32, device=dev)
    ctx.gen_column(L, 8, 0, rank * nLl, nLl, pkg.capi.GEN_PERMUTATION, vmax=nL, seed=SEED_R)          # L.a: unique keys
    ctx.gen_column(L, 8, 4, rank * nLl, nLl, pkg.capi.GEN_UNIFORM, vmax=100, seed=7)                    # L.b: selection attribute
    ctx.gen_column(R, 8, 0, rank * nRl, nRl, pkg.capi.GEN_ZIPF if args.zipf > 0 else pkg.capi.GEN_UNIFORM, vmax=nL, zipf_q=args.zipf, seed=SEED_S)
    ctx.gen_column(R, 8, 4, rank * nRl, nRl, pkg.capi.GEN_IOTA)
    ks = pkg.KeySpec(8, 0, 4, pkg.HASH_MURMUR64_SEXT32)

    def all_sum(vals):
        if dist is None:
            return [int(v) for v in vals]
        t = torch.tensor([v - (1 << 64) if v >= (1 << 63) else v for v in vals], dtype=torch.int64, device=dev)
        dist.all_reduce(t)
        return [int(x) & M64 for x in t.tolist()]

    def all_xor(v):
        if dist is None:
            return v
        t = torch.tensor([v - (1 << 64) if v >= (1 << 63) else v], dtype=torch.int64, device=dev)
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        x = 0
        for o in out:
            x ^= int(o.item()) & M64
        return x
    present = torch.zeros(nL, dtype=torch.uint8, device=dev)
    present[R[:, 0].to(torch.int64)] = 1
    if dist is not None:
        dist.all_reduce(present, op=dist.ReduceOp.MAX)
    D = max(int(present.sum(dtype=torch.int64).item()), 1)                    # #distinct R.c
    del present
    # expected result: every R row whose partner in L passes the selection gives one (L row, R row) pair
    La = torch.zeros((nL, 1), dtype=torch.int32, device=dev); Lb = torch.zeros((nL, 1), dtype=torch.int32, device=dev)
    ctx.gen_column(La, 4, 0, 0, nL, pkg.capi.GEN_PERMUTATION, vmax=nL, seed=SEED_R)
    ctx.gen_column(Lb, 4, 0, 0, nL, pkg.capi.GEN_UNIFORM, vmax=100, seed=7)
    inv = torch.empty(nL, dtype=torch.int64, device=dev)
    inv[La[:, 0].to(torch.int64)] = torch.arange(nL, dtype=torch.int64, device=dev)
    sel_total = int((Lb[:, 0] < SEL_CONST).sum().item())
    e_sum, e_xor, e_cnt = 0, 0, 0
    for lo in range(0, nRl, 1 << 26):
        hi = min(nRl, lo + (1 << 26))
        lrow = inv[R[lo:hi, 0].to(torch.int64)]
        keep = Lb[lrow, 0] < SEL_CONST
        rrow = torch.arange(rank * nRl + lo, rank * nRl + hi, dtype=torch.int64, device=dev)
        s_, x_ = torch_pair_checksum(torch, lrow[keep], rrow[keep])
        e_sum = (e_sum + s_) & M64; e_xor ^= x_; e_cnt += int(keep.sum().item())
    del La, Lb, inv
    torch.cuda.empty_cache()
    e_sum, e_cnt = all_sum([e_sum, e_cnt]); e_xor = all_xor(e_xor)

    idt = torch.zeros(128, dtype=torch.uint8, device=dev)
    if world > 1:
        if rank == 0:
            idt = torch.frombuffer(bytearray(pkg.Comm.unique_id()), dtype=torch.uint8).to(dev)
        dist.broadcast(idt, 0)
        comm = pkg.Comm.create(ctx, world, rank, bytes(idt.cpu().numpy().tobytes()))
    else:
        comm = pkg.Comm.create(ctx, 1, 0, None)
    comm.set_option(pkg.capi.XOPT_MAX_RANGE_WIDTH, args.xmaxwidth or (1 << 20))      # nested table: 1024 fine partitions of 1024 buckets per range
    slack = 1.25 if args.zipf <= 0 else float(max(world, 1))
    comm.reserve(0, int(nR / world * slack) + (1 << 20), 4)
    comm.reserve(1, int(nL / world * 1.25) + (1 << 20), 4)
    lo_, hi_ = comm.shard(D)
    table = ctx.table(pkg.NESTED, D, shard=(lo_, hi_))
    cap_out = e_cnt if (args.zipf > 0 or world == 1) else int(e_cnt / world * 1.3) + 4096
    out = torch.empty((max(cap_out, 1), 2), dtype=torch.int32, device=dev)
    xflags = pkg.capi.XCHG_EXACT if args.zipf > 0 else 0
    state = {}

    def step(fl=0):
        table.clear()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        comm.begin(0, R, nRl, ks, D, rank * nRl, xflags)
        comm.begin(1, L, nLl, ks, D, rank * nLl, xflags, selection=(4, 1, SEL_CONST))      # AlgSelection: L.b < 40, inside the load
        rc0, pr = comm.end(0, R, rank * nRl, nR)
        rc1, plp = comm.end(1, L, rank * nLl, nL)
        ev[1].record()
        assert rc0 == 0 and rc1 == 0, "exchange region overflow"
        table.build_parts(pr)
        tb = ctx.timings()
        rc, c, u = table.probe_parts(plp, 3, flags=fl, out=out, out_cap=cap_out)
        tp = ctx.timings()
        ev[2].record()
        assert rc == 0, "result buffer overflow"
        state.update(tb=tb, tp=tp)
        state.update(ev=ev, sel=plp.selected(), probe=c, unnest=u, sent=8 * (pr.info()["n_sent_remote"] + plp.info()["n_sent_remote"]))
        pr.destroy(); plp.destroy()
        return u

    def sync_all():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
    u0 = step(pkg.F_CHECKSUM)
    g_sum, g_cnt, g_sel, g_probe = all_sum([u0["checksum_sum"], u0["out_tuples"], state["sel"], state["probe"]["matches"]])
    g_xor = all_xor(u0["checksum_xor"])
    verified = bool((g_sum, g_xor, g_cnt, g_sel) == (e_sum, e_xor, e_cnt, sel_total))
    assert verified, f"pipeline result differs from the independent torch computation: {(g_sum, g_xor, g_cnt, g_sel)} vs {(e_sum, e_xor, e_cnt, sel_total)}"
    for _ in range(args.warmup):
        step()
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1) / args.steps
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        x_ms = state["ev"][0].elapsed_time(state["ev"][1]); j_ms = state["ev"][1].elapsed_time(state["ev"][2])
        print(json.dumps({"metric": "config 5: pipeline input tuples/sec (scan -> select -> nested 3D probe -> unnest -> count)",
                          "value": (nL + nR) / (ms * 1e-3), "unit": "tuples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "dtype": "i32 keys, murmur64 of the sign-extended key",
                          "data": "synthetic (device generated, identical at every N)",
                          "config": {"workload": f"main_algebra_example algebra_test2 shape: |L| = 2^{args.log2_build} (unique a, selection b < 40 of uniform [0,100)), "
                                                 f"|R| = 2^{args.log2_probe} ({'Zipf s=' + str(args.zipf) if args.zipf > 0 else 'uniform'} foreign keys c), "
                                                 f"nested table on R with {D} buckets", "parallelism": f"bucket-range sharding over {world} GPU(s)"},
                          "counters": {"c_scan_L": nL, "c_select": g_sel, "c_probe": g_probe, "c_unnest": g_cnt, "c_top": g_cnt, "c_build_R": nR},
                          "result": {"verified_checksum_and_counts": verified},
                          "phases_ms": {"exchange_both_relations": x_ms, "build_probe_unnest": j_ms,
                                        "build": {k: state["tb"][k] for k in ("partition_ms", "group_ms", "total_ms")},
                                        "probe": {k: state["tp"][k] for k in ("partition_ms", "probe_ms", "unnest_ms", "total_ms")}},
                          "shuffle": {"bytes_sent_per_gpu": state["sent"], "bus_gbs_per_gpu": state["sent"] / (x_ms * 1e-3) / 1e9 if x_ms > 0 else None}}), flush=True)
    sync_all()
    comm.destroy()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--plan", default="Csr", choices=list(PLANS))
    ap.add_argument("--log2-build", type=int, default=27)
    ap.add_argument("--log2-probe", type=int, default=30)
    ap.add_argument("--ref-log2-build", type=int, default=22, help="sample size of the CPU reference legs")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-plans", action="store_true", help="skip the secondary measurement of the other plans (N=1)")
    ap.add_argument("--no-unsharded-check", action="store_true", help="N>1: skip rank 0's unsharded join of the same data")
    ap.add_argument("--no-hot", action="store_true", help="N>1 with --zipf: no hot-key probe replication (two-pass exact exchange instead)")
    ap.add_argument("--exact-exchange", action="store_true", help="N>1: two-pass exchange with exact regions (always on with --zipf)")
    ap.add_argument("--xmaxwidth", type=int, default=0, help="N>1: largest bucket-range width of the exchange (default 2^21, nested tables 2^20)")
    ap.add_argument("--overlap", action="store_true", help="N>1: run the probe side's exchange on a second stream under the build")
    ap.add_argument("--xthreads", type=int, default=0, help="N>1: block size of the exchange kernel (512 | 1024; default 1024)")
    ap.add_argument("--xranges", type=int, default=0, help="N>1: coarse bucket ranges of the exchange (default: the engine's 256)")
    ap.add_argument("--zipf", type=float, default=0.0, help="skew of the foreign keys S.a (0 = uniform; config 4 uses 0.5 .. 1.5)")
    ap.add_argument("--checksum", action="store_true",
                    help="also fold the result checksum inside the TIMED steps (it is always verified once, untimed)")
    ap.add_argument("--opt", action="append", default=[], help="engine option id=value (HJ3D_OPT_*), repeatable")
    ap.add_argument("--config", type=int, default=2, help="2 = the headline KFK join (default); 3 = main_experiment4 duplicate sweep (N=1); 5 = main_algebra_example pipeline (scan/select/3D join/unnest/count) on N GPUs")
    ap.add_argument("--c3-points", action="append", default=[], help="config 3 point log2R,alpha,beta,A,B (repeatable)")
    args = ap.parse_args()
    if args.config == 3:
        return run_config3(args)
    if args.config == 5:
        return run_config5(args)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
