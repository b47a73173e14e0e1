"""Parity at the sizes BASELINE.json's configurations state, with the engine's DEFAULT options (the path the headline
number runs), against the reference itself: oracle/_ref = the unmodified reference templates and generators
(Experiment1::init / Experiment4::init), compiled by oracle/Makefile and shipped to the GPU box as a built library.

  * config 1: main_experiment1 -R 20 -S 23, uniform and --skew, all six plans            (main_experiment1.cc:415-457,624-1285)
  * the default two-level shared-memory path: -R 24 -S 27, plans Csr and Nsr                (numCmps and HtStatistics pinned)
  * config 3: main_experiment4 with duplicates per key A, B in {1, 10, 100, 1000}           (main_experiment4.cc:517-575,831-1043)
Every counter, every HtStatistics field and the order-independent checksum of the result multiset are compared.
"""
import numpy as np
import pytest

import pyoracle as pyo
from helpers import sub, to_dev

pytestmark = pytest.mark.gpu

PLANS = {  # plan -> (mode, build relation)          modes: 0 chaining, 1 chaining unique, 2 nested, 3 nested + unnest
    "Csr": (1, "R"), "CsrUU": (0, "R"), "Crs": (0, "S"), "Nsr": (3, "R"), "Nrs": (3, "S"), "NrsNU": (2, "S"),
}


@pytest.fixture(scope="module")
def dctx(pkg):
    """default engine options"""
    import torch
    return pkg.Context(0, stream=torch.cuda.current_stream().cuda_stream)


def run_plan_gpu(pkg, ctx, mode, dB, nB, ksB, D, dP, nP, ksP):
    import torch
    kind = pkg.CHAINING if mode <= 1 else pkg.NESTED
    t = ctx.table(kind, D).build(dB, nB, ksB)
    res = {"stats": t.stats(), "unnest": None}
    if mode <= 1:
        _, c0 = t.probe_chaining(dP, nP, ksP, unique=(mode == 1), flags=pkg.F_CHECKSUM)          # count only
        out = torch.empty((max(c0["out_tuples"], 1), 2), dtype=torch.int32, device="cuda")
        _, c = t.probe_chaining(dP, nP, ksP, unique=(mode == 1), flags=pkg.F_CHECKSUM, out=out, out_cap=c0["out_tuples"])
        assert sub(c0) == sub(c), "count-only and materialising probes disagree"
        res["probe"] = c
    elif mode == 2:
        nest = torch.empty((max(nP, 1), 2), dtype=torch.int32, device="cuda")
        _, c = t.probe_nested(dP, nP, ksP, flags=pkg.F_CHECKSUM, out=nest, out_cap=nP)
        res["probe"] = c
    else:
        nest = torch.empty((max(nP, 1), 2), dtype=torch.int32, device="cuda")
        _, c = t.probe_nested(dP, nP, ksP, flags=pkg.F_CHECKSUM, out=nest, out_cap=nP)
        _, u0 = t.unnest_pairs(nest, c["out_written"], flags=pkg.F_CHECKSUM)
        out = torch.empty((max(u0["out_tuples"], 1), 2), dtype=torch.int32, device="cuda")
        _, u = t.unnest_pairs(nest, c["out_written"], flags=pkg.F_CHECKSUM, out=out, out_cap=u0["out_tuples"])
        # the fused probe + unnest call (what plans Nsr / Nrs run in the drivers and in bench.py)
        _, pc3, uc3 = t.probe_nested_unnest(dP, nP, ksP, flags=pkg.F_CHECKSUM, out=out, out_cap=u0["out_tuples"])
        assert (pc3["matches"], pc3["num_cmps"]) == (c["matches"], c["num_cmps"])
        assert sub(uc3, ("out_tuples", "checksum_sum", "checksum_xor")) == sub(u, ("out_tuples", "checksum_sum", "checksum_xor"))
        res["probe"], res["unnest"] = c, u
    t.destroy()
    return res


def run_plan_ref(ref, mode, B, ksB, D, P, ksP):
    t = ref.build(pyo.CHAINING if mode <= 1 else pyo.NESTED, B, len(B), ksB, D)
    c, cu, _, _ = t.probe(P, len(P), ksP, mode, materialize=False)
    return {"stats": t.stats(), "probe": c, "unnest": cu if mode == 3 else None}


def check_exp1(pkg, ctx, ref, log2R, log2S, skew, plans):
    R, S, dv = ref.gen_exp1(log2R, log2S, skew, 0)
    dR, dS = to_dev(R), to_dev(S)
    for plan in plans:
        mode, brel = PLANS[plan]
        B, dB, kb, P, dP, kp = (R, dR, 0, S, dS, 4) if brel == "R" else (S, dS, 4, R, dR, 0)
        D = len(R) if brel == "R" else dv                        # b = 1: |R| or numDvSa buckets (main_experiment1.cc:651,875)
        g = run_plan_gpu(pkg, ctx, mode, dB, len(B), pkg.KeySpec(12, kb), D, dP, len(P), pkg.KeySpec(12, kp))
        o = run_plan_ref(ref, mode, B, pyo.KeySpec(12, kb), D, P, pyo.KeySpec(12, kp))
        what = f"exp1 -R {log2R} -S {log2S} skew={skew} plan {plan}"
        assert g["stats"] == o["stats"], what
        keys = ("matches", "num_cmps") if mode == 3 else ("matches", "num_cmps", "out_tuples", "checksum_sum", "checksum_xor")
        assert sub(g["probe"], keys) == sub(o["probe"], keys), what
        if mode == 3:
            k3 = ("out_tuples", "checksum_sum", "checksum_xor")
            assert sub(g["unnest"], k3) == sub(o["unnest"], k3), what


@pytest.mark.parametrize("skew", [0, 1])
def test_config1_full_size_all_plans_default_options(pkg, dctx, ref, skew):
    check_exp1(pkg, dctx, ref, 20, 23, skew, list(PLANS))


def test_two_level_default_path_R24_S27(pkg, dctx, ref):
    check_exp1(pkg, dctx, ref, 24, 27, 0, ["Csr", "Nsr"])


def gpu_exp4_counts(pkg, ctx, R, S, T, D):
    """Ndu / Chj of main_experiment4 on the device; the flat (r, s, t) results are folded into the harness' checksum
    mix(mix(r, s) & 0xFFFFFFFF, t) on the device (torch), so 10^7..10^8 results need no host round trip."""
    import torch
    ksR, ksF = pkg.KeySpec(8, 0), pkg.KeySpec(8, 4)
    dR, dS, dT = to_dev(R), to_dev(S), to_dev(T)
    res = {}
    tS = ctx.table(pkg.NESTED, D).build(dS, len(S), ksF)
    tT = ctx.table(pkg.NESTED, D).build(dT, len(T), ksF)
    rc, cnt, trip, n_out = ctx.probe2_unnest2(tS, tT, dR, len(R), ksR, flags=pkg.F_CHECKSUM, want_triples=True)
    res["Ndu"] = dict(c_probe_RS=cnt[0]["matches"], c_probe_RS_cmp=cnt[0]["num_cmps"], c_probe_RT=cnt[1]["matches"],
                      c_probe_RT_cmp=cnt[1]["num_cmps"], c_unnest1=cnt[2]["out_tuples"], c_unnest2=cnt[3]["out_tuples"],
                      c_top=cnt[3]["out_tuples"], checksum_sum=cnt[3]["checksum_sum"], checksum_xor=cnt[3]["checksum_xor"])
    res["Ndu_triples"] = trip[:n_out]
    tS.destroy(); tT.destroy()
    return res


def torch_triple_checksum(trip):
    """mix(mix(r, s) & 0xFFFFFFFF, t) over (r, s, t) rows, in torch int64 with wrap-around (independent of the engine)."""
    import torch
    M = -7046029254386353131                                   # 0x9E3779B97F4A7C15 as int64

    def mix(l, r):
        x = ((l << 32) | r) * M
        return x ^ ((x >> 32) & 0xFFFFFFFF)
    t = trip.to(torch.int64) & 0xFFFFFFFF
    m = mix(mix(t[:, 0], t[:, 1]) & 0xFFFFFFFF, t[:, 2])
    s = int(m.sum().item()) & ((1 << 64) - 1)
    x = 0
    if len(m):
        # xor-reduce: fold halves
        v = m.clone()
        n = len(v)
        while n > 1:
            h = n // 2
            v[:h] ^= v[n - h:n]
            n -= h
        x = int(v[0].item()) & ((1 << 64) - 1)
    return s, x


@pytest.mark.parametrize("log2R,alpha,beta,A,B", [(20, 4, 3, 1, 1), (18, 4, 3, 10, 10), (14, 4, 3, 100, 100), (10, 4, 3, 1000, 1000),
                                                  (16, 4, 3, 100, 1), (16, 4, 3, 1, 100), (12, 2, 2, 1000, 10)])
def test_config3_exp4_duplicate_sweep(pkg, dctx, ref, log2R, alpha, beta, A, B):
    R, S, T = ref.gen_exp4(log2R, alpha, A, beta, B)
    D = (len(R) >> alpha) + (len(R) >> beta)                     # numFkCommon + numFkExclusive (main_experiment4.cc:855)
    want = ref.exp4_run(0, R, S, T, D)                           # plan Ndu on the unmodified reference
    want_chj = ref.exp4_run(1, R, S, T, D)
    got = gpu_exp4_counts(pkg, dctx, R, S, T, D)
    keys = ("c_probe_RS", "c_probe_RS_cmp", "c_probe_RT", "c_probe_RT_cmp", "c_unnest1", "c_unnest2", "c_top",
            "checksum_sum", "checksum_xor")
    assert {k: got["Ndu"][k] for k in keys} == {k: want[k] for k in keys}
    assert want["c_top"] == (len(R) >> alpha) * A * A == want_chj["c_top"]            # main_experiment4.cc:593-597
    assert (want["checksum_sum"], want["checksum_xor"]) == (want_chj["checksum_sum"], want_chj["checksum_xor"])
    # the materialised triples, folded independently of the engine's own checksum
    s, x = torch_triple_checksum(got["Ndu_triples"])
    assert (s, x) == (want["checksum_sum"], want["checksum_xor"])
