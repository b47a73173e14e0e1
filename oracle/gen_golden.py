"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libhj3d_ref.so).

TEST INFRASTRUCTURE.  Runs only in the build container (needs /root/reference to have been
compiled by `make -C oracle ref`); the GPU box only ever reads the committed fixtures.

Each fixture holds the generated input relations (the reference's own generators, i.e.
Experiment1::init / Experiment4::init with std::mt19937's default seed under libstdc++) and
the counters / statistics the reference's operators produce on them, for every plan of the
corresponding driver.  The numbers match SURVEY.md Appendix B (captured from the stock binaries).

    python oracle/gen_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from pyoracle import CHAINING, NESTED, KeySpec, Ref, build_libs  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def exp1_case(ref, log2R, log2S, skew, t, b):
    R, S, dv = ref.gen_exp1(log2R, log2S, skew, t)
    ksRk, ksSa = KeySpec(12, 0), KeySpec(12, 4)
    nR, nS = len(R), len(S)
    dR, dS = max(nR // b, 1), max(dv // b, 1)           # main_experiment1.cc:651,875
    plans = {  # plan -> (kind, build rel, build ks, D, probe rel, probe ks, mode)
        "Csr":   (CHAINING, R, ksRk, dR, S, ksSa, 1),
        "CsrUU": (CHAINING, R, ksRk, dR, S, ksSa, 0),
        "Crs":   (CHAINING, S, ksSa, dS, R, ksRk, 0),
        "Nsr":   (NESTED,   R, ksRk, dR, S, ksSa, 3),
        "Nrs":   (NESTED,   S, ksSa, dS, R, ksRk, 3),
        "NrsNU": (NESTED,   S, ksSa, dS, R, ksRk, 2),
    }
    res = {}
    for name, (kind, B, ksB, D, P, ksP, mode) in plans.items():
        tab = ref.build(kind, B, len(B), ksB, D)
        c, cu, _, _ = tab.probe(P, len(P), ksP, mode)
        res[name] = {"kind": kind, "mode": mode, "D": D, "probe": c, "unnest": cu, "stats": tab.stats()}
    meta = {"log2R": log2R, "log2S": log2S, "skew": bool(skew), "t": t, "b": b, "numDvSa": dv,
            "fkMax": 1 << (log2R - t), "plans": res}
    return R[:, 0].copy(), S[:, 1].copy(), meta


def exp4_case(ref, log2R, a, b, A, B):
    R, S, T = ref.gen_exp4(log2R, a, A, b, B)
    nR = len(R)
    D = (nR >> a) + (nR >> b)                                # main_experiment4.cc:855
    meta = {"log2R": log2R, "alpha": a, "beta": b, "mA": A, "mB": B, "D": D,
            "Ndu": ref.exp4_run(0, R, S, T, D), "Chj": ref.exp4_run(1, R, S, T, D)}
    for plan in ("Ndu", "Chj"):
        for k in ("t_build_S_ns", "t_build_T_ns", "t_probe_ns"):
            meta[plan].pop(k)
    ks = KeySpec(8, 4)
    meta["stats_S_nested"] = ref.build(NESTED, S, len(S), ks, D).stats()
    meta["stats_T_nested"] = ref.build(NESTED, T, len(T), ks, D).stats()
    meta["stats_S_chaining"] = ref.build(CHAINING, S, len(S), ks, D).stats()
    return S[:, 1].copy(), T[:, 1].copy(), meta


def example_case(ref):
    """main_algebra_example.cc:152-158,205-220: int attributes, murmur64 of the sign-extended key, 5 buckets."""
    L = np.array([[1, 11], [2, 21], [3, 31], [4, 41]], np.int32)
    Rr = np.array([[1, -1], [1, -2], [1, -3], [2, -1], [2, -2], [3, -1]], np.int32)
    Lsel = L[L[:, 1] < 40]                                   # SelectionL (:31-38)
    ksL, ksR = KeySpec(8, 0, 4, 2), KeySpec(8, 0, 4, 2)
    out = {}
    tab = ref.build(NESTED, Rr, len(Rr), ksR, 5)
    c, cu, pairs, _ = tab.probe(Lsel, len(Lsel), ksL, 2)
    out["test1_nested_nu"] = {"probe": c, "pairs": pairs.tolist(), "stats": tab.stats()}
    c, cu, pairs, _ = tab.probe(Lsel, len(Lsel), ksL, 3)
    out["test2_nested_unnest"] = {"probe": c, "unnest": cu, "pairs": pairs.tolist()}
    tab = ref.build(CHAINING, Rr, len(Rr), ksR, 5)
    c, cu, pairs, _ = tab.probe(Lsel, len(Lsel), ksL, 0)
    out["test3_chaining"] = {"probe": c, "pairs": pairs.tolist(), "stats": tab.stats()}
    return L, Rr, out


def main():
    build_libs(ref=True)
    ref = Ref()
    os.makedirs(OUT, exist_ok=True)
    for tag, args in (("exp1_R10_S12_uni_t0_b1", (10, 12, False, 0, 1)),
                      ("exp1_R12_S14_zipf_t2_b2", (12, 14, True, 2, 2)),
                      ("exp1_R8_S13_zipf_t0_b4", (8, 13, True, 0, 4))):
        Rk, Sa, meta = exp1_case(ref, *args)
        np.savez_compressed(os.path.join(OUT, tag + ".npz"), Rk=Rk, Sa=Sa, meta=json.dumps(meta))
        print(tag, {k: (v["probe"]["matches"], v["probe"]["num_cmps"]) for k, v in meta["plans"].items()})
    for tag, args in (("exp4_R12_a4_b3_A5_B7", (12, 4, 3, 5, 7)), ("exp4_R10_a2_b2_A10_B1", (10, 2, 2, 10, 1))):
        Sa, Ta, meta = exp4_case(ref, *args)
        np.savez_compressed(os.path.join(OUT, tag + ".npz"), Sa=Sa, Ta=Ta, meta=json.dumps(meta))
        print(tag, meta["Ndu"], meta["Chj"])
    L, Rr, meta = example_case(ref)
    np.savez_compressed(os.path.join(OUT, "algebra_example.npz"), L=L, R=Rr, meta=json.dumps(meta))
    print("algebra_example", meta["test2_nested_unnest"]["pairs"])


if __name__ == "__main__":
    main()
