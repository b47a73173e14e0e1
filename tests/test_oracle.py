"""CPU tests that PIN THE ORACLE: oracle/oracle_join.c against
  * the murmur known answers of SURVEY.md B.4 (util/hasht.hh:52-72),
  * the golden fixtures generated from the unmodified reference (tests/golden/, oracle/gen_golden.py),
  * the reference's own operator templates (oracle/_ref) on random inputs, when that library exists.
"""
import numpy as np
import pytest

import pyoracle as pyo
from helpers import CMP_KEYS, exp1_relations, exp4_relations, load_golden, oracle_plan, sorted_pairs, sub

KS = pyo.KeySpec


def test_murmur_kats(oracle):
    kat32 = {0: 0x00000000, 1: 0x514e28b7, 2: 0x30f4c306, 3: 0x85f0b427, 0x3ff: 0x5ce591a7, 0x400: 0x66cc183d,
             0xdeadbeef: 0x0de5c6a9, 0xffffffff: 0x81f16f39}
    kat64 = {0: 0, 1: 0xa6eaea4b026a3297, 2: 0x1fef84c63323b9a4, 3: 0xe31de9e2e7bcc69d,
             0xdeadbeef: 0xf4bb1a7ddec3ef4d, 2**64 - 1: 0x93a1564dd89219c2}
    for k, v in kat32.items():
        assert oracle.murmur32(k) == v
    for k, v in kat64.items():
        assert oracle.murmur64(k) == v


EXP1 = ["exp1_R10_S12_uni_t0_b1", "exp1_R12_S14_zipf_t2_b2", "exp1_R8_S13_zipf_t0_b4"]


def exp1_plan_args(R, S, meta, plan):
    ksRk, ksSa = KS(12, 0), KS(12, 4)
    p = meta["plans"][plan]
    if plan in ("Csr", "CsrUU", "Nsr"):
        return p["mode"], R, ksRk, p["D"], S, ksSa
    return p["mode"], S, ksSa, p["D"], R, ksRk


@pytest.mark.parametrize("name", EXP1)
@pytest.mark.parametrize("plan", ["Csr", "CsrUU", "Crs", "Nsr", "Nrs", "NrsNU"])
def test_oracle_matches_reference_exp1(oracle, name, plan):
    z, meta = load_golden(name)
    R, S = exp1_relations(z)
    assert len(np.unique(S[:, 1])) == meta["numDvSa"]
    mode, B, ksB, D, P, ksP = exp1_plan_args(R, S, meta, plan)
    o = oracle_plan(oracle, pyo, mode, B, ksB, D, P, ksP)
    g = meta["plans"][plan]
    assert o["probe"]["matches"] == g["probe"]["matches"]
    assert o["probe"]["num_cmps"] == g["probe"]["num_cmps"]
    assert o["stats"] == g["stats"]
    if mode == 3:
        assert sub(o["unnest"]) == sub(g["unnest"])           # incl. checksum of the flat result
        assert o["unnest"]["out_tuples"] == len(S)            # every plan returns |S| tuples
    elif mode == 2:
        assert sub(o["probe"]) == sub(g["probe"])
        assert o["probe"]["matches"] == meta["numDvSa"]
    else:
        assert sub(o["probe"]) == sub(g["probe"])
        assert o["probe"]["out_tuples"] == len(S)


def test_survey_appendix_b_numbers():
    """The fixture values are the ones captured from the stock binaries."""
    _, m = load_golden("exp1_R10_S12_uni_t0_b1")
    assert m["numDvSa"] == 1007 and m["fkMax"] == 1024
    want = {"Csr": 6146, "CsrUU": 8174, "Crs": 8299, "Nsr": 6100, "Nrs": 1532, "NrsNU": 1532}
    assert {k: v["probe"]["num_cmps"] for k, v in m["plans"].items()} == want
    assert m["plans"]["Crs"]["stats"]["cc_max"] == 35 and m["plans"]["Nrs"]["stats"]["cc_max"] == 5
    _, m = load_golden("exp1_R12_S14_zipf_t2_b2")
    assert m["numDvSa"] == 982
    want = {"Csr": 38232, "CsrUU": 50026, "Crs": 169428, "Nsr": 32848, "Nrs": 8247, "NrsNU": 8247}
    assert {k: v["probe"]["num_cmps"] for k, v in m["plans"].items()} == want
    assert m["plans"]["Crs"]["stats"]["cc_max"] == 2279
    _, m = load_golden("exp4_R12_a4_b3_A5_B7")
    assert (m["Ndu"]["c_probe_RS"], m["Ndu"]["c_probe_RS_cmp"], m["Ndu"]["c_probe_RT"], m["Ndu"]["c_probe_RT_cmp"],
            m["Ndu"]["c_unnest1"], m["Ndu"]["c_unnest2"], m["Ndu"]["c_top"]) == (768, 4465, 256, 847, 1280, 6400, 6400)
    assert (m["Chj"]["c_probe_RS"], m["Chj"]["c_probe_RS_cmp"], m["Chj"]["c_probe_RT"], m["Chj"]["c_probe_RT_cmp"],
            m["Chj"]["c_top"]) == (4864, 30623, 6400, 39226, 6400)


def oracle_exp4(oracle, R, S, T, D):
    """Ndu and Chj of main_experiment4.cc:831-1043 composed from the oracle's operators."""
    ksR, ksF = KS(8, 0), KS(8, 4)
    mix = oracle.pair_mix
    out = {}
    # --- Ndu: probe RS (nested) -> probe RT (nested, key reached through r) -> unnest T -> unnest S
    tS, tT = oracle.build(pyo.NESTED, S, len(S), ksF, D), oracle.build(pyo.NESTED, T, len(T), ksF, D)
    c1, n1 = tS.probe_nested(R, len(R), ksR)                          # (r, sgroup)
    c2, n2 = tT.probe_nested(R, len(n1), ksR, gather=n1[:, 0].copy())  # (idx into n1, tgroup)
    u1, f1 = tT.unnest(n2[:, 0], n2[:, 1])                            # (idx into n1, t)
    sg = n1[f1[:, 0], 1]
    u2, f2 = tS.unnest(np.arange(len(f1), dtype=np.uint32), sg)       # (idx into f1, s)
    r = n1[f1[f2[:, 0], 0], 0]; t = f1[f2[:, 0], 1]; s = f2[:, 1]
    ms = [mix(mix(int(a), int(b)) & 0xFFFFFFFF, int(c)) for a, b, c in zip(r, s, t)]
    out["Ndu"] = {"c_probe_RS": c1["matches"], "c_probe_RS_cmp": c1["num_cmps"], "c_probe_RT": c2["matches"],
                  "c_probe_RT_cmp": c2["num_cmps"], "c_unnest1": u1["out_tuples"], "c_unnest2": u2["out_tuples"],
                  "c_top": len(ms), "checksum_sum": sum(ms) % 2**64, "checksum_xor": int(np.bitwise_xor.reduce(np.array(ms, np.uint64))) if ms else 0}
    # --- Chj: two chaining probes, no unnest
    cS, cT = oracle.build(pyo.CHAINING, S, len(S), ksF, D), oracle.build(pyo.CHAINING, T, len(T), ksF, D)
    c1, p1 = cS.probe_chaining(R, len(R), ksR)                        # (r, s)
    c2, p2 = cT.probe_chaining(R, len(p1), ksR, gather=p1[:, 0].copy())  # (idx into p1, t)
    r = p1[p2[:, 0], 0]; s = p1[p2[:, 0], 1]; t = p2[:, 1]
    ms = [mix(mix(int(a), int(b)) & 0xFFFFFFFF, int(c)) for a, b, c in zip(r, s, t)]
    out["Chj"] = {"c_probe_RS": c1["matches"], "c_probe_RS_cmp": c1["num_cmps"], "c_probe_RT": c2["matches"],
                  "c_probe_RT_cmp": c2["num_cmps"], "c_unnest1": 0, "c_unnest2": 0,
                  "c_top": len(ms), "checksum_sum": sum(ms) % 2**64, "checksum_xor": int(np.bitwise_xor.reduce(np.array(ms, np.uint64))) if ms else 0}
    return out


@pytest.mark.parametrize("name", ["exp4_R12_a4_b3_A5_B7", "exp4_R10_a2_b2_A10_B1"])
def test_oracle_matches_reference_exp4(oracle, name):
    z, meta = load_golden(name)
    R, S, T = exp4_relations(z, meta)
    got = oracle_exp4(oracle, R, S, T, meta["D"])
    assert got["Ndu"] == meta["Ndu"]
    assert got["Chj"] == meta["Chj"]
    # analytic identities (main_experiment4.cc:216-223,584-597)
    nR = len(R); nC = nR >> meta["alpha"]; nE = nR >> meta["beta"]
    assert got["Ndu"]["c_probe_RS"] == nC + nE and got["Ndu"]["c_probe_RT"] == nC
    assert got["Ndu"]["c_unnest1"] == nC * meta["mA"] and got["Ndu"]["c_top"] == nC * meta["mA"] ** 2
    assert got["Chj"]["c_probe_RS"] == len(S)
    ks = KS(8, 4)
    assert oracle.build(pyo.NESTED, S, len(S), ks, meta["D"]).stats() == meta["stats_S_nested"]
    assert oracle.build(pyo.NESTED, T, len(T), ks, meta["D"]).stats() == meta["stats_T_nested"]
    assert oracle.build(pyo.CHAINING, S, len(S), ks, meta["D"]).stats() == meta["stats_S_chaining"]


def test_oracle_algebra_example_emission_order(oracle):
    """main_algebra_example.cc algebra_test1..3: also the ORDER of emission is the reference's."""
    z, meta = load_golden("algebra_example")
    L, Rr = z["L"], z["R"]
    Lsel = np.ascontiguousarray(L[L[:, 1] < 40])
    ks = KS(8, 0, 4, pyo.HASH_MURMUR64_SEXT32)
    tn = oracle.build(pyo.NESTED, Rr, len(Rr), ks, 5)
    c, nest = tn.probe_nested(Lsel, len(Lsel), ks)
    cu, flat = tn.unnest(nest[:, 0], nest[:, 1])
    assert flat.tolist() == meta["test2_nested_unnest"]["pairs"] == [[0, 0], [0, 2], [0, 1], [1, 3], [1, 4], [2, 5]]
    assert (c["matches"], cu["out_tuples"]) == (3, 6)
    assert tn.stats() == meta["test1_nested_nu"]["stats"]
    tc = oracle.build(pyo.CHAINING, Rr, len(Rr), ks, 5)
    c, pairs = tc.probe_chaining(Lsel, len(Lsel), ks)
    assert pairs.tolist() == meta["test3_chaining"]["pairs"]
    assert sub(c) == sub(meta["test3_chaining"]["probe"])
    assert tc.stats() == meta["test3_chaining"]["stats"]


# ---------------------------------------------------------------- randomized: oracle == reference templates
def rand_case(rng, layout, nB, nP, kmax, D):
    tb, kb, hid, dt = layout
    cols = tb // kb
    B = rng.integers(0, kmax, (nB, cols)).astype(dt)
    P = rng.integers(0, kmax, (nP, cols)).astype(dt)
    return B, P


LAYOUTS = {"u32x3": (12, 4, 0, np.uint32), "u32x2": (8, 4, 0, np.uint32), "i32x2_m64": (8, 4, 2, np.int32),
           "u64x2": (16, 8, 1, np.uint64), "u64x3": (24, 8, 1, np.uint64)}


@pytest.mark.parametrize("layout", list(LAYOUTS))
@pytest.mark.parametrize("shape", [(0, 50, 10, 7), (300, 0, 10, 7), (1, 1, 1, 1), (500, 400, 50, 1), (2000, 3000, 300, 257),
                                   (4000, 1000, 5000, 1024), (3000, 3000, 40, 4096)])
def test_oracle_equals_reference_templates(oracle, ref, layout, shape):
    nB, nP, kmax, D = shape
    tb, kb, hid, dt = LAYOUTS[layout]
    rng = np.random.default_rng(hash((layout, shape)) % 2**32)
    B, P = rand_case(rng, LAYOUTS[layout], nB, nP, kmax, D)
    if hid == 2:
        B -= kmax // 2; P -= kmax // 2                      # negative ints exercise the sign extension
    if hid == 1:
        B = B * np.uint64(0x9E3779B97F4A7C15); P = P * np.uint64(0x9E3779B97F4A7C15)   # full 64-bit keys
    ksB, ksP = KS(tb, kb, kb, hid), KS(tb, 0, kb, hid)      # build on 2nd attribute, probe on 1st
    for mode in (0, 1, 2, 3):
        o = oracle_plan(oracle, pyo, mode, B, ksB, D, P, ksP)
        rt = ref.build(pyo.CHAINING if mode <= 1 else pyo.NESTED, B, nB, ksB, D)
        c, cu, pairs, _ = rt.probe(P, nP, ksP, mode)
        assert o["stats"] == rt.stats(), (layout, shape, mode)
        assert o["probe"]["matches"] == c["matches"] and o["probe"]["num_cmps"] == c["num_cmps"], (layout, shape, mode)
        if mode == 3:
            assert sub(o["unnest"]) == sub(cu)
        else:
            assert sub(o["probe"]) == sub(c)
        # emission order is the reference's too
        assert np.array_equal(np.asarray(o["pairs"]).reshape(-1, 2), pairs.reshape(-1, 2)), (layout, shape, mode)


def test_oracle_gather_equals_reference(oracle, ref):
    rng = np.random.default_rng(5)
    B = rng.integers(0, 200, (1000, 2)).astype(np.uint32)
    P = rng.integers(0, 200, (300, 2)).astype(np.uint32)
    g = rng.integers(0, 300, 777).astype(np.uint32)
    ksB, ksP = KS(8, 4), KS(8, 0)
    for mode in (0, 1, 2, 3):
        o = oracle_plan(oracle, pyo, mode, B, ksB, 97, P, ksP, gather=g)
        rt = ref.build(pyo.CHAINING if mode <= 1 else pyo.NESTED, B, len(B), ksB, 97)
        c, cu, pairs, _ = rt.probe(P, len(g), ksP, mode, gather=g)
        assert o["probe"]["matches"] == c["matches"] and o["probe"]["num_cmps"] == c["num_cmps"]
        assert np.array_equal(sorted_pairs(o["pairs"]), sorted_pairs(pairs))
