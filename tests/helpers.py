"""Shared helpers for the parity tests: run the same plan through the CPU oracle and the CUDA engine."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CMP_KEYS = ("matches", "num_cmps", "out_tuples", "checksum_sum", "checksum_xor")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return z, json.loads(str(z["meta"]))


def exp1_relations(z):
    """12-byte {k,a,b} tuples of Experiment1 (main_experiment1.cc:86,494-515): R.k shuffled keys, S.k = iota."""
    Rk, Sa = z["Rk"], z["Sa"]
    R = np.zeros((len(Rk), 3), np.uint32); R[:, 0] = Rk
    S = np.zeros((len(Sa), 3), np.uint32); S[:, 0] = np.arange(len(Sa), dtype=np.uint32); S[:, 1] = Sa
    return R, S


def exp4_relations(z, meta):
    """8-byte {k,a} tuples of Experiment4 (main_experiment4.cc:150,730-756)."""
    nR = 1 << meta["log2R"]
    R = np.zeros((nR, 2), np.uint32); R[:, 0] = np.arange(nR, dtype=np.uint32)
    S = np.stack([np.arange(len(z["Sa"]), dtype=np.uint32), z["Sa"]], axis=1)
    T = np.stack([np.arange(len(z["Ta"]), dtype=np.uint32), z["Ta"]], axis=1)
    return R, np.ascontiguousarray(S), np.ascontiguousarray(T)


def sorted_pairs(p):
    p = np.asarray(p, dtype=np.uint32).reshape(-1, 2)
    v = (p[:, 0].astype(np.uint64) << np.uint64(32)) | p[:, 1].astype(np.uint64)
    return np.sort(v)


def sub(d, keys=CMP_KEYS):
    return {k: d[k] for k in keys}


# ---------------------------------------------------------------- oracle side
def oracle_plan(orc, pyo, mode, B, ksB, D, P, ksP, gather=None):
    """mode 0 chaining, 1 chaining unique, 2 nested, 3 nested+unnest.  Returns dict(probe, unnest, stats, pairs)
    where nested pairs are (left, first row of the group)."""
    kind = pyo.CHAINING if mode <= 1 else pyo.NESTED
    t = orc.build(kind, B, len(B), ksB, D)
    res = {"stats": t.stats(), "unnest": None}
    if mode <= 1:
        c, pairs = t.probe_chaining(P, len(gather) if gather is not None else len(P), ksP, unique=(mode == 1), gather=gather)
        res.update(probe=c, pairs=pairs)
    else:
        c, nest = t.probe_nested(P, len(gather) if gather is not None else len(P), ksP, gather=gather)
        res["probe"] = c
        if mode == 2:
            # translate oracle group refs to "first row of the group"
            first = {}
            cu, flat = t.unnest(nest[:, 0], nest[:, 1])
            # first element of each group in emission order is the MainNode's own tuple
            pos = 0
            fr = np.zeros(len(nest), np.uint32)
            for i, g in enumerate(nest[:, 1]):
                fr[i] = flat[pos, 1]
                pos += t.group_len(int(g))
            res["pairs"] = np.stack([nest[:, 0], fr], axis=1) if len(nest) else np.zeros((0, 2), np.uint32)
        else:
            cu, flat = t.unnest(nest[:, 0], nest[:, 1])
            res.update(unnest=cu, pairs=flat)
    return res


# ---------------------------------------------------------------- GPU side
def to_dev(a):
    import torch
    a = np.ascontiguousarray(a)
    return torch.from_numpy(a.view(np.uint8).reshape(-1)).cuda()


def gpu_plan(pkg, ctx, mode, B, ksB, D, P, ksP, gather=None, materialize=True, flags=None, cap=None):
    import torch
    flags = pkg.F_CHECKSUM if flags is None else flags
    kind = pkg.CHAINING if mode <= 1 else pkg.NESTED
    dB, dP = to_dev(B), to_dev(P)
    dG = torch.from_numpy(np.ascontiguousarray(gather).view(np.int32)).cuda() if gather is not None else None
    nP = len(gather) if gather is not None else len(P)
    t = ctx.table(kind, D).build(dB, len(B), ksB)
    res = {"stats": t.stats(), "unnest": None, "pairs": None}
    if mode <= 1:
        rc, c = t.probe_chaining(dP, nP, ksP, unique=(mode == 1), gather=dG, flags=flags)   # count only
        res["probe_count_only"] = c
        if materialize:
            n_out = c["out_tuples"] if cap is None else cap
            out = torch.zeros((max(n_out, 1), 2), dtype=torch.int32, device="cuda")
            rc, c = t.probe_chaining(dP, nP, ksP, unique=(mode == 1), gather=dG, flags=flags, out=out, out_cap=n_out)
            res["pairs"] = out[:c["out_written"]].cpu().numpy().view(np.uint32)
        res["probe"], res["rc"] = c, rc
    else:
        nest = torch.zeros((max(nP, 1), 2), dtype=torch.int32, device="cuda")
        rc, c = t.probe_nested(dP, nP, ksP, gather=dG, flags=flags, out=nest, out_cap=nP)
        res["probe"], res["rc"] = c, rc
        m = c["out_written"]
        left, gref = nest[:m, 0].contiguous(), nest[:m, 1].contiguous()
        if mode == 2:
            fr = torch.zeros(max(m, 1), dtype=torch.int32, device="cuda")
            t.group_first_row(gref, m, fr)
            ctx.sync()
            res["pairs"] = np.stack([left.cpu().numpy().view(np.uint32), fr[:m].cpu().numpy().view(np.uint32)], axis=1)
        else:
            rc, cu = t.unnest(left, gref, m, flags=flags)                                    # count only
            n_out = cu["out_tuples"]
            out = torch.zeros((max(n_out, 1), 2), dtype=torch.int32, device="cuda")
            rc, cu = t.unnest(left, gref, m, flags=flags, out=out, out_cap=n_out)
            res["unnest"] = cu
            # the same unnest fed with the (left, group ref) pairs as the probe wrote them (no column split)
            out2 = torch.zeros((max(n_out, 1), 2), dtype=torch.int32, device="cuda")
            rc2, cu2 = t.unnest_pairs(nest, m, flags=flags, out=out2, out_cap=n_out)
            assert rc2 == rc and cu2 == cu, (cu2, cu)
            assert np.array_equal(sorted_pairs(out2[:cu2["out_written"]].cpu().numpy().view(np.uint32)),
                                  sorted_pairs(out[:cu["out_written"]].cpu().numpy().view(np.uint32)))
            if gather is None:
                # nested probe + unnest in one call (fused kernel on the fine-partition path, composition otherwise)
                out3 = torch.zeros((max(n_out, 1), 2), dtype=torch.int32, device="cuda")
                rc3, pc3, uc3 = t.probe_nested_unnest(dP, nP, ksP, flags=flags, out=out3, out_cap=n_out)
                assert rc3 == rc and (pc3["matches"], pc3["num_cmps"]) == (c["matches"], c["num_cmps"]), (pc3, c)
                assert {k: uc3[k] for k in ("out_tuples", "checksum_sum", "checksum_xor", "out_written")} == \
                       {k: cu[k] for k in ("out_tuples", "checksum_sum", "checksum_xor", "out_written")}, (uc3, cu)
                assert np.array_equal(sorted_pairs(out3[:uc3["out_written"]].cpu().numpy().view(np.uint32)),
                                      sorted_pairs(out[:cu["out_written"]].cpu().numpy().view(np.uint32)))
            res["pairs"] = out[:cu["out_written"]].cpu().numpy().view(np.uint32)
    res["size"] = t.size()
    t.destroy()
    return res


def assert_plan_equal(g, o, what=""):
    assert sub(g["probe"]) == sub(o["probe"]), f"{what}: probe counters {sub(g['probe'])} != {sub(o['probe'])}"
    assert g["stats"] == o["stats"], f"{what}: stats {g['stats']} != {o['stats']}"
    if o.get("unnest") is not None:
        assert sub(g["unnest"]) == sub(o["unnest"]), f"{what}: unnest counters"
    if g.get("pairs") is not None and o.get("pairs") is not None:
        assert np.array_equal(sorted_pairs(g["pairs"]), sorted_pairs(o["pairs"])), f"{what}: result multiset differs"
