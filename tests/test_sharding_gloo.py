"""The N > 1 path on CPU: world_size 2 over gloo.  The device kernel of the path (hj3d_partition_by_owner) is restated
in numpy here; everything after it -- owner ranges from the C ABI, the all-to-all-v of (key, global row id) records,
the local joins, the merge of the counters -- is the code bench.py --gpus N runs (3d-hashjoin_b200/sharding.py), with
the CPU oracle standing in for the per-rank join.  The sharded result must equal the unsharded oracle's: counters and
the pair multiset, chaining (IsBuildKeyUnique) and nested + unnest."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def murmur32_np(x):
    x = x.astype(np.uint32).copy()
    x ^= x >> np.uint32(16); x *= np.uint32(0x85ebca6b); x ^= x >> np.uint32(13); x *= np.uint32(0xc2b2ae35); x ^= x >> np.uint32(16)
    return x


def pair_mix_np(l, r):
    """hj3d_pair_mix / orc_pair_mix on arrays (wrapping uint64 arithmetic)"""
    x = (l.astype(np.uint64) << np.uint64(32)) | r.astype(np.uint64)
    with np.errstate(over="ignore"):
        x = x * np.uint64(0x9E3779B97F4A7C15)
    return x ^ (x >> np.uint64(32))


def relations(nR, nS, seed):
    rng = np.random.default_rng(seed)
    R = np.zeros((nR, 3), np.uint32); R[:, 0] = rng.permutation(nR)
    S = np.zeros((nS, 3), np.uint32); S[:, 0] = np.arange(nS); S[:, 1] = rng.integers(0, nR + nR // 8, nS)   # some probes miss
    return R, S


def partition_by_owner_np(keys, rowid_base, D, width, world):
    """what hj3d_partition_by_owner produces: (key, global row id) records grouped by owner + counts"""
    owner = (murmur32_np(keys) % np.uint32(D)) // np.uint32(width)
    order = np.argsort(owner, kind="stable")
    recs = np.stack([keys[order], (np.arange(len(keys), dtype=np.uint32) + np.uint32(rowid_base))[order]], axis=1).astype(np.uint32)
    return recs, np.bincount(owner, minlength=world).tolist()


def worker(rank, world, port, nR, nS, D, mode, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    import hj3d_loader
    import pyoracle as pyo
    pkg = hj3d_loader.load()
    lib = pkg.capi.load()
    from helpers import sorted_pairs
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dev = torch.device("cpu")
        R, S = relations(nR, nS, 5)
        lo_, hi_ = C.c_uint64(), C.c_uint64()
        assert lib.hj3d_owner_range(D, world, rank, C.byref(lo_), C.byref(hi_)) == 0
        width = (D + world - 1) // world
        assert lo_.value == min(rank * width, D) and hi_.value == min((rank + 1) * width, D)
        nRl, nSl = nR // world, nS // world
        parts, counts = [], []
        for rel, col, nl in ((R, 0, nRl), (S, 1, nSl)):
            sl = rel[rank * nl:(rank + 1) * nl]
            recs, cnt = partition_by_owner_np(sl[:, col], rank * nl, D, width, world)
            parts.append(torch.from_numpy(recs.view(np.int32))); counts.append(cnt)
        # both relations through one exchange (one counts collective), as bench.py --gpus N does; the build side once more
        # through the single-relation call: same records
        got = pkg.sharding.exchange_many(dist, parts, counts, dev, [None, torch.empty((nSl * 2, 2), dtype=torch.int32)])
        again, _ = pkg.sharding.exchange_records(dist, parts[0], counts[0], dev)
        assert torch.equal(again, got[0][0])
        mine = {}
        for name, (recv, rcounts) in zip(("B", "P"), got):
            arr = np.ascontiguousarray(recv.numpy().view(np.uint32))
            assert len(arr) == sum(rcounts)
            b = murmur32_np(arr[:, 0]) % np.uint32(D)
            assert np.all((b >= lo_.value) & (b < hi_.value)), "a record reached a rank that does not own its bucket"
            mine[name] = arr
        orc = pyo.Oracle()
        ks = pyo.KeySpec(8, 0, 4, 0, 4)
        kind = pyo.CHAINING if mode == 1 else pyo.NESTED
        t = orc.build(kind, mine["B"], len(mine["B"]), ks, D)
        if mode == 1:
            c, pairs = t.probe_chaining(mine["P"], len(mine["P"]), ks, unique=True)
        else:
            cp, nest = t.probe_nested(mine["P"], len(mine["P"]), ks)
            c, pairs = t.unnest(nest[:, 0], nest[:, 1])
            c = dict(c); c["num_cmps"] = cp["num_cmps"]
        # the oracle reports the probe-side POSITION as the left id; the engine carries the global row id of the record
        pairs = np.asarray(pairs, dtype=np.uint32).reshape(-1, 2).copy()
        pairs[:, 0] = mine["P"][pairs[:, 0], 1]
        mx = pair_mix_np(pairs[:, 0], pairs[:, 1])
        c = dict(c)
        c["checksum_sum"] = int(mx.sum(dtype=np.uint64)) if len(mx) else 0
        c["checksum_xor"] = int(np.bitwise_xor.reduce(mx)) if len(mx) else 0
        merged = pkg.sharding.merge_counters(dist, c, dev)
        allp = [None] * world
        dist.all_gather_object(allp, pairs)
        if rank == 0:
            t0 = orc.build(kind, R, nR, pyo.KeySpec(12, 0), D)
            if mode == 1:
                c0, p0 = t0.probe_chaining(S, nS, pyo.KeySpec(12, 4), unique=True)
            else:
                cp0, n0 = t0.probe_nested(S, nS, pyo.KeySpec(12, 4))
                c0, p0 = t0.unnest(n0[:, 0], n0[:, 1])
                c0 = dict(c0); c0["num_cmps"] = cp0["num_cmps"]
            for k in ("matches", "num_cmps", "out_tuples", "checksum_sum", "checksum_xor"):
                assert merged[k] == c0[k], (k, merged[k], c0[k])
            assert np.array_equal(sorted_pairs(np.concatenate(allp)), sorted_pairs(p0))
            q.put("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", [1, 3])
@pytest.mark.parametrize("D", [4096, 3001])
def test_world_size_2_sharded_join_equals_unsharded(mode, D):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = 29600 + (os.getpid() % 300) + mode * 7 + (D % 5)
    mp.spawn(worker, args=(2, port, 4096, 16384, D, mode, q), nprocs=2, join=True)
    assert q.get() == "ok"
