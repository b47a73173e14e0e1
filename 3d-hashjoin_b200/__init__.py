"""hj3d_b200 -- Python front-end (tests / bench plumbing) of the B200-native 3D hash-join engine.

The product is ``lib/libhj3d.so`` (hand-written sm_100a CUDA behind the C ABI of ``include/hj3d.h``)
plus the C++20 operator templates in ``hostcpp/hj3d/`` that mirror the reference's ``algebra.hh``.
This module only wraps the C ABI for callers that hold their relations in torch CUDA tensors.
The directory name is not an importable identifier; load it through ``hj3d_loader`` at the repo root.
"""
import ctypes as C

from . import capi
from . import sharding
from .capi import (CHAINING, F_CHECKSUM, HASH_MURMUR32, HASH_MURMUR64, HASH_MURMUR64_SEXT32, NESTED, NO_ROWID,
                   OPT_PARTITION_BYTES, OPT_PARTITION_MIN_PROBE, OPT_PARTITION_WINDOW, OPT_SMEM_CHUNK, OPT_SMEM_MIN_PROBE,
                   OPT_SMEM_PROBE, OPT_SMEM_SLICE_BYTES, OPT_WARP_AGGREGATE, Counters, Hj3dError, KeySpec,
                   Stats, Timings)

__all__ = ["Context", "Table", "Comm", "Parts", "KeySpec", "Hj3dError", "CHAINING", "NESTED", "F_CHECKSUM", "capi"]


def _ptr(t):
    """device pointer of a torch tensor / raw int / None"""
    if t is None:
        return None
    if isinstance(t, int):
        return C.c_void_p(t)
    assert t.is_cuda and t.is_contiguous(), "expected a contiguous CUDA tensor"
    return C.c_void_p(t.data_ptr())


class Context:
    """hj3d_ctx: one device, one stream."""

    def __init__(self, device=0, stream=None):
        self.lib = capi.load()
        h = C.c_void_p()
        capi.check(self.lib.hj3d_ctx_create(int(device), C.byref(h)))
        self.h = h
        self.device = int(device)
        if stream is not None:
            # torch reports the legacy default stream as 0; the C ABI takes NULL as "ctx-owned stream",
            # so name the default stream by its CUDA handle cudaStreamLegacy (0x1)
            stream = int(stream) or 1
            capi.check(self.lib.hj3d_ctx_set_stream(self.h, C.c_void_p(stream)))

    def close(self):
        if getattr(self, "h", None):
            self.lib.hj3d_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, opt, value):
        capi.check(self.lib.hj3d_ctx_set_option(self.h, opt, int(value)))

    def sync(self):
        capi.check(self.lib.hj3d_ctx_sync(self.h))

    def timings(self):
        t = Timings()
        capi.check(self.lib.hj3d_ctx_timings(self.h, C.byref(t)))
        return t.as_dict()

    def table(self, kind, num_buckets, shard=None):
        return Table(self, kind, num_buckets, shard)

    # -- column helpers ------------------------------------------------------------------------
    def split_pairs(self, pairs, n, left, right):
        capi.check(self.lib.hj3d_split_pairs(self.h, _ptr(pairs), n, _ptr(left), _ptr(right)))

    def gather_u32(self, src, idx, n, dst):
        capi.check(self.lib.hj3d_gather_u32(self.h, _ptr(src), _ptr(idx), n, _ptr(dst)))

    def partition_by_owner(self, tuples, n, ks, num_buckets, n_owners, rowid_base, out):
        counts = (C.c_uint64 * n_owners)()
        capi.check(self.lib.hj3d_partition_by_owner(self.h, _ptr(tuples), n, ks, num_buckets, n_owners, rowid_base,
                                                    _ptr(out), counts))
        return [int(x) for x in counts]

    def gen_column(self, tuples, tuple_bytes, offset, first_row, n, kind, vmax=1, zipf_q=0.0, shift=0, seed=0):
        """device-side generation of one uint32 attribute of a row store (hj3d_gen_column_u32)"""
        capi.check(self.lib.hj3d_gen_column_u32(self.h, _ptr(tuples), tuple_bytes, offset, int(first_row), int(n), kind, int(vmax),
                                                float(zipf_q), int(shift), int(seed)))

    def probe2_unnest2(self, table_s, table_t, tuples, n, ks, flags=F_CHECKSUM, want_triples=False, out=None, out_cap=0):
        """exp4's Ndu probe strand as one device pipeline (hj3d_probe2_unnest2): returns (rc, [4 counter dicts], triples, n_out).
        want_triples: run once count-only to size the result, then materialise (r, s, t) row-id triples in a torch tensor."""
        cnt = (Counters * 4)()
        if want_triples and out is None:
            import torch
            capi.check(self.lib.hj3d_probe2_unnest2(self.h, table_s.h, table_t.h, _ptr(tuples), int(n), ks, 0, None, 0, cnt))
            out_cap = int(cnt[3].out_tuples)
            out = torch.empty((max(out_cap, 1), 3), dtype=torch.int32, device=f"cuda:{self.device}")
        rc = capi.check(self.lib.hj3d_probe2_unnest2(self.h, table_s.h, table_t.h, _ptr(tuples), int(n), ks, flags, _ptr(out),
                                                     int(out_cap), cnt))
        return rc, [c.as_dict() for c in cnt], out, int(cnt[3].out_written)

    def join_host(self, mode, h_build, n_build, ks_build, num_buckets, h_probe, n_probe, ks_probe,
                  flags=0, h_out=None, out_cap=0, want_stats=False):
        """hj3d_join_host on host buffers (numpy arrays / pinned torch CPU tensors / raw addresses)."""
        def hp(a):
            if a is None:
                return None
            if isinstance(a, int):
                return C.c_void_p(a)
            if hasattr(a, "data_ptr"):
                return C.c_void_p(a.data_ptr())
            return a.ctypes.data_as(C.c_void_p)
        pc, uc, st = Counters(), Counters(), Stats()
        rc = capi.check(self.lib.hj3d_join_host(self.h, mode, hp(h_build), n_build, ks_build, num_buckets,
                                                hp(h_probe), n_probe, ks_probe, flags, hp(h_out), out_cap,
                                                C.byref(pc), C.byref(uc), C.byref(st) if want_stats else None))
        return rc, pc.as_dict(), uc.as_dict(), (st.as_dict() if want_stats else None)


class Parts:
    """hj3d_parts: this rank's bucket ranges of a relation after the exchange (coarse-partitioned (key, global row id) records)."""

    def __init__(self, lib, h):
        self.lib, self.h = lib, h

    def info(self):
        n, sent, lo, hi, ov = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_int()
        capi.check(self.lib.hj3d_parts_info(self.h, C.byref(n), C.byref(sent), C.byref(lo), C.byref(hi), C.byref(ov)))
        return {"n_records": int(n.value), "n_sent_remote": int(sent.value), "bucket_lo": int(lo.value), "bucket_hi": int(hi.value),
                "overflow": int(ov.value)}

    def selected(self):
        n = C.c_uint64()
        capi.check(self.lib.hj3d_parts_selected(self.h, C.byref(n)))
        return int(n.value)

    def hot(self):
        """tuples of this rank's slice that stayed local as hot-key tuples (XCHG_HOT)"""
        n = C.c_uint64()
        capi.check(self.lib.hj3d_parts_hot(self.h, C.byref(n)))
        return int(n.value)

    def destroy(self):
        if getattr(self, "h", None):
            self.lib.hj3d_parts_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class Comm:
    """hj3d_comm: one rank of the multi-GPU exchange (include/hj3d.h, csrc/exchange.cu)."""

    def __init__(self, ctx, h, world, rank):
        self.ctx, self.lib, self.h, self.world, self.rank = ctx, ctx.lib, h, world, rank

    @staticmethod
    def unique_id():
        buf = (C.c_ubyte * 128)()
        capi.check(capi.load().hj3d_comm_unique_id(buf))
        return bytes(buf)

    @classmethod
    def create(cls, ctx, world, rank, id128=None):
        h = C.c_void_p()
        buf = (C.c_ubyte * 128).from_buffer_copy(id128) if id128 is not None else None
        capi.check(ctx.lib.hj3d_comm_create(ctx.h, world, rank, buf, C.byref(h)))
        return cls(ctx, h, world, rank)

    @classmethod
    def local(cls, ctxs):
        """all ranks in this process (one context per rank; contexts may share a device)"""
        n = len(ctxs)
        arr = (C.c_void_p * n)(*[c.h for c in ctxs])
        out = (C.c_void_p * n)()
        capi.check(ctxs[0].lib.hj3d_comm_create_local(arr, n, out))
        return [cls(ctxs[r], C.c_void_p(out[r]), n, r) for r in range(n)]

    def set_option(self, opt, v):
        capi.check(self.lib.hj3d_comm_set_option(self.h, opt, int(v)))

    def reserve(self, slot, records, key_bytes=4):
        capi.check(self.lib.hj3d_comm_reserve(self.h, slot, int(records), key_bytes))

    def shard(self, num_buckets):
        lo, hi = C.c_uint64(), C.c_uint64()
        capi.check(self.lib.hj3d_comm_shard(self.h, int(num_buckets), C.byref(lo), C.byref(hi)))
        return int(lo.value), int(hi.value)

    def begin(self, slot, tuples, n, ks, num_buckets, rowid_base, flags=0, selection=None):
        """selection: (attr_offset, op, constant) with op 1 <, 2 <=, 3 >, 4 >=, 5 ==, 6 != on an int32 attribute (fused AlgSelection)"""
        if selection is None:
            capi.check(self.lib.hj3d_exchange_begin(self.h, slot, _ptr(tuples), int(n), ks, int(num_buckets), int(rowid_base), flags))
        else:
            sel = capi.Selection(*selection)
            capi.check(self.lib.hj3d_exchange_begin_select(self.h, slot, _ptr(tuples), int(n), ks, int(num_buckets), int(rowid_base), flags,
                                                           C.byref(sel)))

    def begin_host(self, slot, h_tuples, n, ks, num_buckets, rowid_base, flags=0, selection=None):
        """the slice is a HOST buffer (numpy array / pinned torch CPU tensor / raw address): chunked upload, level 1 per chunk"""
        a = h_tuples
        hp = None if a is None else C.c_void_p(a) if isinstance(a, int) else C.c_void_p(a.data_ptr()) if hasattr(a, "data_ptr") \
            else a.ctypes.data_as(C.c_void_p)
        sel = C.byref(capi.Selection(*selection)) if selection is not None else None
        capi.check(self.lib.hj3d_exchange_begin_host(self.h, slot, hp, int(n), ks, int(num_buckets), int(rowid_base), flags, sel))

    def hot_sample(self, slot, tuples, n, ks):
        """hot-key replication, step 1 (every rank, before begin(.., flags=XCHG_HOT)): sample the relation for hot keys"""
        capi.check(self.lib.hj3d_exchange_hot_sample(self.h, slot, _ptr(tuples), int(n), ks))

    def append(self, slot, tuples, n, rowid_base, flags=0):
        """next chunk of a slice begun with XCHG_MORE; the last chunk comes without the flag"""
        capi.check(self.lib.hj3d_exchange_append(self.h, slot, _ptr(tuples), int(n), int(rowid_base), flags))

    def end(self, slot, tuples, rowid_base, rowid_bound=0):
        """returns (rc, Parts); rc == OVERFLOW: a receive region overflowed somewhere (retry with XCHG_EXACT / more room)"""
        h = C.c_void_p()
        rc = capi.check(self.lib.hj3d_exchange_end(self.h, slot, _ptr(tuples), int(rowid_base), int(rowid_bound), C.byref(h)))
        return rc, Parts(self.lib, h)

    def exchange(self, slot, tuples, n, ks, num_buckets, rowid_base, rowid_bound=0, flags=0):
        self.begin(slot, tuples, n, ks, num_buckets, rowid_base, flags)
        return self.end(slot, tuples, rowid_base, rowid_bound)

    def destroy(self):
        if getattr(self, "h", None):
            self.lib.hj3d_comm_destroy(self.h)
        self.h = None


class Table:
    """hj3d_table: the chaining (HtChaining1) or nested (HtNested1) table of one build operator."""

    def __init__(self, ctx, kind, num_buckets, shard=None):
        self.ctx, self.lib, self.kind = ctx, ctx.lib, kind
        h = C.c_void_p()
        if shard is None:
            capi.check(self.lib.hj3d_table_create(ctx.h, kind, int(num_buckets), C.byref(h)))
        else:
            capi.check(self.lib.hj3d_table_create_shard(ctx.h, kind, int(num_buckets), int(shard[0]), int(shard[1]),
                                                        C.byref(h)))
        self.h = h

    def destroy(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.lib.hj3d_table_destroy(self.ctx.h, self.h)
        self.h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    def build(self, tuples, n, ks):
        capi.check(self.lib.hj3d_table_build(self.ctx.h, self.h, _ptr(tuples), int(n), ks))
        return self

    def clear(self):
        capi.check(self.lib.hj3d_table_clear(self.ctx.h, self.h))

    def set_rowid_bound(self, bound):
        capi.check(self.lib.hj3d_table_set_rowid_bound(self.ctx.h, self.h, int(bound)))
        return self

    def stats(self):
        s = Stats()
        capi.check(self.lib.hj3d_table_stats(self.ctx.h, self.h, C.byref(s)))
        return s.as_dict()

    def size(self):
        n, g = C.c_uint64(), C.c_uint64()
        capi.check(self.lib.hj3d_table_size(self.h, C.byref(n), C.byref(g)))
        return int(n.value), int(g.value)

    def build_parts(self, parts):
        capi.check(self.lib.hj3d_table_build_parts(self.ctx.h, self.h, parts.h))
        return self

    def hot_answers(self, parts, mode):
        """hot-key replication, step 3 (every rank, once this table is built): the owners' answers for the hot keys"""
        capi.check(self.lib.hj3d_parts_hot_answers(self.ctx.h, self.h, parts.h, mode))

    def probe_parts(self, parts, mode, flags=F_CHECKSUM, out=None, out_cap=0):
        """probe with an exchanged relation: mode 0 / 1 chaining (1 = IsBuildKeyUnique), 2 nested, 3 nested + unnest"""
        pc, uc = Counters(), Counters()
        rc = capi.check(self.lib.hj3d_probe_parts(self.ctx.h, self.h, parts.h, mode, flags, _ptr(out), int(out_cap), C.byref(pc), C.byref(uc)))
        return rc, pc.as_dict(), uc.as_dict()

    def probe_chaining(self, tuples, n, ks, unique=False, gather=None, flags=F_CHECKSUM, out=None, out_cap=0):
        c = Counters()
        rc = capi.check(self.lib.hj3d_probe_chaining(self.ctx.h, self.h, _ptr(tuples), int(n), ks, _ptr(gather),
                                                     int(bool(unique)), flags, _ptr(out), int(out_cap), C.byref(c)))
        return rc, c.as_dict()

    def probe_nested(self, tuples, n, ks, gather=None, flags=F_CHECKSUM, out=None, out_cap=0):
        c = Counters()
        rc = capi.check(self.lib.hj3d_probe_nested(self.ctx.h, self.h, _ptr(tuples), int(n), ks, _ptr(gather), flags,
                                                   _ptr(out), int(out_cap), C.byref(c)))
        return rc, c.as_dict()

    def unnest(self, left, gref, n, flags=F_CHECKSUM, out=None, out_cap=0):
        c = Counters()
        rc = capi.check(self.lib.hj3d_unnest(self.ctx.h, self.h, _ptr(left), _ptr(gref), int(n), flags, _ptr(out),
                                             int(out_cap), C.byref(c)))
        return rc, c.as_dict()

    def probe_nested_unnest(self, tuples, n, ks, flags=F_CHECKSUM, out=None, out_cap=0):
        """nested probe directly followed by the unnest, in one kernel: (probe counters, unnest counters)"""
        pc, uc = Counters(), Counters()
        rc = capi.check(self.lib.hj3d_probe_nested_unnest(self.ctx.h, self.h, _ptr(tuples), int(n), ks, flags, _ptr(out),
                                                          int(out_cap), C.byref(pc), C.byref(uc)))
        return rc, pc.as_dict(), uc.as_dict()

    def unnest_pairs(self, nested_pairs, n, flags=F_CHECKSUM, out=None, out_cap=0):
        """deferred unnest of the (left, group ref) pairs exactly as probe_nested wrote them"""
        c = Counters()
        rc = capi.check(self.lib.hj3d_unnest_pairs(self.ctx.h, self.h, _ptr(nested_pairs), int(n), flags, _ptr(out),
                                                   int(out_cap), C.byref(c)))
        return rc, c.as_dict()

    def group_first_row(self, gref, n, out):
        capi.check(self.lib.hj3d_group_first_row(self.ctx.h, self.h, _ptr(gref), int(n), _ptr(out)))
