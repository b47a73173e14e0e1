"""ctypes front-end to the CHECKER libraries (TEST INFRASTRUCTURE, NOT PRODUCT CODE).

  * ``Oracle``  -> oracle/liboracle.so         plain-C restatement (oracle/oracle_join.c)
  * ``Ref``     -> oracle/_ref/libhj3d_ref.so  the unmodified reference templates (oracle/ref_harness.cc)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
NO_ROWID = 0xFFFFFFFF

HASH_MURMUR32, HASH_MURMUR64, HASH_MURMUR64_SEXT32 = 0, 1, 2
CHAINING, NESTED = 0, 1


class KeySpec(C.Structure):
    _fields_ = [("tuple_bytes", C.c_uint32), ("key_offset", C.c_uint32), ("key_bytes", C.c_uint32),
                ("hash_id", C.c_uint32), ("rowid_offset", C.c_uint32)]

    def __init__(self, tuple_bytes, key_offset, key_bytes=4, hash_id=0, rowid_offset=NO_ROWID):
        super().__init__(tuple_bytes, key_offset, key_bytes, hash_id, rowid_offset)


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in
                ("matches", "num_cmps", "out_tuples", "checksum_sum", "checksum_xor", "out_written", "overflow")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in
                ("num_buckets", "num_empty", "num_entries", "num_distinct_keys",
                 "cc_min", "cc_max", "cc_sum", "cc_sumsq", "cc_count",
                 "ccne_min", "ccne_max", "ccne_sum", "ccne_sumsq", "ccne_count",
                 "rsv_main", "rsv_sub", "mem_dir", "mem_main", "mem_sub")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


def build_libs(ref=True):
    """(Re)build the checker libraries with oracle/Makefile."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if ref:
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _tuples(a):
    a = np.ascontiguousarray(a)
    return a, _ptr(a)


class Oracle:
    def __init__(self):
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build_libs(ref=False)
        L = self.lib = C.CDLL(path)
        L.orc_murmur32.restype = C.c_uint32; L.orc_murmur32.argtypes = [C.c_uint32]
        L.orc_murmur64.restype = C.c_uint64; L.orc_murmur64.argtypes = [C.c_uint64]
        L.orc_pair_mix.restype = C.c_uint64; L.orc_pair_mix.argtypes = [C.c_uint32, C.c_uint32]
        L.orc_build.restype = C.c_void_p
        L.orc_build.argtypes = [C.c_int, C.c_void_p, C.c_uint64, KeySpec, C.c_uint64]
        L.orc_table_free.argtypes = [C.c_void_p]
        L.orc_table_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
        L.orc_probe_chaining.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, KeySpec, C.c_void_p, C.c_int,
                                         C.c_void_p, C.c_uint64, C.POINTER(Counters)]
        L.orc_probe_nested.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, KeySpec, C.c_void_p,
                                       C.c_void_p, C.c_uint64, C.POINTER(Counters)]
        L.orc_unnest.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64,
                                 C.POINTER(Counters)]
        L.orc_num_groups.restype = C.c_uint64; L.orc_num_groups.argtypes = [C.c_void_p]
        L.orc_group_len.restype = C.c_uint64; L.orc_group_len.argtypes = [C.c_void_p, C.c_uint32]

    def murmur32(self, x): return int(self.lib.orc_murmur32(x))
    def murmur64(self, x): return int(self.lib.orc_murmur64(x))
    def pair_mix(self, l, r): return int(self.lib.orc_pair_mix(l, r))

    def build(self, kind, tuples, n, ks, num_buckets):
        a, p = _tuples(tuples)
        h = self.lib.orc_build(kind, p, n, ks, num_buckets)
        assert h, "orc_build failed"
        return OracleTable(self, h, a)


class OracleTable:
    def __init__(self, o, h, keep):
        self.o, self.h, self._keep = o, h, keep

    def __del__(self):
        if self.h:
            self.o.lib.orc_table_free(self.h); self.h = None

    def stats(self):
        s = Stats(); self.o.lib.orc_table_stats(self.h, C.byref(s)); return s.as_dict()

    def probe_chaining(self, tuples, n, ks, unique=False, gather=None, materialize=True, cap=None):
        a, p = _tuples(tuples)
        c = Counters()
        out = None
        if materialize:
            if cap is None:
                self.o.lib.orc_probe_chaining(self.h, p, n, ks, _ptr(gather), int(unique), None, 0, C.byref(c))
                cap = c.out_tuples
            out = np.zeros((max(cap, 1), 2), dtype=np.uint32)
        self.o.lib.orc_probe_chaining(self.h, p, n, ks, _ptr(gather), int(unique), _ptr(out), cap or 0, C.byref(c))
        return c.as_dict(), (out[:c.out_written] if out is not None else None)

    def probe_nested(self, tuples, n, ks, gather=None):
        a, p = _tuples(tuples)
        c = Counters()
        out = np.zeros((max(n, 1), 2), dtype=np.uint32)
        self.o.lib.orc_probe_nested(self.h, p, n, ks, _ptr(gather), _ptr(out), n, C.byref(c))
        return c.as_dict(), out[:c.out_written]

    def unnest(self, left, gref):
        left = np.ascontiguousarray(left, dtype=np.uint32); gref = np.ascontiguousarray(gref, dtype=np.uint32)
        n = len(left)
        c = Counters()
        self.o.lib.orc_unnest(self.h, _ptr(left), _ptr(gref), n, None, 0, C.byref(c))
        cap = c.out_tuples
        out = np.zeros((max(cap, 1), 2), dtype=np.uint32)
        self.o.lib.orc_unnest(self.h, _ptr(left), _ptr(gref), n, _ptr(out), cap, C.byref(c))
        return c.as_dict(), out[:c.out_written]

    def num_groups(self): return int(self.o.lib.orc_num_groups(self.h))
    def group_len(self, g): return int(self.o.lib.orc_group_len(self.h, g))


class Ref:
    """The reference's own operators (only available where oracle/_ref/libhj3d_ref.so exists)."""

    PATH = os.path.join(HERE, "_ref", "libhj3d_ref.so")

    @classmethod
    def available(cls):
        return os.path.exists(cls.PATH)

    def __init__(self):
        L = self.lib = C.CDLL(self.PATH)
        L.ref_build.restype = C.c_void_p
        L.ref_build.argtypes = [C.c_int, C.c_void_p, C.c_uint64, KeySpec, C.c_uint64]
        L.ref_build_timed.restype = C.c_void_p
        L.ref_build_timed.argtypes = [C.c_int, C.c_void_p, C.c_uint64, KeySpec, C.c_uint64, C.POINTER(C.c_int64)]
        L.ref_free.argtypes = [C.c_void_p]
        L.ref_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
        L.ref_probe.restype = C.c_int
        L.ref_probe.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, KeySpec, C.c_void_p, C.c_int, C.c_void_p,
                                C.c_uint64, C.POINTER(Counters), C.POINTER(Counters), C.c_int, C.POINTER(C.c_int64)]
        L.ref_gen_exp1.restype = C.c_uint64
        L.ref_gen_exp1.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_gen_exp4.argtypes = [C.c_uint32] * 5 + [C.c_void_p] * 5
        L.ref_exp4_run.argtypes = [C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64,
                                   C.c_uint64, C.c_void_p]

    def build(self, kind, tuples, n, ks, num_buckets, timed=False):
        a, p = _tuples(tuples)
        ns = C.c_int64(0)
        h = (self.lib.ref_build_timed(kind, p, n, ks, num_buckets, C.byref(ns)) if timed
             else self.lib.ref_build(kind, p, n, ks, num_buckets))
        assert h, "ref_build: unsupported tuple layout"
        t = RefTable(self, h)
        t.build_ns = ns.value
        return t

    def gen_exp1(self, log2R, log2S, skew, t):
        """Experiment1::init inputs: returns (R tuples [nR,3] u32, S tuples [nS,3] u32, numDvSa)."""
        nR, nS = 1 << log2R, 1 << log2S
        Rk = np.zeros(nR, np.uint32); Sk = np.zeros(nS, np.uint32); Sa = np.zeros(nS, np.uint32)
        dv = self.lib.ref_gen_exp1(log2R, log2S, int(skew), t, _ptr(Rk), _ptr(Sk), _ptr(Sa))
        R = np.zeros((nR, 3), np.uint32); R[:, 0] = Rk
        S = np.zeros((nS, 3), np.uint32); S[:, 0] = Sk; S[:, 1] = Sa
        return R, S, int(dv)

    def gen_exp4(self, log2R, alpha, mA, beta, mB):
        nR = 1 << log2R
        nF = (nR >> alpha) * mA + (nR >> beta) * mB
        Rk = np.zeros(nR, np.uint32)
        Sk = np.zeros(nF, np.uint32); Sa = np.zeros(nF, np.uint32)
        Tk = np.zeros(nF, np.uint32); Ta = np.zeros(nF, np.uint32)
        self.lib.ref_gen_exp4(log2R, alpha, mA, beta, mB, _ptr(Rk), _ptr(Sk), _ptr(Sa), _ptr(Tk), _ptr(Ta))
        R = np.zeros((nR, 2), np.uint32); R[:, 0] = Rk
        S = np.stack([Sk, Sa], axis=1); T = np.stack([Tk, Ta], axis=1)
        return R, np.ascontiguousarray(S), np.ascontiguousarray(T)

    def exp4_run(self, plan, R, S, T, D):
        out = np.zeros(12, np.uint64)
        R = np.ascontiguousarray(R); S = np.ascontiguousarray(S); T = np.ascontiguousarray(T)
        self.lib.ref_exp4_run(plan, _ptr(R), len(R), _ptr(S), len(S), _ptr(T), len(T), D, _ptr(out))
        names = ["c_probe_RS", "c_probe_RS_cmp", "c_probe_RT", "c_probe_RT_cmp", "c_unnest1", "c_unnest2", "c_top",
                 "checksum_sum", "checksum_xor", "t_build_S_ns", "t_build_T_ns", "t_probe_ns"]
        return {k: int(v) for k, v in zip(names, out)}


class RefTable:
    def __init__(self, r, h):
        self.r, self.h = r, h
        self.build_ns = 0

    def __del__(self):
        if self.h:
            self.r.lib.ref_free(self.h); self.h = None

    def stats(self):
        s = Stats(); self.r.lib.ref_stats(self.h, C.byref(s)); return s.as_dict()

    def probe(self, tuples, n, ks, mode, gather=None, materialize=True, timing_top=False):
        """mode 0 chaining, 1 chaining unique, 2 nested, 3 nested+unnest.
        Returns (probe counters, unnest counters, out pairs or None, probe strand ns)."""
        a, p = _tuples(tuples)
        c, cu = Counters(), Counters()
        ns = C.c_int64(0)
        out, cap = None, 0
        if materialize and not timing_top:
            rc = self.r.lib.ref_probe(self.h, p, n, ks, _ptr(gather), mode, None, 0, C.byref(c), C.byref(cu), 0, None)
            assert rc == 0, "ref_probe: unsupported layout / mode"
            cap = max(c.out_tuples, cu.out_tuples)
            out = np.zeros((max(cap, 1), 2), dtype=np.uint32)
        rc = self.r.lib.ref_probe(self.h, p, n, ks, _ptr(gather), mode, _ptr(out), cap, C.byref(c), C.byref(cu),
                                  int(timing_top), C.byref(ns))
        assert rc == 0, "ref_probe: unsupported layout / mode"
        if out is not None:
            out = out[:max(c.out_written, cu.out_written)]
        return c.as_dict(), cu.as_dict(), out, ns.value
