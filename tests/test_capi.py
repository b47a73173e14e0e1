"""CPU tests of the drop-in boundary: libhj3d.so loads, exports every symbol include/hj3d.h declares,
and fails LOUDLY (no CPU fallback) when no GPU is present."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "hj3d.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hj3d_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg.capi.load()
    declared = header_functions()
    assert declared, "no prototypes parsed from include/hj3d.h"
    for name in declared:
        assert hasattr(lib, name), f"libhj3d.so does not export {name}"
    assert sorted(pkg.capi.SYMBOLS) == declared, "capi.SYMBOLS out of sync with include/hj3d.h"


def test_struct_layouts_match_header(pkg):
    assert C.sizeof(pkg.KeySpec) == 20
    assert C.sizeof(pkg.Counters) == 7 * 8
    assert C.sizeof(pkg.Stats) == 19 * 8
    assert C.sizeof(pkg.Timings) == 8 * 4 + 8 + 2 * 4


def test_pair_mix_agrees_with_oracle(pkg, oracle):
    lib = pkg.capi.load()
    for l, r in [(0, 0), (1, 2), (0xFFFFFFFF, 7), (123456789, 987654321)]:
        assert lib.hj3d_pair_mix(l, r) == oracle.pair_mix(l, r)


def test_owner_ranges_partition_the_directory(pkg):
    lib = pkg.capi.load()
    for D, G in [(1, 1), (1007, 8), (1 << 27, 8), (5, 8), (1000, 3)]:
        prev = 0
        for g in range(G):
            lo, hi = C.c_uint64(), C.c_uint64()
            assert lib.hj3d_owner_range(D, G, g, C.byref(lo), C.byref(hi)) == 0
            assert lo.value == prev and hi.value >= lo.value
            prev = hi.value
        assert prev == D


def test_stats_merge_is_exact(pkg):
    lib = pkg.capi.load()
    parts = (pkg.Stats * 2)()
    a, b = parts[0], parts[1]
    a.num_buckets, a.num_empty, a.num_entries = 4, 1, 9
    a.cc_min, a.cc_max, a.cc_sum, a.cc_sumsq, a.cc_count = 0, 5, 9, 35, 4
    a.ccne_min, a.ccne_max, a.ccne_sum, a.ccne_sumsq, a.ccne_count = 1, 5, 9, 35, 3
    b.num_buckets, b.num_empty, b.num_entries = 2, 2, 0
    b.cc_min, b.cc_max, b.cc_sum, b.cc_sumsq, b.cc_count = 0, 0, 0, 0, 2
    b.ccne_min, b.ccne_count = 2**64 - 1, 0
    out = pkg.Stats()
    assert lib.hj3d_stats_merge(parts, 2, C.byref(out)) == 0
    assert (out.num_buckets, out.num_empty, out.num_entries) == (6, 3, 9)
    assert (out.cc_min, out.cc_max, out.cc_sum, out.cc_sumsq, out.cc_count) == (0, 5, 9, 35, 6)
    assert (out.ccne_min, out.ccne_max, out.ccne_count) == (1, 5, 3)


def test_no_cpu_fallback(pkg):
    """Without a CUDA device the engine must refuse to work instead of computing on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.Hj3dError, match="no CPU fallback|CUDA"):
        pkg.Context(0)


def test_product_does_not_touch_the_oracle():
    """Nothing under 3d-hashjoin_b200/ or include/ may reference oracle/ (the checker)."""
    bad = []
    for base in ("3d-hashjoin_b200", "include"):
        for dp, _, fs in os.walk(os.path.join(ROOT, base)):
            for f in fs:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hh", ".cc", ".cpp")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"pyoracle|liboracle|oracle_join|libhj3d_ref", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad
