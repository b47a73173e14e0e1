"""ctypes binding of libhj3d.so (include/hj3d.h).

PyTorch is only used by callers for device memory (``tensor.data_ptr()``), streams and
``torch.distributed``; nothing here routes through a CPU or eager fallback: if the CUDA library is
missing or no sm_100 device is usable, every call raises.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HJ3D_LIB") or os.path.join(HERE, "lib", "libhj3d.so")   # HJ3D_LIB: developer builds (tools/)

OK, OVERFLOW = 0, 1
CHAINING, NESTED = 0, 1
HASH_MURMUR32, HASH_MURMUR64, HASH_MURMUR64_SEXT32 = 0, 1, 2
NO_ROWID = 0xFFFFFFFF
F_CHECKSUM = 1
F_DEVICE_RESULT = 2
OPT_WARP_AGGREGATE, OPT_PARTITION_BYTES, OPT_PARTITION_WINDOW, OPT_PARTITION_MIN_PROBE = 1, 2, 3, 4
OPT_SMEM_PROBE, OPT_SMEM_SLICE_BYTES, OPT_SMEM_MIN_PROBE, OPT_SMEM_CHUNK = 5, 6, 7, 8
OPT_PART_THREADS, OPT_PART_RANK_MATCH, OPT_PROBE_THREADS, OPT_SMEM_BUILD, OPT_SMEM_BUILD_BYTES = 9, 10, 11, 12, 13
OPT_LEAN_PROBE = 18
OPT_UNNEST_HOT_CAP, OPT_PART_SAMPLE = 19, 20
OPT_PACKED_PROBE, OPT_PACKED_MIN_PROBE, OPT_PACKED_SLICE_BYTES = 21, 22, 23
GEN_IOTA, GEN_PERMUTATION, GEN_UNIFORM, GEN_ZIPF, GEN_CONST = 0, 1, 2, 3, 4
XCHG_EXACT, XCHG_MORE, XCHG_HOT, XCHG_ASYNC = 1, 2, 4, 8
OPT_HOST_CHUNK_BYTES = 24
XOPT_TARGET_RANGES, XOPT_MIN_RANGE_WIDTH, XOPT_MAX_RANGE_WIDTH, XOPT_THREADS = 1, 2, 3, 4

# every symbol include/hj3d.h declares (tests check that the library exports all of them)
SYMBOLS = [
    "hj3d_last_error", "hj3d_version", "hj3d_pair_mix",
    "hj3d_ctx_create", "hj3d_ctx_destroy", "hj3d_ctx_set_stream", "hj3d_ctx_set_option", "hj3d_ctx_sync",
    "hj3d_ctx_timings",
    "hj3d_table_create", "hj3d_table_build", "hj3d_table_clear", "hj3d_table_destroy", "hj3d_table_stats",
    "hj3d_table_size", "hj3d_table_set_rowid_bound",
    "hj3d_probe_chaining", "hj3d_probe_nested", "hj3d_unnest", "hj3d_unnest_pairs", "hj3d_probe_nested_unnest", "hj3d_probe2_unnest2", "hj3d_group_first_row", "hj3d_gather_u32",
    "hj3d_split_pairs", "hj3d_join_host",
    "hj3d_comm_unique_id", "hj3d_comm_create", "hj3d_comm_create_local", "hj3d_comm_set_option", "hj3d_comm_destroy",
    "hj3d_comm_reserve", "hj3d_comm_shard", "hj3d_exchange_begin", "hj3d_exchange_append", "hj3d_exchange_begin_host", "hj3d_exchange_hot_sample", "hj3d_parts_hot_answers", "hj3d_parts_hot", "hj3d_exchange_begin_select", "hj3d_parts_selected", "hj3d_exchange_end", "hj3d_parts_info", "hj3d_parts_destroy",
    "hj3d_table_build_parts", "hj3d_probe_parts",
    "hj3d_partition_by_owner", "hj3d_owner_range", "hj3d_table_create_shard", "hj3d_stats_merge",
    "hj3d_gen_column_u32",
    "hj3d_mem_alloc", "hj3d_mem_free", "hj3d_memcpy_h2d", "hj3d_memcpy_d2h", "hj3d_iota_u32",
]


class Hj3dError(RuntimeError):
    pass


class KeySpec(C.Structure):
    _fields_ = [("tuple_bytes", C.c_uint32), ("key_offset", C.c_uint32), ("key_bytes", C.c_uint32),
                ("hash_id", C.c_uint32), ("rowid_offset", C.c_uint32)]

    def __init__(self, tuple_bytes, key_offset, key_bytes=4, hash_id=HASH_MURMUR32, rowid_offset=NO_ROWID):
        super().__init__(tuple_bytes, key_offset, key_bytes, hash_id, rowid_offset)


class Selection(C.Structure):
    _fields_ = [("attr_offset", C.c_uint32), ("op", C.c_uint32), ("constant", C.c_int32)]


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


class Timings(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("partition_ms", "histogram_ms", "scan_ms", "scatter_ms", "group_ms",
                                         "probe_ms", "unnest_ms", "total_ms")] + [("kernel_launches", C.c_uint64),
                                                                                          ("partition_l1_ms", C.c_float), ("reserved_", C.c_float)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


_lib = None


def load():
    """dlopen libhj3d.so and declare the prototypes.  Raises if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Hj3dError(f"{LIB_PATH} is missing: run __graft_entry__.build() / make -C 3d-hashjoin_b200 "
                        "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, u64, u32, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
    L.hj3d_last_error.restype = C.c_char_p
    L.hj3d_version.restype = C.c_char_p
    L.hj3d_pair_mix.restype = u64; L.hj3d_pair_mix.argtypes = [u32, u32]
    L.hj3d_ctx_create.argtypes = [i32, C.POINTER(vp)]
    L.hj3d_ctx_destroy.argtypes = [vp]
    L.hj3d_ctx_set_stream.argtypes = [vp, vp]
    L.hj3d_ctx_set_option.argtypes = [vp, i32, C.c_int64]
    L.hj3d_ctx_sync.argtypes = [vp]
    L.hj3d_ctx_timings.argtypes = [vp, C.POINTER(Timings)]
    L.hj3d_table_create.argtypes = [vp, i32, u64, C.POINTER(vp)]
    L.hj3d_table_create_shard.argtypes = [vp, i32, u64, u64, u64, C.POINTER(vp)]
    L.hj3d_table_build.argtypes = [vp, vp, vp, u64, KeySpec]
    L.hj3d_table_clear.argtypes = [vp, vp]
    L.hj3d_table_destroy.argtypes = [vp, vp]
    L.hj3d_table_stats.argtypes = [vp, vp, C.POINTER(Stats)]
    L.hj3d_table_set_rowid_bound.argtypes = [vp, vp, u64]
    L.hj3d_table_size.argtypes = [vp, C.POINTER(u64), C.POINTER(u64)]
    L.hj3d_probe_chaining.argtypes = [vp, vp, vp, u64, KeySpec, vp, i32, u32, vp, u64, C.POINTER(Counters)]
    L.hj3d_probe_nested.argtypes = [vp, vp, vp, u64, KeySpec, vp, u32, vp, u64, C.POINTER(Counters)]
    L.hj3d_unnest.argtypes = [vp, vp, vp, vp, u64, u32, vp, u64, C.POINTER(Counters)]
    L.hj3d_unnest_pairs.argtypes = [vp, vp, vp, u64, u32, vp, u64, C.POINTER(Counters)]
    L.hj3d_probe_nested_unnest.argtypes = [vp, vp, vp, u64, KeySpec, u32, vp, u64, C.POINTER(Counters), C.POINTER(Counters)]
    L.hj3d_probe2_unnest2.argtypes = [vp, vp, vp, vp, u64, KeySpec, u32, vp, u64, C.POINTER(Counters)]
    L.hj3d_group_first_row.argtypes = [vp, vp, vp, u64, vp]
    L.hj3d_gather_u32.argtypes = [vp, vp, vp, u64, vp]
    L.hj3d_split_pairs.argtypes = [vp, vp, u64, vp, vp]
    L.hj3d_join_host.argtypes = [vp, i32, vp, u64, KeySpec, u64, vp, u64, KeySpec, u32, vp, u64,
                                 C.POINTER(Counters), C.POINTER(Counters), C.POINTER(Stats)]
    L.hj3d_comm_unique_id.argtypes = [vp]
    L.hj3d_comm_create.argtypes = [vp, i32, i32, vp, C.POINTER(vp)]
    L.hj3d_comm_create_local.argtypes = [C.POINTER(vp), i32, C.POINTER(vp)]
    L.hj3d_comm_set_option.argtypes = [vp, i32, C.c_int64]
    L.hj3d_comm_destroy.argtypes = [vp]
    L.hj3d_comm_reserve.argtypes = [vp, i32, u64, u32]
    L.hj3d_comm_shard.argtypes = [vp, u64, C.POINTER(u64), C.POINTER(u64)]
    L.hj3d_exchange_begin.argtypes = [vp, i32, vp, u64, KeySpec, u64, u32, u32]
    L.hj3d_exchange_begin_host.argtypes = [vp, i32, vp, u64, KeySpec, u64, u32, u32, C.POINTER(Selection)]
    L.hj3d_exchange_hot_sample.argtypes = [vp, i32, vp, u64, KeySpec]
    L.hj3d_parts_hot_answers.argtypes = [vp, vp, vp, i32]
    L.hj3d_parts_hot.argtypes = [vp, C.POINTER(u64)]
    L.hj3d_exchange_append.argtypes = [vp, i32, vp, u64, u32, u32]
    L.hj3d_exchange_begin_select.argtypes = [vp, i32, vp, u64, KeySpec, u64, u32, u32, C.POINTER(Selection)]
    L.hj3d_parts_selected.argtypes = [vp, C.POINTER(u64)]
    L.hj3d_exchange_end.argtypes = [vp, i32, vp, u32, u64, C.POINTER(vp)]
    L.hj3d_parts_info.argtypes = [vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64), C.POINTER(u64), C.POINTER(i32)]
    L.hj3d_parts_destroy.argtypes = [vp]
    L.hj3d_table_build_parts.argtypes = [vp, vp, vp]
    L.hj3d_probe_parts.argtypes = [vp, vp, vp, i32, u32, vp, u64, C.POINTER(Counters), C.POINTER(Counters)]
    L.hj3d_partition_by_owner.argtypes = [vp, vp, u64, KeySpec, u64, u32, u32, vp, C.POINTER(u64)]
    L.hj3d_owner_range.argtypes = [u64, u32, u32, C.POINTER(u64), C.POINTER(u64)]
    L.hj3d_stats_merge.argtypes = [C.POINTER(Stats), u32, C.POINTER(Stats)]
    L.hj3d_gen_column_u32.argtypes = [vp, vp, u32, u32, u64, u64, i32, u64, C.c_double, u64, u64]
    L.hj3d_mem_alloc.argtypes = [vp, u64, C.POINTER(vp)]
    L.hj3d_mem_free.argtypes = [vp, vp]
    L.hj3d_memcpy_h2d.argtypes = [vp, vp, vp, u64]
    L.hj3d_memcpy_d2h.argtypes = [vp, vp, vp, u64]
    L.hj3d_iota_u32.argtypes = [vp, vp, u64, u32]
    for name in SYMBOLS:
        fn = getattr(L, name)
        if fn.restype is C.c_int:
            fn.restype = C.c_int
    _lib = L
    return L


def check(rc):
    if rc < 0:
        raise Hj3dError(f"hj3d error {rc}: {load().hj3d_last_error().decode()}")
    return rc
