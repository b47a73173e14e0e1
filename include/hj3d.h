/*
 * hj3d.h -- C ABI of the B200-native 3D hash-join engine (libhj3d.so).
 *
 * The reference (dflaxx/3d-hashjoin) has no FFI: its boundary is the template
 * protocol of algebra.hh.  This header is the device-side replacement of the
 * hot path *below* that protocol; the C++20 operator templates in
 * 3d-hashjoin_b200/hostcpp/hj3d/ (same names and constructor signatures as
 * algebra.hh) are thin shims over it.  Every entry point cites the reference
 * interface it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types cross this boundary
 *   - pointers named d_* are device pointers (cudaMalloc / torch storage on the ctx's device),
 *     h_* are host pointers, everything else (counters, stats, handles) lives on the host
 *   - return value 0 = HJ3D_OK, < 0 = error (hj3d_last_error() has the text),
 *     HJ3D_OVERFLOW (1) = result capacity too small: counters are exact, pairs beyond
 *     out_cap were dropped
 *   - one CUDA stream per ctx; a ctx and its tables are not thread safe (neither is the reference)
 *   - there is NO CPU fallback: every call fails with HJ3D_ERR_CUDA when no sm_100 device is usable
 */
#ifndef HJ3D_H
#define HJ3D_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HJ3D_OK               0
#define HJ3D_OVERFLOW         1
#define HJ3D_ERR_INVALID     -1
#define HJ3D_ERR_CUDA        -2
#define HJ3D_ERR_UNSUPPORTED -3
#define HJ3D_ERR_NOMEM       -4

/* table kinds: ht_chaining.hh:38-158 (HtChaining1) / ht_nested.hh:71-251 (HtNested1) */
#define HJ3D_CHAINING 0
#define HJ3D_NESTED   1

/* hash functions (util/hasht.hh:52-72) as used by the drivers' Hashfun* functors */
#define HJ3D_HASH_MURMUR32        0 /* uint32 key -> ht::murmur_hash<uint32_t>   (main_experiment1.cc:231,288-301) */
#define HJ3D_HASH_MURMUR64        1 /* uint64 key -> ht::murmur_hash<uint64_t>   (tuple_types.hh:13,17)            */
#define HJ3D_HASH_MURMUR64_SEXT32 2 /* int32 key, sign extended -> murmur_hash<uint64_t> (main_algebra_example.cc:48-66) */

#define HJ3D_NO_ROWID 0xFFFFFFFFu

/* probe / unnest flags */
#define HJ3D_F_CHECKSUM      1u /* also fold every result pair into checksum_sum / checksum_xor */
#define HJ3D_F_DEVICE_RESULT 2u /* hj3d_join_host: materialise the result pairs in device memory even when
                                 h_out_pairs is NULL (they are not copied back) */

/* ctx options (hj3d_ctx_set_option) */
#define HJ3D_OPT_WARP_AGGREGATE   1 /* 0/1: warp-aggregate equal buckets before atomics (default 1)            */
#define HJ3D_OPT_PARTITION_BYTES  2 /* table bytes above which inputs are bucket-range partitioned first;
                                       0 = never partition (default: 48 MiB)                                   */
#define HJ3D_OPT_PARTITION_WINDOW 3 /* target table-window bytes per partition (default 8 MiB)                 */
#define HJ3D_OPT_PARTITION_MIN_PROBE 4 /* probe inputs smaller than this are probed in place (default 2^20)     */
#define HJ3D_OPT_SMEM_PROBE       5 /* 0/1: probe through shared-memory resident fine partitions (default 1)    */
#define HJ3D_OPT_SMEM_SLICE_BYTES 6 /* shared memory per block for a fine partition's table slice (default 48 KiB) */
#define HJ3D_OPT_SMEM_MIN_PROBE   7 /* probe inputs smaller than this use the global-memory kernels (default 2^16) */
#define HJ3D_OPT_SMEM_CHUNK       8 /* probe records per work item of the shared-memory probe (default 2^16)     */
#define HJ3D_OPT_PART_THREADS     9 /* partition kernel block size: 256, 512 or 1024 (default 512)               */
#define HJ3D_OPT_SMEM_BUILD      12 /* 0/1: build chaining tables range-by-range in shared memory (default 1)    */
#define HJ3D_OPT_SMEM_BUILD_BYTES 13 /* shared memory budget of one build range (default 64 KiB)               */
#define HJ3D_OPT_PROBE_THREADS   11 /* shared-memory probe block size: 256 or 512 (default 256)                  */
#define HJ3D_OPT_LEAN_PROBE      18 /* 0/1: unique / nested probes of fine partitions use the lean kernel (default 1)  */
#define HJ3D_OPT_PACKED_PROBE    21 /* 0/1: unique chaining probes of large inputs run over compressed table slices (default 0) */
#define HJ3D_OPT_PACKED_MIN_PROBE 22 /* probe inputs smaller than this use the other paths (default 2^22)                     */
#define HJ3D_OPT_PACKED_SLICE_BYTES 23 /* shared memory of one compressed-slice probe block (default 100 KiB; tests shrink it) */
#define HJ3D_OPT_HOST_CHUNK_BYTES 24 /* hj3d_join_host uploads the probe relation in pieces of this size and partitions each piece
                                      * while the next one is in flight (default 256 MiB; 0 = one copy, then the join)        */
#define HJ3D_OPT_UNNEST_HOT_CAP  19 /* entries of the unnest's hot-tuple list (default 2^20; tests shrink it)           */
#define HJ3D_OPT_PART_SAMPLE     20 /* partition regions sized from a sampled histogram: 0 never, 1 after this ctx has seen
                                       an overflow (default), 2 always                                                */
#define HJ3D_OPT_PART_RANK_MATCH 10 /* 0: rank by shared-memory atomics (default), 1: warp-private histograms + match_any (slower on B200) */

/*
 * Device-describable form of the drivers' hash / equality functors (concepts.hh:22-28,49-56):
 * the join attribute is the key_bytes wide integer at key_offset of each tuple_bytes wide
 * row-store tuple (RelationRS, algebra.hh:98-106); it is hashed with hash_id and compared for
 * equality.  rowid_offset != HJ3D_NO_ROWID: the tuple carries its own uint32 row id (used after
 * the multi-GPU exchange, where tuples are (key, global row id) pairs).
 */
typedef struct {
  uint32_t tuple_bytes;
  uint32_t key_offset;
  uint32_t key_bytes;
  uint32_t hash_id;
  uint32_t rowid_offset;
} hj3d_keyspec;

/* What the probe-side operators expose: AlgBase::count() (algebra.hh:178) and numCmps()
 * (algebra.hh:467,666), plus an order-independent checksum of the result multiset. */
typedef struct {
  uint64_t matches;      /* AlgHashJoinProbe / AlgNestJoinProbe / AlgUnnestHt ::count()          */
  uint64_t num_cmps;     /* ::numCmps()                                                          */
  uint64_t out_tuples;   /* result tuples produced (what AlgTop::count() sees, algebra.hh:223-229) */
  uint64_t checksum_sum; /* sum and xor over hj3d_pair_mix(left, right) of every result pair     */
  uint64_t checksum_xor;
  uint64_t out_written;  /* pairs actually stored to d_out_pairs                                 */
  uint64_t overflow;     /* 1 if out_tuples > out_cap                                            */
} hj3d_counters;

/* HtStatistics (ht_statistics.hh:18-54) as filled by makeStatistics (ht_chaining.hh:260-292,
 * ht_nested.hh:450-482), plus the reservoir / memory getters (ht_chaining.hh:113-117,161-177;
 * ht_nested.hh:192-198,262-284).  Aggregate<size_t> (util/aggregate.hh:27-68) is flattened to
 * min,max,sum,sumsq,count. */
typedef struct {
  uint64_t num_buckets, num_empty, num_entries, num_distinct_keys;
  uint64_t cc_min, cc_max, cc_sum, cc_sumsq, cc_count;           /* _collisionChainLen         */
  uint64_t ccne_min, ccne_max, ccne_sum, ccne_sumsq, ccne_count; /* _collisionChainLenNonempty */
  uint64_t rsv_main, rsv_sub;        /* getRsvSize() | getRsvMainSize(), getRsvSubSize()       */
  uint64_t mem_dir, mem_main, mem_sub; /* memoryConsupmtionDir / Chains | MainChains / SubChains */
} hj3d_stats;

/* per-phase device times of the last call on the ctx (CUDA events on the ctx's stream), ms */
typedef struct {
  float partition_ms, histogram_ms, scan_ms, scatter_ms, group_ms, probe_ms, unnest_ms, total_ms;
  uint64_t kernel_launches; /* kernels launched by the engine since ctx creation */
  float partition_l1_ms;    /* the level-1 k_part_scatter launch alone (first one of the call) */
  float reserved_;
} hj3d_timings;

typedef struct hj3d_ctx   hj3d_ctx;
typedef struct hj3d_table hj3d_table;

const char* hj3d_last_error(void);
const char* hj3d_version(void);
/* checksum contribution of one (left,right) pair; host helper so callers / tests use one definition */
uint64_t    hj3d_pair_mix(uint32_t left, uint32_t right);

int hj3d_ctx_create(int device, hj3d_ctx** out);
int hj3d_ctx_destroy(hj3d_ctx* ctx);
int hj3d_ctx_set_stream(hj3d_ctx* ctx, void* cuda_stream); /* cudaStream_t; NULL = ctx-owned stream */
int hj3d_ctx_set_option(hj3d_ctx* ctx, int option, int64_t value);
int hj3d_ctx_sync(hj3d_ctx* ctx);
int hj3d_ctx_timings(hj3d_ctx* ctx, hj3d_timings* out);

/* ---- device memory helpers (so that host code above the ABI needs no CUDA headers) -----------
 * The reference keeps relations in std::vector<tuple_t> (RelationRS, algebra.hh:98-106); the shims
 * copy them to device memory they own through these calls.  Copies are ordered on the ctx's stream;
 * d2h blocks until the data has arrived. */
int hj3d_mem_alloc(hj3d_ctx* ctx, uint64_t bytes, void** d_out);
int hj3d_mem_free(hj3d_ctx* ctx, void* d_ptr);
int hj3d_memcpy_h2d(hj3d_ctx* ctx, void* d_dst, const void* h_src, uint64_t bytes);
int hj3d_memcpy_d2h(hj3d_ctx* ctx, void* h_dst, const void* d_src, uint64_t bytes);
/* d_dst[i] = first + i  (identity column, e.g. the `left` of a second deferred unnest) */
int hj3d_iota_u32(hj3d_ctx* ctx, uint32_t* d_dst, uint64_t n, uint32_t first);

/* ---- device-side input generation for scale runs (csrc/datagen.cu) --------------------------------------------------
 * Fills the uint32 attribute at `offset` of n row-store tuples (tuple_bytes wide) for global rows first_row .. first_row+n:
 *   HJ3D_GEN_IOTA         value = row id                                  (S.k, main_experiment1.cc:438-441)
 *   HJ3D_GEN_PERMUTATION  a bijection of [0, vmax) applied to the row id  (R.k = std::shuffle(iota), main_experiment1.cc:430-436)
 *   HJ3D_GEN_UNIFORM      uniform on [0, vmax)                            (GenRandIntVec::generate_uni, util/GenRandIntVec.cc:71-98)
 *   HJ3D_GEN_ZIPF         (Zipf(vmax, q) - 1 + shift) % vmax by rejection-inversion (util/zipf_distribution.hh:48-58,
 *                         GenRandIntVec.cc:290-293)
 *   HJ3D_GEN_CONST        value = shift
 * The reference's distributions, NOT its libstdc++ bit stream (parity runs use the reference's own generator).  Every value
 * is a pure function of (seed, row id), so any rank of a multi-GPU run generates any slice of one global relation. */
#define HJ3D_GEN_IOTA 0
#define HJ3D_GEN_PERMUTATION 1
#define HJ3D_GEN_UNIFORM 2
#define HJ3D_GEN_ZIPF 3
#define HJ3D_GEN_CONST 4
int hj3d_gen_column_u32(hj3d_ctx* ctx, void* d_tuples, uint32_t tuple_bytes, uint32_t offset, uint64_t first_row, uint64_t n,
                        int kind, uint64_t vmax, double zipf_q, uint64_t shift, uint64_t seed);

/* ---- build side ----------------------------------------------------------------------------
 * hj3d_table_create  <- HtChaining1 / HtNested1 constructors via AlgHashJoinBuild(aHashDirSize, ..)
 *                       / AlgNestJoinBuild(aHashDirSize, .., ..)   (algebra.hh:566-569, 372-380;
 *                       ht_chaining.hh:106-107, ht_nested.hh:255-259).  The reservoir chunk-size
 *                       arguments have no device equivalent (node storage is prefix-sum allocated).
 * hj3d_table_build   <- the build strand: AlgScan::run -> Alg{Hash,Nest}JoinBuild::step -> insert
 *                       for every tuple of the relation (algebra.hh:259-269,574-577,386-389;
 *                       ht_chaining.hh:181-196; ht_nested.hh:287-311).  Bulk: the table must be empty.
 * hj3d_table_clear   <- clear_ht() / HtX::clear() (algebra.hh:398,583; ht_chaining.hh:250-258;
 *                       ht_nested.hh:438-447)
 * hj3d_table_stats   <- makeStatistics() + getRsv*Size() + memoryConsupmtion*()
 */
int hj3d_table_create(hj3d_ctx* ctx, int kind, uint64_t num_buckets, hj3d_table** out);
int hj3d_table_build(hj3d_ctx* ctx, hj3d_table* t, const void* d_tuples, uint64_t n, hj3d_keyspec ks);
int hj3d_table_clear(hj3d_ctx* ctx, hj3d_table* t);
/* promise for build tuples that carry their own row id (rowid_offset != HJ3D_NO_ROWID, e.g. exchanged (key, global row id)
 * records): every row id is < bound.  Lets the engine pack (hash quotient, row id) into one 32-bit slot in shared memory;
 * 0 = unknown (default).  Takes effect at the next hj3d_table_build. */
int hj3d_table_set_rowid_bound(hj3d_ctx* ctx, hj3d_table* t, uint64_t bound);
int hj3d_table_destroy(hj3d_ctx* ctx, hj3d_table* t);
int hj3d_table_stats(hj3d_ctx* ctx, hj3d_table* t, hj3d_stats* out);
int hj3d_table_size(hj3d_table* t, uint64_t* num_entries, uint64_t* num_groups); /* size(); #MainNodes */

/* ---- probe side ----------------------------------------------------------------------------
 * hj3d_probe_chaining <- the probe strand AlgScan::run -> AlgHashJoinProbe<.., IsBuildKeyUnique>::step
 *                        (algebra.hh:625-659) with HtChaining1::findDirEntryByOther (ht_chaining.hh:236-248).
 *                        Result pair = (probe row id, build row id)  [concatfun_t::eval(l, r)].
 * hj3d_probe_nested   <- AlgNestJoinProbe::step (algebra.hh:435-459) with
 *                        HtNested1::findMainNodeByOther (ht_nested.hh:354-382).
 *                        Result pair = (probe row id, group_ref); group_ref names the MainNode
 *                        (one per distinct build key).  The checksum is taken over
 *                        (probe row id, row id of the MainNode's own tuple) because group_ref
 *                        values depend on the physical layout.
 * hj3d_unnest         <- AlgUnnestHt::step (algebra.hh:510-541): (left, group_ref) ->
 *                        (left, build row id) for the MainNode's tuple and every SubNode.
 *                        `left` is an opaque uint32 carried through (deferred unnesting:
 *                        main_experiment4.cc:846-867 chains two of these).
 *
 * d_gather (nullable): indirection for probe inputs that are intermediates: probe tuple i is
 * d_probe[d_gather[i]] and the emitted left id is i (HashfunNestedRS reaches the key through
 * _r, main_experiment4.cc:413-419).
 * d_out_pairs (nullable = count only): uint32 pairs, order unspecified.
 */
int hj3d_probe_chaining(hj3d_ctx* ctx, hj3d_table* t, const void* d_probe, uint64_t n, hj3d_keyspec ks,
                        const uint32_t* d_gather, int build_key_unique, uint32_t flags,
                        uint32_t* d_out_pairs, uint64_t out_cap, hj3d_counters* out);
int hj3d_probe_nested(hj3d_ctx* ctx, hj3d_table* t, const void* d_probe, uint64_t n, hj3d_keyspec ks,
                      const uint32_t* d_gather, uint32_t flags,
                      uint32_t* d_out_pairs, uint64_t out_cap, hj3d_counters* out);
int hj3d_unnest(hj3d_ctx* ctx, hj3d_table* t, const uint32_t* d_left, const uint32_t* d_group_ref, uint64_t n,
                uint32_t flags, uint32_t* d_out_pairs, uint64_t out_cap, hj3d_counters* out);
/* same, reading the (left, group_ref) PAIRS exactly as hj3d_probe_nested wrote them (no column split in between) */
int hj3d_unnest_pairs(hj3d_ctx* ctx, hj3d_table* t, const uint32_t* d_nested_pairs, uint64_t n,
                      uint32_t flags, uint32_t* d_out_pairs, uint64_t out_cap, hj3d_counters* out);
/* AlgNestJoinProbe directly followed by AlgUnnestHt (main_experiment1.cc runNrs / runNsr: algebra.hh:435-459 feeding
 * algebra.hh:510-541): one call, the nested tuples are expanded where they are found and never written.  probe_out gets
 * the nested probe's count / numCmps, unnest_out the flat result count, checksum and out_written. */
int hj3d_probe_nested_unnest(hj3d_ctx* ctx, hj3d_table* t, const void* d_probe, uint64_t n, hj3d_keyspec ks, uint32_t flags,
                             uint32_t* d_out_pairs, uint64_t out_cap, hj3d_counters* probe_out, hj3d_counters* unnest_out);
/* The deferred-unnesting multi-join pipeline of main_experiment4.cc:846-867 (plan Ndu) as one device pipeline:
 * AlgScan(R) -> AlgNestJoinProbe(table_s) -> AlgNestJoinProbe(table_t, the key reached through r: HashfunNestedRS,
 * main_experiment4.cc:413-419) -> AlgUnnestHt(table_t) -> AlgUnnestHt(table_s) -> AlgTop.  Both tables are probed with
 * the probe tuple's own key; only tuples that find a partner in table_s reach table_t (algebra.hh:447-457).  No
 * intermediate but one (r, S-group, T-group) entry per surviving tuple is materialised.
 * out4[0] / out4[1] = count and numCmps of the two probes, out4[2] = count of the first unnest (c_unnest_S column of
 * the driver's CSV: it unpacks T), out4[3] = count of the second unnest = c_top, with the checksum
 * sum / xor of hj3d_pair_mix((uint32_t)hj3d_pair_mix(r, s), t) over the flat (r, s, t) row-id triples.
 * d_out_triples (nullable = count / checksum only): uint32 triples (probe row, table_s row, table_t row), order unspecified. */
int hj3d_probe2_unnest2(hj3d_ctx* ctx, hj3d_table* table_s, hj3d_table* table_t, const void* d_probe, uint64_t n, hj3d_keyspec ks,
                        uint32_t flags, uint32_t* d_out_triples, uint64_t out_cap, hj3d_counters* out4);
/* first build row id (the MainNode's own tuple) of each group_ref: d_out[i] = data(group d_group_ref[i]) */
int hj3d_group_first_row(hj3d_ctx* ctx, hj3d_table* t, const uint32_t* d_group_ref, uint64_t n, uint32_t* d_out);
/* d_dst[i] = d_src[2*d_idx_pairs_col...]: small column helpers for composing deferred-unnest pipelines */
int hj3d_gather_u32(hj3d_ctx* ctx, const uint32_t* d_src, const uint32_t* d_idx, uint64_t n, uint32_t* d_dst);
int hj3d_split_pairs(hj3d_ctx* ctx, const uint32_t* d_pairs, uint64_t n, uint32_t* d_left, uint32_t* d_right);

/* ---- whole join with HOST buffers (what a caller holding std::vector<tuple_t> relations does) --
 * One call = build strand + probe strand (+ unnest) of one plan, host->device copies of both
 * relations and the device->host copy of the result inside.  mode: 0 chaining, 1 chaining with
 * IsBuildKeyUnique, 2 nested (no unnest), 3 nested + unnest  (plans Crs/CsrUU, Csr, NrsNU, Nrs/Nsr
 * of main_experiment1.cc:624-1285).  h_out_pairs nullable.  stats nullable.
 * The probe relation is uploaded in chunks (HJ3D_OPT_HOST_CHUNK_BYTES) on a second stream (hj3d_exchange_begin_host on a
 * one-rank communicator); the build and the first partition pass of every chunk run under the upload of the following
 * chunks, so the call costs the host-to-device transfer plus the second partition pass and the probe kernel.  Pinned host
 * memory is needed for the overlap.  HJ3D_TRACE_HOST=1 in the environment prints host-side time stamps of the phases. */
int hj3d_join_host(hj3d_ctx* ctx, int mode,
                   const void* h_build, uint64_t n_build, hj3d_keyspec ks_build, uint64_t num_buckets,
                   const void* h_probe, uint64_t n_probe, hj3d_keyspec ks_probe, uint32_t flags,
                   uint32_t* h_out_pairs, uint64_t out_cap,
                   hj3d_counters* probe_out, hj3d_counters* unnest_out, hj3d_stats* stats_out);

/* ---- multi-GPU sharding (no reference equivalent) ---------------------------
 * The join shards by bucket range: owner(tuple) is a function of bucket(tuple), so every bucket (chain, key group) lives
 * on exactly one GPU, counters add, checksums add / xor and HtStatistics merge exactly (hj3d_stats_merge).
 *
 * hj3d_exchange_* is the data plane.  The exchange IS partition level 1 of the engine: each rank partitions its slice of
 * a relation into coarse bucket ranges and the partition kernel stores every range directly into the receive buffer of
 * the GPU that owns it (peer-mapped memory: the stores cross NVLink as the kernel writes them); one small all-gather of
 * the per-(source, range) counts is the only collective and also the barrier.  The result (hj3d_parts) is the
 * coarse-partitioned input of the local join: hj3d_table_build_parts / hj3d_probe_parts continue at partition level 2,
 * with no host round trip and no further pass over the received data in between.
 *
 * Communicators:
 *   one process per GPU : rank 0 calls hj3d_comm_unique_id, the host code distributes the 128 bytes (MPI, torch.distributed,
 *                         a file), every rank calls hj3d_comm_create.  NCCL carries the count all-gather, CUDA IPC maps the
 *                         peers' receive buffers.
 *   one process, N GPUs : hj3d_comm_create_local(ctxs, N, comms) -- what a C++ driver holding all contexts uses.  Call
 *                         hj3d_exchange_begin for every rank first, then hj3d_exchange_end for every rank.
 * slot: 0 or 1 -- two relations (build side, probe side) can be in flight at once.
 * Records carry (key, global row id) with global row id = rowid_base + position in the local slice.
 * flags: HJ3D_XCHG_EXACT = two passes (histogram first, regions packed at exact offsets): for skewed keys whose ranges
 *        overflow the uniform regions (hj3d_exchange_end returns HJ3D_OVERFLOW then; nothing is lost by retrying). */
#define HJ3D_XCHG_EXACT 1u
#define HJ3D_XCHG_MORE  2u /* streamed slice: further chunks follow through hj3d_exchange_append (the last one without this flag).
                            * A chunk is partitioned while the next one is still being uploaded; not with HJ3D_XCHG_EXACT. */
#define HJ3D_XOPT_TARGET_RANGES   1 /* coarse bucket ranges over the whole directory (default 256; 128 with more than one rank) */
#define HJ3D_XOPT_MAX_RANGE_WIDTH 3 /* largest range width in buckets (default 2^21 = 1024 fine partitions of a chaining table on unique keys; use 2^20
                                     * for nested tables): the local join continues at partition level 2 only below 1024 fine partitions per range */
#define HJ3D_XOPT_THREADS         4 /* block size of the exchange's partition kernel: 512, 1024, 0 = 1024 with more than one rank (default) */
#define HJ3D_XOPT_MIN_RANGE_WIDTH 2 /* smallest range width in buckets (default 16384: a multiple of every fine-partition width) */
#define HJ3D_XCHG_HOT   4u /* hot-key probe replication for a skewed PROBE side (Zipf foreign keys): tuples of the (at most 128)
                            * most frequent keys are not sent to the owner of their bucket but stay on the GPU that read them;
                            * once the tables are built every rank learns the owners' answers for the hot keys (one small
                            * all-reduce) and joins its hot tuples locally.  Same results and counters, no owner serialises, and
                            * the remaining tuples fit the uniform regions again (no HJ3D_XCHG_EXACT pass).  Call order:
                            *   hj3d_exchange_hot_sample (every rank) -> hj3d_exchange_begin(.., HJ3D_XCHG_HOT) / _end
                            *   -> build the table -> hj3d_parts_hot_answers (every rank) -> hj3d_probe_parts (mode 0, 1 or 3).
                            * Hot keys with more than 8 build partners are refused (HJ3D_ERR_UNSUPPORTED from hj3d_probe_parts):
                            * the feature is for foreign keys probing a (nearly) unique build side. */
#define HJ3D_XCHG_ASYNC 8u /* run this exchange on a stream of the communicator: work queued on the ctx's stream until
                            * hj3d_exchange_end (building the table from the other relation) overlaps it */
typedef struct hj3d_comm  hj3d_comm;
typedef struct hj3d_parts hj3d_parts;
int hj3d_comm_unique_id(void* id128);
int hj3d_comm_create(hj3d_ctx* ctx, int world, int rank, const void* id128, hj3d_comm** out);
int hj3d_comm_create_local(hj3d_ctx** ctxs, int world, hj3d_comm** out_comms);
int hj3d_comm_set_option(hj3d_comm* comm, int option, int64_t value);
int hj3d_comm_destroy(hj3d_comm* comm);
/* collective: (re)allocate this rank's receive buffer of `slot` for `records` records of key_bytes + 4 bytes and map the peers' */
int hj3d_comm_reserve(hj3d_comm* comm, int slot, uint64_t records, uint32_t key_bytes);
/* bucket range [lo, hi) this rank owns in a num_buckets wide directory: create its table with hj3d_table_create_shard */
int hj3d_comm_shard(hj3d_comm* comm, uint64_t num_buckets, uint64_t* lo, uint64_t* hi);
/* AlgSelection (algebra.hh:279-315) fused into the exchange's load: only tuples whose int32 attribute at attr_offset satisfies
 * `attr <op> constant` take part (op: 1 <, 2 <=, 3 >, 4 >=, 5 ==, 6 !=; 0 = no selection).  hj3d_parts_selected reports how
 * many tuples of this rank's slice passed (AlgSelection::count()). */
typedef struct { uint32_t attr_offset; uint32_t op; int32_t constant; } hj3d_selection;
int hj3d_exchange_begin_select(hj3d_comm* comm, int slot, const void* d_tuples, uint64_t n, hj3d_keyspec ks, uint64_t num_buckets,
                               uint32_t rowid_base, uint32_t flags, const hj3d_selection* selection);
int hj3d_parts_selected(hj3d_parts* parts, uint64_t* n_local_selected);
int hj3d_exchange_begin(hj3d_comm* comm, int slot, const void* d_tuples, uint64_t n, hj3d_keyspec ks, uint64_t num_buckets,
                        uint32_t rowid_base, uint32_t flags);
/* hot-key replication, step 1 (collective): sample the slot's relation; the ranks agree on the hot set on the device */
int hj3d_exchange_hot_sample(hj3d_comm* comm, int slot, const void* d_tuples, uint64_t n, hj3d_keyspec ks);
/* next chunk of the local slice after a hj3d_exchange_begin(.., HJ3D_XCHG_MORE); rowid_base = global row id of its first tuple */
int hj3d_exchange_append(hj3d_comm* comm, int slot, const void* d_tuples, uint64_t n, uint32_t rowid_base, uint32_t flags);
/* The local slice is in HOST memory (pinned, for the overlap): it is uploaded in chunks of HJ3D_OPT_HOST_CHUNK_BYTES on a copy
 * stream and every chunk goes through the exchange (partition level 1, peer stores) as soon as it has landed, on a stream
 * of the communicator -- the ctx's stream is not used before hj3d_exchange_end, so work queued on it meanwhile (building the
 * table from the other relation) runs under the upload.  selection nullable.  No HJ3D_XCHG_EXACT / _MORE.  Pass
 * d_tuples = NULL to hj3d_exchange_end. */
int hj3d_exchange_begin_host(hj3d_comm* comm, int slot, const void* h_tuples, uint64_t n, hj3d_keyspec ks, uint64_t num_buckets,
                             uint32_t rowid_base, uint32_t flags, const hj3d_selection* selection);
/* d_tuples / rowid_base: the same as in _begin (the exact mode reads the slice a second time); rowid_bound: global relation
 * size (0 = unknown).  Returns HJ3D_OVERFLOW if a region overflowed anywhere (every rank returns it). */
int hj3d_exchange_end(hj3d_comm* comm, int slot, const void* d_tuples, uint32_t rowid_base, uint64_t rowid_bound, hj3d_parts** out);
int hj3d_parts_info(hj3d_parts* parts, uint64_t* n_records, uint64_t* n_sent_remote, uint64_t* bucket_lo, uint64_t* bucket_hi, int* overflow);
int hj3d_parts_destroy(hj3d_parts* parts);
/* hj3d_table_build / hj3d_probe_chaining (mode 0, 1 = IsBuildKeyUnique) / hj3d_probe_nested (2) / hj3d_probe_nested_unnest (3)
 * on an exchanged relation; the table must be the shard hj3d_comm_shard names.  `left` ids are global row ids. */
int hj3d_table_build_parts(hj3d_ctx* ctx, hj3d_table* t, hj3d_parts* parts);
int hj3d_probe_parts(hj3d_ctx* ctx, hj3d_table* t, hj3d_parts* parts, int mode, uint32_t flags, uint32_t* d_out_pairs, uint64_t out_cap,
                     hj3d_counters* probe_out, hj3d_counters* unnest_out);

/* hot-key replication, step 3 (collective, after the table is built): what a probe with each hot key finds, from its owner */
int hj3d_parts_hot_answers(hj3d_ctx* ctx, hj3d_table* t, hj3d_parts* probe_parts, int mode);
/* number of tuples of this rank's slice that stayed local as hot-key tuples */
int hj3d_parts_hot(hj3d_parts* parts, uint64_t* n_hot_records);

/* Lower-level pieces (kept: callers that move the records themselves, e.g. over another transport).
 * owner(tuple) = bucket(tuple) / ceil(num_buckets / n_owners): contiguous bucket ranges.  Writes (key, global row id) pairs
 * grouped by owner into d_out (key_bytes + 4 bytes each, 8 or 16 byte records) and the per-owner counts
 * (host array of n_owners).  rowid_base is added to the local row position. */
int hj3d_partition_by_owner(hj3d_ctx* ctx, const void* d_tuples, uint64_t n, hj3d_keyspec ks,
                            uint64_t num_buckets, uint32_t n_owners, uint32_t rowid_base,
                            void* d_out, uint64_t* h_counts);
/* bucket range [lo, hi) owned by `owner` */
int hj3d_owner_range(uint64_t num_buckets, uint32_t n_owners, uint32_t owner, uint64_t* lo, uint64_t* hi);
/* a shard table holds only buckets [bucket_lo, bucket_hi) of a num_buckets wide directory */
int hj3d_table_create_shard(hj3d_ctx* ctx, int kind, uint64_t num_buckets, uint64_t bucket_lo, uint64_t bucket_hi,
                            hj3d_table** out);
/* merge shard statistics exactly (Aggregate is mergeable: min,max,sum,sumsq,count) */
int hj3d_stats_merge(const hj3d_stats* parts, uint32_t n, hj3d_stats* out);

#ifdef __cplusplus
}
#endif
#endif /* HJ3D_H */
