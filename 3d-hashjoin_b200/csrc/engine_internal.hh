// engine_internal.hh -- objects and host helpers shared by the translation units of libhj3d.so (engine.cu, pipeline.cu,
// exchange.cu, ...).  Not part of the ABI: include/hj3d.h is.
#pragma once

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"

using namespace hj3d;

// ------------------------------------------------------------------------------------ errors
std::string& hj3d_err_slot();   // thread-local error text (engine.cu)
static inline int fail(int code, const std::string& msg) { hj3d_err_slot() = msg; return code; }

#define CUDA_TRY(expr)                                                                             \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return fail(HJ3D_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));              \
  } while (0)

#define HJ_TRY(expr) do { int _rc = (expr); if (_rc < 0) return _rc; } while (0)

// ------------------------------------------------------------------------------------ objects
enum Phase { PH_PARTITION, PH_HIST, PH_SCAN, PH_SCATTER, PH_GROUP, PH_PROBE, PH_UNNEST, PH_PART1, PH_COUNT };

struct hj3d_ctx {
  int          device = 0;
  cudaStream_t stream = nullptr;
  bool         own_stream = false;
  cudaMemPool_t pool = nullptr;
  // options
  int64_t warp_aggregate = 1;
  int64_t partition_bytes = 48ll << 20;
  int64_t partition_window = 8ll << 20;
  int64_t partition_min_probe = 1ll << 20;
  int64_t smem_build = 1;                    // build chaining tables range-by-range in shared memory
  int64_t smem_build_bytes = 64 << 10;       // shared memory budget of one build range
  int64_t smem_probe = 1;                    // probe through shared-memory resident fine partitions
  int64_t smem_slice_bytes = 48 << 10;      // shared memory per block for a fine partition's table slice
  int64_t smem_min_probe = 1ll << 16;       // smaller probe inputs use the global-memory kernels
  int64_t smem_chunk = 1 << 16;             // probe records per work item
  int64_t probe_threads = 256;              // shared-memory probe block size (256 | 512)
  int64_t part_threads = 512;               // partition kernel block size (256 | 512)
  int64_t part_rank_match = 0;              // rank by warp-private histograms + match_any instead of shared atomics
  int64_t part_sample = 1;                  // regions planned from a sample: 0 never, 1 once this ctx has seen an overflow, 2 always
  bool    seen_skew = false;
  int64_t unnest_hot_cap = 1ll << 20;       // entries of the unnest's hot-tuple list before it is re-run with room for all
  int64_t packed_probe = 0;                 // unique chaining probes over compressed slices (probe_packed.cuh) when the table allows it.
                                            // Off by default: measured at 2^27 x 2^30 it needs 8x fewer fine partitions but its probe
                                            // kernel is issue bound at 7.0 ms against 5.4 ms for k_probe_fine
  int64_t packed_min_probe = 1ll << 22;     // smaller probe inputs use the other paths
  int64_t packed_slice_bytes = 100 << 10;   // shared memory of one k_probe_packed block: two blocks (+ static + reserved) per SM's 228 KB
  int64_t host_chunk_bytes = 256ll << 20;   // hj3d_join_host: probe-side upload chunk (0 = one copy)
  int64_t lean_probe = 1;                   // at-most-one-result probes of fine partitions use k_probe_fine (probe_fine.cuh)
  void*   fused_hot_list = nullptr;          // hot list of the last fused probe + unnest call (arena memory)
  uint32_t fused_hot_cap = 0;
  int     smem_optin = 0;                   // cudaDevAttrMaxSharedMemoryPerBlockOptin
  // per-phase events of the last call
  cudaEvent_t ev[PH_COUNT][2];
  bool        ev_used[PH_COUNT];
  cudaEvent_t ev_total[2];
  uint64_t    launches = 0;
  // small device scratch: counters + stats + scalar, and its pinned host mirror
  DevCounters* d_ctr = nullptr;
  DevStats*    d_stats = nullptr;
  unsigned long long* d_scalar = nullptr;   // 4 scalars
  void*        h_pinned = nullptr;          // >= 256 B
  int          sm_count = 148;
  // grow-only workspace for per-call temporaries: bump allocated, reset at the start of every call, so
  // steady-state calls (the repeat loop of the drivers) never touch the device allocator
  struct Chunk { uint8_t* base; size_t cap, used; };
  std::vector<Chunk> arena;
  struct HostJoinBufs { void *b = nullptr, *p = nullptr, *out = nullptr, *nest = nullptr, *l = nullptr, *g = nullptr;
                        size_t cb = 0, cp = 0, cout = 0, cnest = 0, cl = 0, cg = 0; } hj;
  // streamed upload of hj3d_join_host: copy stream, one event per chunk, a one-rank exchange (partition level 1 per chunk)
  cudaStream_t copy_stream = nullptr;
  std::vector<cudaEvent_t> chunk_ev;
  struct hj3d_comm* host_comm = nullptr;
  uint64_t host_comm_records = 0;
  uint32_t host_comm_key_bytes = 0;
};

struct Buf {  // persistent, grow-only device buffer owned by a table
  void* p = nullptr; size_t cap = 0;
};

struct hj3d_table {
  int      kind = 0;
  uint64_t D = 0, blo = 0, bhi = 0;        // global bucket count and owned range
  Dir      dir{};
  bool     built = false;
  int      hash_id = -1;
  uint32_t key_bytes = 0;
  uint64_t n = 0, n_groups = 0;
  uint32_t* off = nullptr;                 // [n_local + 1] bucket run starts
  void*     slots = nullptr;               // Slot<KeyT>[n]     (chaining; temporary for nested)
  uint32_t* goff = nullptr;                // [n_local + 1]     (nested)
  void*     groups = nullptr;              // Group<KeyT>[G]    (nested)
  uint32_t* rows = nullptr;                // [n]               (nested)
  Buf       b_off, b_slots, b_goff, b_groups, b_rows;   // storage behind the pointers above (kept across clear())
  uint32_t  parts = 1, part_width = 0;     // bucket-range partitioning used by the build (1 = none)
  uint32_t  fine_width = 0, fine_parts = 0; // fine partitions whose table slice fits in shared memory
  uint64_t  rowid_bound = 0;               // row ids stored in the table are < rowid_bound (0 = unknown)
  uint64_t  rowid_bound_user = 0;          // the caller's promise for tuples that carry their own row id (hj3d_table_set_rowid_bound)
  bool      pk_ok = false;                 // compressed-slice geometry (probe_packed.cuh) usable
  uint32_t  pk_width = 0, pk_parts = 0, pk_rowid_bits = 0;
  DevStats  hstats{};                      // bucket statistics captured during the build
  bool      have_stats = false;
};

// A relation that is already bucket-range partitioned at the coarse level (what the multi-GPU exchange delivers,
// exchange.cu): records of range p live in n_src segments recs[seg_start[q] .. +seg_count[q]), q = p * n_src + source.
struct PartsView {
  const unsigned long long* seg_start;   // device [n_seg]
  const unsigned long long* seg_count;   // device [n_seg]
  uint32_t n_seg;
  uint32_t range_width;                  // buckets per coarse range (ranges are aligned to the shard's first bucket)
};

struct hj3d_parts {
  void*    recs = nullptr;               // this rank's receive buffer: Slot<KeyT> records (owned by the comm)
  uint32_t key_bytes = 4, hash_id = 0;
  uint64_t D = 0, bucket_lo = 0, bucket_hi = 0;
  uint32_t n_ranges = 0, n_src = 1, range_width = 0;
  uint64_t cap_seg = 0;
  unsigned long long* d_start = nullptr; // device [n_ranges * n_src]
  unsigned long long* d_count = nullptr; // device [n_ranges * n_src]
  uint64_t n_total = 0;                  // records received
  uint64_t n_sent_remote = 0;            // records this rank wrote into other GPUs' buffers
  uint64_t n_local_selected = 0;         // tuples of this rank's slice that passed the fused selection (all of them without one)
  uint64_t rowid_bound = 0;              // global relation size (row ids are global positions)
  int      overflow = 0;                 // a segment exceeded its capacity
  // hot-key probe replication (hot.cuh): tuples of hot keys never left this GPU
  void*    hot_recs = nullptr;           // Slot<KeyT>[hot_count] (owned by the comm)
  uint64_t hot_count = 0;
  const void* hot_table = nullptr;       // device HotTable<KeyT>
  struct hj3d_comm* comm = nullptr;
  int      slot = 0;
  int      hot_mode = -1;                // probe mode the answers were computed for (hj3d_parts_hot_answers)
  hj3d_parts() = default;
  hj3d_parts(const hj3d_parts&) = delete;
  hj3d_parts& operator=(const hj3d_parts&) = delete;
  ~hj3d_parts() { cudaFree(d_start); cudaFree(d_count); }   // the segment tables are the only memory a parts object owns
};

struct PhaseTimer {
  hj3d_ctx* c; Phase p;
  PhaseTimer(hj3d_ctx* c_, Phase p_) : c(c_), p(p_) {
    if (!c->ev_used[p]) { cudaEventRecord(c->ev[p][0], c->stream); c->ev_used[p] = true; }
  }
  ~PhaseTimer() { cudaEventRecord(c->ev[p][1], c->stream); }
};

static inline void begin_call(hj3d_ctx* c) {
  for (int i = 0; i < PH_COUNT; ++i) c->ev_used[i] = false;
  cudaEventRecord(c->ev_total[0], c->stream);
}
static inline void end_call(hj3d_ctx* c) { cudaEventRecord(c->ev_total[1], c->stream); }

static inline int raw_alloc(void** p, size_t bytes) {
  cudaError_t e = cudaMalloc(p, bytes);
  if (e == cudaErrorMemoryAllocation) { cudaGetLastError(); return fail(HJ3D_ERR_NOMEM, "device out of memory"); }
  if (e != cudaSuccess) return fail(HJ3D_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
  return HJ3D_OK;
}

// start of a call: all temporaries of the previous call are dead (every call ends with a stream sync or
// only enqueues work that is ordered before the next call's work on the same stream).  If the previous
// call had to add chunks, merge them into one so the next call of the same shape bump-allocates only.
static inline int arena_reset(hj3d_ctx* c) {
  if (c->arena.size() > 1) {
    size_t total = 0;
    for (auto& k : c->arena) total += k.cap;
    cudaStreamSynchronize(c->stream);
    for (auto& k : c->arena) cudaFree(k.base);
    c->arena.clear();
    void* p = nullptr;
    HJ_TRY(raw_alloc(&p, total));
    c->arena.push_back({(uint8_t*)p, total, 0});
  }
  for (auto& k : c->arena) k.used = 0;
  return HJ3D_OK;
}

template <class T> int dev_alloc(hj3d_ctx* c, T** p, uint64_t count) {
  *p = nullptr;
  if (count == 0) count = 1;
  const size_t bytes = (count * sizeof(T) + 255) & ~(size_t)255;
  for (auto& k : c->arena)
    if (k.cap - k.used >= bytes) { *p = (T*)(k.base + k.used); k.used += bytes; return HJ3D_OK; }
  const size_t cap = bytes > ((size_t)64 << 20) ? bytes : ((size_t)64 << 20);
  void* q = nullptr;
  HJ_TRY(raw_alloc(&q, cap));
  c->arena.push_back({(uint8_t*)q, cap, bytes});
  *p = (T*)q;
  return HJ3D_OK;
}
static inline void dev_free(hj3d_ctx*, void*) {}   // arena memory is reclaimed wholesale by arena_reset

template <class T> int buf_ensure(hj3d_ctx* c, Buf& b, T** p, uint64_t count) {
  if (count == 0) count = 1;
  const size_t bytes = count * sizeof(T) + 64;              // + padding: bulk copies round a slice's size up to 16 bytes
  if (b.cap < bytes) {
    if (b.p) { cudaStreamSynchronize(c->stream); cudaFree(b.p); b.p = nullptr; b.cap = 0; }
    HJ_TRY(raw_alloc(&b.p, bytes));
    b.cap = bytes;
  }
  *p = (T*)b.p;
  return HJ3D_OK;
}
static inline void buf_release(hj3d_ctx* c, Buf& b) {
  if (b.p) { if (c) cudaStreamSynchronize(c->stream); cudaFree(b.p); }
  b.p = nullptr; b.cap = 0;
}

static inline uint32_t blocks_for(uint64_t n, uint32_t per_block) { return (uint32_t)((n + per_block - 1) / per_block); }

static inline int check_keyspec(const hj3d_keyspec& ks) {
  if (ks.hash_id > 2) return fail(HJ3D_ERR_UNSUPPORTED, "unknown hash_id");
  const uint32_t kb = ks.hash_id == HJ3D_HASH_MURMUR64 ? 8 : 4;
  if (ks.key_bytes != kb) return fail(HJ3D_ERR_INVALID, "key_bytes does not match hash_id");
  if (ks.tuple_bytes == 0 || ks.key_offset + kb > ks.tuple_bytes) return fail(HJ3D_ERR_INVALID, "key outside tuple");
  if (ks.key_offset % kb || ks.tuple_bytes % kb)
    return fail(HJ3D_ERR_UNSUPPORTED, "key must be naturally aligned inside the row-store tuple");
  if (ks.rowid_offset != HJ3D_NO_ROWID && (ks.rowid_offset % 4 || ks.rowid_offset + 4 > ks.tuple_bytes))
    return fail(HJ3D_ERR_INVALID, "rowid_offset outside tuple / misaligned");
  return HJ3D_OK;
}

static inline Src make_src(const void* d_tuples, uint64_t n, const hj3d_keyspec& ks, const uint32_t* gather) {
  Src s; s.base = (const uint8_t*)d_tuples; s.gather = gather; s.n = n; s.stride = ks.tuple_bytes;
  s.key_off = ks.key_offset; s.rowid_off = ks.rowid_offset;
  return s;
}

// exchange.cu, used by engine.cu for the hot-key answers (hot.cuh)
extern "C" {
void* hj3d_comm_hot_ans_buffer(hj3d_comm* comm, int slot);
int   hj3d_comm_hot_reduce_begin(hj3d_comm* comm, int slot);
int   hj3d_comm_hot_reduce_end(hj3d_comm* comm, int slot, const void** d_sum);
}
