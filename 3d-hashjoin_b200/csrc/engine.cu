// engine.cu -- host orchestration + C ABI (include/hj3d.h) of the hj3d engine.
//
// No CPU fallback: every entry point needs a usable sm_100 device and fails with HJ3D_ERR_CUDA
// otherwise.  Memory comes from the device's stream-ordered pool (release threshold = max, so
// steady-state calls do not hit the OS allocator); all kernels run on the ctx's stream.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "build.cuh"
#include "build_nested_fine.cuh"
#include <chrono>

#include "engine_internal.hh"
#include "hot.cuh"
#include "partition.cuh"
#include "probe.cuh"
#include "probe_smem.cuh"
#include "probe_fine.cuh"
#include "probe_packed.cuh"   // PackCfg + launch_probe_packed (kernels are instantiated in probe_packed.cu)
#include "unnest.cuh"
#include "probe_unnest.cuh"
#include "scan.cuh"

std::string& hj3d_err_slot() { static thread_local std::string e; return e; }

namespace {

// ---- scan drivers ---------------------------------------------------------------------------------
template <class T, bool STATS, class Loader, class Storer>
int run_scan(hj3d_ctx* c, Loader load, Storer store, uint64_t n, DevStats* d_stats, T* d_total) {
  const uint32_t nb = blocks_for(n, kScanTile);
  if (nb == 0) {                                   // empty input (e.g. a shard that owns no bucket): nothing to scan
    if (d_total) CUDA_TRY(cudaMemsetAsync(d_total, 0, sizeof(T), c->stream));
    return HJ3D_OK;
  }
  T* sums = nullptr;
  HJ_TRY(dev_alloc(c, &sums, nb));
  k_scan_reduce<T, Loader, STATS><<<nb, kScanThreads, 0, c->stream>>>(load, n, sums, d_stats);
  k_scan_blocksums<T><<<1, 1024, 0, c->stream>>>(sums, nb, d_total);
  k_scan_apply<T, Loader, Storer><<<nb, kScanThreads, 0, c->stream>>>(load, store, n, sums);
  c->launches += 3;
  dev_free(c, sums);
  CUDA_TRY(cudaGetLastError());
  return HJ3D_OK;
}

struct LoadU32 { const uint32_t* p; __device__ uint32_t operator()(uint64_t i) const { return p[i]; } };
struct StoreInclU32 { uint32_t* p; __device__ void operator()(uint64_t i, uint32_t ex, uint32_t v) const { p[i] = ex + v; } };

struct LoadCells {  // packed (claimed ? 1 : 0, gcnt); element n is the sentinel
  const uint32_t* cell; const uint32_t* gcnt; uint64_t n;
  __device__ unsigned long long operator()(uint64_t i) const {
    if (i >= n) return 0ull;
    return ((unsigned long long)(cell[i] != kEmpty32) << 32) | (unsigned long long)gcnt[i];
  }
};
struct StoreCells {
  uint32_t* gidx; uint32_t* gstart;
  __device__ void operator()(uint64_t i, unsigned long long ex, unsigned long long) const {
    gidx[i] = (uint32_t)(ex >> 32); gstart[i] = (uint32_t)ex;
  }
};
struct LoadDiff { const uint32_t* p; __device__ uint32_t operator()(uint64_t i) const { return p[i + 1] - p[i]; } };
struct StoreNone { __device__ void operator()(uint64_t, uint32_t, uint32_t) const {} };

struct LoadU64 { const unsigned long long* p; __device__ unsigned long long operator()(uint64_t i) const { return p[i]; } };
struct StoreExU64 { unsigned long long* p; __device__ void operator()(uint64_t i, unsigned long long ex, unsigned long long) const { p[i] = ex; } };

// The fine build kernels commit their bucket statistics once per block; with 10^5..10^6 blocks the atomics on ONE
// DevStats serialise in the L2 (measured: most of a 47 ms nested build).  Blocks therefore spread over kStatsCopies
// replicas (block id & mask) that one tiny kernel folds into replica 0 afterwards.
constexpr int kStatsCopies = 64;
__global__ void k_stats_fold(DevStats* s, int copies) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  DevStats r = s[0];
  for (int i = 1; i < copies; ++i) {
    const DevStats& o = s[i];
    r.all.mn = o.all.mn < r.all.mn ? o.all.mn : r.all.mn; r.all.mx = o.all.mx > r.all.mx ? o.all.mx : r.all.mx;
    r.all.sum += o.all.sum; r.all.sumsq += o.all.sumsq; r.all.cnt += o.all.cnt;
    r.nonempty.mn = o.nonempty.mn < r.nonempty.mn ? o.nonempty.mn : r.nonempty.mn;
    r.nonempty.mx = o.nonempty.mx > r.nonempty.mx ? o.nonempty.mx : r.nonempty.mx;
    r.nonempty.sum += o.nonempty.sum; r.nonempty.sumsq += o.nonempty.sumsq; r.nonempty.cnt += o.nonempty.cnt;
    r.empty += o.empty;
  }
  s[0] = r;
}

void init_dev_stats_host(DevStats& s) {
  s.all = DevAgg{~0ull, 0, 0, 0, 0}; s.nonempty = DevAgg{~0ull, 0, 0, 0, 0}; s.empty = 0;
}

int init_dev_stats_copies(hj3d_ctx* c) {
  static DevStats h[kStatsCopies];
  for (int i = 0; i < kStatsCopies; ++i) init_dev_stats_host(h[i]);
  CUDA_TRY(cudaMemcpyAsync(c->d_stats, h, sizeof(h), cudaMemcpyHostToDevice, c->stream));
  return HJ3D_OK;
}

// clear(): the table becomes empty but keeps its device buffers for the next build of the repeat loop
void clear_table(hj3d_table* t) {
  t->off = nullptr; t->slots = nullptr; t->goff = nullptr; t->groups = nullptr; t->rows = nullptr;
  t->built = false; t->n = 0; t->n_groups = 0; t->have_stats = false; t->pk_ok = false; t->parts = 1; t->part_width = 0; t->fine_width = 0; t->fine_parts = 0;
}
void free_table_arrays(hj3d_ctx* c, hj3d_table* t) {
  clear_table(t);
  buf_release(c, t->b_off); buf_release(c, t->b_slots); buf_release(c, t->b_goff); buf_release(c, t->b_groups);
  buf_release(c, t->b_rows);
}

// ---- bucket-range partitioning inside one GPU ---------------------------------------------------------
// The records of partition p live in recs[part_start[p] .. part_start[p] + counts[p]) (regions may have
// gaps); tile maps translate block ids of the build / probe kernels to record ranges.
template <class KeyT> struct Partitioned {
  Slot<KeyT>* recs = nullptr;
  unsigned long long* part_start = nullptr;   // device [P]
  unsigned long long* counts = nullptr;       // device [P]
  uint32_t P = 0;
  uint64_t n_kept = 0;                        // records inside the (shard) directory
  bool     fallback = false;
};

inline uint32_t choose_parts(hj3d_ctx* c, uint64_t table_bytes, uint64_t n_local) {
  if (c->partition_bytes <= 0 || (int64_t)table_bytes <= c->partition_bytes) return 1;
  uint64_t want = (table_bytes + c->partition_window - 1) / c->partition_window;
  uint32_t P = 1;
  while (P < want && P < (uint32_t)kMaxParts) P <<= 1;
  while (P > 1 && (uint64_t)P > n_local) P >>= 1;
  return P;
}

// Regions sized from a sampled partition histogram (partition.cuh: k_part_sample / k_plan_caps): fills part_start[P + 1]
// and returns the total capacity.  `tilemap` == nullptr: plain input of n_tiles tiles of `tile` records.
template <int HASH>
int plan_regions(hj3d_ctx* c, Src src, bool recs, const uint2* tilemap, uint32_t tile, uint32_t n_tiles, uint32_t stride, Dir d, PartFn pf,
                 uint32_t P, uint32_t fan, unsigned long long n_total, unsigned long long cap_uniform,
                 unsigned long long* part_start /* [P + 1] */, unsigned long long* total_out) {
  unsigned long long *counts_s = nullptr, *caps = nullptr;
  HJ_TRY(dev_alloc(c, &counts_s, (uint64_t)P + 1));
  HJ_TRY(dev_alloc(c, &caps, (uint64_t)P + 1));
  CUDA_TRY(cudaMemsetAsync(counts_s, 0, ((size_t)P + 1) * 8, c->stream));
  const uint32_t nbs = (n_tiles + stride - 1) / stride;
  if (nbs) {
    if (recs) k_part_sample<HASH, true><<<nbs, 256, 0, c->stream>>>(src, tilemap, tile, stride, n_tiles, d, pf, P, fan, counts_s);
    else      k_part_sample<HASH, false><<<nbs, 256, 0, c->stream>>>(src, tilemap, tile, stride, n_tiles, d, pf, P, fan, counts_s);
  }
  k_plan_caps<<<blocks_for((uint64_t)P + 1, 256), 256, 0, c->stream>>>(counts_s, P, n_total, cap_uniform, caps);
  c->launches += 2;
  HJ_TRY((run_scan<unsigned long long, false>(c, LoadU64{caps}, StoreExU64{part_start}, (uint64_t)P + 1, (DevStats*)nullptr, c->d_scalar + 3)));
  unsigned long long* h = (unsigned long long*)c->h_pinned;
  CUDA_TRY(cudaMemcpyAsync(h, c->d_scalar + 3, 8, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  *total_out = h[0];
  return HJ3D_OK;
}

template <int HASH, bool LEFTID>
int partition_local(hj3d_ctx* c, Src src, Dir d, uint32_t P, uint32_t width, uint32_t rowid_base,
                    Partitioned<typename HashT<HASH>::key_t>* out) {
  using KeyT = typename HashT<HASH>::key_t;
  PhaseTimer pt(c, PH_PARTITION);
  const uint64_t n = src.n;
  const PartFn pf = make_partfn(width, d.lo, d.n_local);
  const unsigned long long cap = (n / P + n / (32ull * P) + 8192) & ~1ull; // expected size + ~3% + constant slack; even: 16-byte aligned regions
  out->P = P;
  const int kTile = (int)c->part_threads * PartCfg<KeyT>::kItems;
  const bool recs = !src.gather && src.stride == sizeof(Slot<KeyT>) && src.key_off == 0 && src.rowid_off == sizeof(KeyT) &&
                    ((uintptr_t)src.base % sizeof(Slot<KeyT>)) == 0 && rowid_base == 0;
  const uint32_t nb = blocks_for(n, kTile);
  // skewed keys seen before (or forced): plan the regions from a sample instead of assuming equal shares
  bool planned = P > 1 && nb > 0 && (c->part_sample == 2 || (c->part_sample == 1 && c->seen_skew));
  HJ_TRY(dev_alloc(c, &out->part_start, (uint64_t)P + 1));
  HJ_TRY(dev_alloc(c, &out->counts, P));
  unsigned long long total = (unsigned long long)P * cap;
  if (planned) {
    const uint32_t stride = nb >= 4096 ? 32u : (nb >= 256 ? 4u : 1u);
    HJ_TRY((plan_regions<HASH>(c, src, recs, nullptr, (uint32_t)kTile, nb, stride, d, pf, P, P, n, cap, out->part_start, &total)));
    // tile maps address records with 32 bits: planned regions (each at least the uniform share) can add up to ~2n
    if (total >= 0xFFFFFFF0ull) { planned = false; total = (unsigned long long)P * cap; }
  }
  if (!planned) {
    k_part_fixed_starts<<<blocks_for((uint64_t)P + 1, 256), 256, 0, c->stream>>>(P + 1, cap, out->part_start);
  }
  HJ_TRY(dev_alloc(c, &out->recs, total));
  CUDA_TRY(cudaMemsetAsync(out->counts, 0, (size_t)P * 8, c->stream));
  auto launch = [&](unsigned long long cap_, unsigned long long* cursor_) -> cudaError_t {
    return launch_part_scatter<HASH, LEFTID>(c->stream, recs, (int)c->part_threads, c->part_rank_match != 0, nb, src, nullptr, d, pf,
                                             P, P, rowid_base, cap_, out->part_start, cursor_, out->recs);
  };
  { PhaseTimer p1(c, PH_PART1); CUDA_TRY(launch(planned ? 0ull : cap, out->counts)); }   // the level-1 scatter kernel alone (bench.py's roofline)
  c->launches += nb ? 2 : 1;
  unsigned long long* h = (unsigned long long*)c->h_pinned;               // P <= 1024 -> 8 KB counts + 8 KB starts
  CUDA_TRY(cudaMemcpyAsync(h, out->counts, (size_t)P * 8, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaMemcpyAsync(h + P, out->part_start, ((size_t)P + 1) * 8, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  bool overflow = false; uint64_t kept = 0;
  for (uint32_t p = 0; p < P; ++p) { kept += h[p]; overflow |= h[p] > h[P + p + 1] - h[P + p]; }
  out->n_kept = kept;
  if (overflow) {
    // skewed partition sizes: the cursors hold the exact histogram, redo the scatter into exact regions
    out->fallback = true;
    c->seen_skew = true;
    unsigned long long* counts2 = nullptr;
    HJ_TRY(dev_alloc(c, &counts2, P));
    k_part_prefix<<<1, 32, 0, c->stream>>>(out->counts, P, out->part_start);
    CUDA_TRY(cudaMemsetAsync(counts2, 0, (size_t)P * 8, c->stream));
    CUDA_TRY(launch(~0ull, counts2));
    c->launches += 2;
  }
  CUDA_TRY(cudaGetLastError());
  return HJ3D_OK;
}

template <class KeyT>
int make_tilemap(hj3d_ctx* c, const Partitioned<KeyT>& pr, uint32_t tile, uint2** tilemap, uint32_t* n_tiles,
                 uint32_t** tile_part = nullptr) {
  uint32_t* tile_prefix = nullptr;
  HJ_TRY(dev_alloc(c, &tile_prefix, (uint64_t)pr.P + 1));
  k_tile_prefix<<<1, 1024, 0, c->stream>>>(pr.counts, pr.P, tile, tile_prefix);
  uint32_t* h = (uint32_t*)c->h_pinned;
  CUDA_TRY(cudaMemcpyAsync(h, tile_prefix + pr.P, 4, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  *n_tiles = *h;
  HJ_TRY(dev_alloc(c, tilemap, (uint64_t)*n_tiles));
  if (tile_part) HJ_TRY(dev_alloc(c, tile_part, (uint64_t)*n_tiles));
  k_make_tilemap<<<pr.P, 256, 0, c->stream>>>(pr.part_start, pr.counts, tile_prefix, tile, *tilemap, tile_part ? *tile_part : nullptr);
  c->launches += 2;
  CUDA_TRY(cudaGetLastError());
  return HJ3D_OK;
}

template <class KeyT> Src records_src(const Partitioned<KeyT>& pr) {
  Src s; s.base = (const uint8_t*)pr.recs; s.gather = nullptr; s.n = 0; s.stride = sizeof(Slot<KeyT>);
  s.key_off = 0; s.rowid_off = sizeof(KeyT);
  return s;
}

// gather the segments of a pre-partitioned relation into one contiguous record array (general paths need a plain input)
template <class KeyT>
__global__ void k_compact_segments(const Slot<KeyT>* __restrict__ recs, const unsigned long long* __restrict__ seg_start,
                                   const unsigned long long* __restrict__ seg_count, const unsigned long long* __restrict__ dst_start,
                                   Slot<KeyT>* __restrict__ out) {
  const unsigned long long n = seg_count[blockIdx.x], s0 = seg_start[blockIdx.x], d0 = dst_start[blockIdx.x];
  for (unsigned long long i = (unsigned long long)blockIdx.y * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.y * blockDim.x)
    out[d0 + i] = recs[s0 + i];
}
template <class KeyT>
int make_contiguous(hj3d_ctx* c, Src* src, const PartsView* pre) {
  Slot<KeyT>* out = nullptr; unsigned long long* dst = nullptr;
  HJ_TRY(dev_alloc(c, &out, src->n));
  HJ_TRY(dev_alloc(c, &dst, (uint64_t)pre->n_seg + 1));
  HJ_TRY((run_scan<unsigned long long, false>(c, LoadU64{pre->seg_count}, StoreExU64{dst}, pre->n_seg, (DevStats*)nullptr, (unsigned long long*)nullptr)));
  if (pre->n_seg && src->n) {
    k_compact_segments<KeyT><<<dim3(pre->n_seg, 16), 256, 0, c->stream>>>((const Slot<KeyT>*)src->base, pre->seg_start, pre->seg_count, dst, out);
    ++c->launches;
  }
  CUDA_TRY(cudaGetLastError());
  src->base = (const uint8_t*)out;
  return HJ3D_OK;
}

// One refinement pass: every partition of `in` (bucket ranges of `fan` x Wout buckets, or segments of such ranges) is split
// into `fan` consecutive ranges of Wout buckets; `out` gets n_out = (#ranges of in) x fan partitions.  F_real = how many of
// them can hold records (sizes the fixed regions).  The ids stored in the records are final (left ids / row ids): the pass
// just carries them (LEFTID + RECS).
template <int HASH>
int refine_partitions(hj3d_ctx* c, const Partitioned<typename HashT<HASH>::key_t>& in, Dir dir, uint32_t fan, uint32_t Wout, uint32_t n_out,
                      uint64_t F_real, uint64_t n, Partitioned<typename HashT<HASH>::key_t>* out_) {
  using KeyT = typename HashT<HASH>::key_t;
  Partitioned<KeyT>& fine = *out_;
  uint2* tm = nullptr; uint32_t n_tiles = 0;
  const int kTile = (int)c->part_threads * PartCfg<KeyT>::kItems;
  HJ_TRY(make_tilemap(c, in, kTile, &tm, &n_tiles));
  PhaseTimer pt(c, PH_PARTITION);
  const uint32_t Fall = n_out;                    // ids past the directory stay empty
  const unsigned long long cap2 = (n / F_real + n / (16ull * F_real) + 2048) & ~1ull;
  fine.P = Fall;
  const PartFn pf = make_partfn(Wout, dir.lo, dir.n_local);
  Src rs = records_src(in);
  bool planned = n_tiles > 0 && (c->part_sample == 2 || (c->part_sample == 1 && c->seen_skew));
  HJ_TRY(dev_alloc(c, &fine.part_start, (uint64_t)Fall + 1));
  HJ_TRY(dev_alloc(c, &fine.counts, Fall));
  unsigned long long total2 = (unsigned long long)Fall * cap2;
  if (planned) {
    const uint32_t stride = n_tiles >= 4096 ? 8u : (n_tiles >= 256 ? 2u : 1u);
    HJ_TRY((plan_regions<HASH>(c, rs, true, tm, (uint32_t)kTile, n_tiles, stride, dir, pf, Fall, fan, in.n_kept, cap2,
                               fine.part_start, &total2)));
    if (total2 >= 0xFFFFFFF0ull) { planned = false; total2 = (unsigned long long)Fall * cap2; }
  }
  if (!planned) {
    k_fixed_starts_u64<<<blocks_for((uint64_t)Fall + 1, 256), 256, 0, c->stream>>>(Fall + 1, cap2, fine.part_start);
  }
  HJ_TRY(dev_alloc(c, &fine.recs, total2));
  CUDA_TRY(cudaMemsetAsync(fine.counts, 0, (size_t)Fall * 8, c->stream));
  CUDA_TRY((launch_part_scatter<HASH, true>(c->stream, true, (int)c->part_threads, c->part_rank_match != 0, n_tiles, rs, tm, dir, pf,
                                            Fall, fan, 0, planned ? 0ull : cap2, fine.part_start, fine.counts, fine.recs)));
  unsigned long long* d_mx = c->d_scalar;
  CUDA_TRY(cudaMemsetAsync(d_mx, 0, 16, c->stream));
  k_part_overflow<<<64, 256, 0, c->stream>>>(fine.counts, fine.part_start, Fall, d_mx);
  c->launches += 3;
  unsigned long long* h = (unsigned long long*)c->h_pinned;
  CUDA_TRY(cudaMemcpyAsync(h, d_mx, 16, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  fine.n_kept = h[1];
  if (h[0]) {                                    // skew: exact regions from the now known histogram
    fine.fallback = true;
    c->seen_skew = true;
    unsigned long long* counts2 = nullptr;
    HJ_TRY(dev_alloc(c, &counts2, Fall));
    HJ_TRY((run_scan<unsigned long long, false>(c, LoadU64{fine.counts}, StoreExU64{fine.part_start}, Fall, (DevStats*)nullptr, (unsigned long long*)nullptr)));
    CUDA_TRY(cudaMemsetAsync(counts2, 0, (size_t)Fall * 8, c->stream));
    CUDA_TRY((launch_part_scatter<HASH, true>(c->stream, true, (int)c->part_threads, c->part_rank_match != 0, n_tiles, rs, tm, dir, pf,
                                              Fall, fan, 0, ~0ull, fine.part_start, counts2, fine.recs)));
    ++c->launches;
  }
  CUDA_TRY(cudaGetLastError());
  return HJ3D_OK;
}

// Partition `src` into the F fine bucket ranges of width Wf (one level if F <= 1024, else coarse
// partitions of P2 consecutive fine ones first).  *ok = false: not representable (caller falls back).
// `pre`: level 1 happened elsewhere (the exchange) and the coarse ranges are given as segments; a range that holds more
// than 1024 fine partitions (heavily duplicated keys: narrow fine partitions) is refined in two passes.
template <int HASH, bool LEFTID>
int partition_fine(hj3d_ctx* c, Src src, Dir dir, uint32_t Wf, uint32_t F,
                   Partitioned<typename HashT<HASH>::key_t>* fine_out, bool* ok, const PartsView* pre = nullptr) {
  using KeyT = typename HashT<HASH>::key_t;
  const uint64_t n = src.n;
  const uint32_t nl = dir.n_local;
  *ok = true;
  if (!pre && F <= (uint32_t)kMaxParts) {
    HJ_TRY((partition_local<HASH, LEFTID>(c, src, dir, F, Wf, 0, fine_out)));
    return HJ3D_OK;
  }
  // two levels: coarse partitions of P2 consecutive fine partitions, then fine inside each coarse region
  uint32_t P2 = 32;
  uint64_t wc;
  Partitioned<KeyT> coarse;
  uint32_t P1;
  if (pre) {
    wc = pre->range_width;
    if (wc == 0 || wc % Wf != 0 || wc / Wf > (uint64_t)kMaxParts * kMaxParts) { *ok = false; return HJ3D_OK; }
    P1 = (uint32_t)(((uint64_t)nl + wc - 1) / wc);
    if ((uint64_t)P1 * (wc / Wf) > 0x7FFFFFFFull) { *ok = false; return HJ3D_OK; }
    P2 = (uint32_t)(wc / Wf);
    coarse.recs = (Slot<KeyT>*)src.base; coarse.part_start = const_cast<unsigned long long*>(pre->seg_start);
    coarse.counts = const_cast<unsigned long long*>(pre->seg_count); coarse.P = pre->n_seg; coarse.n_kept = n;
    if (P2 > (uint32_t)kMaxParts) {
      if (P2 & (P2 - 1)) { *ok = false; return HJ3D_OK; }
      uint32_t P2a = 1; while ((uint64_t)P2a * P2a < P2) P2a <<= 1;        // ranges -> P2a mid ranges -> P2b fine partitions each
      const uint32_t P2b = P2 / P2a;
      Partitioned<KeyT> mid;
      const uint64_t F_mid = ((uint64_t)F + P2b - 1) / P2b;
      HJ_TRY((refine_partitions<HASH>(c, coarse, dir, P2a, Wf * P2b, P1 * P2a, F_mid ? F_mid : 1, n, &mid)));
      return refine_partitions<HASH>(c, mid, dir, P2b, Wf, P1 * P2, F, n, fine_out);
    }
  } else {
    while ((uint64_t)P2 * P2 < F && P2 < (uint32_t)kMaxParts) P2 <<= 1;
    wc = (uint64_t)Wf * P2;
    P1 = (uint32_t)(((uint64_t)nl + wc - 1) / wc);
    if (P1 > (uint32_t)kMaxParts || wc > 0xFFFFFFFFull) { *ok = false; return HJ3D_OK; }
    HJ_TRY((partition_local<HASH, LEFTID>(c, src, dir, P1, (uint32_t)wc, 0, &coarse)));
  }
  return refine_partitions<HASH>(c, coarse, dir, P2, Wf, P1 * P2, F, n, fine_out);
}

// fine partitions: as many consecutive buckets as fit in shared memory together with their slots / groups
inline void set_fine_width(hj3d_ctx* c, hj3d_table* t, double payload_bytes_total) {
  const uint32_t nl = t->dir.n_local;
  const double per_bucket = 4.0 + payload_bytes_total / (double)(nl ? nl : 1);
  double w = 0.85 * (double)c->smem_slice_bytes / per_bucket;
  uint32_t fw = w >= (double)nl ? nl : (w < 1.0 ? 1u : (uint32_t)w);
  if (fw == 0) fw = 1;
  if (fw > 4096) fw = 4096;                                                   // k_build_fine scans 8 buckets per thread
  if (fw < nl) { uint32_t p2 = 1; while (p2 * 2 <= fw) p2 *= 2; fw = p2; }   // power of two -> shifts instead of divisions
  t->fine_width = fw;
  t->fine_parts = nl ? (nl + fw - 1) / fw : 1;
}

// compressed slices (probe_packed.cuh): 2 B of run start + one 32-bit (quotient | row id) word per row
inline void set_packed_geometry(hj3d_ctx* c, hj3d_table* t) {
  t->pk_ok = false;
  const uint32_t nl = t->dir.n_local;
  if (t->kind != HJ3D_CHAINING || t->hash_id != HJ3D_HASH_MURMUR32 || !t->rowid_bound || !nl || !t->n) return;
  uint32_t qbits = 0, rbits = 0;
  for (uint64_t q = 0xFFFFFFFFull / t->D; q; q >>= 1) ++qbits;             // quotient = hash / D
  for (uint64_t r = t->rowid_bound - 1; r; r >>= 1) ++rbits;
  if (rbits == 0) rbits = 1;
  if (qbits + rbits > 32) return;
  const double per_bucket = 2.0 + 4.0 * (double)t->n / (double)nl;
  const double w = 0.985 * ((double)c->packed_slice_bytes - 64.0) / per_bucket;   // a slice that overflows falls back to global lookups
  if (w < 2.0) return;
  uint32_t fw = w >= (double)nl ? nl : (uint32_t)w;
  if (fw < nl) { uint32_t p2 = 1; while (p2 * 2 <= fw) p2 *= 2; fw = p2; }
  if ((double)fw * (double)t->n / (double)nl * 1.02 + 64.0 > 65535.0) {      // 16-bit run starts
    while (fw > 1 && (double)fw * (double)t->n / (double)nl * 1.02 + 64.0 > 65535.0) fw >>= 1;
  }
  t->pk_width = fw; t->pk_parts = (nl + fw - 1) / fw; t->pk_rowid_bits = 32 - qbits;
  t->pk_ok = true;
}

// ---- build ----------------------------------------------------------------------------------------
// Chaining tables over large inputs: partition the build side into fine bucket ranges and let one block
// build each range in shared memory (k_build_fine).  Returns *done = false when a range does not fit
// (skewed / heavily duplicated keys): the caller then uses the global-memory kernels below.
template <int HASH>
int build_chaining_fine(hj3d_ctx* c, hj3d_table* t, Src src, Slot<typename HashT<HASH>::key_t>* slots, bool* done, uint64_t* kept,
                        const PartsView* pre = nullptr) {
  using KeyT = typename HashT<HASH>::key_t;
  *done = false;
  const uint64_t n = src.n;
  const uint32_t nl = t->dir.n_local;
  // bucket ranges of the build kernel: as wide as ~6K build records / 4096 buckets allow
  uint32_t Wf;
  {
    const double per_bucket = 4.0 + (double)n * sizeof(Slot<KeyT>) / (double)(nl ? nl : 1);
    double w = 0.85 * (double)c->smem_build_bytes / per_bucket;
    Wf = w >= (double)nl ? nl : (w < 1.0 ? 1u : (uint32_t)w);
    if (Wf > 2048) Wf = 2048;
    if (Wf < nl) { uint32_t p2 = 1; while (p2 * 2 <= Wf) p2 *= 2; Wf = p2; }
    // the block keeps its records in registers: expected records per range + 25% must fit
    while (Wf > 1 && (double)n / (double)(nl ? nl : 1) * Wf * 1.25 + 256.0 > (double)(kFineBuildThreads * kFineBuildItems)) Wf >>= 1;
  }
  const uint32_t F = nl ? (nl + Wf - 1) / Wf : 1;
  if (!c->smem_build || (int64_t)n < c->smem_min_probe || F < 2) return HJ3D_OK;
  if ((double)n * 1.08 + 4096.0 * (double)F >= 4.0e9) return HJ3D_OK;
  Partitioned<KeyT> fine;
  bool ok = false;
  HJ_TRY((partition_fine<HASH, false>(c, src, t->dir, Wf, F, &fine, &ok, pre)));
  if (!ok) return HJ3D_OK;
  PhaseTimer pt(c, PH_SCATTER);
  unsigned long long* base = nullptr;
  HJ_TRY(dev_alloc(c, &base, (uint64_t)fine.P + 1));
  HJ_TRY((run_scan<unsigned long long, false>(c, LoadU64{fine.counts}, StoreExU64{base}, fine.P, (DevStats*)nullptr, (unsigned long long*)nullptr)));
  const uint32_t off_bytes = ((Wf + 1) * 4 + 15) & ~15u;
  uint32_t cap_recs = kFineBuildThreads * kFineBuildItems;
  const uint32_t smem_max = 160u << 10;
  if (off_bytes + (uint64_t)cap_recs * sizeof(Slot<KeyT>) > smem_max) cap_recs = (smem_max - off_bytes) / sizeof(Slot<KeyT>);
  // right-size: expected records per range + 25% + 512
  const uint64_t expect = n / F + n / (4ull * F) + 512;
  if (expect < cap_recs) cap_recs = (uint32_t)expect;
  const size_t sm = off_bytes + (size_t)cap_recs * sizeof(Slot<KeyT>);
  uint32_t* d_flag = (uint32_t*)(c->d_scalar + 2);
  CUDA_TRY(cudaMemsetAsync(d_flag, 0, 4, c->stream));
  CUDA_TRY(cudaFuncSetAttribute(k_build_fine<HASH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  // ranges past fine.P (two-level rounding never creates them; one-level has exactly F) -> grid = F
  k_build_fine<HASH><<<F, kFineBuildThreads, sm, c->stream>>>(fine.recs, fine.part_start, fine.counts, base, t->dir, Wf, fine.P,
                                                              cap_recs, t->off, slots, c->d_stats, d_flag);
  k_stats_fold<<<1, 32, 0, c->stream>>>(c->d_stats, kStatsCopies);
  c->launches += 2;
  uint32_t* h = (uint32_t*)c->h_pinned;
  CUDA_TRY(cudaMemcpyAsync(h, d_flag, 4, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  CUDA_TRY(cudaGetLastError());
  if (*h) return HJ3D_OK;                                   // some range overflowed shared memory
  (void)nl;
  *kept = fine.n_kept;                                      // tuples of buckets outside a shard's range were dropped
  *done = true;
  return HJ3D_OK;
}

// Nested tables over large inputs: the same fine bucket ranges, grouped by key in shared memory
// (build_nested_fine.cuh).  *done = false: a range or a bucket is too large (skew) -> global-memory kernels.
template <int HASH>
int build_nested_fine(hj3d_ctx* c, hj3d_table* t, Src src, bool* done, const PartsView* pre = nullptr) {
  using KeyT = typename HashT<HASH>::key_t;
  *done = false;
  const uint64_t n = src.n;
  const uint32_t nl = t->dir.n_local;
  uint32_t Wf;
  {
    const double per_bucket = 8.0 + (double)n * sizeof(Slot<KeyT>) / (double)(nl ? nl : 1);
    double w = 0.85 * (double)c->smem_build_bytes / per_bucket;
    Wf = w >= (double)nl ? nl : (w < 1.0 ? 1u : (uint32_t)w);
    if (Wf > 2048) Wf = 2048;
    if (Wf < nl) { uint32_t p2 = 1; while (p2 * 2 <= Wf) p2 *= 2; Wf = p2; }
    // the block keeps its records in registers: expected records per range + 25% must fit
    while (Wf > 1 && (double)n / (double)(nl ? nl : 1) * Wf * 1.25 + 256.0 > (double)(kNfThreads * kNfItems)) Wf >>= 1;
  }
  const uint32_t F = nl ? (nl + Wf - 1) / Wf : 1;
  if (!c->smem_build || (int64_t)n < c->smem_min_probe || F < 2) return HJ3D_OK;
  if ((double)n * 1.08 + 4096.0 * (double)F >= 4.0e9) return HJ3D_OK;
  Partitioned<KeyT> fine;
  bool ok = false;
  HJ_TRY((partition_fine<HASH, false>(c, src, t->dir, Wf, F, &fine, &ok, pre)));
  if (!ok) return HJ3D_OK;
  PhaseTimer pt(c, PH_GROUP);
  unsigned long long* base = nullptr;
  HJ_TRY(dev_alloc(c, &base, (uint64_t)fine.P + 1));
  HJ_TRY((run_scan<unsigned long long, false>(c, LoadU64{fine.counts}, StoreExU64{base}, fine.P, (DevStats*)nullptr, (unsigned long long*)nullptr)));
  const uint32_t hdr_bytes = (2 * (Wf + 1) * 4 + 15) & ~15u;
  uint32_t cap_recs = kNfThreads * kNfItems;
  const uint32_t smem_max = 160u << 10;
  constexpr uint32_t kPerRec = nf_bytes_per_record<KeyT>();                  // record + group word + group start + bucket
  if (hdr_bytes + (uint64_t)cap_recs * kPerRec > smem_max) cap_recs = (smem_max - hdr_bytes) / kPerRec;
  const uint64_t expect = n / F + n / (2ull * F) + 512;                      // expected records per range + 50% + 512
  if (expect < cap_recs) cap_recs = (uint32_t)expect;
  cap_recs &= ~15u;
  const size_t sm = hdr_bytes + (size_t)cap_recs * kPerRec;
  Group<KeyT>* gtmp = nullptr;                                             // one slot per build row: #groups is only known afterwards
  unsigned long long *gcount = nullptr, *gbase = nullptr;
  HJ_TRY(dev_alloc(c, &gtmp, n));
  HJ_TRY(dev_alloc(c, &gcount, (uint64_t)F + 1));
  HJ_TRY(dev_alloc(c, &gbase, (uint64_t)F + 1));
  HJ_TRY(buf_ensure(c, t->b_goff, &t->goff, (uint64_t)nl + 1));
  HJ_TRY(buf_ensure(c, t->b_rows, &t->rows, n));
  uint32_t* d_flag = (uint32_t*)(c->d_scalar + 2);
  CUDA_TRY(cudaMemsetAsync(c->d_scalar, 0, 4 * sizeof(unsigned long long), c->stream));
  HJ_TRY(init_dev_stats_copies(c));
  CUDA_TRY(cudaFuncSetAttribute(k_build_fine_nested<HASH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  k_build_fine_nested<HASH><<<F, kNfThreads, sm, c->stream>>>(fine.recs, fine.part_start, fine.counts, base, t->dir, Wf, F, cap_recs,
                                                              t->goff, gtmp, t->rows, gcount, c->d_stats, d_flag);
  k_stats_fold<<<1, 32, 0, c->stream>>>(c->d_stats, kStatsCopies);
  c->launches += 2;
  // first global group index of every partition; the total goes to the host (the groups array is sized from it)
  HJ_TRY((run_scan<unsigned long long, false>(c, LoadU64{gcount}, StoreExU64{gbase}, F, (DevStats*)nullptr, c->d_scalar)));
  unsigned long long* h = (unsigned long long*)c->h_pinned;
  CUDA_TRY(cudaMemcpyAsync(h, c->d_scalar, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  CUDA_TRY(cudaGetLastError());
  if (*(uint32_t*)(h + 2)) return HJ3D_OK;                                  // a range / bucket overflowed shared memory
  const uint64_t G = h[0];
  Group<KeyT>* groups = nullptr;
  HJ_TRY(buf_ensure(c, t->b_groups, &groups, G));
  k_nested_compact<KeyT><<<F, 256, 0, c->stream>>>(gtmp, base, gbase, gcount, Wf, nl, F, groups, t->goff);
  ++c->launches;
  CUDA_TRY(cudaGetLastError());
  t->groups = groups; t->n_groups = G; t->n = fine.n_kept; t->slots = nullptr;
  t->parts = 1; t->part_width = nl ? nl : 1;
  CUDA_TRY(cudaMemcpyAsync(&t->hstats, c->d_stats, sizeof(DevStats), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  set_fine_width(c, t, (double)G * sizeof(Group<KeyT>));
  t->have_stats = true; t->built = true;
  *done = true;
  return HJ3D_OK;
}

template <int HASH>
int build_impl(hj3d_ctx* c, hj3d_table* t, Src src, const PartsView* pre = nullptr) {
  using KeyT = typename HashT<HASH>::key_t;
  const uint64_t n = src.n;
  const uint32_t nl = t->dir.n_local;
  const Dir d = t->dir;
  const bool agg = c->warp_aggregate != 0;
  Slot<KeyT>* slots = nullptr;
  if (nl == 0) {   // a shard that owns no bucket (hj3d_owner_range with fewer buckets than owners): every tuple is foreign
    HJ_TRY(buf_ensure(c, t->b_off, &t->off, 1));
    CUDA_TRY(cudaMemsetAsync(t->off, 0, 4, c->stream));
    if (t->kind == HJ3D_NESTED) { HJ_TRY(buf_ensure(c, t->b_goff, &t->goff, 1)); CUDA_TRY(cudaMemsetAsync(t->goff, 0, 4, c->stream)); }
    t->n = 0; t->n_groups = 0; t->parts = 1; t->part_width = 1; t->fine_width = 0; t->fine_parts = 1;
    init_dev_stats_host(t->hstats);
    t->have_stats = true; t->built = true;
    return HJ3D_OK;
  }
  if (t->kind == HJ3D_NESTED) {
    bool done = false;
    HJ_TRY(build_nested_fine<HASH>(c, t, src, &done, pre));
    if (done) return HJ3D_OK;
    if (pre) { HJ_TRY(make_contiguous<KeyT>(c, &src, pre)); pre = nullptr; }
  }
  HJ_TRY(buf_ensure(c, t->b_off, &t->off, (uint64_t)nl + 1));
  if (t->kind == HJ3D_CHAINING) HJ_TRY(buf_ensure(c, t->b_slots, &slots, n));
  else                          HJ_TRY(dev_alloc(c, &slots, n));          // nested: only needed during the build
  t->slots = slots;
  if (t->kind == HJ3D_CHAINING) {
    set_fine_width(c, t, (double)n * sizeof(Slot<KeyT>));
    HJ_TRY(init_dev_stats_copies(c));
    bool done = false;
    uint64_t kept = n;
    HJ_TRY(build_chaining_fine<HASH>(c, t, src, slots, &done, &kept, pre));
    if (done) {
      t->n = kept; t->parts = 1; t->part_width = nl ? nl : 1;
      CUDA_TRY(cudaMemcpyAsync(&t->hstats, c->d_stats, sizeof(DevStats), cudaMemcpyDeviceToHost, c->stream));
      CUDA_TRY(cudaGetLastError());
      t->have_stats = true; t->built = true;
      return HJ3D_OK;
    }
    if (pre) { HJ_TRY(make_contiguous<KeyT>(c, &src, pre)); pre = nullptr; }
  }

  // bucket-order the input first when the directory + slots do not fit the L2 window budget
  const uint64_t table_bytes = (uint64_t)nl * 4 + n * sizeof(Slot<KeyT>);
  uint32_t P = choose_parts(c, table_bytes, nl);
  uint32_t width = (uint32_t)(((uint64_t)nl + P - 1) / P);
  if ((double)n * 1.04 + 8192.0 * P >= 4.0e9) P = 1;                      // tile maps address records with 32 bits
  Partitioned<KeyT> pr;
  const uint2* tilemap = nullptr;
  uint32_t n_tiles = blocks_for(n, kBuildTile);
  Src bsrc = src;
  uint64_t n_kept = n;                                                    // < n only for shard tables fed foreign tuples
  if (P > 1 && n) {
    HJ_TRY((partition_local<HASH, false>(c, src, d, P, width, 0, &pr)));
    uint2* tm = nullptr;
    HJ_TRY(make_tilemap(c, pr, kBuildTile, &tm, &n_tiles));
    tilemap = tm;
    bsrc = records_src(pr);
    n_kept = pr.n_kept;
    t->parts = P; t->part_width = width;
  } else {
    P = 1;
    t->parts = 1; t->part_width = nl ? nl : 1;
  }
  DevStats hs; init_dev_stats_host(hs);
  CUDA_TRY(cudaMemcpyAsync(c->d_stats, &hs, sizeof(hs), cudaMemcpyHostToDevice, c->stream));
  {
    PhaseTimer pt(c, PH_HIST);
    CUDA_TRY(cudaMemsetAsync(t->off, 0, ((uint64_t)nl + 1) * 4, c->stream));
    if (n_tiles) {
      if (agg) k_histogram<HASH, true><<<n_tiles, kBuildThreads, 0, c->stream>>>(bsrc, d, tilemap, t->off);
      else     k_histogram<HASH, false><<<n_tiles, kBuildThreads, 0, c->stream>>>(bsrc, d, tilemap, t->off);
      ++c->launches;
    }
  }
  {
    PhaseTimer pt(c, PH_SCAN);
    // chaining statistics are a function of the bucket histogram alone
    const bool want_stats = t->kind == HJ3D_CHAINING;
    if (want_stats) HJ_TRY((run_scan<uint32_t, true>(c, LoadU32{t->off}, StoreInclU32{t->off}, nl, c->d_stats, (uint32_t*)nullptr)));
    else            HJ_TRY((run_scan<uint32_t, false>(c, LoadU32{t->off}, StoreInclU32{t->off}, nl, (DevStats*)nullptr, (uint32_t*)nullptr)));
    // off[nl] = number of tuples kept = inclusive end of the last bucket.  A shard table (its directory is a sub-range
    // of the buckets) drops tuples of foreign buckets, so the kept count comes from the histogram, not from n.
    if (nl) CUDA_TRY(cudaMemcpyAsync(t->off + nl, t->off + nl - 1, 4, cudaMemcpyDeviceToDevice, c->stream));
    else    CUDA_TRY(cudaMemsetAsync(t->off, 0, 4, c->stream));
    if (nl != t->D && P <= 1) {
      uint32_t* h = (uint32_t*)c->h_pinned;
      CUDA_TRY(cudaMemcpyAsync(h, t->off + nl, 4, cudaMemcpyDeviceToHost, c->stream));
      CUDA_TRY(cudaStreamSynchronize(c->stream));
      n_kept = *h;
    }
  }
  {
    PhaseTimer pt(c, PH_SCATTER);
    if (n_tiles) {
      if (agg) k_scatter<HASH, true><<<n_tiles, kBuildThreads, 0, c->stream>>>(bsrc, d, tilemap, t->off, slots);
      else     k_scatter<HASH, false><<<n_tiles, kBuildThreads, 0, c->stream>>>(bsrc, d, tilemap, t->off, slots);
      ++c->launches;
      if (t->kind == HJ3D_CHAINING) {   // short buckets into chain order (probe.cuh walks them like algebra.hh:644-657)
        k_order_slots<KeyT><<<blocks_for(nl, 256), 256, 0, c->stream>>>(t->off, slots, nl);
        ++c->launches;
      }
    }
  }
  t->n = n_kept;
  if (t->kind == HJ3D_NESTED) {
    PhaseTimer pt(c, PH_GROUP);
    const uint64_t nk = n_kept;                                           // slots [0, nk) are valid
    uint32_t *cell = nullptr, *rep = nullptr, *gcnt = nullptr, *gmin = nullptr, *gidx = nullptr, *gstart = nullptr;
    HJ_TRY(dev_alloc(c, &cell, nk)); HJ_TRY(dev_alloc(c, &rep, nk)); HJ_TRY(dev_alloc(c, &gcnt, nk));
    HJ_TRY(dev_alloc(c, &gmin, nk)); HJ_TRY(dev_alloc(c, &gidx, nk + 1)); HJ_TRY(dev_alloc(c, &gstart, nk + 1));
    CUDA_TRY(cudaMemsetAsync(cell, 0xFF, (nk ? nk : 1) * 4, c->stream));
    CUDA_TRY(cudaMemsetAsync(gmin, 0xFF, (nk ? nk : 1) * 4, c->stream));
    CUDA_TRY(cudaMemsetAsync(gcnt, 0, (nk ? nk : 1) * 4, c->stream));
    if (nk) {
      k_group_claim<HASH><<<blocks_for(nk, kBuildTile), kBuildThreads, 0, c->stream>>>(slots, nk, d, t->off, cell, rep, gcnt, gmin);
      ++c->launches;
    }
    unsigned long long* d_tot = c->d_scalar;
    HJ_TRY((run_scan<unsigned long long, false>(c, LoadCells{cell, gcnt, nk}, StoreCells{gidx, gstart}, nk + 1,
                                                  (DevStats*)nullptr, d_tot)));
    unsigned long long* h_tot = (unsigned long long*)c->h_pinned;
    CUDA_TRY(cudaMemcpyAsync(h_tot, d_tot, 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    const uint64_t G = *h_tot >> 32;
    t->n_groups = G;
    Group<KeyT>* groups = nullptr;
    HJ_TRY(buf_ensure(c, t->b_groups, &groups, G));
    t->groups = groups;
    HJ_TRY(buf_ensure(c, t->b_goff, &t->goff, (uint64_t)nl + 1));
    HJ_TRY(buf_ensure(c, t->b_rows, &t->rows, nk));
    if (nk) {
      k_group_emit<HASH><<<blocks_for(nk, 256), 256, 0, c->stream>>>(slots, nk, cell, gcnt, gmin, gidx, gstart, groups);
      ++c->launches;
    }
    k_group_offsets<<<blocks_for((uint64_t)nl + 1, 256), 256, 0, c->stream>>>(t->off, gidx, nl + 1, t->goff);
    ++c->launches;
    if (nk) {   // main chains into first-appearance order (probe.cuh walks them like ht_nested.hh:371-379)
      k_order_groups<KeyT><<<blocks_for(nl, 256), 256, 0, c->stream>>>(t->goff, groups, nl);
      ++c->launches;
    }
    if (nk) {
      k_group_rows<HASH><<<blocks_for(nk, kBuildTile), kBuildThreads, 0, c->stream>>>(slots, nk, rep, gstart, t->rows);
      ++c->launches;
    }
    // nested statistics: main chain length per bucket = #distinct keys (ht_nested.hh:459-479)
    uint32_t* dummy = nullptr;
    const uint32_t nb = blocks_for(nl, kScanTile);
    HJ_TRY(dev_alloc(c, &dummy, nb));
    if (nb) {
      k_scan_reduce<uint32_t, LoadDiff, true><<<nb, kScanThreads, 0, c->stream>>>(LoadDiff{t->goff}, nl, dummy, c->d_stats);
      ++c->launches;
    }
    t->slots = nullptr;                                                   // arena memory, dead after this call
  }
  CUDA_TRY(cudaMemcpyAsync(&t->hstats, c->d_stats, sizeof(DevStats), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaGetLastError());
  if (t->kind == HJ3D_NESTED) set_fine_width(c, t, (double)t->n_groups * sizeof(Group<KeyT>));
  t->have_stats = true;
  t->built = true;
  return HJ3D_OK;
}

int fetch_counters(hj3d_ctx* c, hj3d_counters* out, uint64_t out_cap, bool wrote) {
  DevCounters* h = (DevCounters*)c->h_pinned;
  CUDA_TRY(cudaMemcpyAsync(h, c->d_ctr, sizeof(DevCounters), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  CUDA_TRY(cudaGetLastError());
  out->matches = h->matches; out->num_cmps = h->num_cmps;
  out->checksum_sum = h->checksum_sum; out->checksum_xor = h->checksum_xor;
  out->out_tuples = h->matches;
  out->overflow = (wrote && h->out_cursor > out_cap) ? 1 : 0;
  out->out_written = wrote ? (h->out_cursor > out_cap ? out_cap : h->out_cursor) : 0;
  return HJ3D_OK;
}

// ---- probe planning ---------------------------------------------------------------------------------
// Large probe inputs are bucket-range partitioned (one or two levels) into the table's FINE partitions
// and probed in shared memory (probe_smem.cuh); small ones use the global-memory kernels of probe.cuh,
// optionally bucket-ordered with the build's own (coarse) partition function.
template <class KeyT> struct ProbePlan {
  Src             src;
  bool            recs = false;        // src is an array of Slot<KeyT> records (one vector load per probe)
  bool            smem = false;
  const uint2*    work = nullptr;      // smem: work items; global: tile map (nullable)
  const uint32_t* work_part = nullptr;
  uint32_t        n_work = 0;
  FineCfg         fc{};
};

template <int HASH>
int plan_probe(hj3d_ctx* c, hj3d_table* t, Src src, ProbePlan<typename HashT<HASH>::key_t>* pl, const PartsView* pre = nullptr) {
  using KeyT = typename HashT<HASH>::key_t;
  if (pre) {   // pre-partitioned input: only the fine-partition path continues from the given coarse ranges
    const bool fine_path = c->smem_probe && t->fine_width && (int64_t)src.n >= c->smem_min_probe && t->fine_parts > 1 &&
                           (double)src.n * 1.08 + 4096.0 * (double)t->fine_parts < 4.0e9 &&
                           pre->range_width % t->fine_width == 0 &&
                           (pre->range_width / t->fine_width <= (uint32_t)kMaxParts ||            // more: two refinement passes (powers of two)
                            (pre->range_width / t->fine_width <= (uint32_t)kMaxParts * (uint32_t)kMaxParts &&
                             ((pre->range_width / t->fine_width) & (pre->range_width / t->fine_width - 1)) == 0));
    if (!fine_path) { HJ_TRY(make_contiguous<KeyT>(c, &src, pre)); pre = nullptr; }
  }
  pl->src = src;
  pl->recs = !src.gather && src.stride == sizeof(Slot<KeyT>) && src.key_off == 0 && src.rowid_off == sizeof(KeyT) &&
             ((uintptr_t)src.base % sizeof(Slot<KeyT>)) == 0;
  const uint64_t n = src.n;
  const bool can_index32 = (double)n * 1.08 + 4096.0 * (double)t->fine_parts < 4.0e9;
  if (c->smem_probe && t->fine_width && (int64_t)n >= c->smem_min_probe && can_index32) {
    const uint32_t F = t->fine_parts, Wf = t->fine_width, nl = t->dir.n_local;
    const uint32_t chunk = (uint32_t)c->smem_chunk;
    pl->smem = true;
    {
      const double rows_per_bucket = (t->kind == HJ3D_CHAINING ? (double)t->n * sizeof(Slot<KeyT>) : (double)t->n_groups * sizeof(Group<KeyT>)) / (double)(nl ? nl : 1);
      const double expect = ((double)Wf + 1) * 4.0 + 32.0 + (double)Wf * rows_per_bucket;
      uint64_t want = (uint64_t)(expect * 1.2) + 2048;
      if (want > (uint64_t)c->smem_slice_bytes) want = (uint64_t)c->smem_slice_bytes;
      pl->fc = FineCfg{Wf, nl, (uint32_t)(want & ~15ull)};
    }
    uint2* work = nullptr; uint32_t* wpart = nullptr;
    if (F == 1) {                                    // the whole table fits: no partitioning at all
      pl->n_work = blocks_for(n, chunk);
      HJ_TRY(dev_alloc(c, &work, pl->n_work)); HJ_TRY(dev_alloc(c, &wpart, pl->n_work));
      k_make_chunks<<<blocks_for(pl->n_work, 256), 256, 0, c->stream>>>(n, chunk, pl->n_work, work, wpart);
      ++c->launches;
      pl->work = work; pl->work_part = wpart;
      CUDA_TRY(cudaGetLastError());
      return HJ3D_OK;
    }
    Partitioned<KeyT> fine;
    bool ok = false;
    HJ_TRY((partition_fine<HASH, true>(c, src, t->dir, Wf, F, &fine, &ok, pre)));
    if (!ok) { pl->smem = false; goto global_path; }
    HJ_TRY(make_tilemap(c, fine, chunk, &work, &pl->n_work, &wpart));
    pl->work = work; pl->work_part = wpart;
    pl->src = records_src(fine);
    pl->recs = true;
    return HJ3D_OK;
  }
global_path:
  pl->smem = false;
  pl->work = nullptr;
  pl->n_work = blocks_for(n, kProbeTile);
  if (t->parts <= 1 || src.gather || (int64_t)n < c->partition_min_probe) return HJ3D_OK;
  if ((double)n * 1.04 + 8192.0 * t->parts >= 4.0e9) return HJ3D_OK;
  Partitioned<KeyT> pr;
  HJ_TRY((partition_local<HASH, true>(c, src, t->dir, t->parts, t->part_width, 0, &pr)));
  uint2* tm = nullptr;
  HJ_TRY(make_tilemap(c, pr, kProbeTile, &tm, &pl->n_work));
  pl->work = tm;
  pl->src = records_src(pr);
  pl->recs = true;
  return HJ3D_OK;
}

// Unique chaining probes of large inputs over compressed slices (probe_packed.cuh).  *done = false: not applicable
// (the caller continues with the general paths).
inline int probe_packed_impl(hj3d_ctx* c, hj3d_table* t, Src src, uint32_t flags, uint2* out, uint64_t cap, bool* done, const PartsView* pre) {
  constexpr int HASH = HJ3D_HASH_MURMUR32;
  *done = false;
  const uint32_t F = t->pk_parts, Wf = t->pk_width, nl = t->dir.n_local;
  if (F < 2 || (double)src.n * 1.08 + 4096.0 * (double)F >= 4.0e9) return HJ3D_OK;
  Partitioned<uint32_t> fine;
  bool ok = false;
  HJ_TRY((partition_fine<HASH, true>(c, src, t->dir, Wf, F, &fine, &ok, pre)));
  if (!ok) return HJ3D_OK;
  uint2* work = nullptr; uint32_t* wpart = nullptr; uint32_t n_work = 0;
  HJ_TRY(make_tilemap(c, fine, 1u << 18, &work, &n_work, &wpart));
  if (!n_work) { *done = true; return HJ3D_OK; }
  PhaseTimer pt(c, PH_PROBE);
  PackCfg pc{};
  pc.width = Wf; pc.n_local = nl; pc.smem_bytes = (uint32_t)c->packed_slice_bytes; pc.rowid_bits = t->pk_rowid_bits;
  pc.pow2 = t->dir.is_pow2; pc.qshift = 0;
  while (pc.pow2 && (1ull << pc.qshift) < t->D) ++pc.qshift;
  pc.qmagic = t->dir.magic;
  const bool cs = flags & HJ3D_F_CHECKSUM, wr = out != nullptr;
  const size_t sm = pc.smem_bytes;
  CUDA_TRY(launch_probe_packed(c->stream, cs, wr, n_work, sm, (const Slot<uint32_t>*)fine.recs, t->dir, pc, work, wpart, t->off,
                               (const Slot<uint32_t>*)t->slots, out, cap, c->d_ctr));
  ++c->launches;
  CUDA_TRY(cudaGetLastError());
  *done = true;
  return HJ3D_OK;
}

template <int HASH>
int probe_chaining_impl(hj3d_ctx* c, hj3d_table* t, Src src, bool unique, uint32_t flags, uint2* out, uint64_t cap, const PartsView* pre = nullptr) {
  using KeyT = typename HashT<HASH>::key_t;
  if (!src.n || !t->dir.n_local) return HJ3D_OK;
  if (HASH == HJ3D_HASH_MURMUR32 && unique && !src.gather && c->packed_probe && t->pk_ok && (int64_t)src.n >= c->packed_min_probe) {
    bool done = false;
    HJ_TRY(probe_packed_impl(c, t, src, flags, out, cap, &done, pre));
    if (done) return HJ3D_OK;
  }
  ProbePlan<KeyT> pl;
  HJ_TRY(plan_probe<HASH>(c, t, src, &pl, pre));
  if (!pl.n_work) return HJ3D_OK;
  PhaseTimer pt(c, PH_PROBE);
  const bool cs = flags & HJ3D_F_CHECKSUM, wr = out != nullptr;
  const Slot<KeyT>* slots = (const Slot<KeyT>*)t->slots;
  const uint32_t nb = pl.n_work;
  if (pl.smem && pl.recs && unique && c->lean_probe) {
    const size_t sm = pl.fc.smem_bytes;
    const Slot<KeyT>* recs = (const Slot<KeyT>*)pl.src.base;
#define LAUNCH_PF(C, W) do { \
      CUDA_TRY(cudaFuncSetAttribute(k_probe_fine<HASH, 0, C, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
      k_probe_fine<HASH, 0, C, W><<<nb, kFineThreads, sm, c->stream>>>(recs, t->dir, pl.fc, pl.work, pl.work_part, t->off, (const void*)slots, out, cap, c->d_ctr); } while (0)
    if (cs) { if (wr) LAUNCH_PF(true, true); else LAUNCH_PF(true, false); }
    else    { if (wr) LAUNCH_PF(false, true); else LAUNCH_PF(false, false); }
#undef LAUNCH_PF
  } else if (pl.smem) {
    const size_t sm = pl.fc.smem_bytes;
#define LAUNCH_PS3(U, C, W, R, T) do { \
      CUDA_TRY(cudaFuncSetAttribute(k_probe_chaining_smem<HASH, U, C, W, R, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
      k_probe_chaining_smem<HASH, U, C, W, R, T><<<nb, T, sm, c->stream>>>(pl.src, t->dir, pl.fc, pl.work, pl.work_part, t->off, slots, out, cap, c->d_ctr); } while (0)
#define LAUNCH_PS2(U, C, W, R) do { if (c->probe_threads == 512) LAUNCH_PS3(U, C, W, R, 512); else LAUNCH_PS3(U, C, W, R, 256); } while (0)
#define LAUNCH_PS(U, C, W) do { if (pl.recs) LAUNCH_PS2(U, C, W, true); else LAUNCH_PS2(U, C, W, false); } while (0)
    if (unique) { if (cs) { if (wr) LAUNCH_PS(true, true, true); else LAUNCH_PS(true, true, false); }
                  else    { if (wr) LAUNCH_PS(true, false, true); else LAUNCH_PS(true, false, false); } }
    else        { if (cs) { if (wr) LAUNCH_PS(false, true, true); else LAUNCH_PS(false, true, false); }
                  else    { if (wr) LAUNCH_PS(false, false, true); else LAUNCH_PS(false, false, false); } }
#undef LAUNCH_PS
#undef LAUNCH_PS2
#undef LAUNCH_PS3
  } else {
#define LAUNCH_PC2(U, C, W, R) k_probe_chaining<HASH, U, C, W, R><<<nb, kProbeThreads, 0, c->stream>>>(pl.src, t->dir, pl.work, t->off, slots, out, cap, c->d_ctr)
#define LAUNCH_PC(U, C, W) do { if (pl.recs) LAUNCH_PC2(U, C, W, true); else LAUNCH_PC2(U, C, W, false); } while (0)
    if (unique) { if (cs) { if (wr) LAUNCH_PC(true, true, true); else LAUNCH_PC(true, true, false); }
                  else    { if (wr) LAUNCH_PC(true, false, true); else LAUNCH_PC(true, false, false); } }
    else        { if (cs) { if (wr) LAUNCH_PC(false, true, true); else LAUNCH_PC(false, true, false); }
                  else    { if (wr) LAUNCH_PC(false, false, true); else LAUNCH_PC(false, false, false); } }
#undef LAUNCH_PC
#undef LAUNCH_PC2
  }
  ++c->launches;
  CUDA_TRY(cudaGetLastError());
  return HJ3D_OK;
}

template <int HASH>
int probe_nested_impl(hj3d_ctx* c, hj3d_table* t, Src src, uint32_t flags, uint2* out, uint64_t cap, const PartsView* pre = nullptr) {
  using KeyT = typename HashT<HASH>::key_t;
  if (!src.n || !t->dir.n_local) return HJ3D_OK;
  ProbePlan<KeyT> pl;
  HJ_TRY(plan_probe<HASH>(c, t, src, &pl, pre));
  if (!pl.n_work) return HJ3D_OK;
  PhaseTimer pt(c, PH_PROBE);
  const bool cs = flags & HJ3D_F_CHECKSUM, wr = out != nullptr;
  const Group<KeyT>* groups = (const Group<KeyT>*)t->groups;
  const uint32_t nb = pl.n_work;
  if (pl.smem && pl.recs && c->lean_probe) {
    const size_t sm = pl.fc.smem_bytes;
    const Slot<KeyT>* recs = (const Slot<KeyT>*)pl.src.base;
#define LAUNCH_PF(C, W) do { \
      CUDA_TRY(cudaFuncSetAttribute(k_probe_fine<HASH, 1, C, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
      k_probe_fine<HASH, 1, C, W><<<nb, kFineThreads, sm, c->stream>>>(recs, t->dir, pl.fc, pl.work, pl.work_part, t->goff, (const void*)groups, out, cap, c->d_ctr); } while (0)
    if (cs) { if (wr) LAUNCH_PF(true, true); else LAUNCH_PF(true, false); }
    else    { if (wr) LAUNCH_PF(false, true); else LAUNCH_PF(false, false); }
#undef LAUNCH_PF
  } else if (pl.smem) {
    const size_t sm = pl.fc.smem_bytes;
#define LAUNCH_NS3(C, W, R, T) do { \
      CUDA_TRY(cudaFuncSetAttribute(k_probe_nested_smem<HASH, C, W, R, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
      k_probe_nested_smem<HASH, C, W, R, T><<<nb, T, sm, c->stream>>>(pl.src, t->dir, pl.fc, pl.work, pl.work_part, t->goff, groups, out, cap, c->d_ctr); } while (0)
#define LAUNCH_NS2(C, W, R) do { if (c->probe_threads == 512) LAUNCH_NS3(C, W, R, 512); else LAUNCH_NS3(C, W, R, 256); } while (0)
#define LAUNCH_NS(C, W) do { if (pl.recs) LAUNCH_NS2(C, W, true); else LAUNCH_NS2(C, W, false); } while (0)
    if (cs) { if (wr) LAUNCH_NS(true, true); else LAUNCH_NS(true, false); }
    else    { if (wr) LAUNCH_NS(false, true); else LAUNCH_NS(false, false); }
#undef LAUNCH_NS
#undef LAUNCH_NS2
#undef LAUNCH_NS3
  } else {
#define LAUNCH_PN2(C, W, R) k_probe_nested<HASH, C, W, R><<<nb, kProbeThreads, 0, c->stream>>>(pl.src, t->dir, pl.work, t->goff, groups, out, cap, c->d_ctr)
#define LAUNCH_PN(C, W) do { if (pl.recs) LAUNCH_PN2(C, W, true); else LAUNCH_PN2(C, W, false); } while (0)
    if (cs) { if (wr) LAUNCH_PN(true, true); else LAUNCH_PN(true, false); }
    else    { if (wr) LAUNCH_PN(false, true); else LAUNCH_PN(false, false); }
#undef LAUNCH_PN
#undef LAUNCH_PN2
  }
  ++c->launches;
  CUDA_TRY(cudaGetLastError());
  return HJ3D_OK;
}

// nested probe + unnest in one kernel (probe_unnest.cuh).  *fused = false: the input does not take the fine-partition path
// (small / gathered input); the caller then composes hj3d_probe_nested + hj3d_unnest_pairs.
template <int HASH>
int probe_nested_unnest_impl(hj3d_ctx* c, hj3d_table* t, Src src, uint32_t flags, uint2* out, uint64_t cap, bool* fused,
                             const PartsView* pre = nullptr) {
  using KeyT = typename HashT<HASH>::key_t;
  *fused = false;
  if (!src.n || src.gather || !c->lean_probe || !t->dir.n_local) return HJ3D_OK;
  ProbePlan<KeyT> pl;
  HJ_TRY(plan_probe<HASH>(c, t, src, &pl, pre));
  if (!pl.smem || !pl.recs || !pl.n_work) return HJ3D_OK;               // (a partition pass that was made is simply not used)
  PhaseTimer pt(c, PH_PROBE);
  const bool cs = flags & HJ3D_F_CHECKSUM, wr = out != nullptr;
  const size_t sm = pl.fc.smem_bytes;
  const Slot<KeyT>* recs = (const Slot<KeyT>*)pl.src.base;
  const Group<KeyT>* groups = (const Group<KeyT>*)t->groups;
  // probe records that hit a group of more than kUnnestWarpMax rows are listed and expanded by finish_fused_unnest
  const uint32_t hot_cap = (uint32_t)(c->unnest_hot_cap > 0 ? c->unnest_hot_cap : 1);
  HotGroup* hot_list = nullptr;
  HJ_TRY(dev_alloc(c, &hot_list, hot_cap));
  CUDA_TRY(cudaMemsetAsync(c->d_scalar + 1, 0, 8, c->stream));
  c->fused_hot_list = hot_list; c->fused_hot_cap = hot_cap;
#define LAUNCH_PU(C, W) do { \
    CUDA_TRY(cudaFuncSetAttribute(k_probe_nested_unnest<HASH, C, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
    k_probe_nested_unnest<HASH, C, W><<<pl.n_work, kFineThreads, sm, c->stream>>>(recs, t->dir, pl.fc, pl.work, pl.work_part, t->goff, groups, \
                                                                                   t->rows, out, cap, c->d_ctr, hot_list, hot_cap, c->d_scalar + 1); } while (0)
  if (cs) { if (wr) LAUNCH_PU(true, true); else LAUNCH_PU(true, false); }
  else    { if (wr) LAUNCH_PU(false, true); else LAUNCH_PU(false, false); }
#undef LAUNCH_PU
  ++c->launches;
  CUDA_TRY(cudaGetLastError());
  *fused = true;
  return HJ3D_OK;
}

template <class KeyT>
int unnest_impl(hj3d_ctx* c, hj3d_table* t, const uint32_t* left, const uint32_t* gref, const uint2* pairs, uint64_t n, uint32_t flags,
                uint2* out, uint64_t cap, hj3d_counters* res) {
  const NestedIn in{left, gref, pairs};
  const Group<KeyT>* groups = (const Group<KeyT>*)t->groups;
  const bool cs = flags & HJ3D_F_CHECKSUM, wr = out != nullptr;
  // single pass (unnest.cuh): every block reserves its output range on d_ctr->out_cursor; hot tuples are listed and
  // expanded by k_unnest_hot; if the list overflows (pathological skew) the pass is repeated with a list of n entries
  const uint32_t nb = blocks_for(n, kUxTile);
  uint32_t hot_cap = (uint32_t)c->unnest_hot_cap;
  DevCounters* hc = (DevCounters*)c->h_pinned;
  unsigned long long* h_hot = (unsigned long long*)((char*)c->h_pinned + 512);
  for (int attempt = 0; attempt < 2 && nb; ++attempt) {
    uint32_t* hot_list = nullptr;
    HJ_TRY(dev_alloc(c, &hot_list, hot_cap ? hot_cap : 1u));
    CUDA_TRY(cudaMemsetAsync(c->d_scalar, 0, 8, c->stream));
    if (attempt) CUDA_TRY(cudaMemsetAsync(c->d_ctr, 0, sizeof(DevCounters), c->stream));
#define LAUNCH_UX(C, W) k_unnest_expand<KeyT, C, W><<<nb, kUxThreads, 0, c->stream>>>(in, n, groups, t->rows, out, cap, c->d_ctr, hot_list, hot_cap, c->d_scalar)
    if (cs) { if (wr) LAUNCH_UX(true, true); else LAUNCH_UX(true, false); }
    else    { if (wr) LAUNCH_UX(false, true); else LAUNCH_UX(false, false); }
#undef LAUNCH_UX
    ++c->launches;
    CUDA_TRY(cudaMemcpyAsync(h_hot, c->d_scalar, 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    const unsigned long long n_hot = *h_hot;
    if (n_hot > hot_cap) { hot_cap = (uint32_t)n; continue; }             // list overflow: once more with room for every tuple
    if (n_hot) {
      const uint32_t nbh = n_hot < 4096 ? (uint32_t)n_hot : 4096u;
#define LAUNCH_UH(C, W) k_unnest_hot<KeyT, C, W><<<nbh, kUxThreads, 0, c->stream>>>(in, hot_list, (uint32_t)n_hot, groups, t->rows, out, cap, c->d_ctr)
      if (cs) { if (wr) LAUNCH_UH(true, true); else LAUNCH_UH(true, false); }
      else    { if (wr) LAUNCH_UH(false, true); else LAUNCH_UH(false, false); }
#undef LAUNCH_UH
      ++c->launches;
    }
    break;
  }
  CUDA_TRY(cudaMemcpyAsync(hc, c->d_ctr, sizeof(DevCounters), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  CUDA_TRY(cudaGetLastError());
  const unsigned long long total = nb ? hc->out_cursor : 0ull;
  res->matches = total; res->out_tuples = total; res->num_cmps = 0;     // AlgUnnestHt::_count = #outputs (algebra.hh:486-487)
  res->checksum_sum = hc->checksum_sum; res->checksum_xor = hc->checksum_xor;
  res->overflow = (wr && total > cap) ? 1 : 0;
  res->out_written = wr ? (total > cap ? cap : total) : 0;
  return HJ3D_OK;
}


// After the fused nested probe + unnest kernel: expand the listed hot groups, fetch the counters.  *redo = true: the hot list
// overflowed (more than HJ3D_OPT_UNNEST_HOT_CAP hot probe hits): the caller runs the two-operator composition instead.
int finish_fused_unnest(hj3d_ctx* c, hj3d_table* t, uint32_t flags, uint2* out, uint64_t cap, hj3d_counters* probe_out,
                        hj3d_counters* unnest_out, bool* redo) {
  *redo = false;
  unsigned long long* h_hot = (unsigned long long*)((char*)c->h_pinned + 512);
  CUDA_TRY(cudaMemcpyAsync(h_hot, c->d_scalar + 1, 8, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  const unsigned long long n_hot = *h_hot;
  if (n_hot > c->fused_hot_cap) { *redo = true; return HJ3D_OK; }
  const bool cs = flags & HJ3D_F_CHECKSUM, wr = out != nullptr;
  if (n_hot && (cs || wr)) {
    PhaseTimer pt(c, PH_UNNEST);
    const dim3 grid((unsigned)(n_hot < 1024 ? n_hot : 1024), kHotSplit);
    const HotGroup* hl = (const HotGroup*)c->fused_hot_list;
    if (cs) { if (wr) k_unnest_hot_groups<true, true><<<grid, 256, 0, c->stream>>>(hl, (uint32_t)n_hot, t->rows, out, cap, c->d_ctr);
              else    k_unnest_hot_groups<true, false><<<grid, 256, 0, c->stream>>>(hl, (uint32_t)n_hot, t->rows, out, cap, c->d_ctr); }
    else      k_unnest_hot_groups<false, true><<<grid, 256, 0, c->stream>>>(hl, (uint32_t)n_hot, t->rows, out, cap, c->d_ctr);
    ++c->launches;
    CUDA_TRY(cudaGetLastError());
  }
  DevCounters* h = (DevCounters*)c->h_pinned;
  CUDA_TRY(cudaMemcpyAsync(h, c->d_ctr, sizeof(DevCounters), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  CUDA_TRY(cudaGetLastError());
  probe_out->matches = h->matches; probe_out->num_cmps = h->num_cmps; probe_out->out_tuples = h->matches;
  unnest_out->matches = h->out_cursor; unnest_out->out_tuples = h->out_cursor;
  unnest_out->checksum_sum = h->checksum_sum; unnest_out->checksum_xor = h->checksum_xor;
  unnest_out->overflow = (wr && h->out_cursor > cap) ? 1 : 0;
  unnest_out->out_written = wr ? (h->out_cursor > cap ? cap : h->out_cursor) : 0;
  return HJ3D_OK;
}

int table_matches(hj3d_table* t, const hj3d_keyspec& ks) {
  if (!t->built) return fail(HJ3D_ERR_INVALID, "table has not been built");
  if ((int)ks.hash_id != t->hash_id)
    return fail(HJ3D_ERR_INVALID, "probe hash function differs from the build hash function "
                                  "(static_assert in ht_chaining.hh:243 / ht_nested.hh:361)");
  return HJ3D_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------ C ABI
extern "C" {

const char* hj3d_last_error(void) { return hj3d_err_slot().c_str(); }
const char* hj3d_version(void) { return "hj3d 0.1 (sm_100a)"; }
uint64_t hj3d_pair_mix(uint32_t l, uint32_t r) { return pair_mix(l, r); }

int hj3d_ctx_create(int device, hj3d_ctx** out) {
  if (!out) return fail(HJ3D_ERR_INVALID, "out == NULL");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(HJ3D_ERR_CUDA, std::string("no CUDA device (this engine has no CPU fallback): ") + cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return fail(HJ3D_ERR_INVALID, "device index out of range");
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) return fail(HJ3D_ERR_CUDA, "hj3d kernels are built for sm_100a only; device is older");
  CUDA_TRY(cudaSetDevice(device));
  hj3d_ctx* c = new hj3d_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  CUDA_TRY(cudaDeviceGetAttribute(&c->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
  CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  c->own_stream = true;
  CUDA_TRY(cudaDeviceGetDefaultMemPool(&c->pool, device));
  uint64_t thr = UINT64_MAX;
  CUDA_TRY(cudaMemPoolSetAttribute(c->pool, cudaMemPoolAttrReleaseThreshold, &thr));
  for (int i = 0; i < PH_COUNT; ++i) { CUDA_TRY(cudaEventCreate(&c->ev[i][0])); CUDA_TRY(cudaEventCreate(&c->ev[i][1])); c->ev_used[i] = false; }
  CUDA_TRY(cudaEventCreate(&c->ev_total[0])); CUDA_TRY(cudaEventCreate(&c->ev_total[1]));
  CUDA_TRY(cudaMalloc((void**)&c->d_ctr, sizeof(DevCounters)));
  CUDA_TRY(cudaMalloc((void**)&c->d_stats, kStatsCopies * sizeof(DevStats)));
  CUDA_TRY(cudaMalloc((void**)&c->d_scalar, 4 * sizeof(unsigned long long)));
  CUDA_TRY(cudaMallocHost(&c->h_pinned, 32768));
  *out = c;
  return HJ3D_OK;
}

int hj3d_ctx_destroy(hj3d_ctx* c) {
  if (!c) return HJ3D_OK;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  if (c->host_comm) hj3d_comm_destroy(c->host_comm);
  if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
  for (cudaEvent_t e : c->chunk_ev) cudaEventDestroy(e);
  for (int i = 0; i < PH_COUNT; ++i) { cudaEventDestroy(c->ev[i][0]); cudaEventDestroy(c->ev[i][1]); }
  cudaEventDestroy(c->ev_total[0]); cudaEventDestroy(c->ev_total[1]);
  cudaFree(c->d_ctr); cudaFree(c->d_stats); cudaFree(c->d_scalar); cudaFreeHost(c->h_pinned);
  for (auto& k : c->arena) cudaFree(k.base);
  cudaFree(c->hj.b); cudaFree(c->hj.p); cudaFree(c->hj.out); cudaFree(c->hj.nest); cudaFree(c->hj.l); cudaFree(c->hj.g);
  if (c->own_stream) cudaStreamDestroy(c->stream);
  delete c;
  return HJ3D_OK;
}

int hj3d_ctx_set_stream(hj3d_ctx* c, void* s) {
  if (!c) return fail(HJ3D_ERR_INVALID, "ctx == NULL");
  cudaStreamSynchronize(c->stream);
  if (c->own_stream) { cudaStreamDestroy(c->stream); c->own_stream = false; }
  if (s) c->stream = (cudaStream_t)s;
  else { CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)); c->own_stream = true; }
  return HJ3D_OK;
}

int hj3d_ctx_set_option(hj3d_ctx* c, int opt, int64_t v) {
  if (!c) return fail(HJ3D_ERR_INVALID, "ctx == NULL");
  switch (opt) {
    case HJ3D_OPT_WARP_AGGREGATE:   c->warp_aggregate = v; break;
    case HJ3D_OPT_PARTITION_BYTES:  c->partition_bytes = v; break;
    case HJ3D_OPT_PARTITION_WINDOW: c->partition_window = v > 0 ? v : c->partition_window; break;
    case HJ3D_OPT_PARTITION_MIN_PROBE: c->partition_min_probe = v; break;
    case HJ3D_OPT_SMEM_PROBE: c->smem_probe = v; break;
    case HJ3D_OPT_SMEM_BUILD: c->smem_build = v; break;
    case HJ3D_OPT_SMEM_BUILD_BYTES: if (v >= 2048 && v <= (160 << 10)) c->smem_build_bytes = v; break;
    case HJ3D_OPT_SMEM_SLICE_BYTES: if (v >= 4096 && v <= (200 << 10)) c->smem_slice_bytes = v & ~15ll; break;
    case HJ3D_OPT_SMEM_MIN_PROBE: c->smem_min_probe = v; break;
    case HJ3D_OPT_SMEM_CHUNK: if (v >= 2048) c->smem_chunk = v; break;
    case HJ3D_OPT_PART_THREADS: if (v == 256 || v == 512 || v == 1024) c->part_threads = v; break;
    case HJ3D_OPT_PART_RANK_MATCH: c->part_rank_match = v != 0; break;
    case HJ3D_OPT_PROBE_THREADS: if (v == 256 || v == 512) c->probe_threads = v; break;
    case HJ3D_OPT_LEAN_PROBE: c->lean_probe = v != 0; break;
    case HJ3D_OPT_HOST_CHUNK_BYTES: if (v >= 0) c->host_chunk_bytes = v; break;
    case HJ3D_OPT_PACKED_PROBE: c->packed_probe = v != 0; break;
    case HJ3D_OPT_PACKED_MIN_PROBE: c->packed_min_probe = v; break;
    case HJ3D_OPT_PACKED_SLICE_BYTES: if (v >= 1024 && v <= (110 << 10)) c->packed_slice_bytes = v & ~15ll; break;
    case HJ3D_OPT_PART_SAMPLE: if (v >= 0 && v <= 2) c->part_sample = v; break;
    case HJ3D_OPT_UNNEST_HOT_CAP: if (v >= 0 && v <= (1ll << 30)) c->unnest_hot_cap = v; break;
    default: return fail(HJ3D_ERR_INVALID, "unknown option");
  }
  return HJ3D_OK;
}

int hj3d_ctx_sync(hj3d_ctx* c) {
  if (!c) return fail(HJ3D_ERR_INVALID, "ctx == NULL");
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return HJ3D_OK;
}

int hj3d_ctx_timings(hj3d_ctx* c, hj3d_timings* out) {
  if (!c || !out) return fail(HJ3D_ERR_INVALID, "NULL argument");
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  float v[PH_COUNT];
  for (int i = 0; i < PH_COUNT; ++i) {
    v[i] = 0.f;
    if (c->ev_used[i]) cudaEventElapsedTime(&v[i], c->ev[i][0], c->ev[i][1]);
  }
  out->partition_ms = v[PH_PARTITION]; out->histogram_ms = v[PH_HIST]; out->scan_ms = v[PH_SCAN];
  out->scatter_ms = v[PH_SCATTER]; out->group_ms = v[PH_GROUP]; out->probe_ms = v[PH_PROBE]; out->unnest_ms = v[PH_UNNEST];
  out->partition_l1_ms = v[PH_PART1];
  out->total_ms = 0.f;
  cudaEventElapsedTime(&out->total_ms, c->ev_total[0], c->ev_total[1]);
  cudaGetLastError();
  out->kernel_launches = c->launches;
  return HJ3D_OK;
}

int hj3d_mem_alloc(hj3d_ctx* c, uint64_t bytes, void** out) {
  if (!c || !out) return fail(HJ3D_ERR_INVALID, "NULL argument");
  CUDA_TRY(cudaSetDevice(c->device));
  return raw_alloc(out, bytes ? bytes : 256);
}
int hj3d_mem_free(hj3d_ctx* c, void* p) {
  if (!c) return fail(HJ3D_ERR_INVALID, "NULL argument");
  if (!p) return HJ3D_OK;
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  CUDA_TRY(cudaFree(p));
  return HJ3D_OK;
}
int hj3d_memcpy_h2d(hj3d_ctx* c, void* d, const void* h, uint64_t bytes) {
  if (!c || (bytes && (!d || !h))) return fail(HJ3D_ERR_INVALID, "NULL argument");
  CUDA_TRY(cudaSetDevice(c->device));
  if (bytes) CUDA_TRY(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, c->stream));
  return HJ3D_OK;
}
int hj3d_memcpy_d2h(hj3d_ctx* c, void* h, const void* d, uint64_t bytes) {
  if (!c || (bytes && (!d || !h))) return fail(HJ3D_ERR_INVALID, "NULL argument");
  CUDA_TRY(cudaSetDevice(c->device));
  if (bytes) CUDA_TRY(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return HJ3D_OK;
}
int hj3d_iota_u32(hj3d_ctx* c, uint32_t* d, uint64_t n, uint32_t first) {
  if (!c || (n && !d)) return fail(HJ3D_ERR_INVALID, "NULL argument");
  if (!n) return HJ3D_OK;
  CUDA_TRY(cudaSetDevice(c->device));
  k_iota_u32<<<blocks_for(n, 256), 256, 0, c->stream>>>(d, n, first);
  ++c->launches;
  CUDA_TRY(cudaGetLastError());
  return HJ3D_OK;
}

int hj3d_owner_range(uint64_t D, uint32_t n_owners, uint32_t owner, uint64_t* lo, uint64_t* hi) {
  if (!D || !n_owners || owner >= n_owners || !lo || !hi) return fail(HJ3D_ERR_INVALID, "bad owner range arguments");
  const uint64_t w = (D + n_owners - 1) / n_owners;
  *lo = (uint64_t)owner * w < D ? (uint64_t)owner * w : D;
  *hi = (uint64_t)(owner + 1) * w < D ? (uint64_t)(owner + 1) * w : D;
  return HJ3D_OK;
}

int hj3d_table_create_shard(hj3d_ctx* c, int kind, uint64_t D, uint64_t lo, uint64_t hi, hj3d_table** out) {
  if (!c || !out) return fail(HJ3D_ERR_INVALID, "NULL argument");
  *out = nullptr;
  if (kind != HJ3D_CHAINING && kind != HJ3D_NESTED) return fail(HJ3D_ERR_INVALID, "unknown table kind");
  if (D == 0) return fail(HJ3D_ERR_INVALID, "num_buckets must be >= 1 (h % 0 is undefined in the reference too)");
  if (D > 0xFFFFFFFFull) return fail(HJ3D_ERR_UNSUPPORTED, "num_buckets > 2^32-1 (the reference drivers use uint32_t, main_experiment1.cc:651)");
  if (lo > hi || hi > D) return fail(HJ3D_ERR_INVALID, "bad bucket range");
  hj3d_table* t = new hj3d_table();
  t->kind = kind; t->D = D; t->blo = lo; t->bhi = hi;
  t->dir = make_dir(D, lo, hi);
  *out = t;
  return HJ3D_OK;
}

int hj3d_table_create(hj3d_ctx* c, int kind, uint64_t D, hj3d_table** out) {
  return hj3d_table_create_shard(c, kind, D, 0, D, out);
}

int hj3d_table_build(hj3d_ctx* c, hj3d_table* t, const void* d_tuples, uint64_t n, hj3d_keyspec ks) {
  if (!c || !t) return fail(HJ3D_ERR_INVALID, "NULL argument");
  if (t->built) return fail(HJ3D_ERR_INVALID, "table is not empty: call hj3d_table_clear first (bulk build)");
  if (n && !d_tuples) return fail(HJ3D_ERR_INVALID, "d_tuples == NULL");
  if (n > 0xFFFFFFF0ull) return fail(HJ3D_ERR_UNSUPPORTED, "more than 2^32-16 build tuples (row ids are 32 bit; reference cap is 2^30, main_experiment1.cc:1391)");
  HJ_TRY(check_keyspec(ks));
  CUDA_TRY(cudaSetDevice(c->device));
  HJ_TRY(arena_reset(c));
  begin_call(c);
  Src src = make_src(d_tuples, n, ks, nullptr);
  t->hash_id = (int)ks.hash_id; t->key_bytes = ks.key_bytes;
  int rc;
  switch (ks.hash_id) {
    case HJ3D_HASH_MURMUR32: rc = build_impl<HJ3D_HASH_MURMUR32>(c, t, src); break;
    case HJ3D_HASH_MURMUR64: rc = build_impl<HJ3D_HASH_MURMUR64>(c, t, src); break;
    default:                 rc = build_impl<HJ3D_HASH_MURMUR64_SEXT32>(c, t, src); break;
  }
  end_call(c);
  if (rc < 0) { clear_table(t); return rc; }
  // row ids are positions unless the tuples carry their own (then only the caller can bound them)
  t->rowid_bound = ks.rowid_offset == HJ3D_NO_ROWID ? (n ? n : 1) : t->rowid_bound_user;
  set_packed_geometry(c, t);
  return HJ3D_OK;
}

int hj3d_table_set_rowid_bound(hj3d_ctx* c, hj3d_table* t, uint64_t bound) {
  if (!c || !t) return fail(HJ3D_ERR_INVALID, "NULL argument");
  t->rowid_bound_user = bound;
  if (t->built && t->hash_id >= 0) { /* takes effect at the next build */ }
  return HJ3D_OK;
}

int hj3d_table_clear(hj3d_ctx* c, hj3d_table* t) {
  if (!c || !t) return fail(HJ3D_ERR_INVALID, "NULL argument");
  clear_table(t);
  return HJ3D_OK;
}

int hj3d_table_destroy(hj3d_ctx* c, hj3d_table* t) {
  if (!t) return HJ3D_OK;
  if (c) free_table_arrays(c, t);
  delete t;
  return HJ3D_OK;
}

int hj3d_table_size(hj3d_table* t, uint64_t* n, uint64_t* g) {
  if (!t) return fail(HJ3D_ERR_INVALID, "NULL argument");
  if (n) *n = t->n;
  if (g) *g = t->n_groups;
  return HJ3D_OK;
}

int hj3d_table_stats(hj3d_ctx* c, hj3d_table* t, hj3d_stats* s) {
  if (!c || !t || !s) return fail(HJ3D_ERR_INVALID, "NULL argument");
  memset(s, 0, sizeof(*s));
  const uint64_t nl = t->dir.n_local;
  s->num_buckets = nl;
  if (!t->built) {   // an empty table: every bucket is empty (makeStatistics on a cleared table)
    s->num_empty = nl; s->cc_min = nl ? 0 : ~0ull; s->cc_count = nl; s->ccne_min = ~0ull;
    s->mem_dir = nl * (t->kind == HJ3D_CHAINING ? 24 : 32);
    return HJ3D_OK;
  }
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  const DevStats& h = t->hstats;
  s->num_empty = h.empty; s->num_entries = t->n;
  s->cc_min = h.all.mn; s->cc_max = h.all.mx; s->cc_sum = h.all.sum; s->cc_sumsq = h.all.sumsq; s->cc_count = h.all.cnt;
  s->ccne_min = h.nonempty.mn; s->ccne_max = h.nonempty.mx; s->ccne_sum = h.nonempty.sum;
  s->ccne_sumsq = h.nonempty.sumsq; s->ccne_count = h.nonempty.cnt;
  const uint64_t nonempty = nl - h.empty;
  if (t->kind == HJ3D_CHAINING) {
    // distinct low-32 hash bits via a 2^32-bit bitmap (only here, outside any timed path; the
    // reference's makeStatistics is likewise called after the timed region, main_experiment1.cc:710)
    uint32_t* bitmap = nullptr;
    const uint64_t words = 1ull << 27;
    HJ_TRY(arena_reset(c));
    HJ_TRY(dev_alloc(c, &bitmap, words));
    CUDA_TRY(cudaMemsetAsync(bitmap, 0, words * 4, c->stream));
    CUDA_TRY(cudaMemsetAsync(c->d_scalar, 0, 8, c->stream));
    if (t->n) {
      const uint32_t nb = blocks_for(t->n, 256);
      switch (t->hash_id) {
        case HJ3D_HASH_MURMUR32: k_hash_bitmap<HJ3D_HASH_MURMUR32><<<nb, 256, 0, c->stream>>>((const Slot<uint32_t>*)t->slots, t->n, bitmap); break;
        case HJ3D_HASH_MURMUR64: k_hash_bitmap<HJ3D_HASH_MURMUR64><<<nb, 256, 0, c->stream>>>((const Slot<uint64_t>*)t->slots, t->n, bitmap); break;
        default: k_hash_bitmap<HJ3D_HASH_MURMUR64_SEXT32><<<nb, 256, 0, c->stream>>>((const Slot<uint32_t>*)t->slots, t->n, bitmap); break;
      }
      k_popcount<<<c->sm_count * 8, 256, 0, c->stream>>>(bitmap, words, c->d_scalar);
      c->launches += 2;
    }
    unsigned long long dk = 0;
    CUDA_TRY(cudaMemcpyAsync(&dk, c->d_scalar, 8, cudaMemcpyDeviceToHost, c->stream));
    dev_free(c, bitmap);
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    s->num_distinct_keys = dk;
    s->rsv_main = t->n - nonempty;                       // every tuple but the first of a bucket takes a reservoir node
    s->rsv_sub = 0;
    s->mem_dir = nl * 24; s->mem_main = s->rsv_main * 24; s->mem_sub = 0;
  } else {
    s->num_distinct_keys = t->n_groups;                  // Sum of main chain lengths (ht_nested.hh:470)
    s->rsv_main = t->n_groups - nonempty;
    s->rsv_sub = t->n - t->n_groups;
    s->mem_dir = nl * 32; s->mem_main = s->rsv_main * 32; s->mem_sub = s->rsv_sub * 16;
  }
  CUDA_TRY(cudaGetLastError());
  return HJ3D_OK;
}

int hj3d_stats_merge(const hj3d_stats* p, uint32_t n, hj3d_stats* o) {
  if (!p || !o || !n) return fail(HJ3D_ERR_INVALID, "NULL argument");
  memset(o, 0, sizeof(*o));
  o->cc_min = ~0ull; o->ccne_min = ~0ull;
  for (uint32_t i = 0; i < n; ++i) {
    const hj3d_stats& s = p[i];
    o->num_buckets += s.num_buckets; o->num_empty += s.num_empty; o->num_entries += s.num_entries;
    o->num_distinct_keys += s.num_distinct_keys;   // exact for nested; for chaining exact when the hash is injective on 32 bits
    if (s.cc_count)   { o->cc_min = s.cc_min < o->cc_min ? s.cc_min : o->cc_min; o->cc_max = s.cc_max > o->cc_max ? s.cc_max : o->cc_max; }
    if (s.ccne_count) { o->ccne_min = s.ccne_min < o->ccne_min ? s.ccne_min : o->ccne_min; o->ccne_max = s.ccne_max > o->ccne_max ? s.ccne_max : o->ccne_max; }
    o->cc_sum += s.cc_sum; o->cc_sumsq += s.cc_sumsq; o->cc_count += s.cc_count;
    o->ccne_sum += s.ccne_sum; o->ccne_sumsq += s.ccne_sumsq; o->ccne_count += s.ccne_count;
    o->rsv_main += s.rsv_main; o->rsv_sub += s.rsv_sub;
    o->mem_dir += s.mem_dir; o->mem_main += s.mem_main; o->mem_sub += s.mem_sub;
  }
  return HJ3D_OK;
}

int hj3d_probe_chaining(hj3d_ctx* c, hj3d_table* t, const void* d_probe, uint64_t n, hj3d_keyspec ks,
                        const uint32_t* d_gather, int unique, uint32_t flags,
                        uint32_t* d_out, uint64_t cap, hj3d_counters* out) {
  if (!c || !t || !out) return fail(HJ3D_ERR_INVALID, "NULL argument");
  if (t->kind != HJ3D_CHAINING) return fail(HJ3D_ERR_INVALID, "hj3d_probe_chaining needs a chaining table");
  if (n && !d_probe) return fail(HJ3D_ERR_INVALID, "d_probe == NULL");
  if (n > 0xFFFFFFF0ull) return fail(HJ3D_ERR_UNSUPPORTED, "more than 2^32-16 probe tuples");
  HJ_TRY(check_keyspec(ks));
  HJ_TRY(table_matches(t, ks));
  CUDA_TRY(cudaSetDevice(c->device));
  HJ_TRY(arena_reset(c));
  begin_call(c);
  CUDA_TRY(cudaMemsetAsync(c->d_ctr, 0, sizeof(DevCounters), c->stream));
  Src src = make_src(d_probe, n, ks, d_gather);
  int rc;
  {
    switch (ks.hash_id) {
      case HJ3D_HASH_MURMUR32: rc = probe_chaining_impl<HJ3D_HASH_MURMUR32>(c, t, src, unique != 0, flags, (uint2*)d_out, cap); break;
      case HJ3D_HASH_MURMUR64: rc = probe_chaining_impl<HJ3D_HASH_MURMUR64>(c, t, src, unique != 0, flags, (uint2*)d_out, cap); break;
      default:                 rc = probe_chaining_impl<HJ3D_HASH_MURMUR64_SEXT32>(c, t, src, unique != 0, flags, (uint2*)d_out, cap); break;
    }
  }
  end_call(c);
  if (rc < 0) return rc;
  HJ_TRY(fetch_counters(c, out, cap, d_out != nullptr));
  return out->overflow ? HJ3D_OVERFLOW : HJ3D_OK;
}

int hj3d_probe_nested(hj3d_ctx* c, hj3d_table* t, const void* d_probe, uint64_t n, hj3d_keyspec ks,
                      const uint32_t* d_gather, uint32_t flags, uint32_t* d_out, uint64_t cap, hj3d_counters* out) {
  if (!c || !t || !out) return fail(HJ3D_ERR_INVALID, "NULL argument");
  if (t->kind != HJ3D_NESTED) return fail(HJ3D_ERR_INVALID, "hj3d_probe_nested needs a nested table");
  if (n && !d_probe) return fail(HJ3D_ERR_INVALID, "d_probe == NULL");
  if (n > 0xFFFFFFF0ull) return fail(HJ3D_ERR_UNSUPPORTED, "more than 2^32-16 probe tuples");
  HJ_TRY(check_keyspec(ks));
  HJ_TRY(table_matches(t, ks));
  CUDA_TRY(cudaSetDevice(c->device));
  HJ_TRY(arena_reset(c));
  begin_call(c);
  CUDA_TRY(cudaMemsetAsync(c->d_ctr, 0, sizeof(DevCounters), c->stream));
  Src src = make_src(d_probe, n, ks, d_gather);
  int rc;
  {
    switch (ks.hash_id) {
      case HJ3D_HASH_MURMUR32: rc = probe_nested_impl<HJ3D_HASH_MURMUR32>(c, t, src, flags, (uint2*)d_out, cap); break;
      case HJ3D_HASH_MURMUR64: rc = probe_nested_impl<HJ3D_HASH_MURMUR64>(c, t, src, flags, (uint2*)d_out, cap); break;
      default:                 rc = probe_nested_impl<HJ3D_HASH_MURMUR64_SEXT32>(c, t, src, flags, (uint2*)d_out, cap); break;
    }
  }
  end_call(c);
  if (rc < 0) return rc;
  HJ_TRY(fetch_counters(c, out, cap, d_out != nullptr));
  return out->overflow ? HJ3D_OVERFLOW : HJ3D_OK;
}

int hj3d_unnest(hj3d_ctx* c, hj3d_table* t, const uint32_t* d_left, const uint32_t* d_gref, uint64_t n,
                uint32_t flags, uint32_t* d_out, uint64_t cap, hj3d_counters* out) {
  if (!c || !t || !out) return fail(HJ3D_ERR_INVALID, "NULL argument");
  if (t->kind != HJ3D_NESTED) return fail(HJ3D_ERR_INVALID, "hj3d_unnest needs a nested table");
  if (!t->built) return fail(HJ3D_ERR_INVALID, "table has not been built");
  if (n && (!d_left || !d_gref)) return fail(HJ3D_ERR_INVALID, "NULL input column");
  CUDA_TRY(cudaSetDevice(c->device));
  HJ_TRY(arena_reset(c));
  begin_call(c);
  CUDA_TRY(cudaMemsetAsync(c->d_ctr, 0, sizeof(DevCounters), c->stream));
  memset(out, 0, sizeof(*out));
  int rc;
  {
    PhaseTimer pt(c, PH_UNNEST);
    if (t->key_bytes == 8) rc = unnest_impl<uint64_t>(c, t, d_left, d_gref, nullptr, n, flags, (uint2*)d_out, cap, out);
    else                   rc = unnest_impl<uint32_t>(c, t, d_left, d_gref, nullptr, n, flags, (uint2*)d_out, cap, out);
  }
  end_call(c);
  if (rc < 0) return rc;
  return out->overflow ? HJ3D_OVERFLOW : HJ3D_OK;
}

int hj3d_unnest_pairs(hj3d_ctx* c, hj3d_table* t, const uint32_t* d_nested_pairs, uint64_t n,
                      uint32_t flags, uint32_t* d_out, uint64_t cap, hj3d_counters* out) {
  if (!c || !t || !out) return fail(HJ3D_ERR_INVALID, "NULL argument");
  if (t->kind != HJ3D_NESTED) return fail(HJ3D_ERR_INVALID, "hj3d_unnest_pairs needs a nested table");
  if (!t->built) return fail(HJ3D_ERR_INVALID, "table has not been built");
  if (n && !d_nested_pairs) return fail(HJ3D_ERR_INVALID, "NULL input");
  CUDA_TRY(cudaSetDevice(c->device));
  HJ_TRY(arena_reset(c));
  begin_call(c);
  CUDA_TRY(cudaMemsetAsync(c->d_ctr, 0, sizeof(DevCounters), c->stream));
  memset(out, 0, sizeof(*out));
  int rc;
  {
    PhaseTimer pt(c, PH_UNNEST);
    if (t->key_bytes == 8) rc = unnest_impl<uint64_t>(c, t, nullptr, nullptr, (const uint2*)d_nested_pairs, n, flags, (uint2*)d_out, cap, out);
    else                   rc = unnest_impl<uint32_t>(c, t, nullptr, nullptr, (const uint2*)d_nested_pairs, n, flags, (uint2*)d_out, cap, out);
  }
  end_call(c);
  if (rc < 0) return rc;
  return out->overflow ? HJ3D_OVERFLOW : HJ3D_OK;
}

int hj3d_probe_nested_unnest(hj3d_ctx* c, hj3d_table* t, const void* d_probe, uint64_t n, hj3d_keyspec ks, uint32_t flags,
                             uint32_t* d_out, uint64_t cap, hj3d_counters* probe_out, hj3d_counters* unnest_out) {
  if (!c || !t || !probe_out || !unnest_out) return fail(HJ3D_ERR_INVALID, "NULL argument");
  if (t->kind != HJ3D_NESTED) return fail(HJ3D_ERR_INVALID, "hj3d_probe_nested_unnest needs a nested table");
  if (n && !d_probe) return fail(HJ3D_ERR_INVALID, "d_probe == NULL");
  if (n > 0xFFFFFFF0ull) return fail(HJ3D_ERR_UNSUPPORTED, "more than 2^32-16 probe tuples");
  HJ_TRY(check_keyspec(ks));
  HJ_TRY(table_matches(t, ks));
  CUDA_TRY(cudaSetDevice(c->device));
  HJ_TRY(arena_reset(c));
  begin_call(c);
  CUDA_TRY(cudaMemsetAsync(c->d_ctr, 0, sizeof(DevCounters), c->stream));
  memset(probe_out, 0, sizeof(*probe_out)); memset(unnest_out, 0, sizeof(*unnest_out));
  Src src = make_src(d_probe, n, ks, nullptr);
  bool fused = false;
  int rc;
  switch (ks.hash_id) {
    case HJ3D_HASH_MURMUR32: rc = probe_nested_unnest_impl<HJ3D_HASH_MURMUR32>(c, t, src, flags, (uint2*)d_out, cap, &fused); break;
    case HJ3D_HASH_MURMUR64: rc = probe_nested_unnest_impl<HJ3D_HASH_MURMUR64>(c, t, src, flags, (uint2*)d_out, cap, &fused); break;
    default:                 rc = probe_nested_unnest_impl<HJ3D_HASH_MURMUR64_SEXT32>(c, t, src, flags, (uint2*)d_out, cap, &fused); break;
  }
  end_call(c);
  if (rc < 0) return rc;
  if (fused) {
    bool redo = false;
    HJ_TRY(finish_fused_unnest(c, t, flags, (uint2*)d_out, cap, probe_out, unnest_out, &redo));
    end_call(c);
    if (!redo) return unnest_out->overflow ? HJ3D_OVERFLOW : HJ3D_OK;
    memset(probe_out, 0, sizeof(*probe_out)); memset(unnest_out, 0, sizeof(*unnest_out));
  }
  // composition: nested tuples into a ctx-owned buffer, then the unnest of the pairs
  auto& hj = c->hj;
  const size_t need = (n ? n : 1) * 8;
  if (hj.cnest < need) {
    if (hj.nest) { cudaStreamSynchronize(c->stream); cudaFree(hj.nest); hj.nest = nullptr; hj.cnest = 0; }
    HJ_TRY(raw_alloc(&hj.nest, need));
    hj.cnest = need;
  }
  rc = hj3d_probe_nested(c, t, d_probe, n, ks, nullptr, flags & ~HJ3D_F_CHECKSUM, (uint32_t*)hj.nest, n, probe_out);
  if (rc < 0) return rc;
  return hj3d_unnest_pairs(c, t, (const uint32_t*)hj.nest, probe_out->out_written, flags, d_out, cap, unnest_out);
}

// ---- build / probe continuing from an exchanged (coarse-partitioned) relation: exchange.cu ---------------------------
static Src parts_src(const hj3d_parts* p) {
  Src s; s.base = (const uint8_t*)p->recs; s.gather = nullptr; s.n = p->n_total; s.stride = p->key_bytes == 8 ? 16 : 8;
  s.key_off = 0; s.rowid_off = p->key_bytes;
  return s;
}
static hj3d_keyspec parts_ks(const hj3d_parts* p) {
  hj3d_keyspec ks; ks.tuple_bytes = p->key_bytes == 8 ? 16 : 8; ks.key_offset = 0; ks.key_bytes = p->key_bytes; ks.hash_id = p->hash_id;
  ks.rowid_offset = p->key_bytes;
  return ks;
}
static int parts_fit_table(const hj3d_parts* p, const hj3d_table* t) {
  if (p->overflow) return fail(HJ3D_ERR_INVALID, "the exchange overflowed a receive region: reserve more and exchange again");
  if (p->D != t->D || p->bucket_lo != t->blo || p->bucket_hi != t->bhi)
    return fail(HJ3D_ERR_INVALID, "the exchanged relation was partitioned for another directory / bucket range than the table's (hj3d_comm_shard)");
  return HJ3D_OK;
}

int hj3d_table_build_parts(hj3d_ctx* c, hj3d_table* t, hj3d_parts* p) {
  if (!c || !t || !p) return fail(HJ3D_ERR_INVALID, "NULL argument");
  if (t->built) return fail(HJ3D_ERR_INVALID, "table is not empty: call hj3d_table_clear first (bulk build)");
  HJ_TRY(parts_fit_table(p, t));
  CUDA_TRY(cudaSetDevice(c->device));
  HJ_TRY(arena_reset(c));
  begin_call(c);
  const PartsView pv{p->d_start, p->d_count, p->n_ranges * p->n_src, p->range_width};
  const Src src = parts_src(p);
  t->hash_id = (int)p->hash_id; t->key_bytes = p->key_bytes;
  int rc;
  switch (p->hash_id) {
    case HJ3D_HASH_MURMUR32: rc = build_impl<HJ3D_HASH_MURMUR32>(c, t, src, &pv); break;
    case HJ3D_HASH_MURMUR64: rc = build_impl<HJ3D_HASH_MURMUR64>(c, t, src, &pv); break;
    default:                 rc = build_impl<HJ3D_HASH_MURMUR64_SEXT32>(c, t, src, &pv); break;
  }
  end_call(c);
  if (rc < 0) { clear_table(t); return rc; }
  t->rowid_bound = p->rowid_bound ? p->rowid_bound : t->rowid_bound_user;   // global row ids: bounded by the global relation size
  set_packed_geometry(c, t);
  return HJ3D_OK;
}

// hot-key replication, step 3: every rank looks the hot keys up in its shard; one all-reduce hands everybody the owners' answers
int hj3d_parts_hot_answers(hj3d_ctx* c, hj3d_table* t, hj3d_parts* p, int mode) {
  if (!c || !t || !p) return fail(HJ3D_ERR_INVALID, "NULL argument");
  if (!p->comm) return fail(HJ3D_ERR_INVALID, "the relation was not exchanged with HJ3D_XCHG_HOT");
  if (mode != 0 && mode != 1 && mode != 3) return fail(HJ3D_ERR_UNSUPPORTED, "hot-key replication serves probe modes 0, 1 and 3 (a nested tuple's group reference is local to its shard)");
  if ((mode <= 1) != (t->kind == HJ3D_CHAINING)) return fail(HJ3D_ERR_INVALID, "probe mode does not match the table kind");
  HJ_TRY(parts_fit_table(p, t));
  HJ_TRY(table_matches(t, parts_ks(p)));
  CUDA_TRY(cudaSetDevice(c->device));
  HotAnswers* ans = (HotAnswers*)hj3d_comm_hot_ans_buffer(p->comm, p->slot);
  CUDA_TRY(cudaMemsetAsync(ans, 0, sizeof(HotAnswers), c->stream));
#define HJ_ANS(H) do { using KeyT = typename HashT<H>::key_t; const HotTable<KeyT>* ht = (const HotTable<KeyT>*)p->hot_table; \
    if (mode == 0)      k_hot_answers_chaining<H, false><<<1, kHotMax, 0, c->stream>>>(ht, t->dir, t->off, (const Slot<KeyT>*)t->slots, ans); \
    else if (mode == 1) k_hot_answers_chaining<H, true><<<1, kHotMax, 0, c->stream>>>(ht, t->dir, t->off, (const Slot<KeyT>*)t->slots, ans); \
    else                k_hot_answers_nested<H><<<1, kHotMax, 0, c->stream>>>(ht, t->dir, t->goff, (const Group<KeyT>*)t->groups, t->rows, ans); } while (0)
  if (t->n) {
    switch (p->hash_id) {
      case HJ3D_HASH_MURMUR32: HJ_ANS(HJ3D_HASH_MURMUR32); break;
      case HJ3D_HASH_MURMUR64: HJ_ANS(HJ3D_HASH_MURMUR64); break;
      default:                 HJ_ANS(HJ3D_HASH_MURMUR64_SEXT32); break;
    }
    ++c->launches;
  }
#undef HJ_ANS
  CUDA_TRY(cudaGetLastError());
  HJ_TRY(hj3d_comm_hot_reduce_begin(p->comm, p->slot));
  p->hot_mode = mode;
  return HJ3D_OK;
}

int hj3d_parts_hot(hj3d_parts* p, uint64_t* n_hot) {
  if (!p || !n_hot) return fail(HJ3D_ERR_INVALID, "NULL argument");
  *n_hot = p->hot_count;
  return HJ3D_OK;
}

// step 4: the local hot segment against the all-reduced answers; its counters are added to what the regular probe reported
static int hot_join(hj3d_ctx* c, hj3d_parts* p, int mode, uint32_t flags, uint2* out, uint64_t cap, hj3d_counters* probe_out, hj3d_counters* unnest_out) {
  if (p->hot_mode != mode) return fail(HJ3D_ERR_INVALID, "hj3d_parts_hot_answers has not been called for this probe mode");
  const void* d_sum = nullptr;
  HJ_TRY(hj3d_comm_hot_reduce_end(p->comm, p->slot, &d_sum));
  const HotAnswers* ans = (const HotAnswers*)d_sum;
  hj3d_counters* res = mode == 3 ? unnest_out : probe_out;
  const bool cs = flags & HJ3D_F_CHECKSUM, wr = out != nullptr;
  const uint64_t base = wr ? res->out_written : 0;
  DevCounters* h = (DevCounters*)c->h_pinned;
  memset(h, 0, sizeof(DevCounters)); h->out_cursor = base;
  CUDA_TRY(cudaMemcpyAsync(c->d_ctr, h, sizeof(DevCounters), cudaMemcpyHostToDevice, c->stream));
  const uint32_t nb = (uint32_t)std::min<uint64_t>(blocks_for(p->hot_count, 256), (uint64_t)c->sm_count * 8);
#define HJ_HJ(H, N, C, W) k_hot_join<H, N, C, W><<<nb, 256, 0, c->stream>>>((const Slot<typename HashT<H>::key_t>*)p->hot_recs, p->hot_count, \
    (const HotTable<typename HashT<H>::key_t>*)p->hot_table, ans, out, cap, c->d_ctr)
#define HJ_HJ2(H, N) do { if (cs) { if (wr) HJ_HJ(H, N, true, true); else HJ_HJ(H, N, true, false); } \
                          else    { if (wr) HJ_HJ(H, N, false, true); else HJ_HJ(H, N, false, false); } } while (0)
#define HJ_HJ3(H) do { if (mode == 3) HJ_HJ2(H, true); else HJ_HJ2(H, false); } while (0)
  switch (p->hash_id) {
    case HJ3D_HASH_MURMUR32: HJ_HJ3(HJ3D_HASH_MURMUR32); break;
    case HJ3D_HASH_MURMUR64: HJ_HJ3(HJ3D_HASH_MURMUR64); break;
    default:                 HJ_HJ3(HJ3D_HASH_MURMUR64_SEXT32); break;
  }
#undef HJ_HJ3
#undef HJ_HJ2
#undef HJ_HJ
  ++c->launches;
  CUDA_TRY(cudaGetLastError());
  uint32_t* h_many = (uint32_t*)((char*)c->h_pinned + 512);
  CUDA_TRY(cudaMemcpyAsync(h, c->d_ctr, sizeof(DevCounters), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaMemcpyAsync(h_many, &ans->too_many, 4, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  if (*h_many) return fail(HJ3D_ERR_UNSUPPORTED, "a hot key has more than 8 build partners: exchange the probe side without HJ3D_XCHG_HOT");
  probe_out->matches += h->matches; probe_out->out_tuples += h->matches; probe_out->num_cmps += h->num_cmps;
  const uint64_t flat = h->out_cursor - base;
  if (mode == 3) { unnest_out->matches += flat; unnest_out->out_tuples += flat; }
  res->checksum_sum += h->checksum_sum; res->checksum_xor ^= h->checksum_xor;
  if (wr) {
    if (h->out_cursor > cap) res->overflow = 1;
    res->out_written = h->out_cursor > cap ? cap : h->out_cursor;
  }
  return HJ3D_OK;
}

static int probe_parts_regular(hj3d_ctx* c, hj3d_table* t, hj3d_parts* p, int mode, uint32_t flags, uint32_t* d_out, uint64_t cap,
                               hj3d_counters* probe_out, hj3d_counters* unnest_out);

int hj3d_probe_parts(hj3d_ctx* c, hj3d_table* t, hj3d_parts* p, int mode, uint32_t flags, uint32_t* d_out, uint64_t cap,
                     hj3d_counters* probe_out, hj3d_counters* unnest_out) {
  if (p && p->comm && mode == 2) return fail(HJ3D_ERR_UNSUPPORTED, "hot-key replication serves probe modes 0, 1 and 3");
  if (p && p->comm && p->hot_mode != mode) return fail(HJ3D_ERR_INVALID, "hj3d_parts_hot_answers has not been called for this probe mode");
  int rc = probe_parts_regular(c, t, p, mode, flags, d_out, cap, probe_out, unnest_out);
  if (rc < 0 || !p->comm) return rc;
  if (p->hot_count) HJ_TRY(hot_join(c, p, mode, flags, (uint2*)d_out, cap, probe_out, unnest_out));
  return (mode == 3 ? unnest_out : probe_out)->overflow ? HJ3D_OVERFLOW : HJ3D_OK;
}

static int probe_parts_regular(hj3d_ctx* c, hj3d_table* t, hj3d_parts* p, int mode, uint32_t flags, uint32_t* d_out, uint64_t cap,
                               hj3d_counters* probe_out, hj3d_counters* unnest_out) {
  if (!c || !t || !p || !probe_out) return fail(HJ3D_ERR_INVALID, "NULL argument");
  if (mode < 0 || mode > 3) return fail(HJ3D_ERR_INVALID, "mode must be 0..3");
  if (mode == 3 && !unnest_out) return fail(HJ3D_ERR_INVALID, "unnest_out == NULL");
  if ((mode <= 1) != (t->kind == HJ3D_CHAINING)) return fail(HJ3D_ERR_INVALID, "probe mode does not match the table kind");
  HJ_TRY(parts_fit_table(p, t));
  const hj3d_keyspec ks = parts_ks(p);
  HJ_TRY(table_matches(t, ks));
  CUDA_TRY(cudaSetDevice(c->device));
  HJ_TRY(arena_reset(c));
  begin_call(c);
  CUDA_TRY(cudaMemsetAsync(c->d_ctr, 0, sizeof(DevCounters), c->stream));
  memset(probe_out, 0, sizeof(*probe_out));
  if (unnest_out) memset(unnest_out, 0, sizeof(*unnest_out));
  const PartsView pv{p->d_start, p->d_count, p->n_ranges * p->n_src, p->range_width};
  const Src src = parts_src(p);
  bool fused = false;
  int rc;
#define HJ_BY_HASH(CALL) switch (p->hash_id) { case HJ3D_HASH_MURMUR32: { constexpr int H = HJ3D_HASH_MURMUR32; rc = CALL; } break; \
                                              case HJ3D_HASH_MURMUR64: { constexpr int H = HJ3D_HASH_MURMUR64; rc = CALL; } break; \
                                              default: { constexpr int H = HJ3D_HASH_MURMUR64_SEXT32; rc = CALL; } break; }
  if (mode <= 1)      { HJ_BY_HASH((probe_chaining_impl<H>(c, t, src, mode == 1, flags, (uint2*)d_out, cap, &pv))); }
  else if (mode == 2) { HJ_BY_HASH((probe_nested_impl<H>(c, t, src, flags, (uint2*)d_out, cap, &pv))); }
  else                { HJ_BY_HASH((probe_nested_unnest_impl<H>(c, t, src, flags, (uint2*)d_out, cap, &fused, &pv))); }
#undef HJ_BY_HASH
  end_call(c);
  if (rc < 0) return rc;
  if (mode <= 2) {
    HJ_TRY(fetch_counters(c, probe_out, cap, d_out != nullptr));
    return probe_out->overflow ? HJ3D_OVERFLOW : HJ3D_OK;
  }
  if (fused) {
    bool redo = false;
    HJ_TRY(finish_fused_unnest(c, t, flags, (uint2*)d_out, cap, probe_out, unnest_out, &redo));
    end_call(c);
    if (!redo) return unnest_out->overflow ? HJ3D_OVERFLOW : HJ3D_OK;
    memset(probe_out, 0, sizeof(*probe_out)); memset(unnest_out, 0, sizeof(*unnest_out));
  }
  // not the fine-partition path (small shard): nested tuples into a ctx-owned buffer, then the unnest of the pairs
  auto& hj = c->hj;
  const size_t need = (p->n_total ? p->n_total : 1) * 8;
  if (hj.cnest < need) {
    if (hj.nest) { cudaStreamSynchronize(c->stream); cudaFree(hj.nest); hj.nest = nullptr; hj.cnest = 0; }
    HJ_TRY(raw_alloc(&hj.nest, need));
    hj.cnest = need;
  }
  rc = probe_parts_regular(c, t, p, 2, flags & ~HJ3D_F_CHECKSUM, (uint32_t*)hj.nest, p->n_total, probe_out, nullptr);
  if (rc < 0) return rc;
  return hj3d_unnest_pairs(c, t, (const uint32_t*)hj.nest, probe_out->out_written, flags, d_out, cap, unnest_out);
}

int hj3d_group_first_row(hj3d_ctx* c, hj3d_table* t, const uint32_t* d_gref, uint64_t n, uint32_t* d_out) {
  if (!c || !t) return fail(HJ3D_ERR_INVALID, "NULL argument");
  if (t->kind != HJ3D_NESTED || !t->built) return fail(HJ3D_ERR_INVALID, "needs a built nested table");
  if (!n) return HJ3D_OK;
  CUDA_TRY(cudaSetDevice(c->device));
  if (t->key_bytes == 8) k_group_first_row<uint64_t><<<blocks_for(n, 256), 256, 0, c->stream>>>((const Group<uint64_t>*)t->groups, d_gref, n, d_out);
  else                   k_group_first_row<uint32_t><<<blocks_for(n, 256), 256, 0, c->stream>>>((const Group<uint32_t>*)t->groups, d_gref, n, d_out);
  ++c->launches;
  CUDA_TRY(cudaGetLastError());
  return HJ3D_OK;
}

int hj3d_gather_u32(hj3d_ctx* c, const uint32_t* d_src, const uint32_t* d_idx, uint64_t n, uint32_t* d_dst) {
  if (!c) return fail(HJ3D_ERR_INVALID, "NULL argument");
  if (!n) return HJ3D_OK;
  CUDA_TRY(cudaSetDevice(c->device));
  k_gather_u32<<<blocks_for(n, 256), 256, 0, c->stream>>>(d_src, d_idx, n, d_dst);
  ++c->launches;
  CUDA_TRY(cudaGetLastError());
  return HJ3D_OK;
}

int hj3d_split_pairs(hj3d_ctx* c, const uint32_t* d_pairs, uint64_t n, uint32_t* d_left, uint32_t* d_right) {
  if (!c) return fail(HJ3D_ERR_INVALID, "NULL argument");
  if (!n) return HJ3D_OK;
  CUDA_TRY(cudaSetDevice(c->device));
  k_split_pairs<<<blocks_for(n, 256), 256, 0, c->stream>>>((const uint2*)d_pairs, n, d_left, d_right);
  ++c->launches;
  CUDA_TRY(cudaGetLastError());
  return HJ3D_OK;
}

int hj3d_join_host(hj3d_ctx* c, int mode,
                   const void* h_build, uint64_t nB, hj3d_keyspec ksB, uint64_t D,
                   const void* h_probe, uint64_t nP, hj3d_keyspec ksP, uint32_t flags,
                   uint32_t* h_out, uint64_t cap, hj3d_counters* pc, hj3d_counters* uc, hj3d_stats* st) {
  if (!c || !pc) return fail(HJ3D_ERR_INVALID, "NULL argument");
  if (mode < 0 || mode > 3) return fail(HJ3D_ERR_INVALID, "mode must be 0..3");
  if (mode == 3 && !uc) return fail(HJ3D_ERR_INVALID, "unnest_out == NULL");
  if ((nB && !h_build) || (nP && !h_probe)) return fail(HJ3D_ERR_INVALID, "NULL relation");
  CUDA_TRY(cudaSetDevice(c->device));
  // staging buffers are ctx-owned and grow-only (the sub-calls below reset the temporary arena)
  auto ensure = [&](void** p, size_t* have, size_t bytes) -> int {
    if (bytes == 0) bytes = 256;
    if (*have >= bytes) return HJ3D_OK;
    if (*p) { cudaStreamSynchronize(c->stream); cudaFree(*p); *p = nullptr; *have = 0; }
    HJ_TRY(raw_alloc(p, bytes));
    *have = bytes;
    return HJ3D_OK;
  };
  auto& hj = c->hj;
  const bool trace = getenv("HJ3D_TRACE_HOST") != nullptr;
  auto t0 = std::chrono::steady_clock::now();
  auto stamp = [&](const char* what) {
    if (trace) fprintf(stderr, "[join_host] %-22s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
  };
  // Streamed upload: the build relation goes up first on the copy stream, the probe relation follows in chunks; every probe
  // chunk goes through partition level 1 (a one-rank exchange, hj3d_exchange_begin_host) as soon as it has landed, under the
  // upload of the chunks behind it, and the table is built meanwhile on the caller's stream.  What is left after the last
  // byte is partition level 2 and the probe kernel.
  const uint64_t chunk_rows = c->host_chunk_bytes > 0 ? std::max<uint64_t>(1, (uint64_t)c->host_chunk_bytes / ksP.tuple_bytes) : 0;
  bool streamed = chunk_rows && nP >= 2 * chunk_rows && nP <= 0xFFFFFFF0ull && D <= 0xFFFFFFFFull;
  HJ_TRY(ensure(&hj.b, &hj.cb, nB * ksB.tuple_bytes));
  if (streamed) {
    if (!c->copy_stream) CUDA_TRY(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    if (!c->host_comm) HJ_TRY(hj3d_comm_create(c, 1, 0, nullptr, &c->host_comm));
    if (c->chunk_ev.empty()) { cudaEvent_t e; CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); c->chunk_ev.push_back(e); }
    const uint64_t want = nP + nP / 24 + (4ull << 20);          // 256 regions of n/256 + 4 % + 16 K records
    if (c->host_comm_records < want || c->host_comm_key_bytes != ksP.key_bytes) {
      HJ_TRY(hj3d_comm_reserve(c->host_comm, 1, want, ksP.key_bytes));
      c->host_comm_records = want; c->host_comm_key_bytes = ksP.key_bytes;
    }
    // the copy stream starts where the caller's stream is now (hj.b may still be read by an earlier call's kernels)
    CUDA_TRY(cudaEventRecord(c->chunk_ev[0], c->stream));
    CUDA_TRY(cudaStreamWaitEvent(c->copy_stream, c->chunk_ev[0], 0));
    if (nB) CUDA_TRY(cudaMemcpyAsync(hj.b, h_build, nB * ksB.tuple_bytes, cudaMemcpyHostToDevice, c->copy_stream));
    CUDA_TRY(cudaEventRecord(c->chunk_ev[0], c->copy_stream));
    HJ_TRY(hj3d_exchange_begin_host(c->host_comm, 1, h_probe, nP, ksP, D, 0, 0, nullptr));
    CUDA_TRY(cudaStreamWaitEvent(c->stream, c->chunk_ev[0], 0));
    stamp("copies enqueued");
  } else {
    HJ_TRY(ensure(&hj.p, &hj.cp, nP * ksP.tuple_bytes));
    if (nB) CUDA_TRY(cudaMemcpyAsync(hj.b, h_build, nB * ksB.tuple_bytes, cudaMemcpyHostToDevice, c->stream));
    if (nP) CUDA_TRY(cudaMemcpyAsync(hj.p, h_probe, nP * ksP.tuple_bytes, cudaMemcpyHostToDevice, c->stream));
  }
  hj3d_table* t = nullptr;
  HJ_TRY(hj3d_table_create(c, mode <= 1 ? HJ3D_CHAINING : HJ3D_NESTED, D, &t));
  auto bail = [&](int code) {
    if (streamed) cudaDeviceSynchronize();                    // nothing of this call may still be in flight on the side streams
    hj3d_table_destroy(c, t);
    return code;
  };
  int rc = hj3d_table_build(c, t, hj.b, nB, ksB);
  if (rc < 0) return bail(rc);
  stamp("table built");
  const bool want_pairs = h_out != nullptr || (flags & HJ3D_F_DEVICE_RESULT);
  uint64_t n_out = 0;
  uint32_t* dOut = nullptr;
  if (want_pairs) { rc = ensure(&hj.out, &hj.cout, cap * 8); if (rc < 0) return bail(rc); dOut = (uint32_t*)hj.out; }
  hj3d_parts* parts = nullptr;
  if (streamed) {
    rc = hj3d_exchange_end(c->host_comm, 1, nullptr, 0, nP, &parts);
    stamp("exchange_end");
    if (rc < 0) { hj3d_parts_destroy(parts); return bail(rc); }
    if (rc == HJ3D_OVERFLOW) {     // skewed keys overflowed a fixed region: upload again in one piece and take the general path
      hj3d_parts_destroy(parts); parts = nullptr; streamed = false;
      rc = ensure(&hj.p, &hj.cp, nP * ksP.tuple_bytes); if (rc < 0) return bail(rc);
      cudaError_t e = cudaMemcpyAsync(hj.p, h_probe, nP * ksP.tuple_bytes, cudaMemcpyHostToDevice, c->stream);
      if (e != cudaSuccess) return bail(fail(HJ3D_ERR_CUDA, cudaGetErrorString(e)));
    }
  }
  if (parts) {
    rc = hj3d_probe_parts(c, t, parts, mode, flags, dOut, cap, pc, uc);
    hj3d_parts_destroy(parts);
    if (rc < 0) return bail(rc);
    n_out = mode == 3 ? uc->out_written : pc->out_written;
  } else if (mode <= 1) {
    rc = hj3d_probe_chaining(c, t, hj.p, nP, ksP, nullptr, mode == 1, flags, dOut, cap, pc);
    if (rc < 0) return bail(rc);
    n_out = pc->out_written;
  } else if (mode == 2) {
    rc = hj3d_probe_nested(c, t, hj.p, nP, ksP, nullptr, flags, dOut, cap, pc);
    if (rc < 0) return bail(rc);
    n_out = pc->out_written;
  } else {
    rc = hj3d_probe_nested_unnest(c, t, hj.p, nP, ksP, flags, dOut, cap, pc, uc);
    if (rc < 0) return bail(rc);
    n_out = uc->out_written;
  }
  stamp("probe done");
  const int final_rc = rc;
  if (h_out && n_out) {
    cudaError_t e = cudaMemcpyAsync(h_out, dOut, n_out * 8, cudaMemcpyDeviceToHost, c->stream);
    if (e != cudaSuccess) { hj3d_table_destroy(c, t); return fail(HJ3D_ERR_CUDA, cudaGetErrorString(e)); }
  }
  if (st) { rc = hj3d_table_stats(c, t, st); if (rc < 0) return bail(rc); }
  cudaError_t e = cudaStreamSynchronize(c->stream);
  hj3d_table_destroy(c, t);
  if (e != cudaSuccess) return fail(HJ3D_ERR_CUDA, cudaGetErrorString(e));
  return final_rc;
}

}  // extern "C"


template <int HASH>
static int partition_by_owner_t(hj3d_ctx* c, Src src, Dir d, uint32_t width, uint32_t n_owners, uint32_t rowid_base,
                                void* d_out, uint64_t* h_counts) {
  using KeyT = typename HashT<HASH>::key_t;
  PhaseTimer pt(c, PH_PARTITION);
  const PartFn pf = make_partfn(width, 0, d.D);
  unsigned long long *counts = nullptr, *starts = nullptr, *cursor = nullptr;
  HJ_TRY(dev_alloc(c, &counts, n_owners)); HJ_TRY(dev_alloc(c, &starts, n_owners)); HJ_TRY(dev_alloc(c, &cursor, n_owners));
  CUDA_TRY(cudaMemsetAsync(counts, 0, (size_t)n_owners * 8, c->stream));
  CUDA_TRY(cudaMemsetAsync(cursor, 0, (size_t)n_owners * 8, c->stream));
  const uint32_t nb = blocks_for(src.n, kPartTile);
  const int kTile = (int)c->part_threads * PartCfg<KeyT>::kItems;
  const uint32_t nb2 = blocks_for(src.n, kTile);
  if (nb) k_part_hist<HASH><<<nb, kPartThreads, 0, c->stream>>>(src, d, pf, n_owners, counts);
  k_part_prefix<<<1, 32, 0, c->stream>>>(counts, n_owners, starts);
  CUDA_TRY((launch_part_scatter<HASH, false>(c->stream, false, (int)c->part_threads, c->part_rank_match != 0, nb2, src, nullptr, d, pf,
                                             n_owners, n_owners, rowid_base, ~0ull, starts, cursor, (Slot<KeyT>*)d_out)));
  c->launches += nb ? 3 : 1;
  unsigned long long* h = (unsigned long long*)c->h_pinned;
  CUDA_TRY(cudaMemcpyAsync(h, counts, (size_t)n_owners * 8, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  for (uint32_t i = 0; i < n_owners; ++i) h_counts[i] = h[i];
  CUDA_TRY(cudaGetLastError());
  return HJ3D_OK;
}

extern "C" {

int hj3d_partition_by_owner(hj3d_ctx* c, const void* d_tuples, uint64_t n, hj3d_keyspec ks,
                            uint64_t D, uint32_t n_owners, uint32_t rowid_base, void* d_out, uint64_t* h_counts) {
  if (!c || !h_counts) return fail(HJ3D_ERR_INVALID, "NULL argument");
  if (!D || D > 0xFFFFFFFFull || !n_owners || n_owners > (uint32_t)kMaxParts) return fail(HJ3D_ERR_INVALID, "bad num_buckets / n_owners");
  if (n && (!d_tuples || !d_out)) return fail(HJ3D_ERR_INVALID, "NULL buffer");
  HJ_TRY(check_keyspec(ks));
  CUDA_TRY(cudaSetDevice(c->device));
  HJ_TRY(arena_reset(c));
  begin_call(c);
  Src src = make_src(d_tuples, n, ks, nullptr);
  Dir d = make_dir(D, 0, D);
  const uint32_t width = (uint32_t)((D + n_owners - 1) / n_owners);
  int rc;
  switch (ks.hash_id) {
    case HJ3D_HASH_MURMUR32: rc = partition_by_owner_t<HJ3D_HASH_MURMUR32>(c, src, d, width, n_owners, rowid_base, d_out, h_counts); break;
    case HJ3D_HASH_MURMUR64: rc = partition_by_owner_t<HJ3D_HASH_MURMUR64>(c, src, d, width, n_owners, rowid_base, d_out, h_counts); break;
    default:                 rc = partition_by_owner_t<HJ3D_HASH_MURMUR64_SEXT32>(c, src, d, width, n_owners, rowid_base, d_out, h_counts); break;
  }
  end_call(c);
  return rc;
}

}  // extern "C"
