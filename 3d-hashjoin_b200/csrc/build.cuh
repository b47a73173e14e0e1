// build.cuh -- build-side kernels.
//
// Chaining table (HtChaining1, ht_chaining.hh): insert() appends every tuple to the chain of bucket
// h % D.  On the device the chains become contiguous runs:
//     off[D+1]   start offset of every bucket's run (off[b+1]-off[b] = chain length incl. dir entry)
//     slots[n]   (key, row id) of every build tuple, grouped by bucket
// built by histogram -> prefix sum -> scatter.  Order inside a run is unspecified; everything the
// reference derives from insertion order (numCmps, IsBuildKeyUnique early exit) is recomputed from
// row ids at probe time (row id order == insertion order).
//
// Nested table (HtNested1, ht_nested.hh): one MainNode per distinct key, duplicates hang below it.
// On the device: the slots of every bucket are additionally grouped by key
//     goff[D+1]  first group of every bucket (goff[b+1]-goff[b] = main chain length = #distinct keys)
//     groups[G]  {key, first_row (the MainNode's own tuple = min row id), start, len}
//     rows[n]    build row ids, grouped by group
// Grouping uses the bucket's own slot range as a tiny open-addressing set (capacity = chain length
// >= #distinct keys), so it needs no sort and no per-bucket serial work; hot keys only cost
// warp-aggregated atomics.
#pragma once

#include "common.cuh"
#include "scan.cuh"

namespace hj3d {

constexpr int kBuildThreads = 256;
constexpr int kBuildItems   = 8;
constexpr int kBuildTile    = kBuildThreads * kBuildItems;

// ---- 1. histogram: off[b] += 1 for every build tuple -------------------------------------------
template <int HASH, bool AGG>
__global__ void __launch_bounds__(kBuildThreads) k_histogram(Src s, Dir d, const uint2* __restrict__ tilemap, uint32_t* __restrict__ off) {
  using KeyT = typename HashT<HASH>::key_t;
  uint64_t t0; uint32_t tn;
  block_tile<kBuildTile>(tilemap, s.n, t0, tn);
#pragma unroll
  for (int j = 0; j < kBuildItems; ++j) {
    const uint32_t li = j * kBuildThreads + threadIdx.x;
    const uint64_t i = t0 + li;
    const bool in = li < tn;
    uint32_t b = 0;
    bool ok = in;
    if (in) { b = HashT<HASH>::bucket(src_key<KeyT>(s, i), d) - d.lo; ok = b < d.n_local; }   // shard tables skip foreign buckets
    if (AGG) {
      // warp-aggregated: one atomic per distinct bucket in the warp (hot keys of skewed inputs)
      const uint32_t act = __ballot_sync(0xffffffffu, ok);
      if (ok) {
        const uint32_t peers = __match_any_sync(act, b);
        if ((uint32_t)(__ffs(peers) - 1) == lane_id()) atomicAdd(off + b, (uint32_t)__popc(peers));
      }
    } else {
      if (ok) atomicAdd(off + b, 1u);
    }
  }
}

// ---- 2. (scan.cuh) inclusive prefix sum in place: off[b] = end of bucket b -----------------------

// ---- 3. scatter: claim a position from the end of the bucket's run -------------------------------
// After the pass off[b] has been decremented chain-length times, i.e. holds the START of the run.
template <int HASH, bool AGG>
__global__ void __launch_bounds__(kBuildThreads)
k_scatter(Src s, Dir d, const uint2* __restrict__ tilemap, uint32_t* __restrict__ off,
          Slot<typename HashT<HASH>::key_t>* __restrict__ slots) {
  using KeyT = typename HashT<HASH>::key_t;
  uint64_t t0; uint32_t tn;
  block_tile<kBuildTile>(tilemap, s.n, t0, tn);
#pragma unroll
  for (int j = 0; j < kBuildItems; ++j) {
    const uint32_t li = j * kBuildThreads + threadIdx.x;
    const uint64_t i = t0 + li;
    bool ok = li < tn;
    KeyT key = 0; uint32_t b = 0;
    if (ok) { key = src_key<KeyT>(s, i); b = HashT<HASH>::bucket(key, d) - d.lo; ok = b < d.n_local; }
    uint32_t pos = 0;
    if (AGG) {
      const uint32_t act = __ballot_sync(0xffffffffu, ok);
      if (ok) {
        const uint32_t peers = __match_any_sync(act, b);
        const uint32_t leader = __ffs(peers) - 1, cnt = __popc(peers);
        uint32_t basepos = 0;
        if (leader == lane_id()) basepos = atomicSub(off + b, cnt) - cnt;
        basepos = __shfl_sync(peers, basepos, leader);
        pos = basepos + __popc(peers & ((1u << lane_id()) - 1));
      }
    } else {
      if (ok) pos = atomicSub(off + b, 1u) - 1;
    }
    if (ok) {
      Slot<KeyT> sl; sl.key = key; sl.rowid = src_rowid(s, i);
      slots[pos] = sl;
    }
  }
}

// ---- 4. nested grouping ---------------------------------------------------------------------------
// claim: every slot finds (or becomes) the representative cell of its key inside its bucket's range.
//   cell[c]  = slot index of the first claimer of cell c (kEmpty32 = free)
//   rep[i]   = cell of slot i's key;   gcnt[c] += 1;   gmin[c] = min row id
template <int HASH>
__global__ void __launch_bounds__(kBuildThreads)
k_group_claim(const Slot<typename HashT<HASH>::key_t>* __restrict__ slots, uint64_t n, Dir d,
              const uint32_t* __restrict__ off, uint32_t* cell, uint32_t* __restrict__ rep,
              uint32_t* gcnt, uint32_t* gmin) {
  using KeyT = typename HashT<HASH>::key_t;
  const uint64_t base = (uint64_t)blockIdx.x * kBuildTile + threadIdx.x;
#pragma unroll 2
  for (int j = 0; j < kBuildItems; ++j) {
    const uint64_t i = base + (uint64_t)j * kBuildThreads;
    if (i >= n) continue;
    const Slot<KeyT> me = slots[i];
    const uint32_t b = HashT<HASH>::bucket(me.key, d) - d.lo;
    const uint32_t lo = off[b], len = off[b + 1] - lo;
    uint32_t c;
    if (len == 1) {           // the common case of a key/foreign-key build side: no contention possible
      c = lo;
      cell[c] = (uint32_t)i;
      gcnt[c] = 1; gmin[c] = me.rowid; rep[i] = c;
      continue;
    }
    uint32_t p = __umulhi(mix2(me.key), len);
    for (;;) {
      c = lo + p;
      uint32_t v = cell[c];
      if (v == kEmpty32) {
        v = atomicCAS(cell + c, kEmpty32, (uint32_t)i);
        if (v == kEmpty32) v = (uint32_t)i;
      }
      if (v == (uint32_t)i || slots[v].key == me.key) break;
      if (++p == len) p = 0;
    }
    rep[i] = c;
    // Skewed keys: millions of rows of one key update the same two words.  Slots are bucket ordered, so a hot key fills whole
    // warps: lanes that share a cell combine their update (one atomic pair per warp instead of 32 serialised ones).
    const uint32_t peers = __match_any_sync(__activemask(), c);
    if (__popc(peers) == 1) {
      atomicAdd(gcnt + c, 1u);
      atomicMin(gmin + c, me.rowid);
    } else {
      const uint32_t mn = __reduce_min_sync(peers, me.rowid);
      if ((uint32_t)__ffs(peers) - 1u == lane_id()) { atomicAdd(gcnt + c, (uint32_t)__popc(peers)); atomicMin(gmin + c, mn); }
    }
  }
}

// (scan.cuh) exclusive scan over cells of the packed pair (claimed?1:0 , gcnt) gives for every cell
//   gidx[c]   = dense group index,   gstart[c] = first row position of the group.

// emit: one Group record per claimed cell; gstart doubles as the scatter cursor afterwards.
template <int HASH>
__global__ void __launch_bounds__(kBuildThreads)
k_group_emit(const Slot<typename HashT<HASH>::key_t>* __restrict__ slots, uint64_t n,
             const uint32_t* __restrict__ cell, const uint32_t* __restrict__ gcnt, const uint32_t* __restrict__ gmin,
             const uint32_t* __restrict__ gidx, const uint32_t* __restrict__ gstart,
             Group<typename HashT<HASH>::key_t>* __restrict__ groups) {
  using KeyT = typename HashT<HASH>::key_t;
  const uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const uint32_t v = cell[c];
  if (v == kEmpty32) return;
  Group<KeyT> g;
  g.key = slots[v].key; g.first_row = gmin[c]; g.start = gstart[c]; g.len = gcnt[c];
  groups[gidx[c]] = g;
}

// goff[b] = gidx at the first cell of bucket b (gidx has n+1 entries, the last one = #groups)
__global__ void k_group_offsets(const uint32_t* __restrict__ off, const uint32_t* __restrict__ gidx,
                                uint32_t n_buckets_plus1, uint32_t* __restrict__ goff) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < n_buckets_plus1) goff[b] = gidx[off[b]];
}

// rows: scatter every slot's row id into its group's run
template <int HASH>
__global__ void __launch_bounds__(kBuildThreads)
k_group_rows(const Slot<typename HashT<HASH>::key_t>* __restrict__ slots, uint64_t n,
             const uint32_t* __restrict__ rep, uint32_t* gcursor, uint32_t* __restrict__ rows) {
  const uint64_t base = (uint64_t)blockIdx.x * kBuildTile + threadIdx.x;
#pragma unroll
  for (int j = 0; j < kBuildItems; ++j) {
    const uint64_t i = base + (uint64_t)j * kBuildThreads;
    const bool ok = i < n;
    uint32_t c = 0;
    if (ok) c = rep[i];
    const uint32_t act = __ballot_sync(0xffffffffu, ok);
    if (ok) {
      const uint32_t peers = __match_any_sync(act, c);
      const uint32_t leader = __ffs(peers) - 1, cnt = __popc(peers);
      uint32_t basepos = 0;
      if (leader == lane_id()) basepos = atomicAdd(gcursor + c, cnt);
      basepos = __shfl_sync(peers, basepos, leader);
      rows[basepos + __popc(peers & ((1u << lane_id()) - 1))] = slots[i].rowid;
    }
  }
}

// ---- physical order of short buckets ---------------------------------------------------------------
// Buckets of 2..kOrderedMaxB entries are rewritten in the reference's own order so that the probe is a
// plain early-exit walk (probe.cuh):
//   chaining: chain order  [oldest, newest, .., second oldest]  (ht_chaining.hh:185-194)
//   nested:   main chain in first-appearance order = ascending first_row (ht_nested.hh:303-308)
// One thread per bucket; consecutive threads touch consecutive buckets = consecutive memory.
constexpr uint32_t kOrderedMaxB = 16;   // == kOrderedMax of probe.cuh


// In-place reorder of one short bucket (2 <= n <= kOrderedMaxB) at base[0..n).
// CHAIN = true : chain order of HtChaining1 = [oldest, newest, .., second oldest]   (key = rowid)
// CHAIN = false: ascending by `first_row` (first-appearance order of HtNested1's main chain)
template <class T> __device__ __forceinline__ uint32_t order_key(const T& v);
template <> __device__ __forceinline__ uint32_t order_key(const Slot<uint32_t>& v) { return v.rowid; }
template <> __device__ __forceinline__ uint32_t order_key(const Slot<uint64_t>& v) { return v.rowid; }
template <> __device__ __forceinline__ uint32_t order_key(const Group<uint32_t>& v) { return v.first_row; }
template <> __device__ __forceinline__ uint32_t order_key(const Group<uint64_t>& v) { return v.first_row; }

template <class T, bool CHAIN>
__device__ __forceinline__ void order_short_bucket(T* base, uint32_t n) {
  if (n <= 4) {   // 93% of the multi-entry buckets of a load-factor-1 table: 4-element sorting network
    T v0 = base[0], v1 = base[1], v2 = n > 2 ? base[2] : v1, v3 = n > 3 ? base[3] : v1;
    uint32_t k0 = order_key(v0), k1 = order_key(v1), k2 = n > 2 ? order_key(v2) : 0xFFFFFFFFu, k3 = n > 3 ? order_key(v3) : 0xFFFFFFFFu;
#define HJ_CSWAP(a, b, ka, kb) do { if (kb < ka) { T t_ = a; a = b; b = t_; uint32_t u_ = ka; ka = kb; kb = u_; } } while (0)
    HJ_CSWAP(v0, v1, k0, k1); HJ_CSWAP(v2, v3, k2, k3); HJ_CSWAP(v0, v2, k0, k2); HJ_CSWAP(v1, v3, k1, k3); HJ_CSWAP(v1, v2, k1, k2);
#undef HJ_CSWAP
    if (CHAIN) {      // ascending a<b<c<d  ->  [a, d, c, b]
      base[0] = v0;
      if (n == 2) base[1] = v1;
      else if (n == 3) { base[1] = v2; base[2] = v1; }
      else { base[1] = v3; base[2] = v2; base[3] = v1; }
    } else {
      base[0] = v0; base[1] = v1;
      if (n > 2) base[2] = v2;
      if (n > 3) base[3] = v3;
    }
    return;
  }
  T v[kOrderedMaxB];
#pragma unroll
  for (uint32_t k = 0; k < kOrderedMaxB; ++k) if (k < n) v[k] = base[k];
#pragma unroll
  for (uint32_t k = 0; k < kOrderedMaxB; ++k) {
    if (k >= n) break;
    uint32_t older = 0;                                   // rank by row id = insertion order
#pragma unroll
    for (uint32_t m = 0; m < kOrderedMaxB; ++m) if (m < n) older += order_key(v[m]) < order_key(v[k]);
    base[CHAIN ? (older == 0 ? 0 : n - older) : older] = v[k];
  }
}

template <class KeyT>
__global__ void __launch_bounds__(256)
k_order_slots(const uint32_t* __restrict__ off, Slot<KeyT>* __restrict__ slots, uint32_t n_buckets) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_buckets) return;
  const uint32_t lo = off[b], n = off[b + 1] - lo;
  if (n < 2 || n > kOrderedMaxB) return;
  order_short_bucket<Slot<KeyT>, true>(slots + lo, n);
}

template <class KeyT>
__global__ void __launch_bounds__(256)
k_order_groups(const uint32_t* __restrict__ goff, Group<KeyT>* __restrict__ groups, uint32_t n_buckets) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_buckets) return;
  const uint32_t lo = goff[b], n = goff[b + 1] - lo;
  if (n < 2 || n > kOrderedMaxB) return;
  order_short_bucket<Group<KeyT>, false>(groups + lo, n);
}

// ---- fine-partition build (chaining): one block builds one bucket range completely in shared memory ----
// Input: the build records of fine partition f (bucket range [f*width, (f+1)*width)), contiguous at
// recs[part_start[f] .. +counts[f]).  The block histograms them over its `width` buckets (one shared
// atomic per record, which also yields the record's rank inside its bucket), scans the histogram,
// places the records at off[b] + rank, rewrites short buckets in chain order (k_order_slots) and
// streams the finished directory words and slots to global memory with coalesced stores.  The bucket
// statistics of makeStatistics are reduced on the way.  base[f] = number of build records in
// partitions < f.  A partition with more than `cap_recs` records sets *overflow and is skipped (the
// caller then rebuilds with the global-memory kernels).
constexpr int kFineBuildThreads = 256;   // latency bound: resident warps matter more than registers per thread
constexpr int kFineBuildItems   = 12;

template <int HASH>
#ifndef HJ3D_BUILD_MINBLOCKS
#define HJ3D_BUILD_MINBLOCKS 4   // measured at 2^27 rows: 2 blocks (121 regs) 2.34 ms, 3 blocks (80) 1.84, 4 blocks (64, small spills) 1.72
#endif
__global__ void __launch_bounds__(kFineBuildThreads, HJ3D_BUILD_MINBLOCKS)
k_build_fine(const Slot<typename HashT<HASH>::key_t>* __restrict__ recs,
             const unsigned long long* __restrict__ part_start, const unsigned long long* __restrict__ counts,
             const unsigned long long* __restrict__ base, Dir d, uint32_t width, uint32_t n_fine, uint32_t cap_recs,
             uint32_t* __restrict__ off, Slot<typename HashT<HASH>::key_t>* __restrict__ slots,
             DevStats* stats, uint32_t* overflow) {
  using KeyT = typename HashT<HASH>::key_t;
  using SlotT = Slot<KeyT>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* sm_off = reinterpret_cast<uint32_t*>(smem_raw);                 // [width + 1] counts -> exclusive offsets
  SlotT*    sm_slots = reinterpret_cast<SlotT*>(smem_raw + (((width + 1) * 4 + 15) & ~15u));
  __shared__ uint32_t sm_scan[33];
  __shared__ unsigned long long sm_red[160];

  const uint32_t f = blockIdx.x;
  const uint32_t blo = f * width;
  const uint32_t bhi = (blo + width < d.n_local) ? blo + width : d.n_local;
  const uint32_t nbk = bhi - blo;
  const unsigned long long cnt64 = f < n_fine ? counts[f] : 0ull;
  if (cnt64 > cap_recs) { if (threadIdx.x == 0) atomicExch(overflow, 1u); return; }
  const uint32_t cnt = (uint32_t)cnt64;
  const uint32_t gbase = (uint32_t)base[f];
  const SlotT* in = recs + (f < n_fine ? part_start[f] : 0ull);
  for (uint32_t b = threadIdx.x; b <= nbk; b += kFineBuildThreads) sm_off[b] = 0;
  // ---- load this partition's records (registers) and histogram them
  KeyT     key[kFineBuildItems];
  uint32_t rid[kFineBuildItems], br[kFineBuildItems];      // br = (local bucket << 14) | rank ... rank < 2^14, bucket < 2^18
#pragma unroll
  for (int j = 0; j < kFineBuildItems; ++j) {
    const uint32_t li = j * kFineBuildThreads + threadIdx.x;
    key[j] = 0; rid[j] = 0;
    if (li < cnt) { const SlotT r = in[li]; key[j] = r.key; rid[j] = r.rowid; }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kFineBuildItems; ++j) {
    const uint32_t li = j * kFineBuildThreads + threadIdx.x;
    br[j] = 0xFFFFFFFFu;
    if (li < cnt) {
      const uint32_t b = HashT<HASH>::bucket(key[j], d) - d.lo - blo;
      br[j] = (b << 14) | atomicAdd(&sm_off[b], 1u);
    }
  }
  __syncthreads();
  // ---- statistics over the bucket lengths + exclusive scan (in place)
  {
    DevAgg all{~0ull, 0, 0, 0, 0}, ne{~0ull, 0, 0, 0, 0};
    unsigned long long empty = 0;
    constexpr uint32_t PER = 8;                              // width <= 2048 = 256 threads x 8
    const uint32_t a = threadIdx.x * PER;
    uint32_t v[PER], sum = 0;
#pragma unroll
    for (uint32_t k = 0; k < PER; ++k) {
      v[k] = 0;
      if (a + k < nbk) {
        v[k] = sm_off[a + k];
        stats_step(all, v[k]);
        if (v[k]) stats_step(ne, v[k]); else ++empty;
      }
      sum += v[k];
    }
    uint32_t tot;
    uint32_t ex = block_exscan(sum, sm_scan, &tot);
#pragma unroll
    for (uint32_t k = 0; k < PER; ++k) { if (a + k < nbk) sm_off[a + k] = ex; ex += v[k]; }
    if (threadIdx.x == 0) sm_off[nbk] = cnt;
    DevStats* my_stats = stats + (blockIdx.x & 63u);                   // replica (engine.cu: kStatsCopies = 64)
    agg_commit(all, &my_stats->all, sm_red);
    agg_commit(ne, &my_stats->nonempty, sm_red);
    empty = warp_sum(empty);
    if (lane_id() == 0 && empty) atomicAdd(&my_stats->empty, empty);
  }
  __syncthreads();
  // ---- place the records
#pragma unroll
  for (int j = 0; j < kFineBuildItems; ++j) {
    if (br[j] == 0xFFFFFFFFu) continue;
    SlotT r; r.key = key[j]; r.rowid = rid[j];
    sm_slots[sm_off[br[j] >> 14] + (br[j] & 0x3FFFu)] = r;
  }
  __syncthreads();
  // ---- short buckets into chain order [oldest, newest, .., second oldest] (ht_chaining.hh:185-194)
  for (uint32_t b = threadIdx.x; b < nbk; b += kFineBuildThreads) {
    const uint32_t lo = sm_off[b], n = sm_off[b + 1] - lo;
    if (n < 2 || n > kOrderedMaxB) continue;
    order_short_bucket<SlotT, true>(sm_slots + lo, n);
  }
  __syncthreads();
  // ---- stream out: directory words (global offsets) and slots
  for (uint32_t b = threadIdx.x; b < nbk; b += kFineBuildThreads) off[blo + b] = gbase + sm_off[b];
  if (bhi == d.n_local && threadIdx.x == 0) off[d.n_local] = gbase + cnt;
  for (uint32_t i = threadIdx.x; i < cnt; i += kFineBuildThreads) slots[gbase + i] = sm_slots[i];
}

// ---- statistics helpers ---------------------------------------------------------------------------
// chaining _numDistinctKeys = |{ (int)hashvalue }| (ht_chaining.hh:267,282): distinct low 32 hash bits.
template <int HASH>
__global__ void k_hash_bitmap(const Slot<typename HashT<HASH>::key_t>* __restrict__ slots, uint64_t n, uint32_t* bitmap) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t h = HashT<HASH>::hash_lo32(slots[i].key);
  atomicOr(bitmap + (h >> 5), 1u << (h & 31));
}
__global__ void k_popcount(const uint32_t* __restrict__ bitmap, uint64_t nwords, unsigned long long* out) {
  unsigned long long c = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (uint64_t)gridDim.x * blockDim.x)
    c += __popc(bitmap[i]);
  c = warp_sum(c);
  if (lane_id() == 0 && c) atomicAdd(out, c);
}

}  // namespace hj3d
