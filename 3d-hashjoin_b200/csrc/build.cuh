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
    atomicAdd(gcnt + c, 1u);
    atomicMin(gmin + c, me.rowid);
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

template <class KeyT>
__global__ void __launch_bounds__(256)
k_order_slots(const uint32_t* __restrict__ off, Slot<KeyT>* __restrict__ slots, uint32_t n_buckets) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_buckets) return;
  const uint32_t lo = off[b], n = off[b + 1] - lo;
  if (n < 2 || n > kOrderedMaxB) return;
  Slot<KeyT> v[kOrderedMaxB];
#pragma unroll
  for (uint32_t k = 0; k < kOrderedMaxB; ++k) if (k < n) v[k] = slots[lo + k];
#pragma unroll
  for (uint32_t k = 0; k < kOrderedMaxB; ++k) {
    if (k >= n) break;
    uint32_t older = 0;                                   // rank by row id = insertion order
#pragma unroll
    for (uint32_t m = 0; m < kOrderedMaxB; ++m) if (m < n) older += v[m].rowid < v[k].rowid;
    const uint32_t pos = older == 0 ? 0 : n - older;      // chain position of the tuple with rank `older`
    slots[lo + pos] = v[k];
  }
}

template <class KeyT>
__global__ void __launch_bounds__(256)
k_order_groups(const uint32_t* __restrict__ goff, Group<KeyT>* __restrict__ groups, uint32_t n_buckets) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_buckets) return;
  const uint32_t lo = goff[b], n = goff[b + 1] - lo;
  if (n < 2 || n > kOrderedMaxB) return;
  Group<KeyT> v[kOrderedMaxB];
#pragma unroll
  for (uint32_t k = 0; k < kOrderedMaxB; ++k) if (k < n) v[k] = groups[lo + k];
#pragma unroll
  for (uint32_t k = 0; k < kOrderedMaxB; ++k) {
    if (k >= n) break;
    uint32_t older = 0;
#pragma unroll
    for (uint32_t m = 0; m < kOrderedMaxB; ++m) if (m < n) older += v[m].first_row < v[k].first_row;
    groups[lo + older] = v[k];
  }
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
