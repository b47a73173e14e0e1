// probe.cuh -- probe-side kernels: chaining probe, nested probe, deferred unnest.
//
// Every probe thread recomputes the counters the reference accumulates tuple-at-a-time:
//   matches  = AlgBase::_count of the probe operator,  num_cmps = _numCmps (algebra.hh:449,658)
// from the bucket contents alone:
//   chaining, IsBuildKeyUnique=false : every probe into a non-empty bucket compares the whole chain (n)
//   chaining, IsBuildKeyUnique=true  : chain order is [t0, t_{n-1}, .., t_1] (new nodes are linked right
//                                      after the directory entry, ht_chaining.hh:189-194), so the walk
//                                      stops at position 1 if the oldest tuple matches, else at
//                                      n - rank + 1 of the newest matching tuple; n without a match
//   nested                           : main chain is in first-appearance order (tail append,
//                                      ht_nested.hh:303-308): index(key)+1 on a hit, #distinct keys on a miss
// "oldest/newest" are decided by row id, which is the insertion order of the build strand.
//
// Results are written with block-aggregated allocation: one atomicAdd on the output cursor per
// block tile, pairs of a tile land contiguously (coalesced 8-byte stores).
#pragma once

#include "common.cuh"

namespace hj3d {

constexpr int kProbeThreads = 256;
constexpr int kProbeItems   = 4;
constexpr int kProbeTile    = kProbeThreads * kProbeItems;

struct ProbeAcc {
  unsigned long long matches = 0, cmps = 0, sum = 0, x = 0;
};

// block-reduce the per-thread accumulators and commit them with one set of atomics per block
__device__ __forceinline__ void commit_acc(const ProbeAcc& a, DevCounters* c, bool checksum) {
  __shared__ unsigned long long red[4][32];
  unsigned long long m = warp_sum(a.matches), q = warp_sum(a.cmps), s = warp_sum(a.sum), x = warp_xor(a.x);
  const uint32_t w = threadIdx.x >> 5, l = lane_id(), nw = blockDim.x >> 5;
  __syncthreads();
  if (l == 0) { red[0][w] = m; red[1][w] = q; red[2][w] = s; red[3][w] = x; }
  __syncthreads();
  if (w == 0) {
    m = l < nw ? red[0][l] : 0; q = l < nw ? red[1][l] : 0; s = l < nw ? red[2][l] : 0; x = l < nw ? red[3][l] : 0;
    m = warp_sum(m); q = warp_sum(q); s = warp_sum(s); x = warp_xor(x);
    if (l == 0) {
      if (m) atomicAdd(&c->matches, m);
      if (q) atomicAdd(&c->num_cmps, q);
      if (checksum) { atomicAdd(&c->checksum_sum, s); atomicXor(&c->checksum_xor, x); }
    }
  }
}

// ---- chaining probe -------------------------------------------------------------------------------
template <int HASH, bool UNIQUE, bool CHECKSUM, bool WRITE>
__global__ void __launch_bounds__(kProbeThreads)
k_probe_chaining(Src s, Dir d, const uint2* __restrict__ tilemap, const uint32_t* __restrict__ off,
                 const Slot<typename HashT<HASH>::key_t>* __restrict__ slots,
                 uint2* __restrict__ out, unsigned long long out_cap, DevCounters* ctr) {
  using KeyT = typename HashT<HASH>::key_t;
  __shared__ unsigned long long sm_scan[33];
  __shared__ unsigned long long sm_base;
  uint64_t t0; uint32_t tn;
  block_tile<kProbeTile>(tilemap, s.n, t0, tn);
  ProbeAcc acc;
  KeyT     key[kProbeItems];
  uint32_t lo[kProbeItems], len[kProbeItems], nm[kProbeItems], first[kProbeItems];
  unsigned long long mine = 0;
#pragma unroll
  for (int j = 0; j < kProbeItems; ++j) {
    const uint32_t li = j * kProbeThreads + threadIdx.x;
    const uint64_t i = t0 + li;
    lo[j] = len[j] = nm[j] = first[j] = 0; key[j] = 0;
    if (li < tn) {
      key[j] = src_key<KeyT>(s, i);
      const uint32_t b = HashT<HASH>::bucket(key[j], d);
      if (b - d.lo < d.n_local) {                      // shard tables only own [lo, lo + n_local)
        lo[j] = off[b - d.lo];
        len[j] = off[b - d.lo + 1] - lo[j];
      }
    }
  }
#pragma unroll
  for (int j = 0; j < kProbeItems; ++j) {
    const uint32_t n = len[j];
    if (n == 0) continue;                               // empty bucket: no comparison (algebra.hh:640-643)
    if (!UNIQUE) {
      uint32_t m = 0, f = 0;
      for (uint32_t k = 0; k < n; ++k) {
        const Slot<KeyT> sl = slots[lo[j] + k];
        if (sl.key == key[j]) { if (m == 0) f = sl.rowid; ++m; }
      }
      nm[j] = m; first[j] = f;
      acc.cmps += n;                                    // whole chain is walked (algebra.hh:644-657)
    } else {
      // first match in chain order [oldest, newest, .., second oldest]
      uint32_t min_row = 0xFFFFFFFFu, min_match = 0, best = 0; bool any = false, min_is_match = false;
      for (uint32_t k = 0; k < n; ++k) {
        const Slot<KeyT> sl = slots[lo[j] + k];
        const bool hit = sl.key == key[j];
        if (sl.rowid < min_row) { min_row = sl.rowid; min_is_match = hit; min_match = sl.rowid; }
        if (hit && (!any || sl.rowid > best)) { best = sl.rowid; any = true; }
      }
      if (!any) { acc.cmps += n; }
      else if (min_is_match) { acc.cmps += 1; nm[j] = 1; first[j] = min_match; }
      else {
        uint32_t rank = 0;                              // #tuples of the bucket inserted before `best`
        for (uint32_t k = 0; k < n; ++k) rank += slots[lo[j] + k].rowid < best;
        acc.cmps += n - rank + 1; nm[j] = 1; first[j] = best;
      }
    }
    mine += nm[j];
  }
  acc.matches = mine;
  // ---- output allocation: block exclusive scan + one atomic per tile
  unsigned long long pos = 0;
  if (WRITE) {
    unsigned long long tot;
    pos = block_exscan(mine, sm_scan, &tot);
    if (threadIdx.x == 0) sm_base = tot ? atomicAdd(&ctr->out_cursor, tot) : 0ull;
    __syncthreads();
    pos += sm_base;
  }
  if (WRITE || CHECKSUM) {
#pragma unroll
    for (int j = 0; j < kProbeItems; ++j) {
      if (nm[j] == 0) continue;
      const uint32_t left = src_leftid(s, t0 + j * kProbeThreads + threadIdx.x);
      if (nm[j] == 1) {
        if (CHECKSUM) { const uint64_t mx = pair_mix(left, first[j]); acc.sum += mx; acc.x ^= mx; }
        if (WRITE) { if (pos < out_cap) out[pos] = make_uint2(left, first[j]); ++pos; }
      } else {
        for (uint32_t k = 0; k < len[j]; ++k) {
          const Slot<KeyT> sl = slots[lo[j] + k];
          if (sl.key != key[j]) continue;
          if (CHECKSUM) { const uint64_t mx = pair_mix(left, sl.rowid); acc.sum += mx; acc.x ^= mx; }
          if (WRITE) { if (pos < out_cap) out[pos] = make_uint2(left, sl.rowid); ++pos; }
        }
      }
    }
  }
  commit_acc(acc, ctr, CHECKSUM);
}

// ---- nested probe ---------------------------------------------------------------------------------
template <int HASH, bool CHECKSUM, bool WRITE>
__global__ void __launch_bounds__(kProbeThreads)
k_probe_nested(Src s, Dir d, const uint2* __restrict__ tilemap, const uint32_t* __restrict__ goff,
               const Group<typename HashT<HASH>::key_t>* __restrict__ groups,
               uint2* __restrict__ out, unsigned long long out_cap, DevCounters* ctr) {
  using KeyT = typename HashT<HASH>::key_t;
  __shared__ unsigned long long sm_scan[33];
  __shared__ unsigned long long sm_base;
  uint64_t t0; uint32_t tn;
  block_tile<kProbeTile>(tilemap, s.n, t0, tn);
  ProbeAcc acc;
  uint32_t gref[kProbeItems], frow[kProbeItems];
  bool     hit[kProbeItems];
  unsigned long long mine = 0;
#pragma unroll
  for (int j = 0; j < kProbeItems; ++j) {
    const uint32_t li = j * kProbeThreads + threadIdx.x;
    const uint64_t i = t0 + li;
    hit[j] = false; gref[j] = 0; frow[j] = 0;
    if (li >= tn) continue;
    const KeyT key = src_key<KeyT>(s, i);
    const uint32_t b = HashT<HASH>::bucket(key, d);
    if (b - d.lo >= d.n_local) continue;
    const uint32_t glo = goff[b - d.lo], dk = goff[b - d.lo + 1] - glo;
    if (dk == 0) continue;                               // empty bucket: {nullptr, 0} (ht_nested.hh:372)
    uint32_t my_first = 0, my_g = 0; bool found = false;
    for (uint32_t k = 0; k < dk && !found; ++k) {
      const Group<KeyT> g = groups[glo + k];
      if (g.key == key) { found = true; my_first = g.first_row; my_g = glo + k; }
    }
    if (!found) { acc.cmps += dk; continue; }            // walked the whole main chain (ht_nested.hh:378-381)
    uint32_t before = 0;                                 // groups whose first tuple was inserted earlier
    for (uint32_t k = 0; k < dk; ++k) before += groups[glo + k].first_row < my_first;
    acc.cmps += before + 1;
    hit[j] = true; gref[j] = my_g; frow[j] = my_first; ++mine;
  }
  acc.matches = mine;
  unsigned long long pos = 0;
  if (WRITE) {
    unsigned long long tot;
    pos = block_exscan(mine, sm_scan, &tot);
    if (threadIdx.x == 0) sm_base = tot ? atomicAdd(&ctr->out_cursor, tot) : 0ull;
    __syncthreads();
    pos += sm_base;
  }
#pragma unroll
  for (int j = 0; j < kProbeItems; ++j) {
    if (!hit[j]) continue;
    const uint32_t left = src_leftid(s, t0 + j * kProbeThreads + threadIdx.x);
    if (CHECKSUM) { const uint64_t mx = pair_mix(left, frow[j]); acc.sum += mx; acc.x ^= mx; }
    if (WRITE) { if (pos < out_cap) out[pos] = make_uint2(left, gref[j]); ++pos; }
  }
  commit_acc(acc, ctr, CHECKSUM);
}

// ---- deferred unnest ------------------------------------------------------------------------------
// offsets[i] (exclusive scan of group lengths, n+1 entries) -> load-balanced expansion: every block
// produces kUnnestTile consecutive outputs, locating their source nested tuples by binary search.
constexpr int kUnnestThreads = 256;
constexpr int kUnnestItems   = 8;
constexpr int kUnnestTile    = kUnnestThreads * kUnnestItems;
constexpr int kUnnestSrcCap  = 2048;

__device__ __forceinline__ uint64_t upper_bound_u64(const unsigned long long* a, uint64_t n, unsigned long long v) {
  uint64_t lo = 0, hi = n;       // first index with a[idx] > v
  while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if (a[mid] <= v) lo = mid + 1; else hi = mid; }
  return lo;
}

template <class KeyT, bool CHECKSUM, bool WRITE>
__global__ void __launch_bounds__(kUnnestThreads)
k_unnest(const uint32_t* __restrict__ left, const uint32_t* __restrict__ gref, uint64_t n,
         const unsigned long long* __restrict__ offsets /* n+1 */,
         const Group<KeyT>* __restrict__ groups, const uint32_t* __restrict__ rows,
         uint2* __restrict__ out, unsigned long long out_cap, DevCounters* ctr) {
  __shared__ unsigned long long sm_off[kUnnestSrcCap + 1];
  __shared__ uint64_t sm_s0, sm_s1;
  const unsigned long long total = offsets[n];
  const unsigned long long t0 = (unsigned long long)blockIdx.x * kUnnestTile;
  if (t0 >= total) return;
  const unsigned long long t1 = (t0 + kUnnestTile < total) ? t0 + kUnnestTile : total;
  if (threadIdx.x == 0) sm_s0 = upper_bound_u64(offsets, n + 1, t0) - 1;       // source of output t0
  if (threadIdx.x == 32) sm_s1 = upper_bound_u64(offsets, n + 1, t1 - 1) - 1;  // source of output t1-1
  __syncthreads();
  const uint64_t s0 = sm_s0, s1 = sm_s1;
  const bool staged = (s1 - s0 + 1) <= (uint64_t)kUnnestSrcCap;
  if (staged) {
    for (uint64_t k = threadIdx.x; k <= s1 - s0 + 1; k += kUnnestThreads) sm_off[k] = offsets[s0 + k];
    __syncthreads();
  }
  ProbeAcc acc;
#pragma unroll
  for (int j = 0; j < kUnnestItems; ++j) {
    const unsigned long long o = t0 + threadIdx.x + (unsigned long long)j * kUnnestThreads;
    if (o >= t1) continue;
    uint64_t src;
    unsigned long long src_off;
    if (staged) {
      uint64_t k = upper_bound_u64(sm_off, s1 - s0 + 2, o) - 1;
      src = s0 + k; src_off = sm_off[k];
    } else {
      src = s0 + upper_bound_u64(offsets + s0, s1 - s0 + 2, o) - 1;
      src_off = offsets[src];
    }
    const Group<KeyT> g = groups[gref[src]];
    const uint32_t row = rows[g.start + (uint32_t)(o - src_off)];
    const uint32_t l = left[src];
    if (CHECKSUM) { const uint64_t mx = pair_mix(l, row); acc.sum += mx; acc.x ^= mx; }
    if (WRITE && o < out_cap) out[o] = make_uint2(l, row);
  }
  if (CHECKSUM) commit_acc(acc, ctr, true);
}

// ---- small column helpers -------------------------------------------------------------------------
template <class KeyT>
__global__ void k_group_first_row(const Group<KeyT>* __restrict__ groups, const uint32_t* __restrict__ gref, uint64_t n,
                                  uint32_t* __restrict__ out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = groups[gref[i]].first_row;
}
__global__ void k_gather_u32(const uint32_t* __restrict__ src, const uint32_t* __restrict__ idx, uint64_t n, uint32_t* __restrict__ dst) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[idx[i]];
}
__global__ void k_split_pairs(const uint2* __restrict__ pairs, uint64_t n, uint32_t* __restrict__ l, uint32_t* __restrict__ r) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const uint2 p = pairs[i]; l[i] = p.x; r[i] = p.y; }
}

}  // namespace hj3d
