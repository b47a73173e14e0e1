// probe.cuh -- probe-side kernels: chaining probe, nested probe (the deferred unnest lives in unnest.cuh).
//
// Every probe thread recomputes the counters the reference accumulates tuple-at-a-time:
//   matches  = AlgBase::_count of the probe operator,  num_cmps = _numCmps (algebra.hh:449,658).
// The build keeps short buckets (<= kOrderedMax entries) physically in the reference's own order:
//   chaining : chain order [t0, t_{n-1}, .., t_1] (new nodes are linked right after the directory entry,
//              ht_chaining.hh:189-194)  -> the probe is the same walk as algebra.hh:644-657: ++cmps per
//              node, emit on key match, stop at the first match if IsBuildKeyUnique
//   nested   : main chain in first-appearance order (tail append, ht_nested.hh:303-308) -> walk until the
//              key matches: index+1 comparisons on a hit, #distinct keys on a miss (ht_nested.hh:368-381)
// Longer buckets are left unordered and the same counters are derived from row ids ("oldest/newest" =
// smallest/largest row id = insertion order of the build strand).
//
// Results are written with block-aggregated allocation: one atomicAdd on the output cursor per
// block tile, pairs of a tile land contiguously (coalesced 8-byte stores).
#pragma once

#include "common.cuh"

namespace hj3d {

constexpr int kProbeThreads = 256;
constexpr int kProbeItems   = 4;
constexpr int kProbeTile    = kProbeThreads * kProbeItems;
constexpr uint32_t kOrderedMax = 16;   // buckets up to this length are stored in chain / first-appearance order

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

// (key, left id) of probe tuple i.  RECS: the input is an array of Slot<KeyT> records (partitioned or
// exchanged data) and is read with one vector load; otherwise a strided row-store tuple.
template <class KeyT, bool RECS>
__device__ __forceinline__ void load_probe(const Src& s, uint64_t i, KeyT& key, uint32_t& left) {
  if (RECS) {
    const Slot<KeyT> r = reinterpret_cast<const Slot<KeyT>*>(s.base)[i];
    key = r.key; left = r.rowid;
  } else {
    key = src_key<KeyT>(s, i);
    left = src_leftid(s, i);
  }
}

// One probe of a non-empty bucket of the chaining table: sp[0 .. n) are the bucket's slots.  nm = result tuples,
// first = row of the first one in chain order, cmps += comparisons (algebra.hh:644-657).  Shared by the probe kernels
// and the hot-key answers (hot.cuh).
template <class KeyT, bool UNIQUE>
__device__ __forceinline__ void chain_walk(const Slot<KeyT>* sp, uint32_t n, KeyT key, uint32_t& nm, uint32_t& first, uint32_t& cmps) {
  if (!UNIQUE) {
    uint32_t m = 0, f = 0;
    for (uint32_t k = 0; k < n; ++k) {
      const Slot<KeyT> sl = sp[k];
      if (sl.key == key) { if (m == 0) f = sl.rowid; ++m; }
    }
    nm = m; first = f;
    cmps += n;                                        // whole chain is walked (algebra.hh:644-657)
  } else if (n <= kOrderedMax) {
    uint32_t k = 0;                                   // chain order: stop at the first match (algebra.hh:653-655)
    for (; k < n; ++k) {
      const Slot<KeyT> sl = sp[k];
      if (sl.key == key) { nm = 1; first = sl.rowid; break; }
    }
    cmps += k < n ? k + 1 : n;
  } else {
    // unordered long bucket: first match in chain order [oldest, newest, .., second oldest] from row ids
    uint32_t min_row = 0xFFFFFFFFu, best = 0; bool any = false, min_is_match = false;
    for (uint32_t k = 0; k < n; ++k) {
      const Slot<KeyT> sl = sp[k];
      const bool hit = sl.key == key;
      if (sl.rowid < min_row) { min_row = sl.rowid; min_is_match = hit; }
      if (hit && (!any || sl.rowid > best)) { best = sl.rowid; any = true; }
    }
    if (!any) { cmps += n; }
    else if (min_is_match) { cmps += 1; nm = 1; first = min_row; }
    else {
      uint32_t rank = 0;                              // #tuples of the bucket inserted before `best`
      for (uint32_t k = 0; k < n; ++k) rank += sp[k].rowid < best;
      cmps += n - rank + 1; nm = 1; first = best;
    }
  }
}

// One probe of a non-empty bucket of the nested table: gp[0 .. dk) are the bucket's groups (the main chain).  hit / g
// (index inside the bucket) / frow (first row of the group), cmps += comparisons (ht_nested.hh:368-381).
template <class KeyT>
__device__ __forceinline__ void group_walk(const Group<KeyT>* gp, uint32_t dk, KeyT key, bool& hit, uint32_t& g, uint32_t& frow, uint32_t& cmps) {
  if (dk <= kOrderedMax) {
    uint32_t k = 0;                                    // first-appearance order: the walk of ht_nested.hh:371-379
    for (; k < dk; ++k) {
      const Group<KeyT> gr = gp[k];
      if (gr.key == key) { hit = true; g = k; frow = gr.first_row; break; }
    }
    cmps += k < dk ? k + 1 : dk;
  } else {
    uint32_t my_first = 0, my_g = 0; bool found = false;
    for (uint32_t k = 0; k < dk && !found; ++k) {
      const Group<KeyT> gr = gp[k];
      if (gr.key == key) { found = true; my_first = gr.first_row; my_g = k; }
    }
    if (!found) { cmps += dk; return; }
    uint32_t before = 0;                               // groups whose first tuple was inserted earlier
    for (uint32_t k = 0; k < dk; ++k) before += gp[k].first_row < my_first;
    cmps += before + 1;
    hit = true; g = my_g; frow = my_first;
  }
}

// ---- one tile of chaining probes ----------------------------------------------------------------------
// offp[b] .. offp[b+1] (minus slot_base) delimit bucket b's slots in slotp; both may live in shared or
// global memory (the address space is known at every inlined call site).
template <class KeyT, bool UNIQUE, bool CHECKSUM, bool WRITE, bool RECS, int THREADS, int ITEMS, int HASH>
__device__ __forceinline__ void probe_chaining_tile(const Src& s, const Dir& d, uint64_t t0, uint32_t tn,
                                                    uint32_t bucket_base, uint32_t n_buckets,
                                                    const uint32_t* offp, uint32_t slot_base, const Slot<KeyT>* slotp,
                                                    uint2* __restrict__ out, unsigned long long out_cap, DevCounters* ctr,
                                                    ProbeAcc& acc, unsigned long long* sm_scan, unsigned long long* sm_base) {
  KeyT     key[ITEMS];
  uint32_t left[ITEMS], lo[ITEMS], len[ITEMS], nm[ITEMS], first[ITEMS];
  unsigned long long mine = 0;   // a tile's matches can exceed 2^32 when many probes hit one hot key
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const uint32_t li = j * THREADS + threadIdx.x;
    lo[j] = len[j] = nm[j] = first[j] = left[j] = 0; key[j] = 0;
    if (li < tn) load_probe<KeyT, RECS>(s, t0 + li, key[j], left[j]);
  }
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const uint32_t li = j * THREADS + threadIdx.x;
    if (li < tn) {
      const uint32_t b = HashT<HASH>::bucket(key[j], d) - bucket_base;
      if (b < n_buckets) { const uint32_t o0 = offp[b]; lo[j] = o0 - slot_base; len[j] = offp[b + 1] - o0; }
    }
  }
  uint32_t cmps = 0;
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const uint32_t n = len[j];
    if (n == 0) continue;                               // empty bucket: no comparison (algebra.hh:640-643)
    chain_walk<KeyT, UNIQUE>(slotp + lo[j], n, key[j], nm[j], first[j], cmps);
    mine += nm[j];
  }
  acc.matches += mine;
  acc.cmps += cmps;
  // ---- output allocation: block exclusive scan + one atomic per tile
  unsigned long long pos = 0;
  if (WRITE) {
    unsigned long long tot;
    const unsigned long long ex = block_exscan(mine, sm_scan, &tot);
    if (threadIdx.x == 0) *sm_base = tot ? atomicAdd(&ctr->out_cursor, tot) : 0ull;
    __syncthreads();
    pos = *sm_base + ex;
  }
  if (WRITE || CHECKSUM) {
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      if (nm[j] == 0) continue;
      if (nm[j] == 1) {
        if (CHECKSUM) { const uint64_t mx = pair_mix(left[j], first[j]); acc.sum += mx; acc.x ^= mx; }
        if (WRITE) { if (pos < out_cap) out[pos] = make_uint2(left[j], first[j]); ++pos; }
      } else {
        const Slot<KeyT>* sp = slotp + lo[j];
        for (uint32_t k = 0; k < len[j]; ++k) {
          const Slot<KeyT> sl = sp[k];
          if (sl.key != key[j]) continue;
          if (CHECKSUM) { const uint64_t mx = pair_mix(left[j], sl.rowid); acc.sum += mx; acc.x ^= mx; }
          if (WRITE) { if (pos < out_cap) out[pos] = make_uint2(left[j], sl.rowid); ++pos; }
        }
      }
    }
  }
}

// ---- one tile of nested probes ------------------------------------------------------------------------
template <class KeyT, bool CHECKSUM, bool WRITE, bool RECS, int THREADS, int ITEMS, int HASH>
__device__ __forceinline__ void probe_nested_tile(const Src& s, const Dir& d, uint64_t t0, uint32_t tn,
                                                  uint32_t bucket_base, uint32_t n_buckets,
                                                  const uint32_t* offp, uint32_t group_base, const Group<KeyT>* grp,
                                                  uint2* __restrict__ out, unsigned long long out_cap, DevCounters* ctr,
                                                  ProbeAcc& acc, unsigned long long* sm_scan, unsigned long long* sm_base) {
  uint32_t left[ITEMS], gref[ITEMS], frow[ITEMS];
  bool     hit[ITEMS];
  uint32_t mine = 0, cmps = 0;
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const uint32_t li = j * THREADS + threadIdx.x;
    hit[j] = false; gref[j] = 0; frow[j] = 0; left[j] = 0;
    if (li >= tn) continue;
    KeyT key;
    load_probe<KeyT, RECS>(s, t0 + li, key, left[j]);
    const uint32_t b = HashT<HASH>::bucket(key, d) - bucket_base;
    if (b >= n_buckets) continue;
    const uint32_t o0 = offp[b], dk = offp[b + 1] - o0;
    if (dk == 0) continue;                               // empty bucket: {nullptr, 0} (ht_nested.hh:372)
    uint32_t g = 0;
    group_walk<KeyT>(grp + (o0 - group_base), dk, key, hit[j], g, frow[j], cmps);
    if (hit[j]) gref[j] = o0 + g;
    mine += hit[j];
  }
  acc.matches += mine;
  acc.cmps += cmps;
  unsigned long long pos = 0;
  if (WRITE) {
    unsigned long long tot;
    const unsigned long long ex = block_exscan((unsigned long long)mine, sm_scan, &tot);
    if (threadIdx.x == 0) *sm_base = tot ? atomicAdd(&ctr->out_cursor, tot) : 0ull;
    __syncthreads();
    pos = *sm_base + ex;
  }
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    if (!hit[j]) continue;
    if (CHECKSUM) { const uint64_t mx = pair_mix(left[j], frow[j]); acc.sum += mx; acc.x ^= mx; }
    if (WRITE) { if (pos < out_cap) out[pos] = make_uint2(left[j], gref[j]); ++pos; }
  }
}

// ---- global-memory lookups (small inputs / fallback) ---------------------------------------------------
template <int HASH, bool UNIQUE, bool CHECKSUM, bool WRITE, bool RECS>
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
  probe_chaining_tile<KeyT, UNIQUE, CHECKSUM, WRITE, RECS, kProbeThreads, kProbeItems, HASH>(
      s, d, t0, tn, d.lo, d.n_local, off, 0u, slots, out, out_cap, ctr, acc, sm_scan, &sm_base);
  commit_acc(acc, ctr, CHECKSUM);
}

template <int HASH, bool CHECKSUM, bool WRITE, bool RECS>
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
  probe_nested_tile<KeyT, CHECKSUM, WRITE, RECS, kProbeThreads, kProbeItems, HASH>(
      s, d, t0, tn, d.lo, d.n_local, goff, 0u, groups, out, out_cap, ctr, acc, sm_scan, &sm_base);
  commit_acc(acc, ctr, CHECKSUM);
}

// ---- small column helpers -------------------------------------------------------------------------
template <class KeyT>
__global__ void k_group_first_row(const Group<KeyT>* __restrict__ groups, const uint32_t* __restrict__ gref, uint64_t n,
                                  uint32_t* __restrict__ out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = groups[gref[i]].first_row;
}
static __global__ void k_gather_u32(const uint32_t* __restrict__ src, const uint32_t* __restrict__ idx, uint64_t n, uint32_t* __restrict__ dst) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[idx[i]];
}
static __global__ void k_iota_u32(uint32_t* __restrict__ dst, uint64_t n, uint32_t first) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = first + (uint32_t)i;
}
static __global__ void k_split_pairs(const uint2* __restrict__ pairs, uint64_t n, uint32_t* __restrict__ l, uint32_t* __restrict__ r) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const uint2 p = pairs[i]; l[i] = p.x; r[i] = p.y; }
}

}  // namespace hj3d
