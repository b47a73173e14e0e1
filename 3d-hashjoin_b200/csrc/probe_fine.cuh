// probe_fine.cuh -- the lean shared-memory probe for "at most one result per probe record" probes
// (IsBuildKeyUnique chaining probes, algebra.hh:653-655, and nested probes, algebra.hh:435-459).
//
// Same contract as k_probe_*_smem (probe_smem.cuh): one block handles one work item = (fine partition,
// <= chunk of its (key, id) records), keeps the partition's table slice in shared memory and writes dense
// result pairs.  What is different is the instruction budget.  ncu on the 2^27 x 2^30 join showed the probe
// side to be ISSUE bound, not HBM bound (~350 warp instructions per 32 probe records over the three
// passes, 50-68 % issue-slot utilisation), so this kernel is written for few instructions per record:
//   * 32-bit indexing inside the work item, full tiles take an unguarded path;
//   * the next tile's records are loaded (one LDG.64 per record) before the current tile is probed, so no
//     warp ever waits on a just-issued global load;
//   * results are counted with warp ballots and placed with ONE shared-memory word per warp and ONE global
//     atomic per tile of 1024 probe records (no 64-bit block scan);
//   * the per-thread counters are 32 bit inside a tile and folded into 64 bit once per tile;
//   * long (> kOrderedMax) buckets, which need the row-id based rules of probe.cuh, are out of line.
#pragma once

#include "common.cuh"
#include "probe.cuh"
#include "probe_smem.cuh"

namespace hj3d {

// Tile shape, measured at 2^27 x 2^30 (probe kernel ms): 8 items / 80 registers / 3 blocks per SM 6.21; 8 items / 64
// registers / 4 blocks 5.61; 4 items / 48 registers / 5 blocks 5.39; 4 items / 40 registers / 6 blocks 6.63 (spills).
// The kernel is latency bound (issue slots 56 % busy), so resident warps beat items per thread.
#ifndef HJ3D_FINE_ITEMS
#define HJ3D_FINE_ITEMS 4
#endif
#ifndef HJ3D_FINE_MINBLOCKS
#define HJ3D_FINE_MINBLOCKS 5
#endif
constexpr int kFineThreads = 256;
constexpr int kFineItems   = HJ3D_FINE_ITEMS;
constexpr int kFineTile    = kFineThreads * kFineItems;

// ---- out-of-line rules for buckets longer than kOrderedMax (unordered storage) -------------------
template <class KeyT>
__device__ __noinline__ uint32_t chain_first_long(KeyT key, const Slot<KeyT>* sp, uint32_t n, uint32_t* right, unsigned long long* cmps) {
  uint32_t min_row = 0xFFFFFFFFu, best = 0; bool any = false, min_is_match = false;
  for (uint32_t k = 0; k < n; ++k) {
    const Slot<KeyT> sl = sp[k];
    const bool hit = sl.key == key;
    if (sl.rowid < min_row) { min_row = sl.rowid; min_is_match = hit; }
    if (hit && (!any || sl.rowid > best)) { best = sl.rowid; any = true; }
  }
  if (!any) { *cmps += n; return 0; }
  if (min_is_match) { *cmps += 1; *right = min_row; return 1; }
  uint32_t rk = 0;                                      // #tuples of the bucket inserted before `best`
  for (uint32_t k = 0; k < n; ++k) rk += sp[k].rowid < best;
  *cmps += n - rk + 1; *right = best;
  return 1;
}
template <class KeyT>
__device__ __noinline__ uint32_t group_find_long(KeyT key, const Group<KeyT>* gp, uint32_t dk, uint32_t* gidx, uint32_t* first, unsigned long long* cmps) {
  uint32_t my_first = 0, my_g = 0; bool found = false;
  for (uint32_t k = 0; k < dk && !found; ++k) {
    const Group<KeyT> g = gp[k];
    if (g.key == key) { found = true; my_first = g.first_row; my_g = k; }
  }
  if (!found) { *cmps += dk; return 0; }
  uint32_t before = 0;                                  // groups whose first tuple was inserted earlier
  for (uint32_t k = 0; k < dk; ++k) before += gp[k].first_row < my_first;
  *cmps += before + 1; *gidx = my_g; *first = my_first;
  return 1;
}

// ---- one probe record against a bucket: the ordered walks of algebra.hh:644-657 / ht_nested.hh:368-381 -----
// KIND 0: right = build row id; KIND 1: right = global group index.  `first` = checksum partner.
template <class KeyT, int KIND, class RowT>
__device__ __forceinline__ uint32_t probe_bucket(KeyT key, const RowT* bp, uint32_t o0, uint32_t n, uint32_t& right, uint32_t& first,
                                                 uint32_t& cmps, unsigned long long& cmps_long) {
  if (n == 0) return 0;                                 // empty bucket: no comparison (algebra.hh:640-643, ht_nested.hh:372)
  if (n <= kOrderedMax) {
    uint32_t k = 0;
    for (;;) {
      if (KIND == 0) {
        const Slot<KeyT> sl = reinterpret_cast<const Slot<KeyT>*>(bp)[k];        // one 8 / 16 byte load
        ++k;
        if (sl.key == key) { right = sl.rowid; first = right; cmps += k; return 1; }
      } else {
        const KeyT bk = bp[k].key;
        ++k;
        if (bk == key) { right = o0 + k - 1; first = reinterpret_cast<const Group<KeyT>*>(bp)[k - 1].first_row; cmps += k; return 1; }
      }
      if (k == n) { cmps += n; return 0; }
    }
  }
  if (KIND == 0) {
    const uint32_t hit = chain_first_long<KeyT>(key, reinterpret_cast<const Slot<KeyT>*>(bp), n, &right, &cmps_long);
    first = right;
    return hit;
  } else {
    uint32_t gi = 0;
    const uint32_t hit = group_find_long<KeyT>(key, reinterpret_cast<const Group<KeyT>*>(bp), n, &gi, &first, &cmps_long);
    right = o0 + gi;
    return hit;
  }
}

// ---- the tile loop, instantiated for the shared-memory slice and for the global fallback ---------------
template <int HASH, int KIND, bool CHECKSUM, bool WRITE, class RowT>
__device__ __forceinline__ void probe_fine_items(const Slot<typename HashT<HASH>::key_t>* __restrict__ in, uint32_t n_rec, const Dir& d,
                                                 uint32_t bucket_base, uint32_t nbk, const uint32_t* offp, uint32_t row_base, const RowT* rowp,
                                                 uint2* __restrict__ out, unsigned long long out_cap, DevCounters* ctr,
                                                 ProbeAcc& acc, uint32_t* wsum, unsigned long long* sm_base) {
  using KeyT = typename HashT<HASH>::key_t;
  using SlotT = Slot<KeyT>;
  constexpr int IT = kFineItems, NW = kFineThreads / 32;
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
  KeyT     key[IT], nkey[IT];
  uint32_t id[IT], nid[IT];
  auto fetch = [&](uint32_t t0, KeyT* k, uint32_t* i) {
    if (t0 + kFineTile <= n_rec) {
#pragma unroll
      for (int j = 0; j < IT; ++j) { const SlotT r = in[t0 + j * kFineThreads + threadIdx.x]; k[j] = r.key; i[j] = r.rowid; }
    } else {
#pragma unroll
      for (int j = 0; j < IT; ++j) {
        const uint32_t li = t0 + j * kFineThreads + threadIdx.x;
        k[j] = 0; i[j] = 0;
        if (li < n_rec) { const SlotT r = in[li]; k[j] = r.key; i[j] = r.rowid; }
      }
    }
  };
  fetch(0, key, id);
  for (uint32_t t0 = 0; t0 < n_rec; t0 += kFineTile) {
    if (t0 + kFineTile < n_rec) fetch(t0 + kFineTile, nkey, nid);       // in flight while this tile is probed
    const uint32_t tn = n_rec - t0;                                      // >= kFineTile on full tiles
    uint32_t hitmask = 0, cmps = 0, wtot = 0;
    unsigned long long cmps_long = 0;
#pragma unroll
    for (int j = 0; j < IT; ++j) {
      uint32_t right = 0, first = 0, hit = 0;
      const uint32_t lb = HashT<HASH>::bucket(key[j], d) - bucket_base;
      if ((uint32_t)(j * kFineThreads + threadIdx.x) < tn && lb < nbk) {
        const uint32_t o0 = offp[lb], n = offp[lb + 1] - o0;
        hit = probe_bucket<KeyT, KIND, RowT>(key[j], rowp + (o0 - row_base), o0, n, right, first, cmps, cmps_long);
      }
      if (CHECKSUM && hit) { const uint64_t mx = pair_mix(id[j], first); acc.sum += mx; acc.x ^= mx; }
      hitmask |= hit << j;
      key[j] = (KeyT)right;                                              // the key's register now holds the result
      if (WRITE) wtot += __popc(__ballot_sync(0xffffffffu, hit));
    }
    acc.matches += __popc(hitmask);
    acc.cmps += (unsigned long long)cmps + cmps_long;
    if (WRITE) {
      if (lane == 0) wsum[warp] = wtot;
      __syncthreads();
      uint32_t before = 0, total = 0;
#pragma unroll
      for (int w = 0; w < NW; ++w) { const uint32_t v = wsum[w]; before += w < (int)warp ? v : 0u; total += v; }
      if (threadIdx.x == 0) *sm_base = total ? atomicAdd(&ctr->out_cursor, (unsigned long long)total) : 0ull;
      __syncthreads();
      unsigned long long pos = *sm_base + before;
#pragma unroll
      for (int j = 0; j < IT; ++j) {
        const uint32_t hit = (hitmask >> j) & 1u;
        const uint32_t bal = __ballot_sync(0xffffffffu, hit);
        const unsigned long long mypos = pos + __popc(bal & ((1u << lane) - 1u));
        if (hit && mypos < out_cap) out[mypos] = make_uint2(id[j], (uint32_t)key[j]);
        pos += __popc(bal);
      }
    }
#pragma unroll
    for (int j = 0; j < IT; ++j) { key[j] = nkey[j]; id[j] = nid[j]; }
  }
}

// KIND 0: chaining probe with IsBuildKeyUnique over (off, Slot rows); KIND 1: nested probe over (goff, Group rows).
template <int HASH, int KIND, bool CHECKSUM, bool WRITE>
// (the nested variant walks 16-byte group records and spills at 48 registers: 8.5 ms; 4 blocks / 64 registers suit it)
__global__ void __launch_bounds__(kFineThreads, KIND == 0 ? HJ3D_FINE_MINBLOCKS : HJ3D_FINE_MINBLOCKS - 1)
k_probe_fine(const Slot<typename HashT<HASH>::key_t>* __restrict__ recs, Dir d, FineCfg fc, const uint2* __restrict__ work,
             const uint32_t* __restrict__ work_part, const uint32_t* __restrict__ off, const void* __restrict__ rows_v,
             uint2* __restrict__ out, unsigned long long out_cap, DevCounters* ctr) {
  using KeyT = typename HashT<HASH>::key_t;
  using RowT = typename std::conditional<KIND == 0, Slot<KeyT>, Group<KeyT>>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ uint32_t wsum[kFineThreads / 32];
  __shared__ unsigned long long sm_base;
  const RowT* rows = reinterpret_cast<const RowT*>(rows_v);

  const uint2 w = work[blockIdx.x];
  const uint32_t f = work_part[blockIdx.x];
  const uint32_t blo = f * fc.width;
  const uint32_t bhi = (blo + fc.width < fc.n_local) ? blo + fc.width : fc.n_local;
  const uint32_t nbk = bhi - blo;
  const uint32_t rlo = off[blo], rhi = off[bhi];
  const uint32_t nrows = rhi - rlo;
  const uint32_t pre = sizeof(RowT) >= 16 ? 0u : (rlo & (uint32_t)(16 / sizeof(RowT) - 1));   // copy from the 16-byte aligned predecessor
  const uint32_t off_bytes = ((nbk + 1) * 4 + 15) & ~15u;
  // The slice arrives by two bulk asynchronous copies (TMA engine: cp.async.bulk, completion on an mbarrier) issued by ONE
  // thread -- no LDG / STS pairs, no per-thread address arithmetic.  The engine needs 16-byte aligned addresses and sizes: the
  // rows are copied from their 16-byte aligned predecessor on (`pre`), sizes are rounded up (the table's arrays are padded),
  // and a slice whose directory words are not 16-byte aligned (fine partitions of fewer than 4 buckets: test configurations
  // only) takes the global-memory path like a slice that does not fit.
  constexpr bool kBulk = sizeof(RowT) == 8 || sizeof(RowT) == 16;      // 24-byte group records (64-bit keys) are not 16-byte aligned
  const uint32_t row_bytes = ((nrows + pre) * (uint32_t)sizeof(RowT) + 15u) & ~15u;
  const bool fits = (uint64_t)off_bytes + (uint64_t)row_bytes <= fc.smem_bytes && (!kBulk || (((uintptr_t)(off + blo)) & 15u) == 0);
  uint32_t* sm_off = reinterpret_cast<uint32_t*>(smem_raw);
  RowT*     sm_rows = reinterpret_cast<RowT*>(smem_raw + off_bytes);
  __shared__ unsigned long long sm_bar;
  if (kBulk) {
    if (fits && threadIdx.x == 0) mbar_init(&sm_bar, 1);
    __syncthreads();
    if (fits) {
      if (threadIdx.x == 0) {
        mbar_expect_tx(&sm_bar, off_bytes + row_bytes);
        bulk_g2s(sm_off, off + blo, off_bytes, &sm_bar);
        if (row_bytes) bulk_g2s(sm_rows, rows + (rlo - pre), row_bytes, &sm_bar);
      }
      mbar_wait(&sm_bar, 0);
    }
  } else {
    if (fits) {
      copy_to_smem(sm_off, off + blo, (nbk + 1) * 4);
      copy_to_smem(sm_rows, rows + (rlo - pre), (nrows + pre) * (uint32_t)sizeof(RowT));
    }
    __syncthreads();
  }
  ProbeAcc acc;
  const Slot<KeyT>* in = recs + w.x;
  if (fits) probe_fine_items<HASH, KIND, CHECKSUM, WRITE, RowT>(in, w.y, d, d.lo + blo, nbk, sm_off, rlo - pre, sm_rows, out, out_cap, ctr, acc, wsum, &sm_base);
  else      probe_fine_items<HASH, KIND, CHECKSUM, WRITE, RowT>(in, w.y, d, d.lo + blo, nbk, off + blo, 0u, rows, out, out_cap, ctr, acc, wsum, &sm_base);
  commit_acc(acc, ctr, CHECKSUM);
}

}  // namespace hj3d
