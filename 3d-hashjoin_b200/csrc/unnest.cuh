// unnest.cuh -- deferred unnesting (AlgUnnestHt::step, algebra.hh:510-541): a nested tuple (left, group ref)
// becomes one flat pair (left, build row) for the MainNode's own tuple and every SubNode of its sub chain.
//
// ONE pass, warp-cooperative expansion.  A block owns 2048 consecutive nested tuples (a warp: 256, in 8 rounds of 32):
// it reads (start, len) of every tuple's group, scans the lengths inside the warp, reserves the block's output range with
// one atomic on the result cursor (result order is unspecified, so no global prefix sum is needed) and writes the pairs
// of a round 32 at a time: output lane o finds its source tuple by a 5-step shuffle search over the round's exclusive
// offsets, so the stores are dense and coalesced whatever the group sizes are; a warp whose groups all have one row
// (key/foreign-key build side) skips the search.  Tuples whose group is longer than kUnnestWarpMax rows (hot keys)
// are not expanded by their warp: their indices go to a hot list that k_unnest_hot expands afterwards with one block per
// tuple.  Group records are read through gref, but the nested probe emits its results partition by partition, so these
// reads hit L1/L2.
#pragma once

#include "common.cuh"
#include "probe.cuh"

namespace hj3d {

constexpr int kUxThreads = 256;
constexpr int kUxRounds  = 8;
constexpr int kUxTile    = kUxThreads * kUxRounds;      // nested tuples per block
constexpr uint32_t kUnnestWarpMax = 4096;               // longest group the warp-cooperative kernel takes

template <class KeyT> __device__ __forceinline__ uint2 group_start_len(const Group<KeyT>* groups, uint32_t g) {
  const Group<KeyT>& r = groups[g];
  return make_uint2(r.start, r.len);
}
template <> __device__ __forceinline__ uint2 group_start_len<uint32_t>(const Group<uint32_t>* groups, uint32_t g) {
  return __ldg(reinterpret_cast<const uint2*>(groups + g) + 1);           // {key, first_row | start, len}: one 8-byte load
}

// nested tuples come either as two columns (left[], gref[]) or as the (left, gref) pairs the nested probe wrote
struct NestedIn {
  const uint32_t* left; const uint32_t* gref; const uint2* pairs;
  __device__ __forceinline__ uint32_t g(uint64_t i) const { return pairs ? __ldg(pairs + i).y : __ldg(gref + i); }
  __device__ __forceinline__ uint2 lg(uint64_t i) const { return pairs ? __ldg(pairs + i) : make_uint2(__ldg(left + i), __ldg(gref + i)); }
};

template <class KeyT, bool CHECKSUM, bool WRITE>
__global__ void __launch_bounds__(kUxThreads)
k_unnest_expand(NestedIn in, uint64_t n, const Group<KeyT>* __restrict__ groups, const uint32_t* __restrict__ rows,
                uint2* __restrict__ out, unsigned long long out_cap, DevCounters* ctr,
                uint32_t* __restrict__ hot_list, uint32_t hot_cap, unsigned long long* hot_count) {
  __shared__ unsigned long long sm[kUxThreads / 32];
  __shared__ unsigned long long sm_base;
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
  const uint64_t base = (uint64_t)blockIdx.x * kUxTile + warp * (32 * kUxRounds) + lane;
  uint32_t lf[kUxRounds], st[kUxRounds], ex[kUxRounds], tot[kUxRounds];
  unsigned long long wsum = 0;
  uint32_t all_one = 1, nz = 0;
#pragma unroll
  for (int j = 0; j < kUxRounds; ++j) {
    const uint64_t i = base + j * 32;
    uint32_t len = 0; lf[j] = 0; st[j] = 0;
    if (i < n) { const uint2 t = in.lg(i); const uint2 g = group_start_len<KeyT>(groups, t.y); st[j] = g.x; len = g.y; lf[j] = t.x; }
    if (len > kUnnestWarpMax) {                           // hot key: expanded by a whole block later (k_unnest_hot)
      const unsigned long long h = atomicAdd(hot_count, 1ull);
      if (h < hot_cap) hot_list[h] = (uint32_t)i;
      len = 0;
    }
    nz |= (len != 0 ? 1u : 0u) << j;
    uint32_t inc = len;                                   // inclusive scan of the lengths of this round
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= (uint32_t)o) inc += v; }
    ex[j] = inc - len;
    tot[j] = __shfl_sync(0xffffffffu, inc, 31);
    wsum += tot[j];
    all_one &= (uint32_t)__all_sync(0xffffffffu, len <= 1u);
  }
  if (lane == 0) sm[warp] = wsum;
  __syncthreads();
  unsigned long long before = 0, total = 0;
#pragma unroll
  for (int w = 0; w < kUxThreads / 32; ++w) { const unsigned long long v = sm[w]; before += w < (int)warp ? v : 0ull; total += v; }
  if (threadIdx.x == 0) sm_base = total ? atomicAdd(&ctr->out_cursor, total) : 0ull;   // one reservation per 2048 nested tuples
  __syncthreads();
  unsigned long long pos = sm_base + before;
  ProbeAcc acc;
  if (WRITE || CHECKSUM) {
#pragma unroll
    for (int j = 0; j < kUxRounds; ++j) {
      const uint32_t T = tot[j];
      if (all_one) {                                      // every group of the warp's tile has one row: position = rank
        if ((nz >> j) & 1u) {                             // my length is 1 (0: padding or a deferred hot tuple)
          const uint32_t row = __ldg(rows + st[j]);
          if (CHECKSUM) { const uint64_t mx = pair_mix(lf[j], row); acc.sum += mx; acc.x ^= mx; }
          if (WRITE && pos + ex[j] < out_cap) out[pos + ex[j]] = make_uint2(lf[j], row);
        }
      } else {
        for (uint32_t o = 0; o < T; o += 32) {
          const uint32_t idx = o + lane;
          uint32_t s = 0;                                 // largest s with ex[s] <= idx
#pragma unroll
          for (int step = 16; step > 0; step >>= 1) {
            const uint32_t v = __shfl_sync(0xffffffffu, ex[j], (s + step) & 31);
            if (v <= idx) s += step;
          }
          const uint32_t e = __shfl_sync(0xffffffffu, ex[j], s);
          const uint32_t b = __shfl_sync(0xffffffffu, st[j], s);
          const uint32_t l = __shfl_sync(0xffffffffu, lf[j], s);
          if (idx < T) {
            const uint32_t row = __ldg(rows + b + (idx - e));
            if (CHECKSUM) { const uint64_t mx = pair_mix(l, row); acc.sum += mx; acc.x ^= mx; }
            if (WRITE && pos + idx < out_cap) out[pos + idx] = make_uint2(l, row);
          }
        }
      }
      pos += T;
    }
  }
  if (CHECKSUM) commit_acc(acc, ctr, true);
}

// hot groups: one block per listed nested tuple, rows copied cooperatively
template <class KeyT, bool CHECKSUM, bool WRITE>
__global__ void __launch_bounds__(kUxThreads)
k_unnest_hot(NestedIn in, const uint32_t* __restrict__ hot_list, uint32_t n_hot, const Group<KeyT>* __restrict__ groups,
             const uint32_t* __restrict__ rows, uint2* __restrict__ out, unsigned long long out_cap, DevCounters* ctr) {
  __shared__ unsigned long long sm_base;
  ProbeAcc acc;
  for (uint32_t h = blockIdx.x; h < n_hot; h += gridDim.x) {
    const uint2 t = in.lg(hot_list[h]);
    const uint2 g = group_start_len<KeyT>(groups, t.y);
    __syncthreads();
    if (threadIdx.x == 0) sm_base = atomicAdd(&ctr->out_cursor, (unsigned long long)g.y);
    __syncthreads();
    const unsigned long long pos = sm_base;
    if (WRITE || CHECKSUM) {
      for (uint32_t r = threadIdx.x; r < g.y; r += kUxThreads) {
        const uint32_t row = __ldg(rows + g.x + r);
        if (CHECKSUM) { const uint64_t mx = pair_mix(t.x, row); acc.sum += mx; acc.x ^= mx; }
        if (WRITE && pos + r < out_cap) out[pos + r] = make_uint2(t.x, row);
      }
    }
  }
  if (CHECKSUM) commit_acc(acc, ctr, true);
}

}  // namespace hj3d
