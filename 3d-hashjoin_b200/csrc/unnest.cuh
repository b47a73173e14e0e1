// unnest.cuh -- deferred unnesting (AlgUnnestHt::step, algebra.hh:510-541): a nested tuple (left, group ref)
// becomes one flat pair (left, build row) for the MainNode's own tuple and every SubNode of its sub chain.
//
// Warp-cooperative expansion.  A block owns 2048 consecutive nested tuples (a warp: 256, in 8 rounds of 32):
//   k_unnest_count   sums the group lengths per block (+ the longest group),
//   (device scan of the block sums -> every block's first output position),
//   k_unnest_expand  reloads (start, len), scans the lengths inside the warp and writes the pairs of a round
//                    32 at a time: output lane o finds its source tuple by a 5-step shuffle search over the
//                    round's exclusive offsets, so the stores are dense and coalesced whatever the group
//                    sizes are; a round whose groups all have one row (key/foreign-key build side) skips the
//                    search.
// Group records are read through gref, but the nested probe emits its results partition by partition, so
// these reads hit L1/L2.  Inputs with a group longer than kUnnestWarpMax rows (hot keys) take the
// element-balanced kernel of probe.cuh (k_unnest) instead.
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

template <class KeyT>
__global__ void __launch_bounds__(kUxThreads)
k_unnest_count(NestedIn in, uint64_t n, const Group<KeyT>* __restrict__ groups,
               unsigned long long* __restrict__ block_sums, unsigned long long* __restrict__ max_len) {
  __shared__ unsigned long long sm[kUxThreads / 32];
  const uint64_t base = (uint64_t)blockIdx.x * kUxTile + (threadIdx.x >> 5) * (32 * kUxRounds) + lane_id();
  unsigned long long sum = 0; uint32_t mx = 0;
#pragma unroll
  for (int j = 0; j < kUxRounds; ++j) {
    const uint64_t i = base + j * 32;
    if (i < n) { const uint32_t len = group_start_len<KeyT>(groups, in.g(i)).y; sum += len; mx = len > mx ? len : mx; }
  }
  sum = warp_sum(sum); mx = warp_max(mx);
  if (lane_id() == 0) { sm[threadIdx.x >> 5] = sum; if (mx > kUnnestWarpMax) atomicMax(max_len, (unsigned long long)mx); }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
#pragma unroll
    for (int w = 0; w < kUxThreads / 32; ++w) t += sm[w];
    block_sums[blockIdx.x] = t;
  }
}

template <class KeyT, bool CHECKSUM, bool WRITE>
__global__ void __launch_bounds__(kUxThreads)
k_unnest_expand(NestedIn in, uint64_t n,
                const Group<KeyT>* __restrict__ groups, const uint32_t* __restrict__ rows,
                const unsigned long long* __restrict__ block_base, uint2* __restrict__ out, unsigned long long out_cap, DevCounters* ctr) {
  __shared__ unsigned long long sm[kUxThreads / 32];
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
  const uint64_t base = (uint64_t)blockIdx.x * kUxTile + warp * (32 * kUxRounds) + lane;
  uint32_t lf[kUxRounds], st[kUxRounds], ex[kUxRounds], tot[kUxRounds];
  unsigned long long wsum = 0;
  uint32_t all_one = 1;
#pragma unroll
  for (int j = 0; j < kUxRounds; ++j) {
    const uint64_t i = base + j * 32;
    uint32_t len = 0; lf[j] = 0; st[j] = 0;
    if (i < n) { const uint2 t = in.lg(i); const uint2 g = group_start_len<KeyT>(groups, t.y); st[j] = g.x; len = g.y; lf[j] = t.x; }
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
  unsigned long long pos = block_base[blockIdx.x];
#pragma unroll
  for (int w = 0; w < kUxThreads / 32; ++w) pos += w < (int)warp ? sm[w] : 0ull;
  ProbeAcc acc;
#pragma unroll
  for (int j = 0; j < kUxRounds; ++j) {
    const uint32_t T = tot[j];
    if (all_one) {                                        // every group of the warp's tile has one row: position = rank
      const bool have = (base + j * 32) < n;
      if (have) {
        const uint32_t row = __ldg(rows + st[j]);
        if (CHECKSUM) { const uint64_t mx = pair_mix(lf[j], row); acc.sum += mx; acc.x ^= mx; }
        if (WRITE && pos + ex[j] < out_cap) out[pos + ex[j]] = make_uint2(lf[j], row);
      }
    } else {
      for (uint32_t o = 0; o < T; o += 32) {
        const uint32_t idx = o + lane;
        uint32_t s = 0;                                   // largest s with ex[s] <= idx
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
  if (CHECKSUM) commit_acc(acc, ctr, true);
}

}  // namespace hj3d
