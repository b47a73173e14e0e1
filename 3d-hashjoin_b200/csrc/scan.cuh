// scan.cuh -- device-wide prefix sum (reduce / scan-block-sums / apply), with the statistics
// reduction of makeStatistics (ht_chaining.hh:260-292, ht_nested.hh:450-482) fused into the
// reduce pass.  HBM-bound streaming kernels: 2 reads + 1 write of the scanned array.
#pragma once

#include "common.cuh"

namespace hj3d {

constexpr int kScanThreads = 512;
constexpr int kScanItems   = 16;
constexpr int kScanTile    = kScanThreads * kScanItems;

// Loader:  T operator()(uint64_t i) const
// Storer:  void operator()(uint64_t i, T exclusive_prefix, T element) const

template <class T> __device__ __forceinline__ void stats_step(DevAgg& a, T x) {
  unsigned long long v = (unsigned long long)x;
  a.mn = v < a.mn ? v : a.mn;
  a.mx = v > a.mx ? v : a.mx;
  a.sum += v; a.sumsq += v * v; a.cnt += 1;
}

__device__ __forceinline__ void agg_commit(DevAgg& loc, DevAgg* glob, unsigned long long* smem /*>= 5*32*/) {
  // warp reduce, then one set of atomics per block
  DevAgg r;
  r.mn = warp_min(loc.mn); r.mx = warp_max(loc.mx); r.sum = warp_sum(loc.sum);
  r.sumsq = warp_sum(loc.sumsq); r.cnt = warp_sum(loc.cnt);
  const uint32_t w = threadIdx.x >> 5, l = lane_id(), nw = blockDim.x >> 5;
  __syncthreads();
  if (l == 0) { smem[w] = r.mn; smem[32 + w] = r.mx; smem[64 + w] = r.sum; smem[96 + w] = r.sumsq; smem[128 + w] = r.cnt; }
  __syncthreads();
  if (w == 0) {
    unsigned long long mn = l < nw ? smem[l] : ~0ull, mx = l < nw ? smem[32 + l] : 0ull;
    unsigned long long s = l < nw ? smem[64 + l] : 0ull, sq = l < nw ? smem[96 + l] : 0ull, c = l < nw ? smem[128 + l] : 0ull;
    mn = warp_min(mn); mx = warp_max(mx); s = warp_sum(s); sq = warp_sum(sq); c = warp_sum(c);
    if (l == 0 && c) {
      atomicMin(&glob->mn, mn); atomicMax(&glob->mx, mx);
      atomicAdd(&glob->sum, s); atomicAdd(&glob->sumsq, sq); atomicAdd(&glob->cnt, c);
    }
  }
}

template <class T, class Loader, bool STATS>
__global__ void __launch_bounds__(kScanThreads) k_scan_reduce(Loader load, uint64_t n, T* block_sums, DevStats* stats) {
  __shared__ unsigned long long sm[160];
  const uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanItems;
  T sum = 0;
  DevAgg all{~0ull, 0, 0, 0, 0}, ne{~0ull, 0, 0, 0, 0};
  unsigned long long empty = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    uint64_t i = base + j;
    if (i < n) {
      T v = load(i);
      sum += v;
      if (STATS) {
        stats_step(all, v);
        if (v != 0) stats_step(ne, v); else ++empty;
      }
    }
  }
  // block reduce of sum
  T ws = warp_sum(sum);
  __shared__ T wsum[32];
  const uint32_t w = threadIdx.x >> 5, l = lane_id();
  if (l == 0) wsum[w] = ws;
  __syncthreads();
  if (w == 0) {
    T s = l < (kScanThreads >> 5) ? wsum[l] : T(0);
    s = warp_sum(s);
    if (l == 0) block_sums[blockIdx.x] = s;
  }
  if (STATS) {
    agg_commit(all, &stats->all, sm);
    agg_commit(ne, &stats->nonempty, sm);
    unsigned long long e = warp_sum(empty);
    if (l == 0 && e) atomicAdd(&stats->empty, e);
  }
}

// single block: exclusive scan of block_sums in place, total -> *total
template <class T>
__global__ void __launch_bounds__(1024) k_scan_blocksums(T* block_sums, uint32_t nb, T* total) {
  __shared__ T sm[33];
  T carry = 0;
  for (uint32_t base = 0; base < nb; base += 1024) {
    uint32_t i = base + threadIdx.x;
    T v = i < nb ? block_sums[i] : T(0);
    T tot;
    T ex = block_exscan(v, sm, &tot);
    if (i < nb) block_sums[i] = carry + ex;
    carry += tot;
  }
  if (threadIdx.x == 0 && total) *total = carry;
}

template <class T, class Loader, class Storer>
__global__ void __launch_bounds__(kScanThreads) k_scan_apply(Loader load, Storer store, uint64_t n, const T* block_offsets) {
  __shared__ T sm[33];
  const uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanItems;
  T v[kScanItems];
  T tsum = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    uint64_t i = base + j;
    v[j] = i < n ? load(i) : T(0);
    tsum += v[j];
  }
  T tot;
  T ex = block_exscan(tsum, sm, &tot) + block_offsets[blockIdx.x];
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    uint64_t i = base + j;
    if (i < n) store(i, ex, v[j]);
    ex += v[j];
  }
}

}  // namespace hj3d
