// partition.cuh -- bucket-range partitioning of a row-store relation into (key, row id) records.
//
// part(t) = bucket(t) / width with width = ceil(D / n_parts): contiguous bucket ranges, so a
// whole bucket (chain / key group) always falls into one partition.  Used
//   - inside one GPU: bucket-ordering both inputs turns every later pass (histogram, scatter,
//     grouping, probe, unnest) from random HBM sector traffic over a multi-GB directory into
//     accesses to a moving window of a few MB that stays resident in the 126 MB L2;
//   - across GPUs (hj3d_partition_by_owner): the records of partition g are sent to GPU g, which
//     owns directory range [g*width, (g+1)*width)   (SURVEY.md 8(e)).
//
// One streaming pass in the common case: every partition gets a fixed-capacity region
// (expected size + slack); a block ranks its tile's records per partition in shared memory,
// reserves one contiguous range per (block, partition) with a single atomic and writes the runs.
// If a region overflows (skewed keys), the per-partition cursors still hold the exact counts, so
// the pass is simply repeated into exact-size regions (classic histogram + scatter, with the
// histogram already known).
#pragma once

#include "common.cuh"

namespace hj3d {

constexpr int kPartThreads = 256;
constexpr int kPartItems   = 16;
constexpr int kPartTile    = kPartThreads * kPartItems;
constexpr int kMaxParts    = 1024;

struct PartFn {
  uint32_t width;      // buckets per partition
  uint32_t shift;      // log2(width) if pow2
  uint32_t is_pow2;
  uint32_t lo;         // first bucket of the (shard) directory
  __device__ __forceinline__ uint32_t operator()(uint32_t bucket) const {
    const uint32_t b = bucket - lo;
    return is_pow2 ? (b >> shift) : (b / width);
  }
};

inline PartFn make_partfn(uint32_t width, uint32_t lo) {
  PartFn f; f.width = width; f.is_pow2 = (width & (width - 1)) == 0; f.shift = 0; f.lo = lo;
  while ((1u << f.shift) < width) ++f.shift;
  return f;
}

template <int HASH>
__global__ void __launch_bounds__(kPartThreads)
k_part_hist(Src s, Dir d, PartFn pf, uint32_t n_parts, unsigned long long* __restrict__ counts) {
  using KeyT = typename HashT<HASH>::key_t;
  __shared__ uint32_t h[kMaxParts];
  for (uint32_t p = threadIdx.x; p < n_parts; p += kPartThreads) h[p] = 0;
  __syncthreads();
  const uint64_t base = (uint64_t)blockIdx.x * kPartTile + threadIdx.x;
#pragma unroll 4
  for (int j = 0; j < kPartItems; ++j) {
    const uint64_t i = base + (uint64_t)j * kPartThreads;
    if (i < s.n) {
      const uint32_t p = pf(HashT<HASH>::bucket(src_key<KeyT>(s, i), d));
      if (p < n_parts) atomicAdd(&h[p], 1u);
    }
  }
  __syncthreads();
  for (uint32_t p = threadIdx.x; p < n_parts; p += kPartThreads)
    if (h[p]) atomicAdd(&counts[p], (unsigned long long)h[p]);
}

// part_start[p] = exclusive prefix of counts (single thread; n_parts <= 1024)
__global__ void k_part_prefix(const unsigned long long* __restrict__ counts, uint32_t n_parts,
                              unsigned long long* __restrict__ part_start) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    unsigned long long run = 0;
    for (uint32_t p = 0; p < n_parts; ++p) { part_start[p] = run; run += counts[p]; }
  }
}
__global__ void k_part_fixed_starts(uint32_t n_parts, unsigned long long cap, unsigned long long* __restrict__ part_start) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n_parts) part_start[p] = (unsigned long long)p * cap;
}

// records of partition p go to out[part_start[p] + k], k = running cursor[p]; k >= cap is dropped
// (the cursor keeps counting, so it ends up holding the exact partition size either way).
// LEFTID: the stored id is the probe-side "left" id (position, or the tuple's own id) instead of the
// build-side row id + rowid_base.  Records of buckets outside the shard (p >= n_parts) are dropped.
template <int HASH, bool LEFTID>
__global__ void __launch_bounds__(kPartThreads)
k_part_scatter(Src s, Dir d, PartFn pf, uint32_t n_parts, uint32_t rowid_base, unsigned long long cap,
               const unsigned long long* __restrict__ part_start, unsigned long long* __restrict__ cursor,
               Slot<typename HashT<HASH>::key_t>* __restrict__ out) {
  using KeyT = typename HashT<HASH>::key_t;
  __shared__ uint32_t h[kMaxParts];
  __shared__ unsigned long long basepos[kMaxParts];
  for (uint32_t p = threadIdx.x; p < n_parts; p += kPartThreads) h[p] = 0;
  __syncthreads();
  const uint64_t base = (uint64_t)blockIdx.x * kPartTile + threadIdx.x;
  KeyT     key[kPartItems];
  uint32_t part[kPartItems], rank[kPartItems];
#pragma unroll
  for (int j = 0; j < kPartItems; ++j) {
    const uint64_t i = base + (uint64_t)j * kPartThreads;
    part[j] = 0xFFFFFFFFu; key[j] = 0; rank[j] = 0;
    if (i < s.n) {
      key[j] = src_key<KeyT>(s, i);
      const uint32_t p = pf(HashT<HASH>::bucket(key[j], d));
      if (p < n_parts) { part[j] = p; rank[j] = atomicAdd(&h[p], 1u); }
    }
  }
  __syncthreads();
  for (uint32_t p = threadIdx.x; p < n_parts; p += kPartThreads)
    basepos[p] = h[p] ? atomicAdd(&cursor[p], (unsigned long long)h[p]) : 0ull;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kPartItems; ++j) {
    if (part[j] == 0xFFFFFFFFu) continue;
    const uint64_t i = base + (uint64_t)j * kPartThreads;
    const unsigned long long k = basepos[part[j]] + rank[j];
    if (k >= cap) continue;
    Slot<KeyT> r; r.key = key[j];
    r.rowid = LEFTID ? src_leftid(s, i) : src_rowid(s, i) + rowid_base;
    out[part_start[part[j]] + k] = r;
  }
}

// ---- tile maps: block -> (first record, count) over partition regions with gaps --------------------
// tile_prefix[p] = number of tiles of partitions < p;  tile_prefix[n_parts] = total (single block)
__global__ void __launch_bounds__(1024)
k_tile_prefix(const unsigned long long* __restrict__ counts, uint32_t n_parts, uint32_t tile, uint32_t* __restrict__ tile_prefix) {
  __shared__ uint32_t sm[33];
  uint32_t carry = 0;
  for (uint32_t base = 0; base < n_parts; base += 1024) {
    const uint32_t p = base + threadIdx.x;
    const uint32_t v = p < n_parts ? (uint32_t)((counts[p] + tile - 1) / tile) : 0u;
    uint32_t tot;
    const uint32_t ex = block_exscan(v, sm, &tot);
    if (p < n_parts) tile_prefix[p] = carry + ex;
    carry += tot;
  }
  if (threadIdx.x == 0) tile_prefix[n_parts] = carry;
}
// grid = n_parts blocks
__global__ void k_make_tilemap(const unsigned long long* __restrict__ part_start, const unsigned long long* __restrict__ counts,
                               const uint32_t* __restrict__ tile_prefix, uint32_t tile, uint2* __restrict__ tilemap) {
  const uint32_t p = blockIdx.x;
  const unsigned long long cnt = counts[p], st = part_start[p];
  const uint32_t nt = (uint32_t)((cnt + tile - 1) / tile), t0 = tile_prefix[p];
  for (uint32_t t = threadIdx.x; t < nt; t += blockDim.x) {
    const unsigned long long off = (unsigned long long)t * tile;
    const unsigned long long left = cnt - off;
    tilemap[t0 + t] = make_uint2((uint32_t)(st + off), (uint32_t)(left < tile ? left : tile));
  }
}

}  // namespace hj3d
