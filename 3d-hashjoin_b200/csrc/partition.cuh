// partition.cuh -- bucket-range partitioning of a row-store relation into (key, row id) records.
//
// owner(t) = bucket(t) / width with width = ceil(D / n_parts): contiguous bucket ranges, so a
// whole bucket (chain / key group) always falls into one partition.  Used
//   - across GPUs (hj3d_partition_by_owner): the records of partition g are sent to GPU g, which
//     owns directory range [g*width, (g+1)*width)   (SURVEY.md 8(e));
//   - inside one GPU: bucket-ordering both inputs makes the build/probe kernels touch a moving,
//     L2-resident window of the directory instead of random HBM sectors.
//
// Two streaming passes: per-partition histogram (shared-memory privatised), then a scatter that
// reserves one contiguous range per (block, partition) with a single atomic, so writes of a
// block to one partition are contiguous.
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
  __device__ __forceinline__ uint32_t operator()(uint32_t bucket) const {
    return is_pow2 ? (bucket >> shift) : (bucket / width);
  }
};

inline PartFn make_partfn(uint32_t width) {
  PartFn f; f.width = width; f.is_pow2 = (width & (width - 1)) == 0; f.shift = 0;
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
    if (i < s.n) atomicAdd(&h[pf(HashT<HASH>::bucket(src_key<KeyT>(s, i), d))], 1u);
  }
  __syncthreads();
  for (uint32_t p = threadIdx.x; p < n_parts; p += kPartThreads)
    if (h[p]) atomicAdd(&counts[p], (unsigned long long)h[p]);
}

// counts[0..P) -> cursor[P..2P) = exclusive prefix
__global__ void k_part_offsets(unsigned long long* counts, uint32_t n_parts) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    unsigned long long run = 0;
    for (uint32_t p = 0; p < n_parts; ++p) { counts[n_parts + p] = run; run += counts[p]; }
  }
}

template <int HASH>
__global__ void __launch_bounds__(kPartThreads)
k_part_scatter(Src s, Dir d, PartFn pf, uint32_t n_parts, uint32_t rowid_base,
               unsigned long long* __restrict__ cursor, Slot<typename HashT<HASH>::key_t>* __restrict__ out) {
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
      part[j] = pf(HashT<HASH>::bucket(key[j], d));
      rank[j] = atomicAdd(&h[part[j]], 1u);
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
    Slot<KeyT> r; r.key = key[j]; r.rowid = src_rowid(s, i) + rowid_base;
    out[basepos[part[j]] + rank[j]] = r;
  }
}

template <int HASH>
int partition_by_owner_impl(cudaStream_t stream, Src src, Dir d, uint32_t width, uint32_t n_parts, uint32_t rowid_base,
                            void* d_out, unsigned long long* d_counts /* 2 * n_parts, zeroed */, uint64_t* launches) {
  using KeyT = typename HashT<HASH>::key_t;
  if (n_parts > kMaxParts) return HJ3D_ERR_INVALID;
  const PartFn pf = make_partfn(width);
  const uint32_t nb = (uint32_t)((src.n + kPartTile - 1) / kPartTile);
  if (nb) k_part_hist<HASH><<<nb, kPartThreads, 0, stream>>>(src, d, pf, n_parts, d_counts);
  k_part_offsets<<<1, 32, 0, stream>>>(d_counts, n_parts);
  if (nb) k_part_scatter<HASH><<<nb, kPartThreads, 0, stream>>>(src, d, pf, n_parts, rowid_base, d_counts + n_parts,
                                                                 (Slot<KeyT>*)d_out);
  *launches += nb ? 3 : 1;
  return cudaGetLastError() == cudaSuccess ? HJ3D_OK : HJ3D_ERR_CUDA;
}

}  // namespace hj3d
