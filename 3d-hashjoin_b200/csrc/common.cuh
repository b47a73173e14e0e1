// common.cuh -- shared device/host helpers of the hj3d engine (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/hj3d.h"

namespace hj3d {

constexpr uint32_t kEmpty32 = 0xFFFFFFFFu;

// ---------------------------------------------------------------- hashing (util/hasht.hh:52-72)
__host__ __device__ __forceinline__ uint32_t murmur32(uint32_t x) {
  x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16;
  return x;
}
__host__ __device__ __forceinline__ uint64_t murmur64(uint64_t x) {
  x ^= (x >> 33); x *= 0xFF51AFD7ED558CCDull; x ^= (x >> 33); x *= 0xC4CEB9FE1A95EC63ull; x ^= (x >> 33);
  return x;
}
__host__ __device__ __forceinline__ uint64_t pair_mix(uint32_t l, uint32_t r) {
  uint64_t x = ((uint64_t)l << 32) | (uint64_t)r;
  x *= 0x9E3779B97F4A7C15ull;
  x ^= x >> 32;
  return x;
}

// ---------------------------------------------------------------- bucket index: h % numBuckets
// (getDirIndex, ht_chaining.hh:139-140 / ht_nested.hh:221-222).  numBuckets is arbitrary
// (max(#dv / b, 1), main_experiment1.cc:651,875): exact Lemire fastmod for 32-bit hashes, plain
// 64-bit remainder for 64-bit hashes, a mask when numBuckets is a power of two.
struct Dir {
  uint64_t magic;      // ceil(2^64 / D) (0 for D == 1)
  uint32_t D;          // global number of buckets
  uint32_t pow2_mask;  // D - 1 if D is a power of two else 0xFFFFFFFF marker via is_pow2
  uint32_t is_pow2;
  uint32_t lo;         // shard: first owned bucket (local index = bucket - lo)
  uint32_t n_local;    // shard: number of owned buckets
};

inline Dir make_dir(uint64_t D, uint64_t lo, uint64_t hi) {
  Dir d;
  d.D = (uint32_t)D;
  d.is_pow2 = (D & (D - 1)) == 0;
  d.pow2_mask = (uint32_t)(D - 1);
  d.magic = D == 1 ? 0 : (0xFFFFFFFFFFFFFFFFull / D + 1);
  d.lo = (uint32_t)lo;
  d.n_local = (uint32_t)(hi - lo);
  return d;
}

__device__ __forceinline__ uint32_t mod_u32(uint32_t h, const Dir& d) {
  if (d.is_pow2) return h & d.pow2_mask;
  uint64_t lowbits = d.magic * h;
  return (uint32_t)__umul64hi(lowbits, (uint64_t)d.D);
}
__device__ __forceinline__ uint32_t mod_u64(uint64_t h, const Dir& d) {
  if (d.is_pow2) return (uint32_t)h & d.pow2_mask;
  // Barrett: q = floor(h * floor((2^64-1)/D) / 2^64) underestimates floor(h / D) by at most 2 (a hardware 64-bit `%` is a
  // ~100 instruction subroutine and the nested build / probe evaluate the bucket several times per tuple)
  const uint64_t q = __umul64hi(h, d.magic - 1ull);                     // magic = ceil(2^64 / D) = floor((2^64-1)/D) + 1 for D not a power of two
  uint64_t r = h - q * (uint64_t)d.D;
  if (r >= d.D) r -= d.D;
  if (r >= d.D) r -= d.D;
  return (uint32_t)r;
}

// Key/hash traits per hash id.
template <int HASH> struct HashT;
template <> struct HashT<HJ3D_HASH_MURMUR32> {
  using key_t = uint32_t;
  __device__ __forceinline__ static uint32_t bucket(key_t k, const Dir& d) { return mod_u32(murmur32(k), d); }
  __device__ __forceinline__ static uint32_t hash_lo32(key_t k) { return murmur32(k); }
};
template <> struct HashT<HJ3D_HASH_MURMUR64> {
  using key_t = uint64_t;
  __device__ __forceinline__ static uint32_t bucket(key_t k, const Dir& d) { return mod_u64(murmur64(k), d); }
  __device__ __forceinline__ static uint32_t hash_lo32(key_t k) { return (uint32_t)murmur64(k); }
};
template <> struct HashT<HJ3D_HASH_MURMUR64_SEXT32> {
  using key_t = uint32_t;  // raw bits of the int32 attribute; equality on raw bits == equality on ints
  __device__ __forceinline__ static uint32_t bucket(key_t k, const Dir& d) {
    return mod_u64(murmur64((uint64_t)(int64_t)(int32_t)k), d);
  }
  __device__ __forceinline__ static uint32_t hash_lo32(key_t k) { return (uint32_t)murmur64((uint64_t)(int64_t)(int32_t)k); }
};

// secondary mix used to place a key inside its bucket's own slot range (nested grouping)
__device__ __forceinline__ uint32_t mix2(uint32_t k) { return murmur32(k ^ 0x9E3779B9u); }
__device__ __forceinline__ uint32_t mix2(uint64_t k) { return (uint32_t)(murmur64(k ^ 0x9E3779B97F4A7C15ull) >> 32); }

// ---------------------------------------------------------------- row-store source
struct Src {
  const uint8_t*  base;
  const uint32_t* gather;     // nullable
  uint64_t        n;
  uint32_t        stride;
  uint32_t        key_off;
  uint32_t        rowid_off;  // HJ3D_NO_ROWID: row id = position
  // optional selection fused into the load (AlgSelection, algebra.hh:279-315): tuples whose int32 attribute at sel_off does
  // not satisfy `attr <op> sel_cst` are dropped by the partition / exchange pass (sel_op 0 = none)
  uint32_t        sel_off = 0, sel_op = 0;
  int32_t         sel_cst = 0;
};

__device__ __forceinline__ bool src_selected(const Src& s, uint64_t i) {
  if (s.sel_op == 0) return true;
  const uint64_t idx = s.gather ? (uint64_t)__ldg(s.gather + i) : i;
  const int32_t v = __ldg(reinterpret_cast<const int32_t*>(s.base + idx * s.stride + s.sel_off));
  switch (s.sel_op) {
    case 1: return v < s.sel_cst;
    case 2: return v <= s.sel_cst;
    case 3: return v > s.sel_cst;
    case 4: return v >= s.sel_cst;
    case 5: return v == s.sel_cst;
    default: return v != s.sel_cst;
  }
}

// A block's work: records [t0, t0 + tn).  tilemap == nullptr: block b owns tile b of the whole input;
// otherwise the input is bucket-range partitioned with gaps and tilemap[b] = (first record, count).
template <int TILE>
__device__ __forceinline__ void block_tile(const uint2* __restrict__ tilemap, uint64_t n, uint64_t& t0, uint32_t& tn) {
  if (tilemap) { const uint2 e = tilemap[blockIdx.x]; t0 = e.x; tn = e.y; }
  else { t0 = (uint64_t)blockIdx.x * TILE; tn = (n - t0) < (uint64_t)TILE ? (uint32_t)(n - t0) : (uint32_t)TILE; }
}

// id reported as the `left` of a result pair: the tuple's own row id when it carries one (partitioned /
// exchanged (key,row id) records), else its position in the probe sequence.
__device__ __forceinline__ uint32_t src_leftid(const struct Src& s, uint64_t i);

template <class KeyT>
__device__ __forceinline__ KeyT src_key(const Src& s, uint64_t i) {
  uint64_t idx = s.gather ? (uint64_t)__ldg(s.gather + i) : i;
  return __ldg(reinterpret_cast<const KeyT*>(s.base + idx * s.stride + s.key_off));
}
__device__ __forceinline__ uint32_t src_rowid(const Src& s, uint64_t i) {
  if (s.rowid_off == HJ3D_NO_ROWID) return (uint32_t)i;
  uint64_t idx = s.gather ? (uint64_t)__ldg(s.gather + i) : i;
  return __ldg(reinterpret_cast<const uint32_t*>(s.base + idx * s.stride + s.rowid_off));
}

__device__ __forceinline__ uint32_t src_leftid(const Src& s, uint64_t i) {
  if (s.rowid_off == HJ3D_NO_ROWID || s.gather) return (uint32_t)i;
  return __ldg(reinterpret_cast<const uint32_t*>(s.base + i * s.stride + s.rowid_off));
}

// (key, row id) slot of the bucket-ordered build side
template <class KeyT> struct Slot;
template <> struct __align__(8) Slot<uint32_t> { uint32_t key; uint32_t rowid; };
template <> struct __align__(16) Slot<uint64_t> { uint64_t key; uint32_t rowid; uint32_t pad; };

// one directory record per distinct key of the nested table (the MainNode of ht_nested.hh:111-160)
template <class KeyT> struct Group;
template <> struct __align__(16) Group<uint32_t> { uint32_t key; uint32_t first_row; uint32_t start; uint32_t len; };
template <> struct __align__(8)  Group<uint64_t> { uint64_t key; uint32_t first_row; uint32_t start; uint32_t len; uint32_t pad; };

// device-side accumulators -------------------------------------------------------------------
struct DevCounters {
  unsigned long long matches, num_cmps, out_tuples, checksum_sum, checksum_xor, out_cursor, overflow;
};
struct DevAgg {  // Aggregate<size_t>, util/aggregate.hh:27-52
  unsigned long long mn, mx, sum, sumsq, cnt;
};
struct DevStats {
  DevAgg all, nonempty;
  unsigned long long empty;
};

// ---------------------------------------------------------------- warp / block helpers
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

template <class T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <class T> __device__ __forceinline__ T warp_xor(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v ^= __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <class T> __device__ __forceinline__ T warp_min(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { T w = __shfl_xor_sync(0xffffffffu, v, o); v = w < v ? w : v; }
  return v;
}
template <class T> __device__ __forceinline__ T warp_max(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { T w = __shfl_xor_sync(0xffffffffu, v, o); v = w > v ? w : v; }
  return v;
}
// inclusive warp scan
template <class T> __device__ __forceinline__ T warp_iscan(T v) {
  const uint32_t l = lane_id();
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { T w = __shfl_up_sync(0xffffffffu, v, o); if (l >= (uint32_t)o) v += w; }
  return v;
}

// Block-wide exclusive scan of one value per thread (blockDim.x <= 1024); returns the exclusive
// prefix, *total receives the block total.  `smem` needs 33 elements of T.
template <class T>
__device__ __forceinline__ T block_exscan(T v, T* smem, T* total) {
  const uint32_t w = threadIdx.x >> 5, l = lane_id(), nw = (blockDim.x + 31) >> 5;
  T inc = warp_iscan(v);
  if (l == 31) smem[w] = inc;
  __syncthreads();
  if (w == 0) {
    T s = l < nw ? smem[l] : T(0);
    T si = warp_iscan(s);
    smem[l] = si - s;
    if (l == 31) smem[32] = si;
  }
  __syncthreads();
  T res = inc - v + smem[w];
  *total = smem[32];
  __syncthreads();
  return res;
}

}  // namespace hj3d
