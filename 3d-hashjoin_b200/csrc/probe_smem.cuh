// probe_smem.cuh -- probe kernels whose random accesses hit SHARED MEMORY instead of L2 / HBM.
//
// Measured on B200 (profiles/): a probe that looks its bucket up in global memory is bound by the
// REQUEST rate of the memory system, not by bytes: ~50 G random sectors/s from HBM, ~100 G random
// requests/s from L2 -- 2^30 probes x 2 lookups = 21 ms even when the directory is L2 resident.
// Shared memory serves the same lookups at several lanes per clock per SM.
//
// So the probe input is bucket-range partitioned (partition.cuh, one or two levels) into FINE
// partitions whose slice of the table -- directory words off[blo..bhi] plus the (key,row id) slots
// (chaining) or the group records (nested) of those buckets -- fits in shared memory.  One block
// handles one work item = (fine partition, chunk of its probe records): it copies the slice into shared
// memory with 128-bit loads, then streams the chunk (coalesced (key,id) records), probes in shared
// memory and writes result pairs through block-aggregated allocation.  The per-tile logic (and hence
// every counter) is the one of probe.cuh.
//
// Work items whose table slice does not fit (a hot key owning millions of build rows) run the same
// tile code against the global arrays.
#pragma once

#include "bulk_copy.cuh"
#include "common.cuh"
#include "probe.cuh"

namespace hj3d {

constexpr int kSmProbeItems   = 4;

struct FineCfg {
  uint32_t width;       // buckets per fine partition
  uint32_t n_local;     // buckets of the (shard) directory
  uint32_t smem_bytes;  // dynamic shared memory available for the slice
};

// copy `bytes` (multiple of 4, both sides 4-byte aligned) global -> shared with the widest aligned loads
__device__ __forceinline__ void copy_to_smem(void* dst, const void* src, uint32_t bytes) {
  const unsigned char* s = (const unsigned char*)src;
  unsigned char* d = (unsigned char*)dst;
  if ((((uintptr_t)src) & 15) == 0 && (((uintptr_t)dst) & 15) == 0) {
    const uint32_t n16 = bytes >> 4;
    for (uint32_t i = threadIdx.x; i < n16; i += blockDim.x)
      reinterpret_cast<uint4*>(d)[i] = __ldg(reinterpret_cast<const uint4*>(s) + i);
    for (uint32_t i = (n16 << 4) + threadIdx.x * 4; i < bytes; i += blockDim.x * 4)
      *reinterpret_cast<uint32_t*>(d + i) = __ldg(reinterpret_cast<const uint32_t*>(s + i));
  } else if ((((uintptr_t)src) & 7) == 0 && (((uintptr_t)dst) & 7) == 0) {
    const uint32_t n8 = bytes >> 3;
    for (uint32_t i = threadIdx.x; i < n8; i += blockDim.x)
      reinterpret_cast<uint2*>(d)[i] = __ldg(reinterpret_cast<const uint2*>(s) + i);
    for (uint32_t i = (n8 << 3) + threadIdx.x * 4; i < bytes; i += blockDim.x * 4)
      *reinterpret_cast<uint32_t*>(d + i) = __ldg(reinterpret_cast<const uint32_t*>(s + i));
  } else {
    for (uint32_t i = threadIdx.x * 4; i < bytes; i += blockDim.x * 4)
      *reinterpret_cast<uint32_t*>(d + i) = __ldg(reinterpret_cast<const uint32_t*>(s + i));
  }
}

// Stage a fine partition's slice -- its directory words and its rows -- in shared memory.  When the four addresses are
// 16-byte aligned both ranges arrive by bulk asynchronous copies (TMA engine, mbarrier completion; sizes rounded up to 16:
// the engine pads its arrays) issued by ONE thread; otherwise by the vector-load loop.  Called by all threads of the block;
// returns with the data visible to all of them.
__device__ __forceinline__ void stage_slice(void* dst_a, const void* src_a, uint32_t bytes_a, void* dst_b, const void* src_b,
                                            uint32_t bytes_b, unsigned long long* bar) {
  const uint32_t ra = (bytes_a + 15u) & ~15u, rb = (bytes_b + 15u) & ~15u;
#ifdef HJ3D_NO_BULK
  const bool bulk = false;
#else
  const bool bulk = ((((uintptr_t)src_a) | ((uintptr_t)src_b) | ((uintptr_t)dst_a) | ((uintptr_t)dst_b)) & 15u) == 0;   // block uniform
#endif
  if (bulk) {
    if (threadIdx.x == 0) mbar_init(bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
      mbar_expect_tx(bar, ra + rb);
      if (ra) bulk_g2s(dst_a, src_a, ra, bar);
      if (rb) bulk_g2s(dst_b, src_b, rb, bar);
    }
    mbar_wait(bar, 0);
  } else {
    copy_to_smem(dst_a, src_a, bytes_a);
    copy_to_smem(dst_b, src_b, bytes_b);
    __syncthreads();
  }
}

// ---- chaining ---------------------------------------------------------------------------------------
template <int HASH, bool UNIQUE, bool CHECKSUM, bool WRITE, bool RECS, int THREADS>
__global__ void __launch_bounds__(THREADS)
k_probe_chaining_smem(Src s, Dir d, FineCfg fc, const uint2* __restrict__ work, const uint32_t* __restrict__ work_part,
                      const uint32_t* __restrict__ off, const Slot<typename HashT<HASH>::key_t>* __restrict__ slots,
                      uint2* __restrict__ out, unsigned long long out_cap, DevCounters* ctr) {
  using KeyT = typename HashT<HASH>::key_t;
  using SlotT = Slot<KeyT>;
  constexpr int kSmProbeThreads = THREADS, kSmProbeTile = THREADS * kSmProbeItems;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ unsigned long long sm_scan[33];
  __shared__ unsigned long long sm_base;

  const uint2 w = work[blockIdx.x];
  const uint32_t f = work_part[blockIdx.x];
  const uint32_t blo = f * fc.width;
  const uint32_t bhi = (blo + fc.width < fc.n_local) ? blo + fc.width : fc.n_local;
  const uint32_t nbk = bhi - blo;
  const uint32_t slo = off[blo], shi = off[bhi];
  const uint32_t nrows = shi - slo;
  // the slot array is copied from its 16-byte aligned predecessor on, so the copy uses 128-bit loads
  const uint32_t pre = slo & (uint32_t)(16 / sizeof(SlotT) - 1);
  const uint32_t off_bytes = ((nbk + 1) * 4 + 15) & ~15u;
  const bool fits = (uint64_t)off_bytes + (uint64_t)(nrows + pre) * sizeof(SlotT) <= fc.smem_bytes;
  uint32_t* sm_off = reinterpret_cast<uint32_t*>(smem_raw);
  SlotT*    sm_slots = reinterpret_cast<SlotT*>(smem_raw + off_bytes);
  if (fits) {
    copy_to_smem(sm_off, off + blo, (nbk + 1) * 4);
    copy_to_smem(sm_slots, slots + (slo - pre), (nrows + pre) * (uint32_t)sizeof(SlotT));
  }
  __syncthreads();
  ProbeAcc acc;
  for (uint32_t sub = 0; sub < w.y; sub += kSmProbeTile) {
    const uint64_t t0 = (uint64_t)w.x + sub;
    const uint32_t tn = (w.y - sub) < (uint32_t)kSmProbeTile ? (w.y - sub) : (uint32_t)kSmProbeTile;
    if (fits)
      probe_chaining_tile<KeyT, UNIQUE, CHECKSUM, WRITE, RECS, kSmProbeThreads, kSmProbeItems, HASH>(
          s, d, t0, tn, d.lo + blo, nbk, sm_off, slo - pre, sm_slots, out, out_cap, ctr, acc, sm_scan, &sm_base);
    else
      probe_chaining_tile<KeyT, UNIQUE, CHECKSUM, WRITE, RECS, kSmProbeThreads, kSmProbeItems, HASH>(
          s, d, t0, tn, d.lo + blo, nbk, off + blo, 0u, slots, out, out_cap, ctr, acc, sm_scan, &sm_base);
    if (WRITE) __syncthreads();                       // sm_base is rewritten by the next sub tile
  }
  commit_acc(acc, ctr, CHECKSUM);
}

// ---- nested -----------------------------------------------------------------------------------------
template <int HASH, bool CHECKSUM, bool WRITE, bool RECS, int THREADS>
__global__ void __launch_bounds__(THREADS)
k_probe_nested_smem(Src s, Dir d, FineCfg fc, const uint2* __restrict__ work, const uint32_t* __restrict__ work_part,
                    const uint32_t* __restrict__ goff, const Group<typename HashT<HASH>::key_t>* __restrict__ groups,
                    uint2* __restrict__ out, unsigned long long out_cap, DevCounters* ctr) {
  using KeyT = typename HashT<HASH>::key_t;
  using GroupT = Group<KeyT>;
  constexpr int kSmProbeThreads = THREADS, kSmProbeTile = THREADS * kSmProbeItems;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ unsigned long long sm_scan[33];
  __shared__ unsigned long long sm_base;

  const uint2 w = work[blockIdx.x];
  const uint32_t f = work_part[blockIdx.x];
  const uint32_t blo = f * fc.width;
  const uint32_t bhi = (blo + fc.width < fc.n_local) ? blo + fc.width : fc.n_local;
  const uint32_t nbk = bhi - blo;
  const uint32_t glo = goff[blo], ghi = goff[bhi];
  const uint32_t ngr = ghi - glo;
  const uint32_t off_bytes = ((nbk + 1) * 4 + 15) & ~15u;
  const bool fits = (uint64_t)off_bytes + (uint64_t)ngr * sizeof(GroupT) <= fc.smem_bytes;
  uint32_t* sm_off = reinterpret_cast<uint32_t*>(smem_raw);
  GroupT*   sm_groups = reinterpret_cast<GroupT*>(smem_raw + off_bytes);
  if (fits) {
    copy_to_smem(sm_off, goff + blo, (nbk + 1) * 4);
    copy_to_smem(sm_groups, groups + glo, ngr * (uint32_t)sizeof(GroupT));
  }
  __syncthreads();
  ProbeAcc acc;
  for (uint32_t sub = 0; sub < w.y; sub += kSmProbeTile) {
    const uint64_t t0 = (uint64_t)w.x + sub;
    const uint32_t tn = (w.y - sub) < (uint32_t)kSmProbeTile ? (w.y - sub) : (uint32_t)kSmProbeTile;
    if (fits)
      probe_nested_tile<KeyT, CHECKSUM, WRITE, RECS, kSmProbeThreads, kSmProbeItems, HASH>(
          s, d, t0, tn, d.lo + blo, nbk, sm_off, glo, sm_groups, out, out_cap, ctr, acc, sm_scan, &sm_base);
    else
      probe_nested_tile<KeyT, CHECKSUM, WRITE, RECS, kSmProbeThreads, kSmProbeItems, HASH>(
          s, d, t0, tn, d.lo + blo, nbk, goff + blo, 0u, groups, out, out_cap, ctr, acc, sm_scan, &sm_base);
    if (WRITE) __syncthreads();
  }
  commit_acc(acc, ctr, CHECKSUM);
}

}  // namespace hj3d
