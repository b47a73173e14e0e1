// build_nested_fine.cuh -- nested ("3D") table build, one bucket range per block, completely in shared memory.
//
// HtNested1::insert (ht_nested.hh:287-311) keeps one MainNode per distinct key of a bucket and hangs the
// duplicates below it.  On the device (build.cuh header) that is
//     goff[D+1]  first group of every bucket,   groups[G] {key, first_row, start, len},   rows[n] grouped by key.
// Input: the build records of fine partition f (bucket range [f*width, (f+1)*width)), contiguous in recs.
// The block
//   1. histograms the records over its buckets (one shared atomic per record = rank inside the bucket),
//      scans the histogram and places the records bucket by bucket in shared memory,
//   2. ranks every record inside its bucket by (key, row id) -- one thread per RECORD scanning its bucket in shared
//      memory (buckets are a handful of records; a serial per-bucket sort leaves 255 threads idle behind the longest
//      bucket) -- so that every key group becomes a run whose first record is the MainNode's tuple (smallest row id
//      = first inserted, ht_nested.hh:386-396); the group leaders count the distinct keys of the bucket,
//   3. scans the distinct-key counts (the bucket statistics of makeStatistics, ht_nested.hh:450-482, are reduced
//      on the way) and obtains its first global group index with a decoupled look-back over the partitions
//      before it (single pass, no second kernel, groups stay dense and in bucket order),
//   4. emits the Group records, every main chain in first-appearance order (rank of the leader's row id among the
//      bucket's leaders = the order findMainNodeByOther walks, ht_nested.hh:354-382), the directory words and the
//      row ids.
// A partition with more than cap_recs records sets *overflow (skewed keys); the caller then builds with the
// global-memory kernels of build.cuh.
#pragma once

#include "build.cuh"
#include "common.cuh"
#include "scan.cuh"

namespace hj3d {

constexpr int kNfThreads = 256;   // 96 registers x 256 threads: two or more blocks per SM (512 threads fit only one)
constexpr int kNfItems   = 12;

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
  return *reinterpret_cast<const volatile unsigned long long*>(p);
}

template <class KeyT>
__device__ __forceinline__ bool rec_less(const Slot<KeyT>& a, const Slot<KeyT>& b) {
  return a.key < b.key || (a.key == b.key && a.rowid < b.rowid);
}

template <int HASH>
__global__ void __launch_bounds__(kNfThreads, 4)
k_build_fine_nested(const Slot<typename HashT<HASH>::key_t>* __restrict__ recs,
                    const unsigned long long* __restrict__ part_start, const unsigned long long* __restrict__ counts,
                    const unsigned long long* __restrict__ base, Dir d, uint32_t width, uint32_t n_fine, uint32_t cap_recs,
                    uint32_t* __restrict__ goff, Group<typename HashT<HASH>::key_t>* __restrict__ groups, uint32_t* __restrict__ rows,
                    unsigned long long* lookback /* [n_fine], zeroed */, DevStats* stats, uint32_t* overflow,
                    unsigned long long* g_total) {
  using KeyT = typename HashT<HASH>::key_t;
  using SlotT = Slot<KeyT>;
  using GroupT = Group<KeyT>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* sm_cnt = reinterpret_cast<uint32_t*>(smem_raw);                  // [width + 1] rows per bucket -> exclusive offsets
  uint32_t* sm_dk  = sm_cnt + (width + 1);                                   // [width + 1] distinct keys per bucket -> exclusive offsets
  SlotT*    srec   = reinterpret_cast<SlotT*>(smem_raw + ((2 * (width + 1) * 4 + 15) & ~15u));
  unsigned char* sm_flag = reinterpret_cast<unsigned char*>(srec + cap_recs);   // [cap_recs] 1 = group leader
  __shared__ uint32_t sm_scan[33];
  __shared__ unsigned long long sm_red[160];
  __shared__ unsigned long long sm_gbase;

  const uint32_t f = blockIdx.x;
  const uint32_t blo = f * width;
  const uint32_t bhi = (blo + width < d.n_local) ? blo + width : d.n_local;
  const uint32_t nbk = bhi - blo;
  const unsigned long long cnt64 = counts[f];
  const bool too_big = cnt64 > cap_recs;
  if (too_big && threadIdx.x == 0) atomicExch(overflow, 1u);
  const uint32_t cnt = too_big ? 0u : (uint32_t)cnt64;                       // an overflowing partition takes part as an empty one
  const uint32_t rbase = (uint32_t)base[f];
  const SlotT* in = recs + part_start[f];
  for (uint32_t b = threadIdx.x; b <= nbk; b += kNfThreads) { sm_cnt[b] = 0; sm_dk[b] = 0; }
  // ---- 1. records -> registers, histogram (rank inside the bucket), scan, place
  KeyT     key[kNfItems];
  uint32_t rid[kNfItems], br[kNfItems];
#pragma unroll
  for (int j = 0; j < kNfItems; ++j) {
    const uint32_t li = j * kNfThreads + threadIdx.x;
    key[j] = 0; rid[j] = 0;
    if (li < cnt) { const SlotT r = in[li]; key[j] = r.key; rid[j] = r.rowid; }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kNfItems; ++j) {
    const uint32_t li = j * kNfThreads + threadIdx.x;
    br[j] = 0xFFFFFFFFu;
    if (li < cnt) {
      const uint32_t b = HashT<HASH>::bucket(key[j], d) - d.lo - blo;
      br[j] = (b << 14) | atomicAdd(&sm_cnt[b], 1u);
    }
  }
  __syncthreads();
  constexpr uint32_t PER = 8;                                                // width <= 2048 = 256 threads x 8
  {
    const uint32_t a = threadIdx.x * PER;
    uint32_t v[PER], sum = 0;
#pragma unroll
    for (uint32_t k = 0; k < PER; ++k) { v[k] = (a + k < nbk) ? sm_cnt[a + k] : 0u; sum += v[k]; }
    uint32_t tot;
    uint32_t ex = block_exscan(sum, sm_scan, &tot);
#pragma unroll
    for (uint32_t k = 0; k < PER; ++k) { if (a + k < nbk) sm_cnt[a + k] = ex; ex += v[k]; }
    if (threadIdx.x == 0) sm_cnt[nbk] = cnt;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kNfItems; ++j) {
    if (br[j] == 0xFFFFFFFFu) continue;
    SlotT r; r.key = key[j]; r.rowid = rid[j];
    srec[sm_cnt[br[j] >> 14] + (br[j] & 0x3FFFu)] = r;
  }
  __syncthreads();
  // ---- 2. rank inside the bucket by (key, row id); leaders (smallest row id of their key) count the distinct keys
  // info = leader << 31 | local bucket << 14 | rank            (records of thread t: t, t + kNfThreads, ...)
  uint32_t info[kNfItems];
#pragma unroll
  for (int j = 0; j < kNfItems; ++j) {
    const uint32_t i = j * kNfThreads + threadIdx.x;
    info[j] = 0;
    if (i >= cnt) continue;
    const SlotT me = srec[i];
    const uint32_t b = HashT<HASH>::bucket(me.key, d) - d.lo - blo;
    const uint32_t lo = sm_cnt[b], hi = sm_cnt[b + 1];
    uint32_t less = 0, same_before = 0;
    for (uint32_t q = lo; q < hi; ++q) {
      const SlotT o = srec[q];
      less += (o.key < me.key) ? 1u : 0u;
      same_before += (o.key == me.key && o.rowid < me.rowid) ? 1u : 0u;
    }
    const uint32_t leader = same_before == 0 ? 1u : 0u;
    if (leader) atomicAdd(&sm_dk[b], 1u);
    sm_flag[i] = (unsigned char)leader;
    info[j] = (leader << 31) | (b << 14) | (less + same_before);
  }
  __syncthreads();
  // ---- 3. statistics over the main chain lengths + exclusive scan, first global group index by look-back
  uint32_t g_here;
  {
    DevAgg all{~0ull, 0, 0, 0, 0}, ne{~0ull, 0, 0, 0, 0};
    unsigned long long empty = 0;
    const uint32_t a = threadIdx.x * PER;
    uint32_t v[PER], sum = 0;
#pragma unroll
    for (uint32_t k = 0; k < PER; ++k) {
      v[k] = 0;
      if (a + k < nbk) {
        v[k] = sm_dk[a + k];
        stats_step(all, v[k]);
        if (v[k]) stats_step(ne, v[k]); else ++empty;
      }
      sum += v[k];
    }
    uint32_t ex = block_exscan(sum, sm_scan, &g_here);
#pragma unroll
    for (uint32_t k = 0; k < PER; ++k) { if (a + k < nbk) sm_dk[a + k] = ex; ex += v[k]; }
    if (threadIdx.x == 0) sm_dk[nbk] = g_here;
    DevStats* my_stats = stats + (blockIdx.x & 63u);                   // replica (engine.cu: kStatsCopies = 64)
    agg_commit(all, &my_stats->all, sm_red);
    agg_commit(ne, &my_stats->nonempty, sm_red);
    empty = warp_sum(empty);
    if (lane_id() == 0 && empty) atomicAdd(&my_stats->empty, empty);
  }
  if (threadIdx.x < 32) {
    // decoupled look-back by warp 0, 32 predecessors per step: word = flag << 62 | value; flag 1 = this partition's
    // group count, 2 = inclusive prefix.  The polling loop is warp uniform (all lanes re-read until every word of the
    // window is published), so the warp stays converged for the ballots / shuffles behind it.
    const unsigned long long kMask = (1ull << 62) - 1;
    const uint32_t lane = threadIdx.x;
    unsigned long long prefix = 0;
    if (f > 0) {
      if (lane == 0) atomicExch(lookback + f, (1ull << 62) | (unsigned long long)g_here);
      for (long long p = (long long)f - 1;; p -= 32) {
        const long long idx = p - (long long)lane;
        unsigned long long v = 2ull << 62;                                    // before the first partition: inclusive prefix 0
        do { if (idx >= 0) v = ld_volatile_u64(lookback + idx); } while (__any_sync(0xffffffffu, (v >> 62) == 0));
        const uint32_t incl = __ballot_sync(0xffffffffu, (v >> 62) == 2);
        const uint32_t first = incl ? (uint32_t)(__ffs(incl) - 1) : 31u;      // nearest predecessor that holds an inclusive prefix
        prefix += warp_sum(lane <= first ? (v & kMask) : 0ull);
        if (incl) break;
      }
    }
    if (lane == 0) {
      __threadfence();
      atomicExch(lookback + f, (2ull << 62) | (prefix + g_here));
      sm_gbase = prefix;
      if (f + 1 == n_fine) *g_total = prefix + g_here;
    }
  }
  __syncthreads();
  const uint32_t gbase = (uint32_t)sm_gbase;
  // ---- 4. row ids into (key, row id) order; leaders emit their Group at the rank of their row id among the leaders
#pragma unroll
  for (int j = 0; j < kNfItems; ++j) {
    const uint32_t i = j * kNfThreads + threadIdx.x;
    if (i >= cnt) continue;
    const SlotT me = srec[i];
    const uint32_t b = (info[j] >> 14) & 0x1FFFFu;
    const uint32_t lo = sm_cnt[b], hi = sm_cnt[b + 1];
    const uint32_t pos = rbase + lo + (info[j] & 0x3FFFu);
    rows[pos] = me.rowid;
    if (info[j] >> 31) {
      uint32_t gi = 0, len = 0;                                              // leaders of this bucket inserted before me; my group's length
      for (uint32_t q = lo; q < hi; ++q) {
        const SlotT o = srec[q];
        gi += (sm_flag[q] && o.rowid < me.rowid) ? 1u : 0u;
        len += o.key == me.key ? 1u : 0u;
      }
      GroupT g; g.key = me.key; g.first_row = me.rowid; g.start = pos; g.len = len;
      groups[gbase + sm_dk[b] + gi] = g;
    }
  }
  for (uint32_t b = threadIdx.x; b < nbk; b += kNfThreads) goff[blo + b] = gbase + sm_dk[b];
  if (bhi == d.n_local && threadIdx.x == 0) goff[d.n_local] = gbase + g_here;
}

}  // namespace hj3d
