// build_nested_fine.cuh -- nested ("3D") table build, one bucket range per block, completely in shared memory.
//
// HtNested1::insert (ht_nested.hh:287-311) keeps one MainNode per distinct key of a bucket and hangs the
// duplicates below it.  On the device (build.cuh header) that is
//     goff[D+1]  first group of every bucket,   groups[G] {key, first_row, start, len},   rows[n] grouped by key.
// Input: the build records of fine partition f (bucket range [f*width, (f+1)*width)), contiguous in recs.
// The block
//   1. histograms the records over its buckets (one shared atomic per record = rank inside the bucket),
//      scans the histogram and places the records bucket by bucket in shared memory,
//   2. groups every bucket by key, warp-cooperatively (round 1 let every RECORD scan its whole bucket: O(bucket length)
//      per record, 37 ms per 2^30 rows at 8 duplicates per key and hopeless at the 100..1000 duplicates of
//      main_experiment4).  A warp walks its share of the buckets; as many consecutive buckets as hold <= 32 records are
//      handled at once, one record per lane: __match_any_sync on the key finds a record's key group, whose smallest row id is
//      the MainNode (first inserted, ht_nested.hh:386-396), and a short shuffle loop over the bucket's
//      records ranks the groups by first appearance.  A bucket of more than 32 records is sorted by (key, row id) with a
//      bitonic network in shared memory by its warp; groups are then runs.  Either way every record learns its slot in
//      rows[] and every leader its group's (index inside the bucket, start, length),
//   3. scans the distinct-key counts (the bucket statistics of makeStatistics, ht_nested.hh:450-482, are reduced
//      on the way),
//   4. emits the Group records, every main chain in first-appearance order (rank of the leader's row id among the
//      bucket's leaders = the order findMainNodeByOther walks, ht_nested.hh:354-382), to a staging array at the
//      partition's record base, and partition-local directory words; k_nested_compact makes both dense / global.
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

template <class KeyT> __device__ __forceinline__ uint32_t match_key(uint32_t mask, KeyT k);
template <> __device__ __forceinline__ uint32_t match_key<uint32_t>(uint32_t mask, uint32_t k) { return __match_any_sync(mask, k); }
template <> __device__ __forceinline__ uint32_t match_key<uint64_t>(uint32_t mask, uint64_t k) { return __match_any_sync(mask, (unsigned long long)k); }

// bytes of dynamic shared memory per record: the record, its group word, its group's start inside the bucket, its bucket
template <class KeyT> constexpr uint32_t nf_bytes_per_record() { return (uint32_t)sizeof(Slot<KeyT>) + 4u + 2u + 2u; }

template <int HASH>
__global__ void __launch_bounds__(kNfThreads, 4)
k_build_fine_nested(const Slot<typename HashT<HASH>::key_t>* __restrict__ recs,
                    const unsigned long long* __restrict__ part_start, const unsigned long long* __restrict__ counts,
                    const unsigned long long* __restrict__ base, Dir d, uint32_t width, uint32_t n_fine, uint32_t cap_recs,
                    uint32_t* __restrict__ goff /* partition-local group indices */, Group<typename HashT<HASH>::key_t>* __restrict__ gtmp,
                    uint32_t* __restrict__ rows, unsigned long long* __restrict__ g_count /* [n_fine] groups of every partition */,
                    DevStats* stats, uint32_t* overflow) {
  using KeyT = typename HashT<HASH>::key_t;
  using SlotT = Slot<KeyT>;
  using GroupT = Group<KeyT>;
  constexpr uint32_t NW = kNfThreads / 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* sm_cnt = reinterpret_cast<uint32_t*>(smem_raw);                  // [width + 1] rows per bucket -> exclusive offsets
  uint32_t* sm_dk  = sm_cnt + (width + 1);                                   // [width + 1] distinct keys per bucket -> exclusive offsets
  SlotT*    srec   = reinterpret_cast<SlotT*>(smem_raw + ((2 * (width + 1) * 4 + 15) & ~15u));
  uint32_t* sm_ginfo = reinterpret_cast<uint32_t*>(srec + cap_recs);         // [cap_recs] leader << 31 | group index in bucket << 16 | group length
  uint16_t* sm_gpos  = reinterpret_cast<uint16_t*>(sm_ginfo + cap_recs);     // [cap_recs] start of the record's group inside its bucket
  uint16_t* sm_bkt   = sm_gpos + cap_recs;                                   // [cap_recs] local bucket of the record
  __shared__ uint32_t sm_scan[33];
  __shared__ unsigned long long sm_red[160];
  __shared__ uint32_t sm_lead_first[NW][kOrderedMaxB], sm_lead_pos[NW][kOrderedMaxB];

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
  {
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
    constexpr uint32_t PER = 8;                                              // width <= 2048 = 256 threads x 8
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
      const uint32_t b = br[j] >> 14, pos = sm_cnt[b] + (br[j] & 0x3FFFu);
      srec[pos] = r;
      sm_bkt[pos] = (uint16_t)b;
    }
  }
  __syncthreads();
  // ---- 2. group every bucket by key, one warp per share of the buckets
  {
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    const uint32_t per = (nbk + NW - 1) / NW;
    uint32_t b = warp * per;
    const uint32_t bend = (b + per < nbk) ? b + per : nbk;
    while (b < bend) {                                                       // warp uniform
      const uint32_t lo = sm_cnt[b];
      const uint32_t cand = b + lane + 1;                                    // buckets [b, cand) fit one warp?
      const bool fits = cand <= bend && (sm_cnt[cand < nbk ? cand : nbk] - lo) <= 32u;
      const uint32_t fm = __ballot_sync(0xffffffffu, fits);
      const uint32_t e = fm == 0xFFFFFFFFu ? 32u : (uint32_t)(__ffs(~fm) - 1);
      if (e != 0) {
        // -- short: buckets [b, b + e) hold n <= 32 records, one per lane
        const uint32_t n = sm_cnt[b + e] - lo;
        const bool act = lane < n;
        const uint32_t i = lo + lane;
        SlotT me; me.key = 0; me.rowid = 0xFFFFFFFFu;
        uint32_t bb = 0, first = 0xFFFFFFFFu, glen = 0, pig = 0, leader = 0, bs = 0, bl = 0;
        const uint32_t am = __ballot_sync(0xffffffffu, act);
        if (act) {
          me = srec[i]; bb = sm_bkt[i];
          const uint32_t mask = match_key<KeyT>(am, me.key);                 // lanes holding my key (equal keys share a bucket)
          // the MainNode's tuple = first inserted = smallest row id of the group.  (__reduce_min_sync over the group mask
          // compiles to one serialised CREDUX per DISTINCT mask: 32 rounds per warp for unique keys, half of the kernel's
          // instructions in the ncu capture; the other members are read from shared memory instead.)
          const uint32_t n_groups = __popc(__ballot_sync(am, (uint32_t)__ffs(mask) - 1u == lane));
          const uint32_t max_glen = __reduce_max_sync(am, (uint32_t)__popc(mask));
          first = me.rowid;
          if (max_glen > 1u) {
            if (n_groups < max_glen) {                                       // few large groups: one CREDUX round per group
              first = __reduce_min_sync(mask, me.rowid);
            } else {                                                         // many small groups: walk my group's other members
              for (uint32_t m = mask & ~(1u << lane); m; m &= m - 1u) {
                const uint32_t o = srec[lo + (uint32_t)__ffs(m) - 1u].rowid;
                first = o < first ? o : first;
              }
            }
          }
          leader = me.rowid == first ? 1u : 0u;
          glen = __popc(mask); pig = __popc(mask & ((1u << lane) - 1u));
          bs = sm_cnt[bb] - lo; bl = sm_cnt[bb + 1] - sm_cnt[bb];
        }
        const uint32_t maxbl = __reduce_max_sync(0xffffffffu, bl);
        const uint32_t lb = __ballot_sync(0xffffffffu, leader);             // the pack's MainNodes
        const uint32_t bm = bl >= 32u ? 0xFFFFFFFFu : (((1u << bl) - 1u) << bs);   // lanes of my bucket
        const uint32_t dk = __popc(lb & bm);                                 // distinct keys of my bucket
        uint32_t gi = 0, gstart = 0;                                         // groups of my bucket that appeared before mine; their rows
        if ((uint32_t)__popc(lb) <= maxbl) {                                 // few groups (duplicates): visit the pack's leaders
          for (uint32_t m = lb; m; m &= m - 1u) {
            const uint32_t src = (uint32_t)__ffs(m) - 1u;
            const uint32_t of = __shfl_sync(0xffffffffu, first, src), on = __shfl_sync(0xffffffffu, glen, src);
            if (((bm >> src) & 1u) && of < first) { ++gi; gstart += on; }
          }
        } else {                                                             // short buckets (unique keys): visit my bucket's records
          for (uint32_t dd = 0; dd < maxbl; ++dd) {
            const uint32_t src = (bs + dd) & 31u;
            const uint32_t of = __shfl_sync(0xffffffffu, first, src), on = __shfl_sync(0xffffffffu, glen, src);
            if (dd < bl && ((lb >> src) & 1u) && of < first) { ++gi; gstart += on; }
          }
        }
        if (act) {
          sm_ginfo[i] = (leader << 31) | (gi << 16) | glen;
          sm_gpos[i] = (uint16_t)gstart;
          rows[rbase + sm_cnt[bb] + gstart + pig] = me.rowid;
          if (lane == bs) sm_dk[bb] = dk;
        }
        b += e;
        continue;
      }
      // -- long: bucket b alone holds L > 32 records: sort by (key, row id), groups are runs
      const uint32_t L = sm_cnt[b + 1] - lo;
      SlotT* r = srec + lo;
      uint32_t P = 64; while (P < L) P <<= 1;
      for (uint32_t k = 2; k <= P; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
          const uint32_t flip = (j == (k >> 1)) ? (2u * j - 1u) : j;         // normalised network: every comparator ascending
          for (uint32_t cmp = lane; cmp < (P >> 1); cmp += 32) {
            const uint32_t lo_i = ((cmp & ~(j - 1u)) << 1) | (cmp & (j - 1u));
            const uint32_t hi_i = lo_i ^ flip;
            if (hi_i < L) {                                                 // elements past L act as +infinity and never move
              const SlotT x = r[lo_i], y = r[hi_i];
              if (rec_less(y, x)) { r[lo_i] = y; r[hi_i] = x; }
            }
          }
          __syncwarp();
        }
      }
      uint32_t n_lead = 0, cur_lead = 0;                                     // leaders so far; position of the run covering the chunk start
      for (uint32_t c0 = 0; c0 < L; c0 += 32) {
        const uint32_t sidx = c0 + lane;
        const bool act = sidx < L;
        SlotT me; me.key = 0; me.rowid = 0;
        bool is_lead = false, is_last = false;
        if (act) {
          me = r[sidx];
          is_lead = sidx == 0 || r[sidx - 1].key != me.key;
          is_last = sidx + 1 == L || r[sidx + 1].key != me.key;
        }
        const uint32_t lb = __ballot_sync(0xffffffffu, is_lead);
        const uint32_t before = lb & ((2u << lane) - 1u);                    // leaders at or before me in this chunk
        const uint32_t my_lead = before ? c0 + (31u - (uint32_t)__clz(before)) : cur_lead;
        const uint32_t lidx = n_lead + __popc(lb & ((1u << lane) - 1u));
        if (act) {
          rows[rbase + lo + sidx] = me.rowid;
          sm_gpos[lo + sidx] = (uint16_t)my_lead;
          if (is_lead) {
            sm_ginfo[lo + sidx] = (1u << 31) | ((lidx & 0x7FFFu) << 16);
            if (lidx < kOrderedMaxB) { sm_lead_first[warp][lidx] = me.rowid; sm_lead_pos[warp][lidx] = sidx; }
          } else {
            sm_ginfo[lo + sidx] = 0;
          }
        }
        __syncwarp();
        if (act && is_last) sm_ginfo[lo + my_lead] |= (sidx - my_lead + 1u) & 0xFFFFu;   // the run's last record knows its length
        __syncwarp();
        if (lb) cur_lead = c0 + (31u - (uint32_t)__clz(lb));
        n_lead += __popc(lb);
      }
      if (n_lead <= kOrderedMaxB) {                                          // short main chain: first-appearance order matters (probe walks it)
        const uint32_t mine = lane < n_lead ? sm_lead_first[warp][lane] : 0xFFFFFFFFu;
        uint32_t rank = 0;
        for (uint32_t o = 0; o < n_lead; ++o) rank += __shfl_sync(0xffffffffu, mine, o) < mine ? 1u : 0u;
        if (lane < n_lead) {
          const uint32_t at = lo + sm_lead_pos[warp][lane];
          sm_ginfo[at] = (sm_ginfo[at] & 0x8000FFFFu) | (rank << 16);
        }
      }
      if (lane == 0) sm_dk[b] = n_lead;
      __syncwarp();
      b += 1;
    }
  }
  __syncthreads();
  // ---- 3. statistics over the main chain lengths + exclusive scan, first global group index by look-back
  uint32_t g_here;
  {
    constexpr uint32_t PER = 8;
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
  if (threadIdx.x == 0) g_count[f] = g_here;
  // (round 1 obtained the partition's first global group index here with a decoupled look-back.  With several hundred
  //  resident blocks every block walked ~14 windows of 32 predecessors at one global-memory round trip each, ~20 us per
  //  block of 2048 records, more than all the rest of the kernel.  Groups now go to a staging array at the partition's
  //  RECORD base (#groups <= #records, so ranges cannot overlap); k_nested_compact moves them to their dense place once a
  //  device scan of g_count has the bases -- the copy replaces the memcpy the staging array needed anyway.)
  const uint32_t gbase = rbase;
  // ---- 4. the leaders emit their Group records: main chains in first-appearance order (findMainNodeByOther walks them)
  for (uint32_t i = threadIdx.x; i < cnt; i += kNfThreads) {
    const uint32_t info = sm_ginfo[i];
    if (!(info >> 31)) continue;
    const SlotT me = srec[i];
    const uint32_t b = sm_bkt[i];
    GroupT g; g.key = me.key; g.first_row = me.rowid; g.start = rbase + sm_cnt[b] + sm_gpos[i]; g.len = info & 0xFFFFu;
    gtmp[gbase + sm_dk[b] + ((info >> 16) & 0x7FFFu)] = g;
  }
  for (uint32_t b = threadIdx.x; b < nbk; b += kNfThreads) goff[blo + b] = sm_dk[b];
}

// second step: partition f's groups gtmp[rec_base[f] .. + g_count[f]) -> groups[g_base[f] ..), directory words made global
template <class KeyT>
__global__ void __launch_bounds__(256)
k_nested_compact(const Group<KeyT>* __restrict__ gtmp, const unsigned long long* __restrict__ rec_base, const unsigned long long* __restrict__ g_base,
                 const unsigned long long* __restrict__ g_count, uint32_t width, uint32_t n_local, uint32_t n_fine,
                 Group<KeyT>* __restrict__ groups, uint32_t* __restrict__ goff) {
  const uint32_t f = blockIdx.x;
  const uint32_t gb = (uint32_t)g_base[f], n = (uint32_t)g_count[f];
  const Group<KeyT>* src = gtmp + rec_base[f];
  for (uint32_t i = threadIdx.x; i < n; i += 256) groups[gb + i] = src[i];
  const uint32_t blo = f * width, bhi = (blo + width < n_local) ? blo + width : n_local;
  for (uint32_t b = blo + threadIdx.x; b < bhi; b += 256) goff[b] += gb;
  if (f + 1 == n_fine && threadIdx.x == 0) goff[n_local] = gb + n;
}

}  // namespace hj3d
