// probe_cluster.cuh -- chaining / nested probe of a COARSE bucket-range partition by a thread-block cluster.
//
// Why: the shared-memory probe (probe_smem.cuh) needs the probe input partitioned so finely that one
// partition's table slice fits one SM's shared memory (fan-out ~2^13..2^16 for a 2^27-bucket table); a
// single partition pass cannot write that many streams efficiently (the L2 accepts ~50 requests/clk of any
// size, so 16-byte runs cost as much as 128-byte runs), hence a second streaming pass over the probe
// side (+17 GB of HBM traffic at 2^30 probe tuples).
//
// Here ONE partition pass of fan-out <= 1024 suffices.  A cluster of C CTAs (C SMs) owns one coarse
// partition at a time: CTA r keeps the table slice of sub-range r (1/C of the partition's buckets: 16-bit
// relative directory words + the slots / group records) in its shared memory, so the cluster holds
// C x ~180 KB of table.  The partition's probe records are streamed once from HBM.  Every CTA is warp
// specialised:
//   router warps (8): per round take a tile of the partition, sort it by destination CTA in the CTA's own
//       shared memory (warp ballots + one shared atomic per warp and destination), record the run
//       boundaries and bump the destinations' "round staged" counters;
//   prober warps (24, independent): take (round, source CTA) tickets, PULL the run that source staged for
//       this CTA through distributed shared memory (ld.shared::cluster, coalesced), probe the CTA's slice and
//       write the result pairs with one output reservation per warp and run.
// Hand-off is a pair of monotonic counters per staging buffer bumped remotely (red.shared::cluster) -- no
// cluster-wide barrier in steady state; routers, the prober warps and their in-flight loads overlap each other.  The exchange is exact-fit (runs are as long as they are), so skewed inputs
// stay correct; a slice that does not fit shared memory (hot keys) is probed in global memory by its CTA.
// At most one result per probe record (IsBuildKeyUnique chaining probes and nested probes).
#pragma once

#include <cooperative_groups.h>

#include "common.cuh"
#include "probe.cuh"
#include "probe_smem.cuh"

namespace hj3d {

// ---- cluster / mbarrier PTX -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_map(uint32_t saddr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(cta_rank));
  return r;
}
// Monotonic hand-off counters in shared memory: a remote CTA bumps them with a relaxed cluster-scope reduction, the
// owner polls them with plain volatile loads and a nanosleep back-off (cluster-scope acquire polling costs a
// CCTL.IVALL per poll).  Counters need no phase parity, so prober warps that hold tickets several rounds ahead
// cannot alias the way an mbarrier parity wait would; 8 arrivals per round never wrap 32 bits in practice.
// Shared memory is not cached: what a CTA stored before its CTA barrier is what a peer's ld.shared::cluster returns
// once the peer has seen the bump.
__device__ __forceinline__ void counter_bump_remote(uint32_t cluster_addr) {
  asm volatile("red.relaxed.cluster.shared::cluster.add.u32 [%0], 1;" ::"r"(cluster_addr) : "memory");
}
// Called by ALL lanes of a warp (same address: one broadcast load): a wait loop run by a single lane leaves the warp
// diverged behind it, and every later ballot / shuffle then takes the slow WARPSYNC.COLLECTIVE path (measured: a
// ranking pass of ~600 instructions took 17000 cycles after a lane-0-only spin, 3000 otherwise).
__device__ __forceinline__ void counter_wait(const uint32_t* ctr, uint32_t want) {
  const volatile uint32_t* p = ctr;
  while ((int32_t)(*p - want) < 0) __nanosleep(64);
  __threadfence_block();
}
__device__ __forceinline__ void st_cluster_v2(uint32_t cluster_addr, uint32_t x, uint32_t y) {
  asm volatile("st.shared::cluster.v2.u32 [%0], {%1, %2};" ::"r"(cluster_addr), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ uint2 ld_cluster_v2(uint32_t cluster_addr) {
  uint2 v;
  asm volatile("ld.shared::cluster.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(cluster_addr) : "memory");
  return v;
}
__device__ __forceinline__ uint4 ld_cluster_v4(uint32_t cluster_addr) {
  uint4 v;
  asm volatile("ld.shared::cluster.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(cluster_addr) : "memory");
  return v;
}
template <class KeyT> __device__ __forceinline__ Slot<KeyT> ld_cluster_slot(uint32_t cluster_addr);
template <> __device__ __forceinline__ Slot<uint32_t> ld_cluster_slot<uint32_t>(uint32_t a) {
  const uint2 v = ld_cluster_v2(a); Slot<uint32_t> s; s.key = v.x; s.rowid = v.y; return s;
}
template <> __device__ __forceinline__ Slot<uint64_t> ld_cluster_slot<uint64_t>(uint32_t a) {
  const uint4 v = ld_cluster_v4(a); Slot<uint64_t> s; s.key = ((uint64_t)v.y << 32) | v.x; s.rowid = v.z; s.pad = 0; return s;
}

// ---- geometry ------------------------------------------------------------------------------------------
struct ClusterCfg {
  uint32_t sub_shift;     // log2(buckets per CTA slice)
  uint32_t n_local;       // buckets of the (shard) directory
  uint32_t n_parts;       // coarse partitions
  uint32_t slice_bytes;   // dynamic shared memory available for one slice
};

constexpr int kClC        = 8;                // CTAs per cluster
constexpr int kClProbers  = 768;              // threads   0..767  (24 independent prober warps)
constexpr int kClRouters  = 256;              // threads 768..1023 (8 warps, lock step per round; the highest warp ids: the
                                              // scheduler prefers them, so waiting probers cannot starve the producers)
constexpr int kClThreads  = kClRouters + kClProbers;
constexpr int kClItemsA   = 8;                // records routed per router thread and round
constexpr int kClItemsB   = 10;               // records per prober lane and pass (mean run = TILE / C = 256 = 8 per lane)
template <class KeyT> struct ClTile { static constexpr int kTile = kClRouters * kClItemsA * 8 / (int)sizeof(Slot<KeyT>); };

__device__ __forceinline__ void named_bar(uint32_t id, uint32_t count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// one (partition, round) cursor over the coarse partitions of a cluster
struct RoundIter {
  uint32_t p, j, nr;
  unsigned long long start, cnt;
};
template <int C, int TILE>
__device__ __forceinline__ void iter_load(RoundIter& it, const ClusterCfg& cc, const unsigned long long* __restrict__ part_start,
                                          const unsigned long long* __restrict__ counts, uint32_t stride) {
  while (it.p < cc.n_parts) {
    it.cnt = counts[it.p];
    if (it.cnt) { it.start = part_start[it.p]; it.nr = (uint32_t)((it.cnt + (unsigned long long)C * TILE - 1) / ((unsigned long long)C * TILE)); it.j = 0; return; }
    it.p += stride;
  }
}

// ---- the table slice of one CTA ----------------------------------------------------------------------
template <class RowT>
struct SliceView {
  const uint16_t* off16;     // [nbk + 1] relative to row_lo (a slice holds < 2^16 rows)
  const RowT*     rows;      // shared-memory copy, rows[0] = global row `row_lo`
  const uint32_t* goff;      // global directory (fallback when the slice does not fit)
  const RowT*     grows;
  uint32_t        blo, nbk, row_lo;
  bool            fits;
};

// cooperative load by `nthr` threads (tid = 0..nthr-1)
template <class RowT>
__device__ __forceinline__ void slice_load(SliceView<RowT>& v, unsigned char* smem, uint32_t smem_bytes, const uint32_t* __restrict__ off,
                                           const RowT* __restrict__ rows, uint32_t blo, uint32_t bhi, uint32_t tid, uint32_t nthr) {
  v.goff = off; v.grows = rows; v.blo = blo; v.nbk = bhi - blo;
  const uint32_t rlo = off[blo], rhi = off[bhi];
  const uint32_t nrows = rhi - rlo;
  v.row_lo = rlo;
  const uint32_t pre = (sizeof(RowT) >= 16) ? 0u : (rlo & (uint32_t)(16 / sizeof(RowT) - 1));   // copy from the 16-byte aligned predecessor
  const uint32_t off_bytes = ((v.nbk + 1) * 2 + 15) & ~15u;
  v.fits = nrows < 65536u && (uint64_t)off_bytes + (uint64_t)(nrows + pre) * sizeof(RowT) <= smem_bytes;
  uint16_t* so = reinterpret_cast<uint16_t*>(smem);
  v.off16 = so;
  v.rows = reinterpret_cast<const RowT*>(smem + off_bytes) + pre;
  if (!v.fits) return;
  for (uint32_t b = tid; b <= v.nbk; b += nthr) so[b] = (uint16_t)(__ldg(off + blo + b) - rlo);
  const unsigned char* s = reinterpret_cast<const unsigned char*>(rows + (rlo - pre));
  unsigned char* d = smem + off_bytes;
  const uint32_t bytes = (nrows + pre) * (uint32_t)sizeof(RowT);
  if ((((uintptr_t)s) & 15) == 0) {
    for (uint32_t i = tid; i < (bytes >> 4); i += nthr) reinterpret_cast<uint4*>(d)[i] = __ldg(reinterpret_cast<const uint4*>(s) + i);
    for (uint32_t i = (bytes & ~15u) + tid * 4; i < bytes; i += nthr * 4) *reinterpret_cast<uint32_t*>(d + i) = __ldg(reinterpret_cast<const uint32_t*>(s + i));
  } else {
    for (uint32_t i = tid * 4; i < bytes; i += nthr * 4) *reinterpret_cast<uint32_t*>(d + i) = __ldg(reinterpret_cast<const uint32_t*>(s + i));
  }
}

// ---- per-record lookups: the ordered walks of probe.cuh on one bucket, at most one result ----------------------
// chaining, IsBuildKeyUnique: first match in chain order (algebra.hh:644-657); returns 1 on a hit, `right` = build row
template <class KeyT>
__device__ __forceinline__ uint32_t chain_first(KeyT key, const Slot<KeyT>* sp, uint32_t n, uint32_t& right, uint32_t& cmps) {
  if (n == 0) return 0;                                 // empty bucket: no comparison (algebra.hh:640-643)
  if (n <= kOrderedMax) {
    for (uint32_t k = 0; k < n; ++k) {
      const Slot<KeyT> sl = sp[k];
      if (sl.key == key) { right = sl.rowid; cmps += k + 1; return 1; }
    }
    cmps += n;
    return 0;
  }
  uint32_t min_row = 0xFFFFFFFFu, best = 0; bool any = false, min_is_match = false;
  for (uint32_t k = 0; k < n; ++k) {
    const Slot<KeyT> sl = sp[k];
    const bool hit = sl.key == key;
    if (sl.rowid < min_row) { min_row = sl.rowid; min_is_match = hit; }
    if (hit && (!any || sl.rowid > best)) { best = sl.rowid; any = true; }
  }
  if (!any) { cmps += n; return 0; }
  if (min_is_match) { cmps += 1; right = min_row; return 1; }
  uint32_t rk = 0;
  for (uint32_t k = 0; k < n; ++k) rk += sp[k].rowid < best;
  cmps += n - rk + 1; right = best;
  return 1;
}
// nested: the walk of ht_nested.hh:354-382; `right` = index of the group inside the bucket, `first` = its first row
template <class KeyT>
__device__ __forceinline__ uint32_t group_find(KeyT key, const Group<KeyT>* gp, uint32_t dk, uint32_t& right, uint32_t& first, uint32_t& cmps) {
  if (dk == 0) return 0;                                // empty bucket: {nullptr, 0} (ht_nested.hh:372)
  if (dk <= kOrderedMax) {
    for (uint32_t k = 0; k < dk; ++k) {
      const Group<KeyT> g = gp[k];
      if (g.key == key) { right = k; first = g.first_row; cmps += k + 1; return 1; }
    }
    cmps += dk;
    return 0;
  }
  uint32_t my_first = 0, my_g = 0; bool found = false;
  for (uint32_t k = 0; k < dk && !found; ++k) {
    const Group<KeyT> g = gp[k];
    if (g.key == key) { found = true; my_first = g.first_row; my_g = k; }
  }
  if (!found) { cmps += dk; return 0; }
  uint32_t before = 0;
  for (uint32_t k = 0; k < dk; ++k) before += gp[k].first_row < my_first;
  cmps += before + 1; right = my_g; first = my_first;
  return 1;
}

// one probe record against the CTA's slice; FITS selects the shared-memory copy (LDS) or the global arrays
template <int HASH, int KIND, bool FITS, class RowT>
__device__ __forceinline__ uint32_t probe_record(typename HashT<HASH>::key_t key, const Dir& d, uint32_t sub_mask, const SliceView<RowT>& sv,
                                                 uint32_t& right, uint32_t& first, uint32_t& cmps) {
  using KeyT = typename HashT<HASH>::key_t;
  const uint32_t lb = (HashT<HASH>::bucket(key, d) - d.lo) & sub_mask;
  if (lb >= sv.nbk) return 0;
  uint32_t lo, n; const RowT* bp;
  if (FITS) { const uint32_t o0 = sv.off16[lb]; lo = sv.row_lo + o0; n = (uint32_t)sv.off16[lb + 1] - o0; bp = sv.rows + o0; }
  else      { const uint32_t o0 = __ldg(sv.goff + sv.blo + lb); lo = o0; n = __ldg(sv.goff + sv.blo + lb + 1) - o0; bp = sv.grows + o0; }
  if (KIND == 0) {
    const uint32_t hit = chain_first<KeyT>(key, reinterpret_cast<const Slot<KeyT>*>(bp), n, right, cmps);
    first = right;
    return hit;
  } else {
    uint32_t gi = 0;
    const uint32_t hit = group_find<KeyT>(key, reinterpret_cast<const Group<KeyT>*>(bp), n, gi, first, cmps);
    right = lo + gi;                                     // group ref = global index of the group record
    return hit;
  }
}

// ---- optional in-kernel timeline (compile with -DHJ3D_CL_TRACE): clock64 stamps of CTA 0 ---------------------
#ifdef HJ3D_CL_TRACE
constexpr int kTraceRows = 4096, kTraceCols = 8;
__device__ long long g_cl_trace[kTraceRows][kTraceCols];   // row: (who, k, src, t0..t4)
__device__ unsigned int g_cl_trace_n;
#define HJ_TRACE_DECL long long tr_[6]; (void)tr_
#define HJ_TRACE_AT(i) do { if (blockIdx.x == 0) tr_[i] = clock64(); } while (0)
#define HJ_TRACE_EMIT(who, k, src) do { if (blockIdx.x == 0) { const unsigned r_ = atomicAdd(&g_cl_trace_n, 1u); if (r_ < (unsigned)kTraceRows) { \
    g_cl_trace[r_][0] = (who); g_cl_trace[r_][1] = (k); g_cl_trace[r_][2] = (src); for (int q_ = 0; q_ < 5; ++q_) g_cl_trace[r_][3 + q_] = tr_[q_]; } } } while (0)
#else
#define HJ_TRACE_DECL
#define HJ_TRACE_AT(i)
#define HJ_TRACE_EMIT(who, k, src)
#endif

// ---- the kernel -----------------------------------------------------------------------------------------
// grid = n_clusters * C CTAs of kClThreads threads, cluster dimension C.  Cluster c handles coarse
// partitions c, c + n_clusters, ...; records of partition p: recs[part_start[p] .. + counts[p]).
// KIND 0: chaining probe with IsBuildKeyUnique, result pairs (left, build row);
// KIND 1: nested probe, result pairs (left, group ref).
template <int HASH, int KIND, bool CHECKSUM, bool WRITE>
__global__ void __launch_bounds__(kClThreads, 1)
k_probe_cluster(const Slot<typename HashT<HASH>::key_t>* __restrict__ recs, const unsigned long long* __restrict__ part_start,
                const unsigned long long* __restrict__ counts, Dir d, ClusterCfg cc,
                const uint32_t* __restrict__ off, const void* __restrict__ rows_v,
                uint2* __restrict__ out, unsigned long long out_cap, DevCounters* ctr) {
  using KeyT = typename HashT<HASH>::key_t;
  using SlotT = Slot<KeyT>;
  using RowT = typename std::conditional<KIND == 0, Slot<KeyT>, Group<KeyT>>::type;
  constexpr int C = kClC, LOGC = 3;
  constexpr int TILE = ClTile<KeyT>::kTile;
  constexpr int ITA = TILE / kClRouters;
  constexpr int ITB = kClItemsB;
  namespace cg = cooperative_groups;
  const cg::cluster_group cluster = cg::this_cluster();
  const uint32_t rank = cluster.block_rank();
  const uint32_t n_clusters = gridDim.x / C, cid = blockIdx.x / C;
  const uint32_t lane = lane_id();

  extern __shared__ __align__(16) unsigned char smem_raw[];
  SlotT* staging = reinterpret_cast<SlotT*>(smem_raw);                       // [2][TILE]
  unsigned char* slice_mem = smem_raw + 2 * TILE * sizeof(SlotT);
  // full[k & 3]: round k is staged in every CTA of the cluster.  Four barriers, because prober warps hold tickets up to
  // three rounds ahead: a parity wait on full[k & 3] is alias free as long as round k - 4 has been consumed, which
  // the ticket order guarantees (24 tasks in flight = 3 rounds).  empty[k & 1]: staging[k & 1] of round k was read by all.
  __shared__ uint32_t full_cnt[4], empty_cnt[2];
  __shared__ uint2    runs[2][C];           // [slot][destination] = (first record, count) of the run staged for that CTA
  __shared__ uint32_t cnt[2][C];            // routing histogram of the round being staged
  __shared__ uint32_t ticket;               // next (round, source) task of the prober warps
  static_assert(kClProbers / 32 <= 3 * C, "full[k & 3] needs at most three rounds of tickets in flight");

  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) full_cnt[i] = 0;
    empty_cnt[0] = empty_cnt[1] = 0;
    ticket = 0;
  }
  if (threadIdx.x < 2 * C) cnt[threadIdx.x / C][threadIdx.x % C] = 0;
  cluster.sync();

  const uint32_t sub_mask = (1u << cc.sub_shift) - 1u;
  ProbeAcc acc;

  if (threadIdx.x >= kClProbers) {
    // =========================== routers ===========================
    const uint32_t tid = threadIdx.x - kClProbers;
    RoundIter it;
    it.p = cid; it.j = 0; it.nr = 0; it.cnt = 0; it.start = 0;
    iter_load<C, TILE>(it, cc, part_start, counts, n_clusters);
    KeyT     key[ITA];
    uint32_t id[ITA];
    uint32_t tn = 0;
    auto fetch = [&]() {                     // my records of round (it.p, it.j) -> registers
      tn = 0;
      if (it.p >= cc.n_parts) return;
      const unsigned long long t0 = ((unsigned long long)it.j * C + rank) * TILE;
      if (t0 >= it.cnt) return;
      tn = (it.cnt - t0) < (unsigned long long)TILE ? (uint32_t)(it.cnt - t0) : (uint32_t)TILE;
      const SlotT* in = recs + it.start + t0;
#pragma unroll
      for (int j = 0; j < ITA; ++j) {
        const uint32_t li = j * kClRouters + tid;
        if (li < tn) { const SlotT r = in[li]; key[j] = r.key; id[j] = r.rowid; }
      }
    };
    fetch();
    for (uint32_t k = 0; it.p < cc.n_parts; ++k) {
      const uint32_t slot = k & 1;
      HJ_TRACE_DECL;
      HJ_TRACE_AT(0);
      if (k >= 2) counter_wait(&empty_cnt[slot], (uint32_t)C * (k >> 1));                  // every reader of rounds < k of this slot is done
      HJ_TRACE_AT(1);
      const uint32_t pbase = it.p << (cc.sub_shift + LOGC);                              // first bucket of the coarse partition
      uint32_t dr[ITA];                      // (destination << 16) | rank inside the run, 0xFFFFFFFF = no record
#pragma unroll
      for (int j = 0; j < ITA; ++j) {
        const uint32_t li = j * kClRouters + tid;
        const bool ok = li < tn;
        uint32_t dd = 0;
        if (ok) dd = ((HashT<HASH>::bucket(key[j], d) - d.lo - pbase) >> cc.sub_shift) & (uint32_t)(C - 1);
        uint32_t peers = __ballot_sync(0xffffffffu, ok);                                 // lanes with my destination: LOGC ballots
#pragma unroll
        for (int b = 0; b < LOGC; ++b) {
          const uint32_t m = __ballot_sync(0xffffffffu, (dd >> b) & 1u);
          peers &= ((dd >> b) & 1u) ? m : ~m;
        }
        const uint32_t leader = ok ? (uint32_t)(__ffs(peers) - 1) : 0u;
        uint32_t base = 0;
        if (ok && lane == leader) base = atomicAdd(&cnt[slot][dd], (uint32_t)__popc(peers));
        base = __shfl_sync(0xffffffffu, base, leader);
        dr[j] = ok ? ((dd << 16) | (base + __popc(peers & ((1u << lane) - 1u)))) : 0xFFFFFFFFu;
      }
      HJ_TRACE_AT(2);
      named_bar(1, kClRouters);              // cnt[slot] complete; staging[slot] is free (thread 0 waited)
      HJ_TRACE_AT(3);
      uint32_t c_me = lane < (uint32_t)C ? cnt[slot][lane] : 0u, inc = c_me;              // lanes 0..C-1: exclusive prefix of cnt
#pragma unroll
      for (int o = 1; o < C; o <<= 1) { const uint32_t w = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= (uint32_t)o) inc += w; }
      const uint32_t start_me = inc - c_me;
      if (tid < (uint32_t)C) runs[slot][tid] = make_uint2(start_me, c_me);
      SlotT* st = staging + slot * TILE;
#pragma unroll
      for (int j = 0; j < ITA; ++j) {
        const uint32_t s0 = __shfl_sync(0xffffffffu, start_me, (dr[j] >> 16) & (uint32_t)(C - 1));
        if (dr[j] != 0xFFFFFFFFu) { SlotT r; r.key = key[j]; r.rowid = id[j]; st[s0 + (dr[j] & 0xFFFFu)] = r; }
      }
      named_bar(1, kClRouters);              // staging[slot] and runs[slot] are complete (the barrier drains the stores)
      HJ_TRACE_AT(4);
      if (tid < (uint32_t)C) {               // lane t tells CTA t that its run of round k is staged
        __threadfence_block();
        counter_bump_remote(cluster_map(smem_u32(&full_cnt[k & 3]), tid));
        cnt[slot][tid] = 0;                  // next use is two rounds (four barriers) away
      }
      if (++it.j >= it.nr) { it.p += n_clusters; iter_load<C, TILE>(it, cc, part_start, counts, n_clusters); }
      fetch();
      if (tid == 0) HJ_TRACE_EMIT(0, k, 0);
      if (tid == 7 * 32) HJ_TRACE_EMIT(100, k, 0);
    }
  } else {
    // =========================== probers: 24 independent warps, task = (round, source CTA) ===========================
    const uint32_t ptid = threadIdx.x;
    const RowT* rows = reinterpret_cast<const RowT*>(rows_v);
    SliceView<RowT> sv;
    RoundIter it;                            // partition whose slice is (about to be) loaded
    it.p = cid; it.j = 0; it.nr = 0; it.cnt = 0; it.start = 0;
    iter_load<C, TILE>(it, cc, part_start, counts, n_clusters);
    uint32_t k_end = 0;                      // cluster-wide round numbers [k_end - nr, k_end) belong to the loaded partition
    bool loaded = false;
    for (;;) {
      uint32_t t = 0;
      if (lane == 0) t = atomicAdd(&ticket, 1u);
      t = __shfl_sync(0xffffffffu, t, 0);
      const uint32_t k = t >> LOGC, src = t & (uint32_t)(C - 1);
      // make the slice of round k's partition current (all prober warps walk the same partition sequence)
      while (it.p < cc.n_parts && (!loaded || k >= k_end)) {
        if (loaded) { it.p += n_clusters; iter_load<C, TILE>(it, cc, part_start, counts, n_clusters); if (it.p >= cc.n_parts) break; }
        named_bar(2, kClProbers);                            // every warp is done with the previous slice
        const uint32_t pbase = it.p << (cc.sub_shift + LOGC);
        uint32_t blo = pbase + (rank << cc.sub_shift);
        blo = blo < cc.n_local ? blo : cc.n_local;
        const uint32_t bhi = (blo + sub_mask + 1u) < cc.n_local ? (blo + sub_mask + 1u) : cc.n_local;
        slice_load<RowT>(sv, slice_mem, cc.slice_bytes, off, rows, blo, bhi, ptid, kClProbers);
        named_bar(2, kClProbers);
        k_end += it.nr;
        loaded = true;
      }
      if (it.p >= cc.n_parts) break;                         // no rounds left
      const uint32_t slot = k & 1;
      HJ_TRACE_DECL;
      HJ_TRACE_AT(0);
      counter_wait(&full_cnt[k & 3], (uint32_t)C * ((k >> 2) + 1));                      // all C routers staged round k
      HJ_TRACE_AT(1);
      const uint2 run = ld_cluster_v2(cluster_map(smem_u32(&runs[slot][rank]), src));   // what CTA `src` staged for me
      const uint32_t n_in = run.y;
      const uint32_t src0 = cluster_map(smem_u32(staging + slot * TILE), src) + run.x * (uint32_t)sizeof(SlotT);
      const uint32_t empty_remote = cluster_map(smem_u32(&empty_cnt[slot]), src);
      for (uint32_t base0 = 0; base0 == 0 || base0 < n_in; base0 += 32 * ITB) {
        KeyT     pk[ITB];
        uint32_t left[ITB];
#pragma unroll
        for (int q = 0; q < ITB; ++q) {
          const uint32_t i = base0 + q * 32 + lane;
          pk[q] = 0; left[q] = 0;
          if (i < n_in) { const SlotT r = ld_cluster_slot<KeyT>(src0 + i * (uint32_t)sizeof(SlotT)); pk[q] = r.key; left[q] = r.rowid; }
        }
        uint32_t hitmask = 0, cmps = 0, wtot = 0;
#pragma unroll
        for (int q = 0; q < ITB; ++q) {
          const uint32_t i = base0 + q * 32 + lane;
          uint32_t right = 0, first = 0, hit = 0;
          if (i < n_in) {
            hit = sv.fits ? probe_record<HASH, KIND, true, RowT>(pk[q], d, sub_mask, sv, right, first, cmps)
                          : probe_record<HASH, KIND, false, RowT>(pk[q], d, sub_mask, sv, right, first, cmps);
          }
          if (CHECKSUM && hit) { const uint64_t mx = pair_mix(left[q], first); acc.sum += mx; acc.x ^= mx; }
          hitmask |= hit << q;
          pk[q] = (KeyT)right;                               // the key's register now holds the result
          wtot += __popc(__ballot_sync(0xffffffffu, hit));
        }
        // the last remote read of this run is complete (its values were consumed above): hand the buffer back
        if (base0 + 32 * ITB >= n_in) { __syncwarp(); if (lane == 0) counter_bump_remote(empty_remote); }
        HJ_TRACE_AT(2);
        acc.matches += __popc(hitmask);
        acc.cmps += cmps;
        if (WRITE && wtot) {                 // one output reservation per warp and pass
          unsigned long long pos = 0;
          if (lane == 0) pos = atomicAdd(&ctr->out_cursor, (unsigned long long)wtot);
          pos = __shfl_sync(0xffffffffu, pos, 0);
          HJ_TRACE_AT(3);
#pragma unroll
          for (int q = 0; q < ITB; ++q) {
            const uint32_t hit = (hitmask >> q) & 1u;
            const uint32_t bal = __ballot_sync(0xffffffffu, hit);
            const unsigned long long mypos = pos + __popc(bal & ((1u << lane) - 1u));
            if (hit && mypos < out_cap) out[mypos] = make_uint2(left[q], (uint32_t)pk[q]);
            pos += __popc(bal);
          }
        }
        HJ_TRACE_AT(4);
        if (lane == 0) HJ_TRACE_EMIT(1 + (threadIdx.x >> 5), k, src);
      }
    }
  }
  cluster.sync();                            // nobody leaves while a peer may still read its staging buffers
  commit_acc(acc, ctr, CHECKSUM);
}

}  // namespace hj3d
