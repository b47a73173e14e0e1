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
#include "hot_set.cuh"

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
  uint32_t nl;         // buckets of the (shard) directory: tuples of other buckets belong to no partition
  __device__ __forceinline__ uint32_t operator()(uint32_t bucket) const {
    const uint32_t b = bucket - lo;
    if (b >= nl) return 0xFFFFFFFFu;       // (the last partition may reach past the directory's end)
    return is_pow2 ? (b >> shift) : (b / width);
  }
};

inline PartFn make_partfn(uint32_t width, uint32_t lo, uint32_t nl) {
  PartFn f; f.width = width; f.is_pow2 = (width & (width - 1)) == 0; f.shift = 0; f.lo = lo; f.nl = nl;
  while ((1u << f.shift) < width) ++f.shift;
  return f;
}

// Lanes of the warp whose partition id (< 8) equals mine, found with three ballots.  With a handful of partitions
// (the owner split across 2..8 GPUs) every lane of a block hits the same few shared-memory counters; aggregating per
// warp first turns 32 conflicting atomics into at most `fan` (measured at 2 owners: 6.7 -> 2.x ms per 2^29 tuples).
__device__ __forceinline__ uint32_t peers_small(uint32_t lp, bool valid) {
  uint32_t peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
  for (int b = 0; b < 3; ++b) {
    const uint32_t m = __ballot_sync(0xffffffffu, (lp >> b) & 1u);
    peers &= ((lp >> b) & 1u) ? m : ~m;
  }
  return peers;
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
    uint32_t p = 0xFFFFFFFFu;
    if (i < s.n && src_selected(s, i)) p = pf(HashT<HASH>::bucket(src_key<KeyT>(s, i), d));
    const bool valid = p < n_parts;
    if (n_parts <= 8) {
      const uint32_t peers = peers_small(valid ? p : 0u, valid);
      if (valid && (uint32_t)(__ffs(peers) - 1) == lane_id()) atomicAdd(&h[p], (uint32_t)__popc(peers));
    } else if (valid) {
      atomicAdd(&h[p], 1u);
    }
  }
  __syncthreads();
  for (uint32_t p = threadIdx.x; p < n_parts; p += kPartThreads)
    if (h[p]) atomicAdd(&counts[p], (unsigned long long)h[p]);
}

// part_start[p] = exclusive prefix of counts (single thread; n_parts <= 1024)
static __global__ void k_part_prefix(const unsigned long long* __restrict__ counts, uint32_t n_parts,
                              unsigned long long* __restrict__ part_start) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    unsigned long long run = 0;
    for (uint32_t p = 0; p < n_parts; ++p) { part_start[p] = run; run += counts[p]; }
  }
}
static __global__ void k_part_fixed_starts(uint32_t n_parts, unsigned long long cap, unsigned long long* __restrict__ part_start) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n_parts) part_start[p] = (unsigned long long)p * cap;
}

// Scatter one tile into per-partition regions: records of partition q go to out[part_start[q] + k],
// k = running cursor[q]; k >= cap is dropped (the cursor keeps counting, so it ends up holding the exact
// partition size either way).
//
// The tile is first sorted by partition in shared memory (rank by shared-memory atomics, exclusive scan of
// the tile histogram), then written out in sorted order, so that consecutive threads store consecutive
// records of one partition: a warp issues a handful of wide L2 write requests instead of 32 scattered
// 8-byte ones.  (Measured on B200: the L2 services ~100 G random requests/s, so one request per record
// caps a partition pass at ~10 ms per 2^30 records no matter how few bytes it moves.)
//
// q = bucket / width is the partition id; a block only ever sees partitions [q0, q0 + fan) with
// q0 = fan * (q / fan) of its first record (level 1: q0 = 0; level 2: all records of a tile come from one
// coarse partition = fan consecutive fine partitions).
// LEFTID: the stored id is the probe-side "left" id (position, or the tuple's own id) instead of the
// build-side row id + rowid_base.  Records of buckets outside the directory (q >= n_parts) are dropped.
// Tile shape: THREADS threads x (64 / sizeof(Slot)) records per thread * 2, i.e. 16 records per thread for
// 8-byte slots, 8 for 16-byte slots.
#ifndef HJ3D_PART_TILE_BYTES
#define HJ3D_PART_TILE_BYTES 128
#endif
#ifndef HJ3D_PART_MINBLOCKS
#define HJ3D_PART_MINBLOCKS 2
#endif
template <class KeyT> struct PartCfg { static constexpr int kItems = HJ3D_PART_TILE_BYTES / (int)sizeof(Slot<KeyT>); };

// dynamic shared memory: sorted tile + partition id per record + (dst, hist, loff, klim) per local
// partition + one private histogram per warp (RANK_MATCH)
template <class KeyT> inline size_t part_smem_bytes(uint32_t fan, int threads, bool rank_match) {
  const size_t tile = (size_t)threads * PartCfg<KeyT>::kItems;
  const size_t f = (fan + 3) & ~3u;
  return tile * (sizeof(Slot<KeyT>) + sizeof(uint16_t)) + f * (sizeof(unsigned long long) + 3 * sizeof(uint32_t)) +
         (rank_match ? (size_t)(threads / 32) * f * sizeof(uint32_t) : 0) + 64;
}

// RANK_MATCH = false: rank by shared-memory atomics on one block histogram (the SM's shared atomic unit
//                     retires ~0.5 lanes/clk -> ~7.6 ms per 2^30 records chip wide, measured);
// RANK_MATCH = true : every warp owns a private histogram; lanes with the same partition are found with
//                     __match_any_sync and the group's leader bumps the counter with a plain load/store.
// PEER (multi-GPU exchange, exchange.cu): partition q belongs to owner q >> peer.owner_shift and its region lives in THAT
// GPU's receive buffer peer.base[owner] (a peer-mapped pointer: the stores below travel over NVLink), at the offset
// part_start[q]; `out` is unused.  The level-1 partition pass and the all-to-all are one kernel.
// HOT (with PEER; hot.cuh): tuples whose key is in the hot set are not sent to the owner of their bucket range; they form
// one more partition, peer.hot_q = the number of ranges, which stays in this GPU's own memory (peer.hot_base).
constexpr int kMaxPeers = 16;
struct PeerOut {
  void* base[kMaxPeers]; uint32_t owner_shift;
  uint32_t hot_q; void* hot_base; unsigned long long hot_cap; const void* hot_table;   // HOT only
};

template <int HASH, bool LEFTID, bool RECS, int THREADS, bool RANK_MATCH, bool PEER = false, bool HOT = false>
__global__ void __launch_bounds__(THREADS, THREADS <= 512 ? HJ3D_PART_MINBLOCKS : 1)
k_part_scatter(Src s, const uint2* __restrict__ tilemap, Dir d, PartFn pf, uint32_t n_parts, uint32_t fan,
               uint32_t rowid_base, unsigned long long cap,
               const unsigned long long* __restrict__ part_start, unsigned long long* __restrict__ cursor,
               Slot<typename HashT<HASH>::key_t>* __restrict__ out, PeerOut peer = PeerOut{}) {
  using KeyT = typename HashT<HASH>::key_t;
  using SlotT = Slot<KeyT>;
  constexpr int ITEMS = PartCfg<KeyT>::kItems, TILE = THREADS * ITEMS, WARPS = THREADS / 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t fpad = (fan + 3) & ~3u;
  SlotT*              tile = reinterpret_cast<SlotT*>(smem_raw);
  unsigned long long* dst  = reinterpret_cast<unsigned long long*>(tile + TILE);   // global index of sorted position 0 of the run
  uint32_t*           hist = reinterpret_cast<uint32_t*>(dst + fpad);
  uint32_t*           loff = hist + fpad;
  uint32_t*           klim = loff + fpad;                                          // sorted positions >= klim overflow the region
  uint32_t*           whist = klim + fpad;                                         // [WARPS][fpad] (RANK_MATCH)
  uint16_t*           pid  = reinterpret_cast<uint16_t*>(whist + (RANK_MATCH ? WARPS * fpad : 0));
  __shared__ uint32_t sm_scan[33];
  __shared__ uint32_t sm_q0;
  __shared__ HotEntry<KeyT> sm_hot[HOT ? kHotSlots : 1];
  const uint32_t warp = threadIdx.x >> 5;

  uint64_t t0; uint32_t tn;
  block_tile<TILE>(tilemap, s.n, t0, tn);
  if (HOT) {   // the hot set into shared memory (visible after the barrier below)
    const HotTable<KeyT>* ht = reinterpret_cast<const HotTable<KeyT>*>(peer.hot_table);
    for (uint32_t p = threadIdx.x; p < (uint32_t)kHotSlots; p += THREADS) sm_hot[p] = ht->slot[p];
  }
  if (RANK_MATCH) { for (uint32_t p = threadIdx.x; p < WARPS * fpad; p += THREADS) whist[p] = 0; }
  else            { for (uint32_t p = threadIdx.x; p < fan; p += THREADS) hist[p] = 0; }

  // ---- load the tile's keys (and ids): all loads are issued before the first use
  KeyT     key[ITEMS];
  uint32_t id[ITEMS];
  uint32_t dropmask = 0;
  // (ncu: the generic loads below -- 64-bit index arithmetic, gather / row-id checks and a bounds check per tuple -- were
  //  31 % of the level-1 kernel's instructions; full tiles of a plain row store take the short paths)
  if (RECS && tn == (uint32_t)TILE) {
    const SlotT* in = reinterpret_cast<const SlotT*>(s.base) + t0 + threadIdx.x;
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) { const SlotT r = in[j * THREADS]; key[j] = r.key; id[j] = r.rowid; }
  } else if (RECS) {
    const SlotT* in = reinterpret_cast<const SlotT*>(s.base) + t0;
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      const uint32_t li = j * THREADS + threadIdx.x;
      key[j] = 0; id[j] = 0;
      if (li < tn) { const SlotT r = in[li]; key[j] = r.key; id[j] = r.rowid; }
    }
  } else if (!s.gather && s.rowid_off == HJ3D_NO_ROWID && tn == (uint32_t)TILE && s.sel_op == 0) {
    const uint8_t* p0 = s.base + (t0 + threadIdx.x) * (uint64_t)s.stride + s.key_off;
    const uint32_t step = (uint32_t)THREADS * s.stride;
    const uint32_t id0 = (uint32_t)t0 + threadIdx.x + (LEFTID ? 0u : rowid_base);
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      key[j] = __ldg(reinterpret_cast<const KeyT*>(p0 + (size_t)((uint32_t)j * step)));
      id[j] = id0 + (uint32_t)j * THREADS;
    }
  } else {
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      const uint32_t li = j * THREADS + threadIdx.x;
      key[j] = 0; id[j] = 0;
      if (li < tn) {
        key[j] = src_key<KeyT>(s, t0 + li);
        id[j] = LEFTID ? src_leftid(s, t0 + li) : src_rowid(s, t0 + li) + rowid_base;
        if (!src_selected(s, t0 + li)) dropmask |= 1u << j;             // fused AlgSelection: the tuple joins no partition
      }
    }
  }
  uint32_t q0 = 0;
  if (n_parts > fan) {                               // level 2: the tile's coarse partition = fan consecutive fine ones from q0 on
    if (threadIdx.x == 0) {
      const uint32_t q = tn ? pf(HashT<HASH>::bucket(key[0], d)) : 0u;
      sm_q0 = q < n_parts ? (q / fan) * fan : 0u;
    }
    __syncthreads();
    q0 = sm_q0;
  } else {
    __syncthreads();                                 // the zeroed histogram is visible
  }

  // ---- rank every record inside its partition of this tile
  uint32_t pr[ITEMS];                               // (local partition << 16) | rank, 0xFFFFFFFF = dropped
  // the partition of a key: generic (exact fast-mod bucket, division by the range width) or, when the directory size and
  // the range width are powers of two (the reference's -R / -b 1 shapes), a mask and a shift
  auto rank_all = [&](auto part_of) {
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const uint32_t li = j * THREADS + threadIdx.x;
    pr[j] = 0xFFFFFFFFu;
    uint32_t lp = 0xFFFFFFFFu;
    if (li < tn && !((dropmask >> j) & 1u)) {
      const uint32_t q = part_of(key[j]);
      if (q < n_parts && q - q0 < fan) lp = q - q0;
    }
    if (RANK_MATCH) {
      const uint32_t act = __ballot_sync(0xffffffffu, lp != 0xFFFFFFFFu);
      if (lp != 0xFFFFFFFFu) {
        const uint32_t peers = __match_any_sync(act, lp);
        const uint32_t leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (lane_id() == leader) { old = whist[warp * fpad + lp]; whist[warp * fpad + lp] = old + __popc(peers); }
        old = __shfl_sync(peers, old, leader);
        pr[j] = (lp << 16) | (old + __popc(peers & ((1u << lane_id()) - 1)));   // rank inside this warp so far
      }
      __syncwarp();
    } else if (fan <= 8) {                                                  // few partitions: aggregate per warp first
      const bool valid = lp != 0xFFFFFFFFu;
      const uint32_t peers = peers_small(valid ? lp : 0u, valid);
      const uint32_t leader = valid ? (uint32_t)(__ffs(peers) - 1) : 0u;
      uint32_t base = 0;
      if (valid && lane_id() == leader) base = atomicAdd(&hist[lp], (uint32_t)__popc(peers));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (valid) pr[j] = (lp << 16) | (base + __popc(peers & ((1u << lane_id()) - 1u)));
    } else {
      if (lp != 0xFFFFFFFFu) pr[j] = (lp << 16) | atomicAdd(&hist[lp], 1u);
    }
  }
  };
  if (HOT)                     rank_all([&](KeyT k) { return hot_find(sm_hot, k, HashT<HASH>::hash_lo32(k)) >= 0 ? peer.hot_q : pf(HashT<HASH>::bucket(k, d)); });
  else if (d.is_pow2 && pf.is_pow2) rank_all([&](KeyT k) { const uint32_t b = (HashT<HASH>::hash_lo32(k) & d.pow2_mask) - pf.lo; return b < pf.nl ? b >> pf.shift : 0xFFFFFFFFu; });
  else                         rank_all([&](KeyT k) { return pf(HashT<HASH>::bucket(k, d)); });
  __syncthreads();
  if (RANK_MATCH) {  // per partition: warp counts -> exclusive prefix over warps (each warp's base), total -> hist
    for (uint32_t p = threadIdx.x; p < fan; p += THREADS) {
      uint32_t run = 0;
#pragma unroll
      for (int w = 0; w < WARPS; ++w) { const uint32_t c = whist[w * fpad + p]; whist[w * fpad + p] = run; run += c; }
      hist[p] = run;
    }
    __syncthreads();
  }
  // exclusive scan of the tile histogram (fan <= 1024: entries threadIdx*k.. ) + one range reservation per partition
  {
    constexpr int PER = (kMaxParts + (HOT ? 1 : 0) + THREADS - 1) / THREADS;      // HOT: one partition past the 1024 ranges
    const uint32_t a = PER * threadIdx.x;
    uint32_t v[PER], sum = 0;
#pragma unroll
    for (int k = 0; k < PER; ++k) { v[k] = (a + k) < fan ? hist[a + k] : 0u; sum += v[k]; }
    // block-wide exclusive scan of `sum` (sm_scan is used once per block: no trailing barrier needed)
    uint32_t ex;
    {
      const uint32_t w = threadIdx.x >> 5, l = lane_id();
      const uint32_t inc = warp_iscan(sum);
      if (l == 31) sm_scan[w] = inc;
      __syncthreads();
      if (w == 0) { const uint32_t t = l < (uint32_t)WARPS ? sm_scan[l] : 0u; const uint32_t ti = warp_iscan(t); sm_scan[l] = ti - t; }
      __syncthreads();
      ex = inc - sum + sm_scan[w];
    }
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      if ((a + k) < fan) {
        const unsigned long long g = v[k] ? atomicAdd(&cursor[q0 + a + k], (unsigned long long)v[k]) : 0ull;
        loff[a + k] = ex;
        dst[a + k] = part_start[q0 + a + k] + g - ex;
        const unsigned long long cap_p = (HOT && q0 + a + k == peer.hot_q) ? peer.hot_cap
                                         : cap ? cap : part_start[q0 + a + k + 1] - part_start[q0 + a + k];   // cap == 0: planned regions
        const unsigned long long room = g < cap_p ? cap_p - g : 0ull;         // records of this run that still fit
        klim[a + k] = room >= (unsigned long long)v[k] ? 0xFFFFFFFFu : ex + (uint32_t)room;
      }
      ex += v[k];
    }
  }
  __syncthreads();
  // ---- sort the tile by partition in shared memory
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    if (pr[j] == 0xFFFFFFFFu) continue;
    const uint32_t lp = pr[j] >> 16;
    const uint32_t pos = loff[lp] + (pr[j] & 0xFFFFu) + (RANK_MATCH ? whist[warp * fpad + lp] : 0u);
    SlotT r; r.key = key[j]; r.rowid = id[j];
    tile[pos] = r;
    pid[pos] = (uint16_t)lp;
  }
  __syncthreads();
  // ---- coalesced write-out: consecutive threads store consecutive records of a run
  const uint32_t kept = loff[fan - 1] + hist[fan - 1];
  // (a 32-bit index variant of this loop with a no-overflow fast path measured 0.5 ms SLOWER per 2^30 records)
  for (uint32_t k = threadIdx.x; k < kept; k += THREADS) {
    const uint32_t lp = pid[k];
    SlotT* o = PEER ? reinterpret_cast<SlotT*>((HOT && q0 + lp == peer.hot_q) ? peer.hot_base : peer.base[(q0 + lp) >> peer.owner_shift]) : out;
    if (k < klim[lp]) o[dst[lp] + k] = tile[k];
  }
}

// ---- skew: regions planned from a sample -----------------------------------------------------------------------
// Fixed-capacity regions overflow when the keys are skewed and the pass is then repeated into exact regions.  Once a
// context has seen that happen, it sizes the regions from a sampled partition histogram instead: every `stride`-th tile
// contributes its first kSampleChunk records, the estimate is scaled up and padded by four standard deviations, and no
// region is smaller than the uniform one -- one streaming pass again (Zipf s = 1 at 2^30 probe tuples: 20.7 -> ~11 ms).
constexpr int kSampleChunk = 4096;

template <int HASH, bool RECS>
__global__ void __launch_bounds__(256)
k_part_sample(Src s, const uint2* __restrict__ tilemap, uint32_t tile, uint32_t stride, uint32_t n_tiles, Dir d, PartFn pf,
              uint32_t n_parts, uint32_t fan, unsigned long long* __restrict__ counts /* [n_parts + 1]; last = #records sampled */) {
  using KeyT = typename HashT<HASH>::key_t;
  __shared__ uint32_t h[kMaxParts];
  __shared__ uint32_t sm_q0;
  const uint32_t t = blockIdx.x * stride;
  if (t >= n_tiles) return;
  uint64_t t0; uint32_t tn;
  if (tilemap) { const uint2 e = tilemap[t]; t0 = e.x; tn = e.y; }
  else { t0 = (uint64_t)t * tile; tn = (s.n - t0) < (uint64_t)tile ? (uint32_t)(s.n - t0) : tile; }
  tn = tn < (uint32_t)kSampleChunk ? tn : (uint32_t)kSampleChunk;
  for (uint32_t p = threadIdx.x; p < fan; p += 256) h[p] = 0;
  auto key_of = [&](uint32_t li) -> KeyT {
    if (RECS) return reinterpret_cast<const Slot<KeyT>*>(s.base)[t0 + li].key;
    return src_key<KeyT>(s, t0 + li);
  };
  if (threadIdx.x == 0) {
    const uint32_t q = tn ? pf(HashT<HASH>::bucket(key_of(0), d)) : 0u;
    sm_q0 = (n_parts > fan && q < n_parts) ? (q / fan) * fan : 0u;
  }
  __syncthreads();
  const uint32_t q0 = sm_q0;
  for (uint32_t li = threadIdx.x; li < tn; li += 256) {
    const uint32_t q = pf(HashT<HASH>::bucket(key_of(li), d));
    if (q < n_parts && q - q0 < fan) atomicAdd(&h[q - q0], 1u);
  }
  __syncthreads();
  for (uint32_t p = threadIdx.x; p < fan; p += 256)
    if (h[p]) atomicAdd(&counts[q0 + p], (unsigned long long)h[p]);
  if (threadIdx.x == 0 && tn) atomicAdd(&counts[n_parts], (unsigned long long)tn);
}

// caps[p] = max(uniform capacity, estimate + 3 % + 4 sigma + 1024); caps[n_parts] = 0 (scan sentinel)
static __global__ void k_plan_caps(const unsigned long long* __restrict__ counts_s, uint32_t n_parts, unsigned long long n_total,
                            unsigned long long cap_uniform, unsigned long long* __restrict__ caps) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p > n_parts) return;
  if (p == n_parts) { caps[p] = 0; return; }
  const double sampled = (double)counts_s[n_parts];
  const double scale = sampled > 0.0 ? (double)n_total / sampled : 0.0;
  const double est = (double)counts_s[p] * scale;
  const double want = est * 1.03 + 4.0 * sqrt(est * (scale > 1.0 ? scale : 1.0)) + 1024.0;
  const unsigned long long w = (unsigned long long)want + 1ull;
  caps[p] = w > cap_uniform ? w : cap_uniform;
}

// out[0] = 1 if some partition received more records than its planned region holds, out[1] = records kept
static __global__ void k_part_overflow(const unsigned long long* __restrict__ counts, const unsigned long long* __restrict__ part_start,
                                uint32_t n_parts, unsigned long long* out) {
  unsigned long long over = 0, sum = 0;
  for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < n_parts; p += gridDim.x * blockDim.x) {
    const unsigned long long c = counts[p];
    sum += c;
    over |= c > part_start[p + 1] - part_start[p] ? 1ull : 0ull;
  }
  sum = warp_sum(sum); over = warp_max(over);
  if (lane_id() == 0) { if (over) atomicMax(&out[0], 1ull); atomicAdd(&out[1], sum); }
}

// host-side launcher: picks the instantiation for (recs, threads, rank_match)
template <int HASH, bool LEFTID>
inline cudaError_t launch_part_scatter(cudaStream_t st, bool recs, int threads, bool rank_match, uint32_t n_tiles,
                                       Src s, const uint2* tilemap, Dir d, PartFn pf, uint32_t n_parts, uint32_t fan,
                                       uint32_t rowid_base, unsigned long long cap, const unsigned long long* part_start,
                                       unsigned long long* cursor, Slot<typename HashT<HASH>::key_t>* out) {
  using KeyT = typename HashT<HASH>::key_t;
  const size_t sm = part_smem_bytes<KeyT>(fan, threads, rank_match);
#define HJ_PS(R, T, M)                                                                                            \
  do {                                                                                                            \
    cudaError_t e = cudaFuncSetAttribute(k_part_scatter<HASH, LEFTID, R, T, M>,                                   \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);                   \
    if (e != cudaSuccess) return e;                                                                               \
    if (n_tiles) k_part_scatter<HASH, LEFTID, R, T, M><<<n_tiles, T, sm, st>>>(s, tilemap, d, pf, n_parts, fan,   \
                                                                              rowid_base, cap, part_start, cursor, out); \
    return cudaGetLastError();                                                                                    \
  } while (0)
  if (threads == 1024) { if (recs) HJ_PS(true, 1024, false); else HJ_PS(false, 1024, false); }
  if (recs) { if (threads == 512) { if (rank_match) HJ_PS(true, 512, true); else HJ_PS(true, 512, false); }
              else                { if (rank_match) HJ_PS(true, 256, true); else HJ_PS(true, 256, false); } }
  else      { if (threads == 512) { if (rank_match) HJ_PS(false, 512, true); else HJ_PS(false, 512, false); }
              else                { if (rank_match) HJ_PS(false, 256, true); else HJ_PS(false, 256, false); } }
#undef HJ_PS
}

// ---- tile maps: block -> (first record, count) over partition regions with gaps --------------------
// tile_prefix[p] = number of tiles of partitions < p;  tile_prefix[n_parts] = total (single block)
static __global__ void __launch_bounds__(1024)
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
static __global__ void k_make_tilemap(const unsigned long long* __restrict__ part_start, const unsigned long long* __restrict__ counts,
                               const uint32_t* __restrict__ tile_prefix, uint32_t tile, uint2* __restrict__ tilemap,
                               uint32_t* __restrict__ tile_part /* nullable: partition of every tile */) {
  const uint32_t p = blockIdx.x;
  const unsigned long long cnt = counts[p], st = part_start[p];
  const uint32_t nt = (uint32_t)((cnt + tile - 1) / tile), t0 = tile_prefix[p];
  for (uint32_t t = threadIdx.x; t < nt; t += blockDim.x) {
    const unsigned long long off = (unsigned long long)t * tile;
    const unsigned long long left = cnt - off;
    tilemap[t0 + t] = make_uint2((uint32_t)(st + off), (uint32_t)(left < tile ? left : tile));
    if (tile_part) tile_part[t0 + t] = p;
  }
}

// work list over an unpartitioned input: chunk i = records [i*chunk, ..), all in fine partition 0
static __global__ void k_make_chunks(uint64_t n, uint32_t chunk, uint32_t n_chunks, uint2* __restrict__ work, uint32_t* __restrict__ work_part) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_chunks) return;
  const uint64_t st = (uint64_t)i * chunk, left = n - st;
  work[i] = make_uint2((uint32_t)st, (uint32_t)(left < chunk ? left : chunk));
  work_part[i] = 0;
}

static __global__ void k_fixed_starts_u64(uint32_t n, unsigned long long cap, unsigned long long* __restrict__ st) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) st[p] = (unsigned long long)p * cap;
}

}  // namespace hj3d
