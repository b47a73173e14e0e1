// probe_unnest.cuh -- nested probe and unnest in one kernel, for plans in which AlgUnnestHt directly follows
// AlgNestJoinProbe (main_experiment1.cc runNrs / runNsr: scan -> nested probe -> unnest -> top).  The nested tuples
// (probe row, group) never go to memory: a probe record that finds its group is expanded on the spot into
// (probe row, build row) pairs.  Deferred unnesting proper -- other operators between the probe and the unnest,
// main_experiment4.cc:846-867 -- keeps using hj3d_probe_nested + hj3d_unnest(_pairs).
//
// Same work decomposition as k_probe_fine (one block = one fine partition's slice in shared memory + a chunk of its
// probe records); the expansion is the warp-cooperative one of unnest.cuh (shuffle search over the round's exclusive
// offsets, dense coalesced stores), groups longer than kUnnestWarpMax are copied by the whole warp one after the other.
// Counters: matches / num_cmps are those of the nested probe (algebra.hh:449), out_cursor counts the flat results
// (AlgUnnestHt::_count), the checksum is over the flat pairs.
#pragma once

#include "common.cuh"
#include "probe.cuh"
#include "probe_fine.cuh"
#include "probe_smem.cuh"
#include "unnest.cuh"

namespace hj3d {

// A probe record whose group is longer than kUnnestWarpMax rows (skewed keys: one Zipf key owns millions of build rows): its
// output range is reserved with the tile's, the copy is left to k_unnest_hot_groups, which spreads one group over many blocks.
struct __align__(8) HotGroup { uint32_t left, start, len, pad; unsigned long long pos; };
constexpr uint32_t kHotChunk = 8192;       // rows one block copies at a time
constexpr uint32_t kHotSplit = 64;         // blocks that share one hot group

template <int HASH, bool CHECKSUM, bool WRITE>
__device__ __forceinline__ void probe_unnest_items(const Slot<typename HashT<HASH>::key_t>* __restrict__ in, uint32_t n_rec, const Dir& d,
                                                   uint32_t bucket_base, uint32_t nbk, const uint32_t* offp, uint32_t row_base,
                                                   const Group<typename HashT<HASH>::key_t>* grp, const uint32_t* __restrict__ rows,
                                                   uint2* __restrict__ out, unsigned long long out_cap, DevCounters* ctr,
                                                   ProbeAcc& acc, unsigned long long* wsum64, unsigned long long* sm_base,
                                                   HotGroup* __restrict__ hot_list, uint32_t hot_cap, unsigned long long* hot_count) {
  using KeyT = typename HashT<HASH>::key_t;
  using SlotT = Slot<KeyT>;
  using GroupT = Group<KeyT>;
  constexpr int IT = kFineItems, NW = kFineThreads / 32;
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
  for (uint32_t t0 = 0; t0 < n_rec; t0 += kFineTile) {
    KeyT     key[IT];
    uint32_t id[IT], st[IT], len[IT];
#pragma unroll
    for (int j = 0; j < IT; ++j) {
      const uint32_t li = t0 + j * kFineThreads + threadIdx.x;
      key[j] = 0; id[j] = 0;
      if (li < n_rec) { const SlotT r = in[li]; key[j] = r.key; id[j] = r.rowid; }
    }
    uint32_t cmps = 0, nhit = 0;
    unsigned long long cmps_long = 0, mine = 0;
#pragma unroll
    for (int j = 0; j < IT; ++j) {
      st[j] = 0; len[j] = 0;
      const uint32_t lb = HashT<HASH>::bucket(key[j], d) - bucket_base;
      if (t0 + j * kFineThreads + threadIdx.x < n_rec && lb < nbk) {
        const uint32_t o0 = offp[lb], n = offp[lb + 1] - o0;
        uint32_t gidx = 0, first = 0;
        if (probe_bucket<KeyT, 1, GroupT>(key[j], grp + (o0 - row_base), o0, n, gidx, first, cmps, cmps_long)) {
          const GroupT& g = grp[gidx - row_base];
          st[j] = g.start; len[j] = g.len;
          ++nhit;
        }
      }
      mine += len[j];
    }
    acc.matches += nhit;
    acc.cmps += (unsigned long long)cmps + cmps_long;
    // one reservation of the block's flat results per tile
    const unsigned long long wtot = warp_sum(mine);
    if (lane == 0) wsum64[warp] = wtot;
    __syncthreads();
    unsigned long long before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) { const unsigned long long v = wsum64[w]; before += w < (int)warp ? v : 0ull; total += v; }
    if (threadIdx.x == 0) *sm_base = total ? atomicAdd(&ctr->out_cursor, total) : 0ull;
    __syncthreads();
    if (WRITE || CHECKSUM) {
      unsigned long long pos = *sm_base + before;
#pragma unroll
      for (int j = 0; j < IT; ++j) {
        const uint32_t l = len[j] > kUnnestWarpMax ? 0u : len[j];
        uint32_t inc = l;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= (uint32_t)o) inc += v; }
        const uint32_t ex = inc - l, T = __shfl_sync(0xffffffffu, inc, 31);
        if (__all_sync(0xffffffffu, l <= 1u)) {
          if (l) {
            const uint32_t row = __ldg(rows + st[j]);
            if (CHECKSUM) { const uint64_t mx = pair_mix(id[j], row); acc.sum += mx; acc.x ^= mx; }
            if (WRITE && pos + ex < out_cap) out[pos + ex] = make_uint2(id[j], row);
          }
        } else {
          for (uint32_t o = 0; o < T; o += 32) {
            const uint32_t idx = o + lane;
            uint32_t s = 0;                               // largest s with ex[s] <= idx
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) {
              const uint32_t v = __shfl_sync(0xffffffffu, ex, (s + step) & 31);
              if (v <= idx) s += step;
            }
            const uint32_t e = __shfl_sync(0xffffffffu, ex, s);
            const uint32_t b = __shfl_sync(0xffffffffu, st[j], s);
            const uint32_t lf = __shfl_sync(0xffffffffu, id[j], s);
            if (idx < T) {
              const uint32_t row = __ldg(rows + b + (idx - e));
              if (CHECKSUM) { const uint64_t mx = pair_mix(lf, row); acc.sum += mx; acc.x ^= mx; }
              if (WRITE && pos + idx < out_cap) out[pos + idx] = make_uint2(lf, row);
            }
          }
        }
        pos += T;
      }
#pragma unroll
      for (int j = 0; j < IT; ++j) {                      // hot groups: listed with their reserved output position, copied later
        uint32_t hot = __ballot_sync(0xffffffffu, len[j] > kUnnestWarpMax);
        while (hot) {
          const uint32_t s = __ffs(hot) - 1;
          hot &= hot - 1;
          const uint32_t L = __shfl_sync(0xffffffffu, len[j], s);
          if (lane == s) {
            const unsigned long long h = atomicAdd(hot_count, 1ull);
            if (h < hot_cap) { HotGroup g; g.left = id[j]; g.start = st[j]; g.len = L; g.pad = 0; g.pos = pos; hot_list[h] = g; }
          }
          pos += L;
        }
      }
    }
    __syncthreads();                                      // wsum64 / sm_base are rewritten by the next tile
  }
}

template <int HASH, bool CHECKSUM, bool WRITE>
__global__ void __launch_bounds__(kFineThreads, 3)
k_probe_nested_unnest(const Slot<typename HashT<HASH>::key_t>* __restrict__ recs, Dir d, FineCfg fc, const uint2* __restrict__ work,
                      const uint32_t* __restrict__ work_part, const uint32_t* __restrict__ goff,
                      const Group<typename HashT<HASH>::key_t>* __restrict__ groups, const uint32_t* __restrict__ rows,
                      uint2* __restrict__ out, unsigned long long out_cap, DevCounters* ctr,
                      HotGroup* __restrict__ hot_list, uint32_t hot_cap, unsigned long long* hot_count) {
  using KeyT = typename HashT<HASH>::key_t;
  using GroupT = Group<KeyT>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ unsigned long long wsum64[kFineThreads / 32];
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
  __shared__ unsigned long long sm_bar;
  if (fits) {
    stage_slice(sm_off, goff + blo, (nbk + 1) * 4, sm_groups, groups + glo, ngr * (uint32_t)sizeof(GroupT), &sm_bar);
  } else {
    __syncthreads();
  }
  ProbeAcc acc;
  const Slot<KeyT>* in = recs + w.x;
  if (fits) probe_unnest_items<HASH, CHECKSUM, WRITE>(in, w.y, d, d.lo + blo, nbk, sm_off, glo, sm_groups, rows, out, out_cap, ctr, acc, wsum64, &sm_base, hot_list, hot_cap, hot_count);
  else      probe_unnest_items<HASH, CHECKSUM, WRITE>(in, w.y, d, d.lo + blo, nbk, goff + blo, 0u, groups, rows, out, out_cap, ctr, acc, wsum64, &sm_base, hot_list, hot_cap, hot_count);
  commit_acc(acc, ctr, CHECKSUM);
}

// the listed hot groups: group e is copied by the kHotSplit blocks (e, 0..kHotSplit), kHotChunk rows at a time
template <bool CHECKSUM, bool WRITE>
__global__ void __launch_bounds__(256)
k_unnest_hot_groups(const HotGroup* __restrict__ hot_list, uint32_t n_hot, const uint32_t* __restrict__ rows,
                    uint2* __restrict__ out, unsigned long long out_cap, DevCounters* ctr) {
  ProbeAcc acc;
  for (uint32_t e = blockIdx.x; e < n_hot; e += gridDim.x) {
    const HotGroup g = hot_list[e];
    for (uint32_t c0 = blockIdx.y * kHotChunk; c0 < g.len; c0 += kHotSplit * kHotChunk) {
      const uint32_t c1 = c0 + kHotChunk < g.len ? c0 + kHotChunk : g.len;
      for (uint32_t r = c0 + threadIdx.x; r < c1; r += 256) {
        const uint32_t row = __ldg(rows + g.start + r);
        if (CHECKSUM) { const uint64_t mx = pair_mix(g.left, row); acc.sum += mx; acc.x ^= mx; }
        if (WRITE && g.pos + r < out_cap) out[g.pos + r] = make_uint2(g.left, row);
      }
    }
  }
  if (CHECKSUM) commit_acc(acc, ctr, true);
}

}  // namespace hj3d
