// hot.cuh -- hot-key probe replication for skewed probe sides (Zipf caveat).
//
// With Zipf-distributed foreign keys a handful of keys carry a large share of the probe relation; the bucket-range
// owner of such a key would receive (and answer) all of those tuples.  The table cannot split a bucket -- it is bit
// exact with the reference's -- but the PROBE tuples of a hot key need not travel: every one of them gets the same
// answer (the same build rows, the same number of comparisons), so
//   1. the ranks agree on a small set of hot keys from a sample of the probe side (k_hot_sample on every rank, one
//      all-gather of the sample, k_hot_select on every rank: same sample -> same set, no host round trip),
//   2. the exchange's partition kernel keeps tuples of hot keys in a local segment instead of sending them to the owner
//      (k_part_scatter<.., HOT>, partition.cuh),
//   3. once the tables are built, every rank looks the hot keys up in ITS shard (k_hot_answers_*: non-owners answer
//      "nothing"), one all-reduce (sum) of the answers gives every rank the owner's answer for every hot key,
//   4. k_hot_join expands the local hot segment with those answers: the same result pairs, the same count / numCmps as
//      if the tuples had been probed at the owner, only the work is spread over all GPUs.
// Keys with more than kHotAns partners are not supported (the probe call reports it): the feature is meant for foreign
// keys probing a (nearly) unique build side, plans Csr / CsrUU / Nsr of main_experiment1.cc with --skew.
#pragma once

#include "common.cuh"
#include "hot_set.cuh"
#include "probe.cuh"

namespace hj3d {

// ---- 1. sample: m keys of the local slice at a fixed stride (as 64-bit values; all ones = no tuple)
template <class KeyT>
__global__ void k_hot_sample(Src s, uint32_t m, unsigned long long* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const uint64_t stride = s.n / m ? s.n / m : 1;
  const uint64_t at = (uint64_t)i * stride;
  out[i] = at < s.n ? (unsigned long long)src_key<KeyT>(s, at) : ~0ull;
}

// ---- select: sort the gathered sample, take the (at most kHotMax) most frequent keys that occur >= kHotMinCount times.
// One block; deterministic (every rank computes the same table from the same sample).
template <int HASH>
__global__ void __launch_bounds__(1024)
k_hot_select(const unsigned long long* __restrict__ sample, uint32_t m, HotTable<typename HashT<HASH>::key_t>* __restrict__ out) {
  using KeyT = typename HashT<HASH>::key_t;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned long long* a = reinterpret_cast<unsigned long long*>(smem_raw);          // [kHotSample]
  uint16_t* cnt = reinterpret_cast<uint16_t*>(a + kHotSample);                       // [kHotSample] run length at run starts
  __shared__ uint32_t sm_n;
  for (uint32_t i = threadIdx.x; i < (uint32_t)kHotSample; i += blockDim.x) a[i] = i < m ? sample[i] : ~0ull;
  __syncthreads();
  for (uint32_t k = 2; k <= (uint32_t)kHotSample; k <<= 1)
    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
      for (uint32_t i = threadIdx.x; i < (uint32_t)kHotSample; i += blockDim.x) {
        const uint32_t p = i ^ j;
        if (p > i) {
          const unsigned long long x = a[i], y = a[p];
          const bool up = (i & k) == 0;
          if ((x > y) == up) { a[i] = y; a[p] = x; }
        }
      }
      __syncthreads();
    }
  for (uint32_t i = threadIdx.x; i < (uint32_t)kHotSample; i += blockDim.x) {
    const unsigned long long v = a[i];
    uint32_t c = 0;
    if (v != ~0ull && (i == 0 || a[i - 1] != v)) {
      uint32_t lo = i, hi = kHotSample;                      // first position past the run (binary search)
      while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (a[mid] == v) lo = mid; else hi = mid; }
      c = hi - i;
    }
    cnt[i] = (uint16_t)(c > 0xFFFFu ? 0xFFFFu : c);
  }
  __syncthreads();
  // the smallest threshold >= kHotMinCount that leaves at most kHotMax keys
  uint32_t thr = kHotMinCount;
  while (true) {
    if (threadIdx.x == 0) sm_n = 0;
    __syncthreads();
    uint32_t mine = 0;
    for (uint32_t i = threadIdx.x; i < (uint32_t)kHotSample; i += blockDim.x) mine += cnt[i] >= thr;
    mine = warp_sum(mine);
    if (lane_id() == 0 && mine) atomicAdd(&sm_n, mine);
    __syncthreads();
    const uint32_t n = sm_n;
    __syncthreads();
    if (n <= (uint32_t)kHotMax) break;
    thr += thr >> 1 ? thr >> 1 : 1;
  }
  for (uint32_t i = threadIdx.x; i < (uint32_t)kHotSlots; i += blockDim.x) { HotEntry<KeyT> e{}; out->slot[i] = e; }
  __syncthreads();
  if (threadIdx.x == 0) {                                    // in sample order: the same table on every rank
    uint32_t n = 0;
    for (uint32_t i = 0; i < (uint32_t)kHotSample && n < (uint32_t)kHotMax; ++i) {
      if (cnt[i] < thr) continue;
      const KeyT key = (KeyT)a[i];
      uint32_t s = hot_slot_of(HashT<HASH>::hash_lo32(key));
      while (out->slot[s].idx) s = (s + 1) & (kHotSlots - 1);
      HotEntry<KeyT> e{}; e.key = key; e.idx = n + 1;
      out->slot[s] = e;
      out->key[n] = key;
      ++n;
    }
    out->n = n;
  }
}

// ---- 3. answers: what a probe tuple with hot key k finds in THIS shard.  One thread per hot key; zeros if the key's
// bucket belongs to another rank.  The walks are the probe kernels' own (probe.cuh), so count / numCmps agree.
template <int HASH, bool UNIQUE>
__global__ void k_hot_answers_chaining(const HotTable<typename HashT<HASH>::key_t>* __restrict__ hot, Dir d,
                                       const uint32_t* __restrict__ off, const Slot<typename HashT<HASH>::key_t>* __restrict__ slots,
                                       HotAnswers* __restrict__ out) {
  using KeyT = typename HashT<HASH>::key_t;
  const uint32_t k = threadIdx.x;
  if (k >= (uint32_t)kHotMax) return;
  HotAns r{};
  if (k < hot->n) {
    const KeyT key = hot->key[k];
    const uint32_t b = HashT<HASH>::bucket(key, d) - d.lo;
    if (b < d.n_local) {
      const uint32_t o0 = off[b], n = off[b + 1] - o0;
      if (n) {
        uint32_t nm = 0, first = 0, cmps = 0;
        chain_walk<KeyT, UNIQUE>(slots + o0, n, key, nm, first, cmps);
        r.nm = nm; r.cmps = cmps;
        if (nm == 1) r.row[0] = first;
        else if (nm > 1 && nm <= (uint32_t)kHotAns) {
          uint32_t w = 0;
          for (uint32_t j = 0; j < n; ++j) { const Slot<KeyT> sl = slots[o0 + j]; if (sl.key == key) r.row[w++] = sl.rowid; }
        } else if (nm > (uint32_t)kHotAns) atomicAdd(&out->too_many, 1u);
      }
    }
  }
  out->a[k] = r;
}

// nested table: nm = 1 if the key has a group, cmps = the main-chain walk, rows = the group's members (for the unnest)
template <int HASH>
__global__ void k_hot_answers_nested(const HotTable<typename HashT<HASH>::key_t>* __restrict__ hot, Dir d,
                                     const uint32_t* __restrict__ goff, const Group<typename HashT<HASH>::key_t>* __restrict__ groups,
                                     const uint32_t* __restrict__ rows, HotAnswers* __restrict__ out) {
  using KeyT = typename HashT<HASH>::key_t;
  const uint32_t k = threadIdx.x;
  if (k >= (uint32_t)kHotMax) return;
  HotAns r{};
  if (k < hot->n) {
    const KeyT key = hot->key[k];
    const uint32_t b = HashT<HASH>::bucket(key, d) - d.lo;
    if (b < d.n_local) {
      const uint32_t o0 = goff[b], dk = goff[b + 1] - o0;
      if (dk) {
        bool hit = false; uint32_t g = 0, frow = 0, cmps = 0;
        group_walk<KeyT>(groups + o0, dk, key, hit, g, frow, cmps);
        r.cmps = cmps;
        if (hit) {
          const Group<KeyT> gr = groups[o0 + g];
          if (gr.len > (uint32_t)kHotAns) atomicAdd(&out->too_many, 1u);
          else { r.nm = gr.len; for (uint32_t j = 0; j < gr.len; ++j) r.row[j] = rows[gr.start + j]; }
        }
      }
    }
  }
  out->a[k] = r;
}

// ---- 4. join of the local hot segment (its own small pass after the regular probe; the host adds the counters up).
// ctr->out_cursor starts at the number of pairs already written and always advances by the number of flat results;
// ctr->matches counts what the PROBE operator counts: result tuples (chaining) or nested tuples = hits (NESTED).
template <int HASH, bool NESTED, bool CHECKSUM, bool WRITE>
__global__ void __launch_bounds__(256)
k_hot_join(const Slot<typename HashT<HASH>::key_t>* __restrict__ recs, unsigned long long n,
           const HotTable<typename HashT<HASH>::key_t>* __restrict__ hot, const HotAnswers* __restrict__ ans,
           uint2* __restrict__ out, unsigned long long out_cap, DevCounters* ctr) {
  using KeyT = typename HashT<HASH>::key_t;
  __shared__ HotEntry<KeyT> sm_slot[kHotSlots];
  __shared__ HotAns sm_ans[kHotMax];
  __shared__ unsigned long long sm_scan[33];
  __shared__ unsigned long long sm_base;
  for (uint32_t i = threadIdx.x; i < (uint32_t)kHotSlots; i += blockDim.x) sm_slot[i] = hot->slot[i];
  for (uint32_t i = threadIdx.x; i < (uint32_t)kHotMax; i += blockDim.x) sm_ans[i] = ans->a[i];
  __syncthreads();
  ProbeAcc acc;
  const unsigned long long step = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long base = (unsigned long long)blockIdx.x * blockDim.x; base < n; base += step) {
    const unsigned long long i = base + threadIdx.x;
    uint32_t nm = 0, left = 0; int k = -1;
    if (i < n) {
      const Slot<KeyT> r = recs[i];
      left = r.rowid;
      k = hot_find(sm_slot, r.key, HashT<HASH>::hash_lo32(r.key));
      if (k >= 0) { nm = sm_ans[k].nm; acc.cmps += sm_ans[k].cmps; acc.matches += NESTED ? (nm ? 1u : 0u) : nm; }
    }
    unsigned long long tot;
    const unsigned long long ex = block_exscan((unsigned long long)nm, sm_scan, &tot);
    if (threadIdx.x == 0) sm_base = tot ? atomicAdd(&ctr->out_cursor, tot) : 0ull;
    __syncthreads();
    unsigned long long pos = sm_base + ex;
    for (uint32_t j = 0; j < nm; ++j) {
      const uint32_t row = sm_ans[k].row[j];
      if (CHECKSUM) { const uint64_t mx = pair_mix(left, row); acc.sum += mx; acc.x ^= mx; }
      if (WRITE) { if (pos < out_cap) out[pos] = make_uint2(left, row); ++pos; }
    }
    __syncthreads();
  }
  commit_acc(acc, ctr, CHECKSUM);
}

}  // namespace hj3d
