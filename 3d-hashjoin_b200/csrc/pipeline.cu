// pipeline.cu -- the deferred-unnesting multi-join pipeline of main_experiment4 (plan Ndu) as ONE device pipeline:
//
//   AlgScan(R) -> AlgNestJoinProbe(S table) -> AlgNestJoinProbe(T table, key reached through r) -> AlgUnnestHt(T)
//              -> AlgUnnestHt(S) -> AlgTop                                   (main_experiment4.cc:846-867)
//
// The reference pushes one tuple at a time through the five operators; composing the single-operator C-ABI calls
// reproduces that but materialises every intermediate (nested RS tuples, nested RST tuples, the half-unnested
// (r, S-group, t) tuples, plus a gather between the two unnests).  Here a probe tuple looks its key up in BOTH nested
// tables, the only intermediate is one (r, S-group, T-group) entry per surviving tuple, and the two unnests run as one
// output-centric expansion: output o of an entry with groups of lenS x lenT rows is (r, rowsS[o / lenT], rowsT[o % lenT]).
// Work is split by OUTPUT ranges (binary search over the scanned product lengths), so one key with 10^6 results
// (duplicates per key 1000 x 1000, BASELINE config 3) is expanded by many blocks.
//
// Counters are the reference's: count / numCmps of both probes (algebra.hh:435-459, ht_nested.hh:354-382), the first
// unnest's count = sum of lenT, the second's = sum of lenS * lenT = c_top (algebra.hh:510-541, main_experiment4.cc:593-597).
#include "engine_internal.hh"
#include "probe.cuh"
#include "scan.cuh"

namespace {

struct P2Counters {
  unsigned long long match_s, cmps_s, match_t, cmps_t, unnest1, entries;
};

struct P2Entry { uint32_t left, gs, gt, pad; };

// the main-chain walk of findMainNodeByOther (ht_nested.hh:354-382) over a bucket's group records: first-appearance order
// for chains of <= kOrderedMax groups, position derived from first_row for longer (unordered) ones -- probe.cuh
template <class KeyT>
__device__ __forceinline__ bool nested_find(const Group<KeyT>* __restrict__ gp, uint32_t dk, KeyT key, uint32_t* k_out, uint32_t* cmps) {
  if (dk == 0) return false;
  if (dk <= kOrderedMax) {
    for (uint32_t k = 0; k < dk; ++k)
      if (gp[k].key == key) { *k_out = k; *cmps += k + 1; return true; }
    *cmps += dk;
    return false;
  }
  uint32_t my_first = 0, my_k = 0; bool found = false;
  for (uint32_t k = 0; k < dk && !found; ++k) {
    if (gp[k].key == key) { found = true; my_first = gp[k].first_row; my_k = k; }
  }
  if (!found) { *cmps += dk; return false; }
  uint32_t before = 0;
  for (uint32_t k = 0; k < dk; ++k) before += gp[k].first_row < my_first;
  *cmps += before + 1; *k_out = my_k;
  return true;
}

constexpr int kP2Threads = 256;

template <int HASH, bool RECS>
__global__ void __launch_bounds__(kP2Threads)
k_probe2(Src s, Dir ds, Dir dt, const uint32_t* __restrict__ goff_s, const Group<typename HashT<HASH>::key_t>* __restrict__ grp_s,
         const uint32_t* __restrict__ goff_t, const Group<typename HashT<HASH>::key_t>* __restrict__ grp_t,
         P2Entry* __restrict__ entries, unsigned long long* __restrict__ plen, P2Counters* ctr) {
  using KeyT = typename HashT<HASH>::key_t;
  __shared__ unsigned long long red[5];
  __shared__ uint32_t wcnt[kP2Threads / 32];
  __shared__ unsigned long long sm_base;
  if (threadIdx.x < 5) red[threadIdx.x] = 0;
  __syncthreads();
  const uint64_t i = (uint64_t)blockIdx.x * kP2Threads + threadIdx.x;
  uint32_t ms = 0, mt = 0, cs = 0, ct = 0, gs = 0, gt = 0, left = 0;
  unsigned long long u1 = 0, prod = 0;
  if (i < s.n) {
    KeyT key;
    load_probe<KeyT, RECS>(s, i, key, left);
    const uint32_t bs = HashT<HASH>::bucket(key, ds) - ds.lo;
    if (bs < ds.n_local) {
      const uint32_t o0 = goff_s[bs], dk = goff_s[bs + 1] - o0;
      uint32_t k = 0;
      if (nested_find<KeyT>(grp_s + o0, dk, key, &k, &cs)) {
        ms = 1; gs = o0 + k;
        // the second probe sees only the tuples the first one let through (algebra.hh:447-457)
        const uint32_t bt = HashT<HASH>::bucket(key, dt) - dt.lo;
        if (bt < dt.n_local) {
          const uint32_t p0 = goff_t[bt], dk2 = goff_t[bt + 1] - p0;
          uint32_t k2 = 0;
          if (nested_find<KeyT>(grp_t + p0, dk2, key, &k2, &ct)) {
            mt = 1; gt = p0 + k2;
            const uint32_t ls = grp_s[gs].len, lt = grp_t[gt].len;
            u1 = lt; prod = (unsigned long long)ls * lt;
          }
        }
      }
    }
  }
  // compact the surviving entries (order is free)
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
  const uint32_t bal = __ballot_sync(0xffffffffu, mt);
  if (lane == 0) wcnt[warp] = __popc(bal);
  // counters: warp sums, one shared atomic per warp, one global atomic per block
  const unsigned long long v0 = warp_sum((unsigned long long)ms), v1 = warp_sum((unsigned long long)cs),
                           v2 = warp_sum((unsigned long long)mt), v3 = warp_sum((unsigned long long)ct), v4 = warp_sum(u1);
  if (lane == 0) { atomicAdd(&red[0], v0); atomicAdd(&red[1], v1); atomicAdd(&red[2], v2); atomicAdd(&red[3], v3); atomicAdd(&red[4], v4); }
  __syncthreads();
  uint32_t before = 0, total = 0;
#pragma unroll
  for (int w = 0; w < kP2Threads / 32; ++w) { const uint32_t v = wcnt[w]; before += w < (int)warp ? v : 0u; total += v; }
  if (threadIdx.x == 0) {
    sm_base = total ? atomicAdd(&ctr->entries, (unsigned long long)total) : 0ull;
    if (red[0]) atomicAdd(&ctr->match_s, red[0]);
    if (red[1]) atomicAdd(&ctr->cmps_s, red[1]);
    if (red[2]) atomicAdd(&ctr->match_t, red[2]);
    if (red[3]) atomicAdd(&ctr->cmps_t, red[3]);
    if (red[4]) atomicAdd(&ctr->unnest1, red[4]);
  }
  __syncthreads();
  if (mt) {
    const unsigned long long pos = sm_base + before + __popc(bal & ((1u << lane) - 1u));
    P2Entry e; e.left = left; e.gs = gs; e.gt = gt; e.pad = 0;
    entries[pos] = e;
    plen[pos] = prod;
  }
}

struct LoadBoundedU64 {   // plen[i] for i < *m (the compacted entry count lives on the device), 0 beyond
  const unsigned long long* p; const unsigned long long* m;
  __device__ unsigned long long operator()(uint64_t i) const { return i < *m ? p[i] : 0ull; }
};
struct StoreExclU64 { unsigned long long* p; __device__ void operator()(uint64_t i, unsigned long long ex, unsigned long long) const { p[i] = ex; } };

constexpr int kX2Threads = 256;
constexpr int kX2Chunk   = 4096;      // outputs per block

// outputs [blockIdx * kX2Chunk, +kX2Chunk): entry of an output = last e with offs[e] <= o.  The block's entries (at most
// one per output) are staged in shared memory, so the per-output search and the group lookups stay on chip.
template <class KeyT, bool CHECKSUM, bool WRITE>
__global__ void __launch_bounds__(kX2Threads)
k_expand2(const P2Entry* __restrict__ entries, const unsigned long long* __restrict__ offs, const unsigned long long* __restrict__ n_entries_p,
          unsigned long long total, const Group<KeyT>* __restrict__ grp_s, const uint32_t* __restrict__ rows_s,
          const Group<KeyT>* __restrict__ grp_t, const uint32_t* __restrict__ rows_t,
          uint32_t* __restrict__ out, unsigned long long out_cap, DevCounters* ctr) {
  __shared__ unsigned long long sm_off[kX2Chunk + 1];     // offsets of the staged entries, relative to global output 0
  __shared__ uint32_t sm_first;
  const unsigned long long o_lo = (unsigned long long)blockIdx.x * kX2Chunk;
  const unsigned long long o_hi = o_lo + kX2Chunk < total ? o_lo + kX2Chunk : total;
  const unsigned long long m = *n_entries_p;
  if (threadIdx.x == 0) {   // first entry overlapping the chunk: last e with offs[e] <= o_lo (entries have >= 1 output each)
    unsigned long long lo = 0, hi = m;
    while (hi - lo > 1) { const unsigned long long mid = (lo + hi) >> 1; if (offs[mid] <= o_lo) lo = mid; else hi = mid; }
    sm_first = (uint32_t)lo;
  }
  __syncthreads();
  const uint32_t e0 = sm_first;
  // stage offsets of entries e0 .. e0 + cnt (until one starts at or beyond o_hi)
  uint32_t cnt = 0;
  for (uint32_t k = threadIdx.x; k <= (uint32_t)kX2Chunk; k += kX2Threads) {
    const unsigned long long e = (unsigned long long)e0 + k;
    sm_off[k] = e < m ? offs[e] : total;
  }
  __syncthreads();
  ProbeAcc acc;
  for (unsigned long long o = o_lo + threadIdx.x; o < o_hi; o += kX2Threads) {
    // largest k with sm_off[k] <= o; k <= o - o_lo because every entry has at least one output
    uint32_t lo = 0, hi = (uint32_t)(o - o_lo) + 1;
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (sm_off[mid] <= o) lo = mid; else hi = mid; }
    const P2Entry e = entries[e0 + lo];
    const uint64_t within = o - sm_off[lo];
    const Group<KeyT> gs = grp_s[e.gs];
    const Group<KeyT> gt = grp_t[e.gt];
    uint32_t a, b;
    if (gt.len == 1) { a = (uint32_t)within; b = 0; }
    else { a = (uint32_t)(within / gt.len); b = (uint32_t)(within - (uint64_t)a * gt.len); }
    const uint32_t srow = __ldg(rows_s + gs.start + a), trow = __ldg(rows_t + gt.start + b);
    if (CHECKSUM) { const uint64_t mx = pair_mix((uint32_t)pair_mix(e.left, srow), trow); acc.sum += mx; acc.x ^= mx; }
    if (WRITE && o < out_cap) { out[3 * o] = e.left; out[3 * o + 1] = srow; out[3 * o + 2] = trow; }
  }
  (void)cnt;
  if (CHECKSUM) commit_acc(acc, ctr, true);
}

template <int HASH>
int probe2_unnest2_impl(hj3d_ctx* c, hj3d_table* ts, hj3d_table* tt, Src src, uint32_t flags, uint32_t* d_out, uint64_t cap,
                        hj3d_counters* out4) {
  using KeyT = typename HashT<HASH>::key_t;
  const uint64_t n = src.n;
  P2Counters* d_ctr = nullptr; P2Entry* entries = nullptr; unsigned long long *plen = nullptr, *offs = nullptr;
  HJ_TRY(dev_alloc(c, &d_ctr, 1));
  HJ_TRY(dev_alloc(c, &entries, n)); HJ_TRY(dev_alloc(c, &plen, n)); HJ_TRY(dev_alloc(c, &offs, n + 1));
  CUDA_TRY(cudaMemsetAsync(d_ctr, 0, sizeof(P2Counters), c->stream));
  CUDA_TRY(cudaMemsetAsync(c->d_ctr, 0, sizeof(DevCounters), c->stream));
  const bool recs = !src.gather && src.stride == sizeof(Slot<KeyT>) && src.key_off == 0 && src.rowid_off == sizeof(KeyT) &&
                    ((uintptr_t)src.base % sizeof(Slot<KeyT>)) == 0;
  const uint32_t nb = blocks_for(n, kP2Threads);
  {
    PhaseTimer pt(c, PH_PROBE);
    if (nb) {
      if (recs) k_probe2<HASH, true><<<nb, kP2Threads, 0, c->stream>>>(src, ts->dir, tt->dir, ts->goff, (const Group<KeyT>*)ts->groups, tt->goff,
                                                                       (const Group<KeyT>*)tt->groups, entries, plen, d_ctr);
      else      k_probe2<HASH, false><<<nb, kP2Threads, 0, c->stream>>>(src, ts->dir, tt->dir, ts->goff, (const Group<KeyT>*)ts->groups, tt->goff,
                                                                        (const Group<KeyT>*)tt->groups, entries, plen, d_ctr);
      ++c->launches;
    }
  }
  // exclusive scan of the product lengths over the worst-case n slots (entries beyond the compacted count read as 0)
  const unsigned long long* d_m = &d_ctr->entries;
  unsigned long long* d_total = c->d_scalar;
  {
    PhaseTimer pt(c, PH_SCAN);
    const uint32_t nbs = blocks_for(n + 1, kScanTile);
    unsigned long long* sums = nullptr;
    HJ_TRY(dev_alloc(c, &sums, nbs ? nbs : 1));
    if (nbs) {
      k_scan_reduce<unsigned long long, LoadBoundedU64, false><<<nbs, kScanThreads, 0, c->stream>>>(LoadBoundedU64{plen, d_m}, n + 1, sums, nullptr);
      k_scan_blocksums<unsigned long long><<<1, 1024, 0, c->stream>>>(sums, nbs, d_total);
      k_scan_apply<unsigned long long, LoadBoundedU64, StoreExclU64><<<nbs, kScanThreads, 0, c->stream>>>(LoadBoundedU64{plen, d_m}, StoreExclU64{offs}, n + 1, sums);
      c->launches += 3;
    } else {
      CUDA_TRY(cudaMemsetAsync(d_total, 0, 8, c->stream));
    }
  }
  struct HostBlock { P2Counters p; unsigned long long total; };
  HostBlock* h = (HostBlock*)c->h_pinned;
  CUDA_TRY(cudaMemcpyAsync(&h->p, d_ctr, sizeof(P2Counters), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaMemcpyAsync(&h->total, d_total, 8, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  CUDA_TRY(cudaGetLastError());
  const P2Counters pc = h->p;
  const unsigned long long total = h->total;
  const bool cs = flags & HJ3D_F_CHECKSUM, wr = d_out != nullptr;
  if (total && (cs || wr)) {
    PhaseTimer pt(c, PH_UNNEST);
    const unsigned long long nbx = (total + kX2Chunk - 1) / kX2Chunk;
    if (nbx > 0x7FFFFFFFull) return fail(HJ3D_ERR_UNSUPPORTED, "more than 2^43 flat results");
#define LAUNCH_X2(C, W) k_expand2<KeyT, C, W><<<(uint32_t)nbx, kX2Threads, 0, c->stream>>>(entries, offs, d_m, total, (const Group<KeyT>*)ts->groups, ts->rows, \
                                                                                          (const Group<KeyT>*)tt->groups, tt->rows, d_out, cap, c->d_ctr)
    if (cs) { if (wr) LAUNCH_X2(true, true); else LAUNCH_X2(true, false); }
    else    { LAUNCH_X2(false, true); }
#undef LAUNCH_X2
    ++c->launches;
    CUDA_TRY(cudaGetLastError());
  }
  DevCounters* hc = (DevCounters*)((char*)c->h_pinned + 1024);
  CUDA_TRY(cudaMemcpyAsync(hc, c->d_ctr, sizeof(DevCounters), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  CUDA_TRY(cudaGetLastError());
  memset(out4, 0, 4 * sizeof(hj3d_counters));
  out4[0].matches = out4[0].out_tuples = pc.match_s; out4[0].num_cmps = pc.cmps_s;
  out4[1].matches = out4[1].out_tuples = pc.match_t; out4[1].num_cmps = pc.cmps_t;
  out4[2].matches = out4[2].out_tuples = pc.unnest1;
  out4[3].matches = out4[3].out_tuples = total;
  out4[3].checksum_sum = hc->checksum_sum; out4[3].checksum_xor = hc->checksum_xor;
  out4[3].overflow = (wr && total > cap) ? 1 : 0;
  out4[3].out_written = wr ? (total > cap ? cap : total) : 0;
  return HJ3D_OK;
}

}  // namespace

extern "C" int hj3d_probe2_unnest2(hj3d_ctx* c, hj3d_table* ts, hj3d_table* tt, const void* d_probe, uint64_t n, hj3d_keyspec ks,
                                   uint32_t flags, uint32_t* d_out_triples, uint64_t out_cap, hj3d_counters* out4) {
  if (!c || !ts || !tt || !out4) return fail(HJ3D_ERR_INVALID, "NULL argument");
  if (ts->kind != HJ3D_NESTED || tt->kind != HJ3D_NESTED) return fail(HJ3D_ERR_INVALID, "hj3d_probe2_unnest2 needs two nested tables");
  if (!ts->built || !tt->built) return fail(HJ3D_ERR_INVALID, "table has not been built");
  if ((int)ks.hash_id != ts->hash_id || (int)ks.hash_id != tt->hash_id)
    return fail(HJ3D_ERR_INVALID, "probe hash function differs from a build hash function (static_assert in ht_nested.hh:361)");
  if (n && !d_probe) return fail(HJ3D_ERR_INVALID, "d_probe == NULL");
  if (n > 0xFFFFFFF0ull) return fail(HJ3D_ERR_UNSUPPORTED, "more than 2^32-16 probe tuples");
  HJ_TRY(check_keyspec(ks));
  CUDA_TRY(cudaSetDevice(c->device));
  HJ_TRY(arena_reset(c));
  begin_call(c);
  Src src = make_src(d_probe, n, ks, nullptr);
  int rc;
  switch (ks.hash_id) {
    case HJ3D_HASH_MURMUR32: rc = probe2_unnest2_impl<HJ3D_HASH_MURMUR32>(c, ts, tt, src, flags, d_out_triples, out_cap, out4); break;
    case HJ3D_HASH_MURMUR64: rc = probe2_unnest2_impl<HJ3D_HASH_MURMUR64>(c, ts, tt, src, flags, d_out_triples, out_cap, out4); break;
    default:                 rc = probe2_unnest2_impl<HJ3D_HASH_MURMUR64_SEXT32>(c, ts, tt, src, flags, d_out_triples, out_cap, out4); break;
  }
  end_call(c);
  if (rc < 0) return rc;
  return out4[3].overflow ? HJ3D_OVERFLOW : HJ3D_OK;
}
