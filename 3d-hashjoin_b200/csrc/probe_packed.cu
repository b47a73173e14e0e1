// probe_packed.cu -- kernels of the compressed-slice probe (see probe_packed.cuh for the design) and their launcher.
#include "probe_packed.cuh"

#ifndef HJ3D_PK_PREFETCH
#define HJ3D_PK_PREFETCH 0
#endif

namespace hj3d {

__device__ __forceinline__ uint32_t pk_quotient(uint32_t h, const PackCfg& pc) {
  return pc.pow2 ? (pc.qshift >= 32 ? 0u : h >> pc.qshift) : (uint32_t)__umul64hi(pc.qmagic, (uint64_t)h);
}

// long (unordered) bucket: first match in chain order [oldest, newest, .., second oldest] from row ids (probe.cuh)
__device__ __noinline__ uint32_t pk_first_long(uint32_t q, const uint32_t* pk, uint32_t n, uint32_t rb, uint32_t* right, unsigned long long* cmps) {
  const uint32_t rmask = rb >= 32 ? 0xFFFFFFFFu : (1u << rb) - 1u;
  uint32_t min_row = 0xFFFFFFFFu, best = 0; bool any = false, min_is_match = false;
  for (uint32_t k = 0; k < n; ++k) {
    const uint32_t w = pk[k], row = w & rmask; const bool hit = (rb >= 32 ? 0u : w >> rb) == q;
    if (row < min_row) { min_row = row; min_is_match = hit; }
    if (hit && (!any || row > best)) { best = row; any = true; }
  }
  if (!any) { *cmps += n; return 0; }
  if (min_is_match) { *cmps += 1; *right = min_row; return 1; }
  uint32_t rk = 0;
  for (uint32_t k = 0; k < n; ++k) rk += (pk[k] & rmask) < best;
  *cmps += n - rk + 1; *right = best;
  return 1;
}
__device__ __noinline__ uint32_t pk_first_long_global(uint32_t key, const Slot<uint32_t>* sp, uint32_t n, uint32_t* right, unsigned long long* cmps) {
  uint32_t min_row = 0xFFFFFFFFu, best = 0; bool any = false, min_is_match = false;
  for (uint32_t k = 0; k < n; ++k) {
    const Slot<uint32_t> sl = sp[k]; const bool hit = sl.key == key;
    if (sl.rowid < min_row) { min_row = sl.rowid; min_is_match = hit; }
    if (hit && (!any || sl.rowid > best)) { best = sl.rowid; any = true; }
  }
  if (!any) { *cmps += n; return 0; }
  if (min_is_match) { *cmps += 1; *right = min_row; return 1; }
  uint32_t rk = 0;
  for (uint32_t k = 0; k < n; ++k) rk += sp[k].rowid < best;
  *cmps += n - rk + 1; *right = best;
  return 1;
}

// the tile loop; SMEM: lookups go to the compressed slice, else (a slice that does not fit: skewed keys) to the global table
template <bool CHECKSUM, bool WRITE, bool SMEM>
__device__ __forceinline__ void pk_items(const Slot<uint32_t>* __restrict__ in, uint32_t n_rec, const Dir& d, const PackCfg& pc,
                                         uint32_t bucket_base, uint32_t nbk, const uint16_t* soff, const uint32_t* packed,
                                         const uint32_t* __restrict__ goff, const Slot<uint32_t>* __restrict__ gslots,
                                         uint2* __restrict__ out, unsigned long long out_cap, DevCounters* ctr, ProbeAcc& acc,
                                         uint32_t* wsum, unsigned long long* sm_base) {
  constexpr int IT = kPkItems, NW = kPkThreads / 32;
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
  const uint32_t rb = pc.rowid_bits, rmask = rb >= 32 ? 0xFFFFFFFFu : (1u << rb) - 1u;
  const bool aligned = ((uintptr_t)in & 15) == 0;                       // block-uniform
  uint32_t key[IT], id[IT], nkey[HJ3D_PK_PREFETCH ? IT : 1], nid[HJ3D_PK_PREFETCH ? IT : 1];
  auto fetch = [&](uint32_t t0, uint32_t* k, uint32_t* i) {
    if (aligned && t0 + kPkTile <= n_rec) {                             // two records per 128-bit load
      const uint4* p = reinterpret_cast<const uint4*>(in + t0) + threadIdx.x;
#pragma unroll
      for (int j = 0; j < IT / 2; ++j) { const uint4 r = __ldg(p + j * kPkThreads); k[2 * j] = r.x; i[2 * j] = r.y; k[2 * j + 1] = r.z; i[2 * j + 1] = r.w; }
    } else {
#pragma unroll
      for (int j = 0; j < IT; ++j) {
        const uint32_t li = t0 + (uint32_t)(j >> 1) * (2 * kPkThreads) + 2 * threadIdx.x + (j & 1);   // same record -> thread map as above
        k[j] = 0; i[j] = 0;
        if (li < n_rec) { const Slot<uint32_t> r = in[li]; k[j] = r.key; i[j] = r.rowid; }
      }
    }
  };
  if (HJ3D_PK_PREFETCH) fetch(0, key, id);
  for (uint32_t t0 = 0; t0 < n_rec; t0 += kPkTile) {
    if (HJ3D_PK_PREFETCH) { if (t0 + kPkTile < n_rec) fetch(t0 + kPkTile, nkey, nid); }   // in flight while this tile is probed
    else fetch(t0, key, id);
    uint32_t hitmask = 0, cmps = 0, wtot = 0;
    unsigned long long cmps_long = 0;
#pragma unroll
    for (int j = 0; j < IT; ++j) {
      const uint32_t li = t0 + (uint32_t)(j >> 1) * (2 * kPkThreads) + 2 * threadIdx.x + (j & 1);
      uint32_t right = 0, hit = 0;
      const uint32_t h = murmur32(key[j]);
      const uint32_t lb = mod_u32(h, d) - bucket_base;
      if (li < n_rec && lb < nbk) {
        if (SMEM) {
          const uint32_t o0 = soff[lb], n = (uint32_t)soff[lb + 1] - o0;
          const uint32_t q = pk_quotient(h, pc);
          if (n != 0) {                                                 // empty bucket: no comparison (algebra.hh:640-643)
            if (n <= kOrderedMax) {
              uint32_t k = 0;
              for (;;) {
                const uint32_t w = packed[o0 + k];
                ++k;
                if ((rb >= 32 ? 0u : w >> rb) == q) { right = w & rmask; hit = 1; cmps += k; break; }
                if (k == n) { cmps += n; break; }
              }
            } else {
              hit = pk_first_long(q, packed + o0, n, rb, &right, &cmps_long);
            }
          }
        } else {
          const uint32_t o0 = goff[lb], n = goff[lb + 1] - o0;
          if (n != 0) {
            if (n <= kOrderedMax) {
              uint32_t k = 0;
              for (;;) {
                const Slot<uint32_t> sl = gslots[o0 + k];
                ++k;
                if (sl.key == key[j]) { right = sl.rowid; hit = 1; cmps += k; break; }
                if (k == n) { cmps += n; break; }
              }
            } else {
              hit = pk_first_long_global(key[j], gslots + o0, n, &right, &cmps_long);
            }
          }
        }
      }
      if (CHECKSUM && hit) { const uint64_t mx = pair_mix(id[j], right); acc.sum += mx; acc.x ^= mx; }
      hitmask |= hit << j;
      key[j] = right;                                                   // the key's register now holds the result
      if (WRITE) wtot += __popc(__ballot_sync(0xffffffffu, hit));
    }
    acc.matches += __popc(hitmask);
    acc.cmps += (unsigned long long)cmps + cmps_long;
    if (WRITE) {
      if (lane == 0) wsum[warp] = wtot;
      __syncthreads();
      uint32_t before = 0, total = 0;
#pragma unroll
      for (int w = 0; w < NW; ++w) { const uint32_t v = wsum[w]; before += w < (int)warp ? v : 0u; total += v; }
      if (threadIdx.x == 0) *sm_base = total ? atomicAdd(&ctr->out_cursor, (unsigned long long)total) : 0ull;
      __syncthreads();
      unsigned long long pos = *sm_base + before;
#pragma unroll
      for (int j = 0; j < IT; ++j) {
        const uint32_t hit = (hitmask >> j) & 1u;
        const uint32_t bal = __ballot_sync(0xffffffffu, hit);
        const unsigned long long mypos = pos + __popc(bal & ((1u << lane) - 1u));
        if (hit && mypos < out_cap) out[mypos] = make_uint2(id[j], key[j]);
        pos += __popc(bal);
      }
    }
    if (HJ3D_PK_PREFETCH) {
#pragma unroll
      for (int j = 0; j < IT; ++j) { key[j] = nkey[j]; id[j] = nid[j]; }
    }
  }
}

template <bool CHECKSUM, bool WRITE>
__global__ void __launch_bounds__(kPkThreads, 2)
k_probe_packed(const Slot<uint32_t>* __restrict__ recs, Dir d, PackCfg pc, const uint2* __restrict__ work,
               const uint32_t* __restrict__ work_part, const uint32_t* __restrict__ off, const Slot<uint32_t>* __restrict__ slots,
               uint2* __restrict__ out, unsigned long long out_cap, DevCounters* ctr) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ uint32_t wsum[kPkThreads / 32];
  __shared__ unsigned long long sm_base;
  const uint2 w = work[blockIdx.x];
  const uint32_t f = work_part[blockIdx.x];
  const uint32_t blo = f * pc.width;
  const uint32_t bhi = (blo + pc.width < pc.n_local) ? blo + pc.width : pc.n_local;
  const uint32_t nbk = bhi - blo;
  const uint32_t rlo = off[blo], rhi = off[bhi];
  const uint32_t nrows = rhi - rlo;
  const uint32_t off_bytes = ((nbk + 1) * 2 + 15) & ~15u;
  const bool fits = nrows <= 0xFFFFu && (uint64_t)off_bytes + (uint64_t)nrows * 4 <= pc.smem_bytes;
  uint16_t* soff   = reinterpret_cast<uint16_t*>(smem_raw);
  uint32_t* packed = reinterpret_cast<uint32_t*>(smem_raw + off_bytes);
  if (fits) {
    // stage + compress the slice: run starts relative to the slice, slots as (quotient << rowid_bits) | row id
    const uint32_t* go = off + blo;
    for (uint32_t b = threadIdx.x; b <= nbk; b += kPkThreads) soff[b] = (uint16_t)(__ldg(go + b) - rlo);
    const uint32_t rb = pc.rowid_bits;
    const Slot<uint32_t>* gs = slots + rlo;
    if ((rlo & 1u) == 0) {                                              // 16-byte aligned: two slots per 128-bit load
      const uint4* gp = reinterpret_cast<const uint4*>(gs);
      const uint32_t n2 = nrows >> 1;
      for (uint32_t k = threadIdx.x; k < n2; k += kPkThreads) {
        const uint4 r = __ldg(gp + k);
        const uint32_t q0 = pk_quotient(murmur32(r.x), pc), q1 = pk_quotient(murmur32(r.z), pc);
        reinterpret_cast<uint2*>(packed)[k] = make_uint2((rb >= 32 ? 0u : q0 << rb) | r.y, (rb >= 32 ? 0u : q1 << rb) | r.w);
      }
      if ((nrows & 1u) && threadIdx.x == 0) {
        const Slot<uint32_t> r = gs[nrows - 1];
        packed[nrows - 1] = (rb >= 32 ? 0u : pk_quotient(murmur32(r.key), pc) << rb) | r.rowid;
      }
    } else {
      for (uint32_t k = threadIdx.x; k < nrows; k += kPkThreads) {
        const Slot<uint32_t> r = gs[k];
        packed[k] = (rb >= 32 ? 0u : pk_quotient(murmur32(r.key), pc) << rb) | r.rowid;
      }
    }
  }
  __syncthreads();
  ProbeAcc acc;
  const Slot<uint32_t>* in = recs + w.x;
  if (fits) pk_items<CHECKSUM, WRITE, true>(in, w.y, d, pc, d.lo + blo, nbk, soff, packed, nullptr, nullptr, out, out_cap, ctr, acc, wsum, &sm_base);
  else      pk_items<CHECKSUM, WRITE, false>(in, w.y, d, pc, d.lo + blo, nbk, nullptr, nullptr, off + blo, slots, out, out_cap, ctr, acc, wsum, &sm_base);
  commit_acc(acc, ctr, CHECKSUM);
}


cudaError_t launch_probe_packed(cudaStream_t st, bool checksum, bool write, uint32_t n_work, size_t smem, const Slot<uint32_t>* recs, Dir d,
                                PackCfg pc, const uint2* work, const uint32_t* work_part, const uint32_t* off, const Slot<uint32_t>* slots,
                                uint2* out, unsigned long long out_cap, DevCounters* ctr) {
#define LAUNCH_PK(C, W) do { \
    cudaError_t e = cudaFuncSetAttribute(k_probe_packed<C, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e != cudaSuccess) return e; \
    k_probe_packed<C, W><<<n_work, kPkThreads, smem, st>>>(recs, d, pc, work, work_part, off, slots, out, out_cap, ctr); } while (0)
  if (checksum) { if (write) LAUNCH_PK(true, true); else LAUNCH_PK(true, false); }
  else          { if (write) LAUNCH_PK(false, true); else LAUNCH_PK(false, false); }
#undef LAUNCH_PK
  return cudaGetLastError();
}

}  // namespace hj3d
