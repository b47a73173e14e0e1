// ubench2.cu -- partition-pass design space (developer microbenchmark, not part of the product).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/ubench2.out tools/ubench2.cu
// Pass-1 shape: 12-byte tuples -> 8-byte records of P ranges.  Pass-2 shape: 8-byte records of a coarse range -> records of
// its P2 sub-ranges.  Variants: block shape / occupancy, ranking (shared atomics | ballot aggregation | match_any + private
// histograms), persistent blocks with register prefetch of the next tile.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__host__ __device__ __forceinline__ uint32_t murmur32(uint32_t x) {
  x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ uint32_t warp_iscan(uint32_t v) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { uint32_t w = __shfl_up_sync(0xffffffffu, v, o); if (lane_id() >= (uint32_t)o) v += w; }
  return v;
}
__global__ void k_gen(uint32_t* t, uint64_t n, uint32_t keymask) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    t[3 * i] = (uint32_t)i; t[3 * i + 1] = murmur32((uint32_t)i * 2654435761u + 17u) & keymask; t[3 * i + 2] = 0;
  }
}

__global__ void k_gen_recs(uint32_t* t, uint64_t n, uint32_t keymask) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    t[2 * i] = murmur32((uint32_t)i * 2654435761u + 99u) & keymask; t[2 * i + 1] = (uint32_t)i;
  }
}

// RECS: input is 8-byte records (pass 2) else 12-byte tuples with the key at word 1 (pass 1)
// RANK 0: shared atomics; 1: ballot aggregation over PB bits then one atomic per distinct value per warp; 2: match_any + warp-private histograms
// PERSIST: grid-stride over tiles with the next tile's loads issued before the current one is ranked
template <int TH, int IT, int MINB, bool RECS, int RANK, bool PERSIST>
__global__ void __launch_bounds__(TH, MINB)
k_part(const uint32_t* __restrict__ in, uint64_t n, uint32_t dbits, uint32_t pshift, uint32_t P, uint32_t pbits, uint32_t cap,
       uint32_t tiles_per_coarse, unsigned int* __restrict__ cursor, uint2* __restrict__ out) {
  constexpr int TILE = TH * IT, WARPS = TH / 32;
  extern __shared__ __align__(16) unsigned char smem[];
  uint2*    tile = reinterpret_cast<uint2*>(smem);
  uint32_t* hist = reinterpret_cast<uint32_t*>(tile + TILE);
  uint32_t* dst  = hist + P;
  uint16_t* pid  = reinterpret_cast<uint16_t*>(dst + P);
  uint32_t* whist = reinterpret_cast<uint32_t*>(pid + TILE);       // RANK 2: [WARPS][P]
  uint16_t* wtag = reinterpret_cast<uint16_t*>(pid + TILE);        // RANK 3: [WARPS][P] (count << 5 | lane)
  __shared__ uint32_t sm_scan[33];
  const uint32_t dmask = (1u << dbits) - 1u, warp = threadIdx.x >> 5;
  const uint64_t n_tiles = (n + TILE - 1) / TILE;
  uint32_t key[IT], id[IT], nkey[IT], nid[IT];
  auto fetch = [&](uint64_t t, uint32_t* k, uint32_t* i) {
    const uint64_t t0 = t * TILE;
#pragma unroll
    for (int j = 0; j < IT; ++j) {
      const uint64_t li = t0 + j * TH + threadIdx.x;
      if (RECS) { const uint2 r = li < n ? __ldg(reinterpret_cast<const uint2*>(in) + li) : make_uint2(0, 0); k[j] = r.x; i[j] = r.y; }
      else { k[j] = li < n ? __ldg(in + 3 * li + 1) : 0u; }
    }
  };
  uint64_t t = blockIdx.x;
  if (t < n_tiles) fetch(t, key, id);
  for (; t < n_tiles; t += gridDim.x) {
    const uint64_t t0 = t * TILE;
    const uint32_t tn = (n - t0) < (uint64_t)TILE ? (uint32_t)(n - t0) : (uint32_t)TILE;
    const uint32_t q0 = RECS ? (uint32_t)(t / tiles_per_coarse) * P : 0u;
    if (RANK == 2) { for (uint32_t p = threadIdx.x; p < WARPS * P; p += TH) whist[p] = 0; }
    else if (RANK == 3) { for (uint32_t p = threadIdx.x; p < WARPS * P / 2; p += TH) reinterpret_cast<uint32_t*>(wtag)[p] = 0; }
    else           { for (uint32_t p = threadIdx.x; p < P; p += TH) hist[p] = 0; }
    if (PERSIST && t + gridDim.x < n_tiles) fetch(t + gridDim.x, nkey, nid);
    __syncthreads();
    uint32_t pr[IT];
#pragma unroll
    for (int j = 0; j < IT; ++j) {
      const uint32_t li = j * TH + threadIdx.x;
      const uint32_t q = ((murmur32(key[j]) & dmask) >> pshift) & (P - 1);
      if (RANK == 1) {
        uint32_t peers = 0xffffffffu;
        for (uint32_t b = 0; b < pbits; ++b) { const uint32_t m = __ballot_sync(0xffffffffu, (q >> b) & 1u); peers &= ((q >> b) & 1u) ? m : ~m; }
        const uint32_t leader = __ffs(peers) - 1;
        uint32_t base = 0;
        if (lane_id() == leader) base = atomicAdd(&hist[q], (uint32_t)__popc(peers));
        base = __shfl_sync(0xffffffffu, base, leader);
        pr[j] = (q << 16) | (base + __popc(peers & ((1u << lane_id()) - 1u)));
      } else if (RANK == 2) {
        const uint32_t peers = __match_any_sync(0xffffffffu, q);
        const uint32_t leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (lane_id() == leader) { old = whist[warp * P + q]; whist[warp * P + q] = old + __popc(peers); }
        old = __shfl_sync(0xffffffffu, old, leader);
        pr[j] = (q << 16) | (old + __popc(peers & ((1u << lane_id()) - 1u)));
        __syncwarp();
      } else if (RANK == 3) {
        uint16_t* w = wtag + warp * P + q;
        uint32_t old = *w;
        const uint32_t lane = lane_id();
        bool done = false;
        uint32_t rank = 0;
        for (;;) {
          if (!done) *w = (uint16_t)((((old >> 5) + 1u) << 5) | lane);
          __syncwarp();
          uint32_t chk = 0;
          if (!done) { chk = *w; if ((chk & 31u) == lane) { done = true; rank = old >> 5; } else old = chk; }
          if (__all_sync(0xffffffffu, done)) break;
          __syncwarp();
        }
        pr[j] = (q << 16) | rank;
      } else {
        pr[j] = li < tn ? ((q << 16) | atomicAdd(&hist[q], 1u)) : 0xFFFFFFFFu;
      }
    }
    __syncthreads();
    if (RANK == 2) {
      for (uint32_t p = threadIdx.x; p < P; p += TH) {
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) { const uint32_t c = whist[w * P + p]; whist[w * P + p] = run; run += c; }
        hist[p] = run;
      }
      __syncthreads();
    }
    if (RANK == 3) {
      for (uint32_t p = threadIdx.x; p < P; p += TH) {
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) { const uint32_t c = wtag[w * P + p] >> 5; wtag[w * P + p] = (uint16_t)run; run += c; }
        hist[p] = run;
      }
      __syncthreads();
    }
    {
      const uint32_t PER = (P + TH - 1) / TH;
      const uint32_t a = PER * threadIdx.x;
      uint32_t sum = 0;
      for (uint32_t k = 0; k < PER; ++k) sum += (a + k) < P ? hist[a + k] : 0u;
      const uint32_t w = threadIdx.x >> 5, l = lane_id();
      const uint32_t inc = warp_iscan(sum);
      if (l == 31) sm_scan[w] = inc;
      __syncthreads();
      if (w == 0) { const uint32_t x = l < (uint32_t)WARPS ? sm_scan[l] : 0u; const uint32_t xi = warp_iscan(x); sm_scan[l] = xi - x; }
      __syncthreads();
      uint32_t ex = inc - sum + sm_scan[w];
      for (uint32_t k = 0; k < PER; ++k) {
        if ((a + k) < P) {
          const uint32_t v = hist[a + k];
          const uint32_t g = v ? atomicAdd(&cursor[q0 + a + k], v) : 0u;
          hist[a + k] = ex;
          dst[a + k] = (q0 + a + k) * cap + g - ex;
          ex += v;
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < IT; ++j) {
      if (pr[j] == 0xFFFFFFFFu) continue;
      const uint32_t lp = pr[j] >> 16;
      const uint32_t pos = hist[lp] + (pr[j] & 0xFFFFu) + (RANK == 2 ? whist[warp * P + lp] : (RANK == 3 ? (uint32_t)wtag[warp * P + lp] : 0u));
      tile[pos] = make_uint2(key[j], RECS ? id[j] : (uint32_t)(t0 + j * TH + threadIdx.x));
      pid[pos] = (uint16_t)lp;
    }
    __syncthreads();
    for (uint32_t k = threadIdx.x; k < tn; k += TH) {
      const uint32_t lp = pid[k];
      const uint32_t idx = dst[lp] + k;
      if (idx < (q0 + lp + 1) * cap) out[idx] = tile[k];
    }
    if (PERSIST) {
#pragma unroll
      for (int j = 0; j < IT; ++j) { key[j] = nkey[j]; id[j] = nid[j]; }
      __syncthreads();
    }
  }
}

template <int TH, int IT, int MINB, bool RECS, int RANK, bool PERSIST>
void run(const uint32_t* in, uint64_t n, uint32_t P, uint2* out, unsigned int* cursor) {
  const uint32_t dbits = 27;
  uint32_t pb = 0; while ((1u << pb) < P) ++pb;
  const uint32_t coarse = RECS ? 256u : 1u;
  const uint32_t pshift = RECS ? dbits - 8 - pb : dbits - pb;
  const uint64_t tiles = (n + (uint64_t)TH * IT - 1) / ((uint64_t)TH * IT);
  const uint32_t tiles_per_coarse = (uint32_t)((tiles + coarse - 1) / coarse);
  const uint64_t regions = (uint64_t)coarse * P;
  const uint32_t cap = (uint32_t)(n / regions + n / (8ull * regions) + 2048);
  const size_t sm = (size_t)TH * IT * 10 + (size_t)P * 8 + (RANK == 2 ? (size_t)(TH / 32) * P * 4 : (RANK == 3 ? (size_t)(TH / 32) * P * 2 : 0));
  auto kfn = k_part<TH, IT, MINB, RECS, RANK, PERSIST>;
  if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) != cudaSuccess) { cudaGetLastError(); printf("smem %zu too large\n", sm); return; }
  int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kfn, TH, sm));
  if (occ < 1) { printf("does not fit\n"); return; }
  cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, kfn));
  uint32_t grid = PERSIST ? 148u * occ : (uint32_t)tiles;
  if (grid > tiles) grid = (uint32_t)tiles;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e9f;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaMemset(cursor, 0, regions * 4));
    CK(cudaEventRecord(e0));
    kfn<<<grid, TH, sm>>>(in, n, dbits, pshift, P, pb, cap, tiles_per_coarse, cursor, out);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  std::vector<unsigned int> h(regions); CK(cudaMemcpy(h.data(), cursor, regions * 4, cudaMemcpyDeviceToHost));
  uint64_t s = 0; unsigned mx = 0; for (auto v : h) { s += v; mx = v > mx ? v : mx; }
  const double bytes = (RECS ? 16.0 : 20.0) * n;
  printf("%s TH=%4d IT=%2d minb=%d rank=%d persist=%d P=%4u regs=%3d occ=%d smem=%6zu : %7.3f ms (%.2f ms per 2^30, %.0f GB/s)%s\n",
         RECS ? "pass2" : "pass1", TH, IT, MINB, RANK, (int)PERSIST, P, fa.numRegs, occ, sm, best, best * (double)(1ull << 30) / n, bytes / best / 1e6,
         (s != n || mx > cap) ? "  !! count mismatch / overflow" : "");
  fflush(stdout);
}

int main(int argc, char** argv) {
  const int log2n = argc > 1 ? atoi(argv[1]) : 29;
  const uint64_t n = 1ull << log2n;
  uint32_t* tuples; uint2* out; unsigned int* cursor;
  CK(cudaMalloc(&tuples, n * 12)); CK(cudaMalloc(&out, (n + n / 4 + (256ull << 20)) * 8)); CK(cudaMalloc(&cursor, 65536 * 4));
  k_gen<<<148 * 8, 256>>>(tuples, n, (1u << 27) - 1u);
  CK(cudaDeviceSynchronize());
  printf("n = 2^%d\n", log2n);
  uint32_t* recs; CK(cudaMalloc(&recs, n * 8));
  k_gen_recs<<<148 * 8, 256>>>(recs, n, (1u << 27) - 1u);
  CK(cudaDeviceSynchronize());
  printf("--- pass 1 P=256: atomics vs tagged warp-private counters\n");
  run<512, 16, 2, false, 0, false>(tuples, n, 256, out, cursor);
  run<512, 16, 2, false, 3, false>(tuples, n, 256, out, cursor);
  run<512, 8, 3, false, 0, false>(tuples, n, 256, out, cursor);
  run<512, 8, 3, false, 3, false>(tuples, n, 256, out, cursor);
  run<512, 8, 4, false, 3, false>(tuples, n, 256, out, cursor);
  run<256, 16, 4, false, 3, false>(tuples, n, 256, out, cursor);
  run<256, 8, 6, false, 3, false>(tuples, n, 256, out, cursor);
  run<1024, 8, 2, false, 3, false>(tuples, n, 256, out, cursor);
  run<512, 8, 3, false, 3, true>(tuples, n, 256, out, cursor);
  run<512, 16, 2, false, 3, false>(tuples, n, 1024, out, cursor);
  run<512, 16, 2, false, 0, false>(tuples, n, 1024, out, cursor);
  run<512, 8, 3, false, 3, false>(tuples, n, 128, out, cursor);
  printf("--- pass 2\n");
  run<256, 8, 6, true, 0, false>(recs, n, 32, out, cursor);
  run<256, 8, 6, true, 3, false>(recs, n, 32, out, cursor);
  run<256, 8, 6, true, 3, false>(recs, n, 64, out, cursor);
  run<512, 8, 3, true, 3, false>(recs, n, 64, out, cursor);
  run<512, 8, 3, true, 3, false>(recs, n, 256, out, cursor);
  run<512, 8, 3, true, 0, false>(recs, n, 256, out, cursor);
  run<512, 16, 2, true, 3, false>(recs, n, 256, out, cursor);
  return 0;
}
