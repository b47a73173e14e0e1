// ubench.cu -- developer microbenchmarks behind the round-2 design decisions (not part of the product).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/ubench.out tools/ubench.cu
//   tools/ubench.out [log2n]
//
// (a) one-pass wide partition prototype: 12-byte tuples -> 8-byte (key,id) records of P bucket ranges, P up to 8192,
//     with ablations (no atomics / no global stores) that separate the shared-atomic, tile-sort and scattered-write costs;
// (b) fat-slice probe prototype: packed 6-byte-per-bucket table slice in shared memory, streaming records;
// (c) DSMEM random lookups (ld.shared::cluster) for cluster sizes 1/2/4.
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

// ------------------------------------------------------------------------------------------------ (a)
// MODE 0: full; 1: no atomics (rank = 0 -> collisions in the tile, timing only); 2: no global stores; 3: ballot ranking (P <= 64)
template <int TH, int IT, int MODE>
__global__ void __launch_bounds__(TH, (TH * IT * 10 <= 100 * 1024) ? 2 : 1)
k_part_wide(const uint32_t* __restrict__ tuples, uint64_t n, uint32_t dbits, uint32_t pshift, uint32_t P, uint32_t cap,
            unsigned int* __restrict__ cursor, uint2* __restrict__ out) {
  constexpr int TILE = TH * IT, WARPS = TH / 32;
  extern __shared__ __align__(16) unsigned char smem[];
  uint2*    tile = reinterpret_cast<uint2*>(smem);
  uint32_t* hist = reinterpret_cast<uint32_t*>(tile + TILE);
  uint32_t* dst  = hist + P;
  uint16_t* pid  = reinterpret_cast<uint16_t*>(dst + P);
  __shared__ uint32_t sm_scan[33];
  const uint32_t dmask = (1u << dbits) - 1u;
  for (uint64_t t0 = (uint64_t)blockIdx.x * TILE; t0 < n; t0 += (uint64_t)gridDim.x * TILE) {
    for (uint32_t p = threadIdx.x; p < P; p += TH) hist[p] = 0;
    uint32_t key[IT], pr[IT];
    const uint32_t tn = (n - t0) < (uint64_t)TILE ? (uint32_t)(n - t0) : (uint32_t)TILE;
#pragma unroll
    for (int j = 0; j < IT; ++j) {
      const uint32_t li = j * TH + threadIdx.x;
      key[j] = li < tn ? __ldg(tuples + 3 * (t0 + li) + 1) : 0u;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < IT; ++j) {
      const uint32_t li = j * TH + threadIdx.x;
      const uint32_t q = (murmur32(key[j]) & dmask) >> pshift;
      if (MODE == 1) { pr[j] = (q << 16) | (li & 1u); }
      else if (MODE == 3) {
        uint32_t peers = 0xffffffffu;
#pragma unroll
        for (int b = 0; b < 6; ++b) { const uint32_t m = __ballot_sync(0xffffffffu, (q >> b) & 1u); peers &= ((q >> b) & 1u) ? m : ~m; }
        const uint32_t leader = __ffs(peers) - 1;
        uint32_t base = 0;
        if (lane_id() == leader) base = atomicAdd(&hist[q], (uint32_t)__popc(peers));
        base = __shfl_sync(0xffffffffu, base, leader);
        pr[j] = (q << 16) | (base + __popc(peers & ((1u << lane_id()) - 1u)));
      } else {
        pr[j] = li < tn ? ((q << 16) | atomicAdd(&hist[q], 1u)) : 0xFFFFFFFFu;
      }
    }
    __syncthreads();
    {   // exclusive scan of hist + one reservation per (tile, partition)
      const uint32_t PER = (P + TH - 1) / TH;
      const uint32_t a = PER * threadIdx.x;
      uint32_t sum = 0;
      for (uint32_t k = 0; k < PER; ++k) sum += (a + k) < P ? hist[a + k] : 0u;
      const uint32_t w = threadIdx.x >> 5, l = lane_id();
      const uint32_t inc = warp_iscan(sum);
      if (l == 31) sm_scan[w] = inc;
      __syncthreads();
      if (w == 0) { const uint32_t t = l < (uint32_t)WARPS ? sm_scan[l] : 0u; const uint32_t ti = warp_iscan(t); sm_scan[l] = ti - t; }
      __syncthreads();
      uint32_t ex = inc - sum + sm_scan[w];
      for (uint32_t k = 0; k < PER; ++k) {
        if ((a + k) < P) {
          const uint32_t v = hist[a + k];
          const uint32_t g = v ? atomicAdd(&cursor[a + k], v) : 0u;
          hist[a + k] = ex;
          dst[a + k] = (a + k) * cap + g - ex;
          ex += v;
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < IT; ++j) {
      if (pr[j] == 0xFFFFFFFFu) continue;
      const uint32_t lp = pr[j] >> 16;
      uint32_t pos = hist[lp] + (pr[j] & 0xFFFFu);
      if (MODE == 1) pos = (j * TH + threadIdx.x);
      tile[pos] = make_uint2(key[j], (uint32_t)(t0 + j * TH + threadIdx.x));
      pid[pos] = (uint16_t)lp;
    }
    __syncthreads();
    if (MODE != 2) {
      for (uint32_t k = threadIdx.x; k < tn; k += TH) {
        const uint32_t lp = pid[k];
        const uint32_t idx = dst[lp] + k;
        if (MODE == 1 || idx < (lp + 1) * cap) out[MODE == 1 ? (t0 + k) : idx] = tile[k];
      }
    } else if (tile[threadIdx.x].x == 0xdeadbeefu) out[0] = tile[0];
    __syncthreads();
  }
}

template <int TH, int IT, int MODE>
float run_part(const uint32_t* tuples, uint64_t n, uint32_t dbits, uint32_t P, uint2* out, unsigned int* cursor, int grid_mult) {
  uint32_t pb = 0; while ((1u << pb) < P) ++pb;
  const uint32_t pshift = dbits - pb;
  const uint32_t cap = (uint32_t)(n / P + n / (16ull * P) + 4096);
  const size_t sm = (size_t)TH * IT * 10 + (size_t)P * 8;
  auto kfn = k_part_wide<TH, IT, MODE>;
  if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) != cudaSuccess) { cudaGetLastError(); return -1.f; }
  int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kfn, TH, sm));
  if (occ < 1) return -1.f;
  const uint64_t tiles = (n + (uint64_t)TH * IT - 1) / ((uint64_t)TH * IT);
  uint32_t grid = grid_mult ? 148u * occ * grid_mult : (uint32_t)tiles;
  if (grid > tiles) grid = (uint32_t)tiles;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e9f;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaMemset(cursor, 0, P * 4));
    CK(cudaEventRecord(e0));
    kfn<<<grid, TH, sm>>>(tuples, n, dbits, pshift, P, cap, cursor, out);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  if (MODE == 0) {   // sanity: cursors sum to n and none overflowed
    std::vector<unsigned int> h(P); CK(cudaMemcpy(h.data(), cursor, P * 4, cudaMemcpyDeviceToHost));
    uint64_t s = 0; unsigned mx = 0; for (auto v : h) { s += v; mx = v > mx ? v : mx; }
    if (s != n || mx > cap) printf("   !! sanity: sum %llu (n %llu) max %u cap %u\n", (unsigned long long)s, (unsigned long long)n, mx, cap);
  }
  printf("part_wide TH=%4d IT=%2d MODE=%d P=%5u occ=%d grid=%6u smem=%6zu : %7.3f ms  (%.2f ms per 2^30, %.0f GB/s of 20 B/tuple)\n",
         TH, IT, MODE, P, occ, grid, sm, best, best * (double)(1ull << 30) / n, 20.0 * n / best / 1e6);
  fflush(stdout);
  return best;
}

// ------------------------------------------------------------------------------------------------ (b)
// fat-slice probe: per partition a slice of NB buckets: soff[NB+1] u16 + packed[rows] u32 (quotient << rb | rowid).
// Records (key,id) stream in, (id,rowid) pairs stream out.  Slice content is synthetic (1 row per bucket).
template <int TH, int IT, bool BULK>
__global__ void __launch_bounds__(TH, (TH <= 256 ? 4 : (TH <= 512 ? 2 : 1)))
k_probe_fat(const uint2* __restrict__ recs, uint64_t n_per_part, uint32_t n_parts, uint32_t nb_log2, uint32_t dbits,
            const uint32_t* __restrict__ packed_g, uint2* __restrict__ out, unsigned long long* __restrict__ cursor, uint32_t slice_blocks_per_sm) {
  extern __shared__ __align__(16) unsigned char smem[];
  const uint32_t NB = 1u << nb_log2;
  uint32_t* packed = reinterpret_cast<uint32_t*>(smem);
  uint16_t* soff = reinterpret_cast<uint16_t*>(packed + NB);
  __shared__ uint32_t wsum[TH / 32];
  __shared__ unsigned long long sm_base;
  const uint32_t rb = dbits;                 // rowid bits; quotient = hash >> dbits
  const uint32_t dmask = (1u << dbits) - 1u, bmask = NB - 1u;
  unsigned long long matches = 0;
  for (uint32_t f = blockIdx.x; f < n_parts; f += gridDim.x) {
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < NB; i += TH) { packed[i] = __ldg(packed_g + (size_t)f * NB + i); }
    for (uint32_t i = threadIdx.x; i <= NB; i += TH) soff[i] = (uint16_t)i;
    __syncthreads();
    const uint2* in = recs + (size_t)f * n_per_part;
    constexpr int TILE = TH * IT;
    for (uint32_t t0 = 0; t0 < n_per_part; t0 += TILE) {
      uint2 r[IT];
#pragma unroll
      for (int j = 0; j < IT; ++j) { const uint32_t li = t0 + j * TH + threadIdx.x; r[j] = li < n_per_part ? in[li] : make_uint2(0, 0); }
      uint32_t hitmask = 0, wtot = 0, right[IT];
#pragma unroll
      for (int j = 0; j < IT; ++j) {
        const uint32_t h = murmur32(r[j].x);
        const uint32_t lb = (h & dmask) & bmask;
        const uint32_t o0 = soff[lb], o1 = soff[lb + 1];
        uint32_t hit = 0; right[j] = 0;
        for (uint32_t k = o0; k < o1; ++k) { const uint32_t pk = packed[k]; if (((pk >> rb) == (h >> dbits)) | (pk != 0u)) { hit = 1; right[j] = pk & dmask; break; } }
        hitmask |= hit << j;
        wtot += __popc(__ballot_sync(0xffffffffu, hit));
      }
      matches += __popc(hitmask);
      if (lane_id() == 0) wsum[threadIdx.x >> 5] = wtot;
      __syncthreads();
      uint32_t before = 0, total = 0;
#pragma unroll
      for (int w = 0; w < TH / 32; ++w) { const uint32_t v = wsum[w]; before += w < (int)(threadIdx.x >> 5) ? v : 0u; total += v; }
      if (threadIdx.x == 0) sm_base = atomicAdd(cursor, (unsigned long long)total);
      __syncthreads();
      unsigned long long pos = sm_base + before;
#pragma unroll
      for (int j = 0; j < IT; ++j) {
        const uint32_t hit = (hitmask >> j) & 1u;
        const uint32_t bal = __ballot_sync(0xffffffffu, hit);
        if (hit) out[pos + __popc(bal & ((1u << lane_id()) - 1u))] = make_uint2(r[j].y, right[j]);
        pos += __popc(bal);
      }
    }
  }
  if (matches == 0xffffffffffull) out[0] = make_uint2(1, 1);
}

template <int TH, int IT>
void run_probe_fat(const uint2* recs, uint64_t n, uint32_t nb_log2, uint2* out, const uint32_t* packed_g) {
  const uint32_t dbits = 27;
  const uint32_t NB = 1u << nb_log2;
  const uint32_t n_parts = (1u << dbits) >> nb_log2;
  const uint64_t n_per_part = n / n_parts;
  const size_t sm = (size_t)NB * 4 + (size_t)(NB + 2) * 2;
  auto kfn = k_probe_fat<TH, IT, false>;
  if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) != cudaSuccess) { cudaGetLastError(); printf("probe_fat smem %zu too large\n", sm); return; }
  int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kfn, TH, sm));
  if (occ < 1) { printf("probe_fat TH=%d nb=%u: does not fit\n", TH, nb_log2); return; }
  unsigned long long* cursor; CK(cudaMalloc(&cursor, 8));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e9f;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaMemset(cursor, 0, 8));
    CK(cudaEventRecord(e0));
    kfn<<<148 * occ, TH, sm>>>(recs, n_per_part, n_parts, nb_log2, dbits, packed_g, out, cursor, occ);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  printf("probe_fat TH=%4d IT=%d slice=2^%u buckets (%zu B) occ=%d parts=%u : %7.3f ms (%.2f ms per 2^30, %.0f GB/s of 16 B/probe + table)\n",
         TH, IT, nb_log2, sm, occ, n_parts, best, best * (double)(1ull << 30) / n, (16.0 * n + 6.0 * (1u << dbits)) / best / 1e6);
  fflush(stdout);
  CK(cudaFree(cursor));
}

// ------------------------------------------------------------------------------------------------ (c)
template <int CL>
__global__ void __launch_bounds__(1024, 1)
k_dsmem(uint32_t words_log2, uint32_t iters, unsigned long long* sink) {
  extern __shared__ __align__(16) unsigned char smem[];
  uint32_t* tab = reinterpret_cast<uint32_t*>(smem);
  const uint32_t W = 1u << words_log2;
  for (uint32_t i = threadIdx.x; i < W; i += blockDim.x) tab[i] = i * 2654435761u;
  uint32_t rank = 0;
  if (CL > 1) {
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    asm volatile("barrier.cluster.arrive.aligned; barrier.cluster.wait.aligned;" ::: "memory");
  } else __syncthreads();
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(tab);
  uint32_t x = murmur32(threadIdx.x + blockIdx.x * 1024u + 1u);
  uint32_t acc = 0;
  for (uint32_t it = 0; it < iters; ++it) {
    uint32_t v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      x = x * 1664525u + 1013904223u;
      const uint32_t h = x >> 4;
      const uint32_t addr = base + ((h & (W - 1)) << 2);
      if (CL > 1) {
        const uint32_t cta = (h >> 20) & (CL - 1);
        uint32_t ra;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(addr), "r"(cta));
        asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v[j]) : "r"(ra));
      } else {
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v[j]) : "r"(addr));
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc += v[j];
  }
  if (CL > 1) asm volatile("barrier.cluster.arrive.aligned; barrier.cluster.wait.aligned;" ::: "memory");
  if (acc == 0x12345u) sink[0] = acc + rank;
}

template <int CL>
void run_dsmem(unsigned long long* sink) {
  const uint32_t words_log2 = 15;   // 128 KB table per CTA
  const size_t sm = (size_t)4 << words_log2;
  auto kfn = k_dsmem<CL>;
  CK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.gridDim = dim3(CL * (148 / CL)); cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = sm; cfg.stream = 0; cfg.attrs = at; cfg.numAttrs = 1;
  const uint32_t iters = 2048;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e9f;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaEventRecord(e0));
    CK(cudaLaunchKernelEx(&cfg, kfn, words_log2, iters, sink));
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  const double lookups = (double)cfg.gridDim.x * 1024.0 * iters * 8.0;
  printf("dsmem CL=%d: %u CTAs, %.3f ms, %.1f G lookups/s chip (%.2f lookups/clk/SM at 1.9 GHz)\n", CL, cfg.gridDim.x, best,
         lookups / best / 1e6, lookups / best / 1e6 / cfg.gridDim.x / 1.9);
  fflush(stdout);
}

int main(int argc, char** argv) {
  const int log2n = argc > 1 ? atoi(argv[1]) : 28;
  const uint64_t n = 1ull << log2n;
  uint32_t* tuples; uint2* out; unsigned int* cursor; unsigned long long* sink;
  CK(cudaMalloc(&tuples, n * 12)); CK(cudaMalloc(&out, (n + n / 8 + (64ull << 20)) * 8)); CK(cudaMalloc(&cursor, 8192 * 4)); CK(cudaMalloc(&sink, 64));
  k_gen<<<148 * 8, 256>>>(tuples, n, (1u << 27) - 1u);
  CK(cudaDeviceSynchronize());
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s, %d SMs, n = 2^%d tuples\n", prop.name, prop.multiProcessorCount, log2n);
  // plain copy reference: 12 B read... use cudaMemcpy d2d of 20 B/tuple equivalent
  {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0)); CK(cudaMemcpyAsync(out, tuples, n * 10, cudaMemcpyDeviceToDevice)); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    CK(cudaEventRecord(e0)); CK(cudaMemcpyAsync(out, tuples, n * 10, cudaMemcpyDeviceToDevice)); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("memcpy d2d of %llu B (20 B/tuple traffic): %.3f ms = %.0f GB/s\n", (unsigned long long)(n * 10), ms, 20.0 * n / ms / 1e6);
  }
  printf("--- (a) wide partition, grid = tiles\n");
  for (uint32_t P : {256u, 1024u, 2048u, 4096u, 8192u}) run_part<1024, 16, 0>(tuples, n, 27, P, out, cursor, 0);
  for (uint32_t P : {256u, 1024u, 2048u, 4096u}) run_part<512, 16, 0>(tuples, n, 27, P, out, cursor, 0);
  for (uint32_t P : {256u, 1024u}) run_part<256, 16, 0>(tuples, n, 27, P, out, cursor, 0);
  for (uint32_t P : {1024u, 4096u}) run_part<1024, 8, 0>(tuples, n, 27, P, out, cursor, 0);
  printf("--- (a) ablations at TH=1024 IT=16\n");
  for (uint32_t P : {256u, 4096u}) { run_part<1024, 16, 1>(tuples, n, 27, P, out, cursor, 0); run_part<1024, 16, 2>(tuples, n, 27, P, out, cursor, 0); }
  printf("--- (a) ablations at TH=512 IT=16\n");
  for (uint32_t P : {256u, 4096u}) { run_part<512, 16, 1>(tuples, n, 27, P, out, cursor, 0); run_part<512, 16, 2>(tuples, n, 27, P, out, cursor, 0); }
  printf("--- (a) ballot ranking, P=64\n");
  run_part<256, 16, 3>(tuples, n, 27, 64, out, cursor, 0);
  run_part<512, 16, 3>(tuples, n, 27, 64, out, cursor, 0);
  run_part<256, 16, 0>(tuples, n, 27, 64, out, cursor, 0);
  printf("--- (a) persistent grids (TH=1024 IT=16)\n");
  for (uint32_t P : {1024u, 4096u}) { run_part<1024, 16, 0>(tuples, n, 27, P, out, cursor, 1); run_part<1024, 16, 0>(tuples, n, 27, P, out, cursor, 4); }
  printf("--- (b) fat-slice probe\n");
  {
    uint2* recs = reinterpret_cast<uint2*>(tuples);          // reuse: n*12 B >= n*8 B; content arbitrary
    uint32_t* packed_g; CK(cudaMalloc(&packed_g, (size_t)4 << 27)); CK(cudaMemset(packed_g, 0x5a, (size_t)4 << 27));
    run_probe_fat<1024, 4>(recs, n, 15, out, packed_g);
    run_probe_fat<1024, 8>(recs, n, 15, out, packed_g);
    run_probe_fat<512, 8>(recs, n, 15, out, packed_g);
    run_probe_fat<1024, 4>(recs, n, 14, out, packed_g);
    run_probe_fat<512, 4>(recs, n, 14, out, packed_g);
    run_probe_fat<512, 8>(recs, n, 14, out, packed_g);
    run_probe_fat<256, 4>(recs, n, 13, out, packed_g);
    run_probe_fat<512, 4>(recs, n, 13, out, packed_g);
    CK(cudaFree(packed_g));
  }
  printf("--- (c) DSMEM random lookups\n");
  run_dsmem<1>(sink); run_dsmem<2>(sink); run_dsmem<4>(sink);
  return 0;
}
