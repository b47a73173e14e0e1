// datagen.cu -- device-side input generation for scale runs.
//
// The reference generates its relations on the host with std::mt19937 (Experiment1::init, main_experiment1.cc:415-457;
// util/GenRandIntVec.cc:71-98,167-200,335-340; util/zipf_distribution.hh:48-58).  At 2^30 rows that is minutes of CPU
// time (the Zipf constructor alone loops over n for pmf_denom, zipf_distribution.hh:42-45) plus a 13 GB upload.  These
// kernels draw from the SAME distributions -- a uniform random permutation of [0, n) for R.k (std::shuffle of iota),
// uniform ints on [0, max) (std::uniform_int_distribution), Zipf(n, q) by Hoermann / Derflinger rejection-inversion exactly
// as zipf_distribution::operator() does, (v - 1 + shift) % max as genval_zipf does -- but NOT the libstdc++ bit stream:
// parity runs always use the reference's own generator (oracle/_ref, hostcpp/hj3d/datagen.hh).
//
// Every value is a pure function of (seed, global row id): a counter-based generator, so any rank of a multi-GPU run
// evaluates any slice of the one global relation, and a run on N GPUs joins exactly the data a run on one GPU joins.
// vec_permute after the Zipf / uniform draw (GenRandIntVec.cc:91-93,193-195) reorders i.i.d. values and is
// distribution-neutral; it is not reproduced.
#include "engine_internal.hh"

namespace {

__device__ __forceinline__ uint64_t splitmix(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__device__ __forceinline__ double u01(uint64_t bits) { return (double)(bits >> 11) * (1.0 / 9007199254740992.0); }   // [0,1)

// A bijection of [0, 2^bits): four rounds of (odd multiply, xor-shift) -- each step is invertible mod 2^bits.
__device__ __forceinline__ uint64_t permute_pow2(uint64_t i, uint32_t bits, uint64_t seed) {
  const uint64_t mask = bits >= 64 ? ~0ull : ((1ull << bits) - 1ull);
  const uint32_t sh = bits > 1 ? bits / 2 : 1;
  uint64_t x = i & mask;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const uint64_t k = splitmix(seed + r);
    x = (x * (k | 1ull)) & mask;
    x ^= x >> sh;
    x = (x + (k >> 17)) & mask;
  }
  return x;
}
// ... of [0, n) for any n: cycle walking over the next power of two (expected < 2 steps)
__device__ __forceinline__ uint64_t permute_n(uint64_t i, uint64_t n, uint32_t bits, uint64_t seed) {
  uint64_t x = permute_pow2(i, bits, seed);
  while (x >= n) x = permute_pow2(x, bits, seed);
  return x;
}

// zipf_distribution.hh:84-128, with the same helper functions
struct Zipf {
  double q, H_x1, H_n; uint64_t n;
  static constexpr double eps = 1e-8;
  __host__ __device__ static double expxm1bx(double x) { return fabs(x) > eps ? expm1(x) / x : (1.0 + x / 2.0 * (1.0 + x / 3.0 * (1.0 + x / 4.0))); }
  __host__ __device__ static double log1pxbx(double x) { return fabs(x) > eps ? log1p(x) / x : 1.0 - x * ((1 / 2.0) - x * ((1 / 3.0) - x * (1 / 4.0))); }
  __host__ __device__ double H(double x) const { const double lx = log(x); return expxm1bx((1.0 - q) * lx) * lx; }
  __host__ __device__ double H_inv(double x) const { double t = x * (1.0 - q); if (t < -1.0) t = -1.0; return exp(log1pxbx(t) * x); }
  __host__ __device__ double h(double x) const { return exp(-q * log(x)); }
};

enum { GEN_IOTA = 0, GEN_PERMUTATION = 1, GEN_UNIFORM = 2, GEN_ZIPF = 3, GEN_CONST = 4 };

__global__ void k_gen_column(uint8_t* base, uint32_t stride, uint32_t offset, uint64_t first_row, uint64_t n, int kind,
                             uint64_t vmax, uint32_t bits, Zipf z, uint64_t shift, uint64_t seed) {
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t i = first_row + j;
    uint64_t v;
    switch (kind) {
      case GEN_IOTA: v = i; break;
      case GEN_PERMUTATION: v = permute_n(i, vmax, bits, seed); break;
      case GEN_UNIFORM: {
        // unbiased: multiply-shift of 64 random bits (bias < 2^-32 for max < 2^32)
        v = __umul64hi(splitmix(seed ^ (i * 0xD1342543DE82EF95ull)), vmax);
      } break;
      case GEN_ZIPF: {
        uint64_t k = 1;
        for (uint32_t t = 0;; ++t) {                      // zipf_distribution::operator(): u ~ U(H_x1, H_n)
          const double u = z.H_x1 + (z.H_n - z.H_x1) * u01(splitmix(splitmix(seed + t) ^ (i * 0xD1342543DE82EF95ull)));
          const double x = z.H_inv(u);
          double r = rint(x);                             // std::round differs from rint only at exact .5 (measure zero)
          if (r < 1.0) r = 1.0;
          if (r > (double)z.n) r = (double)z.n;
          k = (uint64_t)r;
          if (u >= z.H((double)k + 0.5) - z.h((double)k)) break;
        }
        v = (k - 1 + shift) % vmax;                       // genval_zipf, GenRandIntVec.cc:290-293
      } break;
      default: v = shift; break;
    }
    *reinterpret_cast<uint32_t*>(base + j * stride + offset) = (uint32_t)v;
  }
}

}  // namespace

extern "C" int hj3d_gen_column_u32(hj3d_ctx* c, void* d_tuples, uint32_t tuple_bytes, uint32_t offset, uint64_t first_row, uint64_t n,
                                   int kind, uint64_t vmax, double zipf_q, uint64_t shift, uint64_t seed) {
  if (!c || (n && !d_tuples)) return fail(HJ3D_ERR_INVALID, "NULL argument");
  if (kind < 0 || kind > 4) return fail(HJ3D_ERR_INVALID, "unknown generator kind");
  if (tuple_bytes % 4 || offset % 4 || offset + 4 > tuple_bytes) return fail(HJ3D_ERR_INVALID, "attribute outside tuple / misaligned");
  if ((kind == GEN_PERMUTATION || kind == GEN_UNIFORM || kind == GEN_ZIPF) && (vmax == 0 || vmax > 0x100000000ull))
    return fail(HJ3D_ERR_INVALID, "value range must be 1 .. 2^32");
  if (kind == GEN_ZIPF && !(zipf_q >= 0.0)) return fail(HJ3D_ERR_INVALID, "zipf exponent must be >= 0");
  if (!n) return HJ3D_OK;
  CUDA_TRY(cudaSetDevice(c->device));
  Zipf z{};
  if (kind == GEN_ZIPF) {   // zipf_distribution.hh:32-41 (pmf_denom is only needed by pmf()/cdf(), not by the sampler)
    z.q = zipf_q; z.n = vmax; z.H_x1 = z.H(1.5) - 1.0; z.H_n = z.H((double)vmax + 0.5);
  }
  uint32_t bits = 0;
  while (bits < 63 && (1ull << bits) < vmax) ++bits;
  const uint32_t nb = (uint32_t)(n / 1024 + 1 > (uint64_t)c->sm_count * 16 ? (uint64_t)c->sm_count * 16 : n / 1024 + 1);
  k_gen_column<<<nb, 256, 0, c->stream>>>((uint8_t*)d_tuples, tuple_bytes, offset, first_row, n, kind, vmax, bits, z, shift, seed);
  ++c->launches;
  CUDA_TRY(cudaGetLastError());
  return HJ3D_OK;
}
