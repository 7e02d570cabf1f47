// Shared device/host helpers for the bayex_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/bayex_b200.h"

namespace bx {

// ---- error plumbing -----------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
#define BX_CHECK_ARG(cond, code, ...) \
  do {                                \
    if (!(cond)) {                    \
      bx::set_error(__VA_ARGS__);     \
      return (code);                  \
    }                                 \
  } while (0)
#define BX_CUDA(expr)                                                                   \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      bx::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return (int32_t)_e;                                                               \
    }                                                                                   \
  } while (0)
#define BX_LAUNCH_CHECK() BX_CUDA(cudaGetLastError())

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kSqrtLog2e = 1.2011224087864498f;  // sqrt(log2 e): coordinates are pre-scaled so exp(-r^2) = 2^(-r'^2)
constexpr float kJitter = 1e-6f;                   // gp.py:46

// hyp[] slots of bx_factor_t::d_hyp
enum { HYP_AMP = 0, HYP_NOISE = 1, HYP_YMEAN = 2, HYP_LOG2AMP = 3, HYP_NLML = 4, HYP_LOGDET = 5, HYP_QUAD = 6, HYP_STATUS = 7, HYP_T16SCALE = 8 };

__host__ __device__ inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

// ---- math ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float softplus_f(float x) {
  // logaddexp(x, 0) = max(x,0) + log1p(exp(-|x|))   (gp.py:26-27)
  return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x)));
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// jax.scipy.special.ndtr: erf branch for |x| < sqrt(1/2), erfc otherwise (acq.py:49,87 via norm.cdf)
__device__ __forceinline__ float ndtr_f(float x) {
  const float h = 0.70710678118654752440f;
  float w = x * h, z = fabsf(w), y;
  if (z < h) y = 1.f + erff(w);
  else if (w > 0.f) y = 2.f - erfcf(z);
  else y = erfcf(z);
  return 0.5f * y;
}
// jax.scipy.stats.norm.pdf = exp(-(log(2 pi) + x^2) / 2)
__device__ __forceinline__ float npdf_f(float x) { return expf((1.8378770664093453f + x * x) * -0.5f); }

// Closed-form acquisitions on (mu, sigma) - acq.py:45-50 (EI), 84-87 (PI), 120-121 (UCB), 155-156 (LCB).
__device__ __forceinline__ float acquisition(int acq_id, float mu, float sigma, float ystar, float p) {
  switch (acq_id) {
    case BX_ACQ_EI: {
      float a = mu - ystar - p;
      float z = a / (sigma + 1e-3f);
      return a * ndtr_f(z) + sigma * npdf_f(z);
    }
    case BX_ACQ_PI:
      return ndtr_f((mu - ystar - p) / sigma);
    case BX_ACQ_UCB:
      return mu + p * sigma;
    default:
      return mu - p * sigma;
  }
}

// ---- packed arg-max keys ------------------------------------------------------------------------------------------
// orderable(value) << 32 | ~index : unsigned max == (largest value, NaN largest of all, lowest index among ties),
// i.e. jnp.argmax semantics (optimizer.py:205).  -0.0 is canonicalised to +0.0 (they compare equal in jnp).
__host__ __device__ inline uint32_t orderable_u32(float v) {
  if (v != v) return 0xFFFFFFFFu;
  if (v == 0.f) v = 0.f;
  union { float f; uint32_t u; } c;
  c.f = v;
  return (c.u & 0x80000000u) ? ~c.u : (c.u | 0x80000000u);
}
__host__ __device__ inline float from_orderable_u32(uint32_t o) {
  union { float f; uint32_t u; } c;
  if (o == 0xFFFFFFFFu) { c.u = 0x7FC00000u; return c.f; }
  c.u = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
  return c.f;
}
__host__ __device__ inline uint64_t pack_best(float v, uint32_t idx) {
  return ((uint64_t)orderable_u32(v) << 32) | (uint64_t)(0xFFFFFFFFu - idx);
}

// ---- Threefry-2x32-20 (JAX PRNG, partitionable mode) ------------------------------------------------
__host__ __device__ inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
__host__ __device__ inline void threefry2x32(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t& o0, uint32_t& o1) {
  const uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
  uint32_t x0 = c0 + ks[0], x1 = c1 + ks[1];
#define BX_TF_ROUND(r) x0 += x1; x1 = rotl32(x1, r); x1 ^= x0;
  BX_TF_ROUND(13) BX_TF_ROUND(15) BX_TF_ROUND(26) BX_TF_ROUND(6)
  x0 += ks[1]; x1 += ks[2] + 1u;
  BX_TF_ROUND(17) BX_TF_ROUND(29) BX_TF_ROUND(16) BX_TF_ROUND(24)
  x0 += ks[2]; x1 += ks[0] + 2u;
  BX_TF_ROUND(13) BX_TF_ROUND(15) BX_TF_ROUND(26) BX_TF_ROUND(6)
  x0 += ks[0]; x1 += ks[1] + 3u;
  BX_TF_ROUND(17) BX_TF_ROUND(29) BX_TF_ROUND(16) BX_TF_ROUND(24)
  x0 += ks[1]; x1 += ks[2] + 4u;
  BX_TF_ROUND(13) BX_TF_ROUND(15) BX_TF_ROUND(26) BX_TF_ROUND(6)
  x0 += ks[2]; x1 += ks[0] + 5u;
#undef BX_TF_ROUND
  o0 = x0; o1 = x1;
}

// Per-dimension sampler state, prepared on the host (split keys are host-side Threefry calls).
struct DimSampler {
  int32_t kind;      // 0 real, 1 integer
  uint32_t k0, k1;   // real: the dimension key.  integer: first key of split(dim key)
  uint32_t k2, k3;   // integer: second key of split(dim key)
  float lo, scale;   // real: lo, (hi - lo) in float32;  clip bounds lo/hi
  float hi;
  int32_t ilo;       // integer: lower bound
  uint32_t span;     // integer: hi + 1 - lo
  uint32_t mult;     // integer: (2^16 mod span)^2 mod span
};
struct SamplerPack {
  int32_t d;
  DimSampler dim[BX_MAX_D];
};

// One coordinate of candidate `idx` (raw, unscaled units) - domain.py:74-75 / 130-131.
__device__ __forceinline__ float sample_coord(const DimSampler& s, uint64_t idx) {
  const uint32_t chi = (uint32_t)(idx >> 32), clo = (uint32_t)idx;
  uint32_t a, b;
  threefry2x32(s.k0, s.k1, chi, clo, a, b);
  const uint32_t bits = a ^ b;
  if (s.kind == 0) {
    float f = __uint_as_float((bits >> 9) | 0x3F800000u) - 1.0f;
    float v = fmaxf(s.lo, __fadd_rn(__fmul_rn(f, s.scale), s.lo));  // separately rounded: no FMA contraction
    return fminf(fmaxf(v, s.lo), s.hi);                              // Real.transform clip
  }
  threefry2x32(s.k2, s.k3, chi, clo, a, b);
  const uint32_t lob = a ^ b;
  const uint32_t off = ((bits % s.span) * s.mult + (lob % s.span)) % s.span;  // wrapping u32, as jax.random.randint
  return (float)(s.ilo + (int32_t)off);  // already integral and in range: round/clip of Integer.transform are no-ops
}

int32_t make_sampler(const uint32_t key[2], const bx_dim_t* dims, int32_t d, SamplerPack* out, bool split_keys = true);

// ---- reductions ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
    v = w > v ? w : v;
  }
  return v;
}

}  // namespace bx
