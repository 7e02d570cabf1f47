// GP fit path on the device: Gram build, Cholesky, T = L^-1, alpha, NLML, analytic gradient, clip + AdamW.
// Replaces gp.py:41-57 (fit half of gaussian_process), gp.py:78-87 (neg_log_likelihood), gp.py:90-110 (posterior_fit).
#include <cuda_fp16.h>
#include <stdlib.h>

#include <vector>

#include "internal.h"

namespace bx {

constexpr float kPriorNoise = -8.0f, kPriorAmp = 1.0f, kPriorLs = 1.0f;  // gp.py:83
constexpr float kLog2Pi = 1.8378770664093453f;

// raw hypers -> hyp[amp, noise, ymean(untouched), log2 amp], scale[k] = sqrt(log2 e)/softplus(ls_raw[k])   (gp.py:41,45)
__global__ void hyp_prepare_kernel(const float* __restrict__ raw, int d, int d4, float* hyp, float* scale) {
  const int t = threadIdx.x;
  if (t == 0) {
    const float amp = softplus_f(raw[1]);
    hyp[HYP_NOISE] = softplus_f(raw[0]);
    hyp[HYP_AMP] = amp;
    hyp[HYP_LOG2AMP] = log2f(amp);
  }
  for (int k = t; k < d4; k += blockDim.x) scale[k] = (k < d) ? kSqrtLog2e / softplus_f(raw[2 + k]) : 0.f;
}

// U[i][k] = X[i][k] * scale[k] for i < n, zero padding elsewhere.
__global__ void scale_obs_kernel(const float* __restrict__ X, int n, int d, int np, int d4, const float* __restrict__ scale,
                                 float* U) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= np * d4) return;
  const int i = e / d4, k = e % d4;
  U[e] = (i < n && k < d) ? X[(size_t)i * d + k] * scale[k] : 0.f;
}

// ymean = mean(y[:n]) (gp.py:43), yc = y - ymean (gp.py:44), zero padded.  Single block, fixed reduction order.
__global__ void __launch_bounds__(1024) center_y_kernel(const float* __restrict__ y, int n, int np, float* yc, float* hyp) {
  __shared__ float red[32];
  __shared__ float mean_s;
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += y[i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) { mean_s = v / (float)n; hyp[HYP_YMEAN] = mean_s; }
  }
  __syncthreads();
  const float m = mean_s;
  for (int i = threadIdx.x; i < np; i += blockDim.x) yc[i] = (i < n) ? y[i] - m : 0.f;
}

// K = amp * exp(-|u_i - u_j|^2) + (noise + 1e-6) I  on the valid block, identity on the padding (gp.py:46).
// 32x32 output tile per block; observation rows staged in shared memory; coalesced row stores.
__global__ void __launch_bounds__(256) gram_kernel(const float* __restrict__ U, int n, int np, int d4,
                                                   const float* __restrict__ hyp, float* K, int ld) {
  extern __shared__ float sm[];
  float* Ui = sm;                  // [32][d4+1]
  float* Uj = sm + 32 * (d4 + 1);  // [32][d4+1]
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32, tid = threadIdx.x;
  for (int e = tid; e < 32 * d4; e += 256) {
    const int r = e / d4, k = e % d4;
    Ui[r * (d4 + 1) + k] = U[(size_t)(i0 + r) * d4 + k];
    Uj[r * (d4 + 1) + k] = U[(size_t)(j0 + r) * d4 + k];
  }
  __syncthreads();
  const float amp = hyp[HYP_AMP], diag = hyp[HYP_NOISE] + kJitter;
  const int c = tid & 31;
  for (int r = tid >> 5; r < 32; r += 8) {
    const int i = i0 + r, j = j0 + c;
    float s = 0.f;
    for (int k = 0; k < d4; ++k) {
      const float df = Ui[r * (d4 + 1) + k] - Uj[c * (d4 + 1) + k];
      s = fmaf(df, df, s);
    }
    float v;
    if (i < n && j < n) v = amp * exp2f(-s) + (i == j ? diag : 0.f);
    else v = (i == j) ? 1.f : 0.f;
    K[(size_t)i * ld + j] = v;
  }
}

// quad = |v|^2 (v = T yc), logdet = sum of per-block parts, NLML (gp.py:52-57 on the compacted rows).
__global__ void __launch_bounds__(256) nlml_kernel(const float* __restrict__ v, int np, const float* __restrict__ logdet_part,
                                                   int nb, int n, float* hyp, const int* status) {
  __shared__ float red[8];
  float acc = 0.f;
  for (int i = threadIdx.x; i < np; i += 256) acc = fmaf(v[i], v[i], acc);
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float quad = 0.f;
    for (int w = 0; w < 8; ++w) quad += red[w];
    float logdet = 0.f;
    for (int b = 0; b < nb; ++b) logdet += logdet_part[b];
    const float amp = hyp[HYP_AMP], la = logf(amp);
    float nl = 0.5f * quad + logdet;            // gp.py:52-53  (masked rows are compacted away: :54 vanishes)
    nl += ((float)n * 0.5f) * kLog2Pi;          // gp.py:55
    nl += -0.5f * kLog2Pi - la - la * la;       // gp.py:56
    hyp[HYP_QUAD] = quad; hyp[HYP_LOGDET] = logdet; hyp[HYP_NLML] = nl;
    hyp[HYP_STATUS] = (float)status[0];
  }
}

// Padded rows/columns of T <- 0 (so padded observations contribute nothing to V) and max|T| for the fp16 scale.
__global__ void zero_pad_absmax_kernel(float* T, int n, int np, unsigned int* absmax_bits) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  float t = 0.f;
  if (e < (size_t)np * np) {
    const int i = (int)(e / np), j = (int)(e % np);
    t = T[e];
    if (i >= n || j >= n) { t = 0.f; T[e] = 0.f; }
  }
  unsigned int b = __float_as_uint(fabsf(t));  // non-negative floats order like their bit patterns
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) b = max(b, __shfl_xor_sync(0xffffffffu, b, o));
  if ((threadIdx.x & 31) == 0 && b) atomicMax(absmax_bits, b);
}
// fp16 head/tail of T * 2^b, 2^b chosen so max|T| 2^b is in [2^13, 2^14].
__global__ void split_T_kernel(const float* __restrict__ T, int np, __half* H16, __half* L16,
                               const unsigned int* absmax_bits, float* hyp) {
  const float amax = __uint_as_float(*absmax_bits);
  const float sc = (amax > 0.f) ? exp2f(floorf(log2f(16384.f / amax))) : 1.f;
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e == 0) hyp[HYP_T16SCALE] = sc;
  if (e >= (size_t)np * np) return;
  const float t = T[e];
  if (H16) {
    const float ts = t * sc;
    const __half h = __float2half_rn(ts);
    H16[e] = h;
    L16[e] = __float2half_rn(ts - __half2float(h));
  }
}

// Gradient contraction over the lower triangle of W = K^-1 - alpha alpha^T  (SURVEY 8.A3):
//   part[blk] = { tr(W), sum W_ij C_ij, sum W_ij C_ij (u_ik - u_jk)^2  (k < d) }   (off-diagonal pairs counted twice)
__global__ void __launch_bounds__(256) grad_contract_kernel(const float* __restrict__ Kinv, int ld, const float* __restrict__ alpha,
                                                            const float* __restrict__ U, int n, int d, int d4, float* part) {
  extern __shared__ float sm[];
  float* Ui = sm;
  float* Uj = sm + 32 * (d4 + 1);
  __shared__ float red[8];
  // linear lower-triangular tile index -> (bi, bj), bj <= bi
  const int t = blockIdx.x;
  int bi = (int)((sqrtf(8.f * (float)t + 1.f) - 1.f) * 0.5f);
  while ((bi + 1) * (bi + 2) / 2 <= t) ++bi;
  while (bi * (bi + 1) / 2 > t) --bi;
  const int bj = t - bi * (bi + 1) / 2;
  const int i0 = bi * 32, j0 = bj * 32, tid = threadIdx.x;
  for (int e = tid; e < 32 * d4; e += 256) {
    const int r = e / d4, k = e % d4;
    Ui[r * (d4 + 1) + k] = U[(size_t)(i0 + r) * d4 + k];
    Uj[r * (d4 + 1) + k] = U[(size_t)(j0 + r) * d4 + k];
  }
  __syncthreads();
  const int c = tid & 31, r0 = tid >> 5;
  float wc[4], tr = 0.f, samp = 0.f;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int r = r0 + q * 8, i = i0 + r, j = j0 + c;
    wc[q] = 0.f;
    if (i < n && j < n && j <= i) {
      const float w = Kinv[(size_t)i * ld + j] - alpha[i] * alpha[j];
      float s = 0.f;
      for (int k = 0; k < d4; ++k) {
        const float df = Ui[r * (d4 + 1) + k] - Uj[c * (d4 + 1) + k];
        s = fmaf(df, df, s);
      }
      if (i == j) tr += w;
      wc[q] = w * exp2f(-s) * (i == j ? 1.f : 2.f);
      samp += wc[q];
    }
  }
  float* out = part + (size_t)blockIdx.x * (d + 2);
  for (int slot = 0; slot < d + 2; ++slot) {
    float v;
    if (slot == 0) v = tr;
    else if (slot == 1) v = samp;
    else {
      const int k = slot - 2;
      v = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int r = r0 + q * 8;
        const float df = Ui[r * (d4 + 1) + k] - Uj[c * (d4 + 1) + k];
        v = fmaf(wc[q], df * df, v);
      }
    }
    v = warp_sum(v);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    if (tid == 0) {
      float tot = 0.f;
      for (int w = 0; w < 8; ++w) tot += red[w];
      out[slot] = tot;
    }
  }
}

// Sum the per-tile partials (fixed order), apply the chain rule; gbuf = [nlml, loss, dloss/draw (d+2), dnlml/draw (d+2)].
__global__ void __launch_bounds__(256) grad_finalize_kernel(const float* __restrict__ part, int ntiles, int d,
                                                            const float* __restrict__ raw, const float* __restrict__ hyp,
                                                            float* gbuf) {
  __shared__ float red[8];
  __shared__ float sums[BX_MAX_D + 2];
  for (int slot = 0; slot < d + 2; ++slot) {
    float acc = 0.f;
    for (int t = threadIdx.x; t < ntiles; t += 256) acc += part[(size_t)t * (d + 2) + slot];
    acc = warp_sum(acc);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tot = 0.f;
      for (int w = 0; w < 8; ++w) tot += red[w];
      sums[slot] = tot;
    }
  }
  __syncthreads();
  const float amp = hyp[HYP_AMP];
  float prior = 0.f;
  for (int p = threadIdx.x; p < d + 2; p += 256) {
    float g;  // d NLML / d raw_p
    if (p == 0) g = 0.5f * sums[0] * sigmoid_f(raw[0]);
    else if (p == 1) {
      const float la = logf(amp);
      g = (0.5f * sums[1] - 1.f / amp - 2.f * la / amp) * sigmoid_f(raw[1]);
    } else {
      const float ls = softplus_f(raw[p]);
      // d/dls exp(-(dx/ls)^2) = exp(.) 2 (dx/ls)^2 / ls ;  (u_i-u_j)^2 = log2e (dx/ls)^2
      g = 0.5f * amp * sums[p] * (2.f / (kLog2e * ls)) * sigmoid_f(raw[p]);
    }
    const float pr = (p == 0) ? kPriorNoise : (p == 1 ? kPriorAmp : kPriorLs);
    gbuf[2 + (d + 2) + p] = g;
    gbuf[2 + p] = -g + (raw[p] - pr);  // gp.py:78-87: loss = -(NLML - 1/2 sum (raw - prior)^2)
  }
  if (threadIdx.x == 0) {
    for (int p = 0; p < d + 2; ++p) {
      const float pr = (p == 0) ? kPriorNoise : (p == 1 ? kPriorAmp : kPriorLs);
      prior += (raw[p] - pr) * (raw[p] - pr);
    }
    gbuf[0] = hyp[HYP_NLML];
    gbuf[1] = -(hyp[HYP_NLML] - 0.5f * prior);
  }
}

// clip_by_global_norm(10) + adamw(lr) on the d+2 raw hypers (gp.py:99,105-106).  adam = [m | v | t].
// The moment weights are 0.1f / 0.001f: optax writes (1 - decay) with `decay` a Python float, so the subtraction happens in
// double (0.09999999999999998, 0.0010000000000000009) and is then rounded to float32 when it meets the float32 gradient -
// that is float32(0.1) and float32(0.001), not the float32 differences 1.f - 0.9f / 1.f - 0.999f (= 0.00100005).  The
// oracle (oracle/gp.py:adamw_step) does the same; tests/test_gpu_parity.py::test_adamw_kernel_step_by_step holds this
// kernel to it at 1e-6 per step.
__global__ void adamw_kernel(float* raw, const float* __restrict__ gbuf, int d, float lr, float* adam, float* trace,
                             int step) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int P = d + 2;
  float* m = adam;
  float* v = adam + P;
  float* tcount = adam + 2 * P;
  if (trace)
    for (int p = 0; p < P; ++p) trace[(size_t)step * P + p] = raw[p];
  float gn = 0.f;
  for (int p = 0; p < P; ++p) gn = fmaf(gbuf[2 + p], gbuf[2 + p], gn);
  gn = sqrtf(gn);
  const float t = tcount[0] + 1.f;
  tcount[0] = t;
  const float c1 = 1.f - powf(0.9f, t), c2 = 1.f - powf(0.999f, t);
  for (int p = 0; p < P; ++p) {
    float g = gbuf[2 + p];
    if (!(gn < 10.f)) g = g / gn * 10.f;
    m[p] = 0.9f * m[p] + 0.1f * g;
    v[p] = 0.999f * v[p] + 0.001f * g * g;
    const float u = (m[p] / c1) / (sqrtf(v[p] / c2) + 1e-8f) + 1e-4f * raw[p];
    raw[p] = raw[p] - lr * u;
  }
}


// ------------------------------------------------------------------------------------------------------------------
// Small problems (n <= 128): the WHOLE posterior_fit loop (gp.py:90-110) in one kernel launch, one CTA, everything in
// shared memory: softplus -> scaled observations -> Gram -> Cholesky -> T = L^-1 -> alpha, NLML -> K^-1 = T^T T ->
// gradient contraction -> clip + AdamW, `steps` times.  This is the README / test-suite regime (n = 3 .. ~100), where
// the multi-kernel path is pure launch latency.
// ------------------------------------------------------------------------------------------------------------------
constexpr int FS_THREADS = 256;

__device__ __forceinline__ float fs_block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < FS_THREADS / 32; ++w) t += red[w];
  return t;
}

__global__ void __launch_bounds__(FS_THREADS, 1)
fit_small_kernel(const float* __restrict__ X, const float* __restrict__ y, int n, int d, float* raw_g, float lr, int steps,
                 float* trace, float* gbuf_out, int grad_only) {
  extern __shared__ float sm[];
  const int ld = n + 1, d4 = round_up(d, 4), P = d + 2, tid = threadIdx.x;
  float* A = sm;                    // [n][ld]  K -> L -> K^-1
  float* B = A + n * ld;            // [n][ld]  T = L^-1 -> W*C weights
  float* U = B + n * ld;            // [n][d4]  scaled observations
  float* yc = U + n * d4;           // [n]
  float* vv = yc + n;               // [n]
  float* al = vv + n;               // [n]
  float* raw = al + n;              // [P]
  float* grad = raw + P;            // [P]
  float* scale = grad + P;          // [d4]
  float* mom = scale + d4;          // [2P] Adam moments
  __shared__ float red[FS_THREADS / 32];
  __shared__ float sc[8];           // amp, noise+jitter, ymean, quad, logdet, tr, samp, nlml
  for (int p = tid; p < P; p += FS_THREADS) { raw[p] = raw_g[p]; mom[p] = 0.f; mom[P + p] = 0.f; }
  float ysum = 0.f;
  for (int i = tid; i < n; i += FS_THREADS) ysum += y[i];
  ysum = fs_block_sum(ysum, red);
  const float ymean = ysum / (float)n;
  for (int i = tid; i < n; i += FS_THREADS) yc[i] = y[i] - ymean;
  __syncthreads();

  for (int step = 0; step <= steps; ++step) {
    const bool last = (step == steps);  // the extra pass only evaluates (bx_nlml_grad) - no update
    if (last && !grad_only) break;
    // hypers (gp.py:41) and scaled observations (gp.py:45)
    if (tid == 0) { sc[0] = softplus_f(raw[1]); sc[1] = softplus_f(raw[0]) + kJitter; }
    for (int k = tid; k < d4; k += FS_THREADS) scale[k] = (k < d) ? kSqrtLog2e / softplus_f(raw[2 + k]) : 0.f;
    __syncthreads();
    for (int e = tid; e < n * d4; e += FS_THREADS) {
      const int i = e / d4, k = e % d4;
      U[e] = (k < d) ? X[(size_t)i * d + k] * scale[k] : 0.f;
    }
    __syncthreads();
    const float amp = sc[0];
    // Gram, lower triangle (gp.py:46)
    for (int e = tid; e < n * n; e += FS_THREADS) {
      const int i = e / n, j = e % n;
      if (j > i) continue;
      float s = 0.f;
      for (int k = 0; k < d4; ++k) { const float df = U[i * d4 + k] - U[j * d4 + k]; s = fmaf(df, df, s); }
      A[i * ld + j] = amp * exp2f(-s) + (i == j ? sc[1] : 0.f);
    }
    __syncthreads();
    // Cholesky, right-looking (gp.py:48)
    for (int j = 0; j < n; ++j) {
      if (tid == 0) A[j * ld + j] = sqrtf(A[j * ld + j]);
      __syncthreads();
      const float inv = 1.f / A[j * ld + j];
      for (int i = j + 1 + tid; i < n; i += FS_THREADS) A[i * ld + j] *= inv;
      __syncthreads();
      const int rem = n - 1 - j;
      for (int e = tid; e < rem * rem; e += FS_THREADS) {
        const int r = j + 1 + e / rem, c = j + 1 + e % rem;
        if (c <= r) A[r * ld + c] = fmaf(-A[r * ld + j], A[c * ld + j], A[r * ld + c]);
      }
      __syncthreads();
    }
    // T = L^-1: one column per thread, forward substitution
    for (int c = tid; c < n; c += FS_THREADS) {
      for (int i = 0; i < c; ++i) B[i * ld + c] = 0.f;
      B[c * ld + c] = 1.f / A[c * ld + c];
      for (int i = c + 1; i < n; ++i) {
        float acc = 0.f;
        for (int j = c; j < i; ++j) acc = fmaf(A[i * ld + j], B[j * ld + c], acc);
        B[i * ld + c] = -acc / A[i * ld + i];
      }
    }
    float ldp = 0.f;
    for (int i = tid; i < n; i += FS_THREADS) ldp += logf(A[i * ld + i]);
    const float logdet = fs_block_sum(ldp, red);  // (also the barrier after the T columns)
    // v = T yc, alpha = T^T v (gp.py:49), quad = |v|^2
    for (int i = tid; i < n; i += FS_THREADS) {
      float acc = 0.f;
      for (int j = 0; j <= i; ++j) acc = fmaf(B[i * ld + j], yc[j], acc);
      vv[i] = acc;
    }
    __syncthreads();
    float qp = 0.f;
    for (int j = tid; j < n; j += FS_THREADS) {
      float acc = 0.f;
      for (int i = j; i < n; ++i) acc = fmaf(B[i * ld + j], vv[i], acc);
      al[j] = acc;
      qp = fmaf(vv[j], vv[j], qp);
    }
    const float quad = fs_block_sum(qp, red);
    const float la = logf(amp);
    const float nlml = 0.5f * quad + logdet + ((float)n * 0.5f) * kLog2Pi + (-0.5f * kLog2Pi - la - la * la);  // gp.py:52-57
    // K^-1 = T^T T (lower) into A
    for (int e = tid; e < n * n; e += FS_THREADS) {
      const int i = e / n, j = e % n;
      if (j > i) continue;
      float acc = 0.f;
      for (int k = i; k < n; ++k) acc = fmaf(B[k * ld + i], B[k * ld + j], acc);
      A[i * ld + j] = acc;
    }
    __syncthreads();
    // weights W_ij C_ij (x2 off the diagonal) into B; trace and amplitude sums
    float trp = 0.f, sap = 0.f;
    for (int e = tid; e < n * n; e += FS_THREADS) {
      const int i = e / n, j = e % n;
      if (j > i) continue;
      const float w = A[i * ld + j] - al[i] * al[j];
      float s = 0.f;
      for (int k = 0; k < d4; ++k) { const float df = U[i * d4 + k] - U[j * d4 + k]; s = fmaf(df, df, s); }
      const float wc = w * exp2f(-s) * (i == j ? 1.f : 2.f);
      B[i * ld + j] = wc;
      if (i == j) trp += w;
      sap += wc;
    }
    const float tr = fs_block_sum(trp, red);
    const float samp = fs_block_sum(sap, red);
    float prior = 0.f;
    for (int p = 0; p < P; ++p) {
      float g;  // d NLML / d raw_p
      if (p == 0) g = 0.5f * tr * sigmoid_f(raw[0]);
      else if (p == 1) g = (0.5f * samp - 1.f / amp - 2.f * la / amp) * sigmoid_f(raw[1]);
      else {
        const int k = p - 2;
        float sk = 0.f;
        for (int e = tid; e < n * n; e += FS_THREADS) {
          const int i = e / n, j = e % n;
          if (j >= i) continue;
          const float df = U[i * d4 + k] - U[j * d4 + k];
          sk = fmaf(B[i * ld + j], df * df, sk);
        }
        sk = fs_block_sum(sk, red);
        const float ls = softplus_f(raw[p]);
        g = 0.5f * amp * sk * (2.f / (kLog2e * ls)) * sigmoid_f(raw[p]);
      }
      const float pr = (p == 0) ? kPriorNoise : (p == 1 ? kPriorAmp : kPriorLs);
      if (tid == 0) {
        grad[p] = -g + (raw[p] - pr);  // gp.py:78-87
        if (gbuf_out && last) { gbuf_out[2 + p] = grad[p]; gbuf_out[2 + P + p] = g; }
      }
      prior += (raw[p] - pr) * (raw[p] - pr);
    }
    if (tid == 0 && gbuf_out && last) { gbuf_out[0] = nlml; gbuf_out[1] = -(nlml - 0.5f * prior); }
    __syncthreads();
    if (last) break;
    if (tid == 0) {  // clip_by_global_norm(10) + adamw (gp.py:99,105-106)
      if (trace) for (int p = 0; p < P; ++p) trace[(size_t)step * P + p] = raw[p];
      float gn = 0.f;
      for (int p = 0; p < P; ++p) gn = fmaf(grad[p], grad[p], gn);
      gn = sqrtf(gn);
      const float t = (float)(step + 1), c1 = 1.f - powf(0.9f, t), c2 = 1.f - powf(0.999f, t);
      for (int p = 0; p < P; ++p) {
        float g = grad[p];
        if (!(gn < 10.f)) g = g / gn * 10.f;
        mom[p] = 0.9f * mom[p] + 0.1f * g;
        mom[P + p] = 0.999f * mom[P + p] + 0.001f * g * g;
        const float u = (mom[p] / c1) / (sqrtf(mom[P + p] / c2) + 1e-8f) + 1e-4f * raw[p];
        raw[p] = raw[p] - lr * u;
      }
    }
    __syncthreads();
  }
  for (int p = tid; p < P; p += FS_THREADS) raw_g[p] = raw[p];
}

// ---- n <= 128: the whole fit loop in one CTA with the matrix in REGISTERS ------------------------------------------------------
// fit_small_kernel above keeps K in shared memory and walks a textbook Cholesky with three block barriers per column, a
// triangular inverse on one thread per column and T^T T by dot products: it crosses over with the multi-kernel path at n = 56
// (0.087 ms per step).  This kernel holds the lower 32 x 32 tiles of the padded matrix in the registers of 256 threads
// (thread (w, tx) owns rows 32 a + w + 8 r, r < 4, of column 32 b + tx of every tile a >= b; diagonal tiles are kept whole) and turns K
// into -K^-1 IN PLACE with the symmetric sweep operator - per pivot k: A_ij -= A_ik A_kj / p, A_ik <- A_ik / p,
// A_kk <- -1 / p with p = A_kk - one uniform rank-1 update per pivot, ONE block barrier per pivot (the pivot column
// travels through a double-buffered shared-memory vector), no separate factorisation / inverse / product phases.  The
// pivots are the squares of the Cholesky diagonal (gp.py:48: log det = 1/2 sum log p_k; a non-positive pivot becomes NaN
// exactly where cholesky() would produce one); alpha = K^-1 yc, the NLML (gp.py:52-57), the gradient contraction
//  and clip + AdamW (gp.py:99-106) follow in the same launch for all `steps`.  n^3 / 2 FMAs per step instead
// of n^3 / 3 + n^3 / 6 + n^3 / 6, but every one of them on all threads; fp32 error of K^-1 at the level of the Cholesky
// route (both ~ cond(K) eps).  Reductions run in a fixed order: the result does not depend on the launch.
constexpr int SW_RPT = 4;                  // rows of every tile a thread owns
constexpr int SW_NW = 32 / SW_RPT;         // warps: thread (w, tx) owns rows 32 a + w + SW_NW r, column 32 b + tx
constexpr int SW_THREADS = 32 * SW_NW;     // 256: with one row per thread (1024 threads) the pivot loop was issue-bound on its
                                           // per-warp overhead (column loads, reciprocal, fix-ups, publish): 0.44 us per pivot
constexpr int SW_MAX_NT = 4;  // n <= 128
__host__ __device__ constexpr int sw_e(int a, int b) { return a * (a + 1) / 2 + b; }

__device__ __forceinline__ float sw_block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < SW_NW; ++w) t += red[w];
  return t;
}

template <int NT>
__global__ void __launch_bounds__(SW_THREADS, 1)
fit_sweep_kernel(const float* __restrict__ X, const float* __restrict__ y, int n, int d, float* raw_g, float lr, int steps,
                 float* trace, float* gbuf_out, int grad_only) {
  constexpr int NP = 32 * NT, NE = NT * (NT + 1) / 2, R = SW_RPT, NW = SW_NW;
  extern __shared__ float sm[];
  const int d4 = round_up(d, 4), P = d + 2, tid = threadIdx.x, w = tid >> 5, tx = tid & 31;
  float* Ut = sm;                 // [d4][NP] scaled observations, dimension-major
  float* yc = Ut + d4 * NP;       // [NP]
  float* al = yc + NP;            // [NP]
  float* cb = al + NP;            // [2][NP] pivot column, double-buffered
  float* part = cb + 2 * NP;      // [NW][NP] cross-warp partial sums (also: the pivots)
  float* raw = part + NW * NP;    // [P]
  float* grad = raw + P;          // [P]
  float* dsum = grad + P;         // [d4] per-dimension gradient sums
  float* scale = dsum + d4;       // [d4]
  float* mom = scale + d4;        // [2P] Adam moments
  __shared__ float red[NW];
  __shared__ float red4[4][NW];
  __shared__ float sc[8];         // amp, noise+jitter
  for (int p = tid; p < P; p += SW_THREADS) { raw[p] = raw_g[p]; mom[p] = 0.f; mom[P + p] = 0.f; }
  float ysp = 0.f;
  for (int i = tid; i < n; i += SW_THREADS) ysp += y[i];
  const float ymean = sw_block_sum(ysp, red) / (float)n;
  for (int i = tid; i < NP; i += SW_THREADS) yc[i] = (i < n) ? y[i] - ymean : 0.f;
  __syncthreads();

  for (int step = 0; step <= steps; ++step) {
    const bool last = (step == steps);  // the extra pass only evaluates (bx_nlml_grad) - no update
    if (last && !grad_only) break;
    // hypers (gp.py:41) and scaled observations (gp.py:45)
    if (tid == 0) { sc[0] = softplus_f(raw[1]); sc[1] = softplus_f(raw[0]) + kJitter; }
    for (int k = tid; k < d4; k += SW_THREADS) scale[k] = (k < d) ? kSqrtLog2e / softplus_f(raw[2 + k]) : 0.f;
    __syncthreads();
    for (int e = tid; e < d4 * NP; e += SW_THREADS) {
      const int k = e / NP, i = e - k * NP;
      Ut[e] = (i < n && k < d) ? X[(size_t)i * d + k] * scale[k] : 0.f;
    }
    __syncthreads();
    const float amp = sc[0];
    // Gram (gp.py:46) into registers; identity on the padding
    float A[NE][R];
#pragma unroll
    for (int e = 0; e < NE; ++e)
#pragma unroll
      for (int r = 0; r < R; ++r) A[e][r] = 0.f;
    for (int k = 0; k < d4; ++k) {
      float ui[NT][R], uj[NT];
#pragma unroll
      for (int a = 0; a < NT; ++a) {
        uj[a] = Ut[k * NP + 32 * a + tx];
#pragma unroll
        for (int r = 0; r < R; ++r) ui[a][r] = Ut[k * NP + 32 * a + w + NW * r];
      }
#pragma unroll
      for (int a = 0; a < NT; ++a)
#pragma unroll
        for (int b = 0; b <= a; ++b)
#pragma unroll
          for (int r = 0; r < R; ++r) { const float df = ui[a][r] - uj[b]; A[sw_e(a, b)][r] = fmaf(df, df, A[sw_e(a, b)][r]); }
    }
#pragma unroll
    for (int a = 0; a < NT; ++a)
#pragma unroll
      for (int b = 0; b <= a; ++b)
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int i = 32 * a + w + NW * r, j = 32 * b + tx;
          A[sw_e(a, b)][r] = (i < n && j < n) ? amp * exp2f(-A[sw_e(a, b)][r]) + (i == j ? sc[1] : 0.f) : (i == j ? 1.f : 0.f);
        }
    // sweep every valid pivot: A -> -K^-1
#pragma unroll
    for (int kb = 0; kb < NT; ++kb) {
      const int kmax = min(32, n - 32 * kb);
      for (int kx = 0; kx < kmax; ++kx) {
        const int k = 32 * kb + kx;
        const int wk = kx % NW, rk = kx / NW;  // row k belongs to warp wk, register slot rk
        float* buf = cb + (k & 1) * NP;
        if (tx == kx) {  // column k: the diagonal tile's rows and the tiles below it
#pragma unroll
          for (int a = kb; a < NT; ++a)
#pragma unroll
            for (int r = 0; r < R; ++r) buf[32 * a + w + NW * r] = A[sw_e(a, kb)][r];
        }
        if (w == wk) {  // row k (= column k above the diagonal tile): warp-uniform
#pragma unroll
          for (int b = 0; b < kb; ++b) {
            float v = A[sw_e(kb, b)][0];
#pragma unroll
            for (int r = 1; r < R; ++r) v = (rk == r) ? A[sw_e(kb, b)][r] : v;
            buf[32 * b + tx] = v;
          }
        }
        __syncthreads();
        float p = buf[k];
        if (!(p > 0.f)) p = CUDART_NAN_F;  // gp.py:48: cholesky() of a non-PD matrix gives NaN, and so does everything after
        const float ip = __frcp_rn(p);     // correctly rounded, without the division's slow path: it heads every pivot's chain
        if (tid == 0) part[k] = p;         // the logarithms are taken after the loop, one pivot per thread
        float ci[NT][R], cjp[NT];
#pragma unroll
        for (int a = 0; a < NT; ++a) {
          cjp[a] = buf[32 * a + tx] * ip;
#pragma unroll
          for (int r = 0; r < R; ++r) ci[a][r] = buf[32 * a + w + NW * r];
        }
#pragma unroll
        for (int a = 0; a < NT; ++a)
#pragma unroll
          for (int b = 0; b <= a; ++b)
#pragma unroll
            for (int r = 0; r < R; ++r) A[sw_e(a, b)][r] = fmaf(-ci[a][r], cjp[b], A[sw_e(a, b)][r]);
        if (tx == kx) {  // column k below the diagonal tile
#pragma unroll
          for (int a = kb + 1; a < NT; ++a)
#pragma unroll
            for (int r = 0; r < R; ++r) A[sw_e(a, kb)][r] = ci[a][r] * ip;
        }
        if (w == wk) {   // row k left of the diagonal tile
#pragma unroll
          for (int b = 0; b < kb; ++b)
#pragma unroll
            for (int r = 0; r < R; ++r)
              if (rk == r) A[sw_e(kb, b)][r] = cjp[b];
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {  // the diagonal tile
          const bool ik = (w == wk) && (rk == r), jk = (tx == kx);
          if (ik && jk) A[sw_e(kb, kb)][r] = -ip;
          else if (ik) A[sw_e(kb, kb)][r] = cjp[kb];
          else if (jk) A[sw_e(kb, kb)][r] = ci[kb][r] * ip;
        }
      }
    }
    __syncthreads();
    float ldp = 0.f;
    for (int i = tid; i < n; i += SW_THREADS) ldp += logf(part[i]);
    const float logdet = 0.5f * sw_block_sum(ldp, red);  // (ends with every thread past its reads of part)
    // alpha = K^-1 yc (gp.py:49) with K^-1 = -A: row sums inside the warp, transposed contributions of the off-diagonal
    // tiles through per-warp partials added in warp order
#pragma unroll
    for (int a = 0; a < NT; ++a)
#pragma unroll
      for (int r = 0; r < R; ++r) {
        float t = 0.f;
#pragma unroll
        for (int b = 0; b <= a; ++b) t = fmaf(-A[sw_e(a, b)][r], yc[32 * b + tx], t);
        t = warp_sum(t);
        if (tx == 0) al[32 * a + w + NW * r] = t;
      }
#pragma unroll
    for (int b = 0; b < NT; ++b) {
      float t = 0.f;
#pragma unroll
      for (int a = b + 1; a < NT; ++a)
#pragma unroll
        for (int r = 0; r < R; ++r) t = fmaf(-A[sw_e(a, b)][r], yc[32 * a + w + NW * r], t);
      part[w * NP + 32 * b + tx] = t;
    }
    __syncthreads();
    for (int j = tid; j < NP; j += SW_THREADS) {
      float t = al[j];
#pragma unroll
      for (int q = 0; q < NW; ++q) t += part[q * NP + j];
      al[j] = t;
    }
    __syncthreads();
    float qp = 0.f;
    for (int i = tid; i < NP; i += SW_THREADS) qp = fmaf(yc[i], al[i], qp);
    const float quad = sw_block_sum(qp, red);
    const float la = logf(amp);
    const float nlml = 0.5f * quad + logdet + ((float)n * 0.5f) * kLog2Pi + (-0.5f * kLog2Pi - la - la * la);  // gp.py:52-57
    // weights W_ij C_ij (x2 off the diagonal, lower triangle only) into A; trace and amplitude sums
    float S[NE][R];
#pragma unroll
    for (int e = 0; e < NE; ++e)
#pragma unroll
      for (int r = 0; r < R; ++r) S[e][r] = 0.f;
    for (int k = 0; k < d4; ++k) {
      float ui[NT][R], uj[NT];
#pragma unroll
      for (int a = 0; a < NT; ++a) {
        uj[a] = Ut[k * NP + 32 * a + tx];
#pragma unroll
        for (int r = 0; r < R; ++r) ui[a][r] = Ut[k * NP + 32 * a + w + NW * r];
      }
#pragma unroll
      for (int a = 0; a < NT; ++a)
#pragma unroll
        for (int b = 0; b <= a; ++b)
#pragma unroll
          for (int r = 0; r < R; ++r) { const float df = ui[a][r] - uj[b]; S[sw_e(a, b)][r] = fmaf(df, df, S[sw_e(a, b)][r]); }
    }
    float trp = 0.f, sap = 0.f;
#pragma unroll
    for (int a = 0; a < NT; ++a)
#pragma unroll
      for (int b = 0; b <= a; ++b)
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int li = w + NW * r, i = 32 * a + li, j = 32 * b + tx;
          const float wgt = -A[sw_e(a, b)][r] - al[i] * al[j];
          float f = 2.f;
          if (a == b) f = (tx < li) ? 2.f : (tx == li ? 1.f : 0.f);  // a diagonal tile holds both triangles: count the lower one
          if (!(i < n && j < n)) f = 0.f;
          const float wc = (f != 0.f) ? wgt * exp2f(-S[sw_e(a, b)][r]) * f : 0.f;
          A[sw_e(a, b)][r] = wc;
          if (a == b && tx == li && i < n) trp += wgt;
          sap += wc;
        }
    const float tr = sw_block_sum(trp, red);
    const float samp = sw_block_sum(sap, red);
    // per-dimension sums, four dimensions per reduction pass
    for (int k0 = 0; k0 < d; k0 += 4) {
      float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int k = k0 + q;  // < d4
        float ui[NT][R], uj[NT];
#pragma unroll
        for (int a = 0; a < NT; ++a) {
          uj[a] = Ut[k * NP + 32 * a + tx];
#pragma unroll
          for (int r = 0; r < R; ++r) ui[a][r] = Ut[k * NP + 32 * a + w + NW * r];
        }
#pragma unroll
        for (int a = 0; a < NT; ++a)
#pragma unroll
          for (int b = 0; b <= a; ++b)
#pragma unroll
            for (int r = 0; r < R; ++r) { const float df = ui[a][r] - uj[b]; s4[q] = fmaf(A[sw_e(a, b)][r], df * df, s4[q]); }
        s4[q] = warp_sum(s4[q]);
      }
      __syncthreads();
      if (tx == 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q) red4[q][w] = s4[q];
      }
      __syncthreads();
      if (tid < 4) {
        float t = 0.f;
#pragma unroll
        for (int q = 0; q < NW; ++q) t += red4[tid][q];
        dsum[k0 + tid] = t;
      }
    }
    __syncthreads();
    if (tid == 0) {
      float prior = 0.f;
      for (int p = 0; p < P; ++p) {
        float g;  // d NLML / d raw_p
        if (p == 0) g = 0.5f * tr * sigmoid_f(raw[0]);
        else if (p == 1) g = (0.5f * samp - 1.f / amp - 2.f * la / amp) * sigmoid_f(raw[1]);
        else {
          const float ls = softplus_f(raw[p]);
          g = 0.5f * amp * dsum[p - 2] * (2.f / (kLog2e * ls)) * sigmoid_f(raw[p]);
        }
        const float pr = (p == 0) ? kPriorNoise : (p == 1 ? kPriorAmp : kPriorLs);
        grad[p] = -g + (raw[p] - pr);  // gp.py:78-87
        if (gbuf_out && last) { gbuf_out[2 + p] = grad[p]; gbuf_out[2 + P + p] = g; }
        prior += (raw[p] - pr) * (raw[p] - pr);
      }
      if (gbuf_out && last) { gbuf_out[0] = nlml; gbuf_out[1] = -(nlml - 0.5f * prior); }
      if (!last) {  // clip_by_global_norm(10) + adamw (gp.py:99,105-106)
        if (trace) for (int p = 0; p < P; ++p) trace[(size_t)step * P + p] = raw[p];
        float gn = 0.f;
        for (int p = 0; p < P; ++p) gn = fmaf(grad[p], grad[p], gn);
        gn = sqrtf(gn);
        const float t = (float)(step + 1), c1 = 1.f - powf(0.9f, t), c2 = 1.f - powf(0.999f, t);
        for (int p = 0; p < P; ++p) {
          float g = grad[p];
          if (!(gn < 10.f)) g = g / gn * 10.f;
          mom[p] = 0.9f * mom[p] + 0.1f * g;
          mom[P + p] = 0.999f * mom[P + p] + 0.001f * g * g;
          const float u = (mom[p] / c1) / (sqrtf(mom[P + p] / c2) + 1e-8f) + 1e-4f * raw[p];
          raw[p] = raw[p] - lr * u;
        }
      }
    }
    __syncthreads();
    if (last) break;
  }
  for (int p = tid; p < P; p += SW_THREADS) raw_g[p] = raw[p];
}

static size_t fit_sweep_smem(int n, int d) {
  const int np = round_up(n, 32), d4 = round_up(d, 4), P = d + 2;
  return sizeof(float) * ((size_t)d4 * np + 4 * (size_t)np + SW_NW * (size_t)np + 4 * P + 2 * d4);
}
static bool fit_sweep_ok(int n) {
  static const bool off = getenv("BX_NO_FIT_SWEEP") != nullptr;
  return !off && n <= 32 * SW_MAX_NT;
}

static size_t fit_small_smem(int n, int d) {
  const int d4 = round_up(d, 4), P = d + 2;
  return sizeof(float) * ((size_t)2 * n * (n + 1) + (size_t)n * d4 + 3 * n + 2 * P + d4 + 2 * P);
}
// measured on B200: fused 0.027 ms/step at n=24, 0.065 at 48, 0.087 at 56, 0.110 at 64, 0.165 at 80; the multi-kernel
// path (one 128-row tile, CUDA graph) takes 0.085 ms/step for every n <= 128 -> cross-over at n = 56
constexpr int FS_DISPATCH_N = 56;
static bool fit_small_ok(int n, int d) {
  if (fit_sweep_ok(n)) return true;  // register-resident sweep kernel: n <= 128
  return n <= FS_DISPATCH_N && fit_small_smem(n, d) <= 200 * 1024;
}

static int32_t launch_fit_small(cudaStream_t s, const float* d_X, const float* d_y, int n, int d, float* d_raw, float lr,
                                int steps, float* d_trace, float* d_gbuf, int grad_only) {
  if (fit_sweep_ok(n)) {
    const size_t ssm = fit_sweep_smem(n, d);
    const int nt = (n + 31) / 32;
#define BX_SWEEP_CASE(NT)                                                                                                  \
  case NT:                                                                                                                 \
    BX_CUDA(cudaFuncSetAttribute(fit_sweep_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssm));            \
    fit_sweep_kernel<NT><<<1, SW_THREADS, ssm, s>>>(d_X, d_y, n, d, d_raw, lr, steps, d_trace, d_gbuf, grad_only);         \
    break;
    switch (nt) {
      BX_SWEEP_CASE(1) BX_SWEEP_CASE(2) BX_SWEEP_CASE(3) BX_SWEEP_CASE(4)
    }
#undef BX_SWEEP_CASE
    BX_LAUNCH_CHECK();
    return 0;
  }
  const size_t smem = fit_small_smem(n, d);
  BX_CUDA(cudaFuncSetAttribute(fit_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  fit_small_kernel<<<1, FS_THREADS, smem, s>>>(d_X, d_y, n, d, d_raw, lr, steps, d_trace, d_gbuf, grad_only);
  BX_LAUNCH_CHECK();
  return 0;
}

FitWs carve_fit_ws(void* base, int n, int d) {
  FitWs w{};
  w.np = round_up(n, BX_TILE);
  w.d4 = round_up(d, 4);
  w.nb = w.np / NB;
  const int t32 = w.np / 32;
  w.ntiles = t32 * (t32 + 1) / 2;
  size_t off = 0;
  auto take = [&](size_t nfloat) {
    float* p = base ? reinterpret_cast<float*>(reinterpret_cast<char*>(base) + off) : nullptr;
    off += (nfloat * sizeof(float) + 255) / 256 * 256;
    return p;
  };
  w.A = take((size_t)w.np * w.np);
  w.TMP = take((size_t)w.np * w.np);
  w.Dinv = take((size_t)w.nb * NB * NB);
  w.logdet_part = take(w.nb);
  w.yc = take(w.np);
  w.v = take(w.np);
  w.alpha = take(w.np);
  w.U = take((size_t)w.np * w.d4);
  w.scale = take(w.d4);
  w.hyp = take(8);
  w.part = take((size_t)w.ntiles * (d + 2));
  w.gbuf = take(2 + 2 * (d + 2));
  w.adam = take(2 * (d + 2) + 1);
  w.status = reinterpret_cast<int*>(take(2));  // [0] Cholesky status, [1] max|T| bits
  w.bytes = off;
  return w;
}

// One step of iterative refinement for alpha (posterior build only): r = yc - K alpha, alpha += T^T (T r).
// alpha = T^T (T yc) goes through the explicit inverse factor, which is not a backward-stable solve: its error is the
// largest part of the posterior mean's distance to the fp64 reference (relative error of alpha 1.8e-5 at n = 2048,
// noise 0.018; 1.3e-6 after this step in a NumPy model of the path).  K is rebuilt from U entry by entry exactly as
// gram_kernel builds it (the factorised matrix itself was overwritten by its factor) and the row sums are accumulated
// in fp64 - the residual is a difference of O(|y|) quantities that agree to ~5 digits.
__global__ void __launch_bounds__(256) gram_residual_kernel(const float* __restrict__ U, int n, int d4,
                                                            const float* __restrict__ hyp, const float* __restrict__ yc,
                                                            const float* __restrict__ alpha, float* __restrict__ r) {
  extern __shared__ float sm[];
  float* Ui = sm;                  // [32][d4+1]
  float* Uj = sm + 32 * (d4 + 1);  // [32][d4+1]
  float* aj = Uj + 32 * (d4 + 1);  // [32]
  const int i0 = blockIdx.x * 32, tid = threadIdx.x, c = tid & 31, w = tid >> 5;
  for (int e = tid; e < 32 * d4; e += 256) Ui[(e / d4) * (d4 + 1) + e % d4] = U[(size_t)(i0 + e / d4) * d4 + e % d4];
  const float amp = hyp[HYP_AMP], diag = hyp[HYP_NOISE] + kJitter;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (int j0 = 0; j0 < n; j0 += 32) {
    __syncthreads();
    for (int e = tid; e < 32 * d4; e += 256) Uj[(e / d4) * (d4 + 1) + e % d4] = U[(size_t)(j0 + e / d4) * d4 + e % d4];
    if (tid < 32) aj[tid] = (j0 + tid < n) ? alpha[j0 + tid] : 0.f;
    __syncthreads();
    const int j = j0 + c;
    const double a = (double)aj[c];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int rr = w + 8 * q, i = i0 + rr;
      float s = 0.f;
      for (int k = 0; k < d4; ++k) {
        const float df = Ui[rr * (d4 + 1) + k] - Uj[c * (d4 + 1) + k];
        s = fmaf(df, df, s);
      }
      const float v = (i < n && j < n) ? amp * exp2f(-s) + (i == j ? diag : 0.f) : 0.f;
      acc[q] = fma((double)v, a, acc[q]);
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    double t = acc[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    const int i = i0 + w + 8 * q;
    if (c == 0) r[i] = (i < n) ? (float)((double)yc[i] - t) : 0.f;
  }
}
__global__ void add_inplace_kernel(float* __restrict__ x, const float* __restrict__ dx, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] += dx[i];
}

// alpha = T^T (T yc), quad = |T yc|^2, NLML; with `refine` (posterior build / append: the alpha the scoring kernels read, the
// fit loop's gradient does not need it) one step of iterative refinement on alpha.  tmp: 3 np floats.
static int32_t launch_alpha_tail(cudaStream_t a, const float* T, const float* U, float* alpha, float* hyp, int n, int np, int d4,
                                 const float* yc, float* v, float* tmp, const float* logdet_part, int nb, const int* status,
                                 bool refine) {
  int32_t rc = trmv_lower(a, T, np, np, yc, v);
  if (rc) return rc;
  rc = trmv_lower_t(a, T, np, np, v, alpha);
  if (rc) return rc;
  nlml_kernel<<<1, 256, 0, a>>>(v, np, logdet_part, nb, n, hyp, status);
  BX_LAUNCH_CHECK();
  static const bool refine_alpha = [] { const char* e = getenv("BX_ALPHA_REFINE"); return !e || atoi(e) != 0; }();
  if (refine && refine_alpha) {
    float *res = tmp, *tv = tmp + np, *dx = tmp + 2 * (size_t)np;  // (the fit workspace's TMP is free once T is complete)
    const size_t rsm = (2 * 32 * (d4 + 1) + 32) * sizeof(float);
    gram_residual_kernel<<<np / 32, 256, rsm, a>>>(U, n, d4, hyp, yc, alpha, res);
    BX_LAUNCH_CHECK();
    rc = trmv_lower(a, T, np, np, res, tv);
    if (rc) return rc;
    rc = trmv_lower_t(a, T, np, np, tv, dx);
    if (rc) return rc;
    add_inplace_kernel<<<(np + 255) / 256, 256, 0, a>>>(alpha, dx, np);
    BX_LAUNCH_CHECK();
  }
  return 0;
}

// ---- appending one observation to a built factor at unchanged hypers (optimizer.py:261-269) ---------------
// The Gram matrix grows by one bordered row, K' = [K k; k^T kappa], so its factor does too:
//   L' = [L 0; l^T lambda],  l = L^-1 k = T k,  lambda^2 = kappa - |l|^2,   T' = L'^-1 = [T 0; -(l^T T) / lambda   1 / lambda]
// - two triangular matrix-vector products (O(n^2)) instead of Gram + Cholesky + inverse (O(n^3)).
//
// U[n] = x_new * scale, kvec[j] = amp 2^(-|U_j - U_n|^2) for j < n (the arithmetic of gram_kernel, entry by entry), 0 on the padding.
__global__ void __launch_bounds__(256) append_cross_kernel(const float* __restrict__ x_new, int n, int d, int np, int d4,
                                                           const float* __restrict__ scale, const float* __restrict__ hyp,
                                                           float* U, float* __restrict__ kvec) {
  __shared__ float un[BX_MAX_D + 4];
  for (int k = threadIdx.x; k < d4; k += 256) un[k] = (k < d) ? x_new[k] * scale[k] : 0.f;
  __syncthreads();
  if (blockIdx.x == 0)
    for (int k = threadIdx.x; k < d4; k += 256) U[(size_t)n * d4 + k] = un[k];  // row n is read by nobody in this launch
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= np) return;
  float v = 0.f;
  if (j < n) {
    float s = 0.f;
    for (int k = 0; k < d4; ++k) {
      const float df = un[k] - U[(size_t)j * d4 + k];
      s = fmaf(df, df, s);
    }
    v = hyp[HYP_AMP] * exp2f(-s);
  }
  kvec[j] = v;
}
// lambda, row n of T, log det.  One block; |l|^2 is accumulated in fp64 in a fixed order (lambda^2 = kappa - |l|^2 cancels
// to the posterior variance at the new point plus the noise, exactly as the last pivot of a Cholesky does).
__global__ void __launch_bounds__(1024) append_row_kernel(float* T, int n, int np, const float* __restrict__ l,
                                                          const float* __restrict__ wv, const float* hyp, float* logdet_out,
                                                          int* status) {
  __shared__ double red[32];
  __shared__ float lam_s;
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += 1024) acc = fma((double)l[i], (double)l[i], acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double ss = 0.0;
    for (int w = 0; w < 32; ++w) ss += red[w];
    const float kappa = hyp[HYP_AMP] + (hyp[HYP_NOISE] + kJitter);  // the diagonal entry gram_kernel writes
    const double lam2 = (double)kappa - ss;
    const bool bad = !(lam2 > 0.0) || !isfinite(lam2);
    const float lam = (float)sqrt(lam2);  // NaN on a bad pivot: propagates like the reference's cholesky (gp.py:48)
    lam_s = lam;
    logdet_out[0] = hyp[HYP_LOGDET] + logf(lam);
    status[0] = (bad || hyp[HYP_STATUS] != 0.f) ? 1 : 0;
  }
  __syncthreads();
  const float inv = 1.f / lam_s;
  float* row = T + (size_t)n * np;
  for (int j = threadIdx.x; j < np; j += 1024) row[j] = (j < n) ? -wv[j] * inv : (j == n ? inv : 0.f);
}


int32_t launch_factor(cudaStream_t s, const float* d_X, const float* d_y, int n, int d, const float* d_raw, FitWs& w,
                      float* T_out, float* U_out, float* alpha_out, float* scale_out, float* hyp_out) {
  float* T = T_out ? T_out : w.A;
  float* U = U_out ? U_out : w.U;
  float* alpha = alpha_out ? alpha_out : w.alpha;
  float* scale = scale_out ? scale_out : w.scale;
  float* hyp = hyp_out ? hyp_out : w.hyp;
  const int np = w.np, d4 = w.d4;
  BX_CUDA(cudaMemsetAsync(w.status, 0, sizeof(int), s));
  hyp_prepare_kernel<<<1, 64, 0, s>>>(d_raw, d, d4, hyp, scale);
  BX_LAUNCH_CHECK();
  scale_obs_kernel<<<(np * d4 + 255) / 256, 256, 0, s>>>(d_X, n, d, np, d4, scale, U);
  BX_LAUNCH_CHECK();
  center_y_kernel<<<1, 1024, 0, s>>>(d_y, n, np, w.yc, hyp);
  BX_LAUNCH_CHECK();
  const size_t gsm = 2 * 32 * (d4 + 1) * sizeof(float);
  gram_kernel<<<dim3(np / 32, np / 32), 256, gsm, s>>>(U, n, np, d4, hyp, T, np);
  BX_LAUNCH_CHECK();
  int32_t rc = cholesky_inplace(s, T, np, np, w.Dinv, w.logdet_part, w.status, lookahead_wanted(np) ? w.la : nullptr, w.TMP);
  if (rc) return rc;
  rc = tri_inverse_inplace(s, T, np, np, w.Dinv, w.TMP);
  if (rc) return rc;
  // alpha = T^T (T yc) and the NLML are three latency-bound launches that need only T: with a helper stream they run
  // beside whatever the caller queues on `s` next (K^-1 = T^T T, the operand splits); join_alpha() orders `s` after them.
  cudaStream_t a = s;
  if (w.la) {
    a = w.la->chain;
    BX_CUDA(cudaEventRecord(w.la->ev_bulk, s));
    BX_CUDA(cudaStreamWaitEvent(a, w.la->ev_bulk, 0));
  }
  rc = launch_alpha_tail(a, T, U, alpha, hyp, n, np, d4, w.yc, w.v, w.TMP, w.logdet_part, w.nb, w.status, alpha_out != nullptr);
  if (rc) return rc;
  if (w.la) BX_CUDA(cudaEventRecord(w.la->ev_chain, a));
  return 0;
}

// Orders `s` after the alpha / NLML launches of launch_factor (no-op without a helper stream).
static int32_t join_alpha(cudaStream_t s, const FitWs& w) {
  if (w.la) BX_CUDA(cudaStreamWaitEvent(s, w.la->ev_chain, 0));
  return 0;
}

// One NLML + gradient evaluation at d_raw; leaves gbuf filled.  Uses ws.A for T and ws.TMP for K^-1.
static int32_t launch_grad(cudaStream_t s, const float* d_X, const float* d_y, int n, int d, const float* d_raw, FitWs& w) {
  int32_t rc = launch_factor(s, d_X, d_y, n, d, d_raw, w, nullptr, nullptr, nullptr, nullptr, nullptr);
  if (rc) return rc;
  GemmArgs g{};  // K^-1 = T^T T on the lower 64x64 tiles
  g.A = w.A; g.lda = w.np; g.B = w.A; g.ldb = w.np; g.C = w.TMP; g.ldc = w.np;
  g.M = g.N = g.K = w.np; g.alpha = 1.f; g.beta = 0.f; g.kmode = KM_TN_BOTH; g.lower_only = 1;
  rc = gemm(s, true, false, g);
  if (rc) return rc;
  if ((rc = join_alpha(s, w))) return rc;
  const size_t gsm = 2 * 32 * (w.d4 + 1) * sizeof(float);
  grad_contract_kernel<<<w.ntiles, 256, gsm, s>>>(w.TMP, w.np, w.alpha, w.U, n, d, w.d4, w.part);
  BX_LAUNCH_CHECK();
  grad_finalize_kernel<<<1, 256, 0, s>>>(w.part, w.ntiles, d, d_raw, w.hyp, w.gbuf);
  BX_LAUNCH_CHECK();
  return 0;
}

}  // namespace bx

using namespace bx;

extern "C" size_t bx_fit_workspace_bytes(int32_t n, int32_t d) {
  if (n <= 0 || d <= 0 || d > BX_MAX_D) return 0;
  return carve_fit_ws(nullptr, n, d).bytes;
}

static int32_t check_fit_args(const void* X, const void* y, int32_t n, int32_t d, const void* raw, void* ws, size_t ws_bytes) {
  BX_CHECK_ARG(X && y && raw && ws, BX_E_ARG, "null pointer argument");
  BX_CHECK_ARG(n > 0 && d > 0, BX_E_ARG, "n and d must be positive (n=%d d=%d)", n, d);
  BX_CHECK_ARG(d <= BX_MAX_D, BX_E_RANGE, "d=%d exceeds BX_MAX_D=%d", d, BX_MAX_D);
  BX_CHECK_ARG(ws_bytes >= bx_fit_workspace_bytes(n, d), BX_E_WORKSPACE, "workspace too small: %zu < %zu", ws_bytes,
               bx_fit_workspace_bytes(n, d));
  return 0;
}

extern "C" int32_t bx_nlml_grad(const float* d_X, const float* d_y, int32_t n, int32_t d, const float* d_raw, float* d_out,
                                void* d_ws, size_t ws_bytes, void* stream) {
  int32_t rc = check_fit_args(d_X, d_y, n, d, d_raw, d_ws, ws_bytes);
  if (rc) return rc;
  BX_CHECK_ARG(d_out, BX_E_ARG, "null output");
  cudaStream_t s = (cudaStream_t)stream;
  FitWs w = carve_fit_ws(d_ws, n, d);
  if (fit_small_ok(n, d) && !getenv("BX_NO_FIT_SMALL"))  // one-kernel evaluation (raw is left untouched: steps = 0)
    return launch_fit_small(s, d_X, d_y, n, d, const_cast<float*>(d_raw), 0.f, 0, nullptr, d_out, 1);
  if ((rc = cholesky_prepare())) return rc;
  Lookahead la;
  if (lookahead_wanted(w.np)) { if ((rc = la.init())) return rc; w.la = &la; }
  rc = launch_grad(s, d_X, d_y, n, d, d_raw, w);
  if (rc) return rc;
  BX_CUDA(cudaMemcpyAsync(d_out, w.gbuf, sizeof(float) * (2 + 2 * (d + 2)), cudaMemcpyDeviceToDevice, s));
  return 0;
}

extern "C" int32_t bx_fit_adamw(const float* d_X, const float* d_y, int32_t n, int32_t d, float* d_raw, float lr,
                                int32_t steps, float* d_trace, void* d_ws, size_t ws_bytes, int32_t use_graph,
                                void* stream) {
  int32_t rc = check_fit_args(d_X, d_y, n, d, d_raw, d_ws, ws_bytes);
  if (rc) return rc;
  BX_CHECK_ARG(steps >= 0, BX_E_ARG, "steps must be >= 0");
  cudaStream_t s = (cudaStream_t)stream;
  FitWs w = carve_fit_ws(d_ws, n, d);
  if (steps == 0) return 0;
  if (fit_small_ok(n, d) && !getenv("BX_NO_FIT_SMALL"))  // whole loop in one launch
    return launch_fit_small(s, d_X, d_y, n, d, d_raw, lr, steps, d_trace, nullptr, 0);
  BX_CUDA(cudaMemsetAsync(w.adam, 0, sizeof(float) * (2 * (d + 2) + 1), s));  // fresh moments per call (gp.py:100)
  if ((rc = cholesky_prepare())) return rc;
  Lookahead la;  // helper stream: its cost is amortised over the loop, so always
  if (!getenv("BX_NO_HELPER_STREAM")) { if ((rc = la.init())) return rc; w.la = &la; }
  if (use_graph && !d_trace) {
    // One step is captured once and replayed `steps` times: the loop never returns to the host.
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    BX_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    rc = launch_grad(s, d_X, d_y, n, d, d_raw, w);
    if (rc == 0) {
      adamw_kernel<<<1, 32, 0, s>>>(d_raw, w.gbuf, d, lr, w.adam, nullptr, 0);
      cudaError_t le = cudaGetLastError();
      if (le != cudaSuccess) { set_error("adamw launch failed: %s", cudaGetErrorString(le)); rc = (int32_t)le; }
    }
    cudaError_t ce = cudaStreamEndCapture(s, &graph);
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (ce != cudaSuccess) {
      if (graph) cudaGraphDestroy(graph);
      BX_CUDA(ce);
    }
    {
      const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
      if (ie != cudaSuccess) {  // the captured graph must not leak when instantiation fails
        cudaGraphDestroy(graph);
        BX_CUDA(ie);
      }
    }
    for (int i = 0; i < steps; ++i) {
      cudaError_t e = cudaGraphLaunch(exec, s);
      if (e != cudaSuccess) {
        set_error("cudaGraphLaunch failed: %s", cudaGetErrorString(e));
        cudaGraphExecDestroy(exec); cudaGraphDestroy(graph);
        return (int32_t)e;
      }
    }
    // The exec object must outlive its queued launches.
    BX_CUDA(cudaStreamSynchronize(s));
    cudaGraphExecDestroy(exec);
    cudaGraphDestroy(graph);
    return 0;
  }
  for (int i = 0; i < steps; ++i) {
    rc = launch_grad(s, d_X, d_y, n, d, d_raw, w);
    if (rc) return rc;
    adamw_kernel<<<1, 32, 0, s>>>(d_raw, w.gbuf, d, lr, w.adam, d_trace, i);
    BX_LAUNCH_CHECK();
  }
  return 0;
}

// Concurrent branches of one batched call: branch 0 is the caller's stream, branches 1.. are side streams that fork
// from it and join it again through events (inside a stream capture they become parallel branches of the graph).
// Created and destroyed with the call: the library keeps no persistent state.
namespace {
struct Branches {
  std::vector<cudaStream_t> st;
  std::vector<cudaEvent_t> done;
  cudaEvent_t forked = nullptr;
  cudaStream_t main = nullptr;
  int32_t init(int R, cudaStream_t s) {
    main = s;
    st.assign(R, nullptr);
    done.assign(R, nullptr);
    st[0] = s;
    BX_CUDA(cudaEventCreateWithFlags(&forked, cudaEventDisableTiming));
    for (int r = 1; r < R; ++r) {
      BX_CUDA(cudaStreamCreateWithFlags(&st[r], cudaStreamNonBlocking));
      BX_CUDA(cudaEventCreateWithFlags(&done[r], cudaEventDisableTiming));
    }
    return 0;
  }
  int32_t fork() {
    BX_CUDA(cudaEventRecord(forked, main));
    for (size_t r = 1; r < st.size(); ++r) BX_CUDA(cudaStreamWaitEvent(st[r], forked, 0));
    return 0;
  }
  int32_t join() {
    for (size_t r = 1; r < st.size(); ++r) {
      BX_CUDA(cudaEventRecord(done[r], st[r]));
      BX_CUDA(cudaStreamWaitEvent(main, done[r], 0));
    }
    return 0;
  }
  ~Branches() {
    for (size_t r = 1; r < st.size(); ++r) {
      if (done[r]) cudaEventDestroy(done[r]);
      if (st[r]) cudaStreamDestroy(st[r]);
    }
    if (forked) cudaEventDestroy(forked);
  }
};
}  // namespace

extern "C" int32_t bx_fit_adamw_batch(const float* d_X, const float* d_y, int32_t n, int32_t d, float* d_raws, int32_t R,
                                      float lr, int32_t steps, float* d_out, void* d_ws, size_t ws_bytes, void* stream) {
  BX_CHECK_ARG(d_X && d_y && d_raws && d_ws, BX_E_ARG, "null pointer argument");
  BX_CHECK_ARG(n > 0 && d > 0, BX_E_ARG, "n and d must be positive (n=%d d=%d)", n, d);
  BX_CHECK_ARG(d <= BX_MAX_D, BX_E_RANGE, "d=%d exceeds BX_MAX_D=%d", d, BX_MAX_D);
  BX_CHECK_ARG(R >= 1 && R <= BX_FIT_BATCH_MAX, BX_E_RANGE, "R=%d outside [1, %d]", R, BX_FIT_BATCH_MAX);
  BX_CHECK_ARG(steps >= 0, BX_E_ARG, "steps must be >= 0");
  const size_t slot = bx_fit_workspace_bytes(n, d);
  BX_CHECK_ARG(ws_bytes >= slot * (size_t)R, BX_E_WORKSPACE, "workspace too small: %zu < %zu", ws_bytes, slot * (size_t)R);
  cudaStream_t s = (cudaStream_t)stream;
  const int P = d + 2, W = 2 + 2 * P;
  std::vector<FitWs> w(R);
  for (int r = 0; r < R; ++r) w[r] = carve_fit_ws(reinterpret_cast<char*>(d_ws) + slot * r, n, d);
  Branches br;
  int32_t rc = br.init(R, s);
  if (rc) return rc;
  if ((rc = cholesky_prepare())) return rc;

  if (fit_small_ok(n, d) && !getenv("BX_NO_FIT_SMALL")) {  // one launch per restart and phase, all restarts side by side
    if ((rc = br.fork())) return rc;
    for (int r = 0; r < R && rc == 0; ++r) {
      if (steps > 0) rc = launch_fit_small(br.st[r], d_X, d_y, n, d, d_raws + (size_t)r * P, lr, steps, nullptr, nullptr, 0);
      if (rc == 0 && d_out)
        rc = launch_fit_small(br.st[r], d_X, d_y, n, d, d_raws + (size_t)r * P, 0.f, 0, nullptr, d_out + (size_t)r * W, 1);
    }
    const int32_t jr = br.join();
    cudaStreamSynchronize(s);  // the side streams are destroyed on return
    return rc ? rc : jr;
  }

  if (steps > 0) {
    for (int r = 0; r < R; ++r)
      BX_CUDA(cudaMemsetAsync(w[r].adam, 0, sizeof(float) * (2 * P + 1), s));  // fresh moments per call (gp.py:100)
    // One step of ALL restarts is captured as one graph with R parallel branches and replayed `steps` times.
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    BX_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    rc = br.fork();
    for (int r = 0; r < R && rc == 0; ++r) {
      rc = launch_grad(br.st[r], d_X, d_y, n, d, d_raws + (size_t)r * P, w[r]);
      if (rc == 0) {
        adamw_kernel<<<1, 32, 0, br.st[r]>>>(d_raws + (size_t)r * P, w[r].gbuf, d, lr, w[r].adam, nullptr, 0);
        cudaError_t le = cudaGetLastError();
        if (le != cudaSuccess) { set_error("adamw launch failed: %s", cudaGetErrorString(le)); rc = (int32_t)le; }
      }
    }
    const int32_t jr = br.join();
    cudaError_t ce = cudaStreamEndCapture(s, &graph);
    if (rc == 0) rc = jr;
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    BX_CUDA(ce);
    cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
    for (int i = 0; i < steps && e == cudaSuccess; ++i) e = cudaGraphLaunch(exec, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);  // the exec object must outlive its queued launches
    if (exec) cudaGraphExecDestroy(exec);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { set_error("batched fit graph failed: %s", cudaGetErrorString(e)); return (int32_t)e; }
  }
  if (d_out) {  // NLML, loss and gradients at the fitted hypers, again side by side
    if ((rc = br.fork())) return rc;
    for (int r = 0; r < R && rc == 0; ++r) {
      rc = launch_grad(br.st[r], d_X, d_y, n, d, d_raws + (size_t)r * P, w[r]);
      if (rc == 0) {
        cudaError_t ce = cudaMemcpyAsync(d_out + (size_t)r * W, w[r].gbuf, sizeof(float) * W, cudaMemcpyDeviceToDevice, br.st[r]);
        if (ce != cudaSuccess) { set_error("cudaMemcpyAsync failed: %s", cudaGetErrorString(ce)); rc = (int32_t)ce; }
      }
    }
    const int32_t jr = br.join();
    cudaStreamSynchronize(s);
    if (rc == 0) rc = jr;
  }
  return rc;
}

extern "C" int32_t bx_factor_build(const float* d_X, const float* d_y, int32_t n, int32_t d, const float* d_raw,
                                   float* d_U, float* d_T, void* d_T16hi, void* d_T16lo,
                                   float* d_alpha, float* d_scale, float* d_hyp, void* d_ws, size_t ws_bytes, void* stream) {
  int32_t rc = check_fit_args(d_X, d_y, n, d, d_raw, d_ws, ws_bytes);
  if (rc) return rc;
  BX_CHECK_ARG(d_U && d_T && d_alpha && d_scale && d_hyp, BX_E_ARG, "null factor output");
  BX_CHECK_ARG((d_T16hi == nullptr) == (d_T16lo == nullptr), BX_E_ARG, "T16hi and T16lo must be given together");
  cudaStream_t s = (cudaStream_t)stream;
  FitWs w = carve_fit_ws(d_ws, n, d);
  if ((rc = cholesky_prepare())) return rc;
  Lookahead la;
  if (lookahead_wanted(w.np)) { if ((rc = la.init())) return rc; w.la = &la; }
  rc = launch_factor(s, d_X, d_y, n, d, d_raw, w, d_T, d_U, d_alpha, d_scale, d_hyp);
  if (rc) return rc;
  if ((rc = join_alpha(s, w))) return rc;  // zero_pad_absmax_kernel below rewrites the padding of T
  const size_t tot = (size_t)w.np * w.np;
  unsigned int* amax = reinterpret_cast<unsigned int*>(w.status) + 1;  // second word of the status slot
  BX_CUDA(cudaMemsetAsync(amax, 0, sizeof(unsigned int), s));
  zero_pad_absmax_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(d_T, n, w.np, amax);
  BX_LAUNCH_CHECK();
  split_T_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(d_T, w.np, reinterpret_cast<__half*>(d_T16hi),
                                                               reinterpret_cast<__half*>(d_T16lo), amax, d_hyp);
  BX_LAUNCH_CHECK();
  return 0;
}

namespace {
struct AppendWs {
  float *kvec, *l, *wv, *yc, *v, *tmp, *logdet;  // tmp: 3 np floats
  int* status;                          // [0] status, [1] max|T| bits
  size_t bytes;
};
AppendWs carve_append_ws(void* base, int np) {
  AppendWs w{};
  size_t off = 0;
  auto take = [&](size_t nfloat) {
    float* p = base ? reinterpret_cast<float*>(reinterpret_cast<char*>(base) + off) : nullptr;
    off += (nfloat * sizeof(float) + 255) / 256 * 256;
    return p;
  };
  w.kvec = take(np); w.l = take(np); w.wv = take(np); w.yc = take(np); w.v = take(np); w.tmp = take(3 * (size_t)np);
  w.logdet = take(1);
  w.status = reinterpret_cast<int*>(take(2));
  w.bytes = off;
  return w;
}
}  // namespace

extern "C" size_t bx_factor_append_workspace_bytes(int32_t np) {
  if (np <= 0 || np % BX_TILE) return 0;
  return carve_append_ws(nullptr, np).bytes;
}

extern "C" int32_t bx_factor_append(const bx_factor_t* f, const float* d_x_new, const float* d_y, void* d_ws, size_t ws_bytes,
                                    void* stream) {
  BX_CHECK_ARG(f && d_x_new && d_y && d_ws, BX_E_ARG, "null pointer argument");
  BX_CHECK_ARG(f->d_U && f->d_T && f->d_alpha && f->d_scale && f->d_hyp, BX_E_ARG, "incomplete factor");
  BX_CHECK_ARG((f->d_T16hi == nullptr) == (f->d_T16lo == nullptr), BX_E_ARG, "T16hi and T16lo must be given together");
  BX_CHECK_ARG(f->n > 0 && f->d > 0 && f->d <= BX_MAX_D && f->np % BX_TILE == 0 && f->d4 == round_up(f->d, 4), BX_E_ARG,
               "malformed factor (n=%d np=%d d=%d d4=%d)", f->n, f->np, f->d, f->d4);
  BX_CHECK_ARG(f->n < f->np, BX_E_RANGE, "no free row: n=%d fills the padded size np=%d (rebuild the factor with bx_factor_build)",
               f->n, f->np);
  BX_CHECK_ARG(ws_bytes >= bx_factor_append_workspace_bytes(f->np), BX_E_WORKSPACE, "workspace too small: %zu < %zu", ws_bytes,
               bx_factor_append_workspace_bytes(f->np));
  cudaStream_t s = (cudaStream_t)stream;
  const int n = f->n, np = f->np, d = f->d, d4 = f->d4;
  AppendWs w = carve_append_ws(d_ws, np);
  float *T = const_cast<float*>(f->d_T), *U = const_cast<float*>(f->d_U), *alpha = const_cast<float*>(f->d_alpha),
        *hyp = const_cast<float*>(f->d_hyp);
  append_cross_kernel<<<(np + 255) / 256, 256, 0, s>>>(d_x_new, n, d, np, d4, f->d_scale, hyp, U, w.kvec);
  BX_LAUNCH_CHECK();
  int32_t rc = trmv_lower(s, T, np, np, w.kvec, w.l);  // l = T k   (rows >= n of T are zero)
  if (rc) return rc;
  rc = trmv_lower_t(s, T, np, np, w.l, w.wv);          // T^T l
  if (rc) return rc;
  append_row_kernel<<<1, 1024, 0, s>>>(T, n, np, w.l, w.wv, hyp, w.logdet, w.status);
  BX_LAUNCH_CHECK();
  // the mean of y moved: centre again, alpha and the NLML from the grown factor
  center_y_kernel<<<1, 1024, 0, s>>>(d_y, n + 1, np, w.yc, hyp);
  BX_LAUNCH_CHECK();
  rc = launch_alpha_tail(s, T, U, alpha, hyp, n + 1, np, d4, w.yc, w.v, w.tmp, w.logdet, 1, w.status, true);
  if (rc) return rc;
  // the new row may move max|T|, i.e. the power-of-two scale of the fp16 copies: split again (O(n^2) bytes, no arithmetic)
  const size_t tot = (size_t)np * np;
  unsigned int* amax = reinterpret_cast<unsigned int*>(w.status) + 1;
  BX_CUDA(cudaMemsetAsync(amax, 0, sizeof(unsigned int), s));
  zero_pad_absmax_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(T, n + 1, np, amax);
  BX_LAUNCH_CHECK();
  split_T_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(T, np, reinterpret_cast<__half*>(const_cast<void*>(f->d_T16hi)),
                                                               reinterpret_cast<__half*>(const_cast<void*>(f->d_T16lo)), amax, hyp);
  BX_LAUNCH_CHECK();
  return 0;
}

extern "C" int32_t bx_check_status(const float* d_hyp, void* stream) {
  BX_CHECK_ARG(d_hyp, BX_E_ARG, "null pointer argument");
  float st = 0.f;
  BX_CUDA(cudaMemcpyAsync(&st, d_hyp + HYP_STATUS, sizeof(float), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  BX_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  if (st != 0.f) {
    set_error("Cholesky met a non-finite or non-positive pivot (status %g): the Gram matrix is not positive definite", (double)st);
    return BX_E_NOTPD;
  }
  return 0;
}
