// Multi-start refinement of the top-k candidates - the role of optimizer.py:39-73,198-199.
//
// The reference runs optax L-BFGS (unpinned dependency, mis-signed value_fn: SURVEY quirk Q12), so its trajectory is
// not a parity target.  What is kept: the same place in sample(), unconstrained iterates clipped to +-1e6
// (optimizer.py:68), scoring of the raw iterates (quirk Q11).  The iteration itself is a monotone ascent with
// analytic gradients, batched over all starts: per round one K* build, two triangular GEMMs
// (V = T K*, W = T^T V = K^-1 k) and one gradient/accept kernel; no host synchronisation anywhere.
// The NumPy restatement is oracle/optimizer.py:refine_ascent.
#include "internal.h"

namespace bx {

constexpr float kLn2 = 0.6931471805599453f;

struct RefineWs {
  float *Kst, *V, *Vp, *Wp, *xt, *xb, *gb, *fb, *step;
  int S64, ksplit, nsplit;  // the two skinny triangular products are split along K into nsplit chunks of ksplit
  size_t bytes;
};
static RefineWs carve_refine_ws(void* base, int np, int d, int S) {
  RefineWs w{};
  w.S64 = round_up(S, 64);
  // np/64 row tiles x 64 starts is too few CTAs and the last row tile carries the whole K range: split K into <= 8
  // chunks (>= 128 long), partial products summed in a fixed order (deterministic: every rank must refine identically).
  // CUPTI at n = 512 with one chunk: 20 us per product (8 CTAs, 32 k-steps each), 1.2 of the 2.8 ms of a C2 sample().
  w.ksplit = max(128, round_up((np + 7) / 8, 64));  // (the products are computed transposed: [S64 x np] = starts x observations)
  w.nsplit = (np + w.ksplit - 1) / w.ksplit;
  size_t off = 0;
  auto take = [&](size_t nfloat) {
    float* p = base ? reinterpret_cast<float*>(reinterpret_cast<char*>(base) + off) : nullptr;
    off += (nfloat * sizeof(float) + 255) / 256 * 256;
    return p;
  };
  w.Kst = take((size_t)np * w.S64);
  w.V = take((size_t)np * w.S64);
  w.Vp = take((size_t)w.nsplit * np * w.S64);
  w.Wp = take((size_t)w.nsplit * np * w.S64);
  w.xt = take((size_t)w.S64 * d);
  w.xb = take((size_t)S * d);
  w.gb = take((size_t)S * d);
  w.fb = take(S);
  w.step = take(S);
  w.bytes = off;
  return w;
}

// Kst[s][j] = amp * exp(-|u_j - c_s|^2), START-major ([S64 x np]: every later kernel walks one start's row contiguously;
// the observation-major layout cost the update kernel a 32-byte sector per 4-byte read); zero for padded observations
// and unused start slots.
__global__ void __launch_bounds__(256) refine_kstar_kernel(bx_factor_t f, const float* __restrict__ xt, int S, int S64, float* Kst) {
  const int e = blockIdx.x * 256 + threadIdx.x;
  if (e >= f.np * S64) return;
  const int s = e / f.np, j = e % f.np;
  float v = 0.f;
  if (j < f.n && s < S) {
    float acc = 0.f;
    for (int k = 0; k < f.d; ++k) {
      const float df = f.d_U[(size_t)j * f.d4 + k] - xt[(size_t)s * f.d + k] * f.d_scale[k];
      acc = fmaf(df, df, acc);
    }
    v = f.d_hyp[HYP_AMP] * ex2_approx(-acc);  // same ex2 as the scoring kernels
  }
  Kst[e] = v;
}

// V = sum of the K-chunk partials, in chunk order
__global__ void __launch_bounds__(256) refine_reduce_kernel(const float* __restrict__ P, int nsplit, size_t stride, float* V) {
  const size_t e = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (e >= stride) return;
  float acc = P[e];
  for (int z = 1; z < nsplit; ++z) acc += P[(size_t)z * stride + e];
  V[e] = acc;
}

// refine_update runs RU_T threads per start: 512 (4 j-iterations at n = 2048: the loops are global-latency bound) or, for
// np <= 256, 128 (fewer warps to reduce across)
template <int RU_W>
__device__ __forceinline__ float block_sum_128(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < RU_W; ++w) t += red[w];  // fixed order: deterministic
  return t;
}

// One block per start: value + gradient of the acquisition at the trial point, accept/reject, next trial point.
template <int RU_T>
__global__ void __launch_bounds__(RU_T) refine_update_kernel(bx_factor_t f, const float* __restrict__ Kst, const float* __restrict__ V,
                                                            const float* __restrict__ Wp, int nsplit, int S64, int acq_id, float acq_param,
                                                            float ystar, int first, float* xt, float* xb, float* gb,
                                                            float* fb, float* step) {
  extern __shared__ float sm[];
  float* ak = sm;              // [np] alpha_j k_j
  float* wk = sm + f.np;       // [np] w_j k_j
  constexpr int RU_W = RU_T / 32;
  __shared__ float red[RU_W];
  __shared__ float dmu[BX_MAX_D], dvar[BX_MAX_D], ct[BX_MAX_D];
  const int s = blockIdx.x, tid = threadIdx.x, d = f.d;
  if (tid < BX_MAX_D) ct[tid] = (tid < d) ? xt[(size_t)s * d + tid] * f.d_scale[tid] : 0.f;
  float mu_p = 0.f, ss_p = 0.f;
  for (int j = tid; j < f.np; j += RU_T) {
    const float k = Kst[(size_t)s * f.np + j], v = V[(size_t)s * f.np + j];
    float w = Wp[(size_t)s * f.np + j];  // W^T = V^T T: sum of the K-chunk partials, in chunk order
    for (int z = 1; z < nsplit; ++z) w += Wp[((size_t)z * S64 + s) * f.np + j];
    const float a = f.d_alpha[j] * k;
    ak[j] = a;
    wk[j] = w * k;
    mu_p += a;
    ss_p = fmaf(v, v, ss_p);
  }
  const float mu = block_sum_128<RU_W>(mu_p, red) + f.d_hyp[HYP_YMEAN];
  const float ssq = block_sum_128<RU_W>(ss_p, red);
  // gradient sums over the observations, 16 dimensions per pass: every thread walks its rows of U once per pass with
  // float4 loads (the per-dimension loop read U with a stride of d4 floats: 25 % sector efficiency, d passes)
  __shared__ float red2[RU_W][32];
  for (int k0 = 0; k0 < d; k0 += 16) {
    float a[16], b[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = b[i] = 0.f;
    for (int j = tid; j < f.n; j += RU_T) {
      const float akj = ak[j], wkj = wk[j];
      const float* urow = f.d_U + (size_t)j * f.d4 + k0;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (k0 + 4 * q < f.d4) {
          const float4 u = *reinterpret_cast<const float4*>(urow + 4 * q);
          const float g0 = u.x - ct[k0 + 4 * q], g1 = u.y - ct[k0 + 4 * q + 1], g2 = u.z - ct[k0 + 4 * q + 2], g3 = u.w - ct[k0 + 4 * q + 3];
          a[4 * q] = fmaf(akj, g0, a[4 * q]); a[4 * q + 1] = fmaf(akj, g1, a[4 * q + 1]);
          a[4 * q + 2] = fmaf(akj, g2, a[4 * q + 2]); a[4 * q + 3] = fmaf(akj, g3, a[4 * q + 3]);
          b[4 * q] = fmaf(wkj, g0, b[4 * q]); b[4 * q + 1] = fmaf(wkj, g1, b[4 * q + 1]);
          b[4 * q + 2] = fmaf(wkj, g2, b[4 * q + 2]); b[4 * q + 3] = fmaf(wkj, g3, b[4 * q + 3]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      a[i] = warp_sum(a[i]);
      b[i] = warp_sum(b[i]);
    }
    __syncthreads();
    if ((tid & 31) == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) { red2[tid >> 5][i] = a[i]; red2[tid >> 5][16 + i] = b[i]; }
    }
    __syncthreads();
    if (tid < 32) {
      const int k = k0 + (tid & 15);
      if (k < d) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < RU_W; ++w) v += red2[w][tid];
        const float gs = 2.f * kLn2 * f.d_scale[k];  // d k_j / d x_k = k_j * 2 ln2 (u_jk - c_k) scale_k
        if (tid < 16) dmu[k] = v * gs; else dvar[k] = -2.f * v * gs;
      }
    }
  }
  __syncthreads();
  // scalar part on one thread (value, gradient coefficients, accept / step control) ...
  __shared__ float ctl[4];   // accept, step, c_mu, c_sd * d(sd)/d(var)
  __shared__ float gl2[BX_MAX_D];
  if (tid == 0) {
    const float amp = f.d_hyp[HYP_AMP];
    const float var = amp - ssq, sd = sqrtf(fmaxf(var, 1e-10f));
    const float dsd_scale = (var > 1e-10f) ? 0.5f / sd : 0.f;
    float val, c_mu, c_sd;  // gradient = c_mu * dmu + c_sd * dstd
    if (acq_id == BX_ACQ_EI) {
      const float A = mu - ystar - acq_param, sp = sd + 1e-3f, z = A / sp, cdf = ndtr_f(z), pdf = npdf_f(z);
      val = A * cdf + sd * pdf;
      const float q = (A - sd * z) * pdf;  // coefficient of dz
      c_mu = cdf + q / sp;
      c_sd = pdf - q * A / (sp * sp);
    } else if (acq_id == BX_ACQ_PI) {
      const float A = mu - ystar - acq_param, z = A / sd, pdf = npdf_f(z);
      val = ndtr_f(z);
      c_mu = pdf / sd;
      c_sd = -pdf * A / (sd * sd);
    } else {
      const float kap = (acq_id == BX_ACQ_UCB) ? acq_param : -acq_param;
      val = mu + kap * sd;
      c_mu = 1.f;
      c_sd = kap;
    }
    const bool accept = first || (val > fb[s]);
    const float st = first ? 0.1f : step[s] * (accept ? 2.f : 0.5f);
    step[s] = st;
    if (accept) fb[s] = val;
    ctl[0] = accept ? 1.f : 0.f; ctl[1] = st; ctl[2] = c_mu; ctl[3] = c_sd * dsd_scale;
  }
  __syncthreads();
  // ... the per-dimension part on d threads (the single-thread version walked xb / gb / xt through ~6 d dependent global
  // accesses: half of this kernel's 30 us at n = 2048)
  const bool accept = ctl[0] != 0.f;
  const float st = ctl[1];
  float gk = 0.f, xbk = 0.f, ls = 0.f;
  if (tid < d) {
    const int k = tid;
    ls = kSqrtLog2e / f.d_scale[k];
    if (accept) {
      xbk = xt[(size_t)s * d + k];
      gk = ctl[2] * dmu[k] + ctl[3] * dvar[k];
      xb[(size_t)s * d + k] = xbk;
      gb[(size_t)s * d + k] = gk;
    } else {
      xbk = xb[(size_t)s * d + k];
      gk = gb[(size_t)s * d + k];
    }
    const float gl = gk * ls;
    gl2[k] = gl * gl;
  }
  __syncthreads();
  if (tid < d) {
    // next trial: xb + st * (g * ls^2) / |g * ls|,  ls_k = sqrt(log2 e) / scale_k   (sum in index order: deterministic)
    float nrm = 0.f;
    for (int k = 0; k < d; ++k) nrm += gl2[k];
    nrm = sqrtf(nrm);
    const bool ok = (nrm > 0.f) && (nrm < CUDART_INF_F);
    const float dir = ok ? gk * ls * ls / nrm : 0.f;
    xt[(size_t)s * d + tid] = fminf(fmaxf(xbk + st * dir, -1e6f), 1e6f);  // optimizer.py:68
  }
}

int32_t check_factor(const bx_factor_t* f);

}  // namespace bx

using namespace bx;

extern "C" size_t bx_refine_workspace_bytes(const bx_factor_t* f, int32_t S) {
  if (!f || S <= 0 || f->np <= 0) return 0;
  return carve_refine_ws(nullptr, f->np, f->d, S).bytes;
}

extern "C" int32_t bx_refine(const bx_factor_t* f, float* d_x, int32_t S, int32_t rounds, int32_t acq_id, float acq_param,
                             float ystar, float* d_val, void* d_ws, size_t ws_bytes, void* stream) {
  int32_t rc = check_factor(f);
  if (rc) return rc;
  BX_CHECK_ARG(d_x && d_val && d_ws, BX_E_ARG, "null pointer argument");
  BX_CHECK_ARG(S > 0 && S <= BX_MAX_K && rounds >= 0, BX_E_RANGE, "S=%d rounds=%d out of range", S, rounds);
  BX_CHECK_ARG(acq_id >= 0 && acq_id <= 3, BX_E_ARG, "unknown acquisition id %d", acq_id);
  BX_CHECK_ARG(ws_bytes >= bx_refine_workspace_bytes(f, S), BX_E_WORKSPACE, "workspace too small");
  const size_t usm = 2 * (size_t)f->np * sizeof(float);
  BX_CHECK_ARG(usm <= 200 * 1024, BX_E_RANGE, "n=%d too large for the refinement kernel", f->n);
  cudaStream_t s = (cudaStream_t)stream;
  RefineWs w = carve_refine_ws(d_ws, f->np, f->d, S);
  BX_CUDA(cudaMemsetAsync(w.xt, 0, sizeof(float) * (size_t)w.S64 * f->d, s));
  BX_CUDA(cudaMemcpyAsync(w.xt, d_x, sizeof(float) * (size_t)S * f->d, cudaMemcpyDeviceToDevice, s));
  const bool wide = f->np > 256;
  BX_CUDA(cudaFuncSetAttribute(wide ? refine_update_kernel<512> : refine_update_kernel<128>,
                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)usm));
  for (int r = 0; r <= rounds; ++r) {
    refine_kstar_kernel<<<(f->np * w.S64 + 255) / 256, 256, 0, s>>>(*f, w.xt, S, w.S64, w.Kst);
    BX_LAUNCH_CHECK();
    const size_t pstride = (size_t)f->np * w.S64;
    GemmArgs g{};  // V^T [S64 x np] = Kst [S64 x np] * T^T   (T lower triangular: k <= i), K-chunk partials
    g.A = w.Kst; g.lda = f->np; g.B = f->d_T; g.ldb = f->np; g.C = w.Vp; g.ldc = f->np;
    g.M = w.S64; g.N = f->np; g.K = f->np; g.alpha = 1.f; g.beta = 0.f; g.kmode = KM_BT_LOWER;
    g.ksplit = w.ksplit; g.csplit = pstride;
    if (w.nsplit == 1) g.C = w.V;  // a single chunk is the product itself: no reduction launch
    rc = gemm(s, false, true, g, w.nsplit);
    if (rc) return rc;
    if (w.nsplit > 1) {
      refine_reduce_kernel<<<(unsigned)((pstride + 255) / 256), 256, 0, s>>>(w.Vp, w.nsplit, pstride, w.V);
      BX_LAUNCH_CHECK();
    }
    GemmArgs h{};  // W^T [S64 x np] = V^T * T   (k >= j), K-chunk partials (summed by the update kernel)
    h.A = w.V; h.lda = f->np; h.B = f->d_T; h.ldb = f->np; h.C = w.Wp; h.ldc = f->np;
    h.M = w.S64; h.N = f->np; h.K = f->np; h.alpha = 1.f; h.beta = 0.f; h.kmode = KM_B_LOWER;
    h.ksplit = w.ksplit; h.csplit = pstride;
    rc = gemm(s, false, false, h, w.nsplit);
    if (rc) return rc;
    if (wide)
      refine_update_kernel<512><<<S, 512, usm, s>>>(*f, w.Kst, w.V, w.Wp, w.nsplit, w.S64, acq_id, acq_param, ystar, r == 0 ? 1 : 0,
                                                    w.xt, w.xb, w.gb, w.fb, w.step);
    else
      refine_update_kernel<128><<<S, 128, usm, s>>>(*f, w.Kst, w.V, w.Wp, w.nsplit, w.S64, acq_id, acq_param, ystar, r == 0 ? 1 : 0,
                                                    w.xt, w.xb, w.gb, w.fb, w.step);
    BX_LAUNCH_CHECK();
  }
  BX_CUDA(cudaMemcpyAsync(d_x, w.xb, sizeof(float) * (size_t)S * f->d, cudaMemcpyDeviceToDevice, s));
  BX_CUDA(cudaMemcpyAsync(d_val, w.fb, sizeof(float) * S, cudaMemcpyDeviceToDevice, s));
  return 0;
}
