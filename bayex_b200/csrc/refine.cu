// Multi-start refinement of the top-k candidates - the role of optimizer.py:39-73,198-199.
//
// The reference runs optax L-BFGS (unpinned dependency, mis-signed value_fn: SURVEY quirk Q12), so its trajectory is
// not a parity target.  What is kept: the same place in sample(), unconstrained iterates clipped to +-1e6
// (optimizer.py:68), scoring of the raw iterates (quirk Q11).  The iteration itself is a monotone ascent with
// analytic gradients, batched over all starts: per round one K* build, two skinny triangular products
// (V = T K*, W = T^T V = K^-1 k; refine_blk_kernel + refine_partsum_kernel) and one gradient/accept kernel; no host
// synchronisation anywhere.
// The NumPy restatement is oracle/optimizer.py:refine_ascent.
#include "internal.h"

namespace bx {

constexpr float kLn2 = 0.6931471805599453f;

struct RefineWs {
  float *Kst, *Kst2, *V, *W, *P, *xt, *xb, *gb, *fb, *step;
  int S16;  // starts rounded up to the group size of the skinny products
  size_t bytes;
};
constexpr int RG = 16;  // starts per group of the skinny triangular products
static RefineWs carve_refine_ws(void* base, int np, int d, int S) {
  RefineWs w{};
  w.S16 = round_up(S, RG);
  size_t off = 0;
  auto take = [&](size_t nfloat) {
    float* p = base ? reinterpret_cast<float*>(reinterpret_cast<char*>(base) + off) : nullptr;
    off += (nfloat * sizeof(float) + 255) / 256 * 256;
    return p;
  };
  w.Kst = take((size_t)np * w.S16);
  w.Kst2 = take((size_t)np * w.S16);  // K* of the next trial point, written by the update kernel
  w.V = take((size_t)np * w.S16);
  w.W = take((size_t)np * w.S16);
  w.P = take((size_t)(np / 128) * np * w.S16);  // per-block partials of the triangular products
  w.xt = take((size_t)w.S16 * d);
  w.xb = take((size_t)S * d);
  w.gb = take((size_t)S * d);
  w.fb = take(S);
  w.step = take(S);
  w.bytes = off;
  return w;
}

// Kst[s][j] = amp * exp(-|u_j - c_s|^2), START-major ([S16 x np]: every later kernel walks one start's row contiguously;
// the observation-major layout cost the update kernel a 32-byte sector per 4-byte read); zero for padded observations
// and unused start slots.
__global__ void __launch_bounds__(256) refine_kstar_kernel(bx_factor_t f, const float* __restrict__ xt, int S, int S16, float* Kst) {
  const int e = blockIdx.x * 256 + threadIdx.x;
  if (e >= f.np * S16) return;
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

// ---- the two triangular products of a round ----------------------------------------------------------------------------
// V[s][i] = sum_{j <= i} T[i][j] K*[s][j]   and   W[s][j] = sum_{i >= j} T[i][j] V[s][i]   (W = T^T V = K^-1 k*) for up to
// 50 starts (ceil(50 / N) per rank when the multi-start axis is sharded): skinny products, bound by the stream of T rather
// than by the FMA pipe.  T is cut into 128 x 128 blocks; one CTA per lower block loads its block once (coalesced, from L2:
// T stays resident across the 31 rounds), multiplies it with the 128-long slice of every group of RG = 16 starts and writes
// the block's contribution to a partial buffer; refine_partsum_kernel adds the partials of a row (column) of blocks in
// block order.  Nothing depends on how many starts ride along: a start's sums are formed in the same order whatever else
// is in the batch, so sharding the starts cannot change a bit (test_refinement_contract), and every rank of a sharded
// sample() refines exactly as one GPU would.
constexpr int RBS = 128;         // block size
constexpr int RLD = RBS + 1;     // padded row of the staged T block [k][out]: scalar accesses, conflict-free both ways

// Programmatic dependent launch: the ~155 kernels of a refinement are each a few microseconds long, so the gaps between
// them (drain + launch latency, ~5 us each) were a third of its time.  Every kernel of the loop lets its successor
// start at once (launch_dependents) and waits for its predecessor's results (wait: predecessor grid complete and its
// writes visible) before it reads or writes anything a predecessor touches; the block kernels stage their T block, which
// no kernel of the loop writes, ahead of that wait.  Launched without the attribute (BX_NO_PDL), both are no-ops.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ void lower_block(int b, int& I, int& J) {  // b-th block of the lower triangle, row-major
  I = (int)((sqrtf(8.f * (float)b + 1.f) - 1.f) * 0.5f);
  while ((I + 1) * (I + 2) / 2 <= b) ++I;
  while (I * (I + 1) / 2 > b) --I;
  J = b - I * (I + 1) / 2;
}

// TRANS = false: P[J][s][I rows] = T[I,J] In[s][J cols];   TRANS = true: P[I][s][J cols] = T[I,J]^T In[s][I rows]
template <bool TRANS>
__global__ void __launch_bounds__(256) refine_blk_kernel(const float* __restrict__ T, int np, const float* __restrict__ In, int S16,
                                                         float* __restrict__ P) {
  extern __shared__ float sm[];
  float* Tb = sm;                  // [128 k][RLD out]
  float* Ib = Tb + RBS * RLD;      // [128 k][RG]
  float* Rb = Ib + RBS * RG;       // [2 k-halves][RG][128 out]
  int I, J;
  lower_block((int)blockIdx.x, I, J);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_launch_dependents();
  // stage the block k-major: k = column of T for V (out = row), k = row of T for W (out = column).  8 loads in flight per
  // thread before the first store (one load per iteration left the staging latency-bound: 64 x an L2 round trip per CTA;
  // 16-byte loads with the block kept [row][col] measured 15 % slower end to end, r2v)
  {
    const float* Tblk = T + (size_t)(I * RBS) * np + J * RBS;
#pragma unroll 1
    for (int b = 0; b < RBS * RBS / 256; b += 8) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int e = tid + 256 * (b + u);
        v[u] = Tblk[(size_t)(e >> 7) * np + (e & 127)];  // zero above the diagonal
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int e = tid + 256 * (b + u), r = e >> 7, c = e & 127;
        if (TRANS) Tb[r * RLD + c] = v[u]; else Tb[c * RLD + r] = v[u];
      }
    }
  }
  const int kblk = TRANS ? I : J, oblk = TRANS ? J : I, part = TRANS ? I : J;
  const int q4 = (warp & 3) * 4, kh = warp >> 2;  // this warp: starts q4 .. q4 + 3 of the group, k half kh; lanes = out
  pdl_wait();  // In is the predecessor's output, P was read by the kernel before it
  for (int g = 0; g < S16 / RG; ++g) {
    __syncthreads();  // Tb staged / the previous group's Ib and Rb consumed
    {
      float v[RG * RBS / 256];
#pragma unroll
      for (int u = 0; u < RG * RBS / 256; ++u) {
        const int e = tid + 256 * u;
        v[u] = In[(size_t)(g * RG + (e >> 7)) * np + kblk * RBS + (e & 127)];
      }
#pragma unroll
      for (int u = 0; u < RG * RBS / 256; ++u) {
        const int e = tid + 256 * u;
        Ib[(e & 127) * RG + (e >> 7)] = v[u];
      }
    }
    __syncthreads();
    float acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
#pragma unroll 4
    for (int kk = 0; kk < RBS / 2; ++kk) {
      const int k = kh * (RBS / 2) + kk;
      const float4 in = *reinterpret_cast<const float4*>(&Ib[k * RG + q4]);  // warp-uniform address: one broadcast
      const float iv[4] = {in.x, in.y, in.z, in.w};
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float t = Tb[k * RLD + lane + 32 * r];
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(t, iv[c], acc[r][c]);
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) Rb[(kh * RG + q4 + c) * RBS + lane + 32 * r] = acc[r][c];
    __syncthreads();
    for (int e = tid; e < RG * RBS; e += 256) {  // the two k halves, in order
      const int sI = e >> 7, o = e & 127;
      P[((size_t)part * S16 + g * RG + sI) * np + oblk * RBS + o] = Rb[sI * RBS + o] + Rb[(RG + sI) * RBS + o];
    }
  }
}

// Out[s][x] = sum of the partials of x's block row (V: parts 0 .. I) or block column (W: parts J .. nb - 1), in part order
template <bool TRANS>
__global__ void __launch_bounds__(256) refine_partsum_kernel(const float* __restrict__ P, int np, int S16, float* __restrict__ Out) {
  const size_t e = (size_t)blockIdx.x * 256 + threadIdx.x;
  pdl_launch_dependents();
  pdl_wait();
  if (e >= (size_t)S16 * np) return;
  const int x = (int)(e % np), blk = x / RBS, nb = np / RBS;
  const int p0 = TRANS ? blk : 0, p1 = TRANS ? nb : blk + 1;
  float acc = 0.f;
  for (int p = p0; p < p1; ++p) acc += P[(size_t)p * S16 * np + e];
  Out[e] = acc;
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

// One block per start: value + gradient of the acquisition at the trial point, accept/reject, next trial point.  The body is
// shared by refine_update_kernel (one launch per round) and refine_small_kernel (all rounds in one launch: the caller puts a
// block barrier between two calls); sm = 2 np floats of shared memory.
template <int RU_T>
__device__ __forceinline__ void refine_update_body(const bx_factor_t& f, const float* Kst, const float* V, const float* W, int acq_id,
                                                   float acq_param, float ystar, int first, float* xt, float* xb, float* gb,
                                                   float* fb, float* step, float* Kst_next, float* sm, int s) {
  float* ak = sm;              // [np] alpha_j k_j
  float* wk = sm + f.np;       // [np] w_j k_j
  constexpr int RU_W = RU_T / 32;
  __shared__ float red[RU_W];
  __shared__ float dmu[BX_MAX_D], dvar[BX_MAX_D], ct[BX_MAX_D];
  const int tid = threadIdx.x, d = f.d;
  if (tid < BX_MAX_D) ct[tid] = (tid < d) ? xt[(size_t)s * d + tid] * f.d_scale[tid] : 0.f;
  float mu_p = 0.f, ss_p = 0.f;
  for (int j = tid; j < f.np; j += RU_T) {
    const float k = Kst[(size_t)s * f.np + j], v = V[(size_t)s * f.np + j];
    const float w = W[(size_t)s * f.np + j];  // (K^-1 k*)_j
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
    const float xn = fminf(fmaxf(xbk + st * dir, -1e6f), 1e6f);  // optimizer.py:68
    xt[(size_t)s * d + tid] = xn;
    ct[tid] = xn * f.d_scale[tid];
  }
  // K* of the next trial point, for the next round's products (saves the separate K* launch of every round but the first)
  __syncthreads();
  for (int j = tid; j < f.np; j += RU_T) {
    float v = 0.f;
    if (j < f.n) {
      float acc = 0.f;
      for (int k = 0; k < d; ++k) {
        const float df = f.d_U[(size_t)j * f.d4 + k] - ct[k];
        acc = fmaf(df, df, acc);
      }
      v = f.d_hyp[HYP_AMP] * ex2_approx(-acc);
    }
    Kst_next[(size_t)s * f.np + j] = v;
  }
}

template <int RU_T>
__global__ void __launch_bounds__(RU_T) refine_update_kernel(bx_factor_t f, const float* __restrict__ Kst, const float* __restrict__ V,
                                                            const float* __restrict__ W, int acq_id, float acq_param,
                                                            float ystar, int first, float* xt, float* xb, float* gb,
                                                            float* fb, float* step, float* Kst_next) {
  extern __shared__ float sm[];
  pdl_launch_dependents();  // the next round's block kernel stages its T blocks under this kernel
  pdl_wait();
  refine_update_body<RU_T>(f, Kst, V, W, acq_id, acq_param, ystar, first, xt, xb, gb, fb, step, Kst_next, sm, (int)blockIdx.x);
}

// n <= 256 (np = 128 or 256: the README / test-suite regime and most real optimisation budgets): the WHOLE refinement of a
// start in one launch.  The lower triangle of T is at most 129 KB - it is staged once in shared memory, packed row by row
// (row i at i (i + 1) / 2: contiguous along a row, consecutive along a column), and a start's 31 rounds run back to back
// in its own CTA of np threads (thread i forms row i of V = T k*, then column i of W = T^T V, then the block runs the
// update body), with block barriers where the general path has kernel boundaries: ~95-160 launches of a few microseconds
// become one.  A start's arithmetic involves nobody else, so sharding the starts still cannot change a bit.
constexpr int RSM_MAX_NP = 256;
template <int NP>
__global__ void __launch_bounds__(NP) refine_small_kernel(bx_factor_t f, float* Kst, float* Kst2, float* V, float* W, int acq_id,
                                                          float acq_param, float ystar, int rounds, float* xt, float* xb,
                                                          float* gb, float* fb, float* step) {
  extern __shared__ float sm[];  // [2 NP] scratch of the update body | packed lower triangle of T
  float* Ts = sm + 2 * NP;
  const int s = blockIdx.x, i = threadIdx.x;
  for (int r = 0; r < NP; ++r)  // row r: entries 0 .. r (coalesced; everything above the diagonal is zero and never read)
    if (i <= r) Ts[r * (r + 1) / 2 + i] = f.d_T[(size_t)r * NP + i];
  float *kcur = Kst, *knext = Kst2;
  float* Vs = V + (size_t)s * NP;
  float* Ws = W + (size_t)s * NP;
  const float* Ti = Ts + i * (i + 1) / 2;
  for (int r = 0; r <= rounds; ++r) {
    __syncthreads();  // T staged / the previous round's K* written
    const float* k = kcur + (size_t)s * NP;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int j = 0;
    for (; j + 3 <= i; j += 4) {  // V[i] = sum_{j <= i} T[i][j] k[j]   (gp.py:68)
      a0 = fmaf(Ti[j], k[j], a0);
      a1 = fmaf(Ti[j + 1], k[j + 1], a1);
      a2 = fmaf(Ti[j + 2], k[j + 2], a2);
      a3 = fmaf(Ti[j + 3], k[j + 3], a3);
    }
    for (; j <= i; ++j) a0 = fmaf(Ti[j], k[j], a0);
    Vs[i] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    a0 = a1 = a2 = a3 = 0.f;
    int q = i;
    for (; q + 3 < NP; q += 4) {  // W[i] = sum_{q >= i} T[q][i] V[q]   (K^-1 k*)
      a0 = fmaf(Ts[q * (q + 1) / 2 + i], Vs[q], a0);
      a1 = fmaf(Ts[(q + 1) * (q + 2) / 2 + i], Vs[q + 1], a1);
      a2 = fmaf(Ts[(q + 2) * (q + 3) / 2 + i], Vs[q + 2], a2);
      a3 = fmaf(Ts[(q + 3) * (q + 4) / 2 + i], Vs[q + 3], a3);
    }
    for (; q < NP; ++q) a0 = fmaf(Ts[q * (q + 1) / 2 + i], Vs[q], a0);
    Ws[i] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    refine_update_body<NP>(f, kcur, V, W, acq_id, acq_param, ystar, r == 0 ? 1 : 0, xt, xb, gb, fb, step, knext, sm, s);
    float* t = kcur; kcur = knext; knext = t;
  }
}

int32_t check_factor(const bx_factor_t* f);

}  // namespace bx

using namespace bx;

namespace {
// launch with programmatic stream serialisation (see pdl_wait above); plain stream order with BX_NO_PDL set
template <typename... KArgs, typename... Args>
cudaError_t launch_chain(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
  static const bool no_pdl = getenv("BX_NO_PDL") != nullptr;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = no_pdl ? 0 : 1;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
}  // namespace

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
  BX_CUDA(cudaMemsetAsync(w.xt, 0, sizeof(float) * (size_t)w.S16 * f->d, s));
  BX_CUDA(cudaMemcpyAsync(w.xt, d_x, sizeof(float) * (size_t)S * f->d, cudaMemcpyDeviceToDevice, s));
  const bool wide = f->np > 256;
  BX_CUDA(cudaFuncSetAttribute(wide ? refine_update_kernel<512> : refine_update_kernel<128>,
                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)usm));
  const int nb = f->np / RBS, nblk = nb * (nb + 1) / 2;
  const size_t bsm = sizeof(float) * (RBS * RLD + RBS * RG + 2 * RG * RBS);
  BX_CUDA(cudaFuncSetAttribute(refine_blk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsm));
  BX_CUDA(cudaFuncSetAttribute(refine_blk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsm));
  const unsigned gsum = (unsigned)(((size_t)w.S16 * f->np + 255) / 256);
  BX_CUDA(cudaMemsetAsync(w.Kst2, 0, sizeof(float) * (size_t)w.S16 * f->np, s));  // unused start slots stay zero
  refine_kstar_kernel<<<(f->np * w.S16 + 255) / 256, 256, 0, s>>>(*f, w.xt, S, w.S16, w.Kst);
  BX_LAUNCH_CHECK();
  const bool no_small = getenv("BX_NO_REFINE_SMALL") != nullptr;  // read per call: the A/B of the two paths in one process
  if (f->np <= RSM_MAX_NP && !no_small) {  // n <= 256: one launch for all rounds
    const size_t ssm = sizeof(float) * (2 * (size_t)f->np + (size_t)f->np * (f->np + 1) / 2);
    if (f->np == 128) {
      BX_CUDA(cudaFuncSetAttribute(refine_small_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssm));
      refine_small_kernel<128><<<S, 128, ssm, s>>>(*f, w.Kst, w.Kst2, w.V, w.W, acq_id, acq_param, ystar, rounds, w.xt, w.xb, w.gb,
                                                   w.fb, w.step);
    } else {
      BX_CUDA(cudaFuncSetAttribute(refine_small_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssm));
      refine_small_kernel<256><<<S, 256, ssm, s>>>(*f, w.Kst, w.Kst2, w.V, w.W, acq_id, acq_param, ystar, rounds, w.xt, w.xb, w.gb,
                                                   w.fb, w.step);
    }
    BX_LAUNCH_CHECK();
    rounds = -1;  // skip the per-round loop below
  }
  float *kcur = w.Kst, *knext = w.Kst2;
  for (int r = 0; r <= rounds; ++r) {
    if (nb == 1) {  // a single block: its partial product is the product
      BX_CUDA(launch_chain(refine_blk_kernel<false>, nblk, 256, bsm, s, f->d_T, f->np, kcur, w.S16, w.V));   // V = T K*  (gp.py:68)
      BX_CUDA(launch_chain(refine_blk_kernel<true>, nblk, 256, bsm, s, f->d_T, f->np, w.V, w.S16, w.W));     // W = T^T V = K^-1 k*
    } else {
      BX_CUDA(launch_chain(refine_blk_kernel<false>, nblk, 256, bsm, s, f->d_T, f->np, kcur, w.S16, w.P));
      BX_CUDA(launch_chain(refine_partsum_kernel<false>, gsum, 256, 0, s, w.P, f->np, w.S16, w.V));
      BX_CUDA(launch_chain(refine_blk_kernel<true>, nblk, 256, bsm, s, f->d_T, f->np, w.V, w.S16, w.P));
      BX_CUDA(launch_chain(refine_partsum_kernel<true>, gsum, 256, 0, s, w.P, f->np, w.S16, w.W));
    }
    if (wide)
      BX_CUDA(launch_chain(refine_update_kernel<512>, S, 512, usm, s, *f, kcur, w.V, w.W, acq_id, acq_param, ystar, r == 0 ? 1 : 0,
                           w.xt, w.xb, w.gb, w.fb, w.step, knext));
    else
      BX_CUDA(launch_chain(refine_update_kernel<128>, S, 128, usm, s, *f, kcur, w.V, w.W, acq_id, acq_param, ystar, r == 0 ? 1 : 0,
                           w.xt, w.xb, w.gb, w.fb, w.step, knext));
    float* t = kcur; kcur = knext; knext = t;
  }
  BX_CUDA(cudaMemcpyAsync(d_x, w.xb, sizeof(float) * (size_t)S * f->d, cudaMemcpyDeviceToDevice, s));
  BX_CUDA(cudaMemcpyAsync(d_val, w.fb, sizeof(float) * S, cudaMemcpyDeviceToDevice, s));
  return 0;
}
