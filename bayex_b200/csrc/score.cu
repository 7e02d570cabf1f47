// Candidate generation and the SIMT (fp32 FFMA) scoring engine.
//
// score_simt_kernel fuses, per tile of 64 candidates:  Threefry generation or load (domain.py:74,130,159-161) ->
// cross-Gram tile K* (gp.py:60-64) -> mean K*^T alpha (gp.py:66-67) -> V = T K* with T = L^-1 (gp.py:68, as a
// triangular GEMM instead of a solve) -> column sums of V^2 (diag of gp.py:69, quirk Q8) -> std (gp.py:70) ->
// EI/PI/UCB/LCB (acq.py) -> packed arg-max (optimizer.py:205).  K* and V live only in shared memory / registers.
// This engine backs bx_score_points (explicit candidates: gp.predict, acq.*, refinement scoring) and is the
// reference implementation the tcgen05 engine (score_tc.cu) is validated against on the GPU.
#include "internal.h"
#include "topk_fused.cuh"

namespace bx {

int32_t make_sampler(const uint32_t key[2], const bx_dim_t* dims, int32_t d, SamplerPack* out, bool split_keys) {
  BX_CHECK_ARG(key && dims && out, BX_E_ARG, "null pointer argument");
  BX_CHECK_ARG(d > 0 && d <= BX_MAX_D, BX_E_RANGE, "d=%d outside [1, %d]", d, BX_MAX_D);
  out->d = d;
  for (int j = 0; j < d; ++j) {
    DimSampler& s = out->dim[j];
    uint32_t k0, k1;
    if (split_keys) threefry2x32(key[0], key[1], 0u, (uint32_t)j, k0, k1);  // split(key, d)[j]  (domain.py:160)
    else { k0 = key[0]; k1 = key[1]; }                                      // Domain.sample(key): the key itself
    s.kind = dims[j].kind;
    if (dims[j].kind == 0) {
      BX_CHECK_ARG(dims[j].lo < dims[j].hi, BX_E_ARG, "Lower bound must be less than upper bound (dim %d)", j);
      s.k0 = k0; s.k1 = k1; s.k2 = s.k3 = 0;
      s.lo = dims[j].lo; s.hi = dims[j].hi; s.scale = dims[j].hi - dims[j].lo;
      s.ilo = 0; s.span = 1; s.mult = 0;
    } else {
      BX_CHECK_ARG(dims[j].ilo < dims[j].ihi, BX_E_ARG, "Lower bound must be less than upper bound (dim %d)", j);
      threefry2x32(k0, k1, 0u, 0u, s.k0, s.k1);  // randint: k1, k2 = split(key)  (domain.py:130)
      threefry2x32(k0, k1, 0u, 1u, s.k2, s.k3);
      s.ilo = dims[j].ilo;
      s.span = (uint32_t)((int64_t)dims[j].ihi + 1 - (int64_t)dims[j].ilo);
      uint32_t m = 65536u % s.span;
      s.mult = (uint32_t)(((uint64_t)m * m) % s.span);
      s.lo = (float)dims[j].ilo; s.hi = (float)dims[j].ihi; s.scale = 0.f;
    }
  }
  return 0;
}

__global__ void threefry_fill_kernel(const __grid_constant__ SamplerPack sp, uint64_t m_begin, uint64_t m_count, float* out) {
  const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m_count * (uint64_t)sp.d) return;
  const uint64_t i = e / sp.d;
  const int k = (int)(e % sp.d);
  out[e] = sample_coord(sp.dim[k], m_begin + i);
}
__global__ void threefry_gather_kernel(const __grid_constant__ SamplerPack sp, const uint32_t* __restrict__ idx, int k,
                                       float* out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= k * sp.d) return;
  out[e] = sample_coord(sp.dim[e % sp.d], (uint64_t)idx[e / sp.d]);
}

// ------------------------------------------------------------------------------------------------------------------
constexpr int CT = 64;   // candidates per block
constexpr int RB = 64;   // T rows per accumulator tile
constexpr int RP = 4;    // accumulator tiles (row blocks) per pass over the observations
constexpr int KS = 16;   // observations per k-step
constexpr int SP = CT + 4;
constexpr int TSP = RP * RB + 4;


// Persistent over 64-candidate tiles.  Two ways to name the candidates of a launch:
//   direct (a.in_list == nullptr): rows [0, m_count) of the launch;
//   list   (a.in_list != nullptr): rows in_list[0 .. *in_count) - candidates another engine handed over (score_tc16p.cu);
//           count read on the device.
// A row's candidate is generated from the key (GEN, index m_begin + row) or read from a.cand[row]; outputs go to slot
// `row`.  With a.topk_k > 0 threads 0..127 also run the fused top-k collector (topk_fused.cuh), one slab per CTA.
template <bool GEN>
__global__ void __launch_bounds__(256) score_simt_kernel(const __grid_constant__ ScoreArgs a, const __grid_constant__ SamplerPack sp) {
  extern __shared__ __align__(16) float sm[];
  __shared__ tk::Shared tks;
  __shared__ uint32_t rows_s[CT];
  const int d = a.f.d, d4 = a.f.d4, np = a.f.np, cs_ld = d4 + 1;
  float* cs = sm;                         // [CT][d4+1] scaled candidate coordinates
  float* us = cs + CT * cs_ld;            // [KS][d4]   scaled observation rows of the current k-step
  float* Ts = us + KS * d4;               // [KS][TSP]  T tiles of the pass, k-major
  float* Ks = Ts + KS * TSP;              // [KS][SP]   K* tile, k-major
  float* red = Ks + KS * SP;              // [4][CT]    mean partials, then sum v^2
  const int tid = threadIdx.x;
  const uint64_t total = a.in_list ? (uint64_t)*a.in_count : a.m_count;
  const float amp = a.f.d_hyp[HYP_AMP];
  const int topk_k = a.topk_k;
  unsigned long long* const slab = topk_k ? a.topk_slab + (size_t)blockIdx.x * tk::CAP : nullptr;
  if (topk_k && tid < tk::TPB) {
    tk::init(tks, tid);
    tk::bar<1>();
  }
  unsigned long long best = 0ull;

  for (uint64_t c_base = (uint64_t)blockIdx.x * CT; c_base < total; c_base += (uint64_t)gridDim.x * CT) {
  if (tid < CT) rows_s[tid] = (c_base + tid < total) ? (a.in_list ? a.in_list[c_base + tid] : (uint32_t)(c_base + tid)) : 0xFFFFFFFFu;
  __syncthreads();
  for (int e = tid; e < CT * d4; e += 256) {
    const int c = e / d4, k = e % d4;
    float v = 0.f;
    const uint32_t row = rows_s[c];
    if (k < d && row != 0xFFFFFFFFu) {
      const float raw = GEN ? sample_coord(sp.dim[k], a.m_begin + row) : a.cand[(uint64_t)row * d + k];
      v = raw * a.f.d_scale[k];
    }
    cs[c * cs_ld + k] = v;
  }
  __syncthreads();

  const int rg = tid & 15, cg = tid >> 4;    // accumulator mapping: rows 4rg.. of every row block, candidates 4cg..
  const int gc = tid & 63, gk = tid >> 6;    // generation mapping: candidate gc, observations gk + 4q
  float ss[4] = {0.f, 0.f, 0.f, 0.f};
  float mpart = 0.f;
  const int nI = np / RB;
  // RP row blocks (256 rows of T) ride on every generated K* tile: K* is regenerated once per pass, not once per row block
  for (int I0 = 0; I0 < nI; I0 += RP) {
    const int nr = min(RP, nI - I0);
    const bool last = (I0 + RP >= nI);
    float acc[RP][4][4];
#pragma unroll
    for (int p = 0; p < RP; ++p)
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[p][r][c] = 0.f;
    const int ksteps = (I0 + nr) * (RB / KS);
    for (int ks = 0; ks < ksteps; ++ks) {
      const int k0 = ks * KS;
      const int p_lo = max(0, k0 / RB - I0);  // row blocks entirely above the diagonal at this k-step are skipped
      for (int e = tid; e < nr * RB * (KS / 4); e += 256) {  // T tiles: rows (I0 + p) * 64.., cols k0..k0+15 -> Ts[k][p * 64 + row]
        const int r = e >> 2, kk = (e & 3) * 4;
        if (r >= p_lo * RB) {
          const float4 v = *reinterpret_cast<const float4*>(a.f.d_T + (size_t)(I0 * RB + r) * np + k0 + kk);
          Ts[(kk + 0) * TSP + r] = v.x; Ts[(kk + 1) * TSP + r] = v.y; Ts[(kk + 2) * TSP + r] = v.z; Ts[(kk + 3) * TSP + r] = v.w;
        }
      }
      for (int e = tid; e < KS * d4; e += 256) us[e] = a.f.d_U[(size_t)k0 * d4 + e];
      __syncthreads();
#pragma unroll
      for (int q = 0; q < 4; ++q) {  // K* tile (gp.py:64): amp * exp(-|u_j - c|^2), never leaves shared memory
        const int kk = gk + 4 * q;
        float s = 0.f;
        for (int k = 0; k < d4; ++k) {
          const float df = us[kk * d4 + k] - cs[gc * cs_ld + k];
          s = fmaf(df, df, s);
        }
        const float kv = amp * ex2_approx(-s);
        Ks[kk * SP + gc] = kv;
        if (last) mpart = fmaf(a.f.d_alpha[k0 + kk], kv, mpart);  // gp.py:66-67
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < KS; ++kk) {
        const float4 kv = *reinterpret_cast<const float4*>(&Ks[kk * SP + cg * 4]);
        const float bv[4] = {kv.x, kv.y, kv.z, kv.w};
#pragma unroll
        for (int p = 0; p < RP; ++p) {
          if (p >= p_lo && p < nr) {
            const float4 tv = *reinterpret_cast<const float4*>(&Ts[kk * TSP + p * RB + rg * 4]);
            const float av[4] = {tv.x, tv.y, tv.z, tv.w};
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
              for (int c = 0; c < 4; ++c) acc[p][r][c] = fmaf(av[r], bv[c], acc[p][r][c]);
          }
        }
      }
      __syncthreads();
    }
#pragma unroll
    for (int p = 0; p < RP; ++p) {
      if (p < nr) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {  // column sums of V^2 over this row block, reduced across the 16 row groups
          float v = 0.f;
#pragma unroll
          for (int r = 0; r < 4; ++r) v = fmaf(acc[p][r][c], acc[p][r][c], v);
          v += __shfl_xor_sync(0xffffffffu, v, 1);
          v += __shfl_xor_sync(0xffffffffu, v, 2);
          v += __shfl_xor_sync(0xffffffffu, v, 4);
          v += __shfl_xor_sync(0xffffffffu, v, 8);
          ss[c] += v;
        }
      }
    }
  }
  red[gk * CT + gc] = mpart;
  __syncthreads();
  float mu = 0.f;
  if (tid < CT) mu = red[tid] + red[CT + tid] + red[2 * CT + tid] + red[3 * CT + tid] + a.f.d_hyp[HYP_YMEAN];
  __syncthreads();
  if (rg == 0) {
#pragma unroll
    for (int c = 0; c < 4; ++c) red[cg * 4 + c] = ss[c];
  }
  __syncthreads();
  if (tid < CT && rows_s[tid] != 0xFFFFFFFFu) {
    const float var = amp - red[tid];                 // diag(amp*cov(xt,xt) - v^T v)   gp.py:69
    const float sd = sqrtf(fmaxf(var, 1e-10f));       // gp.py:70
    const float av = acquisition(a.acq_id, mu, sd, a.ystar, a.acq_param);
    const uint64_t row = rows_s[tid];
    if (a.mu) a.mu[row] = mu;
    if (a.sigma) a.sigma[row] = sd;
    if (a.acq) a.acq[row] = av;
    const unsigned long long key = pack_best(av, (uint32_t)(a.index_base + row));
    best = key > best ? key : best;
    if (topk_k) tk::push(tks, slab, tk::topk_key(av, (uint32_t)(a.index_base + row)));
  }
  if (topk_k && tid < tk::TPB) tk::end_tile<1>(tks, slab, topk_k, tid);
  __syncthreads();  // rows_s / red are rewritten by the next tile
  }
  if (topk_k && tid < tk::TPB) tk::finish<1>(tks, slab, topk_k, tid, a.topk_count + blockIdx.x);
  if (a.best && tid < CT) {
    best = warp_max_u64(best);
    if ((tid & 31) == 0 && best) atomicMax(a.best, best);
  }
}

static size_t score_simt_smem(int d4) { return sizeof(float) * (CT * (d4 + 1) + KS * d4 + KS * TSP + KS * SP + 4 * CT); }

int32_t check_factor(const bx_factor_t* f) {
  BX_CHECK_ARG(f, BX_E_ARG, "null factor");
  BX_CHECK_ARG(f->d_U && f->d_T && f->d_alpha && f->d_scale && f->d_hyp, BX_E_ARG, "factor has null device pointers");
  BX_CHECK_ARG(f->n > 0 && f->d > 0 && f->d <= BX_MAX_D, BX_E_RANGE, "factor n=%d d=%d out of range", f->n, f->d);
  BX_CHECK_ARG(f->np == round_up(f->n, BX_TILE) && f->d4 == round_up(f->d, 4), BX_E_ARG, "factor np/d4 inconsistent");
  return 0;
}

// CTAs the SIMT engine launches for up to `m_count` candidates: persistent, at most SIMT_CTAS_PER_SM per SM
constexpr int SIMT_CTAS_PER_SM = 2;
unsigned score_simt_grid(uint64_t m_count, int sms) {
  const uint64_t ntiles = (m_count + CT - 1) / CT, cap = (uint64_t)sms * SIMT_CTAS_PER_SM;
  return (unsigned)(ntiles < cap ? (ntiles ? ntiles : 1) : cap);
}

int32_t launch_score_simt(cudaStream_t s, const ScoreArgs& a, const SamplerPack* sp, bool gen) {
  static SamplerPack empty{};
  const size_t smem = score_simt_smem(a.f.d4);
  int dev = 0, sms = 148;
  BX_CUDA(cudaGetDevice(&dev));
  BX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const unsigned grid = score_simt_grid(a.m_count, sms);
  if (gen) {
    BX_CUDA(cudaFuncSetAttribute(score_simt_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    score_simt_kernel<true><<<grid, 256, smem, s>>>(a, *sp);
  } else {
    BX_CUDA(cudaFuncSetAttribute(score_simt_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    score_simt_kernel<false><<<grid, 256, smem, s>>>(a, empty);
  }
  BX_LAUNCH_CHECK();
  return 0;
}

// implemented in score_tc16p.cu (fp16 head/tail, CTA pair, register-tiled producers): the production engine
int32_t launch_score_tc16p(cudaStream_t s, const ScoreArgs& a, const SamplerPack& sp, bool accurate = false);
bool score_tc16_supported(const bx_factor_t& f);

}  // namespace bx

using namespace bx;

extern "C" int32_t bx_threefry_fill(const uint32_t key[2], const bx_dim_t* dims, int32_t d, uint64_t m_begin,
                                    uint64_t m_count, float* d_out, void* stream) {
  SamplerPack sp;
  int32_t rc = make_sampler(key, dims, d, &sp);
  if (rc) return rc;
  BX_CHECK_ARG(d_out, BX_E_ARG, "null output");
  BX_CHECK_ARG(m_begin + m_count <= 0x100000000ull, BX_E_RANGE, "candidate indices must stay below 2^32");
  if (m_count == 0) return 0;
  const uint64_t tot = m_count * (uint64_t)d;
  threefry_fill_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(sp, m_begin, m_count, d_out);
  BX_LAUNCH_CHECK();
  return 0;
}

extern "C" int32_t bx_domain_sample(const uint32_t key[2], const bx_dim_t* dim, uint64_t m_begin, uint64_t m_count,
                                    float* d_out, void* stream) {
  SamplerPack sp;
  int32_t rc = make_sampler(key, dim, 1, &sp, false);
  if (rc) return rc;
  BX_CHECK_ARG(d_out, BX_E_ARG, "null output");
  BX_CHECK_ARG(m_begin + m_count <= 0x100000000ull, BX_E_RANGE, "candidate indices must stay below 2^32");
  if (m_count == 0) return 0;
  threefry_fill_kernel<<<(unsigned)((m_count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(sp, m_begin, m_count, d_out);
  BX_LAUNCH_CHECK();
  return 0;
}

extern "C" int32_t bx_threefry_gather(const uint32_t key[2], const bx_dim_t* dims, int32_t d, const uint32_t* d_idx,
                                      int32_t k, float* d_out, void* stream) {
  SamplerPack sp;
  int32_t rc = make_sampler(key, dims, d, &sp);
  if (rc) return rc;
  BX_CHECK_ARG(d_idx && d_out, BX_E_ARG, "null pointer argument");
  if (k <= 0) return 0;
  threefry_gather_kernel<<<(k * d + 255) / 256, 256, 0, (cudaStream_t)stream>>>(sp, d_idx, k, d_out);
  BX_LAUNCH_CHECK();
  return 0;
}

// ---- the two-tier production path ---------------------------------------------------------------------------------------
namespace bx {

static int sm_count() {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}
static size_t align256(size_t b) { return (b + 255) / 256 * 256; }

// Share of the prior variance below which a candidate of the fp16 tcgen05 engine is handed to the fp32 SIMT engine.
// Measured bias of the tensor engine on sum v^2: 3e-6 .. 6e-5 relative (scripts/dev/diag_bias_sign.py; worst on ill-
// conditioned factors, noise = softplus(-8)); for var >= tau amp that is at most 6e-5 (1 - tau) / tau of var.
static float handover_tau() {
  if (const char* e = getenv("BX_HANDOVER_TAU")) return (float)atof(e);
  return 0.5f;
}

struct HandoverWs {
  unsigned int* count;
  uint32_t* list;
  float* ux;  // the fast mode's copy of the observation rows (ScoreArgs::ux)
  size_t bytes;
};
static HandoverWs carve_handover(void* base, uint64_t m_count) {
  HandoverWs h{};
  char* b = reinterpret_cast<char*>(base);
  const size_t list_bytes = align256(4 * (size_t)m_count);
  h.count = reinterpret_cast<unsigned int*>(base);
  h.list = base ? reinterpret_cast<uint32_t*>(b + 256) : nullptr;
  h.ux = base ? reinterpret_cast<float*>(b + 256 + list_bytes) : nullptr;
  h.bytes = 256 + list_bytes + align256(score_tc16_ux_bytes());
  return h;
}

// Which engine scores the candidates the fast tensor mode hands over: the accurate tensor mode (default) or the fp32 SIMT
// engine (BX_TIER2=simt; also what a factor without fp16 copies or d > 32 gets).
static bool tier2_simt() {
  const char* e = getenv("BX_TIER2");
  return e && e[0] == 's';
}

// The production path: the fast mode of the fp16 tcgen05 engine over every candidate of the launch, then the accurate
// mode (or the SIMT engine) over the rows it handed over (variance in the cancellation regime).  a.topk_* (optional)
// name the slabs of the FIRST pass; the second pass uses the slabs after them.  mode: 0 two-tier, 1 fast mode alone
// (diagnostics), 2 accurate mode alone.  Returns the number of slabs in use through *nslabs.
static int32_t launch_two_tier(cudaStream_t s, ScoreArgs a, const SamplerPack& sp, bool gen, const HandoverWs& h, int mode,
                               int* nslabs) {
  const unsigned grid1 = score_tc16p_grid(a.m_count);
  if (nslabs) *nslabs = (int)grid1;
  a.ux = h.ux;
  if (mode != 0) return launch_score_tc16p(s, a, sp, mode == 2);
  BX_CUDA(cudaMemsetAsync(h.count, 0, sizeof(unsigned int), s));
  a.flag_list = h.list; a.flag_count = h.count; a.flag_tau = handover_tau();
  int32_t rc = launch_score_tc16p(s, a, sp, false);
  if (rc) return rc;
  a.flag_list = nullptr; a.flag_count = nullptr;
  a.in_list = h.list; a.in_count = h.count;
  if (a.topk_k) { a.topk_slab += (size_t)grid1 * tk::CAP; a.topk_count += grid1; }
  if (tier2_simt()) {
    if (nslabs) *nslabs = (int)(grid1 + score_simt_grid(a.m_count, sm_count()));
    return launch_score_simt(s, a, &sp, gen);
  }
  if (nslabs) *nslabs = (int)(2 * grid1);
  return launch_score_tc16p(s, a, sp, true);
}

struct SampleWs {
  unsigned long long *best, *slabs, *scratch;
  unsigned int* counts;
  HandoverWs h;
  size_t bytes;
};
static SampleWs carve_sample_ws(void* base, uint64_t m_count, int k) {
  SampleWs w{};
  const size_t nsl = (size_t)sm_count() * (1 + SIMT_CTAS_PER_SM);  // slabs: tensor pass + SIMT pass
  size_t off = 0;
  auto take = [&](size_t nbytes) {
    char* p = base ? reinterpret_cast<char*>(base) + off : nullptr;
    off += align256(nbytes);
    return p;
  };
  w.best = reinterpret_cast<unsigned long long*>(take(8));
  w.counts = reinterpret_cast<unsigned int*>(take(4 * nsl));
  w.slabs = reinterpret_cast<unsigned long long*>(take(8 * nsl * tk::CAP));
  w.scratch = reinterpret_cast<unsigned long long*>(take(8 * nsl * (size_t)k));
  w.h = carve_handover(take(carve_handover(nullptr, m_count).bytes), m_count);
  w.bytes = off;
  return w;
}

// optimizer.py:203-207: arg-max over concat(refined points, random candidates), first occurrence (a refined point wins a
// tie); the winner's raw coordinates (projection onto the domain happens in to_dict on the host: quirk Q11).
__global__ void propose_kernel(const __grid_constant__ SamplerPack sp, const float* __restrict__ opt_x, const float* __restrict__ opt_val,
                               int S, const unsigned long long* __restrict__ best_rnd, float* out) {
  __shared__ unsigned long long red[32];
  const int t = threadIdx.x, d = sp.d;
  unsigned long long b = 0ull;
  for (int j = t; j < S; j += blockDim.x) {
    const unsigned long long key = pack_best(opt_val[j], (uint32_t)j);
    b = key > b ? key : b;
  }
  b = warp_max_u64(b);
  if ((t & 31) == 0) red[t >> 5] = b;
  __syncthreads();
  if (t < 32) {
    b = t < (int)(blockDim.x >> 5) ? red[t] : 0ull;
    b = warp_max_u64(b);
    if (t == 0) red[0] = b;
  }
  __syncthreads();
  const unsigned long long b_opt = red[0], b_rnd = best_rnd ? *best_rnd : 0ull;
  const bool take_opt = S > 0 && (b_opt >> 32) >= (b_rnd >> 32);
  const uint32_t j = 0xFFFFFFFFu - (uint32_t)((take_opt ? b_opt : b_rnd) & 0xFFFFFFFFull);
  if (t < d) out[t] = take_opt ? opt_x[(size_t)j * d + t] : sample_coord(sp.dim[t], (uint64_t)j);
  if (t == 0) {
    out[d] = from_orderable_u32((uint32_t)((take_opt ? b_opt : b_rnd) >> 32));
    out[d + 1] = __uint_as_float(take_opt ? j : (uint32_t)S + j);  // position in the concatenation, as raw bits
  }
}

}  // namespace bx

extern "C" size_t bx_score_workspace_bytes(uint64_t M) { return carve_handover(nullptr, M).bytes; }

extern "C" int32_t bx_score_points(const bx_factor_t* f, const float* d_cand, uint64_t M, int32_t acq_id, float acq_param,
                                   float ystar, float* d_mu, float* d_sigma, float* d_acq, uint64_t* d_best,
                                   uint64_t index_base, void* d_ws, size_t ws_bytes, void* stream) {
  int32_t rc = check_factor(f);
  if (rc) return rc;
  BX_CHECK_ARG(d_cand, BX_E_ARG, "null candidates");
  BX_CHECK_ARG(acq_id >= 0 && acq_id <= 3, BX_E_ARG, "unknown acquisition id %d", acq_id);
  BX_CHECK_ARG(index_base + M <= 0x100000000ull, BX_E_RANGE, "candidate indices must stay below 2^32");
  if (M == 0) return 0;
  ScoreArgs a{};
  a.f = *f; a.cand = d_cand; a.m_begin = 0; a.m_count = M; a.acq_id = acq_id; a.acq_param = acq_param; a.ystar = ystar;
  a.mu = d_mu; a.sigma = d_sigma; a.acq = d_acq; a.best = reinterpret_cast<unsigned long long*>(d_best); a.index_base = index_base;
  // large explicit point sets (gp.predict / acq.* with M >= 2^16) take the two-tier production path when the factor
  // carries its fp16 copies and the caller gave the hand-over workspace: same fused pipeline, coordinates read from
  // d_cand instead of drawn
  static const SamplerPack no_sampler{};
  if (M >= BX_POINTS_TC_MIN && score_tc16_supported(*f) && d_ws) {
    BX_CHECK_ARG(ws_bytes >= bx_score_workspace_bytes(M), BX_E_WORKSPACE, "workspace too small");
    return launch_two_tier((cudaStream_t)stream, a, no_sampler, false, carve_handover(d_ws, M), 0, nullptr);
  }
  return launch_score_simt((cudaStream_t)stream, a, nullptr, false);
}

extern "C" int32_t bx_score_generated(const bx_factor_t* f, const uint32_t key[2], const bx_dim_t* dims, uint64_t m_begin,
                                      uint64_t m_count, int32_t acq_id, float acq_param, float ystar, float* d_acq,
                                      uint64_t* d_best, int32_t engine, void* d_ws, size_t ws_bytes, void* stream) {
  int32_t rc = check_factor(f);
  if (rc) return rc;
  BX_CHECK_ARG(acq_id >= 0 && acq_id <= 3, BX_E_ARG, "unknown acquisition id %d", acq_id);
  BX_CHECK_ARG(m_begin + m_count <= 0x100000000ull, BX_E_RANGE, "candidate indices must stay below 2^32");
  SamplerPack sp;
  rc = make_sampler(key, dims, f->d, &sp);
  if (rc) return rc;
  if (m_count == 0) return 0;
  ScoreArgs a{};
  a.f = *f; a.cand = nullptr; a.m_begin = m_begin; a.m_count = m_count; a.acq_id = acq_id; a.acq_param = acq_param;
  a.ystar = ystar; a.acq = d_acq; a.best = reinterpret_cast<unsigned long long*>(d_best); a.index_base = m_begin;
  // AUTO: the two-tier production path (fp16 tcgen05 engine + fp32 SIMT engine on the cancellation-regime candidates)
  // whenever the factor carries its fp16 copies and d <= 32; otherwise the fp32 SIMT engine alone.
  if (engine == BX_ENGINE_AUTO) engine = score_tc16_supported(*f) ? BX_ENGINE_TCGEN05_F16 : BX_ENGINE_SIMT;
  if (engine == BX_ENGINE_TCGEN05_F16 || engine == BX_ENGINE_TCGEN05_F16_RAW || engine == BX_ENGINE_TCGEN05_F16_ACC) {
    BX_CHECK_ARG(score_tc16_supported(*f), BX_E_UNSUPPORTED, "the fp16 tcgen05 engine needs the T16hi/T16lo split of the factor and d <= 32");
    const int mode = engine == BX_ENGINE_TCGEN05_F16 ? 0 : (engine == BX_ENGINE_TCGEN05_F16_RAW ? 1 : 2);
    BX_CHECK_ARG(mode == 2 || (d_ws && ws_bytes >= bx_score_workspace_bytes(m_count)), BX_E_WORKSPACE, "scoring workspace missing or too small");
    return launch_two_tier((cudaStream_t)stream, a, sp, true, carve_handover(d_ws, m_count), mode, nullptr);
  }
  BX_CHECK_ARG(engine == BX_ENGINE_SIMT, BX_E_ARG, "unknown engine %d", engine);
  return launch_score_simt((cudaStream_t)stream, a, &sp, true);
}

extern "C" size_t bx_score_topk_workspace_bytes(uint64_t m_count, int32_t k) {
  if (k <= 0 || k > BX_MAX_K) return 0;
  return carve_sample_ws(nullptr, m_count, k).bytes;
}

// ---- fused sample() front half: generate -> score -> top-k -> arg-max (optimizer.py:183-197, :205) ---------------------
extern "C" int32_t bx_score_generated_topk(const bx_factor_t* f, const uint32_t key[2], const bx_dim_t* dims, uint64_t m_begin,
                                           uint64_t m_count, int32_t acq_id, float acq_param, float ystar, int32_t k,
                                           uint64_t* d_list, float* d_vals, uint32_t* d_idx, int32_t* d_count, float* d_acq,
                                           int32_t engine, void* d_ws, size_t ws_bytes, void* stream) {
  int32_t rc = check_factor(f);
  if (rc) return rc;
  BX_CHECK_ARG(d_list && d_ws, BX_E_ARG, "null pointer argument");
  BX_CHECK_ARG(acq_id >= 0 && acq_id <= 3, BX_E_ARG, "unknown acquisition id %d", acq_id);
  BX_CHECK_ARG(k > 0 && k <= BX_MAX_K, BX_E_RANGE, "k=%d outside [1, %d]", k, BX_MAX_K);
  BX_CHECK_ARG(m_count > 0 && m_begin + m_count <= 0x100000000ull, BX_E_RANGE, "candidate indices must stay in [0, 2^32)");
  BX_CHECK_ARG(ws_bytes >= bx_score_topk_workspace_bytes(m_count, k), BX_E_WORKSPACE, "workspace too small");
  SamplerPack sp;
  rc = make_sampler(key, dims, f->d, &sp);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  SampleWs w = carve_sample_ws(d_ws, m_count, k);
  BX_CUDA(cudaMemsetAsync(w.best, 0, 8, s));
  ScoreArgs a{};
  a.f = *f; a.cand = nullptr; a.m_begin = m_begin; a.m_count = m_count; a.acq_id = acq_id; a.acq_param = acq_param;
  a.ystar = ystar; a.acq = d_acq; a.best = w.best; a.index_base = m_begin;
  a.topk_slab = w.slabs; a.topk_count = w.counts; a.topk_k = k;
  if (engine == BX_ENGINE_AUTO) engine = score_tc16_supported(*f) ? BX_ENGINE_TCGEN05_F16 : BX_ENGINE_SIMT;
  int nslabs = 0;
  if (engine == BX_ENGINE_TCGEN05_F16 || engine == BX_ENGINE_TCGEN05_F16_RAW || engine == BX_ENGINE_TCGEN05_F16_ACC) {
    BX_CHECK_ARG(score_tc16_supported(*f), BX_E_UNSUPPORTED, "the fp16 tcgen05 engine needs the T16hi/T16lo split of the factor and d <= 32");
    rc = launch_two_tier(s, a, sp, true, w.h, engine == BX_ENGINE_TCGEN05_F16 ? 0 : (engine == BX_ENGINE_TCGEN05_F16_RAW ? 1 : 2), &nslabs);
  } else {
    BX_CHECK_ARG(engine == BX_ENGINE_SIMT, BX_E_UNSUPPORTED, "the fused top-k runs on the AUTO / SIMT / fp16 tcgen05 engines (engine %d)", engine);
    nslabs = (int)score_simt_grid(m_count, sm_count());
    rc = launch_score_simt(s, a, &sp, true);
  }
  if (rc) return rc;
  rc = launch_topk_merge(s, w.slabs, w.counts, nslabs, tk::CAP, k, w.scratch, reinterpret_cast<unsigned long long*>(d_list), w.best, 1,
                         1, d_vals, d_idx, d_count);
  if (rc) return rc;
  if (d_count) {  // d_count[1]: candidates the fast mode handed to the second tier (0 for single-engine launches)
    if (engine == BX_ENGINE_TCGEN05_F16) BX_CUDA(cudaMemcpyAsync(d_count + 1, w.h.count, sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
    else BX_CUDA(cudaMemsetAsync(d_count + 1, 0, sizeof(int32_t), s));
  }
  return 0;
}

extern "C" int32_t bx_propose(const uint32_t key[2], const bx_dim_t* dims, int32_t d, const float* d_opt_x, const float* d_opt_val,
                              int32_t S, const uint64_t* d_best, float* d_out, void* stream) {
  SamplerPack sp;
  int32_t rc = make_sampler(key, dims, d, &sp);
  if (rc) return rc;
  BX_CHECK_ARG(d_out && (S == 0 || (d_opt_x && d_opt_val)) && (S > 0 || d_best), BX_E_ARG, "null pointer argument");
  BX_CHECK_ARG(S >= 0 && S <= BX_MAX_K, BX_E_RANGE, "S=%d outside [0, %d]", S, BX_MAX_K);
  propose_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(sp, d_opt_x, d_opt_val, S, reinterpret_cast<const unsigned long long*>(d_best), d_out);
  BX_LAUNCH_CHECK();
  return 0;
}

// Masked pairwise kernel matrix - replaces: gp.py:20-23 cov (x1 [n1 x d], x2 [n2 x d] already divided by the length scale).
namespace bx {
__global__ void cov_kernel(const float* __restrict__ x1, int n1, const float* __restrict__ x2, int n2, int d,
                           const float* __restrict__ m1, const float* __restrict__ m2, float* out) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (size_t)n1 * n2) return;
  const int i = (int)(e / n2), j = (int)(e % n2);
  float s = 0.f;
  for (int k = 0; k < d; ++k) {
    const float df = x1[(size_t)i * d + k] - x2[(size_t)j * d + k];
    s = fmaf(df, df, s);
  }
  out[e] = expf(-s) * (m1[i] * m2[j]);  // gp.py:15-17 exp_quadratic * outer(mask1, mask2)
}
}  // namespace bx

extern "C" int32_t bx_cov(const float* d_x1, int32_t n1, const float* d_x2, int32_t n2, int32_t d, const float* d_m1,
                          const float* d_m2, float* d_out, void* stream) {
  BX_CHECK_ARG(d_x1 && d_x2 && d_m1 && d_m2 && d_out, BX_E_ARG, "null pointer argument");
  BX_CHECK_ARG(n1 > 0 && n2 > 0 && d > 0, BX_E_ARG, "sizes must be positive");
  const size_t tot = (size_t)n1 * n2;
  bx::cov_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_x1, n1, d_x2, n2, d, d_m1, d_m2, d_out);
  BX_LAUNCH_CHECK();
  return 0;
}
