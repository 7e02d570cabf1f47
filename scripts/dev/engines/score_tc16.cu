// tcgen05 scoring engine with an fp16 head/tail split ("3xFP16"), one CTA per tile (BX_ENGINE_TCGEN05_F16): same pipeline
// as score_tc.cu, half the tensor work.  Kept selectable; production: score_tc16p.cu.
//
// x = hi + lo with hi = fp16(x), lo = fp16(x - hi) carries 22 significand bits - the same as the TF32 head/tail split -
// but kind::f16 MMAs run at twice the TF32 rate and the operands are half as wide, so tensor time, shared-memory operand
// traffic and L2 -> smem traffic all halve.  fp16's narrow exponent is handled with two power-of-two scales: K* * 2^a
// (a from the amplitude, so the largest entry sits near 2^14) and T * 2^b (b from max|T|, computed when the factor is
// built); fp32 TMEM accumulators hold V * 2^(a+b) and the epilogue multiplies sum(v^2) by 2^(-2(a+b)) - exact.
// One K-block is 64 observations (64 fp16 = 128 B = one SWIZZLE_128B row); UMMA K = 16.
#include <cuda_fp16.h>

#include "tc_common.cuh"

namespace bx {
namespace tc16 {

using namespace bx::tc;

constexpr int BN = 256, NACC = 2, SA = 2, SB = 2;
constexpr int KPR = BN / BKH;
constexpr int BTILE_BYTES = BN * BKH * 2;  // 32 KiB
constexpr int NTHREADS = 448, PROD0 = 6 * 32, NPROD = 256;

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

struct __align__(16) Smem {
  // operand tiles first (1024-byte aligned)
  char a_hi[SA][TILE_BYTES], a_lo[SA][TILE_BYTES];
  char b_hi[SB][BTILE_BYTES], b_lo[SB][BTILE_BYTES];
  float mean_s[2][TM];
  unsigned long long full_a[SA], empty_a[SA], full_b[SB], empty_b[SB], tmem_full[NACC], tmem_empty[NACC], mean_full, mean_empty;
  uint32_t tmem_base;
};

// ------------------------------------------------------------------------------------------------------------------
template <int D4>
__global__ void __launch_bounds__(NTHREADS, 1)
score_tc16_kernel(const __grid_constant__ ScoreArgs a, const __grid_constant__ SamplerPack sp,
                const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  Smem& S = *reinterpret_cast<Smem*>(smem_raw);  // SWIZZLE_128B tiles need the 1024-byte alignment declared above
  constexpr int US = BKH * D4 + BKH;
  // [2][US] observation rows + alpha of the current K-block, right behind the struct (sizeof(Smem) % 16 == 0)
  const uint32_t us_addr = smem_u32(smem_raw) + (uint32_t)sizeof(Smem);
  if ((smem_u32(smem_raw) & 1023u) != 0u) __trap();

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int np = a.f.np;
  const int nRB = (np + BN - 1) / BN;       // row blocks of T (rows past np are zero-filled by TMA)
  const int nJtot = np / BKH;                // K-blocks in total
  const int nSB = (nRB + NACC - 1) / NACC;  // super-blocks
  const uint64_t ntiles = (a.m_count + TM - 1) / TM;

  if (tid == 0) {
    for (int i = 0; i < SA; ++i) { mbar_init(smem_u32(&S.full_a[i]), NPROD / 32); mbar_init(smem_u32(&S.empty_a[i]), 1); }
    for (int i = 0; i < SB; ++i) { mbar_init(smem_u32(&S.full_b[i]), 1); mbar_init(smem_u32(&S.empty_b[i]), 1); }
    for (int i = 0; i < NACC; ++i) { mbar_init(smem_u32(&S.tmem_full[i]), 1); mbar_init(smem_u32(&S.tmem_empty[i]), 4); }
    mbar_init(smem_u32(&S.mean_full), NPROD / 32);
    mbar_init(smem_u32(&S.mean_empty), 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&S.tmem_base), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = S.tmem_base;

  if (warp == 0) {
    // ===== TMA producer: T tiles (head + tail) =====
    if (lane == 0) {
      uint32_t it = 0;
      for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int sb = 0; sb < nSB; ++sb) {
          const int rb0 = sb * NACC, rb1 = min(rb0 + NACC, nRB);
          for (int J = 0; J < min(KPR * rb1, nJtot); ++J) {
            for (int rb = max(rb0, J / KPR); rb < rb1; ++rb, ++it) {
              const uint32_t st = it % SB, ph = (it / SB) & 1;
              mbar_wait<64>(smem_u32(&S.empty_b[st]), ph ^ 1);
              const uint32_t bar = smem_u32(&S.full_b[st]);
              mbar_expect_tx(bar, 2 * BTILE_BYTES);
              tma_load_2d(smem_u32(S.b_hi[st]), &map_hi, bar, J * BKH, rb * BN);
              tma_load_2d(smem_u32(S.b_lo[st]), &map_lo, bar, J * BKH, rb * BN);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(BN);
      uint32_t ita = 0, itb = 0, acc_use[NACC] = {};
      for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int sb = 0; sb < nSB; ++sb) {
          const int rb0 = sb * NACC, rb1 = min(rb0 + NACC, nRB);
          for (int J = 0; J < min(KPR * rb1, nJtot); ++J, ++ita) {
            const uint32_t sa = ita % SA, pa = (ita / SA) & 1;
            mbar_wait(smem_u32(&S.full_a[sa]), pa);
            tc_fence_after();
            const uint64_t ah = make_desc(smem_u32(S.a_hi[sa])), al = make_desc(smem_u32(S.a_lo[sa]));
            for (int rb = max(rb0, J / KPR); rb < rb1; ++rb, ++itb) {
              const int acc = rb - rb0;
              if (J == 0) {  // first write of this accumulator in this super-block: wait until the epilogue drained it
                mbar_wait(smem_u32(&S.tmem_empty[acc]), (acc_use[acc] & 1) ^ 1);
                tc_fence_after();
              }
              const uint32_t st = itb % SB, pb = (itb / SB) & 1;
              mbar_wait(smem_u32(&S.full_b[st]), pb);
              tc_fence_after();
              const uint64_t bh = make_desc(smem_u32(S.b_hi[st])), bl = make_desc(smem_u32(S.b_lo[st]));
              const uint32_t d = tmem + (uint32_t)(acc * BN);
              if (!(a.debug & 2))
#pragma unroll
              for (int ks = 0; ks < BKH / 16; ++ks) {  // UMMA K = 16 for f16 = 32 bytes inside the swizzle atom
                const uint64_t o = (uint64_t)(ks * 2);
                umma_f16(d, ah + o, bh + o, idesc, (J | ks) != 0);
                umma_f16(d, al + o, bh + o, idesc, 1u);
                umma_f16(d, ah + o, bl + o, idesc, 1u);
              }
              umma_commit(smem_u32(&S.empty_b[st]));
              if (J == min(KPR * rb + KPR, nJtot) - 1) {  // last K-block below the diagonal: this row block of V is complete
                umma_commit(smem_u32(&S.tmem_full[acc]));
                ++acc_use[acc];
              }
            }
            umma_commit(smem_u32(&S.empty_a[sa]));
          }
        }
      }
    }
  } else if (warp < 6) {
    // ===== epilogue: TMEM -> sum of squares -> std -> acquisition -> arg-max =====
    const int q = warp & 3, cl = q * 32 + lane;  // TMEM lane quarter of this warp; candidate within the tile
    const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
    const float amp = a.f.d_hyp[HYP_AMP], ymean = a.f.d_hyp[HYP_YMEAN];
    const float unscale = 1.f / (a.f.d_hyp[HYP_T16SCALE] * kstar_scale(amp));  // operands were scaled by powers of two
    const float unscale2 = unscale * unscale;
    uint32_t acc_use[NACC] = {}, tiles_done = 0;
    unsigned long long best = 0ull;
    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tiles_done) {
      float ss = 0.f;
      for (int sb = 0; sb < nSB; ++sb) {
        const int rb0 = sb * NACC, rb1 = min(rb0 + NACC, nRB);
        for (int acc = 0; acc < rb1 - rb0; ++acc) {
          mbar_wait<128>(smem_u32(&S.tmem_full[acc]), acc_use[acc] & 1);
          ++acc_use[acc];
          tc_fence_after();
#pragma unroll
          for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t r[32];
            tmem_ld32(lane_addr + (uint32_t)(acc * BN + c0), r);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float v = __uint_as_float(r[i]);
              ss = fmaf(v, v, ss);
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&S.tmem_empty[acc]));
        }
      }
      mbar_wait<128>(smem_u32(&S.mean_full), tiles_done & 1);
      const float mu = S.mean_s[0][cl] + S.mean_s[1][cl] + ymean;  // gp.py:67
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&S.mean_empty));
      const uint64_t row = tile * TM + cl;
      if (row < a.m_count) {
        const float var = amp - ss * unscale2;         // gp.py:69 (diagonal only)
        const float sd = sqrtf(fmaxf(var, 1e-10f));    // gp.py:70
        const float av = acquisition(a.acq_id, mu, sd, a.ystar, a.acq_param);
        if (a.mu) a.mu[row] = mu;
        if (a.sigma) a.sigma[row] = sd;
        if (a.acq) a.acq[row] = av;
        const unsigned long long key = pack_best(av, (uint32_t)(a.index_base + row));
        best = key > best ? key : best;
      }
    }
    if (a.best) {
      best = warp_max_u64(best);
      if (lane == 0 && best) atomicMax(a.best, best);
    }
  } else {
    // ===== K* producers: 2 threads per candidate, 16 observations each per K-block =====
    const int pt = tid - PROD0, c = pt & (TM - 1), h = pt >> 7;
    const float amp = a.f.d_hyp[HYP_AMP];
    const float ksc = kstar_scale(amp);
    uint32_t ita = 0, tiles_done = 0;
    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tiles_done) {
      uint64_t ncr[D4 / 2];  // (-c_k, -c_{k+1}) packed: the squared distance runs on add.f32x2 / fma.f32x2
      {
        const uint64_t gidx = a.m_begin + tile * TM + c;
#pragma unroll
        for (int k = 0; k < D4; k += 2) {
          const float c0 = (k < a.f.d) ? sample_coord(sp.dim[k], gidx) * a.f.d_scale[k] : 0.f;
          const float c1 = (k + 1 < a.f.d) ? sample_coord(sp.dim[k + 1], gidx) * a.f.d_scale[k + 1] : 0.f;
          ncr[k / 2] = pack2(-c0, -c1);
        }
      }
      float mpart = 0.f;
      for (int sb = 0; sb < nSB; ++sb) {
        const int rb1 = min(sb * NACC + NACC, nRB);
        const bool last_sb = (sb == nSB - 1);
        const int nJ = min(KPR * rb1, nJtot);
        // stage observation rows (+ alpha) of K-block 0
        for (int e = pt; e < US; e += NPROD)
          sts32(us_addr + 4u * e, (e < BKH * D4) ? a.f.d_U[e] : a.f.d_alpha[e - BKH * D4]);
        asm volatile("bar.sync 1, 256;" ::: "memory");
        for (int J = 0; J < nJ; ++J, ++ita) {
          const uint32_t u = us_addr + (uint32_t)((J & 1) * US * 4);
          float pre[(US + NPROD - 1) / NPROD];
          if (J + 1 < nJ) {  // prefetch the next K-block's rows into registers
#pragma unroll
            for (int i = 0; i < (US + NPROD - 1) / NPROD; ++i) {
              const int e = pt + i * NPROD;
              pre[i] = (e < BKH * D4) ? __ldg(a.f.d_U + (size_t)(J + 1) * BKH * D4 + e)
                                     : (e < US ? __ldg(a.f.d_alpha + (J + 1) * BKH + e - BKH * D4) : 0.f);
            }
          }
          const uint32_t sa = ita % SA, pa = (ita / SA) & 1;
          mbar_wait(smem_u32(&S.empty_a[sa]), pa ^ 1);
          const uint32_t a_hi = smem_u32(S.a_hi[sa]), a_lo = smem_u32(S.a_lo[sa]);
          if (!(a.debug & 1))
#pragma unroll
          for (int g = 0; g < 4; ++g) {  // 4 groups of 8 observations -> one 16-byte chunk (8 halves) of the A row each
            float kv[8];
            const int kk0 = h * 32 + g * 8;
#pragma unroll
            for (int b = 0; b < 2; ++b) {
              uint64_t s01[4] = {0ull, 0ull, 0ull, 0ull}, s23[4] = {0ull, 0ull, 0ull, 0ull};  // 16 independent chains
#pragma unroll
              for (int k4 = 0; k4 < D4; k4 += 4) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  uint64_t u01, u23;
                  lds_2x64(u + (uint32_t)(((kk0 + b * 4 + i) * D4 + k4) * 4), u01, u23);
                  const uint64_t d01 = add2(u01, ncr[k4 / 2]), d23 = add2(u23, ncr[k4 / 2 + 1]);
                  s01[i] = fma2(d01, d01, s01[i]);
                  s23[i] = fma2(d23, d23, s23[i]);
                }
              }
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                float sx, sy;
                unpack2(add2(s01[i], s23[i]), sx, sy);
                const float kval = amp * ex2_approx(-sx - sy);  // gp.py:64
                if (last_sb) mpart = fmaf(lds32(u + (uint32_t)((BKH * D4 + kk0 + b * 4 + i) * 4)), kval, mpart);  // gp.py:66-67
                kv[b * 4 + i] = kval * ksc;
              }
            }
            split_store_f16(a_hi, a_lo, c, h * 4 + g, kv);
          }
          fence_proxy_async();  // generic-proxy stores -> visible to the tensor core (async proxy)
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&S.full_a[sa]));
          if (J + 1 < nJ) {
            const uint32_t un = us_addr + (uint32_t)(((J + 1) & 1) * US * 4);
#pragma unroll
            for (int i = 0; i < (US + NPROD - 1) / NPROD; ++i) {
              const int e = pt + i * NPROD;
              if (e < US) sts32(un + 4u * e, pre[i]);
            }
          }
          asm volatile("bar.sync 1, 256;" ::: "memory");
        }
      }
      mbar_wait(smem_u32(&S.mean_empty), (tiles_done & 1) ^ 1);
      S.mean_s[h][c] = mpart;
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&S.mean_full));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}


template <int D4>
static int32_t launch_d4(cudaStream_t s, const ScoreArgs& a, const SamplerPack& sp, const CUtensorMap& mh, const CUtensorMap& ml) {
  const size_t smem = sizeof(Smem) + sizeof(float) * 2 * (BKH * D4 + BKH);
  BX_CUDA(cudaFuncSetAttribute(score_tc16_kernel<D4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int dev = 0, sms = 0;
  BX_CUDA(cudaGetDevice(&dev));
  BX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const uint64_t ntiles = (a.m_count + TM - 1) / TM;
  const unsigned grid = (unsigned)(ntiles < (uint64_t)sms ? ntiles : (uint64_t)sms);
  score_tc16_kernel<D4><<<grid, NTHREADS, smem, s>>>(a, sp, mh, ml);
  BX_LAUNCH_CHECK();
  return 0;
}

}  // namespace tc16

bool score_tc16_supported(const bx_factor_t& f) { return f.d_T16hi && f.d_T16lo && f.d4 <= 32; }

int32_t launch_score_tc16(cudaStream_t s, const ScoreArgs& a_in, const SamplerPack& sp) {
  ScoreArgs a = a_in;
  if (const char* e = getenv("BX_TC_ABLATE")) a.debug = atoi(e);
  CUtensorMap mh, ml;
  int32_t rc = tc::make_map_f16(&mh, a.f.d_T16hi, a.f.np, a.f.np, tc16::BN);
  if (rc) return rc;
  rc = tc::make_map_f16(&ml, a.f.d_T16lo, a.f.np, a.f.np, tc16::BN);
  if (rc) return rc;
  switch (a.f.d4) {
    case 4: return tc16::launch_d4<4>(s, a, sp, mh, ml);
    case 8: return tc16::launch_d4<8>(s, a, sp, mh, ml);
    case 12: return tc16::launch_d4<12>(s, a, sp, mh, ml);
    case 16: return tc16::launch_d4<16>(s, a, sp, mh, ml);
    case 20: return tc16::launch_d4<20>(s, a, sp, mh, ml);
    case 24: return tc16::launch_d4<24>(s, a, sp, mh, ml);
    case 28: return tc16::launch_d4<28>(s, a, sp, mh, ml);
    case 32: return tc16::launch_d4<32>(s, a, sp, mh, ml);
  }
  set_error("tcgen05 engines support d <= 32 (d4=%d)", a.f.d4);
  return BX_E_UNSUPPORTED;
}

}  // namespace bx
