// tcgen05 scoring engine, 3xTF32, CTA-pair version (cta_group::2, BX_ENGINE_TCGEN05_PAIR): where the pair protocol was
// developed; kept selectable.  Production: score_tc16p.cu (the same protocol with fp16 head/tail operands).
//
// Same algorithm as score_tc.cu (see there), but two CTAs of a cluster - the two SMs of one TPC - cooperate on one
// UMMA M=256 x N=256 tile: each CTA owns 128 candidates (its own generated K* tile, its own 512 TMEM columns, its own
// epilogue) and only HALF of every T tile (128 of the 256 rows).  The leader CTA issues the MMAs, the tensor core reads
// the other half of the B operand from the peer's shared memory.  Per SM this halves the shared-memory operand reads
// (the measured limiter of the 1-CTA engine: 8-12 KB of smem per 64-128 MMA cycles saturates the 128 B/clk port), halves
// the L2 -> smem traffic for T, and lets the T pipeline be 4 stages deep in the same 128 KB.
//
// Cross-CTA protocol (leader = cluster rank 0):
//   full_a[s]   (leader, count 16)  <- producer warps of both CTAs (remote mbarrier.arrive for rank 1)
//   empty_a[s]  (both, count 1)     <- tcgen05.commit.multicast from the leader
//   full_b[s]   (leader, count 2)   <- leader arrive.expect_tx(64 KB) + remote arrive from rank 1; both CTAs' TMA
//                                      loads complete_tx on the leader's barrier (peer bit cleared)
//   empty_b[s], tmem_full[a] (both) <- tcgen05.commit.multicast
//   tmem_empty[a] (leader, count 8) <- epilogue warps of both CTAs
#include "tc_common.cuh"

namespace bx {
namespace tc2 {

using namespace bx::tc;

constexpr int BN = 256;                    // T rows per accumulator = UMMA N
constexpr int HN = BN / 2;                 // rows of every T tile held by one CTA
constexpr int NACC = 2, SA = 2, SB = 4;
constexpr int KPR = BN / BK;
constexpr int NTHREADS = 448, PROD0 = 6 * 32, NPROD = 256;

struct __align__(16) Smem {
  char a_hi[SA][TILE_BYTES], a_lo[SA][TILE_BYTES];
  char b_hi[SB][TILE_BYTES], b_lo[SB][TILE_BYTES];  // this CTA's 128-row half of each T tile
  float mean_s[2][TM];
  unsigned long long full_a[SA], empty_a[SA], full_b[SB], empty_b[SB], tmem_full[NACC], tmem_empty[NACC], mean_full, mean_empty;
  uint32_t tmem_base;
};

template <int D4>
__global__ void __launch_bounds__(NTHREADS, 1)
score_tc2_kernel(const __grid_constant__ ScoreArgs a, const __grid_constant__ SamplerPack sp,
                 const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  Smem& S = *reinterpret_cast<Smem*>(smem_raw);
  constexpr int US = BK * D4 + BK;
  const uint32_t us_addr = smem_u32(smem_raw) + (uint32_t)sizeof(Smem);
  if ((smem_u32(smem_raw) & 1023u) != 0u) __trap();

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_rank();
  const bool leader = (rank == 0);
  const int np = a.f.np;
  const int nRB = (np + BN - 1) / BN;
  const int nJtot = np / BK;
  const int nSB = (nRB + NACC - 1) / NACC;
  const uint64_t ntiles = (a.m_count + TM - 1) / TM;
  const uint64_t npairs = (ntiles + 1) / 2;            // pair-tiles of 256 candidates
  const uint64_t pair0 = blockIdx.x >> 1, pstride = gridDim.x >> 1;

  if (tid == 0) {
    for (int i = 0; i < SA; ++i) { mbar_init(smem_u32(&S.full_a[i]), 2 * (NPROD / 32)); mbar_init(smem_u32(&S.empty_a[i]), 1); }
    for (int i = 0; i < SB; ++i) { mbar_init(smem_u32(&S.full_b[i]), 2); mbar_init(smem_u32(&S.empty_b[i]), 1); }
    for (int i = 0; i < NACC; ++i) { mbar_init(smem_u32(&S.tmem_full[i]), 1); mbar_init(smem_u32(&S.tmem_empty[i]), 8); }
    mbar_init(smem_u32(&S.mean_full), NPROD / 32);
    mbar_init(smem_u32(&S.mean_empty), 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair(smem_u32(&S.tmem_base), 512);
  tc_fence_before();
  cluster_sync_all();  // barriers of both CTAs initialised before anyone arrives remotely
  tc_fence_after();
  const uint32_t tmem = S.tmem_base;

  if (warp == 0) {
    // ===== TMA producer: this CTA's 128-row half of every T tile (head + tail) =====
    if (lane == 0) {
      uint32_t it = 0;
      for (uint64_t pt = pair0; pt < npairs; pt += pstride) {
        for (int sb = 0; sb < nSB; ++sb) {
          const int rb0 = sb * NACC, rb1 = min(rb0 + NACC, nRB);
          for (int J = 0; J < min(KPR * rb1, nJtot); ++J) {
            for (int rb = max(rb0, J / KPR); rb < rb1; ++rb, ++it) {
              const uint32_t st = it % SB, ph = (it / SB) & 1;
              mbar_wait<64>(smem_u32(&S.empty_b[st]), ph ^ 1);
              const uint32_t bar = smem_u32(&S.full_b[st]);
              if (leader) mbar_expect_tx(bar, 4 * TILE_BYTES);  // both halves, head + tail
              else mbar_arrive_rank(bar, 0);
              tma_load_2d_pair(smem_u32(S.b_hi[st]), &map_hi, bar, J * BK, rb * BN + (int)rank * HN);
              tma_load_2d_pair(smem_u32(S.b_lo[st]), &map_lo, bar, J * BK, rb * BN + (int)rank * HN);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA only) =====
    if (leader && lane == 0) {
      const uint32_t idesc = make_idesc(BN, 2 * TM);
      uint32_t ita = 0, itb = 0, acc_use[NACC] = {};
      for (uint64_t pt = pair0; pt < npairs; pt += pstride) {
        for (int sb = 0; sb < nSB; ++sb) {
          const int rb0 = sb * NACC, rb1 = min(rb0 + NACC, nRB);
          for (int J = 0; J < min(KPR * rb1, nJtot); ++J, ++ita) {
            const uint32_t sa = ita % SA, pa = (ita / SA) & 1;
            mbar_wait_cluster(smem_u32(&S.full_a[sa]), pa);
            tc_fence_after();
            const uint64_t ah = make_desc(smem_u32(S.a_hi[sa])), al = make_desc(smem_u32(S.a_lo[sa]));
            for (int rb = max(rb0, J / KPR); rb < rb1; ++rb, ++itb) {
              const int acc = rb - rb0;
              if (J == 0) {
                mbar_wait_cluster(smem_u32(&S.tmem_empty[acc]), (acc_use[acc] & 1) ^ 1);
                tc_fence_after();
              }
              const uint32_t st = itb % SB, pb = (itb / SB) & 1;
              mbar_wait(smem_u32(&S.full_b[st]), pb);
              tc_fence_after();
              const uint64_t bh = make_desc(smem_u32(S.b_hi[st])), bl = make_desc(smem_u32(S.b_lo[st]));
              const uint32_t d = tmem + (uint32_t)(acc * BN);
              if (!(a.debug & 2))
#pragma unroll
              for (int ks = 0; ks < BK / 8; ++ks) {
                const uint64_t o = (uint64_t)(ks * 2);
                umma_tf32_pair(d, ah + o, bh + o, idesc, (J | ks) != 0);
                umma_tf32_pair(d, al + o, bh + o, idesc, 1u);
                umma_tf32_pair(d, ah + o, bl + o, idesc, 1u);
              }
              umma_commit_pair(smem_u32(&S.empty_b[st]));
              if (J == min(KPR * rb + KPR, nJtot) - 1) {
                umma_commit_pair(smem_u32(&S.tmem_full[acc]));
                ++acc_use[acc];
              }
            }
            umma_commit_pair(smem_u32(&S.empty_a[sa]));
          }
        }
      }
    }
  } else if (warp < 6) {
    // ===== epilogue (both CTAs, own 128 candidates) =====
    const int q = warp & 3, cl = q * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
    const float amp = a.f.d_hyp[HYP_AMP], ymean = a.f.d_hyp[HYP_YMEAN];
    uint32_t acc_use[NACC] = {}, tiles_done = 0;
    unsigned long long best = 0ull;
    for (uint64_t pt = pair0; pt < npairs; pt += pstride, ++tiles_done) {
      const uint64_t tile = 2 * pt + rank;
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
          if (lane == 0) {
            if (leader) mbar_arrive(smem_u32(&S.tmem_empty[acc]));
            else mbar_arrive_rank(smem_u32(&S.tmem_empty[acc]), 0);
          }
        }
      }
      mbar_wait<128>(smem_u32(&S.mean_full), tiles_done & 1);
      const float mu = S.mean_s[0][cl] + S.mean_s[1][cl] + ymean;
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&S.mean_empty));
      const uint64_t row = tile * TM + cl;
      if (row < a.m_count) {
        const float var = amp - ss;
        const float sd = sqrtf(fmaxf(var, 1e-10f));
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
    // ===== K* producers (both CTAs): 2 threads per candidate, 16 observations each per K-block =====
    const int pth = tid - PROD0, c = pth & (TM - 1), h = pth >> 7;
    const float amp = a.f.d_hyp[HYP_AMP];
    uint32_t ita = 0, tiles_done = 0;
    for (uint64_t pt = pair0; pt < npairs; pt += pstride, ++tiles_done) {
      const uint64_t tile = 2 * pt + rank;
      uint64_t ncr[D4 / 2];
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
        for (int e = pth; e < US; e += NPROD)
          sts32(us_addr + 4u * e, (e < BK * D4) ? a.f.d_U[e] : a.f.d_alpha[e - BK * D4]);
        asm volatile("bar.sync 1, 256;" ::: "memory");
        for (int J = 0; J < nJ; ++J, ++ita) {
          const uint32_t u = us_addr + (uint32_t)((J & 1) * US * 4);
          float pre[(US + NPROD - 1) / NPROD];
          if (J + 1 < nJ) {
#pragma unroll
            for (int i = 0; i < (US + NPROD - 1) / NPROD; ++i) {
              const int e = pth + i * NPROD;
              pre[i] = (e < BK * D4) ? __ldg(a.f.d_U + (size_t)(J + 1) * BK * D4 + e)
                                     : (e < US ? __ldg(a.f.d_alpha + (J + 1) * BK + e - BK * D4) : 0.f);
            }
          }
          const uint32_t sa = ita % SA, pa = (ita / SA) & 1;
          mbar_wait(smem_u32(&S.empty_a[sa]), pa ^ 1);
          const uint32_t a_hi = smem_u32(S.a_hi[sa]), a_lo = smem_u32(S.a_lo[sa]);
          if (!(a.debug & 1))
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float kv[4];
            const int kk0 = h * 16 + g * 4;
            uint64_t s01[4] = {0ull, 0ull, 0ull, 0ull}, s23[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
            for (int k4 = 0; k4 < D4; k4 += 4) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                uint64_t u01, u23;
                lds_2x64(u + (uint32_t)(((kk0 + i) * D4 + k4) * 4), u01, u23);
                const uint64_t d01 = add2(u01, ncr[k4 / 2]), d23 = add2(u23, ncr[k4 / 2 + 1]);
                s01[i] = fma2(d01, d01, s01[i]);
                s23[i] = fma2(d23, d23, s23[i]);
              }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float sx, sy;
              unpack2(add2(s01[i], s23[i]), sx, sy);
              kv[i] = amp * ex2_approx(-sx - sy);
              if (last_sb) mpart = fmaf(lds32(u + (uint32_t)((BK * D4 + kk0 + i) * 4)), kv[i], mpart);
            }
            split_store(a_hi, a_lo, c, h * 4 + g, kv[0], kv[1], kv[2], kv[3]);
          }
          fence_proxy_async();  // generic-proxy stores -> visible to the (leader's) tensor core
          __syncwarp();
          if (lane == 0) {
            if (leader) mbar_arrive(smem_u32(&S.full_a[sa]));
            else mbar_arrive_rank(smem_u32(&S.full_a[sa]), 0);
          }
          if (J + 1 < nJ) {
            const uint32_t un = us_addr + (uint32_t)(((J + 1) & 1) * US * 4);
#pragma unroll
            for (int i = 0; i < (US + NPROD - 1) / NPROD; ++i) {
              const int e = pth + i * NPROD;
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
  __syncwarp();
  tc_fence_before();
  cluster_sync_all();  // no CTA may exit (or free TMEM) while its peer can still touch its smem / barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem, 512);
  }
}

template <int D4>
static int32_t launch_d4(cudaStream_t s, const ScoreArgs& a, const SamplerPack& sp, const CUtensorMap& mh, const CUtensorMap& ml) {
  const size_t smem = sizeof(Smem) + sizeof(float) * 2 * (BK * D4 + BK);
  BX_CUDA(cudaFuncSetAttribute(score_tc2_kernel<D4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int dev = 0, sms = 0;
  BX_CUDA(cudaGetDevice(&dev));
  BX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const uint64_t ntiles = (a.m_count + TM - 1) / TM, npairs = (ntiles + 1) / 2;
  const uint64_t max_pairs = (uint64_t)(sms / 2);
  const unsigned grid = 2u * (unsigned)(npairs < max_pairs ? npairs : max_pairs);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(NTHREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  BX_CUDA(cudaLaunchKernelEx(&cfg, score_tc2_kernel<D4>, a, sp, mh, ml));
  return 0;
}

}  // namespace tc2

bool score_tc_supported(const bx_factor_t& f) { return f.d_Thi && f.d_Tlo && f.d4 <= 32; }

int32_t launch_score_tc2(cudaStream_t s, const ScoreArgs& a_in, const SamplerPack& sp) {
  ScoreArgs a = a_in;
  if (const char* e = getenv("BX_TC_ABLATE")) a.debug = atoi(e);
  CUtensorMap mh, ml;
  int32_t rc = tc::make_map(&mh, a.f.d_Thi, a.f.np, a.f.np, a.f.np, tc2::HN);
  if (rc) return rc;
  rc = tc::make_map(&ml, a.f.d_Tlo, a.f.np, a.f.np, a.f.np, tc2::HN);
  if (rc) return rc;
  switch (a.f.d4) {
    case 4: return tc2::launch_d4<4>(s, a, sp, mh, ml);
    case 8: return tc2::launch_d4<8>(s, a, sp, mh, ml);
    case 12: return tc2::launch_d4<12>(s, a, sp, mh, ml);
    case 16: return tc2::launch_d4<16>(s, a, sp, mh, ml);
    case 20: return tc2::launch_d4<20>(s, a, sp, mh, ml);
    case 24: return tc2::launch_d4<24>(s, a, sp, mh, ml);
    case 28: return tc2::launch_d4<28>(s, a, sp, mh, ml);
    case 32: return tc2::launch_d4<32>(s, a, sp, mh, ml);
  }
  set_error("tcgen05 engine supports d <= 32 (d4=%d)", a.f.d4);
  return BX_E_UNSUPPORTED;
}

}  // namespace bx
