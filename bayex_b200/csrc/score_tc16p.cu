// tcgen05 scoring engine, fp16 head/tail split + CTA pair (cta_group::2): the production engine for large candidate sets.
//
// Replaces, for M generated candidates: domain.py:74,130,159-161 (Threefry draws) -> gp.py:60-64 (K*) -> gp.py:66-67 (mean)
// -> gp.py:68-70 (V = L^-1 K*, diag variance, std) -> acq.py closed forms -> optimizer.py:205 (arg-max).
//
// What the ncu capture of the 1-CTA fp16 engine (profiles/r1_score_tc16_ncu_full.csv) said, and what this engine changes:
//   * the shared-memory data pipe was 88 % busy: 45 % broadcast LDS.128 of observation rows by the K* producers
//     (2.3 wavefronts each, one candidate per thread), 32 % tensor-core operand reads, 7 % STS.
//       -> producers register-tile RC candidates per thread, so every observation row read feeds RC entries
//          (LDS per generated entry / RC);
//       -> cta_group::2: each CTA keeps only half (128 rows) of every T tile, the pair's tensor cores share it
//          (operand reads per MMA 12 KB -> 8 KB per SM, T traffic L2 -> smem halved);
//   * the FP32 pipe was 48 % busy (FADD2/FFMA2 occupy it for two cycles: packing saves issue slots, not lanes) and is the
//     next wall; the per-entry epilogue was trimmed (one cvt.rn.f16x2 per pair, amp * 2^a folded into one multiply).
//   * observation rows are staged by bulk-async copies (TMA, mbarrier complete_tx) instead of LDG -> STS -> bar.sync.
//   * registers are re-partitioned with setmaxnreg: 2 producer warpgroups get 176, the epilogue 96, the control warps 56.
//
// Warp roles (512 threads): 0 TMA (T tiles) | 1 MMA issuer (leader CTA) | 2 TMA (observation rows) | 3 idle |
//                           4-7 epilogue (TMEM lane quarter = warp & 3) | 8-15 K* producers.
// Cross-CTA protocol (leader = cluster rank 0), as in score_tc2.cu:
//   full_a[s]   (leader, count 16)  <- producer warps of both CTAs
//   empty_a[s]  (both, count 1)     <- tcgen05.commit.multicast from the leader
//   full_b[s]   (leader, count 2)   <- leader arrive.expect_tx + remote arrive; both CTAs' TMA loads complete_tx there
//   empty_b[s], tmem_full[a] (both) <- tcgen05.commit.multicast;  tmem_empty[a] (leader, count 8) <- epilogue warps
//   full_u[s] / empty_u[s] (CTA-local) <- bulk copy complete_tx / 8 producer warps
#include "tc_common.cuh"
#include "topk_fused.cuh"

namespace bx {
namespace tc16p {

using namespace bx::tc;

constexpr int BN = 256;                    // T rows per accumulator = UMMA N
constexpr int HN = BN / 2;                 // rows of every T tile held by one CTA
// Ring depths.  A stage of K* is refilled only after the MMAs that read it have completed, so with P = producer time and
// Tm = MMA time per K-block (both ~3000 clk) a 2-deep ring runs at (P + Tm + latencies) / 2 per K-block; 3 stages reach
// max(P, Tm).  T tiles: 3 stages = 1.5 K-blocks of TMA look-ahead.  Observation rows: 2 stages (profiles/
// r1c_ring_depth_sweep.csv: all of 3/3/3, 4/2/3, 4/2/2, 3/2/3, 3/3/2 within 0.6 % - the kernel is energy-bound).
#ifndef BX_P_SA
#define BX_P_SA 3
#endif
#ifndef BX_P_SB
#define BX_P_SB 3
#endif
#ifndef BX_P_NU
#define BX_P_NU 2
#endif
#ifndef BX_P_WAIT_NS
#define BX_P_WAIT_NS 64
#endif
constexpr int SA = BX_P_SA, SB = BX_P_SB, NU = BX_P_NU;
constexpr int HB = 128;                    // accurate mode: rows per half of a row block = UMMA N; each CTA holds HB / 2 of them
#ifndef BX_ACC_CHJ
#define BX_ACC_CHJ 1
#endif
constexpr int CHJ = BX_ACC_CHJ;            // accurate mode: K-blocks summed inside the tensor core before a chunk is folded
constexpr int KPR = BN / BKH;              // K-blocks per row block
constexpr int ATILE = TM * BKH * 2;        // 16 KiB: 128 candidates x 64 halves
constexpr int BTILE = HN * BKH * 2;        // 16 KiB: 128 T rows x 64 halves
constexpr int NTHREADS = 512, PROD0 = 8 * 32, NPROD = 256;
constexpr int MAXOP = 8;                   // most observation parts (producer threads per candidate group)
constexpr int UXP = 4;                     // fast mode: floats appended to a staged observation row (-|u'|^2, alpha, 2 pad)

constexpr bool share_coords(int d4, bool acc) { return acc ? d4 <= 28 : d4 <= 24; }

struct __align__(16) Smem {
  char a_hi[SA][ATILE], a_lo[SA][ATILE];
  char b_hi[SB][BTILE], b_lo[SB][BTILE];
  float mean_s[MAXOP][TM];
  unsigned long long full_a[SA], empty_a[SA], full_b[SB], empty_b[SB], full_u[NU], empty_u[NU], tmem_full[2],
      tmem_empty[2], mean_full, mean_empty;
  uint32_t tmem_base;
  tk::Shared topk;  // fused per-CTA top-k (epilogue warps)
};

// ------------------------------------------------------------------------------------------------------------------
// RC = candidates per producer thread.  128 / RC candidate groups x (2 RC) observation parts = 256 producer threads;
// a thread generates RC candidates x (32 / RC) observations of every K-block.
//
// ACC = true is the ACCURATE mode of the same engine (score_tc16a_kernel below): tcgen05.mma adds into its fp32 accumulator
// with truncation (toward zero, once per instruction, every aligned addend at a quarter ulp: scripts/dev/umma_probe.py), so a
// chain of n / 16 * 3 instructions biases V by up to 6e-5 relative.  Here no sum longer than ONE K-block (12 instructions)
// is formed inside the tensor core: a row block (256 rows of T) is processed as two halves of 128 rows, each K-block's
// product lands in its own 128-column chunk accumulator X[h] (TMEM columns 256 + 128 h), and the epilogue warps add the
// chunk into a running sum R (TMEM columns 0..255) with round-to-nearest FADDs - tcgen05.ld X, tcgen05.ld R, add,
// tcgen05.st R - while the tensor core fills the other half's chunk.  That is the Ootomo-Yokota correction ("accumulate
// outside the tensor core") with the running sums resident in TMEM; measured on real T / K* data it brings the error of
// sum v^2 to the level of an fp32 FFMA loop (30-44x below the plain chain).  Price: one row block in flight instead of
// two (K* is regenerated 8.5 n / 256... times instead of 4.5), N = 128 MMAs (the A operand is read twice).
template <int D4, int RC, bool ACC>
__global__ void __launch_bounds__(NTHREADS, 1)
score_tc16p_kernel(const __grid_constant__ ScoreArgs a, const __grid_constant__ SamplerPack sp,
                   const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo) {
  constexpr int NACC = ACC ? 1 : 2;  // row blocks in flight
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  Smem& S = *reinterpret_cast<Smem*>(smem_raw);
  // Staged K-block of observations.  Accurate mode: 64 rows of U + 64 alpha (two bulk copies).  Fast mode: 64 rows of the
  // centred copy built per launch by ux_rows_kernel - [D4 coordinates u' = u - mid | -|u'|^2 | alpha | 2 pad] - one copy.
  constexpr int UR = ACC ? D4 : D4 + UXP;                 // floats per staged row
  constexpr int US = ACC ? BKH * D4 + BKH : BKH * UR;     // floats per staged K-block
  const uint32_t us_addr = smem_u32(smem_raw) + (uint32_t)sizeof(Smem);  // [NU][US]
  // scaled candidate coordinates of the current tile [TM][D4], drawn ONCE per tile by the 256 producer threads together:
  // a producer thread needs the RC x D4 coordinates of its candidates, and the 2 RC threads that share a candidate group
  // used to draw them redundantly - at n = 512 those Threefry calls were half of the producers' instructions
  constexpr bool SHARE_COORDS = share_coords(D4, ACC);  // the widest rows do not leave room for the buffer
  const uint32_t cs_addr = us_addr + (uint32_t)(NU * US * 4);
  if ((smem_u32(smem_raw) & 1023u) != 0u) __trap();

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_rank();
  const bool leader = (rank == 0);
  const int np = a.f.np;
  const int nRB = (np + BN - 1) / BN;
  const int nJtot = np / BKH;
  const int nSB = (nRB + NACC - 1) / NACC;
  // list mode (a.in_list): the launch scores rows in_list[0 .. *in_count) - the candidates the fast mode handed over
  const uint64_t total = a.in_list ? (uint64_t)*a.in_count : a.m_count;
  const uint64_t ntiles = (total + TM - 1) / TM;
  const uint64_t npairs = (ntiles + 1) / 2;  // pair-tiles of 256 candidates
  const uint64_t pair0 = blockIdx.x >> 1, pstride = gridDim.x >> 1;

  if (tid == 0) {
    for (int i = 0; i < SA; ++i) { mbar_init(smem_u32(&S.full_a[i]), 2 * (NPROD / 32)); mbar_init(smem_u32(&S.empty_a[i]), 1); }
    for (int i = 0; i < SB; ++i) { mbar_init(smem_u32(&S.full_b[i]), 2); mbar_init(smem_u32(&S.empty_b[i]), 1); }
    for (int i = 0; i < NU; ++i) { mbar_init(smem_u32(&S.full_u[i]), 1); mbar_init(smem_u32(&S.empty_u[i]), NPROD / 32); }
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&S.tmem_full[i]), 1); mbar_init(smem_u32(&S.tmem_empty[i]), 8); }
    mbar_init(smem_u32(&S.mean_full), NPROD / 32);
    mbar_init(smem_u32(&S.mean_empty), 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair(smem_u32(&S.tmem_base), 512);
  tc_fence_before();
  cluster_sync_all();  // barriers of both CTAs initialised before anyone arrives remotely
  tc_fence_after();
  const uint32_t tmem = S.tmem_base;

  if (warp < 4) {
    reg_dec<56>();
    if (warp == 0 && lane == 0) {
      // ===== TMA: this CTA's 128-row half of every T tile (head + tail) =====
      uint32_t it = 0;
      if constexpr (ACC) {
        // rows of a T tile held by this CTA: 64 of each 128-row half (the pair's N = 128 MMA takes 64 rows from each CTA)
        for (uint64_t pt = pair0; pt < npairs; pt += pstride) {
          for (int rb = 0; rb < nRB; ++rb) {
            for (int J = 0; J < min(KPR * (rb + 1), nJtot); ++J, ++it) {
              const uint32_t st = it % SB, ph = (it / SB) & 1;
              mbar_wait<256>(smem_u32(&S.empty_b[st]), ph ^ 1);
              const uint32_t bar = smem_u32(&S.full_b[st]);
              if (leader) mbar_expect_tx(bar, 4 * BTILE);
              else mbar_arrive_rank(bar, 0);
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const int row = rb * BN + h * HB + (int)rank * (HB / 2);
                tma_load_2d_pair(smem_u32(S.b_hi[st]) + (uint32_t)(h * (BTILE / 2)), &map_hi, bar, J * BKH, row);
                tma_load_2d_pair(smem_u32(S.b_lo[st]) + (uint32_t)(h * (BTILE / 2)), &map_lo, bar, J * BKH, row);
              }
            }
          }
        }
      } else {
        for (uint64_t pt = pair0; pt < npairs; pt += pstride) {
          for (int sb = 0; sb < nSB; ++sb) {
            const int rb0 = sb * NACC, rb1 = min(rb0 + NACC, nRB);
            for (int J = 0; J < min(KPR * rb1, nJtot); ++J) {
              for (int rb = max(rb0, J / KPR); rb < rb1; ++rb, ++it) {
                const uint32_t st = it % SB, ph = (it / SB) & 1;
                mbar_wait<256>(smem_u32(&S.empty_b[st]), ph ^ 1);
                const uint32_t bar = smem_u32(&S.full_b[st]);
                if (leader) mbar_expect_tx(bar, 4 * BTILE);  // both halves, head + tail
                else mbar_arrive_rank(bar, 0);
                tma_load_2d_pair(smem_u32(S.b_hi[st]), &map_hi, bar, J * BKH, rb * BN + (int)rank * HN);
                tma_load_2d_pair(smem_u32(S.b_lo[st]), &map_lo, bar, J * BKH, rb * BN + (int)rank * HN);
              }
            }
          }
        }
      }
    } else if (warp == 1 && leader && lane == 0) {
      // ===== MMA issuer (leader CTA only) =====
      if constexpr (ACC) {
        const uint32_t idesc = make_idesc_f16(HB, 2 * TM);
        uint32_t it = 0, xph = 0;  // bit h of xph = phase of chunk accumulator X[h]
        for (uint64_t pt = pair0; pt < npairs; pt += pstride) {
          for (int rb = 0; rb < nRB; ++rb) {
            for (int J = 0; J < min(KPR * (rb + 1), nJtot); ++J, ++it) {
              const uint32_t sa = it % SA, pa = (it / SA) & 1, st = it % SB, pb = (it / SB) & 1;
              mbar_wait_cluster<32>(smem_u32(&S.full_a[sa]), pa);
              mbar_wait<32>(smem_u32(&S.full_b[st]), pb);
              tc_fence_after();
              const uint64_t ah = make_desc(smem_u32(S.a_hi[sa])), al = make_desc(smem_u32(S.a_lo[sa]));
              const bool c_first = (J % CHJ == 0), c_last = (J % CHJ == CHJ - 1) || (J == min(KPR * (rb + 1), nJtot) - 1);
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                if (c_first) {
                  mbar_wait_cluster(smem_u32(&S.tmem_empty[h]), ((xph >> h) & 1) ^ 1);  // the epilogues folded X[h] into R
                  tc_fence_after();
                }
                const uint64_t bh = make_desc(smem_u32(S.b_hi[st]) + (uint32_t)(h * (BTILE / 2)));
                const uint64_t bl = make_desc(smem_u32(S.b_lo[st]) + (uint32_t)(h * (BTILE / 2)));
                const uint32_t d = tmem + (uint32_t)(BN + h * HB);
                if (!(a.debug & 2))
#pragma unroll
                for (int ks = 0; ks < BKH / 16; ++ks) {
                  const uint64_t o = (uint64_t)(ks * 2);
                  umma_f16_pair(d, ah + o, bh + o, idesc, !(c_first && ks == 0));  // a chunk holds at most CHJ K-blocks
                  umma_f16_pair(d, al + o, bh + o, idesc, 1u);
                  umma_f16_pair(d, ah + o, bl + o, idesc, 1u);
                }
                if (c_last) {
                  umma_commit_pair(smem_u32(&S.tmem_full[h]));
                  xph ^= 1u << h;
                }
              }
              umma_commit_pair(smem_u32(&S.empty_b[st]));
              umma_commit_pair(smem_u32(&S.empty_a[sa]));
            }
          }
        }
      } else {
      const uint32_t idesc = make_idesc_f16(BN, 2 * TM);
      uint32_t ita = 0, itb = 0, acc_ph = 0;  // bit a of acc_ph = phase of accumulator a
      for (uint64_t pt = pair0; pt < npairs; pt += pstride) {
        for (int sb = 0; sb < nSB; ++sb) {
          const int rb0 = sb * NACC, rb1 = min(rb0 + NACC, nRB);
          for (int J = 0; J < min(KPR * rb1, nJtot); ++J, ++ita) {
            const uint32_t sa = ita % SA, pa = (ita / SA) & 1;
            mbar_wait_cluster<32>(smem_u32(&S.full_a[sa]), pa);  // back-off: the hot spin took 27 % of this SMSP's issue slots
            tc_fence_after();
            const uint64_t ah = make_desc(smem_u32(S.a_hi[sa])), al = make_desc(smem_u32(S.a_lo[sa]));
            for (int rb = max(rb0, J / KPR); rb < rb1; ++rb, ++itb) {
              const int acc = rb - rb0;
              if (J == 0) {  // first write of this accumulator in this super-block: the epilogues must have drained it
                mbar_wait_cluster(smem_u32(&S.tmem_empty[acc]), ((acc_ph >> acc) & 1) ^ 1);
                tc_fence_after();
              }
              const uint32_t st = itb % SB, pb = (itb / SB) & 1;
              mbar_wait<32>(smem_u32(&S.full_b[st]), pb);
              tc_fence_after();
              const uint64_t bh = make_desc(smem_u32(S.b_hi[st])), bl = make_desc(smem_u32(S.b_lo[st]));
              const uint32_t d = tmem + (uint32_t)(acc * BN);
              if (!(a.debug & 2))
#pragma unroll
              for (int ks = 0; ks < BKH / 16; ++ks) {  // UMMA K = 16 halves = 32 bytes inside the swizzle atom
                const uint64_t o = (uint64_t)(ks * 2);
                umma_f16_pair(d, ah + o, bh + o, idesc, (J | ks) != 0);
                umma_f16_pair(d, al + o, bh + o, idesc, 1u);
                umma_f16_pair(d, ah + o, bl + o, idesc, 1u);
              }
              umma_commit_pair(smem_u32(&S.empty_b[st]));
              if (J == min(KPR * rb + KPR, nJtot) - 1) {  // last K-block below the diagonal: this row block of V is complete
                umma_commit_pair(smem_u32(&S.tmem_full[acc]));
                acc_ph ^= 1u << acc;
              }
            }
            umma_commit_pair(smem_u32(&S.empty_a[sa]));
          }
        }
      }
      }
    } else if (warp == 2 && lane == 0) {
      // ===== TMA: observation rows (scaled coordinates) + alpha of every K-block, NU-deep ring =====
      uint32_t itu = 0;
      for (uint64_t pt = pair0; pt < npairs; pt += pstride) {
        for (int sb = 0; sb < nSB; ++sb) {
          const int nJ = min(KPR * min(sb * NACC + NACC, nRB), nJtot);
          for (int J = 0; J < nJ; ++J, ++itu) {
            const uint32_t su = itu % NU, ph = (itu / NU) & 1;
            mbar_wait<256>(smem_u32(&S.empty_u[su]), ph ^ 1);
            const uint32_t bar = smem_u32(&S.full_u[su]), dst = us_addr + su * (uint32_t)(US * 4);
            mbar_expect_tx(bar, (uint32_t)(US * 4));
            if constexpr (ACC) {
              bulk_load_1d(dst, a.f.d_U + (size_t)J * BKH * D4, (uint32_t)(BKH * D4 * 4), bar);
              bulk_load_1d(dst + (uint32_t)(BKH * D4 * 4), a.f.d_alpha + (size_t)J * BKH, (uint32_t)(BKH * 4), bar);
            } else {
              bulk_load_1d(dst, a.ux + (size_t)J * BKH * UR, (uint32_t)(US * 4), bar);
            }
          }
        }
      }
    }
  } else if (warp < 8) {
    reg_dec<96>();
    // ===== epilogue (both CTAs, own 128 candidates): TMEM -> sum of squares -> std -> acquisition -> arg-max =====
    constexpr int NOP = 2 * RC;
    const int q = warp & 3, cl = q * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
    const float amp = a.f.d_hyp[HYP_AMP], ymean = a.f.d_hyp[HYP_YMEAN];
    const float ksc = kstar_scale(amp);
    const float unscale = 1.f / (a.f.d_hyp[HYP_T16SCALE] * ksc);  // operands were scaled by powers of two
    const float unscale2 = unscale * unscale, mean_unscale = 1.f / ksc;
    uint32_t acc_ph = 0, tiles_done = 0;
    unsigned long long best = 0ull;
    // fused top-k (optimizer.py:196): this CTA's slab of survivors; named barrier 2 = the four epilogue warps
    const int et = tid - 4 * 32;
    const int topk_k = a.topk_k;
    unsigned long long* const slab = topk_k ? a.topk_slab + (size_t)blockIdx.x * tk::CAP : nullptr;
    if (topk_k) {
      tk::init(S.topk, et);
      tk::bar<2>();
    }
    for (uint64_t pt = pair0; pt < npairs; pt += pstride, ++tiles_done) {
      const uint64_t tile = 2 * pt + rank;
      float ss = 0.f;
      if constexpr (ACC) {
        // fold every K-block's chunk X[h] into the running sum R with round-to-nearest adds; the last chunk of a row
        // block is squared and summed instead of stored (its running sum is complete: V for 128 rows of T)
        for (int rb = 0; rb < nRB; ++rb) {
          const int nCh = (min(KPR * (rb + 1), nJtot) + CHJ - 1) / CHJ;
          for (int ch = 0; ch < nCh; ++ch) {
            const bool first = (ch == 0), last = (ch == nCh - 1);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              mbar_wait(smem_u32(&S.tmem_full[h]), (acc_ph >> h) & 1);  // hot: one wait per 768 CHJ clocks of MMAs, no back-off
              acc_ph ^= 1u << h;
              tc_fence_after();
              // The three cases are separate straight-line loops: with `first` / `last` tested per element the fold ran to
              // ~930 SASS instructions per warp (SEL / FSEL around every add: ncu source page, profiles/r2q) - five times
              // the work - and the fold, not the tensor pipe (18 % active), set the pace of the accurate mode.
              // (A software-pipelined version on 16-column groups was measured slower: more TMEM instructions.)
              const uint32_t xa = lane_addr + (uint32_t)(BN + h * HB), ra = lane_addr + (uint32_t)(h * HB);
              if (first && !last) {         // R = X
#pragma unroll
                for (int c0 = 0; c0 < HB; c0 += 32) {
                  uint32_t x[32];
                  tmem_ld32_nowait(xa + (uint32_t)c0, x);
                  tmem_ld_wait();
                  tmem_st32(ra + (uint32_t)c0, x);
                }
              } else if (!last) {           // R += X   (round to nearest)
#pragma unroll
                for (int c0 = 0; c0 < HB; c0 += 32) {
                  uint32_t x[32], r[32];
                  tmem_ld32_nowait(xa + (uint32_t)c0, x);
                  tmem_ld32_nowait(ra + (uint32_t)c0, r);
                  tmem_ld_wait();
#pragma unroll
                  for (int i = 0; i < 32; ++i) x[i] = __float_as_uint(__uint_as_float(r[i]) + __uint_as_float(x[i]));
                  tmem_st32(ra + (uint32_t)c0, x);
                }
              } else if (!first) {          // V = R + X complete for these 128 rows: square and sum
#pragma unroll
                for (int c0 = 0; c0 < HB; c0 += 32) {
                  uint32_t x[32], r[32];
                  tmem_ld32_nowait(xa + (uint32_t)c0, x);
                  tmem_ld32_nowait(ra + (uint32_t)c0, r);
                  tmem_ld_wait();
#pragma unroll
                  for (int i = 0; i < 32; ++i) {
                    const float v = __uint_as_float(r[i]) + __uint_as_float(x[i]);
                    ss = fmaf(v, v, ss);
                  }
                }
              } else {                      // a row block with a single chunk
#pragma unroll
                for (int c0 = 0; c0 < HB; c0 += 32) {
                  uint32_t x[32];
                  tmem_ld32_nowait(xa + (uint32_t)c0, x);
                  tmem_ld_wait();
#pragma unroll
                  for (int i = 0; i < 32; ++i) ss = fmaf(__uint_as_float(x[i]), __uint_as_float(x[i]), ss);
                }
              }
              if (!last) tmem_st_wait();
              tc_fence_before();
              __syncwarp();
              if (lane == 0) {
                // what the MMA issuer must see is ordered by the tcgen05 fences (TMEM only): the remote arrive needs no
                // cluster-scope release (that form compiles to MEMBAR.ALL.GPU - once per 768 clocks here)
                if (leader) mbar_arrive(smem_u32(&S.tmem_empty[h]));
                else mbar_arrive_rank_relaxed(smem_u32(&S.tmem_empty[h]), 0);
              }
            }
          }
        }
      } else
      for (int sb = 0; sb < nSB; ++sb) {
        const int rb0 = sb * NACC, rb1 = min(rb0 + NACC, nRB);
        for (int acc = 0; acc < rb1 - rb0; ++acc) {
          mbar_wait<128>(smem_u32(&S.tmem_full[acc]), (acc_ph >> acc) & 1);
          acc_ph ^= 1u << acc;
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
      float msum = 0.f;
#pragma unroll
      for (int o = 0; o < NOP; ++o) msum += S.mean_s[o][cl];
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&S.mean_empty));
      const float mu = msum * mean_unscale + ymean;  // gp.py:67
      const uint64_t lrow = tile * TM + cl;  // position in this launch; list mode: the row it stands for
      const bool valid = lrow < total;
      const uint64_t row = (a.in_list && valid) ? (uint64_t)a.in_list[lrow] : lrow;
      const float var = amp - ss * unscale2;       // gp.py:69 (diagonal only)
      // tcgen05.mma adds into its fp32 accumulator with truncation (toward zero, once per instruction, and every aligned
      // addend at a quarter ulp: scripts/dev/umma_probe.py), which biases sum v^2 by 1e-6 .. 6e-5 relative.  That is
      // harmless while var is a sizeable part of amp and ruinous where amp - sum v^2 cancels, so candidates with
      // var < tau * amp are handed to the fp32 SIMT engine (launch_score_tc16p runs it on the list) and get no output here.
      const bool hand_over = valid && a.flag_list != nullptr && var < a.flag_tau * amp;
      const unsigned hm = __ballot_sync(0xffffffffu, hand_over);
      if (hm) {
        unsigned int base = 0;
        if (lane == 0) base = atomicAdd(a.flag_count, (unsigned)__popc(hm));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (hand_over) a.flag_list[base + __popc(hm & ((1u << lane) - 1u))] = (uint32_t)row;
      }
      if (valid && !hand_over) {
        const float sd = sqrtf(fmaxf(var, 1e-10f));  // gp.py:70
        const float av = acquisition(a.acq_id, mu, sd, a.ystar, a.acq_param);
        if (a.mu) a.mu[row] = mu;
        if (a.sigma) a.sigma[row] = sd;
        if (a.acq) a.acq[row] = av;
        const unsigned long long key = pack_best(av, (uint32_t)(a.index_base + row));
        best = key > best ? key : best;
        if (topk_k) tk::push(S.topk, slab, tk::topk_key(av, (uint32_t)(a.index_base + row)));
      }
      if (topk_k) tk::end_tile<2>(S.topk, slab, topk_k, et);
    }
    if (topk_k) tk::finish<2>(S.topk, slab, topk_k, et, a.topk_count + blockIdx.x);
    if (a.best) {
      best = warp_max_u64(best);
      if (lane == 0 && best) atomicMax(a.best, best);
    }
  } else {
    reg_inc<176>();
    // ===== K* producers (both CTAs): RC candidates x OPT observations per thread and K-block =====
    constexpr int NCG = TM / RC;         // candidate groups
    constexpr int OPT = BKH / (2 * RC);  // observations per thread per K-block (32 / RC)
    constexpr int OB = (16 / RC < 8) ? 16 / RC : 8;  // observations in flight: RC * OB = 16 independent FMA chains
    const int pth = tid - PROD0, cg = pth % NCG, op = pth / NCG;  // op is warp-uniform
    const float amp = a.f.d_hyp[HYP_AMP];
    const float ampk = amp * kstar_scale(amp);  // exact (power-of-two scale): entries land near 2^14
    // Fast mode: coordinates are centred on the midpoint of the observations' bounding box (tail of a.ux), and when the
    // norms are small the squared distance is formed as |c'|^2 + |u'|^2 - 2 c'.u' - one packed FMA per two dimensions
    // instead of an add and an FMA (the producers, not the tensor pipe, set the kernel's pace).  The price is rounding at
    // the magnitude of the norms instead of the distance: an absolute error of ~2^-24 (|c'|^2 + |u'|^2) in the exponent,
    // i.e. a relative one in K*; so the form is used only while the largest possible |c'|^2 + |u'|^2 (from the domain box
    // and the largest |u'|^2) stays below a.expand_smax (64: error below 4e-6, 1e-6 typical, against the 2^-22 the
    // head/tail split itself carries).  Generated candidates only: explicit points have no box.
    const float* ux_tail = ACC ? nullptr : a.ux + (size_t)np * UR;  // [D4] mid, then max |u'|^2
    bool expand = false;
    if constexpr (!ACC) {
      if (!a.cand && a.expand_smax > 0.f) {
        float cb = 0.f;
        for (int k = 0; k < a.f.d; ++k) {
          const DimSampler& ds = sp.dim[k];
          const float lo = ds.kind == 0 ? ds.lo : (float)ds.ilo, hi = ds.kind == 0 ? ds.hi : (float)ds.ilo + (float)(ds.span - 1u);
          const float a0 = lo * a.f.d_scale[k] - ux_tail[k], a1 = hi * a.f.d_scale[k] - ux_tail[k];
          cb += fmaxf(a0 * a0, a1 * a1);
        }
        expand = (cb + ux_tail[D4] <= a.expand_smax);  // the same on every thread of the grid
      }
    }
    uint32_t ita = 0, itu = 0, tiles_done = 0;
    for (uint64_t pt = pair0; pt < npairs; pt += pstride, ++tiles_done) {
      const uint64_t tile = 2 * pt + rank;
      uint64_t ncr[RC][D4 / 2];  // (-c_k, -c_{k+1}) packed: the squared distance runs on add.f32x2 / fma.f32x2
      if constexpr (SHARE_COORDS) {
        for (int e = pth; e < TM * D4; e += NPROD) {
          const int c = e / D4, k = e % D4;
          float v = 0.f;
          if (k < a.f.d) {
            v = cand_coord(a, sp, tile * TM + c, k) * a.f.d_scale[k];
            if constexpr (!ACC) v -= ux_tail[k];
          }
          sts32(cs_addr + 4u * (uint32_t)e, -v);
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");  // producers only; the K loop below keeps tiles apart
#pragma unroll
        for (int j = 0; j < RC; ++j)
#pragma unroll
          for (int k = 0; k < D4; k += 4) {
            uint64_t c01, c23;
            lds_2x64(cs_addr + (uint32_t)(((cg + j * NCG) * D4 + k) * 4), c01, c23);
            ncr[j][k / 2] = c01;
            ncr[j][k / 2 + 1] = c23;
          }
      } else {
#pragma unroll
        for (int j = 0; j < RC; ++j) {
          const uint64_t grow = tile * TM + cg + j * NCG;
#pragma unroll
          for (int k = 0; k < D4; k += 2) {
            float c0 = (k < a.f.d) ? cand_coord(a, sp, grow, k) * a.f.d_scale[k] : 0.f;
            float c1 = (k + 1 < a.f.d) ? cand_coord(a, sp, grow, k + 1) * a.f.d_scale[k + 1] : 0.f;
            if constexpr (!ACC) {
              if (k < a.f.d) c0 -= ux_tail[k];
              if (k + 1 < a.f.d) c1 -= ux_tail[k + 1];
            }
            ncr[j][k / 2] = pack2(-c0, -c1);
          }
        }
      }
      float mpart[RC];
#pragma unroll
      for (int j = 0; j < RC; ++j) mpart[j] = 0.f;
      float ncn[RC];  // -|c'_j|^2 (expanded form)
#pragma unroll
      for (int j = 0; j < RC; ++j) {
        uint64_t q = 0ull;
        if (expand) {
#pragma unroll
          for (int k = 0; k < D4 / 2; ++k) q = fma2(ncr[j][k], ncr[j][k], q);
        }
        float qx, qy;
        unpack2(q, qx, qy);
        ncn[j] = -(qx + qy);
      }
      for (int sb = 0; sb < nSB; ++sb) {
        const bool last_sb = (sb == nSB - 1);
        const int nJ = min(KPR * min(sb * NACC + NACC, nRB), nJtot);
        for (int J = 0; J < nJ; ++J, ++ita, ++itu) {
          const uint32_t su = itu % NU, pu = (itu / NU) & 1;
          const uint32_t u = us_addr + su * (uint32_t)(US * 4) + (uint32_t)(op * OPT * UR * 4);  // this thread's rows
          const uint32_t al_addr = us_addr + su * (uint32_t)(US * 4) + (uint32_t)((BKH * D4 + op * OPT) * 4);  // accurate mode
          mbar_wait(smem_u32(&S.full_u[su]), pu);
          const uint32_t sa = ita % SA, pa = (ita / SA) & 1;
          // under the power cap the producers run ahead of the tensor pipe and wait here ~10 % of the time: back off
          // instead of spinning (ncu: 25 try_wait iterations per K-block and warp, ~10 % of the producers' instructions)
          mbar_wait<BX_P_WAIT_NS>(smem_u32(&S.empty_a[sa]), pa ^ 1);
          const uint32_t a_hi = smem_u32(S.a_hi[sa]), a_lo = smem_u32(S.a_lo[sa]);
          if (!(a.debug & 1))
#pragma unroll
          for (int g = 0; g < OPT / 8; ++g) {  // 8 observations -> one 16-byte chunk (8 halves) of RC rows of the A tile
            float kv[RC][8];
            float al[8];  // alpha of these 8 observations (the mean rides along in the last super-block)
#pragma unroll
            for (int b = 0; b < 8 / OB; ++b) {
              uint64_t sq[RC][OB];  // one packed accumulator per (candidate, observation): lanes = even / odd dimensions
#pragma unroll
              for (int j = 0; j < RC; ++j)
#pragma unroll
                for (int i = 0; i < OB; ++i) sq[j][i] = 0ull;
              if (!ACC && expand) {
#pragma unroll
                for (int k4 = 0; k4 < D4; k4 += 4) {
#pragma unroll
                  for (int i = 0; i < OB; ++i) {
                    uint64_t u01, u23;
                    lds_2x64(u + (uint32_t)(((g * 8 + b * OB + i) * UR + k4) * 4), u01, u23);
#pragma unroll
                    for (int j = 0; j < RC; ++j) {
                      sq[j][i] = fma2(u01, ncr[j][k4 / 2], sq[j][i]);  // -c'.u'
                      sq[j][i] = fma2(u23, ncr[j][k4 / 2 + 1], sq[j][i]);
                    }
                  }
                }
#pragma unroll
                for (int i = 0; i < OB; ++i) {
                  uint64_t na;
                  asm volatile("ld.shared.b64 %0, [%1];" : "=l"(na) : "r"(u + (uint32_t)(((g * 8 + b * OB + i) * UR + D4) * 4)));
                  float nun;
                  unpack2(na, nun, al[b * OB + i]);
#pragma unroll
                  for (int j = 0; j < RC; ++j) {
                    float sx, sy;
                    unpack2(sq[j][i], sx, sy);
                    kv[j][b * OB + i] = ampk * ex2_approx(fmaf(-2.f, sx + sy, ncn[j] + nun));  // gp.py:64, times 2^a
                  }
                }
              } else {
#pragma unroll
                for (int k4 = 0; k4 < D4; k4 += 4) {
#pragma unroll
                  for (int i = 0; i < OB; ++i) {
                    uint64_t u01, u23;
                    lds_2x64(u + (uint32_t)(((g * 8 + b * OB + i) * UR + k4) * 4), u01, u23);
#pragma unroll
                    for (int j = 0; j < RC; ++j) {
                      const uint64_t d01 = add2(u01, ncr[j][k4 / 2]), d23 = add2(u23, ncr[j][k4 / 2 + 1]);
                      sq[j][i] = fma2(d01, d01, sq[j][i]);
                      sq[j][i] = fma2(d23, d23, sq[j][i]);
                    }
                  }
                }
#pragma unroll
                for (int j = 0; j < RC; ++j)
#pragma unroll
                  for (int i = 0; i < OB; ++i) {
                    float sx, sy;
                    unpack2(sq[j][i], sx, sy);
                    kv[j][b * OB + i] = ampk * ex2_approx(-sx - sy);  // gp.py:64, times 2^a
                  }
                if constexpr (!ACC) {
                  if (last_sb) {
#pragma unroll
                    for (int i = 0; i < OB; ++i) al[b * OB + i] = lds32(u + (uint32_t)(((g * 8 + b * OB + i) * UR + D4 + 1) * 4));
                  }
                }
              }
            }
            if (last_sb) {  // gp.py:66-67: the last super-block visits every K-block, so the mean rides along there
              if constexpr (ACC) {
                const float4 a0 = lds128(al_addr + (uint32_t)(g * 32)), a1 = lds128(al_addr + (uint32_t)(g * 32 + 16));
                al[0] = a0.x; al[1] = a0.y; al[2] = a0.z; al[3] = a0.w; al[4] = a1.x; al[5] = a1.y; al[6] = a1.z; al[7] = a1.w;
              }
#pragma unroll
              for (int j = 0; j < RC; ++j)
#pragma unroll
                for (int i = 0; i < 8; ++i) mpart[j] = fmaf(al[i], kv[j][i], mpart[j]);
            }
#pragma unroll
            for (int j = 0; j < RC; ++j) split_store_f16(a_hi, a_lo, cg + j * NCG, op * (OPT / 8) + g, kv[j]);
          }
          fence_proxy_async();  // generic-proxy stores -> visible to the (leader's) tensor core
          __syncwarp();
          if (lane == 0) {
            if (leader) mbar_arrive(smem_u32(&S.full_a[sa]));
            else mbar_arrive_rank_relaxed(smem_u32(&S.full_a[sa]), 0);
            mbar_arrive(smem_u32(&S.empty_u[su]));
          }
        }
      }
      mbar_wait(smem_u32(&S.mean_empty), (tiles_done & 1) ^ 1);
#pragma unroll
      for (int j = 0; j < RC; ++j) S.mean_s[op][cg + j * NCG] = mpart[j];
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

// ---- the fast mode's copy of the observations, rebuilt in the caller's workspace by every fast-mode launch (two launches
// of a few microseconds): rows [d4 coordinates u' = u - mid | -|u'|^2 | alpha | 0 | 0], then mid[d4], then max |u'|^2.
// mid = midpoint of the observations' bounding box in scaled units: centring keeps the norms - and with them the
// rounding of the expanded squared distance - as small as a translation can make them.
__global__ void __launch_bounds__(256) ux_mid_kernel(const bx_factor_t f, float* __restrict__ ux) {
  __shared__ float smin[8], smax[8];
  const int k = blockIdx.x, tid = threadIdx.x;
  float mn = CUDART_INF_F, mx = -CUDART_INF_F;
  for (int i = tid; i < f.n; i += 256) {
    const float v = f.d_U[(size_t)i * f.d4 + k];
    mn = fminf(mn, v);
    mx = fmaxf(mx, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((tid & 31) == 0) { smin[tid >> 5] = mn; smax[tid >> 5] = mx; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < 8; ++w) { mn = fminf(mn, smin[w]); mx = fmaxf(mx, smax[w]); }
    float* tail = ux + (size_t)f.np * (f.d4 + UXP);
    tail[k] = (k < f.d) ? 0.5f * (mn + mx) : 0.f;
    if (k == 0) tail[f.d4] = 0.f;  // max |u'|^2, raised by ux_rows_kernel
  }
}

__global__ void __launch_bounds__(256) ux_rows_kernel(const bx_factor_t f, float* __restrict__ ux) {
  const int i = blockIdx.x * 256 + threadIdx.x, ur = f.d4 + UXP;
  if (i >= f.np) return;
  float* tail = ux + (size_t)f.np * ur;
  float* row = ux + (size_t)i * ur;
  float un = 0.f;
  for (int k = 0; k < f.d4; ++k) {
    const float c = (i < f.n && k < f.d) ? f.d_U[(size_t)i * f.d4 + k] - tail[k] : 0.f;  // padding rows meet zero columns of T
    row[k] = c;
    un = fmaf(c, c, un);
  }
  row[f.d4] = -un;
  row[f.d4 + 1] = (i < f.n) ? f.d_alpha[i] : 0.f;
  row[f.d4 + 2] = 0.f;
  row[f.d4 + 3] = 0.f;
  atomicMax(reinterpret_cast<unsigned int*>(tail + f.d4), __float_as_uint(un));  // un >= 0: the bit patterns order like the values
}

static unsigned pair_grid(uint64_t m_count, int sms) {
  const uint64_t ntiles = (m_count + TM - 1) / TM, npairs = (ntiles + 1) / 2;
  const uint64_t max_pairs = (uint64_t)(sms / 2);
  return 2u * (unsigned)(npairs < max_pairs ? npairs : max_pairs);
}

template <int D4, int RC, bool ACC>
static int32_t launch_d4(cudaStream_t s, const ScoreArgs& a, const SamplerPack& sp, const CUtensorMap& mh, const CUtensorMap& ml) {
  const size_t us = ACC ? BKH * D4 + BKH : BKH * (D4 + UXP);  // floats per staged K-block of observations (see the kernel)
  const size_t smem = sizeof(Smem) + sizeof(float) * NU * us + (share_coords(D4, ACC) ? sizeof(float) * TM * D4 : 0);
  BX_CUDA(cudaFuncSetAttribute(score_tc16p_kernel<D4, RC, ACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int dev = 0, sms = 0;
  BX_CUDA(cudaGetDevice(&dev));
  BX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const unsigned grid = pair_grid(a.m_count, sms);  // list mode: sized for the whole launch, CTAs beyond the list idle
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
  BX_CUDA(cudaLaunchKernelEx(&cfg, score_tc16p_kernel<D4, RC, ACC>, a, sp, mh, ml));
  return 0;
}

template <bool ACC>
static int32_t launch_any(cudaStream_t s, const ScoreArgs& a, const SamplerPack& sp, const CUtensorMap& mh, const CUtensorMap& ml) {
  // RC = 4 candidates per producer thread while their coordinates fit the 176-register budget (d <= 20), else 2
  switch (a.f.d4) {
    case 4: return launch_d4<4, 4, ACC>(s, a, sp, mh, ml);
    case 8: return launch_d4<8, 4, ACC>(s, a, sp, mh, ml);
    case 12: return launch_d4<12, 4, ACC>(s, a, sp, mh, ml);
    case 16: return launch_d4<16, 4, ACC>(s, a, sp, mh, ml);
    case 20: return launch_d4<20, 4, ACC>(s, a, sp, mh, ml);
    case 24: return launch_d4<24, 2, ACC>(s, a, sp, mh, ml);
    case 28: return launch_d4<28, 2, ACC>(s, a, sp, mh, ml);
    case 32: return launch_d4<32, 2, ACC>(s, a, sp, mh, ml);
  }
  set_error("tcgen05 engines support d <= 32 (d4=%d)", a.f.d4);
  return BX_E_UNSUPPORTED;
}

}  // namespace tc16p

unsigned score_tc16p_grid(uint64_t m_count) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return tc16p::pair_grid(m_count, sms);
}

bool score_tc16_supported(const bx_factor_t& f) { return f.d_T16hi && f.d_T16lo && f.d4 <= 32 && f.np <= BX_UX_MAX_NP; }

size_t score_tc16_ux_bytes() { return sizeof(float) * ((size_t)BX_UX_MAX_NP * (32 + tc16p::UXP) + 32 + 4); }

// accurate = false: the fast mode (two row blocks in flight, sums over whole rows inside the tensor core);
// accurate = true: chunk accumulators folded into TMEM-resident running sums by the epilogue (fp32-level accuracy)
int32_t launch_score_tc16p(cudaStream_t s, const ScoreArgs& a_in, const SamplerPack& sp, bool accurate) {
  ScoreArgs a = a_in;
  if (const char* e = getenv("BX_TC_ABLATE")) a.debug = atoi(e);
  CUtensorMap mh, ml;
  const uint32_t box_rows = accurate ? tc16p::HB / 2 : tc16p::HN;
  int32_t rc = tc::make_map_f16(&mh, a.f.d_T16hi, a.f.np, a.f.np, box_rows);
  if (rc) return rc;
  rc = tc::make_map_f16(&ml, a.f.d_T16lo, a.f.np, a.f.np, box_rows);
  if (rc) return rc;
  if (accurate) return tc16p::launch_any<true>(s, a, sp, mh, ml);
  if (!a.ux) {
    set_error("the fast mode of the fp16 tcgen05 engine needs its workspace");
    return BX_E_WORKSPACE;
  }
  // largest |c'|^2 + |u'|^2 for which the expanded squared distance is used (0: never); see the producers
  static const float smax = [] { const char* e = getenv("BX_EXPAND_SMAX"); return e ? (float)atof(e) : 64.f; }();
  a.expand_smax = smax;
  tc16p::ux_mid_kernel<<<a.f.d4, 256, 0, s>>>(a.f, a.ux);
  BX_LAUNCH_CHECK();
  tc16p::ux_rows_kernel<<<(a.f.np + 255) / 256, 256, 0, s>>>(a.f, a.ux);
  BX_LAUNCH_CHECK();
  return tc16p::launch_any<false>(s, a, sp, mh, ml);
}

}  // namespace bx
