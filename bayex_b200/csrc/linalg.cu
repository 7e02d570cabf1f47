// Dense fp32 linear algebra for the fit path (gp.py:46-49 and its gradient): SIMT GEMM, blocked Cholesky, recursive
// triangular inverse, triangular mat-vecs.  Tensor cores are deliberately not used here (north star: only the
// candidate contraction may use them).
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>

#include "internal.h"

namespace bx {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

// ------------------------------------------------------------------------------------------------------------------
// 64x64x16 SIMT GEMM.  256 threads, 4x4 register tile per thread.
// ------------------------------------------------------------------------------------------------------------------
constexpr int GT = 64, GK = 16, GP = GT + 4;

__device__ __forceinline__ void tree_node(int nb, int level, int z, int& lo, int& mid, int& hi) {
  lo = 0;
  hi = nb;
  for (int b = level - 1; b >= 0; --b) {
    int m = (lo + hi) >> 1;
    if ((z >> b) & 1) lo = m; else hi = m;
  }
  mid = (lo + hi) >> 1;
}

template <bool TA, bool TB>
__global__ void __launch_bounds__(256) gemm64_kernel(GemmArgs g) {
  __shared__ __align__(16) float As[GK][GP];
  __shared__ __align__(16) float Bs[GK][GP];
  const float* A = g.A;
  const float* B = g.B;
  float* C = g.C;
  int M = g.M, N = g.N, K = g.K, kmode = g.kmode;
  float alpha = g.alpha;
  if (g.tree_nb > 0) {
    int lo, mid, hi;
    tree_node(g.tree_nb, g.tree_level, blockIdx.z, lo, mid, hi);
    if (hi - lo < 2) return;
    if (g.tree_pass == 1) {  // TMP[mid:hi, lo:mid] = L[mid:hi, lo:mid] * T[lo:mid, lo:mid]
      A += (size_t)mid * NB * g.lda + (size_t)lo * NB;
      B += (size_t)lo * NB * g.ldb + (size_t)lo * NB;
      C += (size_t)mid * NB * g.ldc + (size_t)lo * NB;
      M = (hi - mid) * NB; N = K = (mid - lo) * NB; kmode = KM_B_LOWER; alpha = 1.f;
    } else {                 // T[mid:hi, lo:mid] = -T[mid:hi, mid:hi] * TMP[mid:hi, lo:mid]
      A += (size_t)mid * NB * g.lda + (size_t)mid * NB;
      B += (size_t)mid * NB * g.ldb + (size_t)lo * NB;
      C += (size_t)mid * NB * g.ldc + (size_t)lo * NB;
      M = K = (hi - mid) * NB; N = (mid - lo) * NB; kmode = KM_A_LOWER; alpha = -1.f;
    }
  }
  const int i0 = blockIdx.y * GT, j0 = blockIdx.x * GT;
  if (i0 >= M || j0 >= N) return;
  if (g.lower_only && j0 > i0) return;
  int k_lo = 0, k_hi = K;
  if (kmode == KM_A_LOWER) k_hi = min(K, i0 + GT);
  else if (kmode == KM_B_LOWER) k_lo = j0;
  else if (kmode == KM_TN_BOTH) k_lo = max(i0, j0);
  else if (kmode == KM_TN_A) k_lo = i0;
  else if (kmode == KM_BT_LOWER) k_hi = min(K, j0 + GT);

  if (g.ksplit > 0) {
    k_lo = max(k_lo, (int)blockIdx.z * g.ksplit);
    k_hi = min(k_hi, ((int)blockIdx.z + 1) * g.ksplit);
    C += (size_t)blockIdx.z * g.csplit;
  }
  const int tid = threadIdx.x, rg = tid & 15, cg = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;

  // Register-staged software pipeline: the global loads of k-step i+1 are in flight while k-step i is multiplied
  // (the fit / refinement GEMMs are skinny and latency-bound without it).
  const int a_r = TA ? (tid >> 4) : (tid >> 2), a_c = TA ? (tid & 15) * 4 : (tid & 3) * 4;
  const int b_r = TB ? (tid >> 2) : (tid >> 4), b_c = TB ? (tid & 3) * 4 : (tid & 15) * 4;
  auto load_a = [&](int k0) -> float4 {
    return TA ? *reinterpret_cast<const float4*>(A + (size_t)(k0 + a_r) * g.lda + i0 + a_c)
              : *reinterpret_cast<const float4*>(A + (size_t)(i0 + a_r) * g.lda + k0 + a_c);
  };
  auto load_b = [&](int k0) -> float4 {
    return TB ? *reinterpret_cast<const float4*>(B + (size_t)(j0 + b_r) * g.ldb + k0 + b_c)
              : *reinterpret_cast<const float4*>(B + (size_t)(k0 + b_r) * g.ldb + j0 + b_c);
  };
  float4 ra = make_float4(0.f, 0.f, 0.f, 0.f), rb = ra;
  if (k_lo < k_hi) { ra = load_a(k_lo); rb = load_b(k_lo); }
  for (int k0 = k_lo; k0 < k_hi; k0 += GK) {
    if (!TA) {  // A is [M x K]: transpose into As[k][row]
      As[a_c + 0][a_r] = ra.x; As[a_c + 1][a_r] = ra.y; As[a_c + 2][a_r] = ra.z; As[a_c + 3][a_r] = ra.w;
    } else {    // A is [K x M]
      *reinterpret_cast<float4*>(&As[a_r][a_c]) = ra;
    }
    if (!TB) {  // B is [K x N]
      *reinterpret_cast<float4*>(&Bs[b_r][b_c]) = rb;
    } else {    // B is [N x K]: transpose into Bs[k][col]
      Bs[b_c + 0][b_r] = rb.x; Bs[b_c + 1][b_r] = rb.y; Bs[b_c + 2][b_r] = rb.z; Bs[b_c + 3][b_r] = rb.w;
    }
    __syncthreads();
    if (k0 + GK < k_hi) { ra = load_a(k0 + GK); rb = load_b(k0 + GK); }
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][rg * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][cg * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(av[r], bv[c], acc[r][c]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    float* cp = C + (size_t)(i0 + rg * 4 + r) * g.ldc + j0 + cg * 4;
    float4 o = make_float4(alpha * acc[r][0], alpha * acc[r][1], alpha * acc[r][2], alpha * acc[r][3]);
    if (g.beta != 0.f) {
      const float4 c = *reinterpret_cast<const float4*>(cp);
      o.x = fmaf(g.beta, c.x, o.x); o.y = fmaf(g.beta, c.y, o.y); o.z = fmaf(g.beta, c.z, o.z); o.w = fmaf(g.beta, c.w, o.w);
    }
    *reinterpret_cast<float4*>(cp) = o;
  }
}


// ------------------------------------------------------------------------------------------------------------------
// 128x128x16 SIMT GEMM for the large products of the fit path (K^-1 = T^T T, the triangular-inverse tree, the
// Cholesky panel / trailing updates).  256 threads, 8x8 register tile per thread (two 4x4 blocks per dimension, 64
// apart, so every shared-memory read is one LDS.128), double-buffered shared tiles, register-staged global loads
// of the next k-step in flight under the FMAs.  Extents are multiples of 64: edge tiles are guarded per float4.
// Same K-range restrictions (triangular operands) and tree-batched mode as gemm64_kernel.
// ------------------------------------------------------------------------------------------------------------------
constexpr int HT = 128, HK = 16, HP = HT + 4;

// NT = 256: 8x8 outputs per thread, two CTAs per SM (many-tile launches).  NT = 512: 4x8 outputs per thread, one CTA of
// 16 warps per SM - a lone 8-warp CTA reaches only ~50 % of the SM's FMA rate, and in single-wave launches (n <= 4096:
// the inverse tree, T^T T) the launch lasts as long as its heaviest tile.
template <bool TA, bool TB, int NT>
__global__ void __launch_bounds__(NT, 512 / NT) gemm128_kernel(GemmArgs g) {
  constexpr int RT = 2048 / NT;       // rows per thread: 8 or 4
  constexpr int LQ = 512 / NT;        // float4 loads per thread and operand tile
  __shared__ __align__(16) float As[2][HK][HP];
  __shared__ __align__(16) float Bs[2][HK][HP];
  const float* A = g.A;
  const float* B = g.B;
  float* C = g.C;
  int M = g.M, N = g.N, K = g.K, kmode = g.kmode;
  float alpha = g.alpha;
  if (g.tree_nb > 0) {
    int lo, mid, hi;
    tree_node(g.tree_nb, g.tree_level, blockIdx.z, lo, mid, hi);
    if (hi - lo < 2) return;
    if (g.tree_pass == 1) {  // TMP[mid:hi, lo:mid] = L[mid:hi, lo:mid] * T[lo:mid, lo:mid]
      A += (size_t)mid * NB * g.lda + (size_t)lo * NB;
      B += (size_t)lo * NB * g.ldb + (size_t)lo * NB;
      C += (size_t)mid * NB * g.ldc + (size_t)lo * NB;
      M = (hi - mid) * NB; N = K = (mid - lo) * NB; kmode = KM_B_LOWER; alpha = 1.f;
    } else {                 // T[mid:hi, lo:mid] = -T[mid:hi, mid:hi] * TMP[mid:hi, lo:mid]
      A += (size_t)mid * NB * g.lda + (size_t)mid * NB;
      B += (size_t)mid * NB * g.ldb + (size_t)lo * NB;
      C += (size_t)mid * NB * g.ldc + (size_t)lo * NB;
      M = K = (hi - mid) * NB; N = (mid - lo) * NB; kmode = KM_A_LOWER; alpha = -1.f;
    }
  }
  // KM_A_LOWER: the K range grows with the row -> schedule the heavy rows first (blocks are dispatched in index order)
  const int nty = (M + HT - 1) / HT;
  if ((int)blockIdx.y >= nty) return;
  const int i0 = ((kmode == KM_A_LOWER) ? nty - 1 - (int)blockIdx.y : (int)blockIdx.y) * HT, j0 = blockIdx.x * HT;
  if (j0 >= N) return;
  if (g.lower_only && j0 > i0) return;
  int k_lo = 0, k_hi = K;
  if (kmode == KM_A_LOWER) k_hi = min(K, i0 + HT);
  else if (kmode == KM_B_LOWER) k_lo = j0;
  else if (kmode == KM_TN_BOTH) k_lo = max(i0, j0);
  else if (kmode == KM_TN_A) k_lo = i0;
  else if (kmode == KM_BT_LOWER) k_hi = min(K, j0 + HT);

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  float acc[RT][8];
#pragma unroll
  for (int r = 0; r < RT; ++r)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[r][c] = 0.f;

  // global -> register staging: 512 float4 per operand tile, two per thread
  // row-major-in-k operands ([rows x K], "transpose on store"): f -> row f/4, k-quad f%4
  // k-major operands ([K x cols], direct store):               f -> k f/32, col-quad f%32
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  auto load_a = [&](int k0, int q) -> float4 {
    const int f = tid + q * NT;
    if (TA) {
      const int kk = f >> 5, c4 = (f & 31) * 4;
      return (i0 + c4 < M) ? *reinterpret_cast<const float4*>(A + (size_t)(k0 + kk) * g.lda + i0 + c4) : zero4;
    } else {
      const int r = f >> 2, c4 = (f & 3) * 4;
      return (i0 + r < M) ? *reinterpret_cast<const float4*>(A + (size_t)(i0 + r) * g.lda + k0 + c4) : zero4;
    }
  };
  auto load_b = [&](int k0, int q) -> float4 {
    const int f = tid + q * NT;
    if (!TB) {
      const int kk = f >> 5, c4 = (f & 31) * 4;
      return (j0 + c4 < N) ? *reinterpret_cast<const float4*>(B + (size_t)(k0 + kk) * g.ldb + j0 + c4) : zero4;
    } else {
      const int r = f >> 2, c4 = (f & 3) * 4;
      return (j0 + r < N) ? *reinterpret_cast<const float4*>(B + (size_t)(j0 + r) * g.ldb + k0 + c4) : zero4;
    }
  };
  auto store_a = [&](int buf, int q, const float4& v) {
    const int f = tid + q * NT;
    if (TA) {
      *reinterpret_cast<float4*>(&As[buf][f >> 5][(f & 31) * 4]) = v;
    } else {
      const int r = f >> 2, c4 = (f & 3) * 4;
      As[buf][c4 + 0][r] = v.x; As[buf][c4 + 1][r] = v.y; As[buf][c4 + 2][r] = v.z; As[buf][c4 + 3][r] = v.w;
    }
  };
  auto store_b = [&](int buf, int q, const float4& v) {
    const int f = tid + q * NT;
    if (!TB) {
      *reinterpret_cast<float4*>(&Bs[buf][f >> 5][(f & 31) * 4]) = v;
    } else {
      const int r = f >> 2, c4 = (f & 3) * 4;
      Bs[buf][c4 + 0][r] = v.x; Bs[buf][c4 + 1][r] = v.y; Bs[buf][c4 + 2][r] = v.z; Bs[buf][c4 + 3][r] = v.w;
    }
  };

  if (k_lo < k_hi) {
    float4 ra[LQ], rb[LQ];
#pragma unroll
    for (int q = 0; q < LQ; ++q) { ra[q] = load_a(k_lo, q); rb[q] = load_b(k_lo, q); }
#pragma unroll
    for (int q = 0; q < LQ; ++q) { store_a(0, q, ra[q]); store_b(0, q, rb[q]); }
    __syncthreads();
    int buf = 0;
    for (int k0 = k_lo; k0 < k_hi; k0 += HK) {
      const bool more = (k0 + HK < k_hi);
      if (more) {
#pragma unroll
        for (int q = 0; q < LQ; ++q) { ra[q] = load_a(k0 + HK, q); rb[q] = load_b(k0 + HK, q); }
      }
#pragma unroll
      for (int kk = 0; kk < HK; ++kk) {
        const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
        const float4 a1 = (RT == 8) ? *reinterpret_cast<const float4*>(&As[buf][kk][64 + (ty & 15) * 4]) : a0;
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
        const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
        const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int r = 0; r < RT; ++r)
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[r][c] = fmaf(av[r], bv[c], acc[r][c]);
      }
      if (more) {
#pragma unroll
        for (int q = 0; q < LQ; ++q) { store_a(buf ^ 1, q, ra[q]); store_b(buf ^ 1, q, rb[q]); }
      }
      __syncthreads();
      buf ^= 1;
    }
  }
#pragma unroll
  for (int r = 0; r < RT; ++r) {
    const int row = i0 + (r < 4 ? ty * 4 + r : 64 + ty * 4 + (r - 4));
    if (row >= M) continue;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int col = j0 + h * 64 + tx * 4;
      if (col >= N) continue;
      float* cp = C + (size_t)row * g.ldc + col;
      float4 o = make_float4(alpha * acc[r][h * 4 + 0], alpha * acc[r][h * 4 + 1], alpha * acc[r][h * 4 + 2], alpha * acc[r][h * 4 + 3]);
      if (g.beta != 0.f) {
        const float4 c = *reinterpret_cast<const float4*>(cp);
        o.x = fmaf(g.beta, c.x, o.x); o.y = fmaf(g.beta, c.y, o.y); o.z = fmaf(g.beta, c.z, o.z); o.w = fmaf(g.beta, c.w, o.w);
      }
      *reinterpret_cast<float4*>(cp) = o;
    }
  }
}

int32_t gemm(cudaStream_t s, bool ta, bool tb, const GemmArgs& g, int grid_z) {
  if (g.M <= 0 || g.N <= 0) return 0;
  // large, squarish products go to the 128x128 kernel; skinny ones (panels of 64, the refinement's 64 starts) stay on 64x64
  static const bool force64 = getenv("BX_GEMM64") != nullptr;
  // Tile-size choice by a two-term model in clocks: a launch takes the longer of its critical path (the tile with the
  // longest K range: K/16 k-steps of ~2300 clocks for a 128x128 tile, ~600 for a 64x64 tile, both with 256 threads) and
  // its total work at the kernel's measured throughput (~45 vs ~18 TFLOP/s).  Few-tile products (n <= 1024, the bottom of
  // the inverse tree, late trailing updates) are latency-bound and finish sooner on 64x64 tiles.
  const double kext = (g.tree_nb > 0) ? (double)g.M : (double)g.K;  // tree mode: extents are per node, ~M
  double flops = 2.0 * (double)g.M * (double)g.N * kext * (double)grid_z;
  if (g.lower_only) flops *= 0.5;
  if (g.kmode != KM_FULL || g.tree_nb > 0) flops *= 0.5;
  const double t128 = fmax(kext / 16.0 * 2300.0, flops * 4.2e-5), t64 = fmax(kext / 16.0 * 600.0, flops * 1.06e-4);
  if (g.M >= HT && g.N >= HT && !force64 && g.ksplit == 0 && t128 < t64) {
    dim3 grid((g.N + HT - 1) / HT, (g.M + HT - 1) / HT, grid_z);
    // at most ~two tiles per SM: one 16-warp CTA per tile; otherwise two 8-warp CTAs per SM
    static const int wide_max = [] { const char* e = getenv("BX_GEMM_WIDE_MAX"); return e ? atoi(e) : 296; }();
    double tiles = (double)grid.x * grid.y * grid.z;
    if (g.lower_only) tiles *= 0.5;
    if (tiles <= (double)wide_max) {
      if (!ta && !tb) gemm128_kernel<false, false, 512><<<grid, 512, 0, s>>>(g);
      else if (ta && !tb) gemm128_kernel<true, false, 512><<<grid, 512, 0, s>>>(g);
      else if (!ta && tb) gemm128_kernel<false, true, 512><<<grid, 512, 0, s>>>(g);
      else gemm128_kernel<true, true, 512><<<grid, 512, 0, s>>>(g);
    } else {
      if (!ta && !tb) gemm128_kernel<false, false, 256><<<grid, 256, 0, s>>>(g);
      else if (ta && !tb) gemm128_kernel<true, false, 256><<<grid, 256, 0, s>>>(g);
      else if (!ta && tb) gemm128_kernel<false, true, 256><<<grid, 256, 0, s>>>(g);
      else gemm128_kernel<true, true, 256><<<grid, 256, 0, s>>>(g);
    }
    BX_LAUNCH_CHECK();
    return 0;
  }
  dim3 grid(g.N / GT, g.M / GT, grid_z);
  if (grid.x == 0 || grid.y == 0) return 0;
  if (!ta && !tb) gemm64_kernel<false, false><<<grid, 256, 0, s>>>(g);
  else if (ta && !tb) gemm64_kernel<true, false><<<grid, 256, 0, s>>>(g);
  else if (!ta && tb) gemm64_kernel<false, true><<<grid, 256, 0, s>>>(g);
  else gemm64_kernel<true, true><<<grid, 256, 0, s>>>(g);
  BX_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------------------------------
// Diagonal-block factorisation: L_kk = chol(A_kk) in shared memory, plus L_kk^-1 and sum(log diag).
// ------------------------------------------------------------------------------------------------------------------
// This kernel is the serial chain of the blocked Cholesky (np / 64 launches per factorisation), so it is built for
// latency.  The 64x64 block is handled as 2x2 blocks of 32; a 32x32 factorisation or triangular solve runs inside ONE
// warp with a matrix row per lane in registers (fully unrolled; per column one __syncwarp and broadcast shared-memory
// reads instead of block barriers); the 32x32x32 products in between use all eight warps.  Seven block barriers.
// clock64 stamps per phase (B200, 1.965 GHz): load 1.2k | chol32 5.5k | solves + rank-32 update 6k | chol32 5.5k |
// solves + products 5.8k | store 1.6k = 25.6k clocks = 13 us (the first shared-memory version: 35 us).
//   A00 = L00 L00^T | T00 = L00^-1, L10 = A10 L00^-T | A11 -= L10 L10^T | A11 = L11 L11^T | T11 = L11^-1,
//   W = L10 T00 | T10 = -T11 W
constexpr int LP = NB + 1;
constexpr unsigned FULLW = 0xffffffffu;

// lane i holds row i of a symmetric positive-definite 32x32 block (upper part zero); on return row i of its Cholesky
// factor; returns l_ii.  Per column the raw column (pivot included) is published to shared memory once and read back as
// broadcast LDS.128 - a first version fetched every l_cj with its own warp shuffle (496 dependent-issue SHFLs per
// factorisation) and measured ~390 clocks per column; the update uses a_ic -= a_ij a_cj / d_j so the loads do not
// wait for the rsqrt chain.  cs = [2][32] floats of shared memory (double buffer: no barrier between columns beyond
// the one __syncwarp).
constexpr int LTP = 36;  // row stride of the transposed factor copies (16-byte aligned rows)
__device__ __forceinline__ float chol32_warp(float (&a)[32], int lane, float* cs, float* LT, bool& bad) {
  float diag = 1.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    float* col = cs + (j & 1) * 32;
    col[lane] = a[j];
    __syncwarp();
    float cv[32];
#pragma unroll
    for (int c4 = (j + 1) & ~3; c4 < 32; c4 += 4) {
      const float4 v = *reinterpret_cast<const float4*>(col + c4);
      cv[c4] = v.x; cv[c4 + 1] = v.y; cv[c4 + 2] = v.z; cv[c4 + 3] = v.w;
    }
    const float d = col[j];
    if (!(d > 0.f) || !(d < CUDART_INF_F)) bad = true;
    float rs = rsqrtf(d);
    rs = rs * fmaf(-0.5f * d, rs * rs, 1.5f);  // one Newton step: full fp32 accuracy without sqrt + divide in the chain
    const float f = a[j] * (rs * rs);          // l_ij / l_jj
    const float lij = (lane == j) ? d * rs : a[j] * rs;
    a[j] = lij;
    LT[j * LTP + lane] = lij;  // column j of L as a contiguous row: the triangular solves read it as broadcast LDS.128
    if (lane == j) diag = lij;
#pragma unroll
    for (int c = j + 1; c < 32; ++c) a[c] = fmaf(-f, cv[c], a[c]);
  }
  return diag;  // l_ii of this lane's row
}
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rsqrt_approx(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// Same contract as chol32_warp, two columns per shared-memory round trip (16 instead of 32 publish -> __syncwarp -> read
// chains; measured 4.0k clocks per 32x32 block against 5.5k): both raw columns j, j+1 are published together; every lane folds the column-j update of column j+1
// (w_c = a_c,j+1 - a_c,j e / d0, e = a_j+1,j) into the factors and applies both rank-1 updates with two FMAs per entry.  The update factors come from
// rcp + one Newton step (a shorter dependent chain than rsqrt -> square); the rsqrt for the entries of L is off the chain.
// cs = [2][2][32] floats.
__device__ __forceinline__ float chol32_warp2(float (&a)[32], int lane, float* cs, float* LT, bool& bad) {
  float diag = 1.f;
#pragma unroll
  for (int j = 0; j < 32; j += 2) {
    float* cA = cs + ((j >> 1) & 1) * 64;
    float* cB = cA + 32;
    cA[lane] = a[j];
    cB[lane] = a[j + 1];
    __syncwarp();
    float va[32], vb[32];
#pragma unroll
    for (int c4 = (j + 2) & ~3; c4 < 32; c4 += 4) {
      const float4 u = *reinterpret_cast<const float4*>(cA + c4);
      const float4 v = *reinterpret_cast<const float4*>(cB + c4);
      va[c4] = u.x; va[c4 + 1] = u.y; va[c4 + 2] = u.z; va[c4 + 3] = u.w;
      vb[c4] = v.x; vb[c4 + 1] = v.y; vb[c4 + 2] = v.z; vb[c4 + 3] = v.w;
    }
    // (fetching this 2x2 pivot block by warp shuffle instead was measured: no faster)
    const float2 de = *reinterpret_cast<const float2*>(cA + j);  // pivot of column j, a_j+1,j
    const float d0 = de.x, e = de.y, d1raw = cB[j + 1];
    float r0 = rcp_approx(d0);
    r0 = r0 * fmaf(-d0, r0, 2.f);
    const float m = e * r0;
    const float d1 = fmaf(-m, e, d1raw);  // pivot of column j+1 (the same operation lane j+1 applies to its own entry)
    float r1 = rcp_approx(d1);
    r1 = r1 * fmaf(-d1, r1, 2.f);
    if (!(d0 > 0.f) || !(d0 < CUDART_INF_F) || !(d1 > 0.f) || !(d1 < CUDART_INF_F)) bad = true;
    const float f0 = a[j] * r0;
    const float a1 = fmaf(-f0, e, a[j + 1]);
    const float f1 = a1 * r1;
    const float g0 = fmaf(-f1, m, f0);  // a_c -= f0 va_c + f1 (vb_c - m va_c)  =  g0 va_c + f1 vb_c : two FMAs per entry
#pragma unroll
    for (int c = j + 2; c < 32; ++c) a[c] = fmaf(-f1, vb[c], fmaf(-g0, va[c], a[c]));
    float rs0 = rsqrt_approx(d0), rs1 = rsqrt_approx(d1);
    rs0 = rs0 * fmaf(-0.5f * d0, rs0 * rs0, 1.5f);  // one Newton step: full fp32 accuracy
    rs1 = rs1 * fmaf(-0.5f * d1, rs1 * rs1, 1.5f);
    const float l0 = a[j] * rs0, l1 = a1 * rs1;  // lanes j / j+1 hold d0 / d1 there: l_jj = d rsqrt(d)
    a[j] = l0;
    a[j + 1] = l1;
    LT[j * LTP + lane] = l0;  // column j of L as a contiguous row: the triangular solves read it as broadcast LDS.128
    LT[(j + 1) * LTP + lane] = l1;
    if (lane == j) diag = l0;
    if (lane == j + 1) diag = l1;
  }
  return diag;  // l_ii of this lane's row
}
// lane r holds row r of B; solves X L^T = B in place.  LT = L^T (row j = column j of L, stride LTP) in shared memory,
// rdiag = 1 / diag(L).  With B = identity this gives row r of L^-T = column r of L^-1.
__device__ __forceinline__ void trsm32_warp(float (&b)[32], const float* LT, const float* rdiag) {
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    b[j] *= rdiag[j];
#pragma unroll
    for (int c4 = (j + 1) & ~3; c4 < 32; c4 += 4) {
      const float4 v = *reinterpret_cast<const float4*>(LT + j * LTP + c4);
      if (c4 + 0 > j) b[c4 + 0] = fmaf(-b[j], v.x, b[c4 + 0]);
      if (c4 + 1 > j) b[c4 + 1] = fmaf(-b[j], v.y, b[c4 + 1]);
      if (c4 + 2 > j) b[c4 + 2] = fmaf(-b[j], v.z, b[c4 + 2]);
      if (c4 + 3 > j) b[c4 + 3] = fmaf(-b[j], v.w, b[c4 + 3]);
    }
  }
}

__global__ void __launch_bounds__(256) potrf_diag_kernel(float* A, int lda, int kb, float* Dinv, float* logdet_part,
                                                         int* status) {
  __shared__ float S[NB][LP];   // A block -> L
  __shared__ float X[NB][LP];   // L^-1
  __shared__ float W[32][33];   // L10 T00
  __shared__ float rdiag[NB];
  __shared__ __align__(16) float colbuf[2][32];
  __shared__ __align__(16) float LT0[32 * LTP], LT1[32 * LTP];  // L00^T, L11^T
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(FULLW, tid >> 5, 0);  // provably warp-uniform: the role branches below stay convergent
  float* blk = A + (size_t)kb * NB * lda + (size_t)kb * NB;
  {
    float4 v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {  // 1024 float4, all four loads of a thread in flight together
      const int e4 = tid + q * 256, r = e4 >> 4, c = (e4 & 15) * 4;
      v[q] = *reinterpret_cast<const float4*>(blk + (size_t)r * lda + c);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int e4 = tid + q * 256, r = e4 >> 4, c = (e4 & 15) * 4;
      S[r][c] = (c <= r) ? v[q].x : 0.f; S[r][c + 1] = (c + 1 <= r) ? v[q].y : 0.f;
      S[r][c + 2] = (c + 2 <= r) ? v[q].z : 0.f; S[r][c + 3] = (c + 3 <= r) ? v[q].w : 0.f;
      X[r][c] = 0.f; X[r][c + 1] = 0.f; X[r][c + 2] = 0.f; X[r][c + 3] = 0.f;
    }
  }
  __syncthreads();
  bool bad = false;
  // 1. A00 = L00 L00^T
  if (warp == 0) {
    float a[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) a[c] = S[lane][c];
    const float dg = chol32_warp(a, lane, &colbuf[0][0], LT0, bad);
#pragma unroll
    for (int c = 0; c < 32; ++c) if (c <= lane) S[lane][c] = a[c];
    rdiag[lane] = 1.f / dg;
  }
  __syncthreads();
  // 2. T00 (warp 1) and L10 = A10 L00^-T (warp 2), independent
  if (warp == 1) {
    float x[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) x[c] = (c == lane) ? 1.f : 0.f;
    trsm32_warp(x, LT0, rdiag);
#pragma unroll
    for (int r = 0; r < 32; ++r) X[r][lane] = x[r];  // x[r] = T00[r][lane], zero for r < lane
  } else if (warp == 2) {
    float b[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) b[c] = S[32 + lane][c];
    trsm32_warp(b, LT0, rdiag);
#pragma unroll
    for (int c = 0; c < 32; ++c) S[32 + lane][c] = b[c];
  }
  __syncthreads();
  // 3. A11 -= L10 L10^T : warp w takes columns 4w .. 4w+3, lane = row
  {
    float l[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) l[k] = S[32 + lane][k];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int c = warp * 4 + q;
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 32; ++k) acc = fmaf(l[k], S[32 + c][k], acc);
      if (c <= lane) S[32 + lane][32 + c] -= acc;
    }
  }
  __syncthreads();
  // 4. A11 = L11 L11^T
  if (warp == 0) {
    float a[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) a[c] = S[32 + lane][32 + c];
    const float dg = chol32_warp(a, lane, &colbuf[0][0], LT1, bad);
#pragma unroll
    for (int c = 0; c < 32; ++c) if (c <= lane) S[32 + lane][32 + c] = a[c];
    rdiag[32 + lane] = 1.f / dg;
    if (bad && lane == 0) atomicOr(status, 1);
  }
  __syncthreads();
  // 5. T11 (warp 1) and W = L10 T00 (warps 4-7: 8 columns each)
  if (warp == 1) {
    float x[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) x[c] = (c == lane) ? 1.f : 0.f;
    trsm32_warp(x, LT1, rdiag + 32);
#pragma unroll
    for (int r = 0; r < 32; ++r) X[32 + r][32 + lane] = x[r];
  } else if (warp >= 4) {
    float l[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) l[k] = S[32 + lane][k];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int c = (warp - 4) * 8 + q;
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 32; ++k) acc = fmaf(l[k], X[k][c], acc);  // T00[k][c] = 0 for k < c
      W[lane][c] = acc;
    }
  }
  __syncthreads();
  // 6. T10 = -T11 W
  {
    float t[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) t[k] = X[32 + lane][32 + k];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int c = warp * 4 + q;
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 32; ++k) acc = fmaf(t[k], W[k][c], acc);
      X[32 + lane][c] = -acc;
    }
  }
  __syncthreads();
  float* dinv = Dinv + (size_t)kb * NB * NB;
  for (int e = tid; e < NB * NB; e += 256) {
    const int r = e >> 6, c = e & 63;
    if (c <= r) blk[(size_t)r * lda + c] = S[r][c];
    dinv[e] = X[r][c];
  }
  if (tid < 32) {
    float l = logf(S[lane][lane]) + logf(S[lane + 32][lane + 32]);
    l = warp_sum(l);
    if (tid == 0) logdet_part[kb] = l;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Fused block column: diagonal factorisation + panel solve in ONE launch (the serial chain of the blocked Cholesky).
// ------------------------------------------------------------------------------------------------------------------
// grid = nb - k CTAs.  EVERY CTA factorises the diagonal block itself - redundant but parallel, it removes the launch
// boundary between potrf and the panel solve - and CTA b >= 1 then solves its own 64x64 block of the panel,
// X L_kk^T = B, by forward substitution (trsm32_warp, a row of B per lane): no inverse is formed on the chain.
// The `t` previous block columns of the same panel are applied first (left-looking inside the panel), which also
// removes the separate rank-64 update launch:  D -= sum_j P_j P_j^T,  B -= sum_j Q_j P_j^T  with P_j = L[k][k-t+j],
// Q_j = L[k+b][k-t+j].  CTA 0 stores L_kk to Ldiag[k] (not in place: the other CTAs still read A_kk), sum(log diag)
// and the pivot status; diag_inverse_kernel below turns Ldiag into the block inverses off the chain, for all blocks
// at once.  Measured phase clocks of one CTA (scripts/dev/bench_potrf_solve.cu, B200): load 1.6k | chol32 3.3-4.0k | L10 / X0
// solves + rank-32 updates 3.9k | chol32 3.5-4.0k | X1 solve 2.0k | store 0.3k = 14.6k (7.4 us); a previous column of the
// panel adds 9.2k (two 64^3 products on one SM: FMA-issue-bound) - against potrf_diag 25.6k + a 5.6 us GEMM launch
// (+ another for the in-panel update) before.
constexpr int LQ = 68;  // shared-memory row stride of the 64x64 tiles: 16-byte aligned rows, LDS.128 along k
constexpr int PS_TILE = NB * LQ;
constexpr size_t PS_SMEM = sizeof(float) * (4 * PS_TILE + 2 * 32 * LTP + 128 + 64);

// C[r][c] -= sum_k A[r][k] B[c][k]   (r < NR, c < NC, k < KK; shared memory, stride LQ; 256 threads, no barrier inside)
template <int NR, int NC, int KK>
__device__ __forceinline__ void mm_nt_sub(float* C, const float* A, const float* B, int tid) {
  constexpr int RI = NR / 16, CJ = NC / 16;
  const int tr = tid >> 4, tc = tid & 15;
  float acc[RI][CJ];
#pragma unroll
  for (int i = 0; i < RI; ++i)
#pragma unroll
    for (int j = 0; j < CJ; ++j) acc[i][j] = 0.f;
#pragma unroll 4
  for (int k = 0; k < KK; k += 4) {
    float4 a[RI], b[CJ];
#pragma unroll
    for (int i = 0; i < RI; ++i) a[i] = *reinterpret_cast<const float4*>(A + (tr + 16 * i) * LQ + k);
#pragma unroll
    for (int j = 0; j < CJ; ++j) b[j] = *reinterpret_cast<const float4*>(B + (tc + 16 * j) * LQ + k);
#pragma unroll
    for (int i = 0; i < RI; ++i)
#pragma unroll
      for (int j = 0; j < CJ; ++j)
        acc[i][j] = fmaf(a[i].x, b[j].x, fmaf(a[i].y, b[j].y, fmaf(a[i].z, b[j].z, fmaf(a[i].w, b[j].w, acc[i][j]))));
  }
#pragma unroll
  for (int i = 0; i < RI; ++i)
#pragma unroll
    for (int j = 0; j < CJ; ++j) C[(tr + 16 * i) * LQ + tc + 16 * j] -= acc[i][j];
}

__device__ __forceinline__ void load_tile64(float* dst, const float* src, int ld, int tid) {
  float4 v[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {  // 1024 float4, all four loads of a thread in flight together
    const int e4 = tid + q * 256, r = e4 >> 4, c = (e4 & 15) * 4;
    v[q] = *reinterpret_cast<const float4*>(src + (size_t)r * ld + c);
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int e4 = tid + q * 256, r = e4 >> 4, c = (e4 & 15) * 4;
    *reinterpret_cast<float4*>(dst + r * LQ + c) = v[q];
  }
}
// lane's row (32 entries starting at column c0) of a stride-LQ tile <-> registers
__device__ __forceinline__ void row_load32(float (&a)[32], const float* row) {
#pragma unroll
  for (int c = 0; c < 32; c += 4) {
    const float4 v = *reinterpret_cast<const float4*>(row + c);
    a[c] = v.x; a[c + 1] = v.y; a[c + 2] = v.z; a[c + 3] = v.w;
  }
}
__device__ __forceinline__ void row_store32(float* row, const float (&a)[32]) {
#pragma unroll
  for (int c = 0; c < 32; c += 4) *reinterpret_cast<float4*>(row + c) = make_float4(a[c], a[c + 1], a[c + 2], a[c + 3]);
}

#ifdef BX_PS_STAMPS  // scripts/dev/bench_potrf_solve.cu: clock64 at the phase boundaries of CTA 1
__device__ long long bx_ps_stamps[8];  // 0 start | 1 loaded | 2 chol | 3 solves + updates | 4 chol | 5 solves | 6 stored
#define PS_STAMP(i) do { if (threadIdx.x == 0 && blockIdx.x == 1) bx_ps_stamps[i] = clock64(); } while (0)
#else
#define PS_STAMP(i)
#endif
__global__ void __launch_bounds__(256) potrf_solve_kernel(float* A, int lda, int k, int t, int next, float* Ldiag,
                                                          float* Xnext, float* logdet_part, int* status) {
  extern __shared__ __align__(16) float ps_smem[];
  float* S = ps_smem;             // D -> L_kk
  float* Bm = S + PS_TILE;        // B -> X (this CTA's block of the panel)
  float* Pm = Bm + PS_TILE;       // staging: P_j
  float* Qm = Pm + PS_TILE;       // staging: Q_j
  float* LT0 = Qm + PS_TILE;      // L00^T, then L11^T (stride LTP)
  float* colbuf = LT0 + 2 * 32 * LTP;  // [2][2][32]
  float* rdiag = colbuf + 128;     // [64]
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(FULLW, tid >> 5, 0);  // provably warp-uniform: the role branches below stay convergent
  const int b = blockIdx.x;                          // 0: diagonal duty only; b >= 1: row block k + b of the panel
  PS_STAMP(0);
  const float* Dg = A + (size_t)k * NB * lda + (size_t)k * NB;
  float* Bg = A + (size_t)(k + b) * NB * lda + (size_t)k * NB;
  load_tile64(S, Dg, lda, tid);
  if (b) load_tile64(Bm, Bg, lda, tid);
  // next = 1 (first column of a two-column panel, t = 0): CTA b also applies this column's update to its block of column
  // k + 1, A[k+b][k+1] -= X_b X_1^T (b = 1: the next diagonal block), so the launch for column k + 1 starts without
  // any pre-update (9.2k clocks there against ~3.5k here).  X_1 = A[k+1][k] L_kk^-T is another CTA's result: every CTA
  // recomputes it from the raw block on two otherwise idle warps - which is why CTA 1 must not overwrite the raw block
  // while a late-starting CTA may still read it: X_1 goes to `Xnext`, diag_inverse_kernel copies it into place.
  const bool fn = next && b;
  if (fn) {
    load_tile64(Pm, A + (size_t)(k + 1) * NB * lda + (size_t)k * NB, lda, tid);
    load_tile64(Qm, A + (size_t)(k + b) * NB * lda + (size_t)(k + 1) * NB, lda, tid);
  }
  for (int j = 0; j < t; ++j) {
    const size_t col = (size_t)(k - t + j) * NB;
    load_tile64(Pm, A + (size_t)k * NB * lda + col, lda, tid);
    if (b) load_tile64(Qm, A + (size_t)(k + b) * NB * lda + col, lda, tid);
    __syncthreads();
    mm_nt_sub<64, 64, 64>(S, Pm, Pm, tid);
    if (b) mm_nt_sub<64, 64, 64>(Bm, Qm, Pm, tid);
    __syncthreads();
  }
  __syncthreads();
  PS_STAMP(1);
  bool bad = false;
  // Both halves of the 2x2 blocking go through the SAME straight-line warp code (a rolled loop over `half` with
  // run-time pointers): the kernel shrinks from 7.0k to 3.6k SASS instructions (112 -> 58 KB of code every CTA has to
  // fetch once), worth 4-6 % of a fit step; the phase clocks themselves did not move, so the warp code is bound by its
  // dependent chain, not by instruction fetch.
#pragma unroll 1
  for (int half = 0; half < 2; ++half) {
    float* Dh = S + half * (32 * LQ + 32);  // D00 / D11
    float* LTh = LT0 + half * (32 * LTP);   // L00^T / L11^T (LT1 follows LT0)
    float* rdh = rdiag + 32 * half;
    if (warp == 0) {  // D_hh = L_hh L_hh^T
      float a[32];
      row_load32(a, Dh + lane * LQ);
#pragma unroll
      for (int c = 0; c < 32; ++c) if (c > lane) a[c] = 0.f;
      const float dg = chol32_warp2(a, lane, colbuf, LTh, bad);
      row_store32(Dh + lane * LQ, a);
      rdh[lane] = 1.f / dg;
    }
    __syncthreads();
    PS_STAMP(2 + 2 * half);
    {  // forward substitution against L_hh: L10 = D10 L00^-T (warp 2, first half), X_h = B_h L_hh^-T (warps 3, 4)
      float* row = nullptr;
      if (warp == 2 && half == 0) row = S + (32 + lane) * LQ;
      else if (b && (warp == 3 || warp == 4)) row = Bm + ((warp - 3) * 32 + lane) * LQ + 32 * half;
      else if (fn && (warp == 5 || warp == 6)) row = Pm + ((warp - 5) * 32 + lane) * LQ + 32 * half;  // X_1
      if (row) {
        float x[32];
        row_load32(x, row);
        trsm32_warp(x, LTh, rdh);
        row_store32(row, x);
      }
    }
    __syncthreads();
    if (half == 0) {  // D11 -= L10 L10^T,  B[:, 32:64] -= X0 L10^T
      mm_nt_sub<32, 32, 32>(S + 32 * LQ + 32, S + 32 * LQ, S + 32 * LQ, tid);
      if (b) mm_nt_sub<64, 32, 32>(Bm + 32, Bm, S + 32 * LQ, tid);
      if (fn) mm_nt_sub<64, 32, 32>(Pm + 32, Pm, S + 32 * LQ, tid);
      __syncthreads();
    }
    PS_STAMP(3 + 2 * half);
  }
  if (warp == 0 && bad && lane == 0 && b == 0) atomicOr(status, 1);
  if (fn) {
    mm_nt_sub<64, 64, 64>(Qm, Bm, Pm, tid);
    __syncthreads();
    float* Qg = A + (size_t)(k + b) * NB * lda + (size_t)(k + 1) * NB;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int e4 = tid + q * 256, r = e4 >> 4, c = (e4 & 15) * 4;
      *reinterpret_cast<float4*>(Qg + (size_t)r * lda + c) = *reinterpret_cast<const float4*>(Qm + r * LQ + c);
    }
  }
  if (b) {
    float* dst = (next && b == 1) ? Xnext : Bg;
    const size_t ldd = (next && b == 1) ? (size_t)NB : (size_t)lda;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int e4 = tid + q * 256, r = e4 >> 4, c = (e4 & 15) * 4;
      *reinterpret_cast<float4*>(dst + r * ldd + c) = *reinterpret_cast<const float4*>(Bm + r * LQ + c);
    }
  } else {
    float* ld = Ldiag + (size_t)k * NB * NB;
    for (int e = tid; e < NB * NB; e += 256) {
      const int r = e >> 6, c = e & 63;
      ld[e] = (c <= r) ? S[r * LQ + c] : 0.f;
    }
    if (tid < 32) {
      float l = logf(S[lane * LQ + lane]) + logf(S[(lane + 32) * LQ + lane + 32]);
      l = warp_sum(l);
      if (tid == 0) logdet_part[k] = l;
    }
  }
  PS_STAMP(6);
}

// Block inverses off the chain: Dinv[k] holds L_kk on entry (from potrf_solve_kernel) and L_kk^-1 on return; L_kk is also
// written in place into the diagonal block of A (and the parked sub-diagonal block of every two-column panel next to it).  One CTA per block, all blocks in one launch.
//   T00 = L00^-1, T11 = L11^-1 (one warp each), W = L10 T00, T10 = -T11 W.
__global__ void __launch_bounds__(256) diag_inverse_kernel(float* A, int lda, float* Dinv, const float* Xnext) {
  __shared__ float S[NB][LP];  // L
  __shared__ float X[NB][LP];  // L^-1
  __shared__ float W[32][33];
  __shared__ float rdiag[NB];
  __shared__ __align__(16) float LT0[32 * LTP], LT1[32 * LTP];
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(FULLW, tid >> 5, 0);
  const int kb = blockIdx.x;
  float* dinv = Dinv + (size_t)kb * NB * NB;
  float* blk = A + (size_t)kb * NB * lda + (size_t)kb * NB;
  if (Xnext && (kb & 1)) {  // the sub-diagonal block of this panel, parked by potrf_solve_kernel(next = 1)
    const float* xn = Xnext + (size_t)(kb >> 1) * NB * NB;
    for (int e = tid; e < NB * NB; e += 256) blk[(size_t)(e >> 6) * lda + (e & 63) - NB] = xn[e];
  }
  for (int e = tid; e < NB * NB; e += 256) {
    const int r = e >> 6, c = e & 63;
    const float v = dinv[e];
    S[r][c] = v;
    X[r][c] = 0.f;
    if (c <= r) blk[(size_t)r * lda + c] = v;
    if (r < 32 && c < 32) LT0[c * LTP + r] = v;
    if (r >= 32 && c >= 32) LT1[(c - 32) * LTP + (r - 32)] = v;
    if (r == c) rdiag[r] = 1.f / v;
  }
  __syncthreads();
  if (warp == 1 || warp == 2) {
    const int h = warp - 1;  // 0: T00, 1: T11
    float x[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) x[c] = (c == lane) ? 1.f : 0.f;
    trsm32_warp(x, h ? LT1 : LT0, rdiag + 32 * h);
#pragma unroll
    for (int r = 0; r < 32; ++r) X[32 * h + r][32 * h + lane] = x[r];  // x[r] = T[r][lane], zero for r < lane
  }
  __syncthreads();
  {  // W = L10 T00 : warp w takes columns 4w .. 4w+3, lane = row
    float l[32];
#pragma unroll
    for (int kk = 0; kk < 32; ++kk) l[kk] = S[32 + lane][kk];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int c = warp * 4 + q;
      float acc = 0.f;
#pragma unroll
      for (int kk = 0; kk < 32; ++kk) acc = fmaf(l[kk], X[kk][c], acc);  // T00[kk][c] = 0 for kk < c
      W[lane][c] = acc;
    }
  }
  __syncthreads();
  {  // T10 = -T11 W
    float tt[32];
#pragma unroll
    for (int kk = 0; kk < 32; ++kk) tt[kk] = X[32 + lane][32 + kk];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int c = warp * 4 + q;
      float acc = 0.f;
#pragma unroll
      for (int kk = 0; kk < 32; ++kk) acc = fmaf(tt[kk], W[kk][c], acc);
      X[32 + lane][c] = -acc;
    }
  }
  __syncthreads();
  for (int e = tid; e < NB * NB; e += 256) dinv[e] = X[e >> 6][e & 63];
}

int32_t Lookahead::init() {
  int least = 0, greatest = 0;
  BX_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
  BX_CUDA(cudaStreamCreateWithPriority(&chain, cudaStreamNonBlocking, greatest));
  BX_CUDA(cudaEventCreateWithFlags(&ev_chain, cudaEventDisableTiming));
  BX_CUDA(cudaEventCreateWithFlags(&ev_bulk, cudaEventDisableTiming));
  return 0;
}
Lookahead::~Lookahead() {
  if (ev_chain) cudaEventDestroy(ev_chain);
  if (ev_bulk) cudaEventDestroy(ev_bulk);
  if (chain) cudaStreamDestroy(chain);
}
// Below ~2048 rows the trailing updates are a few tiles: nothing to overlap the chain with.
bool lookahead_wanted(int np) {
  static const int min_np = [] { const char* e = getenv("BX_CHOL_LOOKAHEAD_MIN"); return e ? atoi(e) : 2048; }();
  return min_np > 0 && np >= min_np;
}

// Per-device function attributes of the factorisation kernels; the entry points call it once per call, before any
// stream capture (a process may drive several devices, so this is not a one-time static).
int32_t cholesky_prepare() {
  BX_CUDA(cudaFuncSetAttribute(potrf_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PS_SMEM));
  return 0;
}

int32_t cholesky_inplace(cudaStream_t s, float* A, int np, int lda, float* Dinv, float* logdet_part, int* status,
                         const Lookahead* la, float* scratch) {
  // Two-level right-looking blocking: 64-wide steps (potrf of the diagonal block, panel solve) are applied only inside a
  // panel of PW blocks; the trailing matrix is updated once per panel with K = 64 PW (a rank-64 update per step moves
  // the whole trailing matrix through the SMs for 64 FMAs per element: 27 TFLOP/s measured at n = 8192).
  // With look-ahead the trailing update of panel p is split: the columns of panel p+1 are updated on the chain stream
  // `c` right away, the rest on `s`; the bulk update of panel p-1 and the next-panel update of panel p both modify the
  // columns of panel p+1, so the chain waits for the former (ev_bulk) before the latter.
  static const int PW = [] { const char* e = getenv("BX_CHOL_PW"); const int v = e ? atoi(e) : 2; return v < 1 ? 1 : v; }();
  static const bool fused = [] { const char* e = getenv("BX_CHOL_FUSED"); return !e || atoi(e) != 0; }();
  static const bool fuse_next_env = [] { const char* e = getenv("BX_CHOL_FUSE_NEXT"); return !e || atoi(e) != 0; }();
  const bool fuse_next = fused && PW == 2 && scratch != nullptr && fuse_next_env;
  const int nb = np / NB;
  cudaStream_t c = la ? la->chain : s;
  bool bulk_pending = false;
  if (la) {
    BX_CUDA(cudaEventRecord(la->ev_bulk, s));
    BX_CUDA(cudaStreamWaitEvent(c, la->ev_bulk, 0));
  }
  for (int k0 = 0; k0 < nb; k0 += PW) {
    const int pw = min(PW, nb - k0);
    for (int t = 0; t < pw; ++t) {
      const int k = k0 + t;
      if (fused) {  // one launch per block column: potrf, panel solve and the update between the panel's two columns
        const int next = (fuse_next && t == 0 && pw == 2) ? 1 : 0;
        potrf_solve_kernel<<<nb - k, 256, PS_SMEM, c>>>(A, lda, k, fuse_next ? 0 : t, next, Dinv,
                                                        fuse_next ? scratch + (size_t)(k >> 1) * NB * NB : nullptr, logdet_part, status);
        BX_LAUNCH_CHECK();
        continue;
      }
      potrf_diag_kernel<<<1, 256, 0, c>>>(A, lda, k, Dinv, logdet_part, status);
      BX_LAUNCH_CHECK();
      const int rem = nb - k - 1;
      if (rem == 0) break;
      float* panel = A + (size_t)(k + 1) * NB * lda + (size_t)k * NB;
      GemmArgs p{};  // panel <- panel * Dinv_k^T   (in place: each CTA owns its 64 rows x 64 cols)
      p.A = panel; p.lda = lda; p.B = Dinv + (size_t)k * NB * NB; p.ldb = NB; p.C = panel; p.ldc = lda;
      p.M = rem * NB; p.N = NB; p.K = NB; p.alpha = 1.f; p.beta = 0.f; p.kmode = KM_FULL;
      int32_t rc = gemm(c, false, true, p);
      if (rc) return rc;
      const int inner = pw - 1 - t;  // block columns of this panel still to be factorised
      if (inner > 0) {
        GemmArgs u{};  // A[k+1:, k+1 : k0+pw] -= L[k+1:, k] * L[k+1 : k0+pw, k]^T
        u.A = panel; u.lda = lda; u.B = panel; u.ldb = lda; u.C = A + (size_t)(k + 1) * NB * lda + (size_t)(k + 1) * NB; u.ldc = lda;
        u.M = rem * NB; u.N = inner * NB; u.K = NB; u.alpha = -1.f; u.beta = 1.f; u.kmode = KM_FULL; u.lower_only = 1;
        rc = gemm(c, false, true, u);
        if (rc) return rc;
      }
    }
    const int rem = nb - k0 - pw;
    if (rem <= 0) break;
    float* wide = A + (size_t)(k0 + pw) * NB * lda + (size_t)k0 * NB;  // L[k0+pw:, k0 : k0+pw]
    float* trail = A + (size_t)(k0 + pw) * NB * lda + (size_t)(k0 + pw) * NB;
    GemmArgs u{};  // trailing(lower tiles) -= wide * wide^T
    u.A = wide; u.lda = lda; u.B = wide; u.ldb = lda; u.C = trail; u.ldc = lda;
    u.M = rem * NB; u.N = rem * NB; u.K = pw * NB; u.alpha = -1.f; u.beta = 1.f; u.kmode = KM_FULL; u.lower_only = 1;
    if (!la) {
      int32_t rc = gemm(s, false, true, u);
      if (rc) return rc;
      continue;
    }
    const int nxt = min(PW, rem);  // block columns of the next panel
    if (rem > nxt) BX_CUDA(cudaEventRecord(la->ev_chain, c));  // panel p is final
    if (bulk_pending) {
      BX_CUDA(cudaStreamWaitEvent(c, la->ev_bulk, 0));
      bulk_pending = false;
    }
    u.N = nxt * NB;
    int32_t rc = gemm(c, false, true, u);
    if (rc) return rc;
    if (rem > nxt) {
      BX_CUDA(cudaStreamWaitEvent(s, la->ev_chain, 0));
      const size_t skip = (size_t)nxt * NB;
      u.A = wide + skip * lda; u.B = u.A; u.C = trail + skip * lda + skip;
      u.M = u.N = (rem - nxt) * NB;
      rc = gemm(s, false, true, u);
      if (rc) return rc;
      BX_CUDA(cudaEventRecord(la->ev_bulk, s));
      bulk_pending = true;
    }
  }
  if (la) {
    BX_CUDA(cudaEventRecord(la->ev_chain, c));
    BX_CUDA(cudaStreamWaitEvent(s, la->ev_chain, 0));
  }
  if (fused) {
    diag_inverse_kernel<<<nb, 256, 0, s>>>(A, lda, Dinv, fuse_next ? scratch : nullptr);
    BX_LAUNCH_CHECK();
  }
  return 0;
}

// Diagonal blocks <- Dinv, everything above the block diagonal <- 0.
__global__ void tri_prep_kernel(float* A, int np, int lda, const float* Dinv) {
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (bj < bi) return;
  float* blk = A + (size_t)bi * NB * lda + (size_t)bj * NB;
  for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) {
    const int r = e >> 6, c = e & 63;
    blk[(size_t)r * lda + c] = (bi == bj) ? Dinv[(size_t)bi * NB * NB + e] : 0.f;
  }
}

int32_t tri_inverse_inplace(cudaStream_t s, float* A, int np, int lda, const float* Dinv, float* TMP) {
  const int nb = np / NB;
  tri_prep_kernel<<<dim3(nb, nb), 256, 0, s>>>(A, np, lda, Dinv);
  BX_LAUNCH_CHECK();
  int levels = 0;
  while ((1 << levels) < nb) ++levels;
  for (int level = levels - 1; level >= 0; --level) {
    const int nodes = 1 << level;
    const int sz = (nb + nodes - 1) / nodes;  // largest node at this level (blocks)
    const int half = (sz + 1) / 2;
    for (int pass = 1; pass <= 2; ++pass) {
      GemmArgs g{};
      g.tree_nb = nb; g.tree_level = level; g.tree_pass = pass;
      g.lda = lda; g.ldb = (pass == 1) ? lda : np; g.ldc = (pass == 1) ? np : lda;
      g.A = A; g.B = (pass == 1) ? A : TMP; g.C = (pass == 1) ? TMP : A;
      g.M = half * NB; g.N = half * NB; g.K = 0; g.beta = 0.f;  // real extents are set per node in the kernel
      int32_t rc = gemm(s, false, false, g, nodes);
      if (rc) return rc;
    }
  }
  return 0;
}

// out[i] = sum_{j<=i} T[i][j] x[j] : one warp per row, float4 loads (512 bytes of the row per warp instruction).
__global__ void trmv_lower_kernel(const float* __restrict__ T, int np, int ld, const float* __restrict__ x, float* out) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= np) return;
  const float4* Tr = reinterpret_cast<const float4*>(T + (size_t)row * ld);
  const float4* x4 = reinterpret_cast<const float4*>(x);
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 4
  for (int j = lane * 4; j <= row; j += 128) {
    const float4 t = Tr[j >> 2], xv = x4[j >> 2];
    a0 = fmaf(t.x, xv.x, a0);                          // j <= row holds for the first element
    a1 = fmaf(j + 1 <= row ? t.y : 0.f, xv.y, a1);
    a2 = fmaf(j + 2 <= row ? t.z : 0.f, xv.z, a2);
    a3 = fmaf(j + 3 <= row ? t.w : 0.f, xv.w, a3);
  }
  const float acc = warp_sum((a0 + a1) + (a2 + a3));
  if (lane == 0) out[row] = acc;
}
// out[j] = sum_{i>=j} T[i][j] x[i] : 64 columns per block, 64 rows per pass (a half-warp reads one 256-byte row segment as
// float4, 32 warps), four passes unrolled so that a block keeps 64 KB in flight - the product is a stream over T from L2 /
// HBM and a block column is walked by ONE block (fixed summation order, no partial-sum buffer); round 1's 32 x 8 version
// kept 1 KB in flight per block and took ~11 us at n = 512, ~0.2 ms at n = 4096.
constexpr int TRT_COLS = 64, TRT_ROWS = 64;
__global__ void __launch_bounds__(1024) trmv_lower_t_kernel(const float* __restrict__ T, int np, int ld,
                                                            const float* __restrict__ x, float* out) {
  __shared__ float red[TRT_ROWS][TRT_COLS + 1];
  const int lane = threadIdx.x & 31, slot = (threadIdx.x >> 5) * 2 + (lane >> 4), c4 = (lane & 15) * 4;
  const int j0 = blockIdx.x * TRT_COLS, j = j0 + c4;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 4
  for (int i = j0 + slot; i < np; i += TRT_ROWS) {
    const float xi = x[i];
    const float4 t = *reinterpret_cast<const float4*>(T + (size_t)i * ld + j);
    a0 = fmaf(i >= j ? t.x : 0.f, xi, a0);  // only the diagonal block has entries above the diagonal
    a1 = fmaf(i >= j + 1 ? t.y : 0.f, xi, a1);
    a2 = fmaf(i >= j + 2 ? t.z : 0.f, xi, a2);
    a3 = fmaf(i >= j + 3 ? t.w : 0.f, xi, a3);
  }
  red[slot][c4] = a0; red[slot][c4 + 1] = a1; red[slot][c4 + 2] = a2; red[slot][c4 + 3] = a3;
  __syncthreads();
  if (threadIdx.x < TRT_COLS) {
    float t = 0.f;
#pragma unroll 8
    for (int r = 0; r < TRT_ROWS; ++r) t += red[r][threadIdx.x];
    out[j0 + threadIdx.x] = t;
  }
}
int32_t trmv_lower(cudaStream_t s, const float* T, int np, int ld, const float* x, float* out) {
  trmv_lower_kernel<<<(np + 7) / 8, 256, 0, s>>>(T, np, ld, x, out);
  BX_LAUNCH_CHECK();
  return 0;
}
int32_t trmv_lower_t(cudaStream_t s, const float* T, int np, int ld, const float* x, float* out) {
  trmv_lower_t_kernel<<<np / TRT_COLS, 1024, 0, s>>>(T, np, ld, x, out);
  BX_LAUNCH_CHECK();
  return 0;
}

}  // namespace bx

extern "C" const char* bx_last_error(void) { return bx::last_error(); }
extern "C" int32_t bx_version(void) { return BX_VERSION; }
extern "C" uint64_t bx_pack(float value, uint32_t index) { return bx::pack_best(value, index); }
extern "C" void bx_unpack(uint64_t key, float* value, uint32_t* index) {
  if (value) *value = bx::from_orderable_u32((uint32_t)(key >> 32));
  if (index) *index = 0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFu);
}
