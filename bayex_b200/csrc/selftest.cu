// tcgen05 self-test: D[128 x 128] = A[128 x K] * B[128 x K]^T (3xTF32) through the same UMMA descriptors, SWIZZLE_128B
// operand layout, TMA tensor map and tcgen05.ld path as the scoring engines - the unit test of the tensor-core plumbing
// (tests/test_gpu_parity.py::test_umma_selftest_3xtf32).
#include "tc_common.cuh"

namespace bx {
namespace tc {
constexpr int SELF_N = 128;  // accumulator width of the self-test
}  // namespace tc

// ---- self-test: D[128 x 128] = A[128 x K] * B[128 x K]^T with the same descriptors / swizzle / TMA / tcgen05.ld -----
namespace tc {

struct SelfSmem {
  char a_hi[TILE_BYTES], a_lo[TILE_BYTES], b_hi[TILE_BYTES], b_lo[TILE_BYTES];
  unsigned long long full_b, mma_done;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(128, 1)
selftest_kernel(const float* __restrict__ A, float* __restrict__ D, int K, const __grid_constant__ CUtensorMap map_hi,
                const __grid_constant__ CUtensorMap map_lo) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  SelfSmem& S = *reinterpret_cast<SelfSmem*>(smem_raw);
  if ((smem_u32(smem_raw) & 1023u) != 0u) __trap();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(smem_u32(&S.full_b), 1);
    mbar_init(smem_u32(&S.mma_done), 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&S.tmem_base), 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = S.tmem_base;
  const uint32_t idesc = make_idesc(SELF_N);
  for (int kb = 0; kb < K / BK; ++kb) {
    if (tid == 0) {
      const uint32_t bar = smem_u32(&S.full_b);
      mbar_expect_tx(bar, 2 * TILE_BYTES);
      tma_load_2d(smem_u32(S.b_hi), &map_hi, bar, kb * BK, 0);
      tma_load_2d(smem_u32(S.b_lo), &map_lo, bar, kb * BK, 0);
    }
    for (int qd = 0; qd < 8; ++qd) {  // row `tid` of A, 8 chunks of 4 floats
      const float4 v = *reinterpret_cast<const float4*>(A + (size_t)tid * K + kb * BK + qd * 4);
      split_store(smem_u32(S.a_hi), smem_u32(S.a_lo), tid, qd, v.x, v.y, v.z, v.w);
    }
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      mbar_wait(smem_u32(&S.full_b), kb & 1);
      tc_fence_after();
      const uint64_t ah = make_desc(smem_u32(S.a_hi)), al = make_desc(smem_u32(S.a_lo));
      const uint64_t bh = make_desc(smem_u32(S.b_hi)), bl = make_desc(smem_u32(S.b_lo));
      for (int ks = 0; ks < BK / 8; ++ks) {
        const uint64_t o = (uint64_t)(ks * 2);
        umma_tf32(tmem, ah + o, bh + o, idesc, (kb | ks) != 0);
        umma_tf32(tmem, al + o, bh + o, idesc, 1u);
        umma_tf32(tmem, ah + o, bl + o, idesc, 1u);
      }
      umma_commit(smem_u32(&S.mma_done));
    }
    mbar_wait(smem_u32(&S.mma_done), kb & 1);  // operands consumed: the tiles may be overwritten
    tc_fence_after();
  }
  const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
  for (int c0 = 0; c0 < SELF_N; c0 += 32) {
    uint32_t r[32];
    tmem_ld32(lane_addr + (uint32_t)c0, r);
    for (int i = 0; i < 32; ++i) D[(size_t)(warp * 32 + lane) * SELF_N + c0 + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 128);
  }
}

__global__ void split_kernel(const float* __restrict__ x, size_t n, float* hi, float* lo) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const float h = __uint_as_float(__float_as_uint(x[e]) & 0xFFFFE000u);
  hi[e] = h;
  lo[e] = x[e] - h;
}

}  // namespace tc
}  // namespace bx

extern "C" int32_t bx_selftest_umma(const float* d_A, const float* d_B, float* d_D, int32_t N, int32_t K, void* stream) {
  using namespace bx;
  BX_CHECK_ARG(d_A && d_B && d_D, BX_E_ARG, "null pointer argument");
  BX_CHECK_ARG(N == tc::SELF_N && K > 0 && K % tc::BK == 0, BX_E_RANGE, "self-test needs N == 128 and K a multiple of 32");
  cudaStream_t s = (cudaStream_t)stream;
  float *hi = nullptr, *lo = nullptr;
  const size_t n = (size_t)N * K;
  BX_CUDA(cudaMallocAsync(&hi, n * sizeof(float), s));
  BX_CUDA(cudaMallocAsync(&lo, n * sizeof(float), s));
  tc::split_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d_B, n, hi, lo);
  BX_LAUNCH_CHECK();
  CUtensorMap mh, ml;
  int32_t rc = tc::make_map(&mh, hi, N, K, K, tc::SELF_N);
  if (rc == 0) rc = tc::make_map(&ml, lo, N, K, K, tc::SELF_N);
  if (rc == 0) {
    const size_t smem = sizeof(tc::SelfSmem);
    cudaFuncSetAttribute(tc::selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    tc::selftest_kernel<<<1, 128, smem, s>>>(d_A, d_D, K, mh, ml);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("selftest launch failed: %s", cudaGetErrorString(e)); rc = (int32_t)e; }
  }
  cudaFreeAsync(hi, s);
  cudaFreeAsync(lo, s);
  return rc;
}

// ---- fp32 SIMT peak: the denominator of the fit path's roofline (MEASURED_PEAKS.json has the HBM and 16-bit tensor
// figures only).  Every thread runs 8 independent FFMA chains; flops = grid * 256 threads * iters * 8 * 2. -------------
namespace bx {
__global__ void __launch_bounds__(256) ffma_peak_kernel(float* out, int iters, float a, float b) {
  float x0 = threadIdx.x, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
#pragma unroll 4
  for (int i = 0; i < iters; ++i) {
    x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
    x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
  }
  const float s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 123.456f) out[0] = s;  // keeps the chains alive without a store in the common case
}
}  // namespace bx

extern "C" int32_t bx_selftest_ffma(float* d_out, int32_t blocks, int32_t iters, void* stream) {
  BX_CHECK_ARG(d_out && blocks > 0 && iters > 0, BX_E_ARG, "bad argument");
  bx::ffma_peak_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_out, iters, 0.999f, 0.001f);
  BX_LAUNCH_CHECK();
  return 0;
}
