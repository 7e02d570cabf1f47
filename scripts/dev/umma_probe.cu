// Probe of the tcgen05.mma (kind::f16, fp32 accumulate) accumulation arithmetic: D[128 x 128] = A[128 x K] B[128 x K]^T on
// fp16 operands given by the host, K-steps of 16 issued in the given order into ONE TMEM accumulator; every `chunk`
// K-elements the accumulator is drained into fp32 registers (round-to-nearest adds) and restarted (chunk = 0: never).
// The host (scripts/dev/umma_probe.py) crafts the operands: head/tail splits interleaved the way the scoring engine issues
// them, real T / K* data, sign patterns.  Build: nvcc -shared (see umma_probe.py).  Dev tool, not part of the library.
#include "../../bayex_b200/csrc/tc_common.cuh"

using namespace bx::tc;

__device__ __forceinline__ void umma_f16_1(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

struct ProbeSmem {
  char a[128 * 128], b[128 * 128];  // 128 rows x 64 halves, SWIZZLE_128B
  unsigned long long done;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(128, 1) probe_kernel(const uint16_t* __restrict__ A, const uint16_t* __restrict__ B,
                                                      float* __restrict__ D, int K, int chunk) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  ProbeSmem& S = *reinterpret_cast<ProbeSmem*>(smem_raw);
  if ((smem_u32(smem_raw) & 1023u) != 0u) __trap();
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(smem_u32(&S.done), 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&S.tmem_base), 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = S.tmem_base;
  const uint32_t idesc = make_idesc_f16(128);
  const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
  float run[128];
#pragma unroll
  for (int i = 0; i < 128; ++i) run[i] = 0.f;
  uint32_t phase = 0;
  int in_chunk = 0;  // K elements accumulated in TMEM since the last drain
  for (int kb = 0; kb < K / 64; ++kb) {
    const uint4* ar = reinterpret_cast<const uint4*>(A + (size_t)tid * K + kb * 64);
    const uint4* br = reinterpret_cast<const uint4*>(B + (size_t)tid * K + kb * 64);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const uint32_t off = (uint32_t)(tid * 128 + ((q ^ (tid & 7)) << 4));
      *reinterpret_cast<uint4*>(S.a + off) = ar[q];
      *reinterpret_cast<uint4*>(S.b + off) = br[q];
    }
    fence_proxy_async();
    __syncthreads();
    for (int ks = 0; ks < 4; ++ks) {
      if (tid == 0) {
        tc_fence_after();
        const uint64_t ad = make_desc(smem_u32(S.a)) + (uint64_t)(ks * 2), bd = make_desc(smem_u32(S.b)) + (uint64_t)(ks * 2);
        umma_f16_1(tmem, ad, bd, idesc, in_chunk != 0);
        umma_commit(smem_u32(&S.done));
      }
      in_chunk += 16;
      const bool last = (kb == K / 64 - 1 && ks == 3);
      const bool drain = last || (chunk > 0 && in_chunk >= chunk);
      mbar_wait(smem_u32(&S.done), phase & 1);  // probe, not a fast kernel: every MMA is waited for
      tc_fence_after();
      ++phase;
      if (drain) {
#pragma unroll
        for (int c0 = 0; c0 < 128; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(lane_addr + (uint32_t)c0, r);
#pragma unroll
          for (int i = 0; i < 32; ++i) run[c0 + i] += __uint_as_float(r[i]);
        }
        tc_fence_before();
        in_chunk = 0;
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 128; ++i) D[(size_t)tid * 128 + i] = run[i];
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 128);
  }
}

extern "C" int umma_probe(const void* dA, const void* dB, float* dD, int K, int chunk, void* stream) {
  if (K <= 0 || K % 64 || (chunk % 16)) return -1;
  const size_t smem = sizeof(ProbeSmem) + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>((const uint16_t*)dA, (const uint16_t*)dB, dD, K, chunk);
  return (int)cudaGetLastError();
}
