// TMEM <-> register bandwidth of one SM: NW warps (4 or 8) read / read-modify-write all 512 columns of their lane
// quarter with tcgen05.ld / tcgen05.st .32x32b.x32, REP times; prints clocks per pass and bytes per clock.
// Dev tool for the drained-accumulator design (how often can an epilogue afford to pull a 128 x 256 fp32 accumulator?).
// Build + run: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu && ./tmem_bw
#include <cstdio>
#include "../../bayex_b200/csrc/tc_common.cuh"

using namespace bx::tc;

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// mode 0: read only (sum into a register); 1: read two regions, add, store one (the running-sum update); 2: store only
template <int MODE>
__global__ void __launch_bounds__(256, 1) bw_kernel(int rep, long long* clocks, float* sink) {
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(smem_u32(&tmem_base), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  const int half = warp >> 2, nhalf = blockDim.x >> 7;  // with 8 warps two warps share a lane quarter and split the columns
  float acc = 0.f;
  uint32_t r[32], q[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = i;
  for (int c0 = 0; c0 < 512; c0 += 32) tmem_st32(lane_addr + c0, r);
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < rep; ++it) {
    if (MODE == 0) {
      for (int c0 = half * 32; c0 < 512; c0 += 32 * nhalf) {
        tmem_ld32_nowait(lane_addr + c0, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 32; ++i) acc += __uint_as_float(r[i]);
      }
    } else if (MODE == 1) {
      for (int c0 = half * 32; c0 < 256; c0 += 32 * nhalf) {
        tmem_ld32_nowait(lane_addr + c0, r);
        tmem_ld32_nowait(lane_addr + 256 + c0, q);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) + __uint_as_float(q[i]));
        tmem_st32(lane_addr + c0, r);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    } else {
      for (int c0 = half * 32; c0 < 512; c0 += 32 * nhalf) tmem_st32(lane_addr + c0, r);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) clocks[0] = t1 - t0;
  sink[threadIdx.x] = acc + __uint_as_float(r[3]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int main() {
  long long* d_clk;
  float* d_sink;
  cudaMalloc(&d_clk, 8);
  cudaMalloc(&d_sink, 4 * 256);
  const int rep = 200;
  for (int nw : {4, 8}) {
    for (int mode = 0; mode < 3; ++mode) {
      if (mode == 0) bw_kernel<0><<<1, nw * 32>>>(rep, d_clk, d_sink);
      if (mode == 1) bw_kernel<1><<<1, nw * 32>>>(rep, d_clk, d_sink);
      if (mode == 2) bw_kernel<2><<<1, nw * 32>>>(rep, d_clk, d_sink);
      cudaError_t e = cudaDeviceSynchronize();
      long long clk = 0;
      cudaMemcpy(&clk, d_clk, 8, cudaMemcpyDeviceToHost);
      const double per = (double)clk / rep;
      const double bytes = mode == 1 ? 128.0 * 256 * 4 * 3 : 128.0 * 512 * 4;
      printf("warps %d mode %d (%s): %.0f clocks per pass, %.1f TMEM bytes per clock  [%s]\n", nw, mode,
             mode == 0 ? "ld 512 cols + 16k FADD" : mode == 1 ? "ld 256 + ld 256 + add + st 256" : "st 512 cols", per, bytes / per,
             cudaGetErrorString(e));
    }
  }
  return 0;
}
