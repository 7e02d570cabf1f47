// Phase clocks and launch time of potrf_solve_kernel (the chain element of the blocked Cholesky).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I include -o bayex_b200/csrc/build/bench_potrf_solve scripts/dev/bench_potrf_solve.cu
#define BX_PS_STAMPS
#include "../../bayex_b200/csrc/linalg.cu"
#include <cstdio>
#include <vector>
using namespace bx;
int main() {
  const int np = 512, nb = np / NB;
  std::vector<float> h((size_t)np * np);
  for (int i = 0; i < np; ++i)
    for (int j = 0; j < np; ++j) h[(size_t)i * np + j] = (i == j ? 2.0f : 0.f) + 0.5f * expf(-0.001f * (i - j) * (i - j));
  float *A, *A0, *Ld, *logdet; int* status;
  cudaMalloc(&A, h.size() * 4); cudaMalloc(&A0, h.size() * 4); cudaMalloc(&Ld, (size_t)nb * NB * NB * 4);
  cudaMalloc(&logdet, nb * 4); cudaMalloc(&status, 8);
  float* Xn; cudaMalloc(&Xn, NB * NB * 4);
  cudaMemcpy(A0, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(status, 0, 8);
  cudaFuncSetAttribute(potrf_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PS_SMEM);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  // mode 0: plain column | 1: column with the in-kernel pre-update from the previous one | 2: column that also updates the next
  for (int mode = 0; mode < 3; ++mode) {
    const int t = mode == 1 ? 1 : 0;
    float best = 1e9f;
    long long st[8];
    for (int rep = 0; rep < 20; ++rep) {
      cudaMemcpy(A, A0, h.size() * 4, cudaMemcpyDeviceToDevice);
      if (t == 1) potrf_solve_kernel<<<nb, 256, PS_SMEM>>>(A, np, 0, 0, 0, Ld, nullptr, logdet, status);
      cudaDeviceSynchronize();
      cudaEventRecord(e0);
      potrf_solve_kernel<<<nb - t, 256, PS_SMEM>>>(A, np, t, mode == 1 ? 1 : 0, mode == 2 ? 1 : 0, Ld, Xn, logdet, status);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); best = fminf(best, ms);
    }
    cudaMemcpyFromSymbol(st, bx_ps_stamps, sizeof(st));
    printf("mode=%d launch %.2f us; phase clocks:", mode, best * 1e3f);
    const char* nm[6] = {"load+preupdate", "chol32", "L10/X0 solves + rank-32 updates", "chol32", "X1 solve", "store"};
    for (int i = 0; i < 6; ++i) printf(" %s %lld |", nm[i], st[i + 1] - st[i]);
    printf(" total %lld\n", st[6] - st[0]);
  }
  int hs[2]; cudaMemcpy(hs, status, 8, cudaMemcpyDeviceToHost);
  printf("status %d err %s\n", hs[0], cudaGetErrorString(cudaGetLastError()));
  return 0;
}
