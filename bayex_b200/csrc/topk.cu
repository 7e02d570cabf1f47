// Exact top-k of the acquisition vector - replaces `jnp.argsort(acq_vals)[-50:]` (optimizer.py:196).
//
// Keys are 64-bit (orderable(value) << 32 | position): unique, ordered by (value, position), NaN largest - the order
// of a stable ascending argsort, so "the last k" is well defined even when millions of values tie (PI == 0).
// Radix select, most-significant byte first: 8 histogram passes over the M values (HBM-bound, 4 B/candidate/pass),
// then one gather pass and a single-block bitonic sort of the k survivors.
#include "internal.h"

namespace bx {

struct TopkState {
  unsigned long long prefix;  // high bytes of the k-th largest key found so far
  unsigned int k_rem;         // rank still to resolve inside the prefix bucket
  unsigned int count;         // gather cursor
  unsigned int hist[256];
};

__device__ __forceinline__ unsigned long long topk_key(float v, unsigned int pos) {
  return ((unsigned long long)orderable_u32(v) << 32) | pos;
}

__global__ void topk_init_kernel(TopkState* st, int k) {
  if (threadIdx.x == 0) { st->prefix = 0ull; st->k_rem = (unsigned)k; st->count = 0u; }
  st->hist[threadIdx.x] = 0u;
}

__global__ void __launch_bounds__(256) topk_hist_kernel(const float* __restrict__ vals, uint64_t M, int pass, TopkState* st) {
  __shared__ unsigned int h[256];
  h[threadIdx.x] = 0u;
  __syncthreads();
  const unsigned long long prefix = st->prefix;
  const int shift = 56 - 8 * pass;
  for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < M; i += (uint64_t)gridDim.x * 256) {
    const unsigned long long key = topk_key(vals[i], (unsigned)i);
    if (pass == 0 || (key >> (shift + 8)) == prefix) atomicAdd(&h[(key >> shift) & 0xFF], 1u);
  }
  __syncthreads();
  if (h[threadIdx.x]) atomicAdd(&st->hist[threadIdx.x], h[threadIdx.x]);
}

__global__ void topk_select_kernel(TopkState* st) {
  if (threadIdx.x == 0) {
    unsigned int rem = st->k_rem, b = 255;
    for (;; --b) {
      const unsigned int c = st->hist[b];
      if (c >= rem || b == 0) break;
      rem -= c;
    }
    st->prefix = (st->prefix << 8) | b;
    st->k_rem = rem;
  }
  __syncthreads();
  st->hist[threadIdx.x] = 0u;
}

__global__ void __launch_bounds__(256) topk_gather_kernel(const float* __restrict__ vals, uint64_t M, TopkState* st,
                                                          unsigned long long* keys, int k) {
  const unsigned long long thr = st->prefix;  // after 8 passes: the k-th largest key itself
  for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < M; i += (uint64_t)gridDim.x * 256) {
    const unsigned long long key = topk_key(vals[i], (unsigned)i);
    if (key >= thr) {
      const unsigned int slot = atomicAdd(&st->count, 1u);
      if (slot < (unsigned)k) keys[slot] = key;
    }
  }
}

// Ascending bitonic sort of k (<= 1024) keys in one block, then unpack.
__global__ void __launch_bounds__(1024) topk_sort_kernel(const unsigned long long* keys, int k, uint64_t index_base,
                                                         float* out_vals, uint32_t* out_idx) {
  __shared__ unsigned long long s[1024];
  const int t = threadIdx.x;
  s[t] = (t < k) ? keys[t] : ~0ull;  // padding sorts last
  __syncthreads();
  for (int size = 2; size <= 1024; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      const int p = t ^ stride;
      if (p > t) {
        const bool up = ((t & size) == 0);
        const unsigned long long a = s[t], b = s[p];
        if ((a > b) == up) { s[t] = b; s[p] = a; }
      }
      __syncthreads();
    }
  }
  if (t < k) {
    out_vals[t] = from_orderable_u32((unsigned)(s[t] >> 32));
    out_idx[t] = (uint32_t)(index_base + (s[t] & 0xFFFFFFFFull));
  }
}

}  // namespace bx

using namespace bx;

extern "C" size_t bx_topk_workspace_bytes(uint64_t M, int32_t k) {
  (void)M;
  if (k <= 0 || k > BX_MAX_K) return 0;
  return 2048 + sizeof(unsigned long long) * BX_MAX_K;
}

extern "C" int32_t bx_topk(const float* d_vals, uint64_t M, int32_t k, uint64_t index_base, float* d_out_vals,
                           uint32_t* d_out_idx, void* d_ws, size_t ws_bytes, void* stream) {
  BX_CHECK_ARG(d_vals && d_out_vals && d_out_idx && d_ws, BX_E_ARG, "null pointer argument");
  BX_CHECK_ARG(k > 0 && k <= BX_MAX_K, BX_E_RANGE, "k=%d outside [1, %d]", k, BX_MAX_K);
  BX_CHECK_ARG((uint64_t)k <= M, BX_E_RANGE, "k=%d exceeds M=%llu", k, (unsigned long long)M);
  BX_CHECK_ARG(M <= 0x100000000ull && index_base + M <= 0x100000000ull, BX_E_RANGE, "indices must stay below 2^32");
  BX_CHECK_ARG(ws_bytes >= bx_topk_workspace_bytes(M, k), BX_E_WORKSPACE, "workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  TopkState* st = reinterpret_cast<TopkState*>(d_ws);
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(d_ws) + 2048);
  const unsigned grid = (unsigned)((M + 256ull * 16 - 1) / (256ull * 16) < 148ull * 8 ? (M + 256ull * 16 - 1) / (256ull * 16) : 148ull * 8);
  topk_init_kernel<<<1, 256, 0, s>>>(st, k);
  BX_LAUNCH_CHECK();
  for (int pass = 0; pass < 8; ++pass) {
    topk_hist_kernel<<<grid, 256, 0, s>>>(d_vals, M, pass, st);
    BX_LAUNCH_CHECK();
    topk_select_kernel<<<1, 256, 0, s>>>(st);
    BX_LAUNCH_CHECK();
  }
  topk_gather_kernel<<<grid, 256, 0, s>>>(d_vals, M, st, keys, k);
  BX_LAUNCH_CHECK();
  topk_sort_kernel<<<1, 1024, 0, s>>>(keys, k, index_base, d_out_vals, d_out_idx);
  BX_LAUNCH_CHECK();
  return 0;
}
