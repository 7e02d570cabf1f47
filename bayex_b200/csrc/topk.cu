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
                                                         float* out_vals, uint32_t* out_idx, unsigned long long* list_out) {
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
    const uint32_t idx = (uint32_t)(index_base + (s[t] & 0xFFFFFFFFull));
    if (out_vals) out_vals[t] = from_orderable_u32((unsigned)(s[t] >> 32));
    if (out_idx) out_idx[t] = idx;
    if (list_out) list_out[t] = (s[t] & 0xFFFFFFFF00000000ull) | idx;  // the list format carries GLOBAL indices
  }
}

// ---- merge of key slabs (per-CTA survivors of the fused top-k, or the per-rank lists of a sharded sample()) ----------
// One block.  Valid keys are first compacted into `scratch` (block-wide prefix over the slab counts), then an exact radix
// select (8 passes, shared-memory histogram) finds the k-th largest, the survivors are sorted ascending (bitonic, 1024
// slots) and written in the list format: list_out[0..k) with empty (0) slots first, list_out[k] = max of the best keys.
__global__ void __launch_bounds__(1024) topk_merge_kernel(const unsigned long long* __restrict__ slabs, const unsigned int* __restrict__ counts,
                                                          int nslabs, int stride, int k, unsigned long long* scratch,
                                                          unsigned long long* list_out, const unsigned long long* __restrict__ best_in,
                                                          int n_best, int best_stride, float* out_vals, uint32_t* out_idx,
                                                          int32_t* out_count) {
  __shared__ unsigned long long srt[1024];
  __shared__ unsigned int hist[256];
  __shared__ unsigned int total_s, fill_s, rem_s;
  __shared__ unsigned long long prefix_s;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int per = min(stride, k);  // a slab never holds more than k valid keys (counts) / list entries (lists)
  if (t == 0) { total_s = 0u; fill_s = 0u; }
  __syncthreads();
  // compaction: warp w takes slabs w, w + 32, ...; order inside scratch is irrelevant (keys are unique)
  for (int b = warp; b < nslabs; b += 32) {
    const int c = counts ? min((int)counts[b], per) : per;
    for (int i0 = 0; i0 < c; i0 += 32) {
      const int i = i0 + lane;
      const unsigned long long key = i < c ? slabs[(size_t)b * stride + i] : 0ull;
      const unsigned int m = __ballot_sync(0xffffffffu, key != 0ull);
      unsigned int base = 0;
      if (lane == 0 && m) base = atomicAdd(&total_s, __popc(m));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (key != 0ull) scratch[base + __popc(m & ((1u << lane) - 1u))] = key;
    }
  }
  __syncthreads();
  const unsigned int N = total_s;
  unsigned long long thr = 0ull;  // keep keys >= thr
  if (N > (unsigned)k) {
    if (t == 0) { prefix_s = 0ull; rem_s = (unsigned)k; }
    for (int pass = 0; pass < 8; ++pass) {
      const int shift = 56 - 8 * pass;
      if (t < 256) hist[t] = 0u;
      __syncthreads();
      const unsigned long long prefix = prefix_s;
      for (unsigned int e = t; e < N; e += 1024) {
        const unsigned long long key = scratch[e];
        if (pass == 0 || (key >> (shift + 8)) == prefix) atomicAdd(&hist[(key >> shift) & 0xFF], 1u);
      }
      __syncthreads();
      if (t == 0) {
        unsigned int rem = rem_s, b = 255;
        for (;; --b) {
          const unsigned int c = hist[b];
          if (c >= rem || b == 0) break;
          rem -= c;
        }
        prefix_s = (prefix << 8) | b;
        rem_s = rem;
      }
      __syncthreads();
    }
    thr = prefix_s;
  }
  srt[t] = 0ull;  // empty slots sort first
  __syncthreads();
  for (unsigned int e = t; e < N; e += 1024) {
    const unsigned long long key = scratch[e];
    if (key >= thr) srt[atomicAdd(&fill_s, 1u)] = key;  // at most min(N, k) <= 1024 of them
  }
  __syncthreads();
  for (int size = 2; size <= 1024; size <<= 1) {
    for (int sd = size >> 1; sd > 0; sd >>= 1) {
      const int p = t ^ sd;
      if (p > t) {
        const bool up = ((t & size) == 0);
        const unsigned long long a = srt[t], b = srt[p];
        if ((a > b) == up) { srt[t] = b; srt[p] = a; }
      }
      __syncthreads();
    }
  }
  const int cnt = (int)fill_s;
  if (t < k) list_out[t] = srt[1024 - k + t];
  if (t < cnt) {
    const unsigned long long key = srt[1024 - cnt + t];
    if (out_vals) out_vals[t] = from_orderable_u32((unsigned)(key >> 32));
    if (out_idx) out_idx[t] = (uint32_t)(key & 0xFFFFFFFFull);
  }
  if (t == 0) {
    if (out_count) *out_count = cnt;
    unsigned long long best = 0ull;
    for (int i = 0; i < n_best; ++i) {
      const unsigned long long b = best_in[(size_t)i * best_stride];
      best = b > best ? b : best;
    }
    list_out[k] = best;
  }
}

}  // namespace bx

using namespace bx;

extern "C" size_t bx_topk_workspace_bytes(uint64_t M, int32_t k) {
  (void)M;
  if (k <= 0 || k > BX_MAX_K) return 0;
  return 2048 + sizeof(unsigned long long) * BX_MAX_K;
}

namespace bx {

int32_t launch_topk_values(cudaStream_t s, const float* d_vals, uint64_t M, int k, uint64_t index_base, float* d_out_vals,
                           uint32_t* d_out_idx, unsigned long long* list_out, void* d_ws) {
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
  topk_sort_kernel<<<1, 1024, 0, s>>>(keys, k, index_base, d_out_vals, d_out_idx, list_out);
  BX_LAUNCH_CHECK();
  return 0;
}

int32_t launch_topk_merge(cudaStream_t s, const unsigned long long* slabs, const unsigned int* counts, int nslabs, int stride,
                          int k, unsigned long long* scratch, unsigned long long* list_out, const unsigned long long* best_in,
                          int n_best, int best_stride, float* out_vals, uint32_t* out_idx, int32_t* out_count) {
  topk_merge_kernel<<<1, 1024, 0, s>>>(slabs, counts, nslabs, stride, k, scratch, list_out, best_in, n_best, best_stride, out_vals,
                                       out_idx, out_count);
  BX_LAUNCH_CHECK();
  return 0;
}

}  // namespace bx

extern "C" int32_t bx_topk(const float* d_vals, uint64_t M, int32_t k, uint64_t index_base, float* d_out_vals,
                           uint32_t* d_out_idx, void* d_ws, size_t ws_bytes, void* stream) {
  BX_CHECK_ARG(d_vals && d_out_vals && d_out_idx && d_ws, BX_E_ARG, "null pointer argument");
  BX_CHECK_ARG(k > 0 && k <= BX_MAX_K, BX_E_RANGE, "k=%d outside [1, %d]", k, BX_MAX_K);
  BX_CHECK_ARG((uint64_t)k <= M, BX_E_RANGE, "k=%d exceeds M=%llu", k, (unsigned long long)M);
  BX_CHECK_ARG(M <= 0x100000000ull && index_base + M <= 0x100000000ull, BX_E_RANGE, "indices must stay below 2^32");
  BX_CHECK_ARG(ws_bytes >= bx_topk_workspace_bytes(M, k), BX_E_WORKSPACE, "workspace too small");
  return launch_topk_values((cudaStream_t)stream, d_vals, M, k, index_base, d_out_vals, d_out_idx, nullptr, d_ws);
}

extern "C" int32_t bx_topk_merge(const uint64_t* d_lists, int32_t n_lists, int32_t k, uint64_t* d_list_out, float* d_vals,
                                 uint32_t* d_idx, int32_t* d_count, void* d_ws, size_t ws_bytes, void* stream) {
  BX_CHECK_ARG(d_lists && d_list_out && d_ws, BX_E_ARG, "null pointer argument");
  BX_CHECK_ARG(k > 0 && k <= BX_MAX_K && n_lists > 0, BX_E_RANGE, "k=%d n_lists=%d out of range", k, n_lists);
  BX_CHECK_ARG(ws_bytes >= sizeof(uint64_t) * (size_t)n_lists * k, BX_E_WORKSPACE, "workspace too small");
  return launch_topk_merge((cudaStream_t)stream, reinterpret_cast<const unsigned long long*>(d_lists), nullptr, n_lists, k + 1, k,
                           reinterpret_cast<unsigned long long*>(d_ws), reinterpret_cast<unsigned long long*>(d_list_out),
                           reinterpret_cast<const unsigned long long*>(d_lists) + k, n_lists, k + 1, d_vals, d_idx, d_count);
}
