// Internal (non-ABI) declarations shared between the .cu translation units.
#pragma once
#include "common.cuh"

namespace bx {

constexpr int NB = 64;  // linear-algebra block size (Cholesky panel, GEMM tile)

// K-range restriction of one 64x64 output tile at (i0, j0); exploits triangular operands.
enum KMode {
  KM_FULL = 0,
  KM_A_LOWER = 1,   // A lower triangular (NN): k < i0 + 64
  KM_B_LOWER = 2,   // B lower triangular (NN): k >= j0
  KM_TN_BOTH = 3,   // A^T B with both lower triangular: k >= max(i0, j0)
  KM_TN_A = 4,      // A^T B with A lower triangular: k >= i0
  KM_BT_LOWER = 5,  // A B^T with B (stored [N x K]) lower triangular: k < j0 + 64
};

struct GemmArgs {
  const float* A;
  const float* B;
  float* C;
  int lda, ldb, ldc;
  int M, N, K;  // multiples of 64
  float alpha, beta;
  int kmode;
  int lower_only;  // compute only tiles with i0 >= j0
  // tree-batched mode for the recursive triangular inverse (tree_nb > 0): blockIdx.z = node of `tree_level`
  int tree_nb, tree_level, tree_pass;
  // deterministic split-K (64x64 kernel only, ksplit > 0, beta = 0): blockIdx.z = K chunk [z ksplit, (z+1) ksplit); chunk z
  // writes its partial product to C + z * csplit (zeros where its K range is empty); the caller sums the chunks in order
  int ksplit;
  size_t csplit;
};

// C = alpha * op(A) * op(B) + beta * C on the SIMT fp32 pipes (row-major, 64x64x16 tiles).
int32_t gemm(cudaStream_t s, bool ta, bool tb, const GemmArgs& g, int grid_z = 1);

// In-place lower Cholesky of the np x np matrix A (row-major, lda); also the inverses of the 64x64 diagonal blocks
// (Dinv [np/64][64][64]) and per-block sum(log diag) (logdet_part [np/64]).  status[0] |= 1 on a bad pivot.
// `la` (optional): look-ahead.  The serial chain of a panel (potrf, panel solves, the update of the NEXT panel's columns)
// runs on la->chain, a high-priority helper stream forked from and joined to `s`, while `s` itself carries the rest of
// every trailing update - the chain of panel p+1 overlaps the bulk update of panel p.
struct Lookahead {
  cudaStream_t chain = nullptr;
  cudaEvent_t ev_chain = nullptr, ev_bulk = nullptr;
  int32_t init();  // created per call and destroyed on return (work still queued keeps the resources alive): no library state
  ~Lookahead();
};
bool lookahead_wanted(int np);
int32_t cholesky_prepare();  // call once per entry point, outside stream capture, before cholesky_inplace
// `scratch` (optional, (np/128 + 1) * 4096 floats): lets the first column of a panel also update the second one.
int32_t cholesky_inplace(cudaStream_t s, float* A, int np, int lda, float* Dinv, float* logdet_part, int* status,
                         const Lookahead* la = nullptr, float* scratch = nullptr);
// In place L -> T = L^-1 (upper triangle zeroed), using Dinv from cholesky_inplace and a np x np scratch matrix.
int32_t tri_inverse_inplace(cudaStream_t s, float* A, int np, int lda, const float* Dinv, float* TMP);
// out = T * x (lower) and out = T^T * x
int32_t trmv_lower(cudaStream_t s, const float* T, int np, int ld, const float* x, float* out);
int32_t trmv_lower_t(cudaStream_t s, const float* T, int np, int ld, const float* x, float* out);

// Workspace carve-up for the fit path.
struct FitWs {
  float *A, *TMP, *Dinv, *logdet_part, *yc, *v, *alpha, *U, *scale, *hyp, *part, *gbuf, *adam;
  int* status;
  int np, d4, nb, ntiles;
  size_t bytes;
  const Lookahead* la;  // optional, set by the entry point (never by carve_fit_ws)
};
FitWs carve_fit_ws(void* base, int n, int d);

// One evaluation of the posterior factor from raw hypers: fills A(=T), alpha, U, scale, hyp.  If T_out != nullptr the
// factor is built directly in T_out instead of ws.A.
int32_t launch_factor(cudaStream_t s, const float* d_X, const float* d_y, int n, int d, const float* d_raw, FitWs& w,
                      float* T_out, float* U_out, float* alpha_out, float* scale_out, float* hyp_out);

// Arguments of the scoring kernels (both engines).
struct ScoreArgs {
  bx_factor_t f;
  const float* cand;  // explicit candidates [M x d] (GEN == false)
  uint64_t m_begin, m_count;
  int acq_id;
  float acq_param, ystar;
  float* mu;
  float* sigma;
  float* acq;
  unsigned long long* best;
  uint64_t index_base;
  int debug;  // profiling ablations of the tcgen05 engine (env BX_TC_ABLATE): 1 = skip K* math, 2 = skip MMA issue
  // fused per-CTA top-k (topk_fused.cuh; fp16 tcgen05 engine only): slab of CTA b = topk_slab + b * tk::CAP keys,
  // topk_count[b] = survivors it left there; topk_k == 0 switches it off
  unsigned long long* topk_slab;
  unsigned int* topk_count;
  int topk_k;
  // hand-over of cancellation-regime candidates from the fp16 tcgen05 engine to the fp32 SIMT engine: the tensor engine
  // appends the launch-relative row of every candidate whose variance is below flag_tau * amp to flag_list (and writes no
  // output for it); the SIMT engine then scores exactly the rows flag_list[0 .. *flag_count)
  uint32_t* flag_list;
  unsigned int* flag_count;
  float flag_tau;
  // list mode of the scoring engines: the launch scores rows in_list[0 .. *in_count) (a hand-over list) instead of
  // [0, m_count); outputs and keys still go by the row
  const uint32_t* in_list;
  const unsigned int* in_count;
  // fast mode of the fp16 tcgen05 engine: its centred copy of the observation rows, rebuilt in the caller's workspace by
  // every fast-mode launch (score_tc16p.cu: ux_rows_kernel), and the norm bound below which it expands the squared distance
  float* ux;
  float expand_smax;
};

constexpr int BX_UX_MAX_NP = 32768;  // largest padded n the fp16 tcgen05 engine takes (sizes the ux part of the workspaces)
size_t score_tc16_ux_bytes();

// Fused-top-k plumbing shared by score.cu / topk.cu / score_tc16p.cu
unsigned score_tc16p_grid(uint64_t m_count);  // CTAs the fp16 engine launches for m_count candidates (slabs to merge)
unsigned score_simt_grid(uint64_t m_count, int sms);
// Reduce `nslabs` key slabs (`stride` keys apart, counts[b] valid keys each; counts == nullptr: every non-zero key is
// valid) to the sorted list format: list_out[0..k) ascending keys (empty slots = 0 first), list_out[k] = max of the
// n_best keys best_in[i * best_stride].  Optional unpacked outputs: the `count` valid entries ascending at the front of
// out_vals / out_idx.  scratch: nslabs * min(stride, k) keys.
int32_t launch_topk_merge(cudaStream_t s, const unsigned long long* slabs, const unsigned int* counts, int nslabs, int stride,
                          int k, unsigned long long* scratch, unsigned long long* list_out, const unsigned long long* best_in,
                          int n_best, int best_stride, float* out_vals, uint32_t* out_idx, int32_t* out_count);
// Top-k of a value vector by radix select (bx_topk) into the same list format (list_out[k] is left untouched).
int32_t launch_topk_values(cudaStream_t s, const float* d_vals, uint64_t M, int k, uint64_t index_base, float* d_out_vals,
                           uint32_t* d_out_idx, unsigned long long* list_out, void* d_ws);

}  // namespace bx
