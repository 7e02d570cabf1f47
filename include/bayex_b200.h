/*
 * bayex_b200 - C ABI of the B200-native GP-posterior + acquisition hot path.
 *
 * This is the drop-in boundary.  bayex (the reference, /root/reference) has no
 * FFI of its own: its boundary is the Python API.  Each
 * entry point below replaces the XLA work behind one reference call site; the
 * citation after "replaces:" is the reference file:line.  The Python package
 * `bayex_b200` mirrors the reference's names on top of these with ctypes
 * (see INTEGRATION.md for the binding a bayex maintainer would add).
 *
 * Conventions
 *   - every pointer named d_* is a DEVICE pointer owned by the caller; the
 *     library allocates nothing persistent and keeps no global mutable state;
 *   - `stream` is a cudaStream_t passed as void*; all work is stream-ordered and
 *     no entry point synchronises the host unless stated;
 *   - return value: 0 = ok, <0 = bad argument (BX_E_*), >0 = cudaError_t;
 *     bx_last_error() gives a thread-local message;
 *   - observations are COMPACTED by the caller (valid rows only, reference mask
 *     applied on the host side: gp.py:43-47 padding rows carry 1e12 variance and
 *     contribute <1e-12, tests/test_gp.py:54-99).
 */
#ifndef BAYEX_B200_H
#define BAYEX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BX_VERSION 100 /* 0.1.0 */

#define BX_E_ARG (-1)       /* null pointer / non-positive size */
#define BX_E_RANGE (-2)     /* size outside supported range (d > BX_MAX_D, k > BX_MAX_K ...) */
#define BX_E_WORKSPACE (-3) /* workspace too small */
#define BX_E_NOTPD (-4)     /* non-finite / non-positive Cholesky pivot: returned by bx_check_status */
#define BX_E_UNSUPPORTED (-5)

#define BX_MAX_D 64   /* parameter-space dimensions */
#define BX_MAX_K 1024 /* top-k */
#define BX_TILE 128   /* observation padding granule: factor matrices are np x np, np = roundup(n, 128) */

/* acquisition ids - names follow bayex.acq (acq.py:8,53,90,124) */
enum { BX_ACQ_EI = 0, BX_ACQ_PI = 1, BX_ACQ_UCB = 2, BX_ACQ_LCB = 3 };
/* scoring engines: one per regime, no zoo (superseded experiments, incl. the 3xTF32 engines, live in scripts/dev/engines) */
enum {
  BX_ENGINE_AUTO = 0,           /* the production path when the factor has its fp16 copies and d <= 32, else SIMT */
  BX_ENGINE_SIMT = 1,           /* fp32 FFMA pipes: explicit point sets, d > 32, cancellation-regime candidates */
  BX_ENGINE_TCGEN05_F16 = 2,    /* the production path: tcgen05 with fp16 head/tail operands (3 MMAs per MAC, the same 22
                                   operand bits as a 3xTF32 split at half the tensor work), cta_group::2, register-tiled K*
                                   producers.  Fast mode over every candidate (whole-row sums inside the tensor core);
                                   the candidates whose variance sits in the cancellation regime (var < amp / 2) are
                                   re-scored in the accurate mode (per-K-block chunks folded into TMEM-resident fp32
                                   running sums with round-to-nearest adds) */
  BX_ENGINE_TCGEN05_F16_RAW = 3, /* the fast mode alone, no hand-over (diagnostics: its own accuracy) */
  BX_ENGINE_TCGEN05_F16_ACC = 4  /* the accurate mode alone over every candidate */
};

/* One dimension of the parameter space - mirrors bayex.domain.Real / Integer (domain.py:22-131). */
typedef struct {
  int32_t kind; /* 0 = Real, 1 = Integer */
  float lo, hi; /* Real bounds as float32 */
  int32_t ilo, ihi; /* Integer bounds, inclusive */
} bx_dim_t;

/*
 * The cached posterior factor (what gp.py:41-49 recomputes on every predict, quirk Q7).
 * All device memory, caller-owned.  np = roundup(n, 128); rows/cols >= n are zero.
 */
typedef struct {
  int32_t n;          /* valid observations */
  int32_t np;         /* padded size, multiple of 128 */
  int32_t d;          /* dimensions */
  int32_t d4;         /* row stride of d_U, multiple of 4 */
  const float* d_U;   /* [np x d4] observations * sqrt(log2 e) / lengthscale */
  const float* d_T;   /* [np x np] row-major, T = L^-1 (lower triangular), K = L L^T */
  const void* d_T16hi;  /* [np x np] fp16 head of T * 2^b (tcgen05 fp16 engine), may be NULL; 2^b is d_hyp[8] */
  const void* d_T16lo;  /* [np x np] fp16 tail, may be NULL */
  const float* d_alpha; /* [np] K^-1 (y - ymean) */
  const float* d_scale; /* [d4] sqrt(log2 e) / lengthscale_k */
  const float* d_hyp;   /* [16] amp, noise, ymean, log2(amp), nlml, logdet(L), quad, status, T16 scale, (7 reserved) */
} bx_factor_t;

int32_t bx_version(void);
const char* bx_last_error(void);

/* Status of the last factorisation recorded in a factor's d_hyp[7] (a device read: synchronises `stream`).  Returns 0 or
 * BX_E_NOTPD.  The reference's cholesky returns NaN silently; the Python layer calls this after every
 * build and warns - it does not raise, the NaNs propagate exactly as they do upstream. */
int32_t bx_check_status(const float* d_hyp, void* stream);

/* Candidate generation - replaces: domain.py:159-161 (split), :74-75 (uniform+clip), :130-131 (randint+round/clip).
 * Writes candidates [m_begin, m_begin+m_count) of the draw `sample_params(key, (M,))` as float32 [m_count x d]. */
int32_t bx_threefry_fill(const uint32_t key[2], const bx_dim_t* dims, int32_t d, uint64_t m_begin, uint64_t m_count,
                         float* d_out, void* stream);
/* One domain drawn with the key itself - replaces: Real.sample (domain.py:63-75) / Integer.sample (domain.py:119-131). */
int32_t bx_domain_sample(const uint32_t key[2], const bx_dim_t* dim, uint64_t m_begin, uint64_t m_count, float* d_out,
                         void* stream);
/* Regenerate the coordinates of selected candidate indices (top-k start points, optimizer.py:197). */
int32_t bx_threefry_gather(const uint32_t key[2], const bx_dim_t* dims, int32_t d, const uint32_t* d_idx, int32_t k,
                           float* d_out, void* stream);

/* Masked pairwise squared-exponential matrix - replaces: gp.py:15-23 exp_quadratic / cov.
 * x1 [n1 x d], x2 [n2 x d] (already divided by the length scale), masks as float; out [n1 x n2] row-major. */
int32_t bx_cov(const float* d_x1, int32_t n1, const float* d_x2, int32_t n2, int32_t d, const float* d_m1,
               const float* d_m2, float* d_out, void* stream);

/* Workspace sizes (bytes) for the fit-path entry points. */
size_t bx_fit_workspace_bytes(int32_t n, int32_t d);

/* NLML and its gradient - replaces: gp.py:74 marginal_likelihood, gp.py:75 grad_fun, gp.py:78-87 neg_log_likelihood.
 * d_raw = [noise_raw, amp_raw, ls_raw[d]]; d_out[0]=nlml, [1]=loss, [2..4+d) = d loss/d raw, [4+d..6+2d) = d nlml/d raw. */
int32_t bx_nlml_grad(const float* d_X, const float* d_y, int32_t n, int32_t d, const float* d_raw, float* d_out,
                     void* d_ws, size_t ws_bytes, void* stream);

/* Hyper-parameter fit - replaces: gp.py:90-110 posterior_fit (clip_by_global_norm(10) + adamw(lr), fresh moments).
 * d_raw is updated in place; the whole loop runs on the device (one captured step replayed `steps` times: no host
 * round-trip inside the loop).  EXCEPTION to the no-sync convention: with use_graph != 0 the call synchronises `stream`
 * before it returns (the executable graph it created must outlive its queued launches; the library keeps no state that
 * could own it).  d_trace (optional) receives [steps x (2+d)] raw params before each step (plain launches, no graph). */
int32_t bx_fit_adamw(const float* d_X, const float* d_y, int32_t n, int32_t d, float* d_raw, float lr, int32_t steps,
                     float* d_trace, void* d_ws, size_t ws_bytes, int32_t use_graph, void* stream);

#define BX_FIT_BATCH_MAX 32 /* restarts fitted side by side by one bx_fit_adamw_batch call */

/* posterior_fit from R starting points side by side - replaces: R independent gp.py:90-110 posterior_fit calls on the same
 * observations (BASELINE config 5: the hyper-parameter restart sweep).  d_raws [R x (d+2)] is updated in place.  One
 * step of all R restarts is one CUDA graph with R parallel branches (side streams forked from and joined to `stream`),
 * so the serial potrf / panel-solve chain of one restart overlaps with the GEMMs of the others; every restart's
 * arithmetic is the one bx_fit_adamw does, bit for bit.  d_out (optional) [R x (2 + 2(d+2))] receives the
 * bx_nlml_grad row at the fitted hypers of every restart.  d_ws: R * bx_fit_workspace_bytes(n, d) bytes. */
int32_t bx_fit_adamw_batch(const float* d_X, const float* d_y, int32_t n, int32_t d, float* d_raws, int32_t R, float lr,
                           int32_t steps, float* d_out, void* d_ws, size_t ws_bytes, void* stream);

/* Posterior factor - replaces: gp.py:41-49 (softplus, centring, Gram, Cholesky, alpha) once per fit.
 * Outputs are the arrays referenced by bx_factor_t (caller allocates: U np*d4, T np*np, T16hi/T16lo
 * optional, alpha np, scale d4, hyp 16). */
int32_t bx_factor_build(const float* d_X, const float* d_y, int32_t n, int32_t d, const float* d_raw, float* d_U,
                        float* d_T, void* d_T16hi, void* d_T16lo, float* d_alpha,
                        float* d_scale, float* d_hyp, void* d_ws, size_t ws_bytes, void* stream);

/* Append one observation to a built factor at UNCHANGED hyper-parameters - SURVEY 8(f) rank 1, the incremental refit of
 * optimizer.py:261-269 (`_fit` inserts one point; gp.py:46-49 then rebuilds and refactorises the whole Gram matrix).  The
 * matrix only grows by one bordered row, so its inverse factor does too: l = T k, lambda^2 = kappa - |l|^2,
 * T' = [T 0; -(l^T T)/lambda  1/lambda] - O(n^2) instead of O(n^3).  `f` is a factor bx_factor_build filled, with a free
 * padding row (f->n < f->np, else BX_E_RANGE: rebuild); d_x_new [d] raw coordinates of the new point; d_y [f->n + 1] all
 * observed values, the new one last (the mean of gp.py:43 moves, so alpha = K^-1 (y - mean) is recomputed in full, with the
 * same refinement step as the build).  Rewrites row f->n of U and T, alpha, hyp (mean, NLML, log det, status, T16 scale) and
 * the fp16 copies in place; afterwards the factor is valid with n = f->n + 1 (the caller bumps its copy of the struct). */
size_t bx_factor_append_workspace_bytes(int32_t np);
int32_t bx_factor_append(const bx_factor_t* f, const float* d_x_new, const float* d_y, void* d_ws, size_t ws_bytes, void* stream);

#define BX_POINTS_TC_MIN 65536 /* bx_score_points: point sets at least this large use the tcgen05 path if the factor has its fp16 copies */

/* Hand-over workspace of the two-tier production path (bytes): the fp16 tcgen05 engine lists the candidates whose variance is
 * below half the prior variance (where amp - sum v^2 cancels and the truncating fp32 accumulation of tcgen05.mma would
 * show), the fp32 SIMT engine scores exactly those. */
size_t bx_score_workspace_bytes(uint64_t M);

/* Posterior + acquisition at explicit points - replaces: gp.py:59-71 predict, acq.py:45-50,84-87,120-121,155-156.
 * d_cand [M x d] raw (unscaled) points.  Any of d_mu / d_sigma / d_acq may be NULL.  d_best (optional) is a
 * uint64 packed arg-max accumulator (see bx_pack); index = index_base + row.  d_ws (optional, bx_score_workspace_bytes(M)):
 * without it every point is scored on the fp32 SIMT engine. */
int32_t bx_score_points(const bx_factor_t* f, const float* d_cand, uint64_t M, int32_t acq_id, float acq_param,
                        float ystar, float* d_mu, float* d_sigma, float* d_acq, uint64_t* d_best,
                        uint64_t index_base, void* d_ws, size_t ws_bytes, void* stream);

/* Fused generate -> score -> arg-max over candidates [m_begin, m_begin+m_count) of sample_params(key, (M,)) -
 * replaces: optimizer.py:183-193 + :205.  No M x n intermediate is written; d_acq (optional, [m_count]) receives the
 * acquisition values.  d_best must be zero-initialised by the caller (or hold a running max).  d_ws:
 * bx_score_workspace_bytes(m_count), needed by BX_ENGINE_AUTO / BX_ENGINE_TCGEN05_F16. */
int32_t bx_score_generated(const bx_factor_t* f, const uint32_t key[2], const bx_dim_t* dims, uint64_t m_begin,
                           uint64_t m_count, int32_t acq_id, float acq_param, float ystar, float* d_acq,
                           uint64_t* d_best, int32_t engine, void* d_ws, size_t ws_bytes, void* stream);

/* Top-k by (value, index) ascending order, i.e. the last k of a stable argsort - replaces: optimizer.py:196.
 * NaN sorts last (largest).  Outputs ascending: d_out_idx[k-1] is the arg-max.  index = index_base + position. */
size_t bx_topk_workspace_bytes(uint64_t M, int32_t k);
int32_t bx_topk(const float* d_vals, uint64_t M, int32_t k, uint64_t index_base, float* d_out_vals,
                uint32_t* d_out_idx, void* d_ws, size_t ws_bytes, void* stream);

/* The list format of the fused top-k: uint64 [k + 1].  Entries [0, k) are the k largest (orderable(value) << 32 | global
 * candidate index) keys in ascending order - the order of `jnp.argsort(acq_vals)[-k:]` (optimizer.py:196: stable, NaN
 * last, the higher index wins among equal values); when fewer than k candidates exist the leading slots are 0.  Entry [k]
 * is the packed arg-max key of the same candidates (bx_pack: first occurrence wins, optimizer.py:205). */

/* Fused generate -> score -> top-k -> arg-max over candidates [m_begin, m_begin + m_count) of sample_params(key, (M,)) -
 * replaces: optimizer.py:183-197 + :205.  On the fp16 tcgen05 engine the top-k is kept per CTA
 * inside the scoring epilogue and nothing per candidate is written to HBM (d_acq == NULL); d_acq (optional, [m_count])
 * additionally receives every acquisition value.  d_vals / d_idx (optional): the valid entries of the list unpacked,
 * ascending, at the front of d_vals [k] / d_idx [k].  d_count (optional, int32 [2]): their number, and the number of
 * candidates the fast mode handed to the second tier (a caller that sees most of a launch handed over can ask for
 * BX_ENGINE_TCGEN05_F16_ACC next time and save the first pass). */
size_t bx_score_topk_workspace_bytes(uint64_t m_count, int32_t k);
int32_t bx_score_generated_topk(const bx_factor_t* f, const uint32_t key[2], const bx_dim_t* dims, uint64_t m_begin,
                                uint64_t m_count, int32_t acq_id, float acq_param, float ystar, int32_t k, uint64_t* d_list,
                                float* d_vals, uint32_t* d_idx, int32_t* d_count, float* d_acq, int32_t engine, void* d_ws,
                                size_t ws_bytes, void* stream);
/* Merge n_lists lists (k + 1 entries each, contiguous) into one - the candidate shards of a multi-GPU sample() after their
 * all-gather.  d_ws: n_lists * k * 8 bytes.  Outputs as above. */
int32_t bx_topk_merge(const uint64_t* d_lists, int32_t n_lists, int32_t k, uint64_t* d_list_out, float* d_vals,
                      uint32_t* d_idx, int32_t* d_count, void* d_ws, size_t ws_bytes, void* stream);
/* The proposal - replaces: optimizer.py:203-207 (concat(refined, random), arg-max with first occurrence, gather).
 * d_opt_x [S x d] / d_opt_val [S]: refined points and their acquisition values; d_best: packed arg-max key of the random
 * candidates (entry [k] of a list).  d_out [d + 2]: the winner's raw coordinates (a random winner is regenerated from
 * the key), its acquisition value, and its position in the concatenation as uint32 bits. */
int32_t bx_propose(const uint32_t key[2], const bx_dim_t* dims, int32_t d, const float* d_opt_x, const float* d_opt_val,
                   int32_t S, const uint64_t* d_best, float* d_out, void* stream);

/* Multi-start refinement - replaces: optimizer.py:39-73,198-199 (L-BFGS, parity unpinned -> monotone analytic-gradient
 * ascent).  d_x [S x d] start points in, refined points out; d_val [S] acquisition at the result. */
size_t bx_refine_workspace_bytes(const bx_factor_t* f, int32_t S);
int32_t bx_refine(const bx_factor_t* f, float* d_x, int32_t S, int32_t rounds, int32_t acq_id, float acq_param,
                  float ystar, float* d_val, void* d_ws, size_t ws_bytes, void* stream);

/* Packed (value, index) key: max over keys == jnp.argmax (first occurrence, NaN wins) - optimizer.py:205. */
uint64_t bx_pack(float value, uint32_t index);
void bx_unpack(uint64_t key, float* value, uint32_t* index);

/* 3xTF32 tcgen05 self-test: D[128 x N] = A[128 x K] * B[N x K]^T through the same descriptors/pipeline pieces the
 * scoring kernel uses.  Device pointers, row-major. */
int32_t bx_selftest_umma(const float* d_A, const float* d_B, float* d_D, int32_t N, int32_t K, void* stream);

/* fp32 SIMT peak probe: `blocks` CTAs of 256 threads, 8 independent FFMA chains of `iters` steps per thread
 * (flops = blocks * 256 * iters * 16).  bench.py times it with CUDA events: the denominator of the fit path's roofline. */
int32_t bx_selftest_ffma(float* d_out, int32_t blocks, int32_t iters, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BAYEX_B200_H */
