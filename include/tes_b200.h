/*
 * tes_b200.h -- C ABI of libtes_b200.so: the B200 (sm_100a) implementation of the
 * Trusted-Maximizers Entropy Search acquisition hot path.
 *
 * The reference (ZhaoxuanWu/Trusted-Maximizers-Entropy-Search-BO) is pure Python on TensorFlow
 * 1.14 and has no FFI; its de-facto operator boundary is the set of Python functions the three
 * drivers call.  Each entry point below names the reference function it
 * replaces (file:line relative to the reference tree).  INTEGRATION.md shows the ctypes binding
 * a maintainer adds on the reference side.
 *
 * Conventions
 *   - all array pointers are DEVICE pointers owned by the caller, row-major (C order), fp64
 *     unless stated; sizes are int64_t; indices are int64_t;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); all work is
 *     enqueued on it and the call returns without synchronising unless stated;
 *   - no hidden device allocation: functions that need scratch take (ws, ws_bytes) and have a
 *     matching *_workspace_bytes() query; workspace must be 256-byte aligned;
 *   - return value: 0 = ok; < 0 = -(1-based position of the offending argument), -100 = workspace
 *     too small; > 0 = cudaError_t of a failed runtime call / launch;
 *   - re-entrant, no global mutable state; safe from several host threads on distinct streams;
 *   - kernel convention: `l` is an INVERSE SQUARED length-scale,
 *     k(x,x') = sigma * exp(-0.5 * sum_d l_d (x_d - x'_d)^2); sigma / sigma0 are variances;
 *   - ordering rule: arg-max / top-k ties resolve to the LOWER index.
 */
#ifndef TES_B200_H
#define TES_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TES_B200_ABI_VERSION 1

int tes_abi_version(void);
/* human-readable text for a return code of this library (static storage) */
const char* tes_error_string(int code);

/* ===================================================================================== Stage A
 * Random-Fourier-feature posterior function samples and their top-k over the candidate set.
 *
 * Replaces utils_for_continuous.sample_xmaxs_fmaxs body (utils_for_continuous.py:172-187:
 * fvals = sqrt(2 sigma/F) theta_j^T cos(W_j X*^T + b_j); tf.math.top_k) and its numpy twin
 * optfunc.make_function_sample_np (optfunc.py:389-402) + argmax.
 *
 *   cand   (N, d)            candidate inputs
 *   W      (S, F, d) or (1, F, d) when w_shared != 0      (optfunc.py:329)
 *   b      (S, F)    or (1, F)    when w_shared != 0      (optfunc.py:334)
 *   theta  (S, F)                                          (optfunc.py:365-383)
 *   out_idx (S, k) int64, out_val (S, k): per function sample the k best candidates, best first;
 *           reported indices are local index + idx_offset (the global index of cand[0] when the
 *           candidate set is sharded); slots beyond N hold (-inf, -1).
 * Arithmetic follows include/tes_spec.h exactly (fp64, fixed summation order), so the indices
 * are bit-identical to the oracle's.  1 <= d <= 64, 1 <= k <= 16.
 */
int64_t tes_rff_topk_workspace_bytes(int64_t N, int64_t d, int64_t S, int64_t F, int k);
int tes_rff_topk_f64(const double* cand, int64_t N, int64_t d, const double* W, const double* b,
                     const double* theta, int64_t S, int64_t F, int w_shared, double sigma, int k,
                     int64_t idx_offset, int64_t* out_idx, double* out_val, void* ws, int64_t ws_bytes,
                     void* stream);

/* Same as tes_rff_topk_f64 with an explicit evaluation method (results are bit-identical for every method):
 *   0 = automatic; 1 = every pair in fp64; 2 = fp32 screen + fp64 refine: all pairs are first scored on the
 *   packed-fp32 + MUFU pipes with a proven per-sample error bound B_j, every candidate within 2 B_j of the
 *   running k-th best screen value is re-evaluated with the exact fp64 sequence of tes_spec.h, and a
 *   (chunk, sample) whose kept set overflows is recomputed in fp64 (needs d <= 10 and k <= 8, else = 1);
 *   3 = as 2, but the d-term inner products of the screen run on the TF32 tensor cores in split precision
 *   (3xTF32, mma.sync m16n8k8), leaving the MUFU cosine as the bound. */
int tes_rff_topk_method_f64(const double* cand, int64_t N, int64_t d, const double* W, const double* b,
                            const double* theta, int64_t S, int64_t F, int w_shared, double sigma, int k,
                            int64_t idx_offset, int64_t* out_idx, double* out_val, void* ws, int64_t ws_bytes,
                            void* stream, int method);

/* Stage A on PRODUCT-STRUCTURED candidate sets (the meshes of the discrete drivers: functions.get_meshgrid,
 * bo_discrete.py:928-941): row-major U x V, the coordinates in columns [s_lo, s_hi) depend on i / nv only and the other
 * columns on i % nv only.  Same result as tes_rff_topk_f64 (bit-identical values, indices, tie rule), computed as one fp64 tensor-pipe
 * GEMM per function sample (cos(a + b) = cos a cos b - sin a sin b) + exact re-evaluation of the k + 8 best.
 * out_flag (S): 1 where the GEMM ranking could not be proven to contain the true top-k (exact ties, NaN); the caller
 * then recomputes with tes_rff_topk_f64 (the Python binding does).  k <= 8.  The structure itself is NOT checked here.
 * tf32 = 1 / 2: the GEMM runs as a 3xTF32 / 3xFP16 screen on mma.sync tensor cores (operands split into hi + lo pairs
 * of 11 significant bits, fp32 accumulation flushed into fp64 every 256 k, a correspondingly wider safety margin; fp16:
 * theta normalised per sample); 3: the 3xFP16 GEMM on tcgen05.mma kind::f16 + TMEM with bulk-async-copy operand
 * staging; 0: fp64 (DMMA).  The result is the same in every mode. */
int64_t tes_rff_topk_product_workspace_bytes(int64_t N, int64_t d, int64_t nv, int64_t S, int64_t F, int k, int tf32);
int tes_rff_topk_product_f64(const double* cand, int64_t N, int64_t d, int64_t s_lo, int64_t s_hi, int64_t nv,
                             const double* W, const double* b, const double* theta, int64_t S, int64_t F, int w_shared, double sigma,
                             int k, int64_t idx_offset, int64_t* out_idx, double* out_val, int* out_flag, int tf32,
                             void* ws, int64_t ws_bytes, void* stream);

/* Resolve pass for the samples tes_rff_topk_product_f64 flags because of NEAR-TIES (more than 8 candidates within the
 * error bound of the k-th best: meshes that are fine along a direction the function sample is flat in).  W / b / theta
 * of the flagged samples only (S of them).  The CTA-pair tcgen05 GEMM runs twice on them: top-k epilogue (k-th best
 * value vk, bound B), then a threshold epilogue that emits every candidate whose GEMM value reaches vk - 2B -- a
 * superset of the true top-k -- followed by the exact fp64 sequence of tes_spec.h on the emitted candidates and the
 * (value desc, index asc) ranking: the bits of tes_rff_topk_f64 at ~3 % of its cost per sample.  out_flag: 1 where more
 * than 2048 candidates pass the threshold or NaN is involved (recompute with tes_rff_topk_f64).  d <= 16, k <= 8.
 * Replaces nothing in the reference (utils_for_continuous.py:172-187 evaluates every candidate in fp64). */
int64_t tes_rff_topk_product_resolve_workspace_bytes(int64_t N, int64_t d, int64_t nv, int64_t S, int64_t F, int k);
int tes_rff_topk_product_resolve_f64(const double* cand, int64_t N, int64_t d, int64_t s_lo, int64_t s_hi, int64_t nv,
                                     const double* W, const double* b, const double* theta, int64_t S, int64_t F,
                                     int w_shared, double sigma, int k, int64_t idx_offset, int64_t* out_idx,
                                     double* out_val, int* out_flag, void* ws, int64_t ws_bytes, void* stream);

/* Detection of the product structure above on the device (the meshes come from functions.get_meshgrid,
 * bo_discrete.py:928-941; the reference has no counterpart: it never exploits the structure).
 * runlen: out_r[k] = first row whose coordinate k differs from row 0 (N if the column is constant) -- the block length nv
 * of a split is the minimum over its slow columns.  check: for a given nv, out_bad[k] = 1 unless column k is constant
 * inside every block of nv rows ("slow"), out_bad[d + k] = 1 unless row i equals row i % nv in column k ("fast").
 * IEEE == comparisons (NaN never equal).  out_r: d int64, out_bad: 2 d int32, device memory. */
int tes_product_runlen_f64(const double* cand, int64_t N, int64_t d, int64_t* out_r, void* stream);
int tes_product_check_f64(const double* cand, int64_t N, int64_t d, int64_t nv, int* out_bad, void* stream);
/* "Mesh + tail": the continuous drivers append top_init_k uniform candidates to the grid they score
 * (utils_for_continuous.py:36-38).  out_first[k] = first row at which column k stops being "slow" for block length nv,
 * out_first[d + k] = first row at which it stops being "fast" (N: never), so that the longest prefix that IS a product
 * set can take the GEMM path and only the remaining rows the general kernel.  d <= 32; out_first: 2 d int64. */
int tes_product_prefix_f64(const double* cand, int64_t N, int64_t d, int64_t nv, int64_t* out_first, void* stream);

/* All function-sample values, out (S, N).  Replaces utils_for_continuous.get_function_samples
 * (utils_for_continuous.py:257-296) / optfunc.make_function_sample_np.  Debug / small-N use. */
int tes_rff_eval_f64(const double* cand, int64_t N, int64_t d, const double* W, const double* b,
                     const double* theta, int64_t S, int64_t F, int w_shared, double sigma, double* out,
                     void* stream);

/* Merge P partial top-k lists (e.g. the NCCL all-gathered per-GPU results of tes_rff_topk_f64):
 * vals/idx (P, S, k) -> out (S, k) under the ordering rule.  Entries with idx < 0 are ignored. */
int tes_topk_merge_f64(const double* vals, const int64_t* idx, int64_t P, int64_t S, int k, double* out_val,
                       int64_t* out_idx, void* stream);

/* Arg-max of the acquisition vector, FIRST maximum wins, the first NaN wins over everything (numpy / tf.argmax
 * semantics).  Replaces bo_discrete.optimize_criterion (bo_discrete.py:794-807) and the chunk merge at
 * :1053-1055.  out_idx = -1 when N == 0. */
int64_t tes_argmax_workspace_bytes(int64_t N);
int tes_argmax_f64(const double* vals, int64_t N, double* out_val, int64_t* out_idx, void* ws, int64_t ws_bytes,
                   void* stream);

/* ===================================================================================== Stage B
 * SE-ARD cross kernel k(x_i, xb_c), out (N, m).  Replaces utils.computeKnm (utils.py:703-719),
 * using the reference's |a|^2 + |b|^2 - 2ab expansion. */
int tes_se_ard_cross_f64(const double* x, int64_t N, const double* xb, int64_t m, int64_t d, const double* l,
                         double sigma, double* out, void* stream);

/* Plain fp64 GEMM C[MxN] = A[MxK] * op(B); op(B) = B (ldb = row stride) or B^T when transb != 0
 * (B stored N x K, ldb = its row stride).  The contraction primitive of stages B/D. */
int tes_gemm_f64(const double* A, int64_t lda, const double* B, int64_t ldb, int transb, double* C, int64_t ldc,
                 int64_t M, int64_t N, int64_t K, void* stream);

/* pNK = K([test_xs; X]) + sigma0 * diag(0_T, I_n), (T+n)^2.  Replaces the matrix assembly of
 * get_pNK_test_obs (criteria/evaluate_mp.py:12-53 and its four copies); invert it with
 * tes_batched_chol2inv_f64 (the reference uses tf.linalg.inv; the matrix is SPD). */
int64_t tes_pnk_workspace_bytes(int64_t T, int64_t n, int64_t d);
int tes_pnk_f64(const double* test_xs, int64_t T, const double* X, int64_t n, int64_t d, const double* l,
                double sigma, double sigma0, double* out_pNK, void* ws, int64_t ws_bytes, void* stream);

/* Common head of every TES criterion: with k = K(x, [test_xs; X]) (N, T+n),
 *   out_M (N,T) = k invpNK[:, :T];  out_b (N) = k invpNK[:, T:] Y;  out_Kq (N) = sigma - rowsum((k invpNK) o k)
 * Replaces evaluate_mp.get_queried_f_stat_given_data_maxidx:76-90 and
 * evaluate_sample_mp.get_queried_f_stat_given_test_samples:77-124 (tmp_test, query_mean_obs, query_var). */
int64_t tes_posterior_stats_workspace_bytes(int64_t N, int64_t T, int64_t n, int64_t d);
int tes_posterior_stats_f64(const double* x, int64_t N, int64_t d, const double* test_xs, int64_t T, const double* X,
                            int64_t n, const double* Y, const double* l, double sigma, const double* invpNK,
                            double* out_M, double* out_b, double* out_Kq, void* ws, int64_t ws_bytes, void* stream);

/* GP posterior mean (N) and marginal variance (N, clipped at 1e-100) given NKinv = (K_XX + sigma0 I)^-1.
 * Replaces utils.compute_mean_var_f with fullcov=False (utils.py:786-815). */
int64_t tes_gp_mean_var_workspace_bytes(int64_t N, int64_t n);
int tes_gp_mean_var_f64(const double* x, int64_t N, int64_t d, const double* X, int64_t n, const double* Y,
                        const double* l, double sigma, const double* NKinv, double* out_mean, double* out_var,
                        void* ws, int64_t ws_bytes, void* stream);

/* Batched SPD factorisation and inversion (all matrices m x m, contiguous (batch, m, m)).
 * tes_batched_chol2inv_f64 replaces utils.chol2inv (utils.py:657-674), utils.multichol2inv
 * (:677-700), utils.precomputeInvKs (:1838-1856), the per-maximizer (n+1)^2 batch of
 * utils.eval_invKmaxsams (:1859-1883) and the tf.linalg.inv of get_pNK_test_obs
 * (criteria/evaluate_mp.py:47; that matrix is SPD).  A^-1 = L^-T L^-1, L = chol(A), no jitter.
 * info[b] = 0, or the 1-based index of the first non-positive pivot (the output then holds NaN).
 * A is not modified.  `want_inverse` selects the workspace size for chol2inv (1) or potrf (0). */
int64_t tes_batched_chol_workspace_bytes(int64_t batch, int64_t m, int want_inverse);
int tes_batched_potrf_f64(const double* A, int64_t batch, int64_t m, double* L, int* info, void* ws,
                          int64_t ws_bytes, void* stream);
int tes_batched_chol2inv_f64(const double* A, int64_t batch, int64_t m, double* Ainv, int* info, void* ws,
                             int64_t ws_bytes, void* stream);

/* ===================================================================================== Stage D
 * TES_ep acquisition for one hyper-parameter sample (the reference loops over nhyp and averages).
 * Replaces criteria/evaluate_mp.py:116-295 (mode 0 deterministic, mode 1 stochastic) and
 * criteria/evaluate_mp_lite.py:116-225 (mode 2); mode 3 only returns the query moments
 * (get_queried_f_stat_given_data_maxidx, evaluate_mp.py:56-113).
 *   p (Sp), post_mean (Sp,T), post_cov (Sp,T,T): the maximizers with max_probs > 0 only
 *   invKobs (n,n): needed by mode 0 (H[y] of the observation-only posterior)
 *   mode 1 draws: z != NULL -> host-supplied standard normals, layout (niter, nysample, Sp, N);
 *                 z == NULL -> Philox (include/tes_spec.h) keyed by (seed, iteration) and the
 *                 logical element (iy*Sp + j)*n_global + (x_offset + x), so results do not depend
 *                 on how candidates are sharded; n_global < 0 -> SHARED draws: every candidate
 *                 consumes the same normals (element iy*Sp + j), i.e. common random numbers for
 *                 the finite-difference gradients of the criterion refinement
 *                 (utils_for_continuous.py:62-87);
 *   out_acq (N); out_mean / out_var (N,Sp), optional (NULL to skip): candidate-major moments.
 * NaN semantics follow the reference (no variance clipping before sqrt). */
int64_t tes_ep_acq_workspace_bytes(int64_t N, int64_t T, int64_t n, int64_t d, int64_t Sp);
int tes_ep_acq_f64(int mode, const double* x, int64_t N, int64_t d, const double* test_xs, int64_t T,
                   const double* X, int64_t n, const double* Y, const double* l, double sigma, double sigma0,
                   const double* invpNK, const double* invKobs, const double* p, int64_t Sp,
                   const double* post_mean, const double* post_cov, int niter, int nysample, const double* z,
                   uint64_t seed, int64_t n_global, int64_t x_offset, double* out_acq, double* out_mean,
                   double* out_var, void* ws, int64_t ws_bytes, void* stream);

/* Tensor-core SCREEN of the deterministic TES_ep criterion (criteria/evaluate_mp.py:199-219): out_acq (N) = the criterion
 * with the query variances' quadratic forms (evaluate_mp.py:103-107) computed as a 3xFP16 tcgen05 GEMM, out_eps (N) = a
 * proven bound on |out_acq - exact| (infinite where the variance could be non-positive or is NaN).  A caller that only
 * needs the arg-max keeps {x : acq~ + eps >= max(acq~ - eps)} and evaluates those with tes_ep_acq_f64: same arg-max,
 * same value (tes_b200.acquisition.tes_step does).  Sp <= 128. */
int64_t tes_ep_acq_screen_workspace_bytes(int64_t N, int64_t T, int64_t n, int64_t d, int64_t Sp);
int tes_ep_acq_screen_f64(const double* x, int64_t N, int64_t d, const double* test_xs, int64_t T, const double* X,
                          int64_t n, const double* Y, const double* l, double sigma, double sigma0,
                          const double* invpNK, const double* invKobs, const double* p, int64_t Sp,
                          const double* post_cov, double* out_acq, double* out_eps, void* ws, int64_t ws_bytes,
                          void* stream);

/* TES_sp acquisition, q = 1 (criteria/evaluate_sample_mp.py:132-336) or batch q <= 8
 * (criteria/evaluate_batch_sample_mp.py:160-382).  x (nx, q, d); post_samples (Sp,T,K);
 * post_mask (Sp,K) bytes (1 = valid); z layout (niter, nysample, Sp, nx, K, q) or NULL for Philox
 * (element ((iy*Sp + j)*n_global + x_offset + x)*K + s)*q + c; n_global < 0: shared draws, the
 * query index drops out of the element).  out_acq (nx).
 * On return the int32 at ws + tes_sp_acq_workspace_bytes(...) - 256 holds the number of queries whose q x q
 * predictive covariance had no Cholesky factor (their acquisition value is NaN; the reference's tf.cholesky raises
 * InvalidArgumentError there, evaluate_batch_sample_mp.py:306-310). */
int64_t tes_sp_acq_workspace_bytes(int64_t nx, int q, int64_t T, int64_t n, int64_t d, int64_t Sp, int64_t K);
int tes_sp_acq_f64(const double* x, int64_t nx, int q, int64_t d, const double* test_xs, int64_t T, const double* X,
                   int64_t n, const double* Y, const double* l, double sigma, double sigma0, const double* invpNK,
                   const double* p, int64_t Sp, const double* post_samples, const unsigned char* post_mask,
                   int64_t K, int niter, int nysample, const double* z, uint64_t seed, int64_t n_global,
                   int64_t x_offset, double* out_acq, void* ws, int64_t ws_bytes, void* stream);

/* Batch TES_ep (criteria/evaluate_batch_mp.py:145-331).  As shipped, the reference scores the
 * samples under the STANDARD normal and builds the marginal from that same log-prob (:302,
 * :317-322), so its value is -(sum p) log(sum p) ~ 0 for every finite input; this entry
 * point returns exactly that closed form, and NaN for a query x (nx, q, d) with a non-finite
 * coordinate (which the reference propagates through its kernel/eigh/log_prob chain, :88-141,
 * :236-241, :296-302 -- the drivers retry on NaN).  p (Sp): max_probs > 0. */
int tes_batch_ep_acq_f64(const double* p, int64_t Sp, const double* x, int64_t nx, int64_t q, int64_t d,
                         double* out_acq, void* stream);

/* ============================================================================ Stage A (draw)
 * optfunc.draw_random_init_weights_features_np (optfunc.py:303-385): the RFF posterior weight draw for S function
 * samples at once.  X (n, d) observations, Y (n) centred targets, l (d) inverse squared length-scales; the three
 * random draws are inputs in the reference's draw order: randn_W (S, F, d) (:329), rand_b (S, F) uniform (:334),
 * randn_noise (S, F) (:344).  Outputs theta (S, F), W (S, F, d) = randn_W * sqrt(l), b (S, F) = 2 pi rand_b.
 * n < F: eigen-decomposition route of :348-374 (batched GEMM for Z^T Z + sigma0 I, batched Jacobi eigen-solver);
 * n >= F: the F x F inverse + Cholesky route of :377-383.  1 <= d <= 64, S <= 65535. */
int64_t tes_rff_draw_workspace_bytes(int64_t S, int64_t F, int64_t n, int64_t d);
int tes_rff_draw_f64(const double* X, const double* Y, int64_t n, int64_t d, const double* l, double sigma,
                     double sigma0, const double* randn_W, const double* rand_b, const double* randn_noise,
                     int64_t S, int64_t F, double* theta, double* W, double* b, void* ws, int64_t ws_bytes,
                     void* stream);

/* Gradient refinement of the trusted maximizers: utils_for_continuous.sample_xmaxs_fmaxs, second half
 * (utils_for_continuous.py:201-232).  Every (function sample, start point) runs `ntrain` TF-1 Adam steps on
 * -f_j(x) with the clip-to-[xmin, xmax] variable constraint applied after each step, all inside one launch.
 *   x0 (S, k, d) start points; W (S or 1, F, d); b (S or 1, F); theta (S, F); xmin, xmax (d) device arrays;
 *   out_x (S, k, d) refined points; out_f (S, k) the function-sample values there (the caller takes the arg-max over
 *   k, first maximum, as :222-229 do).  TF-1 defaults: lr 1e-3, beta1 0.9, beta2 0.999, epsilon 1e-8.  d <= 16. */
int tes_rff_adam_refine_f64(const double* x0, int64_t S, int k, int64_t d, const double* W, const double* b,
                            const double* theta, int64_t F, int w_shared, double sigma, const double* xmin,
                            const double* xmax, int ntrain, double lr, double beta1, double beta2, double epsilon,
                            double* out_x, double* out_f, void* stream);

/* ===================================================================================== Stage C
 * Trusted-maximizer statistics on the device.  The reference does all of this in host numpy
 * (utils.get_testidxs_stats, utils.py:1733-1833); here it stays next to the data so that one
 * acquisition step has no host round trip of samples or (T,T,T) covariance stacks.
 */

/* Counter-based N(0,1) draws of include/tes_spec.h (Philox4x32-10 + Box-Muller), elements
 * first_elem .. first_elem+count-1 of the logical draw tensor (seed, iteration, stream_id).  Replaces the
 * np.random.normal call of utils.py:1674 when the caller passes a seed instead of the draws themselves. */
int tes_normal_fill_f64(uint64_t seed, uint32_t iteration, uint32_t stream_id, uint64_t first_elem,
                        int64_t count, double* out, void* stream);

/* Uniform draws on (0,1) from the same counters: u = (k + 0.5) 2^-53, k = top 53 bits of the Philox words that feed
 * the Box-Muller pair (element e -> pair e>>1, component e&1). */
int tes_uniform_fill_f64(uint64_t seed, uint32_t iteration, uint32_t stream_id, uint64_t first_elem,
                         int64_t count, double* out, void* stream);

/* utils.remove_duplicates_np / its mask (utils.py:1554-1581): walking the rows of xs (S, d) in order, every
 * not-yet-flagged row flags all LATER rows whose Euclidean distance is <= resolution.  The distance is evaluated in
 * numpy's operation order (pairwise np.sum of the squared differences, then sqrt), so the keep/drop decisions are
 * bit-identical.  mask (S) int32: 1 = duplicate; leader (S) int64: the row that flagged it (itself if kept).
 * S <= 49152, d <= 128. */
int tes_dedup_f64(const double* xs, int64_t S, int64_t d, double resolution, int* mask, int64_t* leader,
                  void* stream);

/* Batched symmetric eigen-decomposition A_b = V_b diag(e_b) V_b^T (one-sided Jacobi, one CTA per matrix; matrices
 * up to n = 160 stay in shared memory).  Replaces np.linalg.eig on the symmetric matrices of utils.py:1668
 * (colouring factor of the joint samples) and optfunc.py:355 (n x n feature Gram of the RFF posterior draw).
 * A (batch, n, n); evals (batch, n) in the solver's own (deterministic) order; evecs (batch, n, n) row-major with
 * the eigenvectors as COLUMNS (numpy's convention); sweeps (batch) int32 or NULL. */
int64_t tes_sym_eig_workspace_bytes(int64_t batch, int64_t n);
int tes_sym_eig_f64(const double* A, int64_t batch, int64_t n, double* evals, double* evecs, int* sweeps, void* ws,
                    int64_t ws_bytes, void* stream);

/* out = V diag(sqrt(clip(e, 0, inf)))  (utils.py:1668-1670), (n, n) row-major */
int tes_color_factor_f64(const double* evals, const double* evecs, int64_t n, double* out, void* stream);

/* Core of utils.sample_multivariate_normal_maxidx_np (utils.py:1675-1705) on joint samples (ns, T) that already
 * hold z A^T: adds mean(prior_mean) to every entry (reference quirk Q11; skipped when prior_mean is NULL), flags the
 * correlated dimensions (np.corrcoef > 1 - 1e-6 against an earlier dimension) when remove_correlated != 0, takes
 * the arg-max of every row over the unflagged dimensions (first maximum wins) and counts the buckets.
 * masked (T) int32, maxidx (ns) int64, counts (T) int32. */
int64_t tes_joint_maxidx_workspace_bytes(int64_t ns, int64_t T);
int tes_joint_maxidx_f64(double* samples, int64_t ns, int64_t T, const double* prior_mean, int remove_correlated,
                         int* masked, int64_t* maxidx, int* counts, void* ws, int64_t ws_bytes, void* stream);

/* empirical_approximation.get_empirical_stat_from_samples (empirical_approximation.py:86-107) for every bucket at
 * once: order (ns) = sample rows sorted stably by bucket, starts (nb+1) = bucket boundaries in `order`, cols (nk) =
 * kept test-point columns.  mean (nb, nk); cov (nb, nk, nk) with the (n_j - 1) denominator. */
int tes_bucket_moments_f64(const double* samples, int64_t ns, int64_t T, const int64_t* order, const int64_t* starts,
                           int64_t nb, const int64_t* cols, int64_t nk, double* mean, double* cov, void* stream);

/* Padded per-maximizer sample tensor of mode 'sample' (utils.py:1716-1727): out (nb, nk, K) holds the bucket's
 * samples in draw order and the PRIOR mean in the padded slots; mask (nb, K) bytes. */
int tes_bucket_samples_f64(const double* samples, int64_t ns, int64_t T, const int64_t* order, const int64_t* starts,
                           int64_t nb, const int64_t* cols, int64_t nk, int64_t K, const double* prior_mean,
                           double* out, uint8_t* mask, void* stream);

/* General inverse of `batch` n x n matrices by Gauss-Jordan elimination with partial pivoting -- what tf.linalg.inv /
 * np.linalg.inv (LU) compute; get_pNK_test_obs (criteria/evaluate_mp.py:47) falls back to it when the Cholesky route
 * reports a non-positive pivot (pNK carries no jitter on its test block).  out may not alias A. */
int tes_batched_gj_inverse_f64(const double* A, int64_t batch, int64_t n, double* out, void* stream);

/* ep.approximate_EP_np (ep.py:73-204) for EVERY maximizer of every hyper-parameter sample in one launch (one CTA per
 * maximizer, Gauss-Jordan inverse with partial pivoting in shared memory for T <= 160).  mean (nhyp, T), cov
 * (nhyp, T, T) -> post_mean (nhyp, T, T), post_cov (nhyp, T, T, T); row j = the EP approximation of N(mean, cov)
 * truncated to {f_j >= f_i for all i}.  niter (nhyp, T) int32 or NULL: the reference's iteration counter at exit.
 * Same update schedule, skip rule, NaN early-exit and convergence measure as the reference; where the reference's
 * skip loop would never terminate (every site invalid) the current approximation is returned. */
int64_t tes_ep_moments_workspace_bytes(int64_t nhyp, int64_t T);
int tes_ep_moments_f64(const double* mean, const double* cov, int64_t nhyp, int64_t T, double resolution,
                       int max_niter, double* post_mean, double* post_cov, int* niter, void* ws, int64_t ws_bytes,
                       void* stream);

/* ============================================================================== measurement aids
 * Sustained fp64 FMA rate of the device, used as the roofline denominator for the fp64-pipe-bound
 * kernels (MEASURED_PEAKS.json carries only HBM and bf16 numbers).  Runs `iters` dependent-chain
 * DFMA rounds on `blocks` CTAs; returns the flop count through *flops; time it with CUDA events on
 * `stream`.  `sink` is a device buffer of at least blocks*256 doubles. */
/* Test probe for the fp32 screen's error model: max over EVERY float x in [lo, hi] (both signs) of
 * |c(x) - cos(x)| - e1*|x|, written to *out_max (device double; 0 if never positive); c = __cosf (which = 0) or
 * the packed FMA-pipe polynomial cosine of the tensor-core screen (which = 1). */
int tes_cosf_error_probe(float lo, float hi, double e1, int which, double* out_max, void* stream);

/* Test probe for the tensor-core screen: max |h_mma - h_fp64| / A over 16 candidates x (8*n8) features
 * (x (16,d), w (8*n8,d), b (8*n8), all device fp64; d <= 10), written to *out_max (device double). */
int tes_mma_arg_error_probe(const double* x, const double* w, const double* b, int64_t d, int n8, double* out_max,
                            void* stream);

/* Same probe for the tcgen05/TMEM screen (method 4): 128 candidates x (64*nt) features through the kernel's own
 * operand construction, tcgen05.mma kind::tf32 and tcgen05.ld (x (128,d), w (64*nt,d), b (64*nt)). */
int tes_tc5_arg_error_probe(const double* x, const double* w, const double* b, int64_t d, int nt, double* out_max,
                            void* stream);

/* number of kernels this library has launched in this process so far (monotonic statistic) */
int64_t tes_launch_count(void);

int tes_fp64_fma_burn(int blocks, int64_t iters, double* sink, double* flops, void* stream);
/* Same for the special-function unit: __cosf (FMUL.RZ + MUFU.COS) evaluations per second; the denominator for the
 * tensor-core screen of stage A, which is MUFU-bound.  *ops receives the number of cosines. */
int tes_mufu_cos_burn(int blocks, int64_t iters, float* sink, double* ops, void* stream);
/* Same for the fp64 tensor-core path (mma.sync m8n8k4 f64): measured to decide whether DMMA is worth using. */
int tes_fp64_mma_burn(int blocks, int64_t iters, double* sink, double* flops, void* stream);
/* Same for the packed fp32 pipe (FFMA2, fma.rn.f32x2): the denominator for the fp32 screen of stage A.
 * `sink`: at least blocks*256 floats. */
int tes_fp32_fma_burn(int blocks, int64_t iters, float* sink, double* flops, void* stream);

#ifdef __cplusplus
}
#endif

#endif /* TES_B200_H */
