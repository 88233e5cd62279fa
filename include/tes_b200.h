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

/* All function-sample values, out (S, N).  Replaces utils_for_continuous.get_function_samples
 * (utils_for_continuous.py:257-296) / optfunc.make_function_sample_np.  Debug / small-N use. */
int tes_rff_eval_f64(const double* cand, int64_t N, int64_t d, const double* W, const double* b,
                     const double* theta, int64_t S, int64_t F, int w_shared, double sigma, double* out,
                     void* stream);

/* Merge P partial top-k lists (e.g. the NCCL all-gathered per-GPU results of tes_rff_topk_f64):
 * vals/idx (P, S, k) -> out (S, k) under the ordering rule.  Entries with idx < 0 are ignored. */
int tes_topk_merge_f64(const double* vals, const int64_t* idx, int64_t P, int64_t S, int k, double* out_val,
                       int64_t* out_idx, void* stream);

/* ============================================================================== measurement aids
 * Sustained fp64 FMA rate of the device, used as the roofline denominator for the fp64-pipe-bound
 * kernels (MEASURED_PEAKS.json carries only HBM and bf16 numbers).  Runs `iters` dependent-chain
 * DFMA rounds on `blocks` CTAs; returns the flop count through *flops; time it with CUDA events on
 * `stream`.  `sink` is a device buffer of at least blocks*256 doubles. */
int tes_fp64_fma_burn(int blocks, int64_t iters, double* sink, double* flops, void* stream);

#ifdef __cplusplus
}
#endif

#endif /* TES_B200_H */
