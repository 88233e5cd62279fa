/*
 * tes_spec.h -- arithmetic SPECIFICATION shared by the CUDA kernels (csrc/ *.cu) and the C
 * oracle (oracle/tes_oracle.c).  Plain C, no CUDA types.
 *
 * Why a spec: the reference evaluates f_j(x) = sqrt(2 sigma/F) theta_j^T cos(W_j x + b_j)
 * (utils_for_continuous.py:172-180, optfunc.py:389-402) through BLAS + libm, whose summation
 * order and cos() are implementation details no other program can reproduce bit for bit.  To
 * make "arg-max indices bit-exact" a testable statement, Stage A is defined by the operation
 * sequence below; the CUDA kernel and the C oracle both execute exactly this sequence with
 * IEEE-754 binary64 fma/add/mul (all correctly rounded, hence identical on CPU and GPU), and the
 * oracle is separately checked against the reference's numpy path (|diff| ~ 1e-15, same arg-max
 * on the golden fixtures).
 *
 *   prologue (once per weight set):   wp = W * TES_INV_PI ;  bp = b * TES_INV_PI   (half-turns)
 *   per (candidate x, sample j):      acc = 0
 *       for f = 0 .. F-1 (ascending): h = bp[f]; for k = 0..d-1 (ascending): h = fma(wp[f][k], x[k], h)
 *                                     acc = fma(theta[f], tes_cospi(h), acc)
 *                                     value = scale * acc,   scale = sqrt(2*sigma/F) (host, double)
 *   tes_cospi(h): n = rint(h) (nearest-even); r = h - n (exact); s = r*r;
 *                 p = Horner_fma(TES_COSPI_C[8..0], s);  return (n odd) ? -p : p
 *
 * tes_cospi approximates cos(pi*h): minimax degree-8 polynomial in s on [0, 1/4]
 * (approximation error 3.9e-18, evaluated error < 4e-16 absolute).
 * Ordering rule everywhere: larger value first, ties -> LOWER index.
 */
#ifndef TES_SPEC_H
#define TES_SPEC_H

#include <stdint.h>

#define TES_INV_PI 0.31830988618379067154 /* 0x3FD45F306DC9C883 */
#define TES_COSPI_NCOEF 9
#define TES_COSPI_C0 0x1.0000000000000p+0
#define TES_COSPI_C1 -0x1.3bd3cc9be45dbp+2
#define TES_COSPI_C2 0x1.03c1f081b5993p+2
#define TES_COSPI_C3 -0x1.55d3c7e3bfc44p+0
#define TES_COSPI_C4 0x1.e1f506813c9d9p-3
#define TES_COSPI_C5 -0x1.a6d1efc968c10p-6
#define TES_COSPI_C6 0x1.f9d25488fb381p-10
#define TES_COSPI_C7 -0x1.b69582d77d94dp-14
#define TES_COSPI_C8 0x1.16797132c8c4dp-18

/* ---- counter-based normal draws ("Randomness at scale") -------------------
 * Philox4x32-10 (Salmon et al. 2011), key = (seed_lo, seed_hi),
 * counter = (pair_lo, pair_hi, iteration, stream).  One call yields two N(0,1) via Box-Muller:
 *   u1 = ((x0 | x1<<32) >> 11) + 0.5) * 2^-53 ;  u2 likewise from (x2, x3)
 *   rad = sqrt(-2 ln u1) ;  z_even = rad*cos(2 pi u2) ;  z_odd = rad*sin(2 pi u2)
 * The logical element e of a draw tensor uses pair index e>>1 and component e&1.
 * log/sin/cos come from the platform libm (CUDA / glibc): draws agree to ~1 ulp, not bit-exact,
 * which the 1e-6 acquisition tolerance absorbs.
 */
#define TES_PHILOX_M0 0xD2511F53u
#define TES_PHILOX_M1 0xCD9E8D57u
#define TES_PHILOX_W0 0x9E3779B9u
#define TES_PHILOX_W1 0xBB67AE85u

/* streams (counter word 3) */
#define TES_STREAM_EP_Y 1u    /* evaluate_mp.py:265 draws      z[it][iy][j][x]            */
#define TES_STREAM_SP_Y 2u    /* evaluate_sample_mp.py:268     z[it][iy][j][x][s]         */
#define TES_STREAM_BSP_Y 3u   /* evaluate_batch_sample_mp.py:315  z[it][iy][j][x][s][q]   */
#define TES_STREAM_BEP_Y 4u   /* evaluate_batch_mp.py:294      z[it][iy][j][x][q]         */
#define TES_STREAM_JOINT_Z 5u /* utils.py:1674 joint samples  z[hyp = iteration][sample][t]   */
#define TES_STREAM_RFF_W 6u   /* optfunc.py:329  randn[j][f][k]   (iteration = hyper-parameter index) */
#define TES_STREAM_RFF_B 7u   /* optfunc.py:334  rand[j][f]  uniform: u = (k + 0.5) 2^-53, k = top 53 bits of the word pair */
#define TES_STREAM_RFF_E 8u   /* optfunc.py:344  randn[j][f] */

#endif /* TES_SPEC_H */
