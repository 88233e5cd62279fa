// rff.cu -- Stage A of the TES hot path on sm_100a: evaluate S random-Fourier-feature posterior
// function samples on N candidates and keep each sample's top-k (value desc, index asc).
//
// Replaces utils_for_continuous.py:172-187 (per-sample F x N cos + tf.math.top_k inside a
// sequential tf.while_loop) and optfunc.py:389-402.  Arithmetic is the fixed fp64 sequence of
// include/tes_spec.h, so indices are bit-identical to oracle/tes_oracle.c.
//
// Kernel shape (per-sample weights, the reference's semantics -- no GEMM exists here because
// every function sample owns its W_j, b_j; SURVEY section 7 "Per-sample features"):
//   * work unit = (candidate chunk, group of `sg` samples); grid = #units, the hardware block
//     scheduler load-balances them over the 148 SMs;
//   * each thread keeps C candidates' coordinates in registers; the group's feature records
//     [w*1/pi (d), b*1/pi, theta] stream through a 3-stage shared-memory ring filled by 1-D bulk
//     async copies (TMA engine, cp.async.bulk + mbarrier); every thread reads the same record
//     (shared-memory broadcast), so the loop is bound by the fp64 pipe: d + 13 DFMA-class
//     instructions per (candidate, sample, feature);
//   * at the end of a sample the CTA filters its C*256 values against the running k-th best
//     (one __syncthreads_or); only improving tiles take the slow path (append + warp selection).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "tes_common.cuh"

namespace tes {

constexpr int RFF_THREADS = 256;
constexpr int RFF_NSTAGE = 3;
constexpr int RFF_FCH = 128;  // features per pipeline stage
constexpr int RFF_MAX_K = 16;
constexpr int RFF_MAX_D_FAST = 10;

__host__ __device__ constexpr int rff_rec(int d) { return (d + 3) & ~1; }  // record doubles, even

// ---- prologue: interleave + prescale into feature records ---------------------------------------
__global__ void rff_pack_kernel(const double* __restrict__ W, const double* __restrict__ b,
                                const double* __restrict__ theta, int64_t S, int64_t F, int d, int w_shared,
                                int rec, double* __restrict__ feat) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= S * F) return;
    int64_t j = t / F, f = t % F;
    int64_t jw = w_shared ? 0 : j;
    const double* w = W + (jw * F + f) * d;
    double* o = feat + t * rec;
    for (int k = 0; k < d; ++k) o[k] = __dmul_rn(w[k], TES_INV_PI);
    o[d] = __dmul_rn(b[jw * F + f], TES_INV_PI);
    o[d + 1] = theta[t];
    for (int k = d + 2; k < rec; ++k) o[k] = 0.0;
}

// k rounds of "best entry strictly after the previous winner" over a list in shared memory.
// Executed by one full warp.  Entries with idx < 0 are invalid.  Duplicated (val, idx) pairs are
// naturally collapsed.  Results written to (out_val, out_idx)[0..k).
__device__ __forceinline__ void warp_select_topk(const double* lv, const int64_t* li, int n, int k, double* out_val,
                                                 int64_t* out_idx, int lane) {
    double pv = pos_inf();
    int64_t pi = -1;  // previous winner; (inf, -1) precedes everything
    for (int r = 0; r < k; ++r) {
        double bv = neg_inf();
        int64_t bi = INT64_MAX;
        for (int e = lane; e < n; e += 32) {
            double v = lv[e];
            int64_t i = li[e];
            if (i >= 0 && better(pv, pi, v, i) && better(v, i, bv, bi)) {
                bv = v;
                bi = i;
            }
        }
        warp_argmax(bv, bi);
        bool found = (bi != INT64_MAX);
        if (lane == 0) {
            out_val[r] = found ? bv : neg_inf();
            out_idx[r] = found ? bi : -1;
        }
        if (!found) {
            for (int q = r + 1; q < k; ++q)
                if (lane == 0) {
                    out_val[q] = neg_inf();
                    out_idx[q] = -1;
                }
            break;
        }
        pv = bv;
        pi = bi;
    }
}

template <int D, int C, int THREADS, int UNROLL>
__global__ void __launch_bounds__(THREADS)
    rff_topk_kernel(const double* __restrict__ cand, int64_t N, const double* __restrict__ feat, int S, int F,
                    double scale, int k, int sg, int64_t chunk, int64_t idx_offset, double* __restrict__ part_val,
                    int64_t* __restrict__ part_idx, const int* __restrict__ flags, int out_stride, int out_off) {
    constexpr int REC = rff_rec(D);
    constexpr int TILE = THREADS * C;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* stage = reinterpret_cast<double*>(smem_raw);                       // [NSTAGE][FCH*REC]
    uint64_t* full = reinterpret_cast<uint64_t*>(stage + RFF_NSTAGE * RFF_FCH * REC);  // [NSTAGE] (+pad to 8)
    double* tk_val = reinterpret_cast<double*>(full + 8);                      // [sg*k]
    int64_t* tk_idx = reinterpret_cast<int64_t*>(tk_val + sg * k);             // [sg*k]
    double* ap_val = reinterpret_cast<double*>(tk_idx + sg * k);               // [TILE + MAX_K]
    int64_t* ap_idx = reinterpret_cast<int64_t*>(ap_val + TILE + RFF_MAX_K);   // [TILE + MAX_K]
    int* ap_count = reinterpret_cast<int*>(ap_idx + TILE + RFF_MAX_K);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int nsg = (S + sg - 1) / sg;
    const int64_t chunk_id = blockIdx.x / nsg;
    const int g = blockIdx.x % nsg;
    const int j0 = g * sg;
    const int nj = min(sg, S - j0);
    const int64_t c0 = chunk_id * chunk;
    const int64_t c1 = min(N, c0 + chunk);
    const int ntiles = (int)((c1 - c0 + TILE - 1) / TILE);
    const int nch = (F + RFF_FCH - 1) / RFF_FCH;
    const int64_t total = (int64_t)ntiles * nj * nch;

    if (flags != nullptr && flags[chunk_id * S + j0] == 0) {
        // fallback launch (sg == 1) for a (chunk, sample) whose fp32 screen did not overflow: nothing to do
        for (int e = tid; e < nj * k; e += THREADS) {
            int64_t o = (chunk_id * S + j0 + e / k) * out_stride + out_off + e % k;
            part_val[o] = neg_inf();
            part_idx[o] = -1;
        }
        return;
    }

    for (int e = tid; e < sg * k; e += THREADS) {
        tk_val[e] = neg_inf();
        tk_idx[e] = -1;
    }
    if (tid == 0) {
        *ap_count = 0;
        for (int s = 0; s < RFF_NSTAGE; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    // producer: thread 0 issues the bulk copy of the next stream element into the ring.  The stream
    // (sample-major, then 128-feature chunk) repeats for every candidate tile; producer and consumer
    // walk it with incremental counters (no 64-bit div/mod in the loop).
    int p_jj = 0, p_c = 0, p_slot = 0;
    int64_t p_left = total;
    auto issue = [&]() {
        const int fcnt = min(RFF_FCH, F - p_c * RFF_FCH);
        const uint32_t bytes = (uint32_t)(fcnt * REC * sizeof(double));
        const double* src = feat + ((int64_t)(j0 + p_jj) * F + (int64_t)p_c * RFF_FCH) * REC;
        mbar_expect_tx(&full[p_slot], bytes);
        bulk_g2s(stage + p_slot * RFF_FCH * REC, src, bytes, &full[p_slot]);
        if (++p_c == nch) {
            p_c = 0;
            if (++p_jj == nj) p_jj = 0;
        }
        if (++p_slot == RFF_NSTAGE) p_slot = 0;
        --p_left;
    };
    if (tid == 0) {
        for (int p = 0; p < RFF_NSTAGE - 1 && p_left > 0; ++p) issue();
    }

    double x[C][D];
    double acc[C];
    int64_t ci[C];
    int slot = 0;
    uint32_t phase = 0;

    for (int tile = 0; tile < ntiles; ++tile) {
        {  // new candidate tile: load coordinates
            const int64_t t0 = c0 + (int64_t)tile * TILE;
#pragma unroll
            for (int cc = 0; cc < C; ++cc) {
                ci[cc] = t0 + cc * THREADS + tid;
                if (ci[cc] < c1) {
#pragma unroll
                    for (int kk = 0; kk < D; ++kk) x[cc][kk] = __ldg(cand + ci[cc] * D + kk);
                } else {
#pragma unroll
                    for (int kk = 0; kk < D; ++kk) x[cc][kk] = 0.0;
                }
            }
        }
        for (int jj = 0; jj < nj; ++jj) {
#pragma unroll
            for (int cc = 0; cc < C; ++cc) acc[cc] = 0.0;
            for (int c = 0; c < nch; ++c) {
                if (tid == 0 && p_left > 0) issue();
                mbar_wait(&full[slot], phase);
                const int fcnt = min(RFF_FCH, F - c * RFF_FCH);
                const double* sp = stage + slot * RFF_FCH * REC;
#pragma unroll UNROLL
                for (int f = 0; f < fcnt; ++f) {
                    double w[REC];
                    const double2* rp = reinterpret_cast<const double2*>(sp + f * REC);
#pragma unroll
                    for (int u = 0; u < REC / 2; ++u) {
                        double2 v = rp[u];
                        w[2 * u] = v.x;
                        w[2 * u + 1] = v.y;
                    }
#pragma unroll
                    for (int cc = 0; cc < C; ++cc) {
                        double h = w[D];
#pragma unroll
                        for (int kk = 0; kk < D; ++kk) h = fma(w[kk], x[cc][kk], h);
                        acc[cc] = fma(w[D + 1], cospi_spec(h), acc[cc]);
                    }
                }
                if (++slot == RFF_NSTAGE) {
                    slot = 0;
                    phase ^= 1u;
                }
                if (c != nch - 1) __syncthreads();  // ring slot may be refilled after this point
            }
            // sample jj finished for this tile
            const double thr = tk_val[jj * k + k - 1];
            double v[C];
            bool qual = false;
#pragma unroll
            for (int cc = 0; cc < C; ++cc) {
                v[cc] = (ci[cc] < c1) ? __dmul_rn(scale, acc[cc]) : neg_inf();
                qual |= (v[cc] > thr);
            }
            if (__syncthreads_or(qual)) {
#pragma unroll
                for (int cc = 0; cc < C; ++cc) {
                    if (v[cc] > thr) {
                        int pos = k + atomicAdd(ap_count, 1);
                        ap_val[pos] = v[cc];
                        ap_idx[pos] = ci[cc] + idx_offset;
                    }
                }
                if (tid < k) {
                    ap_val[tid] = tk_val[jj * k + tid];
                    ap_idx[tid] = tk_idx[jj * k + tid];
                }
                __syncthreads();
                if (tid < 32) {
                    warp_select_topk(ap_val, ap_idx, k + *ap_count, k, tk_val + jj * k, tk_idx + jj * k, lane);
                    __syncwarp();
                    if (tid == 0) *ap_count = 0;
                }
                __syncthreads();
            }
        }
    }

    for (int e = tid; e < nj * k; e += THREADS) {
        int jj = e / k, r = e % k;
        int64_t o = (chunk_id * S + j0 + jj) * out_stride + out_off + r;
        part_val[o] = tk_val[e];
        part_idx[o] = tk_idx[e];
    }
}

// generic-d fallback (d > RFF_MAX_D_FAST): one thread per (candidate, sample), records read through L1.
__global__ void rff_value_kernel_generic(const double* __restrict__ cand, int64_t N, int d,
                                         const double* __restrict__ feat, int S, int F, int rec, double scale,
                                         double* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int j = blockIdx.y;
    if (i >= N) return;
    const double* x = cand + i * d;
    const double* fj = feat + (int64_t)j * F * rec;
    double acc = 0.0;
    for (int f = 0; f < F; ++f) {
        const double* r = fj + (int64_t)f * rec;
        double h = __ldg(r + d);
        for (int kk = 0; kk < d; ++kk) h = fma(__ldg(r + kk), __ldg(x + kk), h);
        acc = fma(__ldg(r + d + 1), cospi_spec(h), acc);
    }
    out[(int64_t)j * N + i] = __dmul_rn(scale, acc);
}

// top-k of each row of a dense (S, N) value matrix -> one partial list (generic-d fallback path).
__global__ void rows_topk_kernel(const double* __restrict__ vals, int64_t N, int k, int64_t idx_offset,
                                 double* __restrict__ out_val, int64_t* __restrict__ out_idx) {
    const int j = blockIdx.x;
    const int lane = threadIdx.x;
    const double* row = vals + (int64_t)j * N;
    double pv = pos_inf();
    int64_t pi = -1;
    for (int r = 0; r < k; ++r) {
        double bv = neg_inf();
        int64_t bi = INT64_MAX;
        for (int64_t e = lane; e < N; e += 32) {
            double v = row[e];
            if (better(pv, pi, v, e) && better(v, e, bv, bi)) {
                bv = v;
                bi = e;
            }
        }
        warp_argmax(bv, bi);
        bool found = (bi != INT64_MAX);
        if (lane == 0) {
            out_val[(int64_t)j * k + r] = found ? bv : neg_inf();
            out_idx[(int64_t)j * k + r] = found ? bi + idx_offset : -1;
        }
        if (!found) {
            for (int q = r + 1; q < k; ++q)
                if (lane == 0) {
                    out_val[(int64_t)j * k + q] = neg_inf();
                    out_idx[(int64_t)j * k + q] = -1;
                }
            break;
        }
        pv = bv;
        pi = bi;
    }
}

// merge P partial lists per sample: vals/idx (P, S, k) -> (S, k).  One warp per sample.
__global__ void topk_merge_kernel(const double* __restrict__ vals, const int64_t* __restrict__ idx, int64_t P,
                                  int64_t S, int kin, int k, double* __restrict__ out_val,
                                  int64_t* __restrict__ out_idx) {
    const int64_t j = blockIdx.x;
    const int lane = threadIdx.x;
    const int64_t n = P * kin;
    double pv = pos_inf();
    int64_t pi = -1;
    for (int r = 0; r < k; ++r) {
        double bv = neg_inf();
        int64_t bi = INT64_MAX;
        for (int64_t e = lane; e < n; e += 32) {
            int64_t p = e / kin, q = e % kin;
            double v = vals[(p * S + j) * kin + q];
            int64_t i = idx[(p * S + j) * kin + q];
            if (i >= 0 && better(pv, pi, v, i) && better(v, i, bv, bi)) {
                bv = v;
                bi = i;
            }
        }
        warp_argmax(bv, bi);
        bool found = (bi != INT64_MAX);
        if (lane == 0) {
            out_val[j * k + r] = found ? bv : neg_inf();
            out_idx[j * k + r] = found ? bi : -1;
        }
        if (!found) {
            for (int q = r + 1; q < k; ++q)
                if (lane == 0) {
                    out_val[j * k + q] = neg_inf();
                    out_idx[j * k + q] = -1;
                }
            break;
        }
        pv = bv;
        pi = bi;
    }
}

// =====================================================================================================
// fp32 screen + fp64 refine.  The fp64 pipe (64 lanes/SM) is the bound of rff_topk_kernel; the packed fp32
// pipe (FFMA2, 2 x 128 lanes/SM) plus the MUFU cos is ~4.7x faster.  The screen evaluates every
// (candidate, sample) pair in fp32 with a PROVEN error bound B_j per sample and keeps every candidate whose
// screen value is within 2 B_j of (a lower bound of) the k-th best screen value; only those are re-evaluated
// with the exact fp64 sequence of tes_spec.h, which alone decides values and indices.  If u_i are screen
// values with |u_i - v_i| <= B and t is the k-th largest u, every member of the true top-k satisfies
// u_i >= t - 2B, so the kept set contains the true top-k (ties included) and the result is bit-identical to
// the pure fp64 kernel.  A (chunk, sample) whose kept set overflows its CAP slots is recomputed by the fp64
// kernel (flag-gated launch), so correctness never depends on the list capacity.
//
// Error bound per sample j (u = 2^-24; A_f = |b_f| + sum_k |w_fk| max_i |x_ik| in radians; the screen works in
// radians, h = b + w.x):
//   |h32 - h| <= (d+4) u A_f                      input rounding of w, b, x and the d FFMA roundings
//   direct path (A_f <= 512 for every f): c = __cosf(h32), |c - cos(h32)| <= 3.0e-7 + 1.7e-7 A_f
//        (exhaustively measured model of FMUL.RZ(x, 1/2pi) + MUFU.COS, see SCR_E0/SCR_E1 below)
//   reduced path: n = rint(h/2pi), r = fma(n, -2pi_f32, h), c = __cosf(r):
//        |r - (h - 2pi n)| <= 1.75e-7 |n| + 2e-7,  |n| <= A_f/2pi + 1,  then the model on |r| <= pi
//   |acc - sum| <= 34 u sum_f |theta_f|           theta rounding, products, 32-term fp32 blocks flushed to fp64
//   B_j = 1.25 * scale * sum_f |theta_f| * (per-feature bound above)
// =====================================================================================================
constexpr int SCR_THREADS = 256;
constexpr int SCR_TILE = SCR_THREADS * 4;  // 4 candidates (2 packed pairs) per thread
constexpr int SCR_CAP = 32;                // kept candidates per (chunk, sample) before the fp64 fallback
constexpr int SCR_MAX_K = 8;               // k-th of the 8 warp maxima is the per-tile lower bound

__host__ __device__ constexpr int scr_recp(int d) { return (d + 3) & ~1; }  // 8-byte (v,v) pairs per record

__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) {
    return ((uint64_t)__float_as_uint(hi) << 32) | (uint64_t)__float_as_uint(lo);
}
__device__ __forceinline__ float f2_lo(uint64_t v) { return __uint_as_float((uint32_t)v); }
__device__ __forceinline__ float f2_hi(uint64_t v) { return __uint_as_float((uint32_t)(v >> 32)); }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t fmul2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

__global__ void rff_pack32_kernel(const double* __restrict__ W, const double* __restrict__ b,
                                  const double* __restrict__ theta, int64_t S, int64_t F, int d, int w_shared,
                                  int recp, uint64_t* __restrict__ feat32) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= S * F) return;
    int64_t j = t / F, f = t % F;
    int64_t jw = w_shared ? 0 : j;
    const double* w = W + (jw * F + f) * d;
    uint64_t* o = feat32 + t * recp;
    // the fp32 screen works in RADIANS (no 1/pi prescale): the 2*pi reduction folds into one FFMA2
    for (int k = 0; k < d; ++k) {
        float v = __double2float_rn(w[k]);
        o[k] = f2_pack(v, v);
    }
    float bv = __double2float_rn(b[jw * F + f]);
    o[d] = f2_pack(bv, bv);
    float tv = __double2float_rn(theta[t]);
    o[d + 1] = f2_pack(tv, tv);
    for (int k = d + 2; k < recp; ++k) o[k] = 0ull;
}

// xmax[k] = max_i |x_ik| (bit pattern of a non-negative double orders like an unsigned integer).
// Each thread walks whole rows (stride = grid size), keeps d running maxima in registers via shared memory
// per-dimension atomics, then one global atomic per dimension per block.
__global__ void rff_xmax_kernel(const double* __restrict__ cand, int64_t N, int d, unsigned long long* xmax) {
    __shared__ unsigned long long sm[64];
    if (threadIdx.x < 64) sm[threadIdx.x] = 0ull;
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int kk = 0; kk < d; ++kk) {
        double m = 0.0;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride)
            m = fmax(m, fabs(cand[i * d + kk]));
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, off));
        if ((threadIdx.x & 31) == 0) atomicMax(&sm[kk], (unsigned long long)__double_as_longlong(m));
    }
    __syncthreads();
    if (threadIdx.x < d) atomicMax(xmax + threadIdx.x, sm[threadIdx.x]);
}

// |__cosf(x) - cos(x)| <= SCR_E0 + SCR_E1 |x| for |x| <= SCR_XMAX: __cosf is FMUL.RZ(x, float(1/2pi)) -- up to
// 2^-23 relative rounding plus 4e-8 relative error of the constant, i.e. <= 1.6e-7 |x| radians -- followed by
// MUFU.COS on revolutions.  Measured EXHAUSTIVELY over every float in the domain on B200
// (tes_cosf_error_probe, tests/test_rff_gpu.py): max(err - 1.7e-7 |x|) = 1.50e-7 up to |x| = 4096, and
// max err = 4.04e-7 on [-pi, pi] (the CUDA guide's 2^-21.19).  E0 carries a 2x margin on the residual.
// Samples whose arguments can exceed SCR_XMAX use the explicit 2*pi reduction instead.
constexpr double SCR_E0 = 3.0e-7;
constexpr double SCR_E1 = 1.7e-7;
constexpr double SCR_XMAX = 512.0;

// one warp per sample: B_j from the fp64 feature records [w'(d), b', theta]; mode[j] = 1 when the explicit
// range reduction is needed (arguments may leave the verified domain of __cosf)
__global__ void rff_bound_kernel(const double* __restrict__ feat, int rec, int64_t S, int64_t F, int d,
                                 const unsigned long long* __restrict__ xmax, double scale, double harg,
                                 double* __restrict__ bound, int* __restrict__ mode) {
    const int64_t j = blockIdx.x;
    const int lane = threadIdx.x;
    const double u = 5.9604644775390625e-08;  // 2^-24
    // harg: |h_screen - h| <= harg * A_f  ((d+4) u for the FFMA2 screen, MMA_HARG for the 3xTF32 screen)
    double e_red = 0.0, e_dir = 0.0, amax = 0.0;
    for (int64_t f = lane; f < F; f += 32) {
        const double* r = feat + (j * F + f) * rec;
        double A = fabs(r[d]);  // records hold w/pi, b/pi: A is in half-turns, pi*A in radians
        for (int k = 0; k < d; ++k) A += fabs(r[k]) * __longlong_as_double((long long)xmax[k]);
        const double Arad = 3.14159265358979323846 * A * (1.0 + 1e-6);
        const double nmax = 0.5 * A + 1.0;  // |rint(h / 2pi)| <= A/2 + 1
        const double th = fabs(r[d + 1]);
        // reduced path: argument rounding + (2pi_f32 - 2pi) n + rounding of r + __cosf on [-pi,pi] + accumulation
        e_red += th * (harg * Arad + 1.75e-7 * nmax + 2.0e-7 + (SCR_E0 + SCR_E1 * 3.1416) + 34.0 * u);
        // direct path: argument rounding + verified __cosf model on |x| <= Arad + accumulation
        e_dir += th * (harg * Arad + SCR_E0 + SCR_E1 * Arad + 34.0 * u);
        amax = fmax(amax, Arad);
    }
    e_red = warp_sum(e_red);
    e_dir = warp_sum(e_dir);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, off));
    if (lane == 0) {
        const int m = (amax <= SCR_XMAX) ? 0 : 1;
        mode[j] = m;
        bound[j] = 1.25 * scale * (m ? e_red : e_dir);
    }
}

// test probe: max over every float x in [lo, hi] (both signs) of |__cosf(x) - cos(x)| - e1 |x|
__device__ __forceinline__ uint64_t cos_poly2(uint64_t h);
__global__ void cosf_probe_kernel(unsigned lo_bits, unsigned hi_bits, double e1, int which, double* out) {
    double m = -1.0;
    const unsigned stride = gridDim.x * blockDim.x;
    for (unsigned long long bits = (unsigned long long)lo_bits + blockIdx.x * blockDim.x + threadIdx.x;
         bits <= hi_bits; bits += stride) {
        float x = __uint_as_float((unsigned)bits);
        double ref = cos((double)x);
        float cp = __cosf(x), cm = __cosf(-x);
        if (which == 1) {
            const uint64_t pr = cos_poly2(f2_pack(x, -x));
            cp = f2_lo(pr);
            cm = f2_hi(pr);
        }
        double e = fmax(fabs((double)cp - ref), fabs((double)cm - ref)) - e1 * fabs((double)x);
        m = fmax(m, e);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, off));
    if ((threadIdx.x & 31) == 0 && m > 0.0)
        atomicMax((unsigned long long*)out, (unsigned long long)__double_as_longlong(m));
}

// shared-memory / global state of the screen's per-sample bookkeeping, common to both screen kernels
struct ScreenCtx {
    double* tk_val;
    int64_t* tk_idx;
    double* ap_val;
    int64_t* ap_idx;
    double* wm;
    int* slot_cnt;
    int* ap_count;
    const double* bound;
    double* slot_val;
    int64_t* slot_idx;
    int* flags;
    int k, j0, S, slot_stride;
    int64_t chunk_id, c1, idx_offset;
};

// One full warp folds the n appended entries ap[k .. k+n) (ap[0..k) = the old running list) of sample jj into the
// running screen top-k and the kept-candidate slots of the (chunk, sample); resets the append counter.
__device__ __forceinline__ void screen_fold_warp(const ScreenCtx& cx, int jj, int lane) {
    const int k = cx.k;
    const int n = *cx.ap_count;
    int cnt = cx.slot_cnt[jj];
    const int64_t fo = cx.chunk_id * cx.S + cx.j0 + jj;
    double* sv = cx.slot_val + fo * cx.slot_stride;
    int64_t* si = cx.slot_idx + fo * cx.slot_stride;
    // 1. fold the new entries into the running top-k of screen values (ap[0..k) holds the old list)
    warp_select_topk(cx.ap_val, cx.ap_idx, k + n, k, cx.tk_val + jj * k, cx.tk_idx + jj * k, lane);
    __syncwarp();
    // the running k-th best only grows, so anything below (new k-th best - 2B) can never be needed
    const double thr2 = cx.tk_val[jj * k + k - 1] - 2.0 * cx.bound[cx.j0 + jj];
    if (cnt > SCR_CAP) {
        // already flagged: nothing to keep
    } else if (cnt + n > SCR_CAP) {  // 2. compact the kept list in place before giving up on it
        int w = 0;
        for (int base = 0; base < cnt; base += 32) {
            const int e = base + lane;
            const bool in = e < cnt;
            const double v = in ? sv[e] : 0.0;
            const int64_t i = in ? si[e] : -1;
            const bool keep = in && v >= thr2;
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            __syncwarp();
            if (keep) {
                const int pos = w + __popc(m & ((1u << lane) - 1u));
                sv[pos] = v;
                si[pos] = i;
            }
            w += __popc(m);
            __syncwarp();
        }
        for (int e = w + lane; e < cnt; e += 32) {
            si[e] = -1;
            sv[e] = neg_inf();
        }
        cnt = w;
        __syncwarp();
    }
    // 3. append the new entries that survive thr2
    bool overflow = cnt > SCR_CAP;
    for (int base = 0; base < n && !overflow; base += 32) {
        const int e = base + lane;
        const bool keep = e < n && cx.ap_val[k + e] >= thr2;
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        const int tot = __popc(m);
        if (cnt + tot > SCR_CAP) {
            overflow = true;
            break;
        }
        if (keep) {
            const int pos = cnt + __popc(m & ((1u << lane) - 1u));
            sv[pos] = cx.ap_val[k + e];
            si[pos] = cx.ap_idx[k + e];
        }
        cnt += tot;
    }
    if (overflow) {
        if (lane == 0) cx.flags[fo] = 1;  // the fp64 kernel recomputes this (chunk, sample)
        cnt = SCR_CAP + 1;                // stays flagged
    }
    __syncwarp();
    if (lane == 0) {
        cx.slot_cnt[jj] = cnt;
        *cx.ap_count = 0;
    }
}

// A tile's NC-per-thread screen values u (=-inf for invalid candidates) of sample jj: keep every candidate
// within 2 B_j of max(k-th largest warp maximum of this tile, running k-th best of the chunk).
// barrier policy: the whole CTA (barrier 0)
struct SyncCta {
    static __device__ __forceinline__ int tid() { return threadIdx.x; }
    static __device__ __forceinline__ void sync() { __syncthreads(); }
    static __device__ __forceinline__ bool sync_or(bool p) { return __syncthreads_or(p); }
};

template <int NC, class Sync = SyncCta, int NW = 8>
__device__ __forceinline__ void screen_commit(const ScreenCtx& cx, int jj, const double (&u)[NC],
                                              const int64_t (&ci)[NC]) {
    const int tid = Sync::tid(), lane = tid & 31, warp = tid >> 5;
    const int k = cx.k;
    double tmax = neg_inf();
#pragma unroll
    for (int cc = 0; cc < NC; ++cc) tmax = fmax(tmax, u[cc]);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) tmax = fmax(tmax, __shfl_xor_sync(0xffffffffu, tmax, off));
    if (lane == 0) cx.wm[warp] = tmax;
    Sync::sync();
    // k-th largest (with multiplicity) of the NW warp maxima: a lower bound of the tile's k-th largest value
    double lk;
    {
        double a[NW];
#pragma unroll
        for (int q = 0; q < NW; ++q) a[q] = cx.wm[q];
        lk = neg_inf();
        double prev = pos_inf();
        int taken = 0;
        for (int r = 0; r < k; ++r) {
            double best = neg_inf();
            int cnt = 0;
#pragma unroll
            for (int q = 0; q < NW; ++q) {
                if (a[q] < prev) {
                    if (a[q] > best) {
                        best = a[q];
                        cnt = 1;
                    } else if (a[q] == best) {
                        ++cnt;
                    }
                }
            }
            if (cnt == 0) break;
            taken += cnt;
            lk = best;
            prev = best;
            if (taken >= k) break;
        }
        if (taken < k) lk = neg_inf();
    }
    const double thr = fmax(lk, cx.tk_val[jj * k + k - 1]) - 2.0 * cx.bound[cx.j0 + jj];
    bool qual = false;
#pragma unroll
    for (int cc = 0; cc < NC; ++cc) qual |= (ci[cc] < cx.c1) && (u[cc] >= thr);
    if (Sync::sync_or(qual)) {
#pragma unroll
        for (int cc = 0; cc < NC; ++cc) {
            if ((ci[cc] < cx.c1) && (u[cc] >= thr)) {
                int pos = k + atomicAdd(cx.ap_count, 1);
                cx.ap_val[pos] = u[cc];
                cx.ap_idx[pos] = ci[cc] + cx.idx_offset;
            }
        }
        if (tid < k) {
            cx.ap_val[tid] = cx.tk_val[jj * k + tid];
            cx.ap_idx[tid] = cx.tk_idx[jj * k + tid];
        }
        Sync::sync();
        if (tid < 32) screen_fold_warp(cx, jj, lane);
        Sync::sync();
    }
}

template <int D, int UNROLL, int MINB>
__global__ void __launch_bounds__(SCR_THREADS, MINB)
    rff_screen_kernel(const double* __restrict__ cand, int64_t N, const uint64_t* __restrict__ feat32, int S, int F,
                      double scale, int k, int sg, int64_t chunk, int64_t idx_offset,
                      const double* __restrict__ bound, const int* __restrict__ mode,
                      double* __restrict__ slot_val, int64_t* __restrict__ slot_idx, int slot_stride,
                      int* __restrict__ flags) {
    constexpr int RECP = scr_recp(D);
    constexpr int THREADS = SCR_THREADS;
    constexpr int TILE = SCR_TILE;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* stage = reinterpret_cast<uint64_t*>(smem_raw);                            // [NSTAGE][FCH*RECP]
    uint64_t* full = stage + RFF_NSTAGE * RFF_FCH * RECP;                               // [8]
    double* tk_val = reinterpret_cast<double*>(full + 8);                               // [sg*k] running screen top-k
    int64_t* tk_idx = reinterpret_cast<int64_t*>(tk_val + sg * k);                      // [sg*k]
    double* ap_val = reinterpret_cast<double*>(tk_idx + sg * k);                        // [TILE + 16]
    int64_t* ap_idx = reinterpret_cast<int64_t*>(ap_val + TILE + 16);                   // [TILE + 16]
    double* wm = reinterpret_cast<double*>(ap_idx + TILE + 16);                         // [8] warp maxima
    int* slot_cnt = reinterpret_cast<int*>(wm + 8);                                     // [sg]
    int* ap_count = slot_cnt + sg;

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int nsg = (S + sg - 1) / sg;
    const int64_t chunk_id = blockIdx.x / nsg;
    const int g = blockIdx.x % nsg;
    const int j0 = g * sg;
    const int nj = min(sg, S - j0);
    const int64_t c0 = chunk_id * chunk;
    const int64_t c1 = min(N, c0 + chunk);
    const int ntiles = (int)((c1 - c0 + TILE - 1) / TILE);
    const int nch = (F + RFF_FCH - 1) / RFF_FCH;
    const int64_t total = (int64_t)ntiles * nj * nch;

    for (int e = tid; e < sg * k; e += THREADS) {
        tk_val[e] = neg_inf();
        tk_idx[e] = -1;
    }
    for (int e = tid; e < nj * SCR_CAP; e += THREADS) {
        int64_t o = (chunk_id * S + j0 + e / SCR_CAP) * slot_stride + e % SCR_CAP;
        slot_idx[o] = -1;
        slot_val[o] = neg_inf();
    }
    for (int e = tid; e < nj; e += THREADS) {
        slot_cnt[e] = 0;
        flags[chunk_id * S + j0 + e] = 0;
    }
    if (tid == 0) {
        *ap_count = 0;
        for (int s = 0; s < RFF_NSTAGE; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    const ScreenCtx cx{tk_val, tk_idx, ap_val, ap_idx, wm, slot_cnt, ap_count, bound, slot_val, slot_idx, flags,
                       k, j0, S, slot_stride, chunk_id, c1, idx_offset};
    int p_jj = 0, p_c = 0, p_slot = 0;
    int64_t p_left = total;
    auto issue = [&]() {
        const int fcnt = min(RFF_FCH, F - p_c * RFF_FCH);
        const uint32_t bytes = (uint32_t)(fcnt * RECP * sizeof(uint64_t));
        const uint64_t* src = feat32 + ((int64_t)(j0 + p_jj) * F + (int64_t)p_c * RFF_FCH) * RECP;
        mbar_expect_tx(&full[p_slot], bytes);
        bulk_g2s(stage + p_slot * RFF_FCH * RECP, src, bytes, &full[p_slot]);
        if (++p_c == nch) {
            p_c = 0;
            if (++p_jj == nj) p_jj = 0;
        }
        if (++p_slot == RFF_NSTAGE) p_slot = 0;
        --p_left;
    };
    if (tid == 0) {
        for (int p = 0; p < RFF_NSTAGE - 1 && p_left > 0; ++p) issue();
    }

    const uint64_t INV2PI2 = f2_pack(0.15915494f, 0.15915494f);
    const uint64_t MAGIC2 = f2_pack(12582912.0f, 12582912.0f);  // 1.5 * 2^23
    const uint64_t NMAGIC2 = f2_pack(-12582912.0f, -12582912.0f);
    const uint64_t NTWOPI2 = f2_pack(-6.2831855f, -6.2831855f);

    uint64_t x2[2][D];
    double dacc[4];
    int64_t ci[4];
    int slot = 0;
    uint32_t phase = 0;

    for (int tile = 0; tile < ntiles; ++tile) {
        {
            const int64_t t0 = c0 + (int64_t)tile * TILE;
            float xf[4][D];
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                ci[cc] = t0 + cc * THREADS + tid;
#pragma unroll
                for (int kk = 0; kk < D; ++kk)
                    xf[cc][kk] = (ci[cc] < c1) ? __double2float_rn(__ldg(cand + ci[cc] * D + kk)) : 0.0f;
            }
#pragma unroll
            for (int kk = 0; kk < D; ++kk) {
                x2[0][kk] = f2_pack(xf[0][kk], xf[1][kk]);
                x2[1][kk] = f2_pack(xf[2][kk], xf[3][kk]);
            }
        }
        for (int jj = 0; jj < nj; ++jj) {
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) dacc[cc] = 0.0;
            const int red = mode[j0 + jj];
            for (int c = 0; c < nch; ++c) {
                if (tid == 0 && p_left > 0) issue();
                mbar_wait(&full[slot], phase);
                const int fcnt = min(RFF_FCH, F - c * RFF_FCH);
                const uint64_t* sp = stage + slot * RFF_FCH * RECP;
                auto run_chunk = [&](auto reduce_tag) {
                    constexpr bool REDUCE = decltype(reduce_tag)::value;
                    for (int f0 = 0; f0 < fcnt; f0 += 32) {
                        const int fe = min(fcnt, f0 + 32);
                        uint64_t acc2[2] = {0ull, 0ull};
#pragma unroll UNROLL
                        for (int f = f0; f < fe; ++f) {
                            uint64_t w[RECP];
                            const ulonglong2* rp = reinterpret_cast<const ulonglong2*>(sp + f * RECP);
#pragma unroll
                            for (int q = 0; q < RECP / 2; ++q) {
                                ulonglong2 v = rp[q];
                                w[2 * q] = v.x;
                                w[2 * q + 1] = v.y;
                            }
#pragma unroll
                            for (int pp = 0; pp < 2; ++pp) {
                                uint64_t h = w[D];
#pragma unroll
                                for (int kk = 0; kk < D; ++kk) h = ffma2(w[kk], x2[pp][kk], h);
                                uint64_t xr = h;
                                if (REDUCE) {  // r = h - 2pi*rint(h/2pi) in [-pi,pi] (one FFMA2 after the rint)
                                    uint64_t t = ffma2(h, INV2PI2, MAGIC2);
                                    uint64_t n = fadd2(t, NMAGIC2);
                                    xr = ffma2(n, NTWOPI2, h);
                                }
                                uint64_t cs = f2_pack(__cosf(f2_lo(xr)), __cosf(f2_hi(xr)));
                                acc2[pp] = ffma2(w[D + 1], cs, acc2[pp]);
                            }
                        }
                        dacc[0] += (double)f2_lo(acc2[0]);
                        dacc[1] += (double)f2_hi(acc2[0]);
                        dacc[2] += (double)f2_lo(acc2[1]);
                        dacc[3] += (double)f2_hi(acc2[1]);
                    }
                };
                if (red)
                    run_chunk(std::true_type{});
                else
                    run_chunk(std::false_type{});
                if (++slot == RFF_NSTAGE) {
                    slot = 0;
                    phase ^= 1u;
                }
                __syncthreads();  // ring slot may be refilled after this point
            }
            // ---- sample jj finished for this tile: keep everything within 2B of the k-th best screen value
            double u[4];
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) u[cc] = (ci[cc] < c1) ? scale * dacc[cc] : neg_inf();
            screen_commit<4>(cx, jj, u, ci);
        }
    }
}

// =====================================================================================================
// Tensor-core variant of the screen: the d-term inner products h = b + w.x run on the warp-level
// TF32 tensor-core MMA (mma.sync m16n8k8) in split precision ("3xTF32": x = x_hi + x_lo, w = w_hi + w_lo,
// h ~ x_hi w_hi + x_hi w_lo + x_lo w_hi, bias as two extra K columns against a constant 1), which removes
// the dot products from the fp32 pipe; what is left per (candidate, sample, feature) is the MUFU cosine,
// the FMUL of __cosf and half a packed FFMA2, so the kernel is bound by the MUFU/XU pipe (16 lanes/SM/clk).
//   * K layout (Kp = 8*KS, KS = ceil((3d+2)/8)): [x_hi | x_hi | x_lo | 1 1] against [w_hi | w_lo | w_hi | b_hi b_lo]
//   * a warp owns 64 candidates = 4 row tiles of 16; its A fragments (KS*4 registers per row tile) are built
//     once per candidate tile; B fragments + the two theta values a thread needs arrive pre-arranged in
//     fragment order from the pack kernel (one 16/32/48-byte record per lane per 8 features), streamed through
//     the same bulk-copy ring;
//   * accumulator fragment (c0..c3) = h for rows (g, g+8) x features (2t, 2t+1): cos, packed FFMA2 with
//     (theta_2t, theta_2t+1) into per-row-tile fp32 accumulators, flushed to fp64 every 16 feature tiles;
//     at the end of a sample the 4 lanes of a quad are shuffle-reduced.
// Error model: |h_mma - h| <= 6e-6 A_f (3.5 * 2^-22 from the dropped lo*lo / tf32 residuals / fp32 input
// rounding, 2^-18 for <= 32 truncating fp32 accumulations), verified by tes_mma_arg_error_probe.
// =====================================================================================================
// packed fp32 cosine on the FMA pipe: n = rint(h/pi), r = fma(n, -pi_f32, h) in [-pi/2, pi/2] (+ slack), degree-4
// minimax polynomial in r^2 (approximation error 4.7e-8 on 1.001*pi/2), sign from the parity of n.  It shares
// the (SCR_E0, SCR_E1) error model of __cosf, verified exhaustively by the same probe (which = 1).
__device__ __forceinline__ uint64_t cos_poly2(uint64_t h) {
    const uint64_t INVPI2 = f2_pack(0.31830987f, 0.31830987f);
    const uint64_t MAGIC2 = f2_pack(12582912.0f, 12582912.0f);
    const uint64_t NMAGIC2 = f2_pack(-12582912.0f, -12582912.0f);
    const uint64_t NPI2 = f2_pack(-3.14159274f, -3.14159274f);
    const uint64_t t = ffma2(h, INVPI2, MAGIC2);
    const uint64_t n = fadd2(t, NMAGIC2);
    const uint64_t r = ffma2(n, NPI2, h);
    const uint64_t s2 = fmul2(r, r);
    uint64_t p = ffma2(f2_pack(0x1.8467a8p-16f, 0x1.8467a8p-16f), s2, f2_pack(-0x1.6b29b6p-10f, -0x1.6b29b6p-10f));
    p = ffma2(p, s2, f2_pack(0x1.554ed4p-5f, 0x1.554ed4p-5f));
    p = ffma2(p, s2, f2_pack(-0x1.ffffcp-2f, -0x1.ffffcp-2f));
    p = ffma2(p, s2, f2_pack(0x1.fffffep-1f, 0x1.fffffep-1f));
    const uint32_t lo = (uint32_t)p ^ ((uint32_t)t << 31);
    const uint32_t hi = (uint32_t)(p >> 32) ^ ((uint32_t)(t >> 32) << 31);
    return ((uint64_t)hi << 32) | lo;
}

template <int D>
struct MmaCfg {
    static constexpr int KS = (3 * D + 2 + 7) / 8;
    static constexpr int RL = ((2 * KS + 2) + 3) & ~3;  // floats per lane record
};
constexpr int MMA_RT = 4;                  // row tiles (of 16 candidates) per warp
constexpr int MMA_NPOLY = 0;               // accumulator pairs (of 8 per feature tile) whose cosine runs on the FMA pipe:
                                           // measured slower for every value > 0 here (mma.sync already fills the issue
                                           // slots); the tcgen05 kernel, whose MMA is off the issue path, uses the mix
constexpr int MMA_TILE = 8 * MMA_RT * 16;  // candidates per CTA tile
constexpr double MMA_HARG = 6.0e-6;

__device__ __forceinline__ uint32_t to_tf32(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// value of the B operand (feature side) at K index k for one feature: w, b in radians
template <int D>
__device__ __forceinline__ float mma_bval(const double* __restrict__ w, double b, int k) {
    if (k < 2 * D || (k >= 2 * D && k < 3 * D)) {
        const int kk = k < D ? k : (k < 2 * D ? k - D : k - 2 * D);
        const float wf = __double2float_rn(w[kk]);
        const float hi = __uint_as_float(to_tf32(wf));
        if (k >= D && k < 2 * D) return __uint_as_float(to_tf32(wf - hi));
        return hi;
    }
    if (k == 3 * D || k == 3 * D + 1) {
        const float bf = __double2float_rn(b);
        const float hi = __uint_as_float(to_tf32(bf));
        return k == 3 * D ? hi : __uint_as_float(to_tf32(bf - hi));
    }
    return 0.0f;
}

// fragment-ordered feature records: feat[(j*nt8 + tile)*32 + lane][RL]
template <int D>
__global__ void rff_pack_mma_kernel(const double* __restrict__ W, const double* __restrict__ b,
                                    const double* __restrict__ theta, int64_t S, int64_t F, int w_shared,
                                    float* __restrict__ featm) {
    constexpr int KS = MmaCfg<D>::KS, RL = MmaCfg<D>::RL;
    const int64_t nt8 = (F + 7) / 8;
    int64_t t_ = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t_ >= S * nt8 * 32) return;
    const int lane = (int)(t_ & 31);
    const int64_t tile = (t_ >> 5) % nt8, j = (t_ >> 5) / nt8;
    const int g = lane >> 2, t = lane & 3;
    const int64_t jw = w_shared ? 0 : j;
    // record element i of a lane lives at [(tile*RL/4 + i/4)*32 + lane][i%4]: every LDS.128 of a warp is contiguous
    float* base = featm + (t_ >> 5) * 32 * RL;
    auto at = [&](int i) -> float& { return base[((i >> 2) * 32 + lane) * 4 + (i & 3)]; };
    const int64_t f = tile * 8 + g;
    for (int s = 0; s < KS; ++s)
        for (int e = 0; e < 2; ++e) {
            float v = 0.0f;
            if (f < F) v = mma_bval<D>(W + (jw * F + f) * D, b[jw * F + f], 8 * s + 4 * e + t);
            at(2 * s + e) = v;
        }
    const int64_t fa = tile * 8 + 2 * t;
    at(2 * KS) = fa < F ? __double2float_rn(theta[j * F + fa]) : 0.0f;
    at(2 * KS + 1) = fa + 1 < F ? __double2float_rn(theta[j * F + fa + 1]) : 0.0f;
    for (int q = 2 * KS + 2; q < RL; ++q) at(q) = 0.0f;
}

// A-operand (candidate side) value at compile-time K index k
template <int D>
__device__ __forceinline__ float mma_aval(const float (&xh)[D], const float (&xl)[D], int k) {
    if (k < D) return xh[k];
    if (k < 2 * D) return xh[k - D];
    if (k < 3 * D) return xl[k - 2 * D];
    if (k == 3 * D || k == 3 * D + 1) return 1.0f;
    return 0.0f;
}

template <int D, int RT, int MINB, int NPOLY>
__global__ void __launch_bounds__(SCR_THREADS, MINB)
    rff_screen_mma_kernel(const double* __restrict__ cand, int64_t N, const float* __restrict__ featm, int S, int F,
                          double scale, int k, int sg, int64_t chunk, int64_t idx_offset,
                          const double* __restrict__ bound, const int* __restrict__ mode,
                          double* __restrict__ slot_val, int64_t* __restrict__ slot_idx, int slot_stride,
                          int* __restrict__ flags) {
    constexpr int KS = MmaCfg<D>::KS, RL = MmaCfg<D>::RL;
    constexpr int THREADS = SCR_THREADS;
    constexpr int TILE = 8 * RT * 16;  // candidates per CTA tile: 8 warps x RT row tiles x 16 rows
    constexpr int TPS = RFF_FCH / 8;               // feature tiles per pipeline stage
    constexpr int STAGE_FLOATS = TPS * 32 * RL;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* stage = reinterpret_cast<float*>(smem_raw);                                  // [NSTAGE][STAGE_FLOATS]
    uint64_t* full = reinterpret_cast<uint64_t*>(stage + RFF_NSTAGE * STAGE_FLOATS);    // [8]
    double* tk_val = reinterpret_cast<double*>(full + 8);
    int64_t* tk_idx = reinterpret_cast<int64_t*>(tk_val + sg * k);
    double* ap_val = reinterpret_cast<double*>(tk_idx + sg * k);                        // [TILE + 16]
    int64_t* ap_idx = reinterpret_cast<int64_t*>(ap_val + TILE + 16);
    double* wm = reinterpret_cast<double*>(ap_idx + TILE + 16);
    int* slot_cnt = reinterpret_cast<int*>(wm + 8);
    int* ap_count = slot_cnt + sg;

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int nsg = (S + sg - 1) / sg;
    const int64_t chunk_id = blockIdx.x / nsg;
    const int gq = blockIdx.x % nsg;
    const int j0 = gq * sg;
    const int nj = min(sg, S - j0);
    const int64_t c0 = chunk_id * chunk;
    const int64_t c1 = min(N, c0 + chunk);
    const int ntiles = (int)((c1 - c0 + TILE - 1) / TILE);
    const int nt8 = (F + 7) / 8;                       // feature tiles per sample
    const int nch = (nt8 + TPS - 1) / TPS;             // pipeline stages per sample
    const int64_t total = (int64_t)ntiles * nj * nch;

    for (int e = tid; e < sg * k; e += THREADS) {
        tk_val[e] = neg_inf();
        tk_idx[e] = -1;
    }
    for (int e = tid; e < nj * SCR_CAP; e += THREADS) {
        int64_t o = (chunk_id * S + j0 + e / SCR_CAP) * slot_stride + e % SCR_CAP;
        slot_idx[o] = -1;
        slot_val[o] = neg_inf();
    }
    for (int e = tid; e < nj; e += THREADS) {
        slot_cnt[e] = 0;
        flags[chunk_id * S + j0 + e] = 0;
    }
    if (tid == 0) {
        *ap_count = 0;
        for (int s = 0; s < RFF_NSTAGE; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    const ScreenCtx cx{tk_val, tk_idx, ap_val, ap_idx, wm, slot_cnt, ap_count, bound, slot_val, slot_idx, flags,
                       k, j0, S, slot_stride, chunk_id, c1, idx_offset};

    int p_jj = 0, p_c = 0, p_slot = 0;
    int64_t p_left = total;
    auto issue = [&]() {
        const int tcnt = min(TPS, nt8 - p_c * TPS);
        const uint32_t bytes = (uint32_t)(tcnt * 32 * RL * sizeof(float));
        const float* src = featm + ((int64_t)(j0 + p_jj) * nt8 + (int64_t)p_c * TPS) * 32 * RL;
        mbar_expect_tx(&full[p_slot], bytes);
        bulk_g2s(stage + p_slot * STAGE_FLOATS, src, bytes, &full[p_slot]);
        if (++p_c == nch) {
            p_c = 0;
            if (++p_jj == nj) p_jj = 0;
        }
        if (++p_slot == RFF_NSTAGE) p_slot = 0;
        --p_left;
    };
    if (tid == 0) {
        for (int p = 0; p < RFF_NSTAGE - 1 && p_left > 0; ++p) issue();
    }

    const uint64_t INV2PI2 = f2_pack(0.15915494f, 0.15915494f);
    const uint64_t MAGIC2 = f2_pack(12582912.0f, 12582912.0f);
    const uint64_t NMAGIC2 = f2_pack(-12582912.0f, -12582912.0f);
    const uint64_t NTWOPI2 = f2_pack(-6.2831855f, -6.2831855f);

    uint32_t afr[RT][KS][4];   // A fragments: [row tile][k-step][a0..a3]
    double dacc[RT][2];        // [row tile][row g | row g+8], partial over this lane's feature columns
    int slot = 0;
    uint32_t phase = 0;

    for (int tile = 0; tile < ntiles; ++tile) {
        const int64_t t0 = c0 + (int64_t)tile * TILE + warp * (RT * 16);
#pragma unroll
        for (int r = 0; r < RT; ++r) {
            float xh[2][D], xl[2][D];
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
                const int64_t ci_ = t0 + r * 16 + g + 8 * h2;
#pragma unroll
                for (int kk = 0; kk < D; ++kk) {
                    const float xf = (ci_ < c1) ? __double2float_rn(__ldg(cand + ci_ * D + kk)) : 0.0f;
                    xh[h2][kk] = __uint_as_float(to_tf32(xf));
                    xl[h2][kk] = __uint_as_float(to_tf32(xf - xh[h2][kk]));
                }
            }
#pragma unroll
            for (int s = 0; s < KS; ++s)
#pragma unroll
                for (int e = 0; e < 2; ++e)
#pragma unroll
                    for (int h2 = 0; h2 < 2; ++h2) {
                        const float v0 = mma_aval<D>(xh[h2], xl[h2], 8 * s + 4 * e + 0);
                        const float v1 = mma_aval<D>(xh[h2], xl[h2], 8 * s + 4 * e + 1);
                        const float v2 = mma_aval<D>(xh[h2], xl[h2], 8 * s + 4 * e + 2);
                        const float v3 = mma_aval<D>(xh[h2], xl[h2], 8 * s + 4 * e + 3);
                        const float v = t == 0 ? v0 : (t == 1 ? v1 : (t == 2 ? v2 : v3));
                        afr[r][s][2 * e + h2] = __float_as_uint(v);  // a0:(g,t) a1:(g+8,t) a2:(g,t+4) a3:(g+8,t+4)
                    }
        }
        for (int jj = 0; jj < nj; ++jj) {
#pragma unroll
            for (int r = 0; r < RT; ++r) dacc[r][0] = dacc[r][1] = 0.0;
            const int red = mode[j0 + jj];
            for (int c = 0; c < nch; ++c) {
                if (tid == 0 && p_left > 0) issue();
                mbar_wait(&full[slot], phase);
                const int tcnt = min(TPS, nt8 - c * TPS);
                const float* sp = stage + slot * STAGE_FLOATS + lane * 4;
                auto run_stage = [&](auto reduce_tag) {
                    constexpr bool REDUCE = decltype(reduce_tag)::value;
                    uint64_t acc2[RT][2];
#pragma unroll
                    for (int r = 0; r < RT; ++r) acc2[r][0] = acc2[r][1] = 0ull;
#pragma unroll 2
                    for (int ft = 0; ft < tcnt; ++ft) {
                        float rec[RL];
#pragma unroll
                        for (int q = 0; q < RL / 4; ++q) {
                            const float4 v = *reinterpret_cast<const float4*>(sp + (ft * (RL / 4) + q) * 128);
                            rec[4 * q] = v.x;
                            rec[4 * q + 1] = v.y;
                            rec[4 * q + 2] = v.z;
                            rec[4 * q + 3] = v.w;
                        }
                        const uint64_t th2 = f2_pack(rec[2 * KS], rec[2 * KS + 1]);
                        // all RT x KS MMAs first, k-step major: RT independent accumulator chains in flight
                        float cc[RT][4];
#pragma unroll
                        for (int r = 0; r < RT; ++r) cc[r][0] = cc[r][1] = cc[r][2] = cc[r][3] = 0.f;
#pragma unroll
                        for (int s = 0; s < KS; ++s)
#pragma unroll
                            for (int r = 0; r < RT; ++r)
                                mma_tf32(cc[r], afr[r][s], __float_as_uint(rec[2 * s]), __float_as_uint(rec[2 * s + 1]));
#pragma unroll
                        for (int r = 0; r < RT; ++r) {
                            uint64_t h01 = f2_pack(cc[r][0], cc[r][1]), h23 = f2_pack(cc[r][2], cc[r][3]);
                            if (REDUCE) {
                                uint64_t ta = ffma2(h01, INV2PI2, MAGIC2), tb = ffma2(h23, INV2PI2, MAGIC2);
                                h01 = ffma2(fadd2(ta, NMAGIC2), NTWOPI2, h01);
                                h23 = ffma2(fadd2(tb, NMAGIC2), NTWOPI2, h23);
                            }
                            // the last NPOLY of the 2*RT accumulator pairs of a feature tile take their cosine on the
                            // FMA pipe (polynomial), the others on the MUFU: both pipes work concurrently
                            uint64_t c01, c23;
                            if (2 * r >= 2 * RT - NPOLY)
                                c01 = cos_poly2(f2_pack(cc[r][0], cc[r][1]));
                            else
                                c01 = f2_pack(__cosf(f2_lo(h01)), __cosf(f2_hi(h01)));
                            if (2 * r + 1 >= 2 * RT - NPOLY)
                                c23 = cos_poly2(f2_pack(cc[r][2], cc[r][3]));
                            else
                                c23 = f2_pack(__cosf(f2_lo(h23)), __cosf(f2_hi(h23)));
                            acc2[r][0] = ffma2(th2, c01, acc2[r][0]);
                            acc2[r][1] = ffma2(th2, c23, acc2[r][1]);
                        }
                    }
#pragma unroll
                    for (int r = 0; r < RT; ++r) {
                        dacc[r][0] += (double)f2_lo(acc2[r][0]) + (double)f2_hi(acc2[r][0]);
                        dacc[r][1] += (double)f2_lo(acc2[r][1]) + (double)f2_hi(acc2[r][1]);
                    }
                };
                if (red)
                    run_stage(std::true_type{});
                else
                    run_stage(std::false_type{});
                if (++slot == RFF_NSTAGE) {
                    slot = 0;
                    phase ^= 1u;
                }
                __syncthreads();
            }
            // quad reduction: the 4 lanes sharing g hold partial sums over disjoint feature columns
#pragma unroll
            for (int r = 0; r < RT; ++r)
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2) {
                    double v = dacc[r][h2];
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
                    v += __shfl_xor_sync(0xffffffffu, v, 2);
                    dacc[r][h2] = v;
                }
            // lane t of a quad takes row tile t: candidates (t0 + t*16 + g) and (+8)
            double u[2];
            int64_t ci[2];
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
                double v = dacc[0][h2];
#pragma unroll
                for (int r = 1; r < RT; ++r) v = (t == r) ? dacc[r][h2] : v;
                const bool own = t < RT;  // lanes beyond the row-tile count own no candidate
                ci[h2] = own ? t0 + t * 16 + g + 8 * h2 : c1;
                u[h2] = (ci[h2] < c1) ? scale * v : neg_inf();
            }
            screen_commit<2>(cx, jj, u, ci);
        }
    }
}

// test probe: max over random (x, w, b) of |h_mma - h_fp64| / A for one K-step configuration D
template <int D>
__global__ void mma_arg_probe_kernel(const double* __restrict__ x, const double* __restrict__ w,
                                     const double* __restrict__ b, int n8, double* out) {
    // one warp: 16 candidates x (8*n8) features
    constexpr int KS = MmaCfg<D>::KS;
    const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
    uint32_t afr[KS][4];
    float xh[2][D], xl[2][D];
    for (int h2 = 0; h2 < 2; ++h2)
        for (int kk = 0; kk < D; ++kk) {
            const float xf = __double2float_rn(x[(g + 8 * h2) * D + kk]);
            xh[h2][kk] = __uint_as_float(to_tf32(xf));
            xl[h2][kk] = __uint_as_float(to_tf32(xf - xh[h2][kk]));
        }
#pragma unroll
    for (int s = 0; s < KS; ++s)
#pragma unroll
        for (int e = 0; e < 2; ++e)
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
                const float v0 = mma_aval<D>(xh[h2], xl[h2], 8 * s + 4 * e + 0);
                const float v1 = mma_aval<D>(xh[h2], xl[h2], 8 * s + 4 * e + 1);
                const float v2 = mma_aval<D>(xh[h2], xl[h2], 8 * s + 4 * e + 2);
                const float v3 = mma_aval<D>(xh[h2], xl[h2], 8 * s + 4 * e + 3);
                afr[s][2 * e + h2] = __float_as_uint(t == 0 ? v0 : (t == 1 ? v1 : (t == 2 ? v2 : v3)));
            }
    double worst = 0.0;
    for (int tile = 0; tile < n8; ++tile) {
        float cc[4] = {0.f, 0.f, 0.f, 0.f};
        const int64_t f = tile * 8 + g;
#pragma unroll
        for (int s = 0; s < KS; ++s) {
            const float b0 = mma_bval<D>(w + f * D, b[f], 8 * s + t), b1 = mma_bval<D>(w + f * D, b[f], 8 * s + t + 4);
            mma_tf32(cc, afr[s], __float_as_uint(b0), __float_as_uint(b1));
        }
        for (int q = 0; q < 4; ++q) {
            const int row = g + 8 * (q >> 1);
            const int64_t ff = tile * 8 + 2 * t + (q & 1);
            double h = b[ff], A = fabs(b[ff]);
            for (int kk = 0; kk < D; ++kk) {
                h = fma(w[ff * D + kk], x[row * D + kk], h);
                A += fabs(w[ff * D + kk] * x[row * D + kk]);
            }
            worst = fmax(worst, fabs((double)cc[q] - h) / A);
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) worst = fmax(worst, __shfl_xor_sync(0xffffffffu, worst, off));
    if (lane == 0) *out = worst;
}

// exact fp64 re-evaluation (tes_spec.h sequence) of every kept (chunk, sample, slot) candidate: one warp per
// (chunk, sample); the 32 lanes evaluate cos(pi h_f) of 32 consecutive features in parallel, then every lane replays
// the SEQUENTIAL accumulation acc = fma(theta_f, c_f, acc), f ascending, from a per-warp shared staging buffer -- the
// same operation order as the fp64 kernel, hence the same bits.
constexpr int REFINE_WARPS = 8;
__global__ void __launch_bounds__(32 * REFINE_WARPS)
    rff_refine_kernel(const double* __restrict__ cand, int d, const double* __restrict__ feat, int rec, int64_t S,
                      int64_t F, double scale, int64_t idx_offset, int64_t npairs, int slot_stride,
                      double* __restrict__ slot_val, int64_t* __restrict__ slot_idx, const int* __restrict__ flags) {
    __shared__ double2 stage[REFINE_WARPS][32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t cs = (int64_t)blockIdx.x * REFINE_WARPS + wib;  // chunk*S + j
    if (cs >= npairs) return;
    const int64_t o0 = cs * slot_stride;
    if (flags[cs]) {  // overflowed: the fp64 kernel recomputes this (chunk, sample)
        for (int c = lane; c < SCR_CAP; c += 32) {
            slot_idx[o0 + c] = -1;
            slot_val[o0 + c] = neg_inf();
        }
        return;
    }
    const double* fj = feat + (cs % S) * F * rec;
    for (int c = 0; c < SCR_CAP; ++c) {
        const int64_t gi = slot_idx[o0 + c];
        if (gi < 0) continue;
        const double* x = cand + (gi - idx_offset) * d;
        double xr[RFF_MAX_D_FAST];
#pragma unroll
        for (int kk = 0; kk < RFF_MAX_D_FAST; ++kk) xr[kk] = kk < d ? __ldg(x + kk) : 0.0;
        double acc = 0.0;
        for (int64_t f0 = 0; f0 < F; f0 += 32) {
            const int64_t f = f0 + lane;
            double2 tc = make_double2(0.0, 0.0);
            if (f < F) {
                const double* r = fj + f * rec;
                double h = __ldg(r + d);
#pragma unroll
                for (int kk = 0; kk < RFF_MAX_D_FAST; ++kk)
                    if (kk < d) h = fma(__ldg(r + kk), xr[kk], h);
                tc = make_double2(__ldg(r + d + 1), cospi_spec(h));
            }
            __syncwarp();
            stage[wib][lane] = tc;
            __syncwarp();
            const int nb = (int)(F - f0 < 32 ? F - f0 : 32);
            if (nb == 32) {
#pragma unroll
                for (int l = 0; l < 32; ++l) {
                    const double2 v = stage[wib][l];
                    acc = fma(v.x, v.y, acc);
                }
            } else {
                for (int l = 0; l < nb; ++l) {
                    const double2 v = stage[wib][l];
                    acc = fma(v.x, v.y, acc);
                }
            }
        }
        if (lane == 0) slot_val[o0 + c] = __dmul_rn(scale, acc);
    }
}

#include "rff_tc5.cuh"
#include "rff_tc5t.cuh"

struct RffPlan {
    int C, threads, unroll, tile, sg, rec;
    int64_t chunk, nchunks, nunits;
    int64_t feat_bytes, part_val_bytes, part_idx_bytes, dense_bytes;
    size_t smem;
    // fp32 screen
    int screened, recp, slot_stride;
    int64_t feat32_bytes, aux_bytes, flags_bytes, total_bytes;
    size_t smem_screen, smem_fallback;
    int64_t nunits_screen;
    int sg_screen;
};

constexpr int64_t SCR_MIN_N = 32768;  // below this the screen's fixed costs are not worth it

static size_t rff_smem_fp64(int rec, int sg, int k, int tile) {
    return (size_t)RFF_NSTAGE * RFF_FCH * rec * sizeof(double) + 8 * sizeof(uint64_t) + (size_t)sg * k * 16 +
           (size_t)(tile + RFF_MAX_K) * 16 + 16;
}

// method: 0 = auto, 1 = fp64 only, 2 = packed-fp32 screen + fp64 refine, 3 = tensor-core (3xTF32 mma.sync) screen +
// fp64 refine, 4 = tcgen05/TMEM (3xTF32) screen + fp64 refine, 5 = the same with the transposed (feature-major)
// accumulator layout (2..5 only when the shape allows it)
static RffPlan rff_plan(int64_t N, int64_t d, int64_t S, int64_t F, int k, int method) {
    RffPlan p{};
    p.C = 4;
    p.threads = 256;
    p.unroll = 2;
    if (const char* v = getenv("TES_RFF_VARIANT")) {  // developer knob: "C,threads,unroll"
        int a = 0, b2 = 0, c2 = 0;
        if (sscanf(v, "%d,%d,%d", &a, &b2, &c2) == 3) {
            p.C = a;
            p.threads = b2;
            p.unroll = c2;
        }
    }
    p.rec = rff_rec((int)d);
    p.recp = scr_recp((int)d);
    const bool can_screen = d <= RFF_MAX_D_FAST && k <= SCR_MAX_K && N >= 1;
    p.screened = 0;
    if (can_screen && (method >= 2 || (method == 0 && N >= SCR_MIN_N)))
        p.screened = method == 5 ? 4 : ((method == 4 || (method == 0 && d >= 5)) ? 3 : (method == 3 ? 2 : 1));
    if (p.screened) {
        p.C = 4;
        p.threads = 256;
        p.unroll = 2;
    }
    p.tile = p.threads * p.C;
    if (p.screened == 2) {  // candidate tile of the tensor-core screen (the fp64 fallback tiles itself)
        int rt = MMA_RT, mb = 2;
        if (const char* v = getenv("TES_MMA_VARIANT")) sscanf(v, "%d,%d", &rt, &mb);
        p.tile = 8 * rt * 16;
    }
    if (p.screened == 3) p.tile = TC5_TILE;
    if (p.screened == 4) p.tile = T5T_TILE;
    int64_t ntiles = (N + p.tile - 1) / p.tile;
    if (ntiles < 1) ntiles = 1;
    int64_t tpc = ntiles < 16 ? ntiles : 16;  // tiles per chunk
    int64_t sg = S < 32 ? S : 32;
    if (p.screened >= 3) {
        // tcgen05 screen: a CTA holds its candidate tile while a sample group streams past; 8-sample groups quarter the
        // CTA duration (finer wave quantisation: 13.4 -> 53.6 waves at C3, measured -4 %) at the cost of more tile
        // conversions
        int sgt = 8;
        if (const char* v = getenv("TES_TC5_SG")) sscanf(v, "%d", &sgt);  // developer knob
        if (sgt >= 1 && sgt <= 32 && S > sgt) sg = sgt;
    }
    auto units = [&]() { return ((ntiles + tpc - 1) / tpc) * ((S + sg - 1) / sg); };
    const int64_t want = 8 * 148;
    if (p.screened >= 2) tpc = ntiles < 32 ? ntiles : 32;
    if (p.screened == 4) tpc = ntiles < 64 ? ntiles : 64;  // 256-row tiles: same 16 384-candidate chunks
    while (units() < want && tpc > 4) tpc = (tpc + 1) / 2;
    while (units() < want && sg > 1) sg = (sg + 1) / 2;
    while (units() < want && tpc > 1) tpc = (tpc + 1) / 2;
    p.sg = (int)sg;
    p.chunk = tpc * p.tile;
    p.nchunks = (ntiles + tpc - 1) / tpc;
    p.nunits = p.nchunks * ((S + sg - 1) / sg);
    p.feat_bytes = align_up(S * F * p.rec * (int64_t)sizeof(double), 256);
    p.dense_bytes = d > RFF_MAX_D_FAST ? align_up(S * N * (int64_t)sizeof(double), 256) : 0;
    p.smem = rff_smem_fp64(p.rec, p.sg, k, p.tile);
    if (p.screened) {
        p.slot_stride = SCR_CAP + k;
        p.sg_screen = p.sg;
        p.nunits_screen = p.nunits;
        p.part_val_bytes = align_up(p.nchunks * S * p.slot_stride * (int64_t)sizeof(double), 256);
        p.part_idx_bytes = align_up(p.nchunks * S * p.slot_stride * (int64_t)sizeof(int64_t), 256);
        p.feat32_bytes = align_up(S * F * p.recp * (int64_t)sizeof(uint64_t), 256);
        if (p.screened == 2) {
            const int ks = (3 * (int)d + 2 + 7) / 8, rl = ((2 * ks + 2) + 3) & ~3;
            p.feat32_bytes = align_up(S * ((F + 7) / 8) * 32 * rl * (int64_t)sizeof(float), 256);
            p.smem_screen = (size_t)RFF_NSTAGE * (RFF_FCH / 8) * 32 * rl * sizeof(float) + 8 * sizeof(uint64_t) +
                            (size_t)p.sg * k * 16 + (size_t)(512 + 16) * 16 + 8 * sizeof(double) +
                            (size_t)p.sg * sizeof(int) + 16;
        }
        if (p.screened == 3) {
            const int kc = tc5_kc((int)d);
            const int64_t sf = tc5_stage_floats(kc);
            p.feat32_bytes = align_up(S * ((F + TC5_NF - 1) / TC5_NF) * sf * (int64_t)sizeof(float), 256);
            p.smem_screen = (size_t)TC5_MT * kc * 128 * 16 + (size_t)TC5_NSTAGE * sf * sizeof(float) +
                            (2 * TC5_NSTAGE + 10) * sizeof(uint64_t) + 4 * TC5_TILE * sizeof(double) +
                            (size_t)p.sg * k * 16 + (size_t)(TC5_TILE + 16) * 16 + (size_t)p.sg * sizeof(int) + 16;
        }
        if (p.screened == 4) {
            const int kc = tc5_kc((int)d);
            const int64_t sf = t5t_stage_floats(kc);
            p.feat32_bytes = align_up(S * ((F + T5T_NF - 1) / T5T_NF) * sf * (int64_t)sizeof(float), 256);
            p.smem_screen = (size_t)kc * T5T_TILE * 16 + (size_t)T5T_NSTAGE * sf * sizeof(float) +
                            (2 * T5T_NSTAGE + 10) * sizeof(uint64_t) + 8 * T5T_TILE * sizeof(float) +
                            T5T_TILE * sizeof(double) + (size_t)p.sg * k * 16 + (size_t)(T5T_TILE + 16) * 16 +
                            (size_t)p.sg * sizeof(int) + 16;
        }
        p.aux_bytes = align_up(64 * 8, 256) + 2 * align_up(S * 8, 256);  // xmax[64], bound[S], mode[S]
        p.flags_bytes = align_up(p.nchunks * S * (int64_t)sizeof(int), 256);
        if (p.screened == 1)
            p.smem_screen = (size_t)RFF_NSTAGE * RFF_FCH * p.recp * sizeof(uint64_t) + 8 * sizeof(uint64_t) +
                            (size_t)p.sg * k * 16 + (size_t)(SCR_TILE + 16) * 16 + 8 * sizeof(double) +
                            (size_t)p.sg * sizeof(int) + 16;
        p.smem_fallback = rff_smem_fp64(p.rec, 1, k, 1024);
    } else {
        p.slot_stride = k;
        p.part_val_bytes = align_up(p.nchunks * S * k * (int64_t)sizeof(double), 256);
        p.part_idx_bytes = align_up(p.nchunks * S * k * (int64_t)sizeof(int64_t), 256);
    }
    p.total_bytes = p.feat_bytes + p.part_val_bytes + p.part_idx_bytes + p.dense_bytes + p.feat32_bytes +
                    p.aux_bytes + p.flags_bytes;
    return p;
}

template <int D, int C, int THREADS, int UNROLL>
static int launch_rff_variant(const RffPlan& p, const double* cand, int64_t N, const double* feat, int64_t S,
                              int64_t F, double scale, int k, int64_t idx_offset, double* part_val,
                              int64_t* part_idx, const int* flags, int out_stride, int out_off, cudaStream_t st) {
    auto kern = rff_topk_kernel<D, C, THREADS, UNROLL>;
    if (p.smem > 48 * 1024)
        TES_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
    kern<<<(unsigned)p.nunits, THREADS, p.smem, st>>>(cand, N, feat, (int)S, (int)F, scale, k, p.sg, p.chunk,
                                                      idx_offset, part_val, part_idx, flags, out_stride, out_off);
    TES_LAUNCH_OK();
    return 0;
}

template <int D>
static int launch_rff_topk(const RffPlan& p, const double* cand, int64_t N, const double* feat, int64_t S, int64_t F,
                           double scale, int k, int64_t idx_offset, double* part_val, int64_t* part_idx,
                           const int* flags, int out_stride, int out_off, cudaStream_t st) {
#define TES_V(CC, TT, UU)                                                                                   \
    if (p.C == CC && p.threads == TT && p.unroll == UU)                                                     \
        return launch_rff_variant<D, CC, TT, UU>(p, cand, N, feat, S, F, scale, k, idx_offset, part_val, part_idx, \
                                                 flags, out_stride, out_off, st);
    TES_V(4, 256, 2)
#ifdef TES_RFF_DEV_VARIANTS
    if constexpr (D == 6 || D == 2 || D == 10) {
        TES_V(4, 256, 1) TES_V(2, 256, 2) TES_V(2, 512, 2) TES_V(4, 128, 2) TES_V(3, 256, 2) TES_V(2, 256, 4)
        TES_V(4, 512, 1) TES_V(3, 256, 1) TES_V(2, 256, 1) TES_V(1, 512, 4) TES_V(1, 256, 4) TES_V(2, 128, 2)
    }
#endif
#undef TES_V
    return -100;
}

template <int D>
static int launch_rff_screen(const RffPlan& p, const double* cand, int64_t N, const uint64_t* feat32, int64_t S,
                             int64_t F, double scale, int k, int64_t idx_offset, const double* bound,
                             const int* mode, double* slot_val, int64_t* slot_idx, int* flags, cudaStream_t st) {
    int un = 8, mb = 2;
    if (const char* v = getenv("TES_SCR_VARIANT")) sscanf(v, "%d,%d", &un, &mb);  // developer knob
    void (*kern)(const double*, int64_t, const uint64_t*, int, int, double, int, int, int64_t, int64_t, const double*,
                 const int*, double*, int64_t*, int, int*) = rff_screen_kernel<D, 8, 2>;
#ifdef TES_RFF_DEV_VARIANTS
    if constexpr (D == 6) {
        if (un == 4 && mb == 1) kern = rff_screen_kernel<D, 4, 1>;
        if (un == 2 && mb == 3) kern = rff_screen_kernel<D, 2, 3>;
        if (un == 4 && mb == 3) kern = rff_screen_kernel<D, 4, 3>;
        if (un == 4 && mb == 2) kern = rff_screen_kernel<D, 4, 2>;
        if (un == 2 && mb == 1) kern = rff_screen_kernel<D, 2, 1>;
        if (un == 1 && mb == 3) kern = rff_screen_kernel<D, 1, 3>;
        if (un == 8 && mb == 1) kern = rff_screen_kernel<D, 8, 1>;
    }
#endif
    if (p.smem_screen > 48 * 1024)
        TES_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_screen));
    TES_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    kern<<<(unsigned)p.nunits_screen, SCR_THREADS, p.smem_screen, st>>>(cand, N, feat32, (int)S, (int)F, scale, k,
                                                                        p.sg_screen, p.chunk, idx_offset, bound, mode,
                                                                        slot_val, slot_idx, p.slot_stride, flags);
    TES_LAUNCH_OK();
    return 0;
}

template <int D>
static int launch_rff_screen_mma(const RffPlan& p, const double* cand, int64_t N, const float* featm, int64_t S,
                                 int64_t F, double scale, int k, int64_t idx_offset, const double* bound,
                                 const int* mode, double* slot_val, int64_t* slot_idx, int* flags, cudaStream_t st) {
    int rt = MMA_RT, mb = 2;
    if (const char* v = getenv("TES_MMA_VARIANT")) sscanf(v, "%d,%d", &rt, &mb);  // developer knob
    void (*kern)(const double*, int64_t, const float*, int, int, double, int, int, int64_t, int64_t, const double*,
                 const int*, double*, int64_t*, int, int*) = rff_screen_mma_kernel<D, MMA_RT, 2, MMA_NPOLY>;
#ifdef TES_RFF_DEV_VARIANTS
    if constexpr (D == 6) {  // developer sweep: second knob = number of polynomial pairs
        if (rt == 4 && mb == 0) kern = rff_screen_mma_kernel<D, 4, 2, 0>;
        if (rt == 4 && mb == 1) kern = rff_screen_mma_kernel<D, 4, 2, 1>;
        if (rt == 4 && mb == 2) kern = rff_screen_mma_kernel<D, 4, 2, 2>;
        if (rt == 4 && mb == 3) kern = rff_screen_mma_kernel<D, 4, 2, 3>;
        if (rt == 4 && mb == 4) kern = rff_screen_mma_kernel<D, 4, 2, 4>;
        if (rt == 4 && mb == 5) kern = rff_screen_mma_kernel<D, 4, 2, 5>;
    }
#endif
    if (p.smem_screen > 48 * 1024)
        TES_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_screen));
    TES_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    kern<<<(unsigned)p.nunits_screen, SCR_THREADS, p.smem_screen, st>>>(cand, N, featm, (int)S, (int)F, scale, k,
                                                                        p.sg_screen, p.chunk, idx_offset, bound, mode,
                                                                        slot_val, slot_idx, p.slot_stride, flags);
    TES_LAUNCH_OK();
    return 0;
}

static int launch_rff_screen_tc5(const RffPlan& p, const double* cand, int64_t N, int d, const float* featt, int64_t S,
                                 int64_t F, double scale, int k, int64_t idx_offset, const double* bound,
                                 const int* mode, double* slot_val, int64_t* slot_idx, int* flags, cudaStream_t st) {
    int np = TC5_NPOLY, pc = TC5_PC;
    if (const char* v = getenv("TES_TC5_NPOLY")) sscanf(v, "%d", &np);  // developer knobs (need the dev build)
    if (const char* v = getenv("TES_TC5_PC")) sscanf(v, "%d", &pc);
    void (*kern)(const double*, int64_t, int, const float*, int, int, int, double, int, int, int64_t, int64_t,
                 const double*, const int*, double*, int64_t*, int, int*) = rff_screen_tc5_kernel<TC5_NPOLY, TC5_PC>;
    int threads = TC5_PC > 0 ? TC5_THREADS_SPEC : TC5_THREADS;
#ifdef TES_RFF_DEV_VARIANTS
    threads = pc > 0 ? TC5_THREADS_SPEC : TC5_THREADS;
    if (pc == 0) {
        if (np == 0) kern = rff_screen_tc5_kernel<0, 0>;
        if (np == 1) kern = rff_screen_tc5_kernel<1, 0>;
        if (np == 2) kern = rff_screen_tc5_kernel<2, 0>;
    }
    if (pc == 8) kern = rff_screen_tc5_kernel<0, 8>;
    if (pc == -2) kern = np ? rff_screen_tc5_kernel<1, -2> : rff_screen_tc5_kernel<0, -2>;
    if (pc == -3) kern = np ? rff_screen_tc5_kernel<1, -3> : rff_screen_tc5_kernel<0, -3>;
    if (pc == -4) kern = np ? rff_screen_tc5_kernel<1, -4> : rff_screen_tc5_kernel<0, -4>;
    if (pc == -6) kern = np ? rff_screen_tc5_kernel<1, -6> : rff_screen_tc5_kernel<0, -6>;
    if (pc == 16) kern = rff_screen_tc5_kernel<0, 16>;
#endif
    TES_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_screen));
    kern<<<(unsigned)p.nunits_screen, threads, p.smem_screen, st>>>(
        cand, N, d, featt, tc5_kc(d), (int)S, (int)F, scale, k, p.sg_screen, p.chunk, idx_offset, bound, mode, slot_val,
        slot_idx, p.slot_stride, flags);
    TES_LAUNCH_OK();
    return 0;
}

static int launch_rff_screen_tc5t(const RffPlan& p, const double* cand, int64_t N, int d, const float* featt,
                                  int64_t S, int64_t F, double scale, int k, int64_t idx_offset, const double* bound,
                                  const int* mode, double* slot_val, int64_t* slot_idx, int* flags, cudaStream_t st) {
    int np = 1;
    if (const char* v = getenv("TES_TC5_NPOLY")) sscanf(v, "%d", &np);  // developer knob (needs the dev build)
    void (*kern)(const double*, int64_t, int, const float*, int, int, int, double, int, int, int64_t, int64_t,
                 const double*, const int*, double*, int64_t*, int, int*) = rff_screen_tc5t_kernel<1>;
#ifdef TES_RFF_DEV_VARIANTS
    if (np == 0) kern = rff_screen_tc5t_kernel<0>;
    if (np == 2) kern = rff_screen_tc5t_kernel<2>;
    if (np == 3) kern = rff_screen_tc5t_kernel<3>;
#endif
    TES_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_screen));
    kern<<<(unsigned)p.nunits_screen, T5T_THREADS, p.smem_screen, st>>>(
        cand, N, d, featt, tc5_kc(d), (int)S, (int)F, scale, k, p.sg_screen, p.chunk, idx_offset, bound, mode, slot_val,
        slot_idx, p.slot_stride, flags);
    TES_LAUNCH_OK();
    return 0;
}

static int rff_topk_run(const double* cand, int64_t N, int64_t d, const double* W, const double* b,
                        const double* theta, int64_t S, int64_t F, int w_shared, double sigma, int k,
                        int64_t idx_offset, int64_t* out_idx, double* out_val, void* ws, int64_t ws_bytes,
                        void* stream, int method) {
    if (N < 0 || (N > 0 && !cand)) return -1;
    if (d < 1 || d > 64) return -3;
    if (!W) return -4;
    if (!b) return -5;
    if (!theta) return -6;
    if (S < 1 || S > (1 << 24)) return -7;
    if (F < 1 || F > (1 << 24)) return -8;
    if (!(sigma >= 0.0)) return -10;
    if (k < 1 || k > RFF_MAX_K) return -11;
    if (!out_idx) return -13;
    if (!out_val) return -14;
    if (method < 0 || method > 5) return -18;
    RffPlan p = rff_plan(N, d, S, F, k, method);
    if (!ws || ws_bytes < p.total_bytes) return -100;
    cudaStream_t st = (cudaStream_t)stream;
    char* base = (char*)ws;
    double* feat = (double*)base;
    base += p.feat_bytes;
    double* part_val = (double*)base;
    base += p.part_val_bytes;
    int64_t* part_idx = (int64_t*)base;
    base += p.part_idx_bytes;
    double* dense = (double*)base;
    base += p.dense_bytes;
    uint64_t* feat32 = (uint64_t*)base;
    base += p.feat32_bytes;
    unsigned long long* xmax = (unsigned long long*)base;
    double* bound = (double*)(base + align_up(64 * 8, 256));
    int* mode = (int*)(base + align_up(64 * 8, 256) + align_up(S * 8, 256));
    base += p.aux_bytes;
    int* flags = (int*)base;
    const double scale = sqrt(2.0 * sigma / (double)F);

    const int64_t nsf = S * F;
    rff_pack_kernel<<<(unsigned)((nsf + 255) / 256), 256, 0, st>>>(W, b, theta, S, F, (int)d, w_shared, p.rec, feat);
    TES_LAUNCH_OK();
    int64_t P = p.nchunks;
    int kin = k;
    if (N == 0) {
        P = 0;
    } else if (p.screened) {
        TES_CUDA_OK(cudaMemsetAsync(xmax, 0, 64 * 8, st));
        rff_xmax_kernel<<<148 * 4, 256, 0, st>>>(cand, N, (int)d, xmax);
        TES_LAUNCH_OK();
        int rc = 0;
        if (p.screened == 4) {
            const int64_t nrow = S * ((F + T5T_NF - 1) / T5T_NF) * T5T_NF;
            rff_bound_kernel<<<(unsigned)S, 32, 0, st>>>(feat, p.rec, S, F, (int)d, xmax, scale, MMA_HARG, bound, mode);
            TES_LAUNCH_OK();
            rff_pack_tc5_kernel<T5T_NF><<<(unsigned)((nrow + 255) / 256), 256, 0, st>>>(
                W, b, theta, S, F, (int)d, w_shared, tc5_kc((int)d), (float*)feat32);
            TES_LAUNCH_OK();
            rc = launch_rff_screen_tc5t(p, cand, N, (int)d, (const float*)feat32, S, F, scale, k, idx_offset, bound,
                                        mode, part_val, part_idx, flags, st);
        } else if (p.screened == 3) {
            const int64_t nrow = S * ((F + TC5_NF - 1) / TC5_NF) * TC5_NF;
            rff_bound_kernel<<<(unsigned)S, 32, 0, st>>>(feat, p.rec, S, F, (int)d, xmax, scale, MMA_HARG, bound, mode);
            TES_LAUNCH_OK();
            rff_pack_tc5_kernel<TC5_NF><<<(unsigned)((nrow + 255) / 256), 256, 0, st>>>(W, b, theta, S, F, (int)d, w_shared,
                                                                                 tc5_kc((int)d), (float*)feat32);
            TES_LAUNCH_OK();
            rc = launch_rff_screen_tc5(p, cand, N, (int)d, (const float*)feat32, S, F, scale, k, idx_offset, bound, mode,
                                       part_val, part_idx, flags, st);
        } else if (p.screened == 2) {
            const int64_t nrec = S * ((F + 7) / 8) * 32;
            rff_bound_kernel<<<(unsigned)S, 32, 0, st>>>(feat, p.rec, S, F, (int)d, xmax, scale, MMA_HARG, bound, mode);
            TES_LAUNCH_OK();
            switch (d) {
#define TES_CASE(DD)                                                                                            \
    case DD:                                                                                                    \
        rff_pack_mma_kernel<DD><<<(unsigned)((nrec + 255) / 256), 256, 0, st>>>(W, b, theta, S, F, w_shared,    \
                                                                                (float*)feat32);                \
        TES_LAUNCH_OK();                                                                                        \
        rc = launch_rff_screen_mma<DD>(p, cand, N, (const float*)feat32, S, F, scale, k, idx_offset, bound,     \
                                       mode, part_val, part_idx, flags, st);                                    \
        break;
                TES_CASE(1) TES_CASE(2) TES_CASE(3) TES_CASE(4) TES_CASE(5) TES_CASE(6) TES_CASE(7) TES_CASE(8)
                TES_CASE(9) TES_CASE(10)
#undef TES_CASE
            }
        } else {
            rff_pack32_kernel<<<(unsigned)((nsf + 255) / 256), 256, 0, st>>>(W, b, theta, S, F, (int)d, w_shared,
                                                                              p.recp, feat32);
            TES_LAUNCH_OK();
            rff_bound_kernel<<<(unsigned)S, 32, 0, st>>>(feat, p.rec, S, F, (int)d, xmax, scale,
                                                         (d + 4) * 5.9604644775390625e-08, bound, mode);
            TES_LAUNCH_OK();
            switch (d) {
#define TES_CASE(DD)                                                                                            \
    case DD:                                                                                                    \
        rc = launch_rff_screen<DD>(p, cand, N, feat32, S, F, scale, k, idx_offset, bound, mode, part_val,      \
                                   part_idx, flags, st);                                                        \
        break;
                TES_CASE(1) TES_CASE(2) TES_CASE(3) TES_CASE(4) TES_CASE(5) TES_CASE(6) TES_CASE(7) TES_CASE(8)
                TES_CASE(9) TES_CASE(10)
#undef TES_CASE
            }
        }
        if (rc) return rc;
        const int64_t npairs = p.nchunks * S;
        rff_refine_kernel<<<(unsigned)((npairs + REFINE_WARPS - 1) / REFINE_WARPS), 32 * REFINE_WARPS, 0, st>>>(
            cand, (int)d, feat, p.rec, S, F, scale, idx_offset, npairs, p.slot_stride, part_val, part_idx, flags);
        TES_LAUNCH_OK();
        // flag-gated fp64 recomputation of overflowed (chunk, sample) units, one CTA each
        RffPlan fb = p;
        fb.C = 4;
        fb.threads = 256;
        fb.unroll = 2;
        fb.tile = 1024;
        fb.sg = 1;
        fb.nunits = p.nchunks * S;
        fb.smem = p.smem_fallback;
        switch (d) {
#define TES_CASE(DD)                                                                                            \
    case DD:                                                                                                    \
        rc = launch_rff_topk<DD>(fb, cand, N, feat, S, F, scale, k, idx_offset, part_val, part_idx, flags,     \
                                 p.slot_stride, SCR_CAP, st);                                                   \
        break;
            TES_CASE(1) TES_CASE(2) TES_CASE(3) TES_CASE(4) TES_CASE(5) TES_CASE(6) TES_CASE(7) TES_CASE(8)
            TES_CASE(9) TES_CASE(10)
#undef TES_CASE
        }
        if (rc) return rc;
        kin = p.slot_stride;
    } else if (d <= RFF_MAX_D_FAST) {
        int rc = 0;
        switch (d) {
#define TES_CASE(DD)                                                                                        \
    case DD:                                                                                                \
        rc = launch_rff_topk<DD>(p, cand, N, feat, S, F, scale, k, idx_offset, part_val, part_idx, nullptr, \
                                 k, 0, st);                                                                 \
        break;
            TES_CASE(1) TES_CASE(2) TES_CASE(3) TES_CASE(4) TES_CASE(5) TES_CASE(6) TES_CASE(7) TES_CASE(8)
            TES_CASE(9) TES_CASE(10)
#undef TES_CASE
        }
        if (rc) return rc;
    } else {
        dim3 grid((unsigned)((N + 127) / 128), (unsigned)S);
        rff_value_kernel_generic<<<grid, 128, 0, st>>>(cand, N, (int)d, feat, (int)S, (int)F, p.rec, scale, dense);
        TES_LAUNCH_OK();
        rows_topk_kernel<<<(unsigned)S, 32, 0, st>>>(dense, N, k, idx_offset, part_val, part_idx);
        TES_LAUNCH_OK();
        P = 1;
    }
    topk_merge_kernel<<<(unsigned)S, 32, 0, st>>>(part_val, part_idx, P, S, kin, k, out_val, out_idx);
    TES_LAUNCH_OK();
    return 0;
}

}  // namespace tes

using namespace tes;

extern "C" int64_t tes_rff_topk_workspace_bytes(int64_t N, int64_t d, int64_t S, int64_t F, int k) {
    if (N < 0 || d < 1 || d > 64 || S < 1 || F < 1 || k < 1 || k > RFF_MAX_K) return -1;
    // large enough for every method
    int64_t best = 0;
    for (int m = 1; m <= 5; ++m) {
        int64_t v = rff_plan(N, d, S, F, k, m).total_bytes;
        if (v > best) best = v;
    }
    return best;
}

extern "C" int tes_rff_topk_f64(const double* cand, int64_t N, int64_t d, const double* W, const double* b,
                                const double* theta, int64_t S, int64_t F, int w_shared, double sigma, int k,
                                int64_t idx_offset, int64_t* out_idx, double* out_val, void* ws, int64_t ws_bytes,
                                void* stream) {
    return rff_topk_run(cand, N, d, W, b, theta, S, F, w_shared, sigma, k, idx_offset, out_idx, out_val, ws,
                        ws_bytes, stream, 0);
}

extern "C" int tes_rff_topk_method_f64(const double* cand, int64_t N, int64_t d, const double* W, const double* b,
                                       const double* theta, int64_t S, int64_t F, int w_shared, double sigma, int k,
                                       int64_t idx_offset, int64_t* out_idx, double* out_val, void* ws,
                                       int64_t ws_bytes, void* stream, int method) {
    return rff_topk_run(cand, N, d, W, b, theta, S, F, w_shared, sigma, k, idx_offset, out_idx, out_val, ws,
                        ws_bytes, stream, method);
}

extern "C" int tes_topk_merge_f64(const double* vals, const int64_t* idx, int64_t P, int64_t S, int k,
                                  double* out_val, int64_t* out_idx, void* stream) {
    if (!vals) return -1;
    if (!idx) return -2;
    if (P < 0) return -3;
    if (S < 1) return -4;
    if (k < 1 || k > 64) return -5;
    if (!out_val) return -6;
    if (!out_idx) return -7;
    topk_merge_kernel<<<(unsigned)S, 32, 0, (cudaStream_t)stream>>>(vals, idx, P, S, k, k, out_val, out_idx);
    TES_LAUNCH_OK();
    return 0;
}

namespace tes {
// dense values with on-the-fly prescale (no workspace): used by tes_rff_eval_f64
__global__ void rff_eval_kernel(const double* __restrict__ cand, int64_t N, int d, const double* __restrict__ W,
                                const double* __restrict__ b, const double* __restrict__ theta, int S, int F,
                                int w_shared, double scale, double* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int j = blockIdx.y;
    if (i >= N) return;
    const double* x = cand + i * d;
    const int64_t jw = w_shared ? 0 : j;
    const double* wj = W + jw * F * d;
    const double* bj = b + jw * F;
    const double* tj = theta + (int64_t)j * F;
    double acc = 0.0;
    for (int f = 0; f < F; ++f) {
        double h = __dmul_rn(__ldg(bj + f), TES_INV_PI);
        for (int kk = 0; kk < d; ++kk) h = fma(__dmul_rn(__ldg(wj + (int64_t)f * d + kk), TES_INV_PI), __ldg(x + kk), h);
        acc = fma(__ldg(tj + f), cospi_spec(h), acc);
    }
    out[(int64_t)j * N + i] = __dmul_rn(scale, acc);
}
}  // namespace tes

extern "C" int tes_rff_eval_f64(const double* cand, int64_t N, int64_t d, const double* W, const double* b,
                                const double* theta, int64_t S, int64_t F, int w_shared, double sigma, double* out,
                                void* stream) {
    if (N < 0 || (N > 0 && !cand)) return -1;
    if (d < 1 || d > 64) return -3;
    if (!W) return -4;
    if (!b) return -5;
    if (!theta) return -6;
    if (S < 1 || S > 65535) return -7;
    if (F < 1) return -8;
    if (!(sigma >= 0.0)) return -10;
    if (!out) return -11;
    if (N == 0) return 0;
    const double scale = sqrt(2.0 * sigma / (double)F);
    dim3 grid((unsigned)((N + 127) / 128), (unsigned)S);
    rff_eval_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(cand, N, (int)d, W, b, theta, (int)S, (int)F, w_shared,
                                                            scale, out);
    TES_LAUNCH_OK();
    return 0;
}

extern "C" int tes_cosf_error_probe(float lo, float hi, double e1, int which, double* out_max, void* stream) {
    if (!(lo >= 0.0f) || !(hi >= lo)) return -1;
    if (!out_max) return -5;
    unsigned lb, hb;
    memcpy(&lb, &lo, 4);
    memcpy(&hb, &hi, 4);
    cudaStream_t st = (cudaStream_t)stream;
    TES_CUDA_OK(cudaMemsetAsync(out_max, 0, sizeof(double), st));
    if (which < 0 || which > 1) return -4;
    cosf_probe_kernel<<<148 * 8, 256, 0, st>>>(lb, hb, e1, which, out_max);
    TES_LAUNCH_OK();
    return 0;
}

// arg-max of a vector, first maximum wins (tf.argmax / np.argmax; bo_discrete.py:806, :1053-1055).
namespace tes {
__device__ __forceinline__ int64_t warp_min_i64(int64_t v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const int64_t o = __shfl_xor_sync(0xffffffffu, v, off);
        v = o < v ? o : v;
    }
    return v;
}
// np.argmax / tf.argmax treat NaN as the maximum: the FIRST NaN wins, also over a +inf at a lower index.  A block that
// saw a NaN reports (NaN, index of its first NaN); the final kernel takes the lowest such index.
__global__ void argmax_partial_kernel(const double* __restrict__ v, int64_t N, double* __restrict__ pv,
                                      int64_t* __restrict__ pi) {
    double bv = neg_inf();
    int64_t bi = INT64_MAX;
    int64_t nan_i = INT64_MAX;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < N; e += (int64_t)gridDim.x * blockDim.x) {
        double x = v[e];
        if (x != x) {
            if (e < nan_i) nan_i = e;
        } else if (better(x, e, bv, bi)) {
            bv = x;
            bi = e;
        }
    }
    warp_argmax(bv, bi);
    nan_i = warp_min_i64(nan_i);
    __shared__ double sv[32];
    __shared__ int64_t si[32], sn[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        sv[warp] = bv;
        si[warp] = bi;
        sn[warp] = nan_i;
    }
    __syncthreads();
    if (warp == 0) {
        bv = lane < (blockDim.x >> 5) ? sv[lane] : neg_inf();
        bi = lane < (blockDim.x >> 5) ? si[lane] : INT64_MAX;
        nan_i = lane < (blockDim.x >> 5) ? sn[lane] : INT64_MAX;
        warp_argmax(bv, bi);
        nan_i = warp_min_i64(nan_i);
        if (lane == 0) {
            const bool has_nan = nan_i != INT64_MAX;
            pv[blockIdx.x] = has_nan ? __longlong_as_double(0x7ff8000000000000LL) : bv;
            pi[blockIdx.x] = has_nan ? nan_i : (bi == INT64_MAX ? -1 : bi);
        }
    }
}
__global__ void argmax_final_kernel(const double* __restrict__ v, const double* __restrict__ pv,
                                    const int64_t* __restrict__ pi, int nb, double* out_val, int64_t* out_idx) {
    double bv = neg_inf();
    int64_t bi = INT64_MAX, nan_i = INT64_MAX;
    for (int e = threadIdx.x; e < nb; e += 32) {
        if (pi[e] < 0) continue;
        if (pv[e] != pv[e]) {
            if (pi[e] < nan_i) nan_i = pi[e];
        } else if (better(pv[e], pi[e], bv, bi)) {
            bv = pv[e];
            bi = pi[e];
        }
    }
    warp_argmax(bv, bi);
    nan_i = warp_min_i64(nan_i);
    if (nan_i != INT64_MAX) bi = nan_i;
    if (threadIdx.x == 0) {
        *out_idx = bi == INT64_MAX ? -1 : bi;
        *out_val = bi == INT64_MAX ? neg_inf() : v[bi];  // a NaN winner reports NaN
    }
}
}  // namespace tes

extern "C" int64_t tes_argmax_workspace_bytes(int64_t N) { return N < 0 ? -1 : 2 * align_up(1024 * 8, 256); }

extern "C" int tes_argmax_f64(const double* vals, int64_t N, double* out_val, int64_t* out_idx, void* ws,
                              int64_t ws_bytes, void* stream) {
    if (N < 0 || (N > 0 && !vals)) return -1;
    if (!out_val) return -3;
    if (!out_idx) return -4;
    if (!ws || ws_bytes < tes_argmax_workspace_bytes(N)) return -100;
    cudaStream_t st = (cudaStream_t)stream;
    double* pv = (double*)ws;
    int64_t* pi = (int64_t*)((char*)ws + align_up(1024 * 8, 256));
    int nb = (int)((N + 1023) / 1024);
    if (nb > 1024) nb = 1024;
    if (nb < 1) nb = 1;
    argmax_partial_kernel<<<nb, 256, 0, st>>>(vals, N, pv, pi);
    TES_LAUNCH_OK();
    argmax_final_kernel<<<1, 32, 0, st>>>(vals, pv, pi, nb, out_val, out_idx);
    TES_LAUNCH_OK();
    return 0;
}

extern "C" int tes_mma_arg_error_probe(const double* x, const double* w, const double* b, int64_t d, int n8,
                                       double* out_max, void* stream) {
    if (!x || !w || !b) return -1;
    if (d < 1 || d > 10) return -4;
    if (n8 < 1) return -5;
    if (!out_max) return -6;
    cudaStream_t st = (cudaStream_t)stream;
    switch (d) {
#define TES_CASE(DD)                                                        \
    case DD:                                                                \
        mma_arg_probe_kernel<DD><<<1, 32, 0, st>>>(x, w, b, n8, out_max);   \
        break;
        TES_CASE(1) TES_CASE(2) TES_CASE(3) TES_CASE(4) TES_CASE(5) TES_CASE(6) TES_CASE(7) TES_CASE(8) TES_CASE(9)
        TES_CASE(10)
#undef TES_CASE
    }
    TES_LAUNCH_OK();
    return 0;
}

extern "C" int tes_tc5_arg_error_probe(const double* x, const double* w, const double* b, int64_t d, int nt,
                                       double* out_max, void* stream) {
    if (!x || !w || !b) return -1;
    if (d < 1 || d > 10) return -4;
    if (nt < 1) return -5;
    if (!out_max) return -6;
    cudaStream_t st = (cudaStream_t)stream;
    const int kc = tc5_kc((int)d);
    const size_t smem = (size_t)kc * (128 + TC5_NF) * 16 + 64;
    TES_CUDA_OK(cudaMemsetAsync(out_max, 0, sizeof(double), st));
    TES_CUDA_OK(cudaFuncSetAttribute(tc5_arg_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc5_arg_probe_kernel<<<1, 128, smem, st>>>(x, w, b, (int)d, nt, out_max);
    TES_LAUNCH_OK();
    return 0;
}
