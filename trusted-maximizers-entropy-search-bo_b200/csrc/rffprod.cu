// rffprod.cu -- Stage A on PRODUCT-STRUCTURED candidate sets (every mesh of the discrete drivers: functions.get_meshgrid,
// bo_discrete.py; BASELINE configs C2 / C3).
//
// If the candidates are U x V in row-major order -- the coordinates in columns [s_lo, s_hi) depend on i / nv only (U) and
// the other columns on i % nv only (V) -- then
//     cos(w.x + b) = cos(alpha) cos(beta) - sin(alpha) sin(beta),   alpha = w_U.u + b,  beta = w_V.v
// and one function sample on the whole set is a GEMM
//     f_j(u, v) = scale * sum_f [theta_f cos alpha_f(u)] cos beta_f(v) + [-theta_f sin alpha_f(u)] sin beta_f(v)
//               = scale * (P_j Q_j^T)(u, v),        P_j (nu x 2F),  Q_j (nv x 2F)
// i.e. 2 nu nv 2F flops on the fp64 tensor pipe (DMMA) instead of nu nv F cosines on the MUFU: 4.3e12 flops instead of
// 1.07e12 cosines at C3 (N = 10^6 = 1000 x 1000, F = 1024, S = 1024), and only S (nu + nv) F sincos to build the operands.
//
// The GEMM values differ from the spec sequence of include/tes_spec.h only by rounding (both are fp64 evaluations of
// the same function: |difference| <= B_j = 2 u scale sum_f |theta_f| (4 (d + 2) A_j + 3F + 16) with u = 2^-53 and A_j
// the largest |argument| of the sample, derived in prod_exact_kernel).  Selection therefore works like the other
// screens: every warp of the GEMM keeps the kp = k + 8 best of its 32 x 32 values, the lists are merged, the kp
// survivors per sample are re-evaluated with the EXACT spec sequence and the final top-k is taken by (exact value
// desc, index asc).  A sample whose kp-th best GEMM value is not below its k-th best by more than 2 B_j (exact ties:
// duplicated candidates, constant samples) is FLAGGED and the caller recomputes with the general kernel, so the result
// is bit-identical to tes_rff_topk_f64 in every case.
#include <cuda_fp16.h>

#include <cstdlib>

#include "gemm.cuh"
#include "gp.cuh"

namespace tes {

constexpr int PROD_MAX_KP = 16;
__host__ __device__ constexpr int prod_rec(int d) { return (d + 3) & ~1; }   // same record as rff.cu

__global__ void prod_pack_kernel(const double* __restrict__ W, const double* __restrict__ b,
                                 const double* __restrict__ theta, int64_t S, int64_t F, int d, int w_shared, int rec,
                                 double* __restrict__ feat) {
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

// P[g][iu][2f] = theta cos(alpha), P[g][iu][2f+1] = -theta sin(alpha)      (threads: f fastest)
__global__ void prod_build_p_kernel(const double* __restrict__ cand, int64_t nu, int64_t nv, int d, int s_lo, int s_hi,
                                    const double* __restrict__ W, const double* __restrict__ b,
                                    const double* __restrict__ theta, int64_t j0, int64_t G, int64_t F, int w_shared,
                                    double* __restrict__ P) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= G * nu * F) return;
    const int64_t f = t % F, iu = (t / F) % nu, g = t / (F * nu);
    const int64_t j = j0 + g, jw = w_shared ? 0 : j;
    const double* w = W + (jw * F + f) * d;
    const double* u = cand + (iu * nv) * d;
    double a = b[jw * F + f];
    for (int k = s_lo; k < s_hi; ++k) a = fma(w[k], u[k], a);
    double sn, cs;
    sincos(a, &sn, &cs);
    const double th = theta[j * F + f];
    double2 o = make_double2(th * cs, -th * sn);
    *reinterpret_cast<double2*>(P + ((g * nu + iu) * F + f) * 2) = o;
}

// Qt[g][2f][iv] = cos(beta), Qt[g][2f+1][iv] = sin(beta)                    (threads: iv fastest)
__global__ void prod_build_q_kernel(const double* __restrict__ cand, int64_t nv, int d, int s_lo, int s_hi,
                                    const double* __restrict__ W, int64_t j0, int64_t G, int64_t F, int w_shared,
                                    double* __restrict__ Qt) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= G * F * nv) return;
    const int64_t iv = t % nv, f = (t / nv) % F, g = t / (nv * F);
    const int64_t jw = w_shared ? 0 : j0 + g;
    const double* w = W + (jw * F + f) * d;
    const double* v = cand + iv * d;
    double a = 0.0;
    for (int k = 0; k < s_lo; ++k) a = fma(w[k], v[k], a);
    for (int k = s_hi; k < d; ++k) a = fma(w[k], v[k], a);
    double sn, cs;
    sincos(a, &sn, &cs);
    Qt[((g * F + f) * 2) * nv + iv] = cs;
    Qt[((g * F + f) * 2 + 1) * nv + iv] = sn;
}

// C_g = P_g Qt_g on the DMMA core; instead of storing C every warp keeps the kp best of its 32 x 32 values.
// part lists: (list, g, kp) with list = (tile * 8 + warp).
__global__ void __launch_bounds__(GEMM_THREADS, 2)
    prod_gemm_topk_kernel(const double* __restrict__ P, const double* __restrict__ Qt, int64_t nu, int64_t nv, int64_t K2,
                          double scale, int64_t idx_offset, int kp, int64_t G, double* __restrict__ part_val,
                          int64_t* __restrict__ part_idx) {
    extern __shared__ __align__(16) unsigned char gemm_smem_raw[];
    GemmSmem& sm = *reinterpret_cast<GemmSmem*>(gemm_smem_raw);
    const int64_t g = blockIdx.z;
    const int64_t row0 = (int64_t)blockIdx.x * GEMM_BM;
    const int64_t col0 = (int64_t)blockIdx.y * GEMM_BN;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    APlain ap{P + g * nu * K2, K2, nu, K2};
    gemm_mainloop(ap, Qt + g * K2 * nv, nv, K2, nv, row0, col0, sm, acc);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int64_t list = ((int64_t)blockIdx.x * gridDim.y + blockIdx.y) * 8 + warp;
    double* ov = part_val + (list * G + g) * kp;
    int64_t* oi = part_idx + (list * G + g) * kp;
    // scale as the spec does (one rounding after the feature sum)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) acc[i][j][e] = __dmul_rn(scale, acc[i][j][e]);
    double pv = pos_inf();
    int64_t pi = -1;
    for (int r = 0; r < kp; ++r) {
        double bv = neg_inf();
        int64_t bi = INT64_MAX;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int64_t row = row0 + gemm_row_of(warp, lane, i);
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int64_t col = col0 + gemm_col_of(warp, lane, j, e);
                    const int64_t ci = row * nv + col + idx_offset;
                    const double v = acc[i][j][e];
                    if (row < nu && col < nv && better(pv, pi, v, ci) && better(v, ci, bv, bi)) {
                        bv = v;
                        bi = ci;
                    }
                }
        }
        warp_argmax(bv, bi);
        const bool found = (bi != INT64_MAX);
        if (lane == 0) {
            ov[r] = found ? bv : neg_inf();
            oi[r] = found ? bi : -1;
        }
        if (!found) {
            for (int q = r + 1; q < kp; ++q)
                if (lane == 0) {
                    ov[q] = neg_inf();
                    oi[q] = -1;
                }
            break;
        }
        pv = bv;
        pi = bi;
    }
}

// ---- 3xTF32 screen of the same GEMM on the legacy tensor-core path (mma.sync.m16n8k8.tf32) ------------------------
// p = p_hi + p_lo (two TF32 numbers), q likewise; p q ~ p_hi q_hi + p_hi q_lo + p_lo q_hi.  The operands are split once
// by the build kernels; the fp32 accumulators are flushed into fp64 sums (in shared memory) every PT_FLUSH k so the
// accumulation error stays bounded (see prod_exact_kernel for the bound).  Same tile shape and top-k epilogue as the
// fp64 kernel; operands padded with zeros to whole tiles so the loads need no bounds checks.
constexpr int PT_BM = 128, PT_BN = 64, PT_BK = 32, PT_THREADS = 256;
constexpr int PT_LDA = PT_BK + 4;    // floats: row stride 36 -> the 8 rows x 4 k of an A fragment hit 32 distinct banks
constexpr int PT_LDB = PT_BN + 8;    // floats: row stride 72 -> the 4 k x 8 n of a B fragment hit 32 distinct banks
constexpr int PT_FLUSH = 256;        // k per fp32 accumulation segment
constexpr int PT_NS = 3;             // cp.async stages
struct PtSmem {                      // 3 stages x (A hi/lo 128 x 32, B hi/lo 32 x 64) = 166 KB, filled by cp.async
    float ahi[PT_NS][PT_BM][PT_LDA];
    float alo[PT_NS][PT_BM][PT_LDA];
    float bhi[PT_NS][PT_BK][PT_LDB];
    float blo[PT_NS][PT_BK][PT_LDB];
};
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

__device__ __forceinline__ float to_tf32f(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// Phi/Plo[g][iu][2f], [2f+1] (row stride K2p floats, nup rows per sample); pads are pre-zeroed
__global__ void prod_build_p32_kernel(const double* __restrict__ cand, int64_t nu, int64_t nv, int d, int s_lo, int s_hi,
                                      const double* __restrict__ W, const double* __restrict__ b,
                                      const double* __restrict__ theta, int64_t j0, int64_t G, int64_t F, int w_shared,
                                      int64_t nup, int64_t K2p, float* __restrict__ Phi, float* __restrict__ Plo) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= G * nu * F) return;
    const int64_t f = t % F, iu = (t / F) % nu, g = t / (F * nu);
    const int64_t j = j0 + g, jw = w_shared ? 0 : j;
    const double* w = W + (jw * F + f) * d;
    const double* u = cand + (iu * nv) * d;
    double a = b[jw * F + f];
    for (int k = s_lo; k < s_hi; ++k) a = fma(w[k], u[k], a);
    // the operands are consumed as hi + lo TF32 pairs (2^-21 residual): reduce the argument in fp64, then sincosf
    // (2 ulp = 2^-22, covered by the 1.3 * 2^-20 term of the bound) instead of the ~4x dearer fp64 sincos
    float sn, cs;
    sincosf((float)(a - 6.283185307179586 * rint(a * 0.15915494309189535)), &sn, &cs);
    const double th = theta[j * F + f];
    const float p0 = (float)(th * (double)cs), p1 = (float)(-th * (double)sn);
    const float h0 = to_tf32f(p0), h1 = to_tf32f(p1);
    const int64_t o = (g * nup + iu) * K2p + 2 * f;
    *reinterpret_cast<float2*>(Phi + o) = make_float2(h0, h1);
    *reinterpret_cast<float2*>(Plo + o) = make_float2(to_tf32f(p0 - h0), to_tf32f(p1 - h1));
}

// Qhi/Qlo[g][2f][iv], [2f+1][iv] (row stride nvp floats, K2p rows per sample)
__global__ void prod_build_q32_kernel(const double* __restrict__ cand, int64_t nv, int d, int s_lo, int s_hi,
                                      const double* __restrict__ W, int64_t j0, int64_t G, int64_t F, int w_shared,
                                      int64_t nvp, int64_t K2p, float* __restrict__ Qhi, float* __restrict__ Qlo) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= G * F * nv) return;
    const int64_t iv = t % nv, f = (t / nv) % F, g = t / (nv * F);
    const int64_t jw = w_shared ? 0 : j0 + g;
    const double* w = W + (jw * F + f) * d;
    const double* v = cand + iv * d;
    double a = 0.0;
    for (int k = 0; k < s_lo; ++k) a = fma(w[k], v[k], a);
    for (int k = s_hi; k < d; ++k) a = fma(w[k], v[k], a);
    float sn, cs;
    sincosf((float)(a - 6.283185307179586 * rint(a * 0.15915494309189535)), &sn, &cs);
    const float c0 = cs, c1 = sn;
    const float h0 = to_tf32f(c0), h1 = to_tf32f(c1);
    const int64_t o = (g * K2p + 2 * f) * nvp + iv;
    Qhi[o] = h0;
    Qhi[o + nvp] = h1;
    Qlo[o] = to_tf32f(c0 - h0);
    Qlo[o + nvp] = to_tf32f(c1 - h1);
}

__global__ void __launch_bounds__(PT_THREADS, 1)
    prod_gemm_topk_tf32_kernel(const float* __restrict__ Phi, const float* __restrict__ Plo, const float* __restrict__ Qhi,
                               const float* __restrict__ Qlo, int64_t nu, int64_t nv, int64_t nup, int64_t nvp,
                               int64_t K2p, double scale, int64_t idx_offset, int kp, int64_t G,
                               double* __restrict__ part_val, int64_t* __restrict__ part_idx) {
    extern __shared__ __align__(16) unsigned char pt_smem_raw[];
    PtSmem& sm = *reinterpret_cast<PtSmem*>(pt_smem_raw);
    const int64_t g = blockIdx.z;
    const int64_t row0 = (int64_t)blockIdx.x * PT_BM, col0 = (int64_t)blockIdx.y * PT_BN;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int gq = lane >> 2, tq = lane & 3;
    const int wr = (warp >> 1) * 32, wc = (warp & 1) * 32;       // warp tile origin inside the CTA tile
    const float* pa_hi = Phi + (g * nup + row0) * K2p;
    const float* pa_lo = Plo + (g * nup + row0) * K2p;
    const float* pb_hi = Qhi + g * K2p * nvp + col0;
    const float* pb_lo = Qlo + g * K2p * nvp + col0;
    // global -> shared with cp.async (16-byte chunks, no staging registers): A tile 128 x 32 floats = 1024 chunks per
    // array (4 per thread), B tile 32 x 64 = 512 chunks per array (2 per thread)
    auto issue = [&](int st, int64_t k0) {
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            const int c = t + 256 * h, r = c >> 3, kc = (c & 7) * 4;
            cp_async16(&sm.ahi[st][r][kc], pa_hi + (int64_t)r * K2p + k0 + kc);
            cp_async16(&sm.alo[st][r][kc], pa_lo + (int64_t)r * K2p + k0 + kc);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int c = t + 256 * h, r = c >> 4, cc = (c & 15) * 4;
            cp_async16(&sm.bhi[st][r][cc], pb_hi + (k0 + r) * nvp + cc);
            cp_async16(&sm.blo[st][r][cc], pb_lo + (k0 + r) * nvp + cc);
        }
        cp_async_commit();
    };
    float acc[2][4][4];
    double sum[2][4][4];             // fp64 running sums (registers)
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                acc[i][j][e] = 0.0f;
                sum[i][j][e] = 0.0;
            }
    auto flush = [&]() {
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    sum[i][j][e] += (double)acc[i][j][e];
                    acc[i][j][e] = 0.0f;
                }
    };
    const int64_t nk = K2p / PT_BK;
    issue(0, 0);
    if (nk > 1) issue(1, PT_BK);
    for (int64_t kt = 0; kt < nk; ++kt) {
        const int cur = (int)(kt % PT_NS);
        if (kt + 1 < nk)
            cp_async_wait<1>();                 // slice kt has landed (slice kt + 1 may still be in flight)
        else
            cp_async_wait<0>();
        __syncthreads();                        // ... for every thread; and everyone is done reading slice kt - 1
        if (kt + 2 < nk) issue((int)((kt + 2) % PT_NS), (kt + 2) * PT_BK);   // into the stage slice kt - 1 used
        // fragments of k8-step s + 1 are fetched while the 24 mma of step s issue (register double buffer)
        uint32_t ah[2][2][4], al[2][2][4], bh[2][4][2], bl[2][4][2];
        auto frag = [&](int fb, int ks) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int r = wr + 16 * i + gq;
                ah[fb][i][0] = __float_as_uint(sm.ahi[cur][r][ks + tq]);
                ah[fb][i][1] = __float_as_uint(sm.ahi[cur][r + 8][ks + tq]);
                ah[fb][i][2] = __float_as_uint(sm.ahi[cur][r][ks + tq + 4]);
                ah[fb][i][3] = __float_as_uint(sm.ahi[cur][r + 8][ks + tq + 4]);
                al[fb][i][0] = __float_as_uint(sm.alo[cur][r][ks + tq]);
                al[fb][i][1] = __float_as_uint(sm.alo[cur][r + 8][ks + tq]);
                al[fb][i][2] = __float_as_uint(sm.alo[cur][r][ks + tq + 4]);
                al[fb][i][3] = __float_as_uint(sm.alo[cur][r + 8][ks + tq + 4]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = wc + 8 * j + gq;
                bh[fb][j][0] = __float_as_uint(sm.bhi[cur][ks + tq][c]);
                bh[fb][j][1] = __float_as_uint(sm.bhi[cur][ks + tq + 4][c]);
                bl[fb][j][0] = __float_as_uint(sm.blo[cur][ks + tq][c]);
                bl[fb][j][1] = __float_as_uint(sm.blo[cur][ks + tq + 4][c]);
            }
        };
        frag(0, 0);
#pragma unroll
        for (int s8 = 0; s8 < PT_BK / 8; ++s8) {
            const int fb = s8 & 1;
            if (s8 + 1 < PT_BK / 8) frag(fb ^ 1, (s8 + 1) * 8);
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) mma_tf32(acc[i][j], al[fb][i], bh[fb][j]);   // small terms first
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) mma_tf32(acc[i][j], ah[fb][i], bl[fb][j]);
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) mma_tf32(acc[i][j], ah[fb][i], bh[fb][j]);
        }
        if (((kt + 1) * PT_BK) % PT_FLUSH == 0) flush();
    }
    flush();
    // values in the accumulator layout: acc[i][j][e]: row wr + 16 i + gq + 8 (e >> 1), col wc + 8 j + 2 tq + (e & 1)
    const int64_t list = ((int64_t)blockIdx.x * gridDim.y + blockIdx.y) * 8 + warp;
    double* ov = part_val + (list * G + g) * kp;
    int64_t* oi = part_idx + (list * G + g) * kp;
    double pv = pos_inf();
    int64_t pi = -1;
    for (int r = 0; r < kp; ++r) {
        double bv = neg_inf();
        int64_t bi = INT64_MAX;
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int64_t row = row0 + wr + 16 * i + gq + 8 * (e >> 1);
                    const int64_t col = col0 + wc + 8 * j + 2 * tq + (e & 1);
                    const int64_t ci = row * nv + col + idx_offset;
                    const double v = __dmul_rn(scale, sum[i][j][e]);
                    if (row < nu && col < nv && better(pv, pi, v, ci) && better(v, ci, bv, bi)) {
                        bv = v;
                        bi = ci;
                    }
                }
        warp_argmax(bv, bi);
        const bool found = (bi != INT64_MAX);
        if (lane == 0) {
            ov[r] = found ? bv : neg_inf();
            oi[r] = found ? bi : -1;
        }
        if (!found) {
            for (int q = r + 1; q < kp; ++q)
                if (lane == 0) {
                    ov[q] = neg_inf();
                    oi[q] = -1;
                }
            break;
        }
        pv = bv;
        pi = bi;
    }
}

// ---- 3xFP16 screen (mma.sync.m16n8k16.f16, fp32 accumulate): same precision class as 3xTF32 (fp16 carries 11
// significant bits like TF32; hi + lo = 22 bits) at twice the k per instruction.  Both operands are stored K-contiguous
// (P: (G, nup, K2p), Q: (G, nvp, K2p) halves) because the k16 fragments pack two consecutive k per register.
constexpr int PH_BK = 64;            // halves per slice = 4 k16 steps
constexpr int PH_LD = PH_BK + 8;     // halves: row stride 36 words -> the 8 rows x 4 words of a fragment hit 32 banks
struct PhSmem {                      // 3 stages x (A hi/lo 128 x 64, B hi/lo 64 x 64 halves) = 162 KB
    __half ahi[PT_NS][PT_BM][PH_LD];
    __half alo[PT_NS][PT_BM][PH_LD];
    __half bhi[PT_NS][PT_BN][PH_LD];
    __half blo[PT_NS][PT_BN][PH_LD];
};
__device__ __forceinline__ void mma_f16(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void split_f16(float x, __half& hi, __half& lo) {
    hi = __float2half_rn(x);
    lo = __float2half_rn(x - __half2float(hi));
}

// Phi/Plo[g][iu][2f..2f+1] halves (row stride K2p); which = 0: the U side (theta cos a, -theta sin a),
// which = 1: the V side (cos b, sin b) written to Qhi/Qlo[g][iv][2f..2f+1]
constexpr int PB_ROWS = 8;           // rows per thread of the operand build (the feature's w is loaded once for them)
__global__ void prod_build_f16_kernel(const double* __restrict__ cand, int64_t nrow, int64_t row_stride, int d, int s_lo,
                                      int s_hi, int which, const double* __restrict__ W, const double* __restrict__ b,
                                      const double* __restrict__ theta, int64_t j0, int64_t G, int64_t F, int w_shared,
                                      int64_t nrowp, int64_t K2p, const double* __restrict__ tmax,
                                      __half* __restrict__ Hi, __half* __restrict__ Lo) {
    const int64_t nrb = (nrow + PB_ROWS - 1) / PB_ROWS;
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= G * nrb * F) return;
    const int64_t f = t % F, rb = (t / F) % nrb, g = t / (F * nrb);
    const int64_t j = j0 + g, jw = w_shared ? 0 : j;
    const double* w = W + (jw * F + f) * d;
    double wk[16];
    int nk_ = 0;
    if (which == 0) {
        for (int k = s_lo; k < s_hi; ++k) wk[nk_++] = w[k];
    } else {
        for (int k = 0; k < s_lo; ++k) wk[nk_++] = w[k];
        for (int k = s_hi; k < d; ++k) wk[nk_++] = w[k];
    }
    const double a0 = which == 0 ? b[jw * F + f] : 0.0;
    double th = 1.0;
    if (which == 0) {
        const double tm = tmax[j];                       // fp16 has a 5-bit exponent: theta is normalised per sample
        th = tm > 0.0 ? theta[j * F + f] / tm : 0.0;
    }
    for (int r = 0; r < PB_ROWS; ++r) {
        const int64_t ir = rb * PB_ROWS + r;
        if (ir >= nrow) break;
        const double* x = cand + (ir * row_stride) * d;
        double a = a0;
        int q = 0;
        if (which == 0) {
            for (int k = s_lo; k < s_hi; ++k) a = fma(wk[q++], x[k], a);
        } else {
            for (int k = 0; k < s_lo; ++k) a = fma(wk[q++], x[k], a);
            for (int k = s_hi; k < d; ++k) a = fma(wk[q++], x[k], a);
        }
        float sn, cs;
        sincosf((float)(a - 6.283185307179586 * rint(a * 0.15915494309189535)), &sn, &cs);
        float p0 = cs, p1 = sn;
        if (which == 0) {
            p0 = (float)(th * (double)cs);
            p1 = (float)(-th * (double)sn);
        }
        __half h0, l0, h1, l1;
        split_f16(p0, h0, l0);
        split_f16(p1, h1, l1);
        const int64_t o = (g * nrowp + ir) * K2p + 2 * f;
        *reinterpret_cast<__half2*>(Hi + o) = __halves2half2(h0, h1);
        *reinterpret_cast<__half2*>(Lo + o) = __halves2half2(l0, l1);
    }
}

__global__ void __launch_bounds__(PT_THREADS, 1)
    prod_gemm_topk_f16_kernel(const __half* __restrict__ Phi, const __half* __restrict__ Plo,
                              const __half* __restrict__ Qhi, const __half* __restrict__ Qlo, int64_t nu, int64_t nv,
                              int64_t nup, int64_t nvp, int64_t K2p, double scale0, const double* __restrict__ tmax,
                              int64_t idx_offset, int kp, int64_t G, double* __restrict__ part_val,
                              int64_t* __restrict__ part_idx) {
    extern __shared__ __align__(16) unsigned char pt_smem_raw[];
    const double scale = scale0 * tmax[blockIdx.z];
    PhSmem& sm = *reinterpret_cast<PhSmem*>(pt_smem_raw);
    const int64_t g = blockIdx.z;
    const int64_t row0 = (int64_t)blockIdx.x * PT_BM, col0 = (int64_t)blockIdx.y * PT_BN;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int gq = lane >> 2, tq = lane & 3;
    const int wr = (warp >> 1) * 32, wc = (warp & 1) * 32;
    const __half* pa_hi = Phi + (g * nup + row0) * K2p;
    const __half* pa_lo = Plo + (g * nup + row0) * K2p;
    const __half* pb_hi = Qhi + (g * nvp + col0) * K2p;
    const __half* pb_lo = Qlo + (g * nvp + col0) * K2p;
    // cp.async, 16-byte chunks of 8 halves: A tile 128 rows x 8 chunks = 1024 per array (4 per thread), B tile 64 rows
    // x 8 chunks = 512 per array (2 per thread)
    auto issue = [&](int st, int64_t k0) {
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            const int c = t + 256 * h, r = c >> 3, kc = (c & 7) * 8;
            cp_async16(&sm.ahi[st][r][kc], pa_hi + (int64_t)r * K2p + k0 + kc);
            cp_async16(&sm.alo[st][r][kc], pa_lo + (int64_t)r * K2p + k0 + kc);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int c = t + 256 * h, r = c >> 3, kc = (c & 7) * 8;
            cp_async16(&sm.bhi[st][r][kc], pb_hi + (int64_t)r * K2p + k0 + kc);
            cp_async16(&sm.blo[st][r][kc], pb_lo + (int64_t)r * K2p + k0 + kc);
        }
        cp_async_commit();
    };
    float acc[2][4][4];
    double sum[2][4][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                acc[i][j][e] = 0.0f;
                sum[i][j][e] = 0.0;
            }
    auto flush = [&]() {
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    sum[i][j][e] += (double)acc[i][j][e];
                    acc[i][j][e] = 0.0f;
                }
    };
    const int64_t nk = K2p / PH_BK;
    issue(0, 0);
    if (nk > 1) issue(1, PH_BK);
    for (int64_t kt = 0; kt < nk; ++kt) {
        const int cur = (int)(kt % PT_NS);
        if (kt + 1 < nk)
            cp_async_wait<1>();
        else
            cp_async_wait<0>();
        __syncthreads();
        if (kt + 2 < nk) issue((int)((kt + 2) % PT_NS), (kt + 2) * PH_BK);
        uint32_t ah[2][2][4], al[2][2][4], bh[2][4][2], bl[2][4][2];
        auto ld32 = [](const __half* p) { return *reinterpret_cast<const uint32_t*>(p); };
        auto frag = [&](int fb, int ks) {   // ks in halves; register r of a fragment = 2 consecutive k
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int r = wr + 16 * i + gq;
                ah[fb][i][0] = ld32(&sm.ahi[cur][r][ks + 2 * tq]);
                ah[fb][i][1] = ld32(&sm.ahi[cur][r + 8][ks + 2 * tq]);
                ah[fb][i][2] = ld32(&sm.ahi[cur][r][ks + 2 * tq + 8]);
                ah[fb][i][3] = ld32(&sm.ahi[cur][r + 8][ks + 2 * tq + 8]);
                al[fb][i][0] = ld32(&sm.alo[cur][r][ks + 2 * tq]);
                al[fb][i][1] = ld32(&sm.alo[cur][r + 8][ks + 2 * tq]);
                al[fb][i][2] = ld32(&sm.alo[cur][r][ks + 2 * tq + 8]);
                al[fb][i][3] = ld32(&sm.alo[cur][r + 8][ks + 2 * tq + 8]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = wc + 8 * j + gq;
                bh[fb][j][0] = ld32(&sm.bhi[cur][c][ks + 2 * tq]);
                bh[fb][j][1] = ld32(&sm.bhi[cur][c][ks + 2 * tq + 8]);
                bl[fb][j][0] = ld32(&sm.blo[cur][c][ks + 2 * tq]);
                bl[fb][j][1] = ld32(&sm.blo[cur][c][ks + 2 * tq + 8]);
            }
        };
        frag(0, 0);
#pragma unroll
        for (int s16 = 0; s16 < PH_BK / 16; ++s16) {
            const int fb = s16 & 1;
            if (s16 + 1 < PH_BK / 16) frag(fb ^ 1, (s16 + 1) * 16);
            // term by term over the 8 accumulator tiles: consecutive mma never depend on each other (three back-to-back
            // mma on one accumulator serialise on its latency: ncu `wait` 3.0 per issue, tensor pipe 35 % busy)
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) mma_f16(acc[i][j], al[fb][i], bh[fb][j]);
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) mma_f16(acc[i][j], ah[fb][i], bl[fb][j]);
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) mma_f16(acc[i][j], ah[fb][i], bh[fb][j]);
        }
        if (((kt + 1) * PH_BK) % PT_FLUSH == 0) flush();
    }
    flush();
    const int64_t list = ((int64_t)blockIdx.x * gridDim.y + blockIdx.y) * 8 + warp;
    double* ov = part_val + (list * G + g) * kp;
    int64_t* oi = part_idx + (list * G + g) * kp;
    double pv = pos_inf();
    int64_t pi = -1;
    for (int r = 0; r < kp; ++r) {
        double bv = neg_inf();
        int64_t bi = INT64_MAX;
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int64_t row = row0 + wr + 16 * i + gq + 8 * (e >> 1);
                    const int64_t col = col0 + wc + 8 * j + 2 * tq + (e & 1);
                    const int64_t ci = row * nv + col + idx_offset;
                    const double v = __dmul_rn(scale, sum[i][j][e]);
                    if (row < nu && col < nv && better(pv, pi, v, ci) && better(v, ci, bv, bi)) {
                        bv = v;
                        bi = ci;
                    }
                }
        warp_argmax(bv, bi);
        const bool found = (bi != INT64_MAX);
        if (lane == 0) {
            ov[r] = found ? bv : neg_inf();
            oi[r] = found ? bi : -1;
        }
        if (!found) {
            for (int q = r + 1; q < kp; ++q)
                if (lane == 0) {
                    ov[q] = neg_inf();
                    oi[q] = -1;
                }
            break;
        }
        pv = bv;
        pi = bi;
    }
}

// ---- the same 3xFP16 GEMM on the 5th-generation tensor cores: tcgen05.mma kind::f16 + TMEM + bulk async copies ------
// Operands are laid out by the build kernel in the shared-memory order of the MMA (K-major, no swizzle: core matrix =
// 8 rows x 16 bytes): per (sample, 128-row tile, stage of 64 k) one contiguous 32 KB block
//     [hi | lo][8 chunks of 8 halves][128 rows][8 halves]
// so a stage is filled by two 1-D bulk copies (A block, B block) completing on an mbarrier.
//   warp 0 (one lane): producer -- 3-stage ring, 64 KB per stage;
//   warp 1 (one lane): issues 12 tcgen05.mma (M 128 x N 128 x K 16: 4 k16 steps x {lo*hi, hi*lo, hi*hi}) per stage into
//                      one of two 128-column TMEM accumulators; a SEGMENT = 4 stages = 256 k; tcgen05.commit releases
//                      the stage to the producer and, at the end of a segment, hands the accumulator to the epilogue;
//   warps 2..17      : epilogue -- thread = one row (TMEM lane), 32 columns: tcgen05.ld, fp32 -> fp64 running sums
//                      (the segment flush that bounds the fp32 accumulation error), then the per-warp top-kp.
constexpr int TC_EPI_WARPS = 16, TC_THREADS = 64 + 32 * TC_EPI_WARPS, TC_NS = 3, TC_BK = 64, TC_SEG = 4;
constexpr uint32_t TC_BLOCK_BYTES = 2 * 8 * 128 * 16;          // one operand block of a stage (hi + lo) = 32 KB
constexpr uint32_t TC_STAGE_BYTES = 2 * TC_BLOCK_BYTES;        // A block + B block
#ifndef TC_COPY_SPLIT
#define TC_COPY_SPLIT 1
#endif

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// K-major, no swizzle: LBO = bytes between core matrices adjacent in K, SBO = between core matrices adjacent in M/N
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
}
// D = f32 (bit 4), A = B = f16 (0), K-major, N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr uint32_t tc_idesc_f16(int m, int n) {
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
#define TC_LD32(v, taddr)                                                                                             \
    asm volatile(                                                                                                     \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,"  \
        "%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                                                \
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),             \
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),       \
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),     \
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])      \
        : "r"(taddr))

// operand build in MMA order: block index ((g * ntile + tile) * nst + st), inside [hl][chunk][row][8].
// One CTA per (sample, 128-row tile): the tile's coordinates are staged in shared memory, thread (row, half) walks the
// 8-k chunks (= 4 features: weights read with block-uniform addresses), row-consecutive threads write consecutive
// 16-byte slots of the MMA layout.
constexpr int PB_WS_DOUBLES = 1536;
// NC > 0: the number of coordinate columns of this side is a compile-time constant (the dot product and the record
// offsets unroll, the row's coordinates live in registers); NC = 0: generic.  ncu of the generic kernel at C3: issue
// slots 80 % busy, 56 thread-instructions per feature of which half were index arithmetic, remainder branches of the
// d-loop and per-element masking (profiles/r02_ncu_prod_build_tc.md).  Rows >= nrow and features >= F are NOT written:
// the callers zero the operand arrays once (cudaMemsetAsync) and those positions are padding in every group.
template <int NC>
__global__ void __launch_bounds__(256)
    prod_build_tc_kernel_t(const double* __restrict__ cand, int64_t nrow, int64_t row_stride, int d, int s_lo, int s_hi,
                           int which, const double* __restrict__ W, const double* __restrict__ b,
                           const double* __restrict__ theta, int64_t j0, int64_t F, int w_shared, int64_t ntile,
                           int64_t nst, const double* __restrict__ tmax, __half* __restrict__ out, int chunk_lo,
                           int TR) {
    // chunk_lo > 0: only the 8-k chunks chunk_lo .. 7 of every 64-k stage are built (prod_gemm_gen_kernel generates
    // chunks 0 .. chunk_lo - 1 on the SM).  TR = rows per operand tile: 128, or 64 for the B operand of the CTA-pair
    // kernel (each CTA of a pair stages half of the 128 columns of a tile).
    __shared__ double xs[16][129];                   // [coordinate slot][row]
    __shared__ double ws[PB_WS_DOUBLES];             // [feature of the block][w (ncol), b, theta]
    const int64_t tile = blockIdx.x, g = blockIdx.y;
    const int r = threadIdx.x % TR, half = threadIdx.x / TR, nhalf = (int)blockDim.x / TR;   // 256 or 128 threads
    const int64_t ir = tile * TR + r;
    const int64_t j = j0 + g, jw = w_shared ? 0 : j;
    // coordinate slots: the columns this side uses, in order (arithmetic, not a table: a runtime-indexed array would
    // live in local memory and put an LDL in front of every DFMA)
    const int ncol = NC > 0 ? NC : (which == 0 ? s_hi - s_lo : d - (s_hi - s_lo));
    auto col_of = [&](int c) { return which == 0 ? s_lo + c : (c < s_lo ? c : c + (s_hi - s_lo)); };
    for (int e = threadIdx.x; e < ncol * TR; e += blockDim.x) {
        const int c = e / TR, rr = e % TR;
        const int64_t gr = tile * TR + rr;
        xs[c][rr] = gr < nrow ? cand[(gr * row_stride) * d + col_of(c)] : 0.0;
    }
    __syncthreads();
    double xr[NC > 0 ? NC : 1];
    if (NC > 0) {
#pragma unroll
        for (int c = 0; c < NC; ++c) xr[c] = xs[c][r];
    }
    const bool okrow = ir < nrow;
    const double tm = which == 0 ? tmax[j] : 1.0;
    const double inv_tm = tm > 0.0 ? 1.0 / tm : 0.0;
    const int64_t nkc = nst * (TC_BK / 8);
    const int64_t kc_lo = nkc * blockIdx.z / gridDim.z, kc_hi = nkc * (blockIdx.z + 1) / gridDim.z;
    // feature records [w (ncol), b, theta / tmax] of a block of chunks are staged in shared memory by one coalesced
    // pass, so the chunk loop reads them as broadcast LDS instead of ~20 block-uniform LDGs from L2 per chunk in front
    // of every dependent chain
    const int rs = ncol + 2;
    const int cpb = (PB_WS_DOUBLES / rs) / 4 < 32 ? (PB_WS_DOUBLES / rs) / 4 : 32;      // chunks per block
    // this thread's 16-byte slot inside a block: [hl][chunk][row][8 halves]; blocks of 2 * 8 * TR * 8 halves
    const int64_t blk_halves = (int64_t)(2 * 8 * TR * 8);
    __half* const tile_out = out + ((g * ntile + tile) * nst) * blk_halves + (int64_t)r * 8;
    const int lo_off = 8 * TR * 8;
    for (int64_t cb = kc_lo; cb < kc_hi; cb += cpb) {
        const int64_t ce = cb + cpb < kc_hi ? cb + cpb : kc_hi;
        const int nf = (int)(ce - cb) * 4;
        __syncthreads();
        for (int e = threadIdx.x; e < nf * rs; e += blockDim.x) {
            const int fl = e / rs, c = e % rs;
            const int64_t f = cb * 4 + fl, fc = f < F ? f : F - 1;
            double v;
            if (c < ncol) v = __ldg(W + (jw * F + fc) * d + col_of(c));
            else if (c == ncol) v = which == 0 ? __ldg(b + jw * F + fc) : 0.0;
            else v = __hiloint2double(0, __float_as_int(which == 0 ? (float)(__ldg(theta + j * F + fc) * inv_tm) : 0.0f));
            ws[e] = v;   // theta / tmax as a float (bit pattern in the low word): the P operand is an fp32 product
        }
        __syncthreads();
        if (!okrow) continue;
        const int nck = (int)(ce - cb);
        for (int lc = half; lc < nck; lc += nhalf) {
            const int64_t kc = cb + lc;
            const int st = (int)(kc >> 3), chunk = (int)(kc & 7);      // TC_BK / 8 = 8 chunks per stage
            if (chunk < chunk_lo) continue;
            const int nfeat = F - kc * 4 >= 4 ? 4 : (int)(F - kc * 4);  // features of this chunk that exist (<= 0: padding)
            if (nfeat <= 0) continue;
            __align__(16) __half hi[8], lo[8];
            // the four features of the chunk are evaluated together, branch-free, so their dependent chains -- the fp64
            // dot and reduction, the MUFU pair, the splits -- overlap instead of running back to back
            const double* rec = ws + lc * 4 * rs;
            double a[4];
            float th[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                a[q] = rec[q * rs + ncol];
                th[q] = __int_as_float(__double2loint(rec[q * rs + ncol + 1]));
            }
            if (NC > 0) {
#pragma unroll
                for (int c = 0; c < NC; ++c)
#pragma unroll
                    for (int q = 0; q < 4; ++q) a[q] = fma(rec[q * rs + c], xr[c], a[q]);
            } else {
                for (int c = 0; c < ncol; ++c) {
                    const double x = xs[c][r];
#pragma unroll
                    for (int q = 0; q < 4; ++q) a[q] = fma(rec[q * rs + c], x, a[q]);
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                // argument reduced to [-pi, pi] in fp64, then the MUFU pair (__sincosf: abs error <= 2^-21.2 there):
                // the operand error budget of the screen (prod_exact_kernel) carries 2^-21 for it
                // rint by the 1.5 * 2^52 shift (exact for |t| < 2^51, round-to-nearest-even like rint): two DADDs on
                // the fp64 pipe instead of an FRND on the conversion (XU) pipe (MUFU.SIN / MUFU.COS and the
                // fp64 -> fp32 conversion live there too).  theta / tmax is applied as an fp32 product of floats (no
                // fp32 <-> fp64 round trips: 4 fewer XU conversions per feature); its two roundings are the
                // 0.25 * 2^-20 added to the per-term budget in prod_exact_kernel
                float sn, cs;
                const double tt = a[q] * 0.15915494309189535;
                // (|t| >= 2^51 only loses the reduction: the value bound B of prod_exact_kernel grows with max |a| and
                // flags such a sample for the general kernel long before)
                const double red = __dadd_rn(__dadd_rn(tt, 6755399441055744.0), -6755399441055744.0);
                __sincosf((float)(a[q] - 6.283185307179586 * red), &sn, &cs);
                float p0 = cs, p1 = sn;
                if (which == 0) {
                    p0 = __fmul_rn(th[q], cs);
                    p1 = __fmul_rn(-th[q], sn);
                }
                if (q >= nfeat) p0 = p1 = 0.0f;          // F not a multiple of 4: the tail of the last chunk
                split_f16(p0, hi[2 * q], lo[2 * q]);
                split_f16(p1, hi[2 * q + 1], lo[2 * q + 1]);
            }
            __half* blk = tile_out + (int64_t)st * blk_halves + chunk * TR * 8;
            *reinterpret_cast<uint4*>(blk) = *reinterpret_cast<const uint4*>(hi);
            *reinterpret_cast<uint4*>(blk + lo_off) = *reinterpret_cast<const uint4*>(lo);
        }
    }
}
template <int NC>
static void prod_build_tc_launch_t(dim3 grid, int threads, cudaStream_t st, const double* cand, int64_t nrow,
                                   int64_t row_stride, int d, int s_lo, int s_hi, int which, const double* W, const double* b,
                                   const double* theta, int64_t j0, int64_t F, int w_shared, int64_t ntile, int64_t nst,
                                   const double* tmax, __half* out, int chunk_lo, int TR) {
    static const bool once = [] {     // the whole 228 KB as shared memory: a build CTA fits beside a persistent GEMM CTA
        cudaFuncSetAttribute(prod_build_tc_kernel_t<NC>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        return true;
    }();
    (void)once;
    prod_build_tc_kernel_t<NC><<<grid, threads, 0, st>>>(cand, nrow, row_stride, d, s_lo, s_hi, which, W, b, theta, j0, F,
                                                         w_shared, ntile, nst, tmax, out, chunk_lo, TR);
}
static void launch_prod_build_tc(dim3 grid, int threads, cudaStream_t st, const double* cand, int64_t nrow,
                                 int64_t row_stride, int d, int s_lo, int s_hi, int which, const double* W, const double* b,
                                 const double* theta, int64_t j0, int64_t F, int w_shared, int64_t ntile, int64_t nst,
                                 const double* tmax, __half* out, int chunk_lo, int TR) {
    const int ncol = which == 0 ? s_hi - s_lo : d - (s_hi - s_lo);
#define PB_ARGS grid, threads, st, cand, nrow, row_stride, d, s_lo, s_hi, which, W, b, theta, j0, F, w_shared, ntile, nst, tmax, out, chunk_lo, TR
    switch (ncol) {
        case 1: prod_build_tc_launch_t<1>(PB_ARGS); break;
        case 2: prod_build_tc_launch_t<2>(PB_ARGS); break;
        case 3: prod_build_tc_launch_t<3>(PB_ARGS); break;
        case 4: prod_build_tc_launch_t<4>(PB_ARGS); break;
        case 5: prod_build_tc_launch_t<5>(PB_ARGS); break;
        case 6: prod_build_tc_launch_t<6>(PB_ARGS); break;
        default: prod_build_tc_launch_t<0>(PB_ARGS); break;
    }
#undef PB_ARGS
}

static int tc_sm_count() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1)
            sms = 148;
    }
    return sms;
}

// HALF = 0: a stage is a whole 64-k operand block pair (64 KB), NS stages.  HALF = 1: a stage is HALF a block pair (32 k,
// 32 KB: [A hi | A lo | B hi | B lo], 8 KB each), so that 7 stages = 224 KB fill the shared memory instead of
// 3 x 64 = 192 KB -- the kernel is bound by the latency of the operand pipeline, i.e. by the bytes in flight.
template <int HALF, int NS>
__global__ void __launch_bounds__(TC_THREADS, 1)
    prod_gemm_topk_tc_kernel(const __half* __restrict__ Pt, const __half* __restrict__ Qt, int64_t nu, int64_t nv,
                             int64_t ntu, int64_t ntv, int64_t nst, double scale0, const double* __restrict__ tmax,
                             int64_t idx_offset, int kp, int64_t G, double* __restrict__ part_val,
                             int64_t* __restrict__ part_idx, double* __restrict__ store_out, int64_t ldo,
                             const double* __restrict__ rowscale, const double* __restrict__ colscale) {
    // store_out != nullptr: instead of the top-k epilogue, write rowscale[row] * colscale[col] * sum to
    // store_out[row * ldo + col] (the quadratic-form screen of the TES_ep criterion)
    extern __shared__ __align__(1024) unsigned char tc_smem_raw[];
    constexpr uint32_t STAGE = HALF ? TC_STAGE_BYTES / 2 : TC_STAGE_BYTES;          // bytes per stage
    constexpr uint32_t PART = HALF ? TC_BLOCK_BYTES / 4 : TC_BLOCK_BYTES / 2;       // bytes of the hi (= of the lo) part
    constexpr int SUB = HALF ? 2 : 1;                                                // stages per 64-k block pair
    unsigned char* stages = tc_smem_raw;                                            // [NS][A hi | A lo | B hi | B lo]
    uint64_t* bars = reinterpret_cast<uint64_t*>(tc_smem_raw + NS * STAGE);
    uint64_t* full_b = bars;                 // [NS] producer -> MMA
    uint64_t* empty_b = bars + NS;           // [NS] MMA (commit) -> producer
    uint64_t* acc_full = bars + 2 * NS;      // [2] MMA (commit) -> epilogue
    uint64_t* acc_empty = acc_full + 2;      // [2] epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // persistent: CTA c takes tiles c, c + gridDim.x, ... of the (rt fastest, then ct, then sample g) order; the barrier
    // phases run on across tiles, so the next tile's loads and MMAs start while the epilogue warps still rank this one
    const int64_t ntot = ntu * ntv * G;
    const int64_t ntile = (int64_t)blockIdx.x < ntot ? (ntot - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int64_t nseg = nst / TC_SEG;
    if (tid == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(&full_b[s], 1);
            mbar_init(&empty_b[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&acc_full[a], 1);
            mbar_init(&acc_empty[a], TC_EPI_WARPS);
        }
        mbar_fence_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int64_t gst = 0;
            for (int64_t it = 0; it < ntile; ++it) {
            const int64_t t = blockIdx.x + it * gridDim.x;
            const int64_t rt = t % ntu, ct = (t / ntu) % ntv, g = t / (ntu * ntv);
            const unsigned char* srcA = reinterpret_cast<const unsigned char*>(Pt) + ((g * ntu + rt) * nst) * (int64_t)TC_BLOCK_BYTES;
            const unsigned char* srcB = reinterpret_cast<const unsigned char*>(Qt) + ((g * ntv + ct) * nst) * (int64_t)TC_BLOCK_BYTES;
            for (int64_t st = 0; st < nst; ++st)
                for (int h = 0; h < SUB; ++h, ++gst) {
                    const uint32_t s = (uint32_t)(gst % NS), ph = (uint32_t)((gst / NS) & 1);
                    mbar_wait_relaxed(&empty_b[s], ph ^ 1u, 32);
                    mbar_expect_tx(&full_b[s], STAGE);
                    unsigned char* dst = stages + s * STAGE;
                    const unsigned char* a = srcA + st * (int64_t)TC_BLOCK_BYTES + h * PART;
                    const unsigned char* b = srcB + st * (int64_t)TC_BLOCK_BYTES + h * PART;
                    if (HALF) {     // the hi and the lo half of a 32-k slice are 16 KB apart in the 64-k block
                        bulk_g2s(dst, a, PART, &full_b[s]);
                        bulk_g2s(dst + PART, a + TC_BLOCK_BYTES / 2, PART, &full_b[s]);
                        bulk_g2s(dst + 2 * PART, b, PART, &full_b[s]);
                        bulk_g2s(dst + 3 * PART, b + TC_BLOCK_BYTES / 2, PART, &full_b[s]);
                    } else {
                        bulk_g2s(dst, a, TC_BLOCK_BYTES, &full_b[s]);
                        bulk_g2s(dst + TC_BLOCK_BYTES, b, TC_BLOCK_BYTES, &full_b[s]);
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = tc_idesc_f16(128, 128);
            const uint32_t base = smem_u32(stages);
            int64_t gst = 0;
            for (int64_t seg = 0; seg < ntile * nseg; ++seg) {
                const uint32_t a = (uint32_t)(seg & 1), pa = (uint32_t)((seg >> 1) & 1);
                mbar_wait_relaxed(&acc_empty[a], pa ^ 1u, 32);
                tc_fence_after();
                for (int q4 = 0; q4 < TC_SEG * SUB; ++q4, ++gst) {
                    const uint32_t s = (uint32_t)(gst % NS), ph = (uint32_t)((gst / NS) & 1);
                    mbar_wait_relaxed(&full_b[s], ph, 32);
                    tc_fence_after();
                    const uint32_t sa = base + s * STAGE, sb = sa + 2 * PART;
#pragma unroll
                    for (int ks = 0; ks < TC_BK / 16 / SUB; ++ks) {
                        const uint32_t off = (uint32_t)(2 * ks) * 128u * 16u;      // 2 chunks of 8 halves per k16 step
                        const uint64_t a_hi = tc_desc(sa + off, 128 * 16, 128);
                        const uint64_t a_lo = tc_desc(sa + PART + off, 128 * 16, 128);
                        const uint64_t b_hi = tc_desc(sb + off, 128 * 16, 128);
                        const uint64_t b_lo = tc_desc(sb + PART + off, 128 * 16, 128);
                        const uint32_t first = (q4 == 0 && ks == 0) ? 0u : 1u;
                        tc_mma_f16(tmem + a * 128u, a_lo, b_hi, idesc, first);
                        tc_mma_f16(tmem + a * 128u, a_hi, b_lo, idesc, 1u);
                        tc_mma_f16(tmem + a * 128u, a_hi, b_hi, idesc, 1u);
                    }
                    tc_commit(&empty_b[s]);            // the stage is free once these MMAs have read it
                }
                tc_commit(&acc_full[a]);               // the segment's accumulator is complete
            }
        }
        __syncwarp();
    } else {
        const int ew = warp - 2;                       // 0..15
        const int quarter = warp & 3;                  // TMEM lanes this warp may access
        const int cb = ew >> 2;                        // columns [32 cb, 32 cb + 32)
        const uint32_t t_lane = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(cb * 32);
        int64_t gseg = 0;
        for (int64_t it = 0; it < ntile; ++it) {
        const int64_t t = blockIdx.x + it * gridDim.x;
        const int64_t rt = t % ntu, ct = (t / ntu) % ntv, g = t / (ntu * ntv);
        const double scale = tmax ? scale0 * tmax[g] : scale0;
        double sum[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) sum[c] = 0.0;
        for (int64_t seg = 0; seg < nseg; ++seg, ++gseg) {
            const uint32_t a = (uint32_t)(gseg & 1), pa = (uint32_t)((gseg >> 1) & 1);
            mbar_wait(&acc_full[a], pa);
            tc_fence_after();
            uint32_t v0[32];
            TC_LD32(v0, t_lane + a * 128u);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) tc_arrive(&acc_empty[a]);
#pragma unroll
            for (int c = 0; c < 32; ++c) sum[c] += (double)__uint_as_float(v0[c]);
        }
        // per-warp top-kp of its 32 rows x 32 columns: every lane caches the best of its not yet consumed values
        const int64_t row = rt * 128 + quarter * 32 + lane;
        const int64_t colb = ct * 128 + cb * 32;
        if (store_out) {
            if (row < nu) {
                const double rs = scale * rowscale[row];
#pragma unroll
                for (int c = 0; c < 32; ++c)
                    if (colb + c < nv) store_out[row * ldo + colb + c] = rs * colscale[colb + c] * sum[c];
            }
        } else {
        const int64_t list = (rt * ntv + ct) * TC_EPI_WARPS + ew;
        double* ov = part_val + (list * G + g) * kp;
        int64_t* oi = part_idx + (list * G + g) * kp;
        uint32_t used = 0;
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            sum[c] = __dmul_rn(scale, sum[c]);
            if (!(row < nu && colb + c < nv)) used |= 1u << c;
        }
        const int64_t ci0 = row * nv + colb + idx_offset;
        for (int r = 0; r < kp; ++r) {
            double bv = neg_inf();
            int bc = -1;
#pragma unroll
            for (int c = 0; c < 32; ++c)           // ascending c = ascending index: strict > keeps the lowest index
                if (!((used >> c) & 1u) && (bc < 0 || sum[c] > bv)) {
                    bv = sum[c];
                    bc = c;
                }
            double wv = bv;
            int64_t wi = bc >= 0 ? ci0 + bc : INT64_MAX;
            if (bc < 0) wv = neg_inf();
            warp_argmax(wv, wi);
            const bool found = (wi != INT64_MAX);
            if (lane == 0) {
                ov[r] = found ? wv : neg_inf();
                oi[r] = found ? wi : -1;
            }
            if (!found) {
                for (int q = r + 1; q < kp; ++q)
                    if (lane == 0) {
                        ov[q] = neg_inf();
                        oi[q] = -1;
                    }
                break;
            }
            if (bc >= 0 && wi == ci0 + bc) used |= 1u << bc;   // the winning lane consumes its element
        }
        }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256));
    }
}

// ---- the same GEMM on CTA PAIRS (mode 6): tcgen05.mma.cta_group::2, 256 x 128 tiles per pair ------------------------
// prod_gemm_topk_tc_kernel is bound by operand ingest: 64 KB per 64-k stage and SM (ncu: tensor pipe 38.6 % busy, both
// operands arrive at ~30 B/clk/SM whatever the number or size of the bulk copies).  Here two CTAs of a cluster (the two
// SMs of a TPC) share one 256 x 128 tile: each stages its OWN 128 rows of P (32 KB) but only HALF of the tile's 128
// columns of Q (16 KB, its 64-column half; the tensor cores read the other half from the peer's shared memory), i.e.
// 48 KB per stage and SM for the same MMA work, and four stages fit instead of three.  One lane of the LEADER CTA
// (cluster rank 0) issues the MMAs for both SMs (M = 256: rows 0..127 accumulate in the leader's TMEM, 128..255 in the
// peer's); tcgen05.commit multicasts the stage-empty and accumulator-full arrivals to both CTAs; the peer tells the
// leader that its half of a stage has landed through a relay lane (remote mbarrier arrive), and the peer's epilogue
// warps release the accumulator on the leader's barrier the same way.  Epilogue as in the single-CTA kernel (each CTA
// drains and ranks its own 128 rows).
constexpr int T2_NS = 4;
constexpr uint32_t T2_BHALF_BYTES = TC_BLOCK_BYTES / 2;                       // 64 columns x 64 k x (hi + lo)
constexpr uint32_t T2_STAGE_BYTES = TC_BLOCK_BYTES + T2_BHALF_BYTES;          // 48 KB

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {   // observes remote arrivals
    uint32_t ok;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tc2_commit_mc(uint64_t* bar) {   // arrives on `bar` of BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void tc2_mma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
    prod_gemm_topk_tc2_kernel(const __half* __restrict__ Pt, const __half* __restrict__ Qt, int64_t nu, int64_t nv,
                              int ntu, int ntv, int nst, double scale0, const double* __restrict__ tmax,
                              int64_t idx_offset, int kp, int G, double* __restrict__ part_val,
                              int64_t* __restrict__ part_idx) {
    extern __shared__ __align__(1024) unsigned char t2_smem_raw[];
    unsigned char* stages = t2_smem_raw;                                           // [T2_NS][A block | B half block]
    uint64_t* bars = reinterpret_cast<uint64_t*>(t2_smem_raw + T2_NS * T2_STAGE_BYTES);
    uint64_t* full_b = bars;                     // [T2_NS] this CTA's producer -> (leader: MMA, peer: relay)
    uint64_t* peer_full = full_b + T2_NS;        // [T2_NS] leader only: the peer's relay -> MMA
    uint64_t* empty_b = peer_full + T2_NS;       // [T2_NS] MMA (multicast commit) -> producer of each CTA
    uint64_t* acc_full = empty_b + T2_NS;        // [2] MMA (multicast commit) -> epilogue of each CTA
    uint64_t* acc_empty = acc_full + 2;          // [2] leader only: the epilogue warps of BOTH CTAs -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair = (int)(blockIdx.x >> 1), npair = (int)(gridDim.x >> 1);
    const int ntu2 = (ntu + 1) / 2;              // row-tile PAIRS per sample
    const int tps = ntu2 * ntv;
    const int ntot = tps * G;
    const int ntile = pair < ntot ? (ntot - pair + npair - 1) / npair : 0;
    const int nseg = nst / TC_SEG;
    if (tid == 0) {
        for (int s = 0; s < T2_NS; ++s) {
            mbar_init(&full_b[s], 1);
            mbar_init(&peer_full[s], 1);
            mbar_init(&empty_b[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&acc_full[a], 1);
            mbar_init(&acc_empty[a], 2 * TC_EPI_WARPS);
        }
        mbar_fence_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();            // both CTAs: barriers initialised and TMEM allocated before anything crosses over
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t s = 0, ph = 0;
            for (int it = 0; it < ntile; ++it) {
                const int t = pair + it * npair;
                const int rt = 2 * (t % ntu2) + (int)rank, ct = (t / ntu2) % ntv, g = t / tps;
                // P blocks of this CTA's 128 rows (the array holds 2 ntu2 row tiles per sample: an odd ntu is padded with
                // a zero tile); Q half blocks: 64-column tiles, this CTA's half of column tile ct
                const unsigned char* srcA = reinterpret_cast<const unsigned char*>(Pt) +
                                            ((int64_t)(g * 2 * ntu2 + rt) * nst) * (int64_t)TC_BLOCK_BYTES;
                const unsigned char* srcB = reinterpret_cast<const unsigned char*>(Qt) +
                                            ((int64_t)(g * 2 * ntv + 2 * ct + (int)rank) * nst) * (int64_t)T2_BHALF_BYTES;
                for (int st = 0; st < nst; ++st) {
                    mbar_wait(&empty_b[s], ph ^ 1u);
                    mbar_expect_tx(&full_b[s], T2_STAGE_BYTES);
                    bulk_g2s(stages + s * T2_STAGE_BYTES, srcA + st * (int64_t)TC_BLOCK_BYTES, TC_BLOCK_BYTES, &full_b[s]);
                    bulk_g2s(stages + s * T2_STAGE_BYTES + TC_BLOCK_BYTES, srcB + st * (int64_t)T2_BHALF_BYTES,
                             T2_BHALF_BYTES, &full_b[s]);
                    if (++s == T2_NS) {
                        s = 0;
                        ph ^= 1u;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0 && !leader) {
            // relay: tell the leader's MMA lane that this CTA's half of the stage has landed
            uint32_t s = 0, ph = 0;
            const int total = ntile * nst;
            for (int x = 0; x < total; ++x) {
                mbar_wait(&full_b[s], ph);
                mbar_arrive_remote(mapa_u32(smem_u32(&peer_full[s]), 0));
                if (++s == T2_NS) {
                    s = 0;
                    ph ^= 1u;
                }
            }
        } else if (lane == 0) {
            const uint32_t idesc = tc_idesc_f16(256, 128);
            const uint32_t base = smem_u32(stages);
            uint32_t s = 0, ph = 0;
            const int segs = ntile * nseg;
            for (int seg = 0; seg < segs; ++seg) {
                const uint32_t a = (uint32_t)(seg & 1), pa = (uint32_t)((seg >> 1) & 1);
                mbar_wait_cluster(&acc_empty[a], pa ^ 1u);
                tc_fence_after();
                for (int q4 = 0; q4 < TC_SEG; ++q4) {
                    mbar_wait(&full_b[s], ph);
                    mbar_wait_cluster(&peer_full[s], ph);
                    tc_fence_after();
                    const uint32_t sa = base + s * T2_STAGE_BYTES, sb = sa + TC_BLOCK_BYTES;
#pragma unroll
                    for (int ks = 0; ks < TC_BK / 16; ++ks) {
                        const uint32_t offa = (uint32_t)(2 * ks) * 128u * 16u;     // 2 chunks of 8 halves x 128 rows
                        const uint32_t offb = (uint32_t)(2 * ks) * 64u * 16u;      // ... x 64 columns per CTA
                        const uint64_t a_hi = tc_desc(sa + offa, 128 * 16, 128);
                        const uint64_t a_lo = tc_desc(sa + TC_BLOCK_BYTES / 2 + offa, 128 * 16, 128);
                        const uint64_t b_hi = tc_desc(sb + offb, 64 * 16, 128);
                        const uint64_t b_lo = tc_desc(sb + T2_BHALF_BYTES / 2 + offb, 64 * 16, 128);
                        const uint32_t first = (q4 == 0 && ks == 0) ? 0u : 1u;
                        tc2_mma_f16(tmem + a * 128u, a_lo, b_hi, idesc, first);
                        tc2_mma_f16(tmem + a * 128u, a_hi, b_lo, idesc, 1u);
                        tc2_mma_f16(tmem + a * 128u, a_hi, b_hi, idesc, 1u);
                    }
                    tc2_commit_mc(&empty_b[s]);            // both CTAs: the stage is free once these MMAs have read it
                    if (++s == T2_NS) {
                        s = 0;
                        ph ^= 1u;
                    }
                }
                tc2_commit_mc(&acc_full[a]);               // both CTAs: the segment's accumulator is complete
            }
        }
        __syncwarp();
    } else {
        const int ew = warp - 2;                       // 0..15
        const int quarter = warp & 3;                  // TMEM lanes this warp may access
        const int cb = ew >> 2;                        // columns [32 cb, 32 cb + 32)
        const uint32_t t_lane = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(cb * 32);
        const uint32_t ae_remote0 = mapa_u32(smem_u32(&acc_empty[0]), 0), ae_remote1 = mapa_u32(smem_u32(&acc_empty[1]), 0);
        int gseg = 0;
        for (int it = 0; it < ntile; ++it) {
            const int t = pair + it * npair;
            const int rt = 2 * (t % ntu2) + (int)rank, ct = (t / ntu2) % ntv, g = t / tps;
            const double scale = scale0 * tmax[g];
            double sum[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) sum[c] = 0.0;
            for (int seg = 0; seg < nseg; ++seg, ++gseg) {
                const uint32_t a = (uint32_t)(gseg & 1), pa = (uint32_t)((gseg >> 1) & 1);
                mbar_wait(&acc_full[a], pa);
                tc_fence_after();
                uint32_t v0[32];
                TC_LD32(v0, t_lane + a * 128u);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (leader) tc_arrive(&acc_empty[a]);
                    else mbar_arrive_remote(a ? ae_remote1 : ae_remote0);
                }
#pragma unroll
                for (int c = 0; c < 32; ++c) sum[c] += (double)__uint_as_float(v0[c]);
            }
            if (rt >= ntu) continue;                   // the zero tile that pads an odd number of row tiles
            // per-warp top-kp of its 32 rows x 32 columns: every lane caches the best of its not yet consumed values
            const int64_t row = (int64_t)rt * 128 + quarter * 32 + lane;
            const int64_t colb = (int64_t)ct * 128 + cb * 32;
            const int64_t list = (int64_t)(rt * ntv + ct) * TC_EPI_WARPS + ew;
            double* ov = part_val + (list * G + g) * kp;
            int64_t* oi = part_idx + (list * G + g) * kp;
            uint32_t used = 0;
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                sum[c] = __dmul_rn(scale, sum[c]);
                if (!(row < nu && colb + c < nv)) used |= 1u << c;
            }
            const int64_t ci0 = row * nv + colb + idx_offset;
            for (int r = 0; r < kp; ++r) {
                double bv = neg_inf();
                int bc = -1;
#pragma unroll
                for (int c = 0; c < 32; ++c)           // ascending c = ascending index: strict > keeps the lowest index
                    if (!((used >> c) & 1u) && (bc < 0 || sum[c] > bv)) {
                        bv = sum[c];
                        bc = c;
                    }
                double wv = bv;
                int64_t wi = bc >= 0 ? ci0 + bc : INT64_MAX;
                if (bc < 0) wv = neg_inf();
                warp_argmax(wv, wi);
                const bool found = (wi != INT64_MAX);
                if (lane == 0) {
                    ov[r] = found ? wv : neg_inf();
                    oi[r] = found ? wi : -1;
                }
                if (!found) {
                    for (int q = r + 1; q < kp; ++q)
                        if (lane == 0) {
                            ov[q] = neg_inf();
                            oi[q] = -1;
                        }
                    break;
                }
                if (bc >= 0 && wi == ci0 + bc) used |= 1u << bc;   // the winning lane consumes its element
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();            // neither CTA leaves (or frees its TMEM) while the other may still signal / read it
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256));
    }
}

// ---- CTA pairs with 256 x 256 tiles (mode 8): two operand blocks per SM and stage feed TWICE the MMA work ------------
// scripts/tc5/mma_rate.cu measures 64 cycles per M128 N128 K16 tcgen05.mma: the 12 MMAs of a 64-k stage need 0.39 us,
// while every kernel above delivers a 64 KB stage in ~1.2 us (~54 GB/s of operand ingest per SM) -- the product GEMM is
// bound by operand bytes per MMA, and hi/lo fp16 operands cost 4 bytes per element.  A pair of CTAs on a 256 x 256
// tile halves those bytes: each CTA stages its own 128 rows of P (32 KB) and ONE of the tile's two 128-column Q blocks
// (32 KB; the tensor cores of both SMs read both halves), and the leader issues M256 x N256 x K16 MMAs: 24 per stage
// and SM instead of 12 for the same 64 KB.  The two accumulators of 256 columns fill the tensor memory; every epilogue
// thread then owns 64 columns (32 of each column tile), which only fit in registers as FP32 running sums (the segments
// bound the tensor core's internal accumulation error; adding 2F / 256 segment values in fp32 adds one rounding of
// 2^-24 of the sum of |terms| each -- carried by prod_exact_kernel's bound, +1.5 % at F = 1024).
constexpr uint32_t T8_STAGE_BYTES = 2 * TC_BLOCK_BYTES;                       // own A block + one B block = 64 KB
constexpr int T8_NS = 3;

#define TC_LD16(v, taddr)                                                                                             \
    asm volatile(                                                                                                     \
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"      \
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),             \
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])        \
        : "r"(taddr))

// HALF = 1: stages of 32 k (32 KB: [A hi | A lo | B hi | B lo], 8 KB each), NS of them: the pair hand-offs (relay lane,
// multicast commits) add ~1.6 us to the cycle of a slot, which more, smaller slots hide
template <int HALF, int NS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
    prod_gemm_topk_tc8_kernel(const __half* __restrict__ Pt, const __half* __restrict__ Qt, int64_t nu, int64_t nv,
                              int ntu, int ntv, int nst, double scale0, const double* __restrict__ tmax,
                              int64_t idx_offset, int kp, int G, double* __restrict__ part_val,
                              int64_t* __restrict__ part_idx, const double* __restrict__ thr,
                              int64_t* __restrict__ emit_idx, int* __restrict__ emit_cnt, int emit_cap) {
    // thr != nullptr (tes_rff_topk_product_resolve_f64): instead of the per-warp top-kp lists, every candidate whose
    // value reaches thr[g] is appended to emit_idx[g][.] (emit_cnt[g] counts them, entries beyond emit_cap are dropped)
    extern __shared__ __align__(1024) unsigned char t8_smem_raw[];
    constexpr uint32_t STAGE = HALF ? T8_STAGE_BYTES / 2 : T8_STAGE_BYTES;
    constexpr uint32_t PART = HALF ? TC_BLOCK_BYTES / 4 : TC_BLOCK_BYTES / 2;      // bytes of the hi (= of the lo) part
    constexpr int SUB = HALF ? 2 : 1;
    unsigned char* stages = t8_smem_raw;                                           // [NS][A hi | A lo | B hi | B lo]
    uint64_t* bars = reinterpret_cast<uint64_t*>(t8_smem_raw + NS * STAGE);
    uint64_t* full_b = bars;                     // [NS] this CTA's producer -> (leader: MMA, peer: relay)
    uint64_t* peer_full = full_b + NS;           // [NS] leader only: the peer's relay -> MMA
    uint64_t* empty_b = peer_full + NS;          // [NS] MMA (multicast commit) -> producer of each CTA
    uint64_t* acc_full = empty_b + NS;           // [2] MMA (multicast commit) -> epilogue of each CTA
    uint64_t* acc_empty = acc_full + 2;          // [2] leader only: the epilogue warps of BOTH CTAs -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair = (int)(blockIdx.x >> 1), npair = (int)(gridDim.x >> 1);
    const int ntu2 = (ntu + 1) / 2, ntv2 = (ntv + 1) / 2;     // 256-row / 256-column tile pairs per sample
    const int tps = ntu2 * ntv2;
    const int ntot = tps * G;
    const int ntile = pair < ntot ? (ntot - pair + npair - 1) / npair : 0;
    const int nseg = nst / TC_SEG;
    if (tid == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(&full_b[s], 1);
            mbar_init(&peer_full[s], 1);
            mbar_init(&empty_b[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&acc_full[a], 1);
            mbar_init(&acc_empty[a], 2 * TC_EPI_WARPS);
        }
        mbar_fence_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t s = 0, ph = 0;
            for (int it = 0; it < ntile; ++it) {
                const int t = pair + it * npair;
                const int rt = 2 * (t % ntu2) + (int)rank, ct = 2 * ((t / ntu2) % ntv2) + (int)rank, g = t / tps;
                // both arrays hold an EVEN number of 128-row tiles per sample (an odd count is padded with a zero tile)
                const unsigned char* srcA = reinterpret_cast<const unsigned char*>(Pt) +
                                            ((int64_t)(g * 2 * ntu2 + rt) * nst) * (int64_t)TC_BLOCK_BYTES;
                const unsigned char* srcB = reinterpret_cast<const unsigned char*>(Qt) +
                                            ((int64_t)(g * 2 * ntv2 + ct) * nst) * (int64_t)TC_BLOCK_BYTES;
                for (int st = 0; st < nst; ++st)
                    for (int h = 0; h < SUB; ++h) {
                        mbar_wait(&empty_b[s], ph ^ 1u);
                        mbar_expect_tx(&full_b[s], STAGE);
                        unsigned char* dst = stages + s * STAGE;
                        const unsigned char* a = srcA + st * (int64_t)TC_BLOCK_BYTES + h * PART;
                        const unsigned char* b = srcB + st * (int64_t)TC_BLOCK_BYTES + h * PART;
                        if (HALF) {
                            bulk_g2s(dst, a, PART, &full_b[s]);
                            bulk_g2s(dst + PART, a + TC_BLOCK_BYTES / 2, PART, &full_b[s]);
                            bulk_g2s(dst + 2 * PART, b, PART, &full_b[s]);
                            bulk_g2s(dst + 3 * PART, b + TC_BLOCK_BYTES / 2, PART, &full_b[s]);
                        } else {
                            bulk_g2s(dst, a, TC_BLOCK_BYTES, &full_b[s]);
                            bulk_g2s(dst + TC_BLOCK_BYTES, b, TC_BLOCK_BYTES, &full_b[s]);
                        }
                        if (++s == NS) {
                            s = 0;
                            ph ^= 1u;
                        }
                    }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0 && !leader) {
            uint32_t s = 0, ph = 0;                 // relay: this CTA's half of the stage has landed
            const int total = ntile * nst * SUB;
            for (int x = 0; x < total; ++x) {
                mbar_wait(&full_b[s], ph);
                mbar_arrive_remote(mapa_u32(smem_u32(&peer_full[s]), 0));
                if (++s == NS) {
                    s = 0;
                    ph ^= 1u;
                }
            }
        } else if (lane == 0) {
            const uint32_t idesc = tc_idesc_f16(256, 256);
            const uint32_t base = smem_u32(stages);
            uint32_t s = 0, ph = 0;
            const int segs = ntile * nseg;
            for (int seg = 0; seg < segs; ++seg) {
                const uint32_t a = (uint32_t)(seg & 1), pa = (uint32_t)((seg >> 1) & 1);
                mbar_wait_cluster(&acc_empty[a], pa ^ 1u);
                tc_fence_after();
                for (int q4 = 0; q4 < TC_SEG * SUB; ++q4) {
                    mbar_wait(&full_b[s], ph);
                    mbar_wait_cluster(&peer_full[s], ph);
                    tc_fence_after();
                    const uint32_t sa = base + s * STAGE, sb = sa + 2 * PART;
#pragma unroll
                    for (int ks = 0; ks < TC_BK / 16 / SUB; ++ks) {
                        const uint32_t off = (uint32_t)(2 * ks) * 128u * 16u;
                        const uint64_t a_hi = tc_desc(sa + off, 128 * 16, 128);
                        const uint64_t a_lo = tc_desc(sa + PART + off, 128 * 16, 128);
                        const uint64_t b_hi = tc_desc(sb + off, 128 * 16, 128);
                        const uint64_t b_lo = tc_desc(sb + PART + off, 128 * 16, 128);
                        const uint32_t first = (q4 == 0 && ks == 0) ? 0u : 1u;
                        tc2_mma_f16(tmem + a * 256u, a_lo, b_hi, idesc, first);
                        tc2_mma_f16(tmem + a * 256u, a_hi, b_lo, idesc, 1u);
                        tc2_mma_f16(tmem + a * 256u, a_hi, b_hi, idesc, 1u);
                    }
                    tc2_commit_mc(&empty_b[s]);
                    if (++s == NS) {
                        s = 0;
                        ph ^= 1u;
                    }
                }
                tc2_commit_mc(&acc_full[a]);
            }
        }
        __syncwarp();
    } else {
        const int ew = warp - 2;                       // 0..15
        const int quarter = warp & 3;                  // TMEM lanes this warp may access
        const int cb = ew >> 2;                        // columns [32 cb, 32 cb + 32) of EACH of the two column tiles
        const uint32_t t_lane = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(cb * 32);
        const uint32_t ae_remote0 = mapa_u32(smem_u32(&acc_empty[0]), 0), ae_remote1 = mapa_u32(smem_u32(&acc_empty[1]), 0);
        int gseg = 0;
        for (int it = 0; it < ntile; ++it) {
            const int t = pair + it * npair;
            const int rt = 2 * (t % ntu2) + (int)rank, ctp = (t / ntu2) % ntv2, g = t / tps;
            const double scale = scale0 * tmax[g];
            float acc[2][32];
#pragma unroll
            for (int c = 0; c < 32; ++c) acc[0][c] = acc[1][c] = 0.0f;
            for (int seg = 0; seg < nseg; ++seg, ++gseg) {
                const uint32_t a = (uint32_t)(gseg & 1), pa = (uint32_t)((gseg >> 1) & 1);
                mbar_wait(&acc_full[a], pa);
                tc_fence_after();
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        uint32_t v0[16];
                        TC_LD16(v0, t_lane + a * 256u + (uint32_t)(h * 128 + q * 16));
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                        for (int c = 0; c < 16; ++c) acc[h][q * 16 + c] = __fadd_rn(acc[h][q * 16 + c], __uint_as_float(v0[c]));
                    }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (leader) tc_arrive(&acc_empty[a]);
                    else mbar_arrive_remote(a ? ae_remote1 : ae_remote0);
                }
            }
            if (rt >= ntu) continue;                   // the zero tile that pads an odd number of row tiles
            const int64_t row = (int64_t)rt * 128 + quarter * 32 + lane;
            if (thr) {
                const double th = thr[g];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int ct = 2 * ctp + h;
                    if (ct >= ntv || row >= nu) continue;
                    const int64_t colb = (int64_t)ct * 128 + cb * 32;
#pragma unroll
                    for (int c = 0; c < 32; ++c)
                        if (colb + c < nv && __dmul_rn(scale, (double)acc[h][c]) >= th) {
                            const int slot = atomicAdd(emit_cnt + g, 1);
                            if (slot < emit_cap) emit_idx[(int64_t)g * emit_cap + slot] = row * nv + colb + c + idx_offset;
                        }
                }
                continue;
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int ct = 2 * ctp + h;
                if (ct >= ntv) continue;
                // per-warp top-kp of its 32 rows x 32 columns of column tile ct
                const int64_t colb = (int64_t)ct * 128 + cb * 32;
                const int64_t list = (int64_t)(rt * ntv + ct) * TC_EPI_WARPS + ew;
                double* ov = part_val + (list * G + g) * kp;
                int64_t* oi = part_idx + (list * G + g) * kp;
                uint32_t used = 0;
#pragma unroll
                for (int c = 0; c < 32; ++c)
                    if (!(row < nu && colb + c < nv)) used |= 1u << c;
                const int64_t ci0 = row * nv + colb + idx_offset;
                for (int r = 0; r < kp; ++r) {
                    float bh = 0.0f;
                    int bc = -1;
#pragma unroll
                    for (int c = 0; c < 32; ++c)       // ascending c = ascending index: strict > keeps the lowest index
                        if (!((used >> c) & 1u) && (bc < 0 || acc[h][c] > bh)) {
                            bh = acc[h][c];
                            bc = c;
                        }
                    double wv = bc >= 0 ? __dmul_rn(scale, (double)bh) : neg_inf();
                    int64_t wi = bc >= 0 ? ci0 + bc : INT64_MAX;
                    warp_argmax(wv, wi);
                    const bool found = (wi != INT64_MAX);
                    if (lane == 0) {
                        ov[r] = found ? wv : neg_inf();
                        oi[r] = found ? wi : -1;
                    }
                    if (!found) {
                        for (int q = r + 1; q < kp; ++q)
                            if (lane == 0) {
                                ov[q] = neg_inf();
                                oi[q] = -1;
                            }
                        break;
                    }
                    if (bc >= 0 && wi == ci0 + bc) used |= 1u << bc;   // the winning lane consumes its element
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
    }
}

// ---- the same GEMM with the P operand GENERATED ON THE SM (mode 4) -------------------------------------------------
// prod_gemm_topk_tc_kernel streams both operands: 64 KB per 64-k stage and SM, i.e. more than an SM ingests while the
// tensor pipe works through a stage (ncu, round 1: tensor pipe 38.6 % busy, epilogue warps parked on acc_full), and its
// operands make a round trip through HBM (prod_build_tc_kernel: 16.8 GB written and read back per C3 step).  Here only
// the Q blocks (the V side, no theta) are built ahead and streamed (32 KB per stage); the P block of every stage --
// theta~ cos(alpha), -theta~ sin(alpha), alpha = w_U . u + b, for the tile's 128 rows and the stage's 32 features -- is
// computed by the 16 epilogue warps straight into the stage buffer (generic-proxy stores + fence.proxy.async), with
// exactly the arithmetic of prod_build_tc_kernel (fp64 argument, rint by the 2^52 shift, __sincosf, fp32 theta product,
// hi/lo fp16 split), so the error budget of prod_exact_kernel is unchanged.  The feature records
// [w_U (ncol), b, theta~] reach shared memory through a ring of small bulk copies issued PG_AHEAD stages ahead by the
// producer lane; a ring slot is reused PG_RING stages later, which the stage-empty barrier the producer has just
// passed already orders after the generators' reads (its MMAs consumed an A block written from those records).
// Every fourth stage the same warps drain the previous 256-k segment from TMEM into double-float running sums.
constexpr int PG_NS = 3, PG_RING = 8, PG_AHEAD = 5, PG_MAXCOL = 6, PG_THREADS = 640;
__host__ __device__ inline int pg_rs(int ncol) { return ncol + 2; }
__host__ __device__ inline size_t pg_smem_bytes(int ncol) {
    return (size_t)PG_NS * TC_STAGE_BYTES + (size_t)PG_RING * 32 * pg_rs(ncol) * 8 + (size_t)ncol * 128 * 8 + 512;
}

// featP[(j * Fp + f) * rs + ...] = [w (ncol slow columns), b, theta_j[f] / tmax_j as a float in the low word]; features
// F .. Fp - 1 are zero records (theta~ = 0 -> zero operand columns)
__global__ void prod_featp_kernel(const double* __restrict__ W, const double* __restrict__ b,
                                  const double* __restrict__ theta, const double* __restrict__ tmax, int64_t S,
                                  int64_t F, int64_t Fp, int d, int s_lo, int ncol, int w_shared,
                                  double* __restrict__ featP) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= S * Fp) return;
    const int64_t j = t / Fp, f = t % Fp, jw = w_shared ? 0 : j;
    const int rs = pg_rs(ncol);
    double* o = featP + t * rs;
    if (f >= F) {
        for (int c = 0; c < rs; ++c) o[c] = 0.0;
        return;
    }
    for (int c = 0; c < ncol; ++c) o[c] = W[(jw * F + f) * d + s_lo + c];
    o[ncol] = b[jw * F + f];
    const double tm = tmax[j];
    const double inv_tm = tm > 0.0 ? 1.0 / tm : 0.0;
    o[ncol + 1] = __hiloint2double(0, __float_as_int((float)(theta[j * F + f] * inv_tm)));
}

// GC = chunks of 8 k per 64-k stage generated on the SM (8: the whole P block; 4: the first half, the second half is
// streamed from the array prod_build_tc_kernel wrote -- generation and operand ingest then share the load)
template <int NCOL, int GC>
__global__ void __launch_bounds__(PG_THREADS, 1)
    prod_gemm_gen_kernel(const double* __restrict__ cand, int64_t nu, int64_t nv, int d, int s_lo,
                         const double* __restrict__ featP, int64_t Fp, const __half* __restrict__ Pt,
                         const __half* __restrict__ Qt, int ntu, int ntv,
                         int nst, double scale0, const double* __restrict__ tmax, int64_t j0, int64_t idx_offset, int kp,
                         int G, double* __restrict__ part_val, int64_t* __restrict__ part_idx) {
    extern __shared__ __align__(1024) unsigned char pg_smem_raw[];
    constexpr int RS = NCOL + 2;
    constexpr uint32_t REC_BYTES = 32 * RS * 8;
    unsigned char* stages = pg_smem_raw;                                               // [PG_NS][A block | B block]
    double* ring = reinterpret_cast<double*>(pg_smem_raw + PG_NS * TC_STAGE_BYTES);    // [PG_RING][32][RS]
    double* xs = ring + PG_RING * 32 * RS;                                             // [NCOL][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(xs + NCOL * 128);
    uint64_t* full_b = bars;                    // [PG_NS] producer (B block landed) -> MMA
    uint64_t* a_full = full_b + PG_NS;          // [PG_NS] generators (A block written) -> MMA
    uint64_t* empty_b = a_full + PG_NS;         // [PG_NS] MMA (commit) -> producer, generators
    uint64_t* rec_full = empty_b + PG_NS;       // [PG_RING] producer (records landed) -> generators
    uint64_t* acc_full = rec_full + PG_RING;    // [2] MMA (commit) -> epilogue
    uint64_t* acc_empty = acc_full + 2;         // [2] epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tps = ntu * ntv;                  // tiles per sample
    const int ntot = tps * G;
    const int ntile = (int)blockIdx.x < ntot ? (ntot - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int nseg = nst / TC_SEG;
    if (tid == 0) {
        for (int s = 0; s < PG_NS; ++s) {
            mbar_init(&full_b[s], 1);
            mbar_init(&a_full[s], TC_EPI_WARPS);
            mbar_init(&empty_b[s], 1);
        }
        for (int r = 0; r < PG_RING; ++r) mbar_init(&rec_full[r], 1);
        for (int a = 0; a < 2; ++a) {
            mbar_init(&acc_full[a], 1);
            mbar_init(&acc_empty[a], TC_EPI_WARPS);
        }
        mbar_fence_init();
    }
    if (warp == 17) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    // warps 0..15: generators + epilogue; warp 16 = producer lane, warp 17 = MMA lane, 18 / 19 idle

    if (warp == 16) {
        if (lane == 0) {
            // records run PG_AHEAD stages ahead of the B blocks: (ri, rst) = tile / stage of the next record block
            int ri = 0, rst = 0;
            uint32_t rr = 0;
            auto issue_rec = [&]() {
                const int g = ((int)blockIdx.x + ri * (int)gridDim.x) / tps;
                mbar_expect_tx(&rec_full[rr], REC_BYTES);
                bulk_g2s(ring + (size_t)rr * 32 * RS, featP + ((j0 + g) * Fp + (int64_t)rst * 32) * RS, REC_BYTES,
                         &rec_full[rr]);
                rr = (rr + 1) & (PG_RING - 1);
                if (++rst == nst) {
                    rst = 0;
                    ++ri;
                }
            };
            for (int x = 0; x < PG_AHEAD && ri < ntile; ++x) issue_rec();
            uint32_t s = 0, ph = 0;
            for (int it = 0; it < ntile; ++it) {
                const int t = (int)blockIdx.x + it * (int)gridDim.x;
                const int rt = t % ntu, ct = (t / ntu) % ntv, g = t / tps;
                const unsigned char* srcB = reinterpret_cast<const unsigned char*>(Qt) +
                                            ((int64_t)(g * ntv + ct) * nst) * (int64_t)TC_BLOCK_BYTES;
                const unsigned char* srcA = reinterpret_cast<const unsigned char*>(Pt) +
                                            ((int64_t)(g * ntu + rt) * nst) * (int64_t)TC_BLOCK_BYTES;
                for (int st = 0; st < nst; ++st) {
                    mbar_wait_relaxed(&empty_b[s], ph ^ 1u, 32);
                    constexpr uint32_t A_STREAM = (8 - GC) * 128 * 16;      // bytes of the hi (and of the lo) half streamed
                    mbar_expect_tx(&full_b[s], TC_BLOCK_BYTES + 2 * A_STREAM);
                    bulk_g2s(stages + s * TC_STAGE_BYTES + TC_BLOCK_BYTES, srcB + st * (int64_t)TC_BLOCK_BYTES,
                             TC_BLOCK_BYTES, &full_b[s]);
                    if (GC < 8) {
                        const unsigned char* sa = srcA + st * (int64_t)TC_BLOCK_BYTES + GC * 128 * 16;
                        unsigned char* da = stages + s * TC_STAGE_BYTES + GC * 128 * 16;
                        bulk_g2s(da, sa, A_STREAM, &full_b[s]);
                        bulk_g2s(da + TC_BLOCK_BYTES / 2, sa + TC_BLOCK_BYTES / 2, A_STREAM, &full_b[s]);
                    }
                    // the ring slot about to be refilled last held the records of the stage 3 = PG_RING - PG_AHEAD back,
                    // whose A block the MMAs behind the barrier just passed have consumed
                    if (ri < ntile) issue_rec();
                    if (++s == PG_NS) {
                        s = 0;
                        ph ^= 1u;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 17) {
        if (lane == 0) {
            const uint32_t idesc = tc_idesc_f16(128, 128);
            const uint32_t base = smem_u32(stages);
            uint32_t s = 0, ph = 0;
            const int segs = ntile * nseg;
            for (int seg = 0; seg < segs; ++seg) {
                const uint32_t a = (uint32_t)(seg & 1), pa = (uint32_t)((seg >> 1) & 1);
                mbar_wait_relaxed(&acc_empty[a], pa ^ 1u, 32);
                tc_fence_after();
                for (int q4 = 0; q4 < TC_SEG; ++q4) {
                    mbar_wait(&full_b[s], ph);      // (tight waits: the hand-off latency is on the critical path here)
                    mbar_wait(&a_full[s], ph);
                    tc_fence_after();
                    const uint32_t sa = base + s * TC_STAGE_BYTES, sb = sa + TC_BLOCK_BYTES;
#pragma unroll
                    for (int ks = 0; ks < TC_BK / 16; ++ks) {
                        const uint32_t off = (uint32_t)(2 * ks) * 128u * 16u;
                        const uint64_t a_hi = tc_desc(sa + off, 128 * 16, 128);
                        const uint64_t a_lo = tc_desc(sa + TC_BLOCK_BYTES / 2 + off, 128 * 16, 128);
                        const uint64_t b_hi = tc_desc(sb + off, 128 * 16, 128);
                        const uint64_t b_lo = tc_desc(sb + TC_BLOCK_BYTES / 2 + off, 128 * 16, 128);
                        const uint32_t first = (q4 == 0 && ks == 0) ? 0u : 1u;
                        tc_mma_f16(tmem + a * 128u, a_lo, b_hi, idesc, first);
                        tc_mma_f16(tmem + a * 128u, a_hi, b_lo, idesc, 1u);
                        tc_mma_f16(tmem + a * 128u, a_hi, b_hi, idesc, 1u);
                    }
                    tc_commit(&empty_b[s]);
                    if (++s == PG_NS) {
                        s = 0;
                        ph ^= 1u;
                    }
                }
                tc_commit(&acc_full[a]);
            }
        }
        __syncwarp();
    } else if (warp < 16) {
        const int ew = warp;
        const int quarter = warp & 3, cb = ew >> 2;
        const uint32_t t_lane = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(cb * 32);
        const int te = tid, gr = te & 127, part = te >> 7;     // generator role: row gr, GC / 4 chunks of four features
        constexpr int CPT = GC / 4;                           // chunks per thread and stage
        const double* xrow = xs + gr;
        const uint32_t aoff = (uint32_t)((CPT * part) * 128 + gr) * 16u;
        uint32_t s = 0, ph = 0, rr = 0, pr = 0;
        int gs = 0;
        for (int it = 0; it < ntile; ++it) {
            const int t = (int)blockIdx.x + it * (int)gridDim.x;
            const int rt = t % ntu, ct = (t / ntu) % ntv, g = t / tps;
            // the tile's slow coordinates; the MMAs of the previous tile may still be in flight (they do not read xs)
            asm volatile("bar.sync 1, 512;" ::: "memory");
            for (int e = te; e < NCOL * 128; e += 512) {
                const int c = e >> 7, r_ = e & 127;
                const int64_t grow = (int64_t)rt * 128 + r_;
                xs[e] = grow < nu ? cand[(grow * nv) * d + s_lo + c] : 0.0;
            }
            asm volatile("bar.sync 1, 512;" ::: "memory");
            double x[NCOL];
#pragma unroll
            for (int c = 0; c < NCOL; ++c) x[c] = xrow[c * 128];
            // running sums of the 256-k fp32 segments, in fp32: the segments bound the tensor core's internal accumulation
            // error; adding nseg of them adds nseg roundings of 2^-24 of the sum of |terms| (prod_exact_kernel's B
            // carries 1.5 nseg 2^-24: +1.5 % at F = 1024) -- and costs 32 registers and one FADD per value instead of
            // 64 registers and a TwoSum
            float acc[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) acc[c] = 0.0f;
            auto drain = [&](int dseg) {
                const uint32_t a = (uint32_t)(dseg & 1), pa = (uint32_t)((dseg >> 1) & 1);
                mbar_wait(&acc_full[a], pa);
                tc_fence_after();
                uint32_t v0[32];
                TC_LD32(v0, t_lane + a * 128u);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) tc_arrive(&acc_empty[a]);
#pragma unroll
                for (int c = 0; c < 32; ++c) acc[c] = __fadd_rn(acc[c], __uint_as_float(v0[c]));
            };
            // iteration seg generates segment seg (if any) and then drains segment seg - 1: ONE call site of the drain, so
            // that it is inlined and the running sums stay in registers
            for (int seg = 0; seg <= nseg; ++seg) {
#pragma unroll 1
                for (int q4 = 0; q4 < TC_SEG && seg < nseg; ++q4) {
                    mbar_wait(&rec_full[rr], pr);
                    // (the slot was released by the MMAs of three stages ago: when the generators are the bottleneck this
                    // wait is free, and waiting first lets every chunk be stored as soon as it is formed)
                    mbar_wait(&empty_b[s], ph ^ 1u);
                    const double* rec = ring + (size_t)rr * 32 * RS + (size_t)(4 * CPT * part) * RS;
                    unsigned char* ablk = stages + s * TC_STAGE_BYTES + aoff;
#pragma unroll
                    for (int hq = 0; hq < CPT; ++hq) {
                        uint32_t hi[4], lo[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const double* rq = rec + (4 * hq + q) * RS;
                            double a = rq[NCOL];
#pragma unroll
                            for (int c = 0; c < NCOL; ++c) a = fma(rq[c], x[c], a);
                            const float th = __int_as_float(__double2loint(rq[NCOL + 1]));
                            // same sequence as prod_build_tc_kernel (the operand error budget of prod_exact_kernel)
                            float sn, cs;
                            const double tt = a * 0.15915494309189535;
                            const double red = __dadd_rn(__dadd_rn(tt, 6755399441055744.0), -6755399441055744.0);
                            __sincosf((float)(a - 6.283185307179586 * red), &sn, &cs);
                            const float p0 = __fmul_rn(th, cs), p1 = __fmul_rn(-th, sn);
                            const __half2 h = __floats2half2_rn(p0, p1);
                            const float2 hf = __half22float2(h);
                            const __half2 l = __floats2half2_rn(p0 - hf.x, p1 - hf.y);
                            hi[q] = *reinterpret_cast<const uint32_t*>(&h);
                            lo[q] = *reinterpret_cast<const uint32_t*>(&l);
                        }
                        *reinterpret_cast<uint4*>(ablk + hq * 128 * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                        *reinterpret_cast<uint4*>(ablk + TC_BLOCK_BYTES / 2 + hq * 128 * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic stores -> visible to the MMA
                    __syncwarp();
                    if (lane == 0) tc_arrive(&a_full[s]);
                    if (++s == PG_NS) {
                        s = 0;
                        ph ^= 1u;
                    }
                    rr = (rr + 1) & (PG_RING - 1);
                    if (rr == 0) pr ^= 1u;
                }
                // the PREVIOUS segment: its MMAs were issued a segment ago, so the wait is (nearly) free and the tensor
                // pipe keeps running on the stages just generated
                if (seg > 0) drain(gs++);
            }
            // per-warp top-kp of its 32 rows x 32 columns (same selection as prod_gemm_topk_tc_kernel)
            const double scale = scale0 * tmax[j0 + g];
            const int64_t row = (int64_t)rt * 128 + quarter * 32 + lane;
            const int64_t colb = (int64_t)ct * 128 + cb * 32;
            const int64_t list = (int64_t)(rt * ntv + ct) * TC_EPI_WARPS + ew;
            double* ov = part_val + (list * G + g) * kp;
            int64_t* oi = part_idx + (list * G + g) * kp;
            uint32_t used = 0;
#pragma unroll
            for (int c = 0; c < 32; ++c)
                if (!(row < nu && colb + c < nv)) used |= 1u << c;
            const int64_t ci0 = row * nv + colb + idx_offset;
            for (int r = 0; r < kp; ++r) {
                float bh = 0.0f;
                int bc = -1;
#pragma unroll
                for (int c = 0; c < 32; ++c)           // ascending c = ascending index: strict > keeps the lowest index
                    if (!((used >> c) & 1u) && (bc < 0 || acc[c] > bh)) {
                        bh = acc[c];
                        bc = c;
                    }
                double wv = bc >= 0 ? __dmul_rn(scale, (double)bh) : neg_inf();
                int64_t wi = bc >= 0 ? ci0 + bc : INT64_MAX;
                warp_argmax(wv, wi);
                const bool found = (wi != INT64_MAX);
                if (lane == 0) {
                    ov[r] = found ? wv : neg_inf();
                    oi[r] = found ? wi : -1;
                }
                if (!found) {
                    for (int q = r + 1; q < kp; ++q)
                        if (lane == 0) {
                            ov[q] = neg_inf();
                            oi[q] = -1;
                        }
                    break;
                }
                if (bc >= 0 && wi == ci0 + bc) used |= 1u << bc;   // the winning lane consumes its element
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 17) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256));
    }
}

template <int NCOL, int GC>
static int launch_prod_gemm_gen(int sms, const double* cand, int64_t nu, int64_t nv, int d, int s_lo, const double* featP,
                                int64_t Fp, const __half* Pt, const __half* Qt, int ntu, int ntv, int nst, double scale,
                                const double* tmax, int64_t j0, int64_t idx_offset, int kp, int G, double* pv, int64_t* pi,
                                cudaStream_t st) {
    TES_CUDA_OK(cudaFuncSetAttribute(prod_gemm_gen_kernel<NCOL, GC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)pg_smem_bytes(NCOL)));
    const int ntot = ntu * ntv * G;
    prod_gemm_gen_kernel<NCOL, GC><<<(unsigned)(ntot < sms ? ntot : sms), PG_THREADS, pg_smem_bytes(NCOL), st>>>(
        cand, nu, nv, d, s_lo, featP, Fp, Pt, Qt, ntu, ntv, nst, scale, tmax, j0, idx_offset, kp, G, pv, pi);
    TES_LAUNCH_OK();
    return 0;
}

// ---- TES_ep quadratic forms as a 3xFP16 tcgen05 GEMM (screen; criteria.cu refines the candidates near the arg-max) --
// A[x][k] = M[x,a_k] M[x,b_k] (index pairs of the slice table), normalised per row to max 4096 so that the fp16 hi/lo
// pair keeps 22 significant bits down to products 2^-37 of the largest; B[k][j] = Bpack[k][j] normalised per column to
// max 8.  Both in MMA order (see above), K padded with zeros to whole 256-k segments.
constexpr double QS_AMAX = 4096.0, QS_BMAX = 8.0;
__global__ void quad_rowmax_kernel(const double* __restrict__ G, int64_t ldg, int64_t nc, int64_t T,
                                   double* __restrict__ rowmax) {
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= nc) return;
    double m = 0.0;
    for (int64_t a = lane; a < T; a += 32) m = fmax(m, fabs(G[row * ldg + a]));
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, off));
    if (lane == 0) rowmax[row] = m;
}
// One CTA per 128-row tile: M[128][T] is staged once in shared memory as floats ([a][row]: conflict-free for
// row-consecutive threads; the products only need the 22 bits the hi/lo fp16 pair keeps), then the CTA walks the 8-k
// chunks: thread (row, half) forms 8 products per chunk and writes 16 bytes of hi and of lo -- row-consecutive threads
// write consecutive 16-byte slots of the MMA layout.  (A first version read M straight from global memory with a
// row-sized stride between lanes: 0.86 ms per chunk, 45 ms per C3 step.)
__global__ void __launch_bounds__(256)
    quad_build_a_kernel(const double* __restrict__ G, int64_t ldg, int64_t nc, int64_t T, const uint32_t* __restrict__ ab,
                        int64_t Kdim, int64_t Kp, const double* __restrict__ rowmax, __half* __restrict__ out) {
    extern __shared__ float qa_sm[];                 // [T][129] (odd stride: conflict-free for both access patterns)
    const int64_t tile = blockIdx.x;
    const int r = threadIdx.x & 127, half = threadIdx.x >> 7;
    const int64_t row = tile * 128 + r;
    for (int64_t e = threadIdx.x; e < T * 128; e += 256) {
        const int64_t rr = e / T, a = e % T;          // lanes run along a row of M: coalesced reads
        const int64_t grow = tile * 128 + rr;
        qa_sm[a * 129 + rr] = grow < nc ? (float)G[grow * ldg + a] : 0.0f;
    }
    __syncthreads();
    float sc = 0.0f;
    if (row < nc) {
        const double rm = rowmax[row];
        sc = rm > 0.0 ? (float)(QS_AMAX / (rm * rm)) : 0.0f;
    }
    const int64_t nst = Kp / TC_BK, nkc = Kp / 8;
    const int64_t kc_lo = nkc * blockIdx.y / gridDim.y, kc_hi = nkc * (blockIdx.y + 1) / gridDim.y;   // k range of this CTA
    for (int64_t kc = kc_lo + half; kc < kc_hi; kc += 2) {
        const int64_t k0 = kc * 8, st = k0 / TC_BK, chunk = (k0 % TC_BK) / 8;
        __align__(16) __half hi[8], lo[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) hi[e] = lo[e] = __float2half_rn(0.0f);
        if (k0 < Kdim) {
            // ab[k] = a | b << 16 (0xffffffff: padding), decoded once per call by quad_ab_table_kernel
            const uint4 t0 = __ldg(reinterpret_cast<const uint4*>(ab + k0));
            const uint4 t1 = __ldg(reinterpret_cast<const uint4*>(ab + k0 + 4));
            const uint32_t tb[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
            for (int e = 0; e < 8; ++e)
                if (tb[e] != 0xffffffffu)
                    split_f16(qa_sm[(tb[e] & 0xffffu) * 129 + r] * qa_sm[(tb[e] >> 16) * 129 + r] * sc, hi[e], lo[e]);
        }
        __half* blk = out + (tile * nst + st) * (int64_t)(TC_BLOCK_BYTES / 2);
        const int64_t o = (chunk * 128 + r) * 8;
        *reinterpret_cast<uint4*>(blk + o) = *reinterpret_cast<const uint4*>(hi);
        *reinterpret_cast<uint4*>(blk + 8 * 128 * 8 + o) = *reinterpret_cast<const uint4*>(lo);
    }
}
// K axis of the screen in RUN order: for a = 0 .. T-1 the pairs (a, b), b = a .. T-1, each run padded with empty entries
// to a multiple of 8 -- every 8-k operand chunk has one a and consecutive b (what the fused kernel's generators exploit)
__global__ void quad_ab_runs_kernel(int64_t T, int64_t Kp, uint32_t* __restrict__ ab) {
    for (int64_t k = threadIdx.x; k < Kp; k += blockDim.x) ab[k] = 0xffffffffu;
    __syncthreads();
    for (int64_t a = threadIdx.x; a < T; a += blockDim.x) {
        int64_t off = 0;
        for (int64_t a2 = 0; a2 < a; ++a2) off += (T - a2 + 7) / 8 * 8;
        for (int64_t b = a; b < T; ++b) ab[off + b - a] = (uint32_t)a | ((uint32_t)b << 16);
    }
}
// Bpack[k][j] = Sigma_j[a][a] (a = b), Sigma_j[a][b] + Sigma_j[b][a] (a < b), 0 (padding)
__global__ void quad_pack_ab_kernel(const double* __restrict__ cov, int64_t Sp, int64_t T, const uint32_t* __restrict__ ab,
                                    int64_t Kdim, double* __restrict__ Bpack) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= Kdim * Sp) return;
    const int64_t k = e / Sp, j = e % Sp;
    const uint32_t v = ab[k];
    double r = 0.0;
    if (v != 0xffffffffu) {
        const int64_t a = v & 0xffffu, b = v >> 16;
        const double* c = cov + j * T * T;
        r = c[a * T + b];
        if (a != b) r += c[b * T + a];
    }
    Bpack[e] = r;
}
__global__ void quad_colmax_kernel(const double* __restrict__ Bpack, int64_t Kdim, int64_t Sp,
                                   double* __restrict__ colmax) {
    __shared__ double red[256];
    const int64_t j = blockIdx.x;
    double m = 0.0;
    for (int64_t k = threadIdx.x; k < Kdim; k += blockDim.x) m = fmax(m, fabs(Bpack[k * Sp + j]));
    red[threadIdx.x] = m;
    __syncthreads();
    for (int s = blockDim.x >> 1; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) colmax[j] = red[0];
}
// thread = (column j < 128, 8-k chunk)
__global__ void quad_build_b_kernel(const double* __restrict__ Bpack, int64_t Kdim, int64_t Kp, int64_t Sp,
                                    const double* __restrict__ colmax, __half* __restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 128 * (Kp / 8)) return;
    const int64_t j = t % 128, kc = t / 128;
    const int64_t k0 = kc * 8, st = k0 / TC_BK, chunk = (k0 % TC_BK) / 8;
    __align__(16) __half hi[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) hi[e] = lo[e] = __float2half_rn(0.0f);
    if (j < Sp) {
        const double cm = colmax[j];
        const double sc = cm > 0.0 ? QS_BMAX / cm : 0.0;
#pragma unroll
        for (int e = 0; e < 8; ++e)
            if (k0 + e < Kdim) split_f16((float)(Bpack[(k0 + e) * Sp + j] * sc), hi[e], lo[e]);
    }
    __half* blk = out + st * (int64_t)(TC_BLOCK_BYTES / 2);
    const int64_t o = (chunk * 128 + j) * 8;
    *reinterpret_cast<uint4*>(blk + o) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(blk + 8 * 128 * 8 + o) = *reinterpret_cast<const uint4*>(lo);
}
// out[x][j] *= rowmax[x]^2 / QS_AMAX * colmax[j] / QS_BMAX is folded into the store epilogue through rowscale / colscale
__global__ void quad_scales_kernel(const double* __restrict__ rowmax, int64_t nc, double* __restrict__ rowscale,
                                   const double* __restrict__ colmax, int64_t Sp, double* __restrict__ colscale) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nc) rowscale[t] = rowmax[t] * rowmax[t] / QS_AMAX;
    if (colscale && t < Sp) colscale[t] = colmax[t] / QS_BMAX;
}

// ---- fused variant: the A operand never leaves the SM -------------------------------------------------------------
// Persistent CTAs, one 128-row tile at a time.  M[128][T] (pre-scaled by 64 / rowmax, so products stay below QS_AMAX)
// sits in shared memory as floats next to the (a, b) table; the 16 epilogue warps GENERATE each 64-k stage of A
// (products, hi/lo split, MMA layout) straight into the stage buffer (generic-proxy stores + fence.proxy.async), the
// producer warp only streams B, the MMA warp waits for both.  Every fourth stage the same warps drain the finished
// 256-k segment from TMEM into their fp64 sums.  Saves the 625 MB per chunk that the unfused path writes to and reads
// back from HBM, and its build kernel.
constexpr int QF_NS = 2;
__host__ __device__ inline int64_t qf_msm_floats(int64_t T) { return ((T + 8) * 129 + 3) & ~(int64_t)3; }
__global__ void __launch_bounds__(TC_THREADS, 1)
    quad_screen_fused_kernel(const double* __restrict__ G, int64_t ldg, int64_t nc, int64_t T,
                             const uint32_t* __restrict__ ab, int64_t Kdim, const __half* __restrict__ Bt, int64_t Sp,
                             int64_t nst, int64_t ntu, double* __restrict__ rowmax,
                             const double* __restrict__ colscale, double* __restrict__ Qout) {
    extern __shared__ __align__(1024) unsigned char qf_smem_raw[];
    unsigned char* stages = qf_smem_raw;                                            // [QF_NS][A block | B block]
    float* msm = reinterpret_cast<float*>(qf_smem_raw + QF_NS * TC_STAGE_BYTES);    // [T + 8][129], rows T.. = 0
    uint2* dsc = reinterpret_cast<uint2*>(msm + qf_msm_floats(T));    // [nst * 8] byte offsets of rows a, b0 per chunk
    double* rmax_s = reinterpret_cast<double*>(dsc + nst * 8);     // [128] row maxima of the tile (scale of M, of q~)
    uint64_t* bars = reinterpret_cast<uint64_t*>(rmax_s + 128);
    uint64_t* full_b = bars;                 // [2] producer (B block landed) -> MMA
    uint64_t* a_full = bars + QF_NS;         // [2] generators (A block written) -> MMA
    uint64_t* empty_b = bars + 2 * QF_NS;    // [2] MMA (commit) -> producer, generators
    uint64_t* acc_full = bars + 3 * QF_NS;   // [2] MMA (commit) -> epilogue
    uint64_t* acc_empty = acc_full + 2;      // [2] epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t nseg = nst / TC_SEG;
    const int64_t ntile = blockIdx.x < ntu ? (ntu - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    if (tid == 0) {
        for (int s = 0; s < QF_NS; ++s) {
            mbar_init(&full_b[s], 1);
            mbar_init(&a_full[s], TC_EPI_WARPS);
            mbar_init(&empty_b[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&acc_full[a], 1);
            mbar_init(&acc_empty[a], TC_EPI_WARPS);
        }
        mbar_fence_init();
    }
    // run order (launch_quad_screen_ab): the 8 entries of a chunk are (a, b0), (a, b0 + 1), ... or padding; rows
    // T .. T + 7 of the tile are zero, so padding needs no test
    for (int64_t c = tid; c < nst * 8; c += TC_THREADS) {
        const uint32_t v = c * 8 < Kdim ? ab[c * 8] : 0xffffffffu;
        const uint32_t a = v == 0xffffffffu ? (uint32_t)T : (v & 0xffffu), b = v == 0xffffffffu ? (uint32_t)T : (v >> 16);
        dsc[c] = make_uint2(a * 129u * 4u, b * 129u * 4u);
    }
    for (int e = tid; e < 8 * 129; e += TC_THREADS) msm[T * 129 + e] = 0.0f;
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            const unsigned char* srcB = reinterpret_cast<const unsigned char*>(Bt);
            int64_t g = 0;
            for (int64_t it = 0; it < ntile; ++it)
                for (int64_t st = 0; st < nst; ++st, ++g) {
                    const uint32_t s = (uint32_t)(g % QF_NS), ph = (uint32_t)((g / QF_NS) & 1);
                    mbar_wait_relaxed(&empty_b[s], ph ^ 1u, 32);
                    mbar_expect_tx(&full_b[s], TC_BLOCK_BYTES);
                    bulk_g2s(stages + s * TC_STAGE_BYTES + TC_BLOCK_BYTES, srcB + st * (int64_t)TC_BLOCK_BYTES,
                             TC_BLOCK_BYTES, &full_b[s]);
                }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = tc_idesc_f16(128, 128);
            const uint32_t base = smem_u32(stages);
            int64_t g = 0;
            for (int64_t gs = 0; gs < ntile * nseg; ++gs) {
                const uint32_t a = (uint32_t)(gs & 1), pa = (uint32_t)((gs >> 1) & 1);
                mbar_wait_relaxed(&acc_empty[a], pa ^ 1u, 32);
                tc_fence_after();
                for (int q4 = 0; q4 < TC_SEG; ++q4, ++g) {
                    const uint32_t s = (uint32_t)(g % QF_NS), ph = (uint32_t)((g / QF_NS) & 1);
                    mbar_wait_relaxed(&full_b[s], ph, 32);
                    mbar_wait_relaxed(&a_full[s], ph, 32);
                    tc_fence_after();
                    const uint32_t sa = base + s * TC_STAGE_BYTES, sb = sa + TC_BLOCK_BYTES;
#pragma unroll
                    for (int ks = 0; ks < TC_BK / 16; ++ks) {
                        const uint32_t off = (uint32_t)(2 * ks) * 128u * 16u;
                        const uint64_t a_hi = tc_desc(sa + off, 128 * 16, 128);
                        const uint64_t a_lo = tc_desc(sa + TC_BLOCK_BYTES / 2 + off, 128 * 16, 128);
                        const uint64_t b_hi = tc_desc(sb + off, 128 * 16, 128);
                        const uint64_t b_lo = tc_desc(sb + TC_BLOCK_BYTES / 2 + off, 128 * 16, 128);
                        const uint32_t first = (q4 == 0 && ks == 0) ? 0u : 1u;
                        tc_mma_f16(tmem + a * 128u, a_lo, b_hi, idesc, first);
                        tc_mma_f16(tmem + a * 128u, a_hi, b_lo, idesc, 1u);
                        tc_mma_f16(tmem + a * 128u, a_hi, b_hi, idesc, 1u);
                    }
                    tc_commit(&empty_b[s]);
                }
                tc_commit(&acc_full[a]);
            }
        }
        __syncwarp();
    } else {
        const int ew = warp - 2;
        const int quarter = warp & 3, cb = ew >> 2;
        const uint32_t t_lane = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(cb * 32);
        const int te = tid - 64, gr = te & 127, part = te >> 7;        // generator role: row gr, chunks 2 part, 2 part + 1
        const unsigned char* mrow = reinterpret_cast<const unsigned char*>(msm + gr);
        int64_t g = 0, gs = 0;
        for (int64_t it = 0; it < ntile; ++it) {
            const int64_t rt = blockIdx.x + it * gridDim.x;
            // (re)load the tile of M: the 16 generator warps only; the MMAs of the previous tile may still be in flight
            asm volatile("bar.sync 1, 512;" ::: "memory");
            // a warp per row (8 rows each), lanes over the T columns: one scale (one fp64 division) per row instead of
            // per element, no index divisions, coalesced independent loads the compiler can batch
            // The row maximum (quad_rowmax_kernel of the unfused path: max |M[x, a]|, NaN ignored) is formed here by the
            // same warp, so the fused path needs no separate pass over G and no rowmax launch.
            for (int rr = te >> 5; rr < 128; rr += TC_EPI_WARPS) {
                const int64_t grow = rt * 128 + rr;
                const double* gp = G + grow * ldg;
                double rm = 0.0;
                if (grow < nc)
                    for (int a = lane; a < (int)T; a += 32) rm = fmax(rm, fabs(gp[a]));
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) rm = fmax(rm, __shfl_xor_sync(0xffffffffu, rm, off));
                const double sc = rm > 0.0 ? 64.0 / rm : 0.0;                       // 64^2 = QS_AMAX
                if (lane == 0) {
                    rmax_s[rr] = rm;
                    if (grow < nc) rowmax[grow] = rm;      // the absolute term of the screen's error bound needs it
                }
#pragma unroll 4
                for (int a = lane; a < (int)T; a += 32) msm[a * 129 + rr] = sc != 0.0 ? (float)(gp[a] * sc) : 0.0f;
            }
            asm volatile("bar.sync 1, 512;" ::: "memory");
            // running sums of the 256-k fp32 segments in fp32: one rounding of 2^-24 of the partial sum (<= h^2) per
            // segment, nseg <= 64 of them -- carried by QS_ETA (criteria.cu).  fp64 sums cost a conversion on the XU
            // pipe per value (16 lanes / clk / SM), double-float pairs 7 fp32 operations per value: the drains were
            // 26 % of the kernel's samples
            float shi[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) shi[c] = 0.0f;
            auto drain = [&](int64_t dseg) {
                const uint32_t a = (uint32_t)(dseg & 1), pa = (uint32_t)((dseg >> 1) & 1);
                mbar_wait(&acc_full[a], pa);
                tc_fence_after();
                uint32_t v0[32];
                TC_LD32(v0, t_lane + a * 128u);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) tc_arrive(&acc_empty[a]);
#pragma unroll
                for (int c = 0; c < 32; ++c) shi[c] = __fadd_rn(shi[c], __uint_as_float(v0[c]));
            };
            for (int64_t seg = 0; seg < nseg; ++seg) {
                for (int q4 = 0; q4 < TC_SEG; ++q4, ++g) {
                    const int64_t st = seg * TC_SEG + q4;
                    const uint32_t s = (uint32_t)(g % QF_NS), ph = (uint32_t)((g / QF_NS) & 1);
                    uint32_t hi[8], lo[8];
#pragma unroll
                    for (int cc = 0; cc < 2; ++cc) {
                        const uint2 ds = dsc[st * 8 + 2 * part + cc];
                        const float ma = *reinterpret_cast<const float*>(mrow + ds.x);
                        const unsigned char* pb = mrow + ds.y;
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float p0 = ma * *reinterpret_cast<const float*>(pb + (2 * e) * 516);
                            const float p1 = ma * *reinterpret_cast<const float*>(pb + (2 * e + 1) * 516);
                            const __half2 h = __floats2half2_rn(p0, p1);
                            const float2 hf = __half22float2(h);
                            const __half2 l = __floats2half2_rn(p0 - hf.x, p1 - hf.y);
                            hi[4 * cc + e] = *reinterpret_cast<const uint32_t*>(&h);
                            lo[4 * cc + e] = *reinterpret_cast<const uint32_t*>(&l);
                        }
                    }
                    mbar_wait(&empty_b[s], ph ^ 1u);
                    unsigned char* ablk = stages + s * TC_STAGE_BYTES;
#pragma unroll
                    for (int cc = 0; cc < 2; ++cc) {
                        const uint32_t o = (uint32_t)((2 * part + cc) * 128 + gr) * 16u;
                        *reinterpret_cast<uint4*>(ablk + o) = make_uint4(hi[4 * cc], hi[4 * cc + 1], hi[4 * cc + 2], hi[4 * cc + 3]);
                        *reinterpret_cast<uint4*>(ablk + TC_BLOCK_BYTES / 2 + o) =
                            make_uint4(lo[4 * cc], lo[4 * cc + 1], lo[4 * cc + 2], lo[4 * cc + 3]);
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic stores -> visible to the MMA
                    __syncwarp();
                    if (lane == 0) tc_arrive(&a_full[s]);
                }
                // drain the PREVIOUS segment: its MMAs were issued a segment ago, so this wait is (nearly) free and
                // the tensor pipe keeps running on the stages just generated
                ++gs;
                if (seg > 0) drain(gs - 2);
            }
            drain(gs - 1);
            const int64_t row = rt * 128 + quarter * 32 + lane;
            if (row < nc) {
                const double rm = rmax_s[quarter * 32 + lane];
                const double rs = rm * rm / QS_AMAX;
#pragma unroll
                for (int c = 0; c < 32; ++c)
                    if (cb * 32 + c < Sp)
                        Qout[row * Sp + cb * 32 + c] = rs * colscale[cb * 32 + c] * (double)shi[c];
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256));
    }
}

// ---- triangular variant: no generated operand at all ------------------------------------------------------------------
// q_j(x) = sum_a M_a Y_ja,  Y_ja = sum_{b <= a} c_ab(j) M_b,  c_aa = Sigma_j[a][a], c_ab = Sigma_j[a][b] + Sigma_j[b][a]:
// for every a ("a-block") ONE small GEMM  C_a[x][j] = sum_{b <= a} M[x][b] c_ab(j)  with A = the tile of M itself (fp16
// hi / lo, resident in shared memory for the whole tile: no products to generate) and B_a = the a-th row block of the
// packed covariances, K = a + 1 rounded up to 32 -- the same MMA count as the index-pair formulation (the triangle),
// and the dot with M_a happens in the epilogue: acc[x][j] += M~[x][a] C_a[x][j] (one tcgen05.ld + 32 FFMA per thread
// and a-block).  ncu of the fused kernel above: bound by its generator (8192 products + splits per 64-k stage, tensor
// pipe ~35 %).  Here the 16 epilogue warps only drain.  A CTA owns TWO 128-row tiles at once: both share every B unit
// (the B stream per SM halves: 5 MB per chunk at T = 125), the four 128-column accumulators (2 tiles x 2 buffers) fill
// the tensor memory.  T <= 128, S' <= 128.
// B layout: units of 32 k x 128 columns: [hi | lo][4 chunks of 8 k][128 columns][8 halves] = 16 KB; a-block a has
// ceil((a + 1) / 32) units; unit index u(a, n) = 16 g (g + 1) + (g + 1)(a - 32 g) + n, g = a / 32.
constexpr int QT_NS = 5;                                   // B units in flight (80 KB)
constexpr uint32_t QT_UNIT_BYTES = 2 * 4 * 128 * 16;       // 16 KB
__host__ __device__ inline int64_t qt_unit_index(int64_t a, int64_t n) {
    const int64_t g = a / 32;
    return 16 * g * (g + 1) + (g + 1) * (a - 32 * g) + n;
}
__host__ __device__ inline int64_t qt_units(int64_t T) { return qt_unit_index(T, 0); }
// colmax[j] = max over the packed coefficients |c_ab(j)| (NaN ignored, like quad_colmax_kernel); one CTA per j
__global__ void __launch_bounds__(256) quad_tri_colmax_kernel(const double* __restrict__ cov, int64_t Sp, int64_t T,
                                                              double* __restrict__ colmax, double* __restrict__ colscale) {
    __shared__ double red[256];
    const int64_t j = blockIdx.x;
    const double* c = cov + j * T * T;
    double m = 0.0;
    for (int64_t e = threadIdx.x; e < T * T; e += blockDim.x) {
        const int64_t a = e / T, b = e % T;
        if (b > a) continue;
        double r = c[a * T + b];
        if (a != b) r += c[b * T + a];
        m = fmax(m, fabs(r));
    }
    red[threadIdx.x] = m;
    __syncthreads();
    for (int s = blockDim.x >> 1; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        colmax[j] = red[0];
        colscale[j] = red[0] / QS_BMAX;
    }
}
// thread = (unit, chunk of 8 k, column j): 8 coefficients, scaled to max QS_BMAX per column, hi / lo split
__global__ void quad_tri_build_b_kernel(const double* __restrict__ cov, int64_t Sp, int64_t T,
                                        const double* __restrict__ colmax, __half* __restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nu = qt_units(T);
    if (t >= nu * 4 * 128) return;
    const int64_t j = t % 128, chunk = (t / 128) % 4, u = t / 512;
    // (a, n) of unit u: invert u(a, n) group by group
    int64_t g = 0;
    while (16 * (g + 1) * (g + 2) <= u) ++g;
    const int64_t r = u - 16 * g * (g + 1), a = 32 * g + r / (g + 1), n = r % (g + 1);
    __align__(16) __half hi[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) hi[e] = lo[e] = __float2half_rn(0.0f);
    if (j < Sp && a < T) {
        const double cm = colmax[j];
        const double sc = cm > 0.0 ? QS_BMAX / cm : 0.0;
        const double* c = cov + j * T * T;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int64_t b = 32 * n + 8 * chunk + e;
            if (b <= a) {
                double v = c[a * T + b];
                if (a != b) v += c[b * T + a];
                split_f16((float)(v * sc), hi[e], lo[e]);
            }
        }
    }
    __half* blk = out + u * (int64_t)(QT_UNIT_BYTES / 2);
    const int64_t o = (chunk * 128 + j) * 8;
    *reinterpret_cast<uint4*>(blk + o) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(blk + 4 * 128 * 8 + o) = *reinterpret_cast<const uint4*>(lo);
}

// Orientation: D[j][x] = sum_b c_ab(j) M[x][b] -- the covariance unit is the A operand (M = 128 maximizers), BOTH row
// tiles of the CTA are the B operand (N = 256 candidates): ONE M128 N256 K16 instruction does what two M128 N128 K16
// did.  ncu of the first orientation (tiles as A): the MMAs were accepted at the pipe's rate (61 cycles each), but the
// single issuing thread also pays ~290 cycles per 16 KB unit (barrier check, fence, commit) -- issue, not the tensor
// pipe, bound the a-loop at 48 % pipe-busy; halving the instruction count leaves the issue thread slack.
// Accumulator: 128 lanes (j) x 256 columns (x), two buffers = the whole tensor memory.  Epilogue thread = (lane j,
// 64 columns): acc[x] += M~[x][a] D[j][x]; the multipliers of an a-block are a 256-float line in shared memory,
// written by the epilogue threads one block ahead.
__global__ void __launch_bounds__(TC_THREADS, 1)
    quad_screen_tri_kernel(const double* __restrict__ G, int64_t ldg, int64_t nc, int T, const __half* __restrict__ Bt,
                           int64_t Sp, int64_t npair, double* __restrict__ rowmax, const double* __restrict__ colscale,
                           double* __restrict__ Qout) {
    extern __shared__ __align__(1024) unsigned char qt_smem_raw[];
    unsigned char* Mop = qt_smem_raw;                                   // [2 blocks of 64 k][hi | lo][8 chunks][256 rows][8] = 128 KB
    unsigned char* Cring = qt_smem_raw + 4 * TC_BLOCK_BYTES;            // [QT_NS][16 KB] covariance units
    double* rmax_s = reinterpret_cast<double*>(Cring + QT_NS * QT_UNIT_BYTES);      // [256]
    double* cs_s = rmax_s + 256;                                                    // [128] colscale
    float* mline = reinterpret_cast<float*>(cs_s + 128);                            // [2][256] multipliers of an a-block
    uint64_t* bars = reinterpret_cast<uint64_t*>(mline + 512);
    uint64_t* full_b = bars;                 // [QT_NS] producer -> MMA
    uint64_t* empty_b = bars + QT_NS;        // [QT_NS] MMA (commit) -> producer
    uint64_t* acc_full = bars + 2 * QT_NS;   // [2] MMA (commit) -> epilogue
    uint64_t* acc_empty = acc_full + 2;      // [2] epilogue -> MMA
    uint64_t* a_full = acc_empty + 2;        // [1] epilogue warps (M operand of the pair written) -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_full + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t nmine = blockIdx.x < npair ? (npair - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int nunit_all = (int)qt_units(T);
    constexpr uint32_t MBLK = 2 * TC_BLOCK_BYTES;          // one 64-k block of the 256-row operand (hi + lo) = 64 KB
    if (tid == 0) {
        for (int s = 0; s < QT_NS; ++s) {
            mbar_init(&full_b[s], 1);
            mbar_init(&empty_b[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&acc_full[a], 1);
            mbar_init(&acc_empty[a], TC_EPI_WARPS);
        }
        mbar_init(a_full, TC_EPI_WARPS);
        mbar_fence_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid < 128) cs_s[tid] = tid < Sp ? colscale[tid] : 0.0;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            const unsigned char* srcB = reinterpret_cast<const unsigned char*>(Bt);
            uint32_t s = 0, ph = 0;
            for (int64_t it = 0; it < nmine; ++it)
                for (int u = 0; u < nunit_all; ++u) {
                    mbar_wait_relaxed(&empty_b[s], ph ^ 1u, 32);
                    mbar_expect_tx(&full_b[s], QT_UNIT_BYTES);
                    bulk_g2s(Cring + s * QT_UNIT_BYTES, srcB + (int64_t)u * QT_UNIT_BYTES, QT_UNIT_BYTES, &full_b[s]);
                    if (++s == QT_NS) {
                        s = 0;
                        ph ^= 1u;
                    }
                }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = tc_idesc_f16(128, 256);
            // shared-memory descriptors: only the start-address field (bits 0-13, units of 16 bytes) changes between
            // MMAs, so they are formed once and advanced by 32-bit adds
            const uint64_t dc = tc_desc(smem_u32(Cring), 128 * 16, 128);     // covariance unit: 128 rows per chunk
            const uint64_t dm = tc_desc(smem_u32(Mop), 256 * 16, 128);       // candidate operand: 256 rows per chunk
            const uint32_t dhi = (uint32_t)(dc >> 32), c_lo32 = (uint32_t)dc, m_lo32 = (uint32_t)dm;
            auto mk = [&](uint32_t lo32) { return ((uint64_t)dhi << 32) | (uint64_t)lo32; };
            int64_t gs = 0;
            uint32_t s = 0, ph = 0;                                   // ring slot and phase
            for (int64_t it = 0; it < nmine; ++it) {
                mbar_wait_relaxed(a_full, (uint32_t)(it & 1), 32);
                tc_fence_after();
                for (int a = 0; a < T; ++a, ++gs) {
                    const uint32_t buf = (uint32_t)(gs & 1), pa = (uint32_t)((gs >> 1) & 1);
                    mbar_wait_relaxed(&acc_empty[buf], pa ^ 1u, 32);
                    tc_fence_after();
                    const int nun = a / 32 + 1;
                    const uint32_t d = tmem + buf * 256u;
                    for (int n = 0; n < nun; ++n) {
                        mbar_wait_relaxed(&full_b[s], ph, 32);
                        tc_fence_after();
                        const uint32_t sc = c_lo32 + s * (QT_UNIT_BYTES >> 4);
                        const int nk16 = (a + 1 - 32 * n) > 16 ? 2 : 1;       // the second half of the unit may be all zero
                        for (int h = 0; h < nk16; ++h) {
                            const int kk = 2 * n + h;                          // K16 step of the candidate operand (0 .. 7)
                            const uint32_t sm_ = m_lo32 + (uint32_t)(kk >> 2) * (MBLK >> 4) + (uint32_t)(kk & 3) * 512u;
                            const uint64_t c_hi = mk(sc + (uint32_t)h * 256u), c_lo = mk(sc + (uint32_t)h * 256u + (QT_UNIT_BYTES >> 5));
                            const uint64_t m_hi = mk(sm_), m_lo = mk(sm_ + (MBLK >> 5));
                            tc_mma_f16(d, c_lo, m_hi, idesc, (n == 0 && h == 0) ? 0u : 1u);
                            tc_mma_f16(d, c_hi, m_lo, idesc, 1u);
                            tc_mma_f16(d, c_hi, m_hi, idesc, 1u);
                        }
                        tc_commit(&empty_b[s]);
                        if (++s == QT_NS) {
                            s = 0;
                            ph ^= 1u;
                        }
                    }
                    tc_commit(&acc_full[buf]);
                }
            }
        }
        __syncwarp();
    } else {
        const int ew = warp - 2;
        const int quarter = warp & 3, cg = ew >> 2;                // TMEM lane quarter, group of 64 columns
        const uint32_t t_lane = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(cg * 64);
        const int te = tid - 64;
        const int j = quarter * 32 + lane;                         // this thread's maximizer (TMEM lane)
        int64_t gs = 0;
        // multiplier M~[x][a] = hi + lo exactly as the tensor core saw it; thread te < 256 owns x = te
        auto write_line = [&](int a) {
            if (te < 256) {
                const unsigned char* p = Mop + (a >> 6) * MBLK + ((((a & 63) >> 3) * 256 + te) * 8 + (a & 7)) * 2;
                mline[(a & 1) * 256 + te] = __half2float(*reinterpret_cast<const __half*>(p)) +
                                            __half2float(*reinterpret_cast<const __half*>(p + MBLK / 2));
            }
        };
        for (int64_t it = 0; it < nmine; ++it) {
            const int64_t pr = blockIdx.x + it * gridDim.x;
            // ---- candidate operand of the pair: a warp per row, lanes over the T columns; row maximum, scale 64 / max,
            // hi / lo halves straight into the MMA layout; k >= T and rows >= nc stay zero
            asm volatile("bar.sync 1, 512;" ::: "memory");        // every warp is done reading the previous pair's M~
            for (int e = te; e < (int)(4 * TC_BLOCK_BYTES / 16); e += 512) reinterpret_cast<uint4*>(Mop)[e] = make_uint4(0, 0, 0, 0);
            asm volatile("bar.sync 1, 512;" ::: "memory");
            for (int r0 = te >> 5; r0 < 256; r0 += 2 * TC_EPI_WARPS) {
                double gv[2][4], rm[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int rr = r0 + u * TC_EPI_WARPS;
                    const int64_t grow = pr * 256 + rr;
                    const double* gp = G + grow * ldg;
                    rm[u] = 0.0;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int a = lane + 32 * q;
                        gv[u][q] = (grow < nc && a < T) ? gp[a] : 0.0;
                        rm[u] = fmax(rm[u], fabs(gv[u][q]));      // fmax ignores NaN, like the row maximum of the other paths
                    }
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    rm[0] = fmax(rm[0], __shfl_xor_sync(0xffffffffu, rm[0], off));
                    rm[1] = fmax(rm[1], __shfl_xor_sync(0xffffffffu, rm[1], off));
                }
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int rr = r0 + u * TC_EPI_WARPS;
                    const int64_t grow = pr * 256 + rr;
                    const double sc = rm[u] > 0.0 ? 64.0 / rm[u] : 0.0;                 // 64^2 = QS_AMAX
                    if (lane == 0) {
                        rmax_s[rr] = rm[u];
                        if (grow < nc) rowmax[grow] = rm[u];
                    }
                    if (grow < nc && sc != 0.0) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int a = lane + 32 * q;
                            if (a < T) {
                                __half h, l;
                                split_f16((float)(gv[u][q] * sc), h, l);
                                unsigned char* p = Mop + (a >> 6) * MBLK + ((((a & 63) >> 3) * 256 + rr) * 8 + (a & 7)) * 2;
                                *reinterpret_cast<__half*>(p) = h;
                                *reinterpret_cast<__half*>(p + MBLK / 2) = l;
                            }
                        }
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic stores -> visible to the MMA
            __syncwarp();
            if (lane == 0) tc_arrive(a_full);
            asm volatile("bar.sync 1, 512;" ::: "memory");        // M~ of every row is readable below

            float acc[64];
#pragma unroll
            for (int c = 0; c < 64; ++c) acc[c] = 0.0f;
            write_line(0);
            for (int a = 0; a < T; ++a, ++gs) {
                const uint32_t buf = (uint32_t)(gs & 1), pa = (uint32_t)((gs >> 1) & 1);
                asm volatile("bar.sync 1, 512;" ::: "memory");    // line a complete; line a - 1 no longer read
                if (a + 1 < T) write_line(a + 1);
                const float* ml = mline + (a & 1) * 256 + cg * 64;
                mbar_wait(&acc_full[buf], pa);
                tc_fence_after();
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint32_t v0[16];
                    TC_LD16(v0, t_lane + buf * 256u + (uint32_t)(q * 16));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) {
                        const float4 m4 = *reinterpret_cast<const float4*>(ml + q * 16 + c4 * 4);
                        acc[q * 16 + 4 * c4] = fmaf(m4.x, __uint_as_float(v0[4 * c4]), acc[q * 16 + 4 * c4]);
                        acc[q * 16 + 4 * c4 + 1] = fmaf(m4.y, __uint_as_float(v0[4 * c4 + 1]), acc[q * 16 + 4 * c4 + 1]);
                        acc[q * 16 + 4 * c4 + 2] = fmaf(m4.z, __uint_as_float(v0[4 * c4 + 2]), acc[q * 16 + 4 * c4 + 2]);
                        acc[q * 16 + 4 * c4 + 3] = fmaf(m4.w, __uint_as_float(v0[4 * c4 + 3]), acc[q * 16 + 4 * c4 + 3]);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) tc_arrive(&acc_empty[buf]);
            }
            if (j < Sp) {
                const double cs = cs_s[j];
#pragma unroll
                for (int c = 0; c < 64; ++c) {
                    const int x = cg * 64 + c;
                    const int64_t row = pr * 256 + x;
                    if (row < nc) {
                        const double rm = rmax_s[x];
                        Qout[row * Sp + j] = rm * rm / QS_AMAX * cs * (double)acc[c];
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
    }
}

int64_t quad_tri_b_bytes(int64_t T) { return align_up(qt_units(T) * (int64_t)QT_UNIT_BYTES, 256); }
int64_t quad_tri_terms(int64_t T) { return T * (T + 1) / 2; }   // (a, b <= a) terms; the padding is exactly zero in both operands
// Btri: quad_tri_b_bytes(T); colmax: 256 doubles [colmax (128) | colscale (128)]
int launch_quad_tri_build_b(const double* cov, int64_t Sp, int64_t T, void* Btri, double* colmax, cudaStream_t st) {
    if (T < 1 || T > 128 || Sp < 1 || Sp > 128) return -5;
    quad_tri_colmax_kernel<<<(unsigned)Sp, 256, 0, st>>>(cov, Sp, T, colmax, colmax + 128);
    TES_LAUNCH_OK();
    const int64_t tot = qt_units(T) * 4 * 128;
    quad_tri_build_b_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(cov, Sp, T, colmax, (__half*)Btri);
    TES_LAUNCH_OK();
    return 0;
}
int launch_quad_screen_tri(const double* G, int64_t ldg, int64_t nc, int64_t T, const void* Btri, const double* colmax,
                           int64_t Sp, double* rowmax, double* Qout, cudaStream_t st) {
    if (T < 1 || T > 128 || Sp < 1 || Sp > 128) return -5;
    const size_t smem = 4 * (size_t)TC_BLOCK_BYTES + QT_NS * (size_t)QT_UNIT_BYTES + (256 + 128) * sizeof(double) + 512 * sizeof(float) + 256;
    TES_CUDA_OK(cudaFuncSetAttribute(quad_screen_tri_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t npair = (nc + 255) / 256;
    const int sms = tc_sm_count();
    const unsigned grid = (unsigned)(npair < sms ? npair : sms);
    quad_screen_tri_kernel<<<grid, TC_THREADS, smem, st>>>(G, ldg, nc, (int)T, (const __half*)Btri, Sp, npair, rowmax,
                                                          colmax + 128, Qout);
    TES_LAUNCH_OK();
    return 0;
}

int64_t quad_screen_kp(int64_t Kdim) { return (Kdim + 255) / 256 * 256; }
int64_t quad_screen_a_bytes(int64_t nc, int64_t Kdim) {
    return align_up(((nc + 127) / 128) * (quad_screen_kp(Kdim) / TC_BK) * (int64_t)TC_BLOCK_BYTES, 256) + align_up(nc * 8, 256);
}
int64_t quad_screen_b_bytes(int64_t Kdim) {
    return align_up((quad_screen_kp(Kdim) / TC_BK) * (int64_t)TC_BLOCK_BYTES, 256) + 2 * align_up(128 * 8, 256) +
           align_up(quad_screen_kp(Kdim) * 4, 256);
}
// Btc layout: [operand blocks][colmax (128)][colscale (128)]
int launch_quad_screen_build_b(const double* Bpack, int64_t Kdim, int64_t Sp, void* Btc, double* colmax, cudaStream_t st) {
    const int64_t Kp = quad_screen_kp(Kdim);
    double* colscale = colmax + 128;
    quad_colmax_kernel<<<(unsigned)Sp, 256, 0, st>>>(Bpack, Kdim, Sp, colmax);
    TES_LAUNCH_OK();
    const int64_t tot = 128 * (Kp / 8);
    quad_build_b_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(Bpack, Kdim, Kp, Sp, colmax, (__half*)Btc);
    TES_LAUNCH_OK();
    quad_scales_kernel<<<1, 128, 0, st>>>(colmax, 0, nullptr, colmax, Sp, colscale);
    TES_LAUNCH_OK();
    return 0;
}
int64_t quad_screen_kdim(int64_t T) {
    int64_t k = 0;
    for (int64_t a = 0; a < T; ++a) k += (T - a + 7) / 8 * 8;
    return k;
}
// (a, b) table of the K axis (run order), stored behind colmax / colscale in the B region; Bpack in the same order
int launch_quad_screen_ab(const double* cov, int64_t Sp, int64_t T, double* colmax, double* Bpack, cudaStream_t st) {
    if (T > 0xffff) return -5;
    const int64_t Kdim = quad_screen_kdim(T), Kp = quad_screen_kp(Kdim);
    uint32_t* ab = reinterpret_cast<uint32_t*>(colmax + 256);
    quad_ab_runs_kernel<<<1, 256, 0, st>>>(T, Kp, ab);
    TES_LAUNCH_OK();
    quad_pack_ab_kernel<<<(unsigned)((Kdim * Sp + 255) / 256), 256, 0, st>>>(cov, Sp, T, ab, Kdim, Bpack);
    TES_LAUNCH_OK();
    return 0;
}
// Atc: [operand blocks of the chunk][rowscale (nc)]; rowmax: nc doubles of scratch
int launch_quad_screen(const double* G, int64_t ldg, int64_t nc, int64_t T, int64_t Kdim, const void* Btc,
                       const double* colmax, int64_t Sp, void* Atc, double* rowmax, double* Qout, cudaStream_t st) {
    const int64_t Kp = quad_screen_kp(Kdim), nst = Kp / TC_BK, ntu = (nc + 127) / 128;
    double* rowscale = (double*)((char*)Atc + align_up(ntu * nst * (int64_t)TC_BLOCK_BYTES, 256));
    const double* colscale = colmax + 128;
    {
        // fused kernel (A generated on the SM, row maxima included) whenever M's tile fits beside two stages;
        // TES_QUAD_FUSED=0 disables it
        const size_t fsm = (size_t)QF_NS * TC_STAGE_BYTES + (size_t)qf_msm_floats(T) * sizeof(float) +
                           (size_t)nst * 8 * sizeof(uint2) + 128 * sizeof(double) + 256;
        const char* env = getenv("TES_QUAD_FUSED");
        if (fsm <= 227 * 1024 && T <= 500 && nst / TC_SEG <= 64 && !(env && env[0] == '0')) {
            const uint32_t* ab = reinterpret_cast<const uint32_t*>(colmax + 256);
            const int sms = tc_sm_count();
            TES_CUDA_OK(cudaFuncSetAttribute(quad_screen_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)fsm));
            const unsigned grid = (unsigned)(ntu < sms ? ntu : sms);
            quad_screen_fused_kernel<<<grid, TC_THREADS, fsm, st>>>(G, ldg, nc, T, ab, Kdim, (const __half*)Btc, Sp, nst, ntu,
                                                                   rowmax, colscale, Qout);
            TES_LAUNCH_OK();
            return 0;
        }
    }
    quad_rowmax_kernel<<<(unsigned)((nc * 32 + 255) / 256), 256, 0, st>>>(G, ldg, nc, T, rowmax);
    TES_LAUNCH_OK();
    {
        const size_t smem = (size_t)T * 129 * sizeof(float);
        if (smem > 200 * 1024) return -5;
        TES_CUDA_OK(cudaFuncSetAttribute(quad_build_a_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const uint32_t* ab = reinterpret_cast<const uint32_t*>(colmax + 256);
        quad_build_a_kernel<<<dim3((unsigned)ntu, 6), 256, smem, st>>>(G, ldg, nc, T, ab, Kdim, Kp, rowmax, (__half*)Atc);
        TES_LAUNCH_OK();
    }
    quad_scales_kernel<<<(unsigned)((nc + 255) / 256), 256, 0, st>>>(rowmax, nc, rowscale, nullptr, 0, nullptr);
    TES_LAUNCH_OK();
    TES_CUDA_OK(cudaFuncSetAttribute(prod_gemm_topk_tc_kernel<0, TC_NS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)(TC_NS * TC_STAGE_BYTES + 256)));
    dim3 grid((unsigned)(ntu < tc_sm_count() ? ntu : tc_sm_count()), 1, 1);
    prod_gemm_topk_tc_kernel<0, TC_NS><<<grid, TC_THREADS, TC_NS * TC_STAGE_BYTES + 256, st>>>(
        (const __half*)Atc, (const __half*)Btc, nc, Sp, ntu, 1, nst, 1.0, nullptr, 0, 0, 1, nullptr, nullptr, Qout, Sp,
        rowscale, colscale);
    TES_LAUNCH_OK();
    return 0;
}

// tmax[j] = max_f |theta[j][f]| (one block per sample)
__global__ void prod_tmax_kernel(const double* __restrict__ theta, int64_t F, double* __restrict__ tmax) {
    __shared__ double red[256];
    const int64_t j = blockIdx.x;
    double m = 0.0;
    for (int64_t f = threadIdx.x; f < F; f += blockDim.x) m = fmax(m, fabs(theta[j * F + f]));
    red[threadIdx.x] = m;
    __syncthreads();
    for (int s = blockDim.x >> 1; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) tmax[j] = red[0];
}

// merge `nlist` partial lists per sample: vals/idx (nlist, G, kp) -> (G, kp), one 256-thread CTA per sample, kp rounds of
// "best entry strictly after the previous winner" (same ordering rule as everywhere: value desc, index asc)
__global__ void __launch_bounds__(256)
    prod_merge_kernel(const double* __restrict__ vals, const int64_t* __restrict__ idx, int64_t nlist, int64_t G, int kp,
                      double* __restrict__ out_val, int64_t* __restrict__ out_idx) {
    __shared__ double sv[8];
    __shared__ int64_t si[8];
    __shared__ double wv;
    __shared__ int64_t wi;
    const int64_t g = blockIdx.x;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int64_t n = nlist * kp;
    double pv = pos_inf();
    int64_t pi = -1;
    for (int r = 0; r < kp; ++r) {
        double bv = neg_inf();
        int64_t bi = INT64_MAX;
        for (int64_t e = t; e < n; e += 256) {
            const int64_t p = e / kp, q = e % kp;
            const double v = vals[(p * G + g) * kp + q];
            const int64_t i = idx[(p * G + g) * kp + q];
            if (i >= 0 && better(pv, pi, v, i) && better(v, i, bv, bi)) {
                bv = v;
                bi = i;
            }
        }
        warp_argmax(bv, bi);
        if (lane == 0) {
            sv[warp] = bv;
            si[warp] = bi;
        }
        __syncthreads();
        if (t == 0) {
            double v = sv[0];
            int64_t i = si[0];
            for (int w = 1; w < 8; ++w)
                if (si[w] != INT64_MAX && (i == INT64_MAX || better(sv[w], si[w], v, i))) {
                    v = sv[w];
                    i = si[w];
                }
            const bool found = (i != INT64_MAX);
            out_val[g * kp + r] = found ? v : neg_inf();
            out_idx[g * kp + r] = found ? i : -1;
            wv = v;
            wi = i;
        }
        __syncthreads();
        if (wi == INT64_MAX) {   // exhausted: the remaining slots are empty
            for (int q = r + 1 + t; q < kp; q += 256) {
                out_val[g * kp + q] = neg_inf();
                out_idx[g * kp + q] = -1;
            }
            break;
        }
        pv = wv;
        pi = wi;
        __syncthreads();
    }
}

// Two-level form of the merge for nlist <= 64 * 64 (every mesh the tcgen05 path sees: 1024 lists at C3).  One CTA per
// sample leaves 148 - G SMs idle and walks 9 k entries kp times; here level 1 runs one CTA per (sample, block of Lc
// lists), level 2 one CTA per sample over the level-1 results.  Each CTA reads its <= 1024 entries ONCE into registers
// and runs the same kp rounds ("best entry strictly after the previous winner": value desc, index asc) on them, so the
// result is the one prod_merge_kernel gives.  Level 1 works in place: the winners of block c replace list c * Lc of the
// sample (written after the last round; no other CTA touches the lists of this (sample, block)).
constexpr int MERGE_LC = 64, MERGE_EPT = MERGE_LC * PROD_MAX_KP / 256;
template <bool LEVEL1>
__global__ void __launch_bounds__(256)
    prod_merge_regs_kernel(double* __restrict__ vals, int64_t* __restrict__ idx, int64_t nlist, int64_t G, int kp,
                           int64_t Lc, double* __restrict__ out_val, int64_t* __restrict__ out_idx) {
    __shared__ double sv[8], rv[PROD_MAX_KP];
    __shared__ int64_t si[8], ri[PROD_MAX_KP];
    __shared__ double wv;
    __shared__ int64_t wi;
    const int64_t g = blockIdx.x;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int64_t p0 = LEVEL1 ? (int64_t)blockIdx.y * Lc : 0;
    const int64_t np = LEVEL1 ? (nlist - p0 < Lc ? nlist - p0 : Lc) : (nlist + Lc - 1) / Lc;
    const int64_t pstride = LEVEL1 ? 1 : Lc;
    const int64_t n = np * kp;
    double ev[MERGE_EPT];
    int64_t ei[MERGE_EPT];
#pragma unroll
    for (int u = 0; u < MERGE_EPT; ++u) {
        const int64_t e = t + u * 256;
        ev[u] = 0.0;
        ei[u] = -1;
        if (e < n) {
            const int64_t o = ((p0 + (e / kp) * pstride) * G + g) * kp + e % kp;
            ev[u] = vals[o];
            ei[u] = idx[o];
        }
    }
    if (t < kp) {
        rv[t] = neg_inf();
        ri[t] = -1;
    }
    double pv = pos_inf();
    int64_t pi = -1;
    for (int r = 0; r < kp; ++r) {
        double bv = neg_inf();
        int64_t bi = INT64_MAX;
#pragma unroll
        for (int u = 0; u < MERGE_EPT; ++u)
            if (ei[u] >= 0 && better(pv, pi, ev[u], ei[u]) && better(ev[u], ei[u], bv, bi)) {
                bv = ev[u];
                bi = ei[u];
            }
        warp_argmax(bv, bi);
        if (lane == 0) {
            sv[warp] = bv;
            si[warp] = bi;
        }
        __syncthreads();
        if (t == 0) {
            double v = sv[0];
            int64_t i = si[0];
            for (int w = 1; w < 8; ++w)
                if (si[w] != INT64_MAX && (i == INT64_MAX || better(sv[w], si[w], v, i))) {
                    v = sv[w];
                    i = si[w];
                }
            if (i != INT64_MAX) {
                rv[r] = v;
                ri[r] = i;
            }
            wv = v;
            wi = i;
        }
        __syncthreads();
        if (wi == INT64_MAX) break;   // exhausted: the remaining slots stay empty
        pv = wv;
        pi = wi;
    }
    __syncthreads();
    if (t < kp) {
        if (LEVEL1) {
            vals[(p0 * G + g) * kp + t] = rv[t];
            idx[(p0 * G + g) * kp + t] = ri[t];
        } else {
            out_val[g * kp + t] = rv[t];
            out_idx[g * kp + t] = ri[t];
        }
    }
}
// TES_PROD_MERGE=0 keeps the one-CTA-per-sample kernel (comparison path of the tests)
static int launch_prod_merge(double* pv, int64_t* pi, int64_t nlist, int64_t G, int kp, double* out_val, int64_t* out_idx,
                             cudaStream_t st) {
    const char* env = getenv("TES_PROD_MERGE");
    const bool regs = !(env && env[0] == '0') && kp <= PROD_MAX_KP;
    if (regs && nlist * kp <= 256 * MERGE_EPT) {
        prod_merge_regs_kernel<false><<<(unsigned)G, 256, 0, st>>>(pv, pi, nlist, G, kp, 1, out_val, out_idx);
        TES_LAUNCH_OK();
    } else if (regs && nlist <= (int64_t)MERGE_LC * MERGE_LC) {
        const int64_t C = (nlist + MERGE_LC - 1) / MERGE_LC;
        prod_merge_regs_kernel<true><<<dim3((unsigned)G, (unsigned)C), 256, 0, st>>>(pv, pi, nlist, G, kp, MERGE_LC, nullptr,
                                                                                    nullptr);
        TES_LAUNCH_OK();
        prod_merge_regs_kernel<false><<<(unsigned)G, 256, 0, st>>>(pv, pi, nlist, G, kp, MERGE_LC, out_val, out_idx);
        TES_LAUNCH_OK();
    } else {
        prod_merge_kernel<<<(unsigned)G, 256, 0, st>>>(pv, pi, nlist, G, kp, out_val, out_idx);
        TES_LAUNCH_OK();
    }
    return 0;
}

// exact spec values of the kp survivors of every sample (one warp per (sample, slot): the sequence of
// rff_value_kernel_generic), and the safety flag of the sample (slot 0).
__global__ void prod_exact_kernel(const double* __restrict__ cand, int d, const double* __restrict__ feat, int rec,
                                  int64_t S, int64_t F, double scale, int64_t idx_offset, int k, int kp,
                                  const double* __restrict__ approx_val, const int64_t* __restrict__ idx,
                                  const double* __restrict__ xmax, int tf32, const double* __restrict__ tmax,
                                  double* __restrict__ exact_val,
                                  int* __restrict__ flag, double* __restrict__ thr_out = nullptr) {
    // one WARP per (sample, slot): the lanes evaluate 32 features' arguments and cosines in parallel and park
    // (theta, cos) in shared memory, then every lane replays the spec's sequential fma accumulation over them in feature
    // order -- the same operations in the same order as one thread walking the F features (which left 2 warps per SM
    // and 1.3 ms of dependent-latency at C3), so the value is still the bits of rff_value_kernel_generic
    __shared__ double2 pr[4][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t t = (int64_t)blockIdx.x * 4 + warp;
    if (t >= S * kp) return;
    const int64_t j = t / kp;
    const int r = (int)(t % kp);
    const double* fj = feat + j * F * rec;
    const int64_t gi = idx[t];
    double val = neg_inf();
    if (gi >= 0) {
        const double* x = cand + (gi - idx_offset) * d;
        double acc = 0.0;
        for (int64_t f0 = 0; f0 < F; f0 += 32) {
            const int64_t f = f0 + lane;
            double2 tc = make_double2(0.0, 0.0);
            if (f < F) {
                const double* rr = fj + f * rec;
                double h = __ldg(rr + d);
                for (int kk = 0; kk < d; ++kk) h = fma(__ldg(rr + kk), __ldg(x + kk), h);
                tc = make_double2(__ldg(rr + d + 1), cospi_spec(h));
            }
            __syncwarp();
            pr[warp][lane] = tc;
            __syncwarp();
            const int nb = F - f0 < 32 ? (int)(F - f0) : 32;
            for (int q = 0; q < nb; ++q) {
                const double2 e = pr[warp][q];
                acc = fma(e.x, e.y, acc);
            }
        }
        val = __dmul_rn(scale, acc);
    }
    if (lane == 0) exact_val[t] = val;
    if (r == 0) {
        // B_j: both values are fp64 evaluations of the same real function.  Per feature the argument carries
        // <= 4 (d + 2) u A rounding (prescale by 1/pi, fma chains on both sides; A = largest |argument| of the sample),
        // cos / sincos and the products <= 16 u, and the two accumulations (F sequential fmas; 2F DMMA steps)
        // <= 3 F u sum|theta|.  Doubled for margin.
        double sa = 0.0, am = 0.0;
        for (int64_t f = lane; f < F; f += 32) {
            const double* rr = fj + f * rec;
            sa += fabs(__ldg(rr + d + 1));
            double a = fabs(__ldg(rr + d));
            for (int kk = 0; kk < d; ++kk) a = fma(fabs(__ldg(rr + kk)), xmax[kk], a);
            am = a > am ? a : am;
        }
        sa = warp_sum(sa);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double o = __shfl_xor_sync(0xffffffffu, am, off);
            am = o > am ? o : am;
        }
        if (lane != 0) return;
        const double u = 1.1102230246251565e-16;
        double B = 2.0 * u * scale * sa * (4.0 * (d + 2) * 3.141592653589793 * am + 3.0 * (double)F + 16.0) + 1e-300;
        // 3xTF32 screen: every operand is sincosf of the fp64-reduced argument (argument rounding 2^-22.3, sincosf 2 ulp
        // = 2^-22; the tcgen05 build uses __sincosf: 2^-21) carried as a hi + lo pair (residual 2^-21): <= 2^-19 per
        // operand pair, plus the dropped lo*lo (2^-22): <= 4.5 * 2^-20 per term; fp32 accumulation in segments of PT_FLUSH k = 3 * PT_FLUSH / 8 mma steps,
        // each assumed to err by <= 4 * 2^-24 of the segment's sum of |terms| (the tests compare against the fp64 GEMM
        // path on ties 1e-13 ... 1e-4 apart); the |terms| of a feature sum to <= |theta_f|.  x1.5 margin.
        // (+ 0.25 * 2^-20: the tcgen05 build forms theta / tmax * cos as a product of floats, 2 extra roundings of 2^-24)
        if (tf32) B += 1.5 * scale * sa * (4.75 * 9.5367431640625e-07 + 4.0 * 5.9604644775390625e-08 * 3.0 * (PT_FLUSH / 8));
        // fp16 operands (mode 2; same 11-bit significand as TF32, theta normalised by tmax): subnormal rounding adds an
        // absolute 2^-25 per operand, i.e. <= 2 * 2^-25 per term over 2F terms, in units of scale * tmax
        if (tf32 >= 2) B += 1.5 * scale * tmax[j] * 4.0 * (double)F * 2.98023223876953125e-08;
        // modes 4 / 5 / 8 add the 2F / 256 segment values in fp32: one rounding of 2^-24 of the sum of |terms| per segment
        if (tf32 == 4 || tf32 == 5 || tf32 == 8) B += 1.5 * scale * sa * 5.9604644775390625e-08 * (double)((2 * F + 255) / 256);
        const double vk = approx_val[j * kp + (k - 1)], vl = approx_val[j * kp + (kp - 1)];
        const bool full = idx[j * kp + (kp - 1)] >= 0;   // fewer than kp candidates in total: nothing was cut
        const bool enough = idx[j * kp + (k - 1)] >= 0;  // NaN values are never selected: let the general kernel decide
        // safe iff the kp-th best GEMM value is below the k-th best by more than 2B
        const bool safe = !full || (vl < vk - 2.0 * B);
        flag[j] = (safe && enough && vk == vk && B == B) ? 0 : 1;
        // every member of the true top-k has a GEMM value >= vk - 2B (resolve pass); NaN: not resolvable
        if (thr_out) thr_out[j] = (enough && vk == vk && B == B) ? vk - 2.0 * B : __longlong_as_double(0x7ff8000000000000LL);
    }
}

// ---- resolve pass (tes_rff_topk_product_resolve_f64): samples the screen could not certify ---------------------------
// exact spec values of the emitted candidates (one warp per (sample, slot), the sequence of prod_exact_kernel)
__global__ void prod_exact_list_kernel(const double* __restrict__ cand, int d, const double* __restrict__ feat, int rec,
                                       int64_t S, int64_t F, double scale, int64_t idx_offset, int cap,
                                       const int64_t* __restrict__ idx, const int* __restrict__ cnt,
                                       double* __restrict__ exact_val) {
    __shared__ double2 pr[4][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t t = (int64_t)blockIdx.x * 4 + warp;
    if (t >= S * cap) return;
    const int64_t j = t / cap;
    const int r = (int)(t % cap);
    const int c = cnt[j];
    if (r >= (c < cap ? c : cap)) return;
    const double* fj = feat + j * F * rec;
    const double* x = cand + (idx[t] - idx_offset) * d;
    double acc = 0.0;
    for (int64_t f0 = 0; f0 < F; f0 += 32) {
        const int64_t f = f0 + lane;
        double2 tc = make_double2(0.0, 0.0);
        if (f < F) {
            const double* rr = fj + f * rec;
            double h = __ldg(rr + d);
            for (int kk = 0; kk < d; ++kk) h = fma(__ldg(rr + kk), __ldg(x + kk), h);
            tc = make_double2(__ldg(rr + d + 1), cospi_spec(h));
        }
        __syncwarp();
        pr[warp][lane] = tc;
        __syncwarp();
        const int nb = F - f0 < 32 ? (int)(F - f0) : 32;
        for (int q = 0; q < nb; ++q) {
            const double2 e = pr[warp][q];
            acc = fma(e.x, e.y, acc);
        }
    }
    if (lane == 0) exact_val[t] = __dmul_rn(scale, acc);
}
// top-k of a sample's emitted candidates by (exact value desc, index asc); one CTA per sample.  flag = 1 when the list
// overflowed or the threshold was NaN (the caller falls back to the general kernel for that sample).
__global__ void __launch_bounds__(256) prod_final_list_kernel(const double* __restrict__ vals, const int64_t* __restrict__ idx,
                                                              const int* __restrict__ cnt, const double* __restrict__ thr,
                                                              int cap, int k, double* __restrict__ out_val,
                                                              int64_t* __restrict__ out_idx, int* __restrict__ flag) {
    __shared__ double sv[8];
    __shared__ int64_t si[8];
    const int64_t j = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = cnt[j];
    const bool ok = c <= cap && thr[j] == thr[j];
    if (threadIdx.x == 0) flag[j] = ok ? 0 : 1;
    const int n = c < cap ? c : cap;
    double pv = pos_inf();
    int64_t pi = -1;
    for (int r = 0; r < k; ++r) {
        double bv = neg_inf();
        int64_t bi = INT64_MAX;
        for (int e = threadIdx.x; e < n; e += 256) {
            const double v = vals[j * cap + e];
            const int64_t i = idx[j * cap + e];
            if (better(pv, pi, v, i) && better(v, i, bv, bi)) {
                bv = v;
                bi = i;
            }
        }
        warp_argmax(bv, bi);
        if (lane == 0) {
            sv[warp] = bv;
            si[warp] = bi;
        }
        __syncthreads();
        bv = sv[0];
        bi = si[0];
        for (int q = 1; q < 8; ++q)
            if (si[q] != INT64_MAX && (bi == INT64_MAX || better(sv[q], si[q], bv, bi))) {
                bv = sv[q];
                bi = si[q];
            }
        __syncthreads();
        const bool found = (bi != INT64_MAX);
        if (threadIdx.x == 0) {
            out_val[j * k + r] = found ? bv : neg_inf();
            out_idx[j * k + r] = found ? bi : -1;
        }
        if (found) {
            pv = bv;
            pi = bi;
        }
    }
}

// xmax[k] = max_i |cand[i][k]| over the product set: the slow coordinates only vary with i / nv, the rest with i % nv
__global__ void prod_xmax_kernel(const double* __restrict__ cand, int64_t nu, int64_t nv, int d, int s_lo, int s_hi,
                                 double* __restrict__ xmax) {
    __shared__ double red[256];
    for (int k = 0; k < d; ++k) {
        double m = 0.0;
        if (k >= s_lo && k < s_hi) {
            for (int64_t i = threadIdx.x; i < nu; i += blockDim.x) m = fmax(m, fabs(cand[(i * nv) * d + k]));
        } else {
            for (int64_t i = threadIdx.x; i < nv; i += blockDim.x) m = fmax(m, fabs(cand[i * d + k]));
        }
        red[threadIdx.x] = m;
        __syncthreads();
        for (int s = blockDim.x >> 1; s > 0; s >>= 1) {
            if ((int)threadIdx.x < s) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + s]);
            __syncthreads();
        }
        if (threadIdx.x == 0) xmax[k] = red[0];
        __syncthreads();
    }
}

struct ProdPlan {
    int64_t nu, G, gx, gy, nlist, nup, nvp, K2p;
    int rec, kp;
    int64_t off_feat, off_p, off_q, off_pv, off_pi, off_mv, off_mi, off_ev, off_xmax, off_tmax, off_featp, total;
};

static ProdPlan prod_plan(int64_t N, int64_t d, int64_t nv, int64_t S, int64_t F, int k, int tf32 = 0, int tc = 0,
                          int gen_ncol = 0, int gen_gc = 8) {
    // gen_ncol > 0: modes 4 / 5 -- the P operand is generated on the SM from feature records of gen_ncol slow columns:
    // entirely (gen_gc = 8: no P array) or the first gen_gc of the 8 chunks of every stage (gen_gc = 4: P array as in mode 3)
    ProdPlan p;
    const bool no_p = gen_ncol && gen_gc == 8;
    const bool pairs = tc >= 2;                  // modes 6 / 8: CTA pairs, 256-row tile pairs (an odd tile count is padded)
    p.nu = N / nv;
    p.rec = prod_rec((int)d);
    p.kp = k + 8 > PROD_MAX_KP ? PROD_MAX_KP : k + 8;
    p.nup = (p.nu + PT_BM - 1) / PT_BM * PT_BM;
    p.nvp = (nv + PT_BN - 1) / PT_BN * PT_BN;
    p.K2p = tc ? (2 * F + 255) / 256 * 256 : (2 * F + PH_BK - 1) / PH_BK * PH_BK;
    if (tc) p.nvp = (nv + 127) / 128 * 128;
    if (pairs) p.nup = (p.nu + 255) / 256 * 256;
    if (tc == 3) p.nvp = (nv + 255) / 256 * 256;   // mode 8: 256-column tile pairs too
    const int64_t per_sample = no_p ? p.nvp * p.K2p * 8 : (tf32 ? (p.nup + p.nvp) * p.K2p * 8 : (p.nu + nv) * 2 * F * 8);
    int64_t G = per_sample > 0 ? (int64_t)(1.5e9 / (double)per_sample) : 1;
    if (G < 1) G = 1;
    if (G > S) G = S;
    if (G > 65535) G = 65535;
    if (tc && S > G) {
        // the tcgen05 GEMM is persistent (one CTA per SM walks tiles-per-sample * G tiles): pick the group size in
        // [2G/3, G] that needs the fewest tile rounds over all S samples, the short last group included (C3: 64 tiles
        // per sample; G = 45 -> 22 x 20 + 15 = 455 rounds, G = 37 -> 27 x 16 + 11 = 443 = ceil(65536 / 148)); a launch
        // is charged a quarter of a round for its prologue
        const int64_t sms = tc_sm_count(), t1 = ((p.nu + 127) / 128) * ((nv + 127) / 128);
        double best = 1e300;
        int64_t bg = G;
        for (int64_t g = G; g >= 1 && 3 * g >= 2 * G; --g) {
            const int64_t full = S / g, rem = S % g;
            const double cost = (double)(full * ((t1 * g + sms - 1) / sms) + (t1 * rem + sms - 1) / sms) +
                                0.25 * (double)(full + (rem ? 1 : 0));
            if (cost < best - 1e-9) {
                best = cost;
                bg = g;
            }
        }
        G = bg;
    }
    p.G = G;
    p.gx = (p.nu + GEMM_BM - 1) / GEMM_BM;
    p.gy = (nv + GEMM_BN - 1) / GEMM_BN;
    if (tc) p.gy = (nv + 127) / 128;
    p.nlist = p.gx * p.gy * (tc ? TC_EPI_WARPS : 8);
    int64_t o = 0;
    auto take = [&](int64_t bytes) {
        int64_t r = o;
        o += align_up(bytes, 256);
        return r;
    };
    p.off_feat = take(S * F * p.rec * 8);
    // tf32: hi and lo float arrays
    // mode 8 double-buffers the operand arrays: the build of group g + 1 runs on a second stream under the GEMM of group g
    const int nbuf = tc == 3 ? 2 : 1;
    p.off_p = take(no_p ? 256 : nbuf * (tf32 ? G * p.nup * p.K2p * 8 : G * p.nu * 2 * F * 8));
    p.off_q = take(nbuf * (tf32 ? G * p.K2p * p.nvp * 8 : G * 2 * F * nv * 8));
    p.off_pv = take(p.nlist * G * p.kp * 8);
    p.off_pi = take(p.nlist * G * p.kp * 8);
    p.off_mv = take(S * p.kp * 8);
    p.off_mi = take(S * p.kp * 8);
    p.off_ev = take(S * p.kp * 8);
    p.off_xmax = take(64 * 8);
    p.off_tmax = take(S * 8);
    p.off_featp = take(gen_ncol ? S * (p.K2p / 2) * pg_rs(gen_ncol) * 8 : 0);   // [w_U, b, theta~] records of all S samples
    p.total = o;
    return p;
}

// final top-k of the kp survivors by (exact value desc, index asc); one warp per sample
__global__ void prod_final_kernel(const double* __restrict__ vals, const int64_t* __restrict__ idx, int kp, int k,
                                  double* __restrict__ out_val, int64_t* __restrict__ out_idx) {
    const int64_t j = blockIdx.x;
    const int lane = threadIdx.x;
    double pv = pos_inf();
    int64_t pi = -1;
    for (int r = 0; r < k; ++r) {
        double bv = neg_inf();
        int64_t bi = INT64_MAX;
        for (int e = lane; e < kp; e += 32) {
            const double v = vals[j * kp + e];
            const int64_t i = idx[j * kp + e];
            if (i >= 0 && better(pv, pi, v, i) && better(v, i, bv, bi)) {
                bv = v;
                bi = i;
            }
        }
        warp_argmax(bv, bi);
        const bool found = (bi != INT64_MAX);
        if (lane == 0) {
            out_val[j * k + r] = found ? bv : neg_inf();
            out_idx[j * k + r] = found ? bi : -1;
        }
        if (found) {
            pv = bv;
            pi = bi;
        }
    }
}

}  // namespace tes

using namespace tes;

extern "C" int64_t tes_rff_topk_product_workspace_bytes(int64_t N, int64_t d, int64_t nv, int64_t S, int64_t F, int k,
                                                        int tf32) {
    if (N < 1 || d < 2 || d > 64 || nv < 1 || N % nv != 0 || S < 1 || F < 1 || k < 1 || k > 8) return -1;
    // modes 4 / 5 (P generated on the SM) need the split, which this query does not carry: the largest layout
    const int64_t a = prod_plan(N, d, nv, S, F, k, tf32 != 0, tf32 >= 3).total;   // tf32 / fp16 operand arrays: (hi, lo) x 4 resp. 2 bytes
    if (tf32 == 6) return prod_plan(N, d, nv, S, F, k, 1, 2).total;
    if (tf32 == 8) return prod_plan(N, d, nv, S, F, k, 1, 3).total;
    if (tf32 < 4) return a;
    const int64_t b = prod_plan(N, d, nv, S, F, k, 1, 1, PG_MAXCOL, 8).total;
    const int64_t c = prod_plan(N, d, nv, S, F, k, 1, 1, PG_MAXCOL, 4).total;
    return a > b ? (a > c ? a : c) : (b > c ? b : c);
}

extern "C" int tes_rff_topk_product_f64(const double* cand, int64_t N, int64_t d, int64_t s_lo, int64_t s_hi, int64_t nv,
                                        const double* W, const double* b, const double* theta, int64_t S, int64_t F,
                                        int w_shared, double sigma, int k, int64_t idx_offset, int64_t* out_idx,
                                        double* out_val, int* out_flag, int tf32, void* ws, int64_t ws_bytes,
                                        void* stream) {
    if (N < 1 || !cand) return -1;
    if (d < 2 || d > 64) return -3;
    if (s_lo < 0 || s_hi > d || s_lo >= s_hi || s_hi - s_lo >= d) return -4;
    if (nv < 1 || N % nv != 0) return -5;
    if (!W) return -6;
    if (!b) return -7;
    if (!theta) return -8;
    if (S < 1 || S > (1 << 24)) return -9;
    if (F < 1 || F > (1 << 22)) return -10;
    if (!(sigma >= 0.0)) return -12;
    if (k < 1 || k > 8) return -13;
    if (!out_idx) return -15;
    if (!out_val) return -16;
    if (!out_flag) return -17;
    if (tf32 < 0 || tf32 > 8 || tf32 == 7) return -18;
    if (tf32 >= 2 && d > 16) tf32 = 1;   // the fp16 operand build keeps a feature's weights in 16 registers
    if ((tf32 == 4 || tf32 == 5) && s_hi - s_lo > PG_MAXCOL) tf32 = 3;   // modes 4 / 5 keep the slow coordinates and the record ring in smem
    const int gen_ncol = (tf32 == 4 || tf32 == 5) ? (int)(s_hi - s_lo) : 0;
    const int gen_gc = tf32 == 5 ? 4 : 8;
    ProdPlan pl = prod_plan(N, d, nv, S, F, k, tf32 != 0, tf32 == 8 ? 3 : (tf32 == 6 ? 2 : (tf32 >= 3)), gen_ncol, gen_gc);
    if (!ws || ws_bytes < pl.total) return -100;
    cudaStream_t st = (cudaStream_t)stream;
    char* base = (char*)ws;
    double* feat = (double*)(base + pl.off_feat);
    double* P = (double*)(base + pl.off_p);
    double* Qt = (double*)(base + pl.off_q);
    double* pv = (double*)(base + pl.off_pv);
    int64_t* pi = (int64_t*)(base + pl.off_pi);
    double* mv = (double*)(base + pl.off_mv);
    int64_t* mi = (int64_t*)(base + pl.off_mi);
    double* ev = (double*)(base + pl.off_ev);
    double* xmax = (double*)(base + pl.off_xmax);
    double* tmax = (double*)(base + pl.off_tmax);
    double* featP = (double*)(base + pl.off_featp);
    const double scale = sqrt(2.0 * sigma / (double)F);
    const int64_t nu = pl.nu, K2 = 2 * F;
    prod_pack_kernel<<<(unsigned)((S * F + 255) / 256), 256, 0, st>>>(W, b, theta, S, F, (int)d, w_shared, pl.rec, feat);
    TES_LAUNCH_OK();
    prod_xmax_kernel<<<1, 256, 0, st>>>(cand, nu, nv, (int)d, (int)s_lo, (int)s_hi, xmax);
    TES_LAUNCH_OK();
    TES_CUDA_OK(cudaFuncSetAttribute(prod_gemm_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)sizeof(GemmSmem)));
    TES_CUDA_OK(cudaFuncSetAttribute(prod_gemm_topk_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    float* Phi = (float*)P;
    float* Plo = Phi + pl.G * pl.nup * pl.K2p;
    float* Qhi = (float*)Qt;
    float* Qlo = Qhi + pl.G * pl.K2p * pl.nvp;
    __half* Hhi = (__half*)P;
    __half* Hlo = Hhi + pl.G * pl.nup * pl.K2p;
    __half* Ghi = (__half*)Qt;
    __half* Glo = Ghi + pl.G * pl.nvp * pl.K2p;
    if (tf32) {
        TES_CUDA_OK(cudaFuncSetAttribute(prod_gemm_topk_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(PtSmem)));
        TES_CUDA_OK(cudaFuncSetAttribute(prod_gemm_topk_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(PhSmem)));
        TES_CUDA_OK(cudaFuncSetAttribute(prod_gemm_topk_tc_kernel<0, TC_NS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(TC_NS * TC_STAGE_BYTES + 256)));
        TES_CUDA_OK(cudaFuncSetAttribute(prod_gemm_topk_tc_kernel<0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(2 * TC_STAGE_BYTES + 256)));
        TES_CUDA_OK(cudaFuncSetAttribute(prod_gemm_topk_tc_kernel<1, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(7 * TC_STAGE_BYTES / 2 + 256)));
        TES_CUDA_OK(cudaFuncSetAttribute(prod_gemm_topk_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(T2_NS * T2_STAGE_BYTES + 256)));
        TES_CUDA_OK(cudaFuncSetAttribute(prod_gemm_topk_tc8_kernel<0, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(3 * T8_STAGE_BYTES + 256)));
        // the whole 228 KB as shared memory, so that a build CTA of the second stream (30 KB) fits beside the GEMM CTA
        TES_CUDA_OK(cudaFuncSetAttribute(prod_gemm_topk_tc8_kernel<0, 3>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        TES_CUDA_OK(cudaFuncSetAttribute(prod_gemm_topk_tc8_kernel<1, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(6 * T8_STAGE_BYTES / 2 + 256)));
        TES_CUDA_OK(cudaFuncSetAttribute(prod_gemm_topk_tc8_kernel<1, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(7 * T8_STAGE_BYTES / 2 + 256)));
        // zero the tile padding once (rows >= nu, columns >= nv, k >= 2F are never written by the build kernels)
        const size_t nbuf = tf32 == 8 ? 2 : 1;
        if (!(gen_ncol && gen_gc == 8)) TES_CUDA_OK(cudaMemsetAsync(P, 0, nbuf * (size_t)(pl.G * pl.nup * pl.K2p * 8), st));
        TES_CUDA_OK(cudaMemsetAsync(Qt, 0, nbuf * (size_t)(pl.G * pl.K2p * pl.nvp * 8), st));
        prod_tmax_kernel<<<(unsigned)S, 256, 0, st>>>(theta, F, tmax);
        TES_LAUNCH_OK();
        if (gen_ncol) {
            const int64_t Fp = pl.K2p / 2;
            prod_featp_kernel<<<(unsigned)((S * Fp + 255) / 256), 256, 0, st>>>(W, b, theta, tmax, S, F, Fp, (int)d, (int)s_lo,
                                                                                 gen_ncol, w_shared, featP);
            TES_LAUNCH_OK();
        }
    }
    // second stream + events of the overlapped operand build (mode 8, more than one group); released when the call
    // returns (destruction is deferred by the runtime until the enqueued work has finished)
    struct Overlap {
        bool on = false;
        cudaStream_t s2 = nullptr;
        cudaEvent_t start = nullptr, built[2] = {nullptr, nullptr}, freed[2] = {nullptr, nullptr};
        ~Overlap() {
            for (cudaEvent_t e : {start, built[0], built[1], freed[0], freed[1]})
                if (e) cudaEventDestroy(e);
            if (s2) cudaStreamDestroy(s2);
        }
    } ov;
    {
        static const bool want = [] {
            // developer knob, default OFF: measured at C3 the overlap buys 0.5 ms of 18.6 (the build CTAs mostly wait
            // for the persistent GEMM CTAs to leave), not worth a stream created inside the library by default
            const char* e = getenv("TES_PROD_OVERLAP");
            return e && e[0] == '1';
        }();
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(st, &cap);                  // (a capturing stream cannot fork to a stream created here)
        if (tf32 == 8 && S > pl.G && want && cap == cudaStreamCaptureStatusNone) {
            int lo = 0, hi = 0;
            TES_CUDA_OK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
            TES_CUDA_OK(cudaStreamCreateWithPriority(&ov.s2, cudaStreamNonBlocking, lo));
            for (cudaEvent_t* e : {&ov.start, &ov.built[0], &ov.built[1], &ov.freed[0], &ov.freed[1]})
                TES_CUDA_OK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
            ov.on = true;
        }
    }
    for (int64_t j0 = 0; j0 < S; j0 += pl.G) {
        const int64_t G = (S - j0 < pl.G) ? (S - j0) : pl.G;
        dim3 grid((unsigned)pl.gx, (unsigned)pl.gy, (unsigned)G);
        if (tf32 == 8) {
            // operands of group g + 1 are built on a second, low-priority stream into the other half of the double
            // buffer while the GEMM of group g runs: the build is bound by the conversion / transcendental pipe, the
            // GEMM by operand delivery, and a 128-thread build CTA fits beside the persistent GEMM CTA of an SM
            // (registers 55 296 + 8 192, shared memory 197 + 30 KB)
            const int64_t ntu = pl.gx, ntv = pl.gy, nst = pl.K2p / TC_BK, ntu2 = (ntu + 1) / 2, ntv2 = (ntv + 1) / 2;
            const int64_t gi = j0 / pl.G;
            const size_t pbuf = (size_t)(pl.G * pl.nup * pl.K2p * 4), qbuf = (size_t)(pl.G * pl.K2p * pl.nvp * 4);   // halves
            auto build = [&](int64_t jj0, int64_t GG, int buf, cudaStream_t bs, int threads) -> int {
                __half* ph_ = Hhi + buf * pbuf;
                __half* qh_ = Ghi + buf * qbuf;
                launch_prod_build_tc(dim3((unsigned)ntu, (unsigned)GG, 8), threads, bs, cand, nu, nv, (int)d, (int)s_lo, (int)s_hi, 0, W, b, theta, jj0, F, w_shared, 2 * ntu2, nst, tmax, ph_, 0, 128);
                TES_LAUNCH_OK();
                launch_prod_build_tc(dim3((unsigned)ntv, (unsigned)GG, 8), threads, bs, cand, nv, 1, (int)d, (int)s_lo, (int)s_hi, 1, W, b, theta, jj0, F, w_shared, 2 * ntv2, nst, tmax, qh_, 0, 128);
                TES_LAUNCH_OK();
                return 0;
            };
            const int buf = (int)(gi & 1);
            if (!ov.on) {
                if (int rc = build(j0, G, buf, st, 256)) return rc;
            } else {
                if (gi == 0) {
                    TES_CUDA_OK(cudaEventRecord(ov.start, st));          // tmax, memsets: everything the builds read
                    TES_CUDA_OK(cudaStreamWaitEvent(ov.s2, ov.start, 0));
                    if (int rc = build(j0, G, 0, ov.s2, 128)) return rc;
                    TES_CUDA_OK(cudaEventRecord(ov.built[0], ov.s2));
                }
                const int64_t jn = j0 + pl.G;
                if (jn < S) {                                            // the next group, into the other buffer
                    const int nb = buf ^ 1;
                    if (gi >= 1) TES_CUDA_OK(cudaStreamWaitEvent(ov.s2, ov.freed[nb], 0));   // GEMM gi - 1 read it
                    if (int rc = build(jn, (S - jn < pl.G) ? (S - jn) : pl.G, nb, ov.s2, 128)) return rc;
                    TES_CUDA_OK(cudaEventRecord(ov.built[nb], ov.s2));
                }
                TES_CUDA_OK(cudaStreamWaitEvent(st, ov.built[buf], 0));
            }
            const int64_t ntot = ntu2 * ntv2 * G;
            int64_t ncta = 2 * ntot < tc_sm_count() ? 2 * ntot : (tc_sm_count() / 2) * 2;
            const __half* Pg = Hhi + buf * pbuf;
            const __half* Qg = Ghi + buf * qbuf;
            static const int t8_shape = [] {      // developer knob: 3 whole stages of 64 KB or 6 / 7 half stages of 32 KB
                const char* e = getenv("TES_T8_STAGES");
                return e ? atoi(e) : 3;        // measured at C3: 18.7 ms (3 x 64 KB) / 20.7 ms (6 or 7 x 32 KB)
            }();
            if (t8_shape == 3)
                prod_gemm_topk_tc8_kernel<0, 3><<<(unsigned)ncta, TC_THREADS, 3 * T8_STAGE_BYTES + 256, st>>>(
                    Pg, Qg, nu, nv, (int)ntu, (int)ntv, (int)nst, scale, tmax + j0, idx_offset, pl.kp, (int)G, pv, pi, nullptr, nullptr,
                    nullptr, 0);
            else if (t8_shape == 6)
                prod_gemm_topk_tc8_kernel<1, 6><<<(unsigned)ncta, TC_THREADS, 6 * T8_STAGE_BYTES / 2 + 256, st>>>(
                    Pg, Qg, nu, nv, (int)ntu, (int)ntv, (int)nst, scale, tmax + j0, idx_offset, pl.kp, (int)G, pv, pi, nullptr, nullptr,
                    nullptr, 0);
            else
                prod_gemm_topk_tc8_kernel<1, 7><<<(unsigned)ncta, TC_THREADS, 7 * T8_STAGE_BYTES / 2 + 256, st>>>(
                    Pg, Qg, nu, nv, (int)ntu, (int)ntv, (int)nst, scale, tmax + j0, idx_offset, pl.kp, (int)G, pv, pi, nullptr, nullptr,
                    nullptr, 0);
            TES_LAUNCH_OK();
            if (ov.on) TES_CUDA_OK(cudaEventRecord(ov.freed[buf], st));
        } else if (tf32 == 6) {
            const int64_t ntu = pl.gx, ntv = pl.gy, nst = pl.K2p / TC_BK, ntu2 = (ntu + 1) / 2;
            // P: 128-row tiles, 2 ntu2 of them per sample (an odd count is padded with the zero tile of the memset);
            // Q: 64-column tiles, 2 ntv per sample (each CTA of a pair stages one half of a 128-column tile)
            launch_prod_build_tc(dim3((unsigned)ntu, (unsigned)G, 8), 256, st, cand, nu, nv, (int)d, (int)s_lo, (int)s_hi, 0, W, b, theta, j0, F, w_shared, 2 * ntu2, nst, tmax, Hhi, 0, 128);
            TES_LAUNCH_OK();
            launch_prod_build_tc(dim3((unsigned)(2 * ntv), (unsigned)G, 8), 256, st, cand, nv, 1, (int)d, (int)s_lo, (int)s_hi, 1, W, b, theta, j0, F, w_shared, 2 * ntv, nst, tmax, Ghi, 0, 64);
            TES_LAUNCH_OK();
            const int64_t ntot = ntu2 * ntv * G;
            int64_t ncta = 2 * ntot < tc_sm_count() ? 2 * ntot : (tc_sm_count() / 2) * 2;
            prod_gemm_topk_tc2_kernel<<<(unsigned)ncta, TC_THREADS, T2_NS * T2_STAGE_BYTES + 256, st>>>(
                Hhi, Ghi, nu, nv, (int)ntu, (int)ntv, (int)nst, scale, tmax + j0, idx_offset, pl.kp, (int)G, pv, pi);
            TES_LAUNCH_OK();
        } else if (tf32 >= 4) {
            const int64_t ntu = pl.gx, ntv = pl.gy, nst = pl.K2p / TC_BK;
            if (gen_gc < 8) {   // the streamed half of the P blocks
                launch_prod_build_tc(dim3((unsigned)ntu, (unsigned)G, 8), 256, st, cand, nu, nv, (int)d, (int)s_lo, (int)s_hi, 0, W, b, theta, j0, F, w_shared, ntu, nst, tmax, Hhi, gen_gc, 128);
                TES_LAUNCH_OK();
            }
            launch_prod_build_tc(dim3((unsigned)ntv, (unsigned)G, 8), 256, st, cand, nv, 1, (int)d, (int)s_lo, (int)s_hi, 1, W, b, theta, j0, F, w_shared, ntv, nst, tmax, Ghi, 0, 128);
            TES_LAUNCH_OK();
            int rc = -18;
#define PG_CASE(NC)                                                                                                  \
    case NC:                                                                                                          \
        rc = gen_gc == 8                                                                                              \
                 ? launch_prod_gemm_gen<NC, 8>(tc_sm_count(), cand, nu, nv, (int)d, (int)s_lo, featP, pl.K2p / 2, Hhi, Ghi, \
                                               (int)ntu, (int)ntv, (int)nst, scale, tmax, j0, idx_offset, pl.kp, (int)G,  \
                                               pv, pi, st)                                                            \
                 : launch_prod_gemm_gen<NC, 4>(tc_sm_count(), cand, nu, nv, (int)d, (int)s_lo, featP, pl.K2p / 2, Hhi, Ghi, \
                                               (int)ntu, (int)ntv, (int)nst, scale, tmax, j0, idx_offset, pl.kp, (int)G,  \
                                               pv, pi, st);                                                           \
        break;
            switch (gen_ncol) { PG_CASE(1) PG_CASE(2) PG_CASE(3) PG_CASE(4) PG_CASE(5) PG_CASE(6) }
#undef PG_CASE
            if (rc) return rc;
        } else if (tf32 == 3) {
            const int64_t ntu = pl.gx, ntv = pl.gy, nst = pl.K2p / TC_BK;
            launch_prod_build_tc(dim3((unsigned)ntu, (unsigned)G, 8), 256, st, cand, nu, nv, (int)d, (int)s_lo, (int)s_hi, 0, W, b, theta, j0, F, w_shared, ntu, nst, tmax, Hhi, 0, 128);
            TES_LAUNCH_OK();
            launch_prod_build_tc(dim3((unsigned)ntv, (unsigned)G, 8), 256, st, cand, nv, 1, (int)d, (int)s_lo, (int)s_hi, 1, W, b, theta, j0, F, w_shared, ntv, nst, tmax, Ghi, 0, 128);
            TES_LAUNCH_OK();
            const int64_t ntot = ntu * ntv * G;
            dim3 gtc((unsigned)(ntot < tc_sm_count() ? ntot : tc_sm_count()), 1, 1);
            // stage shape (developer knob TES_TC_STAGES, read once): 3 whole stages of 64 KB (default: 22.5 ms at C3) /
            // 7 half stages of 32 KB (224 KB in flight instead of 192 KB, but 24.9 ms) / 2 whole stages
            static const int tc_shape = [] {
                const char* e = getenv("TES_TC_STAGES");
                return e ? atoi(e) : 3;
            }();
            if (tc_shape == 2)
                prod_gemm_topk_tc_kernel<0, 2><<<gtc, TC_THREADS, 2 * TC_STAGE_BYTES + 256, st>>>(
                    Hhi, Ghi, nu, nv, ntu, ntv, nst, scale, tmax + j0, idx_offset, pl.kp, G, pv, pi, nullptr, 0, nullptr, nullptr);
            else if (tc_shape == 3)
                prod_gemm_topk_tc_kernel<0, TC_NS><<<gtc, TC_THREADS, TC_NS * TC_STAGE_BYTES + 256, st>>>(
                    Hhi, Ghi, nu, nv, ntu, ntv, nst, scale, tmax + j0, idx_offset, pl.kp, G, pv, pi, nullptr, 0, nullptr, nullptr);
            else
                prod_gemm_topk_tc_kernel<1, 7><<<gtc, TC_THREADS, 7 * TC_STAGE_BYTES / 2 + 256, st>>>(
                    Hhi, Ghi, nu, nv, ntu, ntv, nst, scale, tmax + j0, idx_offset, pl.kp, G, pv, pi, nullptr, 0, nullptr, nullptr);
            TES_LAUNCH_OK();
        } else if (tf32 == 2) {
            const int64_t totp = G * ((nu + PB_ROWS - 1) / PB_ROWS) * F, totq = G * ((nv + PB_ROWS - 1) / PB_ROWS) * F;
            prod_build_f16_kernel<<<(unsigned)((totp + 255) / 256), 256, 0, st>>>(
                cand, nu, nv, (int)d, (int)s_lo, (int)s_hi, 0, W, b, theta, j0, G, F, w_shared, pl.nup, pl.K2p, tmax, Hhi, Hlo);
            TES_LAUNCH_OK();
            prod_build_f16_kernel<<<(unsigned)((totq + 255) / 256), 256, 0, st>>>(
                cand, nv, 1, (int)d, (int)s_lo, (int)s_hi, 1, W, b, theta, j0, G, F, w_shared, pl.nvp, pl.K2p, tmax, Ghi, Glo);
            TES_LAUNCH_OK();
            prod_gemm_topk_f16_kernel<<<grid, PT_THREADS, sizeof(PhSmem), st>>>(
                Hhi, Hlo, Ghi, Glo, nu, nv, pl.nup, pl.nvp, pl.K2p, scale, tmax + j0, idx_offset, pl.kp, G, pv, pi);
            TES_LAUNCH_OK();
        } else if (tf32) {
            const int64_t totp = G * nu * F, totq = G * F * nv;
            prod_build_p32_kernel<<<(unsigned)((totp + 255) / 256), 256, 0, st>>>(
                cand, nu, nv, (int)d, (int)s_lo, (int)s_hi, W, b, theta, j0, G, F, w_shared, pl.nup, pl.K2p, Phi, Plo);
            TES_LAUNCH_OK();
            prod_build_q32_kernel<<<(unsigned)((totq + 255) / 256), 256, 0, st>>>(
                cand, nv, (int)d, (int)s_lo, (int)s_hi, W, j0, G, F, w_shared, pl.nvp, pl.K2p, Qhi, Qlo);
            TES_LAUNCH_OK();
            prod_gemm_topk_tf32_kernel<<<grid, PT_THREADS, sizeof(PtSmem), st>>>(
                Phi, Plo, Qhi, Qlo, nu, nv, pl.nup, pl.nvp, pl.K2p, scale, idx_offset, pl.kp, G, pv, pi);
            TES_LAUNCH_OK();
        } else {
            {
                const int64_t tot = G * nu * F;
                prod_build_p_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(cand, nu, nv, (int)d, (int)s_lo, (int)s_hi, W,
                                                                                 b, theta, j0, G, F, w_shared, P);
                TES_LAUNCH_OK();
            }
            {
                const int64_t tot = G * F * nv;
                prod_build_q_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(cand, nv, (int)d, (int)s_lo, (int)s_hi, W, j0,
                                                                                 G, F, w_shared, Qt);
                TES_LAUNCH_OK();
            }
            prod_gemm_topk_kernel<<<grid, GEMM_THREADS, sizeof(GemmSmem), st>>>(P, Qt, nu, nv, K2, scale, idx_offset, pl.kp,
                                                                                G, pv, pi);
            TES_LAUNCH_OK();
        }
        if (int rc = launch_prod_merge(pv, pi, pl.nlist, G, pl.kp, mv + j0 * pl.kp, mi + j0 * pl.kp, st)) return rc;
    }
    {
        const int64_t tot = S * pl.kp;
        prod_exact_kernel<<<(unsigned)((tot + 3) / 4), 128, 0, st>>>(cand, (int)d, feat, pl.rec, S, F, scale,
                                                                       idx_offset, k, pl.kp, mv, mi, xmax, tf32, tmax, ev, out_flag);
        TES_LAUNCH_OK();
    }
    prod_final_kernel<<<(unsigned)S, 32, 0, st>>>(ev, mi, pl.kp, k, out_val, out_idx);
    TES_LAUNCH_OK();
    return 0;
}

// ---- resolve pass for the samples tes_rff_topk_product_f64 flags -----------------------------------------------------
// A sample is flagged when its (k + 8)-th best GEMM value is within 2B of the k-th best: more than 8 near-ties (meshes
// that are fine along a direction the sample is flat in), so the k + 8 survivors need not contain the true top-k.
// Recomputing such a sample with the general kernel costs ~0.4 ms (N = 10^6, F = 1024).  Here instead: the same CTA-pair
// tcgen05 GEMM runs on the flagged samples only, first with the top-kp epilogue (k-th best value vk, bound B), then with
// the THRESHOLD epilogue: every candidate whose GEMM value u reaches vk - 2B is emitted (|u - v| <= B for the exact
// value v, and the k candidates with u >= vk have v >= vk - B, so every member of the true top-k has u >= vk - 2B).
// The emitted candidates are evaluated with the exact spec sequence and ranked by (value desc, index asc): the bits of
// tes_rff_topk_f64.  Still flagged afterwards: more than PR_CAP candidates above the threshold, NaN.
constexpr int PR_CAP = 2048;
extern "C" int64_t tes_rff_topk_product_resolve_workspace_bytes(int64_t N, int64_t d, int64_t nv, int64_t S, int64_t F,
                                                                int k) {
    if (N < 1 || d < 2 || d > 16 || nv < 1 || N % nv != 0 || S < 1 || F < 1 || k < 1 || k > 8) return -1;
    return prod_plan(N, d, nv, S, F, k, 1, 3).total + align_up(S * 8, 256) + align_up(S * 4, 256) +
           2 * align_up(S * (int64_t)PR_CAP * 8, 256);
}
extern "C" int tes_rff_topk_product_resolve_f64(const double* cand, int64_t N, int64_t d, int64_t s_lo, int64_t s_hi,
                                                int64_t nv, const double* W, const double* b, const double* theta,
                                                int64_t S, int64_t F, int w_shared, double sigma, int k,
                                                int64_t idx_offset, int64_t* out_idx, double* out_val, int* out_flag,
                                                void* ws, int64_t ws_bytes, void* stream) {
    if (N < 1 || !cand) return -1;
    if (d < 2 || d > 16) return -3;
    if (s_lo < 0 || s_hi > d || s_lo >= s_hi || s_hi - s_lo >= d) return -4;
    if (nv < 1 || N % nv != 0) return -5;
    if (!W) return -6;
    if (!b) return -7;
    if (!theta) return -8;
    if (S < 1 || S > (1 << 24)) return -9;
    if (F < 1 || F > (1 << 22)) return -10;
    if (!(sigma >= 0.0)) return -12;
    if (k < 1 || k > 8) return -13;
    if (!out_idx) return -15;
    if (!out_val) return -16;
    if (!out_flag) return -17;
    ProdPlan pl = prod_plan(N, d, nv, S, F, k, 1, 3);
    if (!ws || ws_bytes < tes_rff_topk_product_resolve_workspace_bytes(N, d, nv, S, F, k)) return -100;
    cudaStream_t st = (cudaStream_t)stream;
    char* base = (char*)ws;
    double* feat = (double*)(base + pl.off_feat);
    __half* Hhi = (__half*)(base + pl.off_p);
    __half* Ghi = (__half*)(base + pl.off_q);
    double* pv = (double*)(base + pl.off_pv);
    int64_t* pi = (int64_t*)(base + pl.off_pi);
    double* mv = (double*)(base + pl.off_mv);
    int64_t* mi = (int64_t*)(base + pl.off_mi);
    double* ev = (double*)(base + pl.off_ev);
    double* xmax = (double*)(base + pl.off_xmax);
    double* tmax = (double*)(base + pl.off_tmax);
    char* extra = base + pl.total;
    double* thr = (double*)extra;
    extra += align_up(S * 8, 256);
    int* cnt = (int*)extra;
    extra += align_up(S * 4, 256);
    int64_t* lidx = (int64_t*)extra;
    extra += align_up(S * (int64_t)PR_CAP * 8, 256);
    double* lval = (double*)extra;
    const double scale = sqrt(2.0 * sigma / (double)F);
    const int64_t nu = pl.nu;
    prod_pack_kernel<<<(unsigned)((S * F + 255) / 256), 256, 0, st>>>(W, b, theta, S, F, (int)d, w_shared, pl.rec, feat);
    TES_LAUNCH_OK();
    prod_xmax_kernel<<<1, 256, 0, st>>>(cand, nu, nv, (int)d, (int)s_lo, (int)s_hi, xmax);
    TES_LAUNCH_OK();
    TES_CUDA_OK(cudaFuncSetAttribute(prod_gemm_topk_tc8_kernel<0, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)(3 * T8_STAGE_BYTES + 256)));
    TES_CUDA_OK(cudaMemsetAsync(Hhi, 0, (size_t)(pl.G * pl.nup * pl.K2p * 8), st));
    TES_CUDA_OK(cudaMemsetAsync(Ghi, 0, (size_t)(pl.G * pl.K2p * pl.nvp * 8), st));
    TES_CUDA_OK(cudaMemsetAsync(cnt, 0, (size_t)S * 4, st));
    prod_tmax_kernel<<<(unsigned)S, 256, 0, st>>>(theta, F, tmax);
    TES_LAUNCH_OK();
    const int64_t ntu = pl.gx, ntv = pl.gy, nst = pl.K2p / TC_BK, ntu2 = (ntu + 1) / 2, ntv2 = (ntv + 1) / 2;
    for (int64_t j0 = 0; j0 < S; j0 += pl.G) {
        const int64_t G = (S - j0 < pl.G) ? (S - j0) : pl.G;
        launch_prod_build_tc(dim3((unsigned)ntu, (unsigned)G, 8), 256, st, cand, nu, nv, (int)d, (int)s_lo, (int)s_hi, 0, W, b, theta, j0, F, w_shared, 2 * ntu2, nst, tmax, Hhi, 0, 128);
        TES_LAUNCH_OK();
        launch_prod_build_tc(dim3((unsigned)ntv, (unsigned)G, 8), 256, st, cand, nv, 1, (int)d, (int)s_lo, (int)s_hi, 1, W, b, theta, j0, F, w_shared, 2 * ntv2, nst, tmax, Ghi, 0, 128);
        TES_LAUNCH_OK();
        const int64_t ntot = ntu2 * ntv2 * G;
        const int64_t ncta = 2 * ntot < tc_sm_count() ? 2 * ntot : (tc_sm_count() / 2) * 2;
        prod_gemm_topk_tc8_kernel<0, 3><<<(unsigned)ncta, TC_THREADS, 3 * T8_STAGE_BYTES + 256, st>>>(
            Hhi, Ghi, nu, nv, (int)ntu, (int)ntv, (int)nst, scale, tmax + j0, idx_offset, pl.kp, (int)G, pv, pi, nullptr, nullptr,
            nullptr, 0);
        TES_LAUNCH_OK();
        if (int rc = launch_prod_merge(pv, pi, pl.nlist, G, pl.kp, mv + j0 * pl.kp, mi + j0 * pl.kp, st)) return rc;
        // vk and B of the group's samples -> thresholds (the exact values of the kp survivors are not needed here)
        prod_exact_kernel<<<(unsigned)((G * pl.kp + 3) / 4), 128, 0, st>>>(
            cand, (int)d, feat + j0 * F * pl.rec, pl.rec, G, F, scale, idx_offset, k, pl.kp, mv + j0 * pl.kp, mi + j0 * pl.kp, xmax,
            8, tmax + j0, ev + j0 * pl.kp, out_flag + j0, thr + j0);
        TES_LAUNCH_OK();
        prod_gemm_topk_tc8_kernel<0, 3><<<(unsigned)ncta, TC_THREADS, 3 * T8_STAGE_BYTES + 256, st>>>(
            Hhi, Ghi, nu, nv, (int)ntu, (int)ntv, (int)nst, scale, tmax + j0, idx_offset, pl.kp, (int)G, pv, pi, thr + j0,
            lidx + j0 * PR_CAP, cnt + j0, PR_CAP);
        TES_LAUNCH_OK();
    }
    prod_exact_list_kernel<<<(unsigned)((S * PR_CAP + 3) / 4), 128, 0, st>>>(cand, (int)d, feat, pl.rec, S, F, scale,
                                                                              idx_offset, PR_CAP, lidx, cnt, lval);
    TES_LAUNCH_OK();
    prod_final_list_kernel<<<(unsigned)S, 256, 0, st>>>(lval, lidx, cnt, thr, PR_CAP, k, out_val, out_idx, out_flag);
    TES_LAUNCH_OK();
    return 0;
}

// ---- product-structure detection (ops.detect_product_structure) ----------------------------------------------------
// Two passes over the candidate array instead of ~40 framework launches and ~20 host syncs:
//   runlen:  r[k] = first row whose coordinate k differs from row 0 (N if none)  -> the host derives nv of a split
//   check:   for a given nv, bad[k] = 1 if column k is not "slow" (constant inside every block of nv rows),
//            bad[d + k] = 1 if it is not "fast" (row i equals row i % nv)
// Comparisons are IEEE == (NaN never equal, -0 == +0), as the tensor comparisons they replace.
namespace tes {
__global__ void product_runlen_init_kernel(int64_t N, int d, unsigned long long* __restrict__ r) {
    if ((int)threadIdx.x < d) r[threadIdx.x] = (unsigned long long)N;
}
__global__ void __launch_bounds__(256)
    product_runlen_kernel(const double* __restrict__ cand, int64_t N, int d, unsigned long long* __restrict__ r) {
    const int64_t tot = N * d;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(e % d);
        const int64_t i = e / d;
        if (!(cand[e] == cand[k]) && (unsigned long long)i < r[k]) atomicMin(r + k, (unsigned long long)i);
    }
}
__global__ void __launch_bounds__(256)
    product_check_kernel(const double* __restrict__ cand, int64_t N, int d, int64_t nv, int* __restrict__ bad) {
    const int64_t tot = N * d;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(e % d);
        const int64_t i = e / d;
        const double v = cand[e];
        if (!(v == cand[(i / nv) * nv * d + k]) && !bad[k]) bad[k] = 1;
        if (!(v == cand[(i % nv) * d + k]) && !bad[d + k]) bad[d + k] = 1;
    }
}
// first[k] = first row whose column k is not constant inside its block of nv rows ("slow" broken), first[d + k] = first
// row that differs from row i % nv in column k ("fast" broken); N where the column never breaks
__global__ void __launch_bounds__(256)
    product_prefix_kernel(const double* __restrict__ cand, int64_t N, int d, int64_t nv, unsigned long long* __restrict__ first) {
    const int64_t tot = N * d;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(e % d);
        const int64_t i = e / d;
        const double v = cand[e];
        if (!(v == cand[(i / nv) * nv * d + k]) && (unsigned long long)i < first[k]) atomicMin(&first[k], (unsigned long long)i);
        if (!(v == cand[(i % nv) * d + k]) && (unsigned long long)i < first[d + k])
            atomicMin(&first[d + k], (unsigned long long)i);
    }
}
}  // namespace tes

extern "C" int tes_product_prefix_f64(const double* cand, int64_t N, int64_t d, int64_t nv, int64_t* out_first,
                                      void* stream) {
    if (N < 1 || !cand) return -1;
    if (d < 1 || d > 32) return -3;
    if (nv < 1 || nv > N) return -4;
    if (!out_first) return -5;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long* r = reinterpret_cast<unsigned long long*>(out_first);
    product_runlen_init_kernel<<<1, 64, 0, st>>>(N, (int)(2 * d), r);
    TES_LAUNCH_OK();
    const int64_t tot = N * d;
    const int64_t want = (tot + 255) / 256, cap = (int64_t)tc_sm_count() * 8;
    product_prefix_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(cand, N, (int)d, nv, r);
    TES_LAUNCH_OK();
    return 0;
}

extern "C" int tes_product_runlen_f64(const double* cand, int64_t N, int64_t d, int64_t* out_r, void* stream) {
    if (N < 1 || !cand) return -1;
    if (d < 1 || d > 64) return -3;
    if (!out_r) return -4;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long* r = reinterpret_cast<unsigned long long*>(out_r);
    product_runlen_init_kernel<<<1, 64, 0, st>>>(N, (int)d, r);
    TES_LAUNCH_OK();
    const int64_t tot = N * d;
    const int64_t want = (tot + 255) / 256, cap = (int64_t)tc_sm_count() * 8;
    product_runlen_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(cand, N, (int)d, r);
    TES_LAUNCH_OK();
    return 0;
}

extern "C" int tes_product_check_f64(const double* cand, int64_t N, int64_t d, int64_t nv, int* out_bad, void* stream) {
    if (N < 1 || !cand) return -1;
    if (d < 1 || d > 64) return -3;
    if (nv < 1 || nv > N) return -4;
    if (!out_bad) return -5;
    cudaStream_t st = (cudaStream_t)stream;
    TES_CUDA_OK(cudaMemsetAsync(out_bad, 0, (size_t)(2 * d) * sizeof(int), st));
    const int64_t tot = N * d;
    const int64_t want = (tot + 255) / 256, cap = (int64_t)tc_sm_count() * 8;
    product_check_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(cand, N, (int)d, nv, out_bad);
    TES_LAUNCH_OK();
    return 0;
}
