// gemm.cuh -- fp64 tiled GEMM core for the tall-skinny products of stages B and D.
//
// B200 has no fp64 path on the tcgen05 tensor cores worth using (fp64 runs on the 64-lane/SM
// DFMA pipe, 36.3 TFLOP/s measured), so the contraction is a classic register-tiled CUDA-core
// GEMM: 128x64 CTA tile, 16-deep k-slices, 256 threads each owning an 8x4 accumulator block
// laid out with the strided-fragment mapping (rows ty*2+{0,1}+32i, cols tx*2+{0,1}+32j) so that
// every LDS.128 of a warp is conflict-free and A loads broadcast.  Global->shared goes through
// registers (prefetch of slice t+1 while slice t is multiplied); all loads are bounds-checked with
// zero fill so any (M, N, K) and leading dimension works.
//
// The A operand comes from a provider functor, which lets the TES_ep quadratic forms
// (evaluate_mp.py:103-107) generate their operand -- products M[x,a]*M[x,b] -- on the fly.
#pragma once
#include "tes_common.cuh"

namespace tes {

constexpr int GEMM_BM = 128, GEMM_BN = 64, GEMM_BK = 16, GEMM_THREADS = 256;
constexpr int GEMM_LDA_S = GEMM_BM + 2;  // +2 doubles: k-major stores from the global-load mapping stay <= 2-way conflicted

struct GemmSmem {
    double a[2][GEMM_BK][GEMM_LDA_S];
    double b[2][GEMM_BK][GEMM_BN];
};

// acc[i][j]: row = row0 + (i/2)*32 + ty*2 + (i&1),  col = col0 + (j/2)*32 + tx*2 + (j&1)
__device__ __forceinline__ int gemm_row_of(int ty, int i) { return (i >> 1) * 32 + ty * 2 + (i & 1); }
__device__ __forceinline__ int gemm_col_of(int tx, int j) { return (j >> 1) * 32 + tx * 2 + (j & 1); }

__device__ __forceinline__ void gemm_mac_slice(const double (*as)[GEMM_LDA_S], const double (*bs)[GEMM_BN], int ty,
                                               int tx, double acc[8][4]) {
#pragma unroll
    for (int kk = 0; kk < GEMM_BK; ++kk) {
        double av[8], bv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            double2 v = *reinterpret_cast<const double2*>(&as[kk][i * 32 + ty * 2]);
            av[2 * i] = v.x;
            av[2 * i + 1] = v.y;
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            double2 v = *reinterpret_cast<const double2*>(&bs[kk][j * 32 + tx * 2]);
            bv[2 * j] = v.x;
            bv[2 * j + 1] = v.y;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
    }
}

// Plain row-major operand loaders -------------------------------------------------------------
// A slice: BM x BK, thread t loads 8 elements: row = t/2 ... choose mapping for coalescing along K:
// each thread loads rows r = (t / 16) + 16*q (q=0..7), k = t % 16.
struct APlain {
    const double* A;
    int64_t lda;
    int64_t M, K;
    static constexpr bool kTransposed = false;
    __device__ __forceinline__ void load(int64_t row0, int64_t k0, int t, double regs[8]) const {
        const int kk = t & 15;
        const int r = t >> 4;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            int64_t row = row0 + r + 16 * q;
            int64_t k = k0 + kk;
            regs[q] = (row < M && k < K) ? __ldg(A + row * lda + k) : 0.0;
        }
    }
    __device__ __forceinline__ void store(double (*as)[GEMM_LDA_S], int t, const double regs[8]) const {
        const int kk = t & 15;
        const int r = t >> 4;
#pragma unroll
        for (int q = 0; q < 8; ++q) as[kk][r + 16 * q] = regs[q];
    }
};

// Transposed A operand: element (row, k) lives at A[k*ldt + row] (A^T stored row-major, K x M).
// Threads run along rows so the global reads stay coalesced and the smem stores conflict-free.
struct ATrans {
    const double* A;
    int64_t ldt;
    int64_t M, K;
    __device__ __forceinline__ void load(int64_t row0, int64_t k0, int t, double regs[8]) const {
        const int r = t & 127;
        const int kb = t >> 7;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            int64_t row = row0 + r;
            int64_t k = k0 + kb + 2 * q;
            regs[q] = (row < M && k < K) ? __ldg(A + k * ldt + row) : 0.0;
        }
    }
    __device__ __forceinline__ void store(double (*as)[GEMM_LDA_S], int t, const double regs[8]) const {
        const int r = t & 127;
        const int kb = t >> 7;
#pragma unroll
        for (int q = 0; q < 8; ++q) as[kb + 2 * q][r] = regs[q];
    }
};

// B slice: BK x BN, thread t loads 4 elements: k = t/16 (0..15), cols c = (t%16) + 16*q
__device__ __forceinline__ void gemm_load_b(const double* __restrict__ B, int64_t ldb, int64_t K, int64_t N,
                                            int64_t k0, int64_t col0, int t, double regs[4]) {
    const int kk = t >> 4;
    const int c = t & 15;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        int64_t k = k0 + kk, col = col0 + c + 16 * q;
        regs[q] = (k < K && col < N) ? __ldg(B + k * ldb + col) : 0.0;
    }
}
__device__ __forceinline__ void gemm_store_b(double (*bs)[GEMM_BN], int t, const double regs[4]) {
    const int kk = t >> 4;
    const int c = t & 15;
#pragma unroll
    for (int q = 0; q < 4; ++q) bs[kk][c + 16 * q] = regs[q];
}

// Mainloop: acc += A(row0.., :) * B(:, col0..) over K.
template <class AProv>
__device__ __forceinline__ void gemm_mainloop(const AProv& ap, const double* __restrict__ B, int64_t ldb, int64_t K,
                                              int64_t N, int64_t row0, int64_t col0, GemmSmem& sm,
                                              double acc[8][4]) {
    const int t = threadIdx.x;
    const int ty = t >> 4, tx = t & 15;
    double ra[8], rb[4];
    const int64_t nk = (K + GEMM_BK - 1) / GEMM_BK;
    ap.load(row0, 0, t, ra);
    gemm_load_b(B, ldb, K, N, 0, col0, t, rb);
    ap.store(sm.a[0], t, ra);
    gemm_store_b(sm.b[0], t, rb);
    __syncthreads();
    for (int64_t kt = 0; kt < nk; ++kt) {
        const int cur = (int)(kt & 1);
        if (kt + 1 < nk) {
            ap.load(row0, (kt + 1) * GEMM_BK, t, ra);
            gemm_load_b(B, ldb, K, N, (kt + 1) * GEMM_BK, col0, t, rb);
        }
        gemm_mac_slice(sm.a[cur], sm.b[cur], ty, tx, acc);
        if (kt + 1 < nk) {
            ap.store(sm.a[cur ^ 1], t, ra);
            gemm_store_b(sm.b[cur ^ 1], t, rb);
        }
        __syncthreads();
    }
}

}  // namespace tes
