// gemm.cuh -- fp64 tiled GEMM core for the tall-skinny products of stages B and D.
//
// B200's tcgen05 tensor cores have no fp64 path; fp64 contractions run either on the 64-lane/SM DFMA pipe
// (36.3 TFLOP/s measured) or through the legacy warp-level fp64 MMA (mma.sync m8n8k4, SASS DMMA.8x8x4;
// 37.2 TFLOP/s measured with tes_fp64_mma_burn).  The peak is the same, but one DMMA retires 256 FMAs for
// a single issue slot and two 8-byte operand registers per thread, so the GEMM core uses DMMA:
//   CTA tile 128 x 64, k-slices of 16, 256 threads = 8 warps arranged 4 (rows) x 2 (cols), each warp owning a
//   32 x 32 block = 4 x 4 DMMA tiles (32 accumulator doubles per thread); per k-step of 4 a warp loads
//   4 A and 4 B fragments (conflict-free 64-bit LDS thanks to the 132 / 68 strides) and issues 16 DMMAs.
// Global->shared goes through registers (prefetch of slice t+1 while slice t is multiplied); all loads are
// bounds-checked with zero fill so any (M, N, K) and leading dimension works.
//
// The A operand comes from a provider functor, which lets the TES_ep quadratic forms
// (evaluate_mp.py:103-107) generate their operand -- products M[x,a]*M[x,b] -- on the fly.
#pragma once
#include "tes_common.cuh"

namespace tes {

constexpr int GEMM_BM = 128, GEMM_BN = 64, GEMM_BK = 16, GEMM_THREADS = 256;
// strides = 4 (mod 16) doubles: the 4 k-rows of a fragment land 8 banks apart -> conflict-free 64-bit LDS
constexpr int GEMM_LDA_S = GEMM_BM + 4;
constexpr int GEMM_LDB_S = GEMM_BN + 4;

struct GemmSmem {
    double a[2][GEMM_BK][GEMM_LDA_S];
    double b[2][GEMM_BK][GEMM_LDB_S];
};

// accumulator element acc[i][j][e] of a thread: DMMA tile (i, j) of its warp's 32 x 32 block
__device__ __forceinline__ int gemm_row_of(int warp, int lane, int i) { return (warp >> 1) * 32 + i * 8 + (lane >> 2); }
__device__ __forceinline__ int gemm_col_of(int warp, int lane, int j, int e) {
    return (warp & 1) * 32 + j * 8 + (lane & 3) * 2 + e;
}

__device__ __forceinline__ void dmma_8x8x4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// k-steps [K0, K1) (multiples of 4) of one BK-slice
template <int K0, int K1>
__device__ __forceinline__ void gemm_mac_steps(const double (*as)[GEMM_LDA_S], const double (*bs)[GEMM_LDB_S],
                                               int warp, int lane, double acc[4][4][2]) {
    const int rbase = (warp >> 1) * 32 + (lane >> 2);
    const int cbase = (warp & 1) * 32 + (lane >> 2);
    const int kq = lane & 3;
#pragma unroll
    for (int k0 = K0; k0 < K1; k0 += 4) {
        double av[4], bv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) av[i] = as[k0 + kq][rbase + 8 * i];
#pragma unroll
        for (int j = 0; j < 4; ++j) bv[j] = bs[k0 + kq][cbase + 8 * j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) dmma_8x8x4(acc[i][j][0], acc[i][j][1], av[i], bv[j]);
    }
}
__device__ __forceinline__ void gemm_mac_slice(const double (*as)[GEMM_LDA_S], const double (*bs)[GEMM_LDB_S],
                                               int warp, int lane, double acc[4][4][2]) {
    gemm_mac_steps<0, GEMM_BK>(as, bs, warp, lane, acc);
}

// Plain row-major operand loaders -------------------------------------------------------------
// A slice: BM x BK, thread t loads 8 elements: row = t/2 ... choose mapping for coalescing along K:
// each thread loads rows r = (t / 16) + 16*q (q=0..7), k = t % 16.
struct APlain {
    const double* A;
    int64_t lda;
    int64_t M, K;
    static constexpr bool kTransposed = false;
    static constexpr int NREG = 8;
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
    static constexpr int NREG = 8;
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
__device__ __forceinline__ void gemm_store_b(double (*bs)[GEMM_LDB_S], int t, const double regs[4]) {
    const int kk = t >> 4;
    const int c = t & 15;
#pragma unroll
    for (int q = 0; q < 4; ++q) bs[kk][c + 16 * q] = regs[q];
}

// Mainloop: acc += A(row0.., :) * B(:, col0..) over K.
template <class AProv>
__device__ __forceinline__ void gemm_mainloop(const AProv& ap, const double* __restrict__ B, int64_t ldb, int64_t K,
                                              int64_t N, int64_t row0, int64_t col0, GemmSmem& sm,
                                              double acc[4][4][2], int64_t kbeg = 0) {
    const int t = threadIdx.x;
    const int warp = t >> 5, lane = t & 31;
    double ra[AProv::NREG], rb[4];
    const int64_t nk = (K + GEMM_BK - 1) / GEMM_BK;
    const int64_t kt0 = kbeg / GEMM_BK;   // slices below kbeg are known to be zero (triangular operands)
    if (kt0 >= nk) return;
    ap.load(row0, kt0 * GEMM_BK, t, ra);
    gemm_load_b(B, ldb, K, N, kt0 * GEMM_BK, col0, t, rb);
    ap.store(sm.a[kt0 & 1], t, ra);
    gemm_store_b(sm.b[kt0 & 1], t, rb);
    __syncthreads();
    for (int64_t kt = kt0; kt < nk; ++kt) {
        const int cur = (int)(kt & 1);
        if (kt + 1 < nk) {
            ap.load(row0, (kt + 1) * GEMM_BK, t, ra);
            gemm_load_b(B, ldb, K, N, (kt + 1) * GEMM_BK, col0, t, rb);
        }
        // the next slice's operands go to shared memory in the MIDDLE of this slice's DMMA stream: the other buffer
        // is free since the last barrier, the loads issued above have had two k-steps to land, and the warp reaches
        // the barrier straight after its last DMMA, so the fp64 pipe keeps draining queued DMMAs across the barrier
        // (with the store after the slice, its DMULs queued behind the DMMAs and the pipe idled at every barrier:
        // ncu showed 17 % of all stall samples on the first DMUL, DMMA pipe 63 % busy)
        gemm_mac_steps<0, GEMM_BK / 2>(sm.a[cur], sm.b[cur], warp, lane, acc);
        if (kt + 1 < nk) {
            ap.store(sm.a[cur ^ 1], t, ra);
            gemm_store_b(sm.b[cur ^ 1], t, rb);
        }
        gemm_mac_steps<GEMM_BK / 2, GEMM_BK>(sm.a[cur], sm.b[cur], warp, lane, acc);
        __syncthreads();
    }
}

}  // namespace tes
