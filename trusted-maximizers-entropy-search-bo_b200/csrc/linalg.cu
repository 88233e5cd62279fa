// linalg.cu -- batched SPD factorisation / inversion on sm_100a.
//
//   tes_batched_potrf_f64     lower Cholesky factor of `batch` SPD matrices (m x m)
//   tes_batched_chol2inv_f64  A^-1 = L^-T L^-1 with L = chol(A): utils.chol2inv (utils.py:657-674),
//                             utils.multichol2inv (:677-700), precomputeInvKs (:1838-1856), the
//                             maximizer-conditioned batch of eval_invKmaxsams (:1859-1883), and the
//                             SPD inverse get_pNK_test_obs needs (evaluate_mp.py:47).
//
// Layout of the algorithm (NB = 64):
//   * m <= 64: ONE CTA PER MATRIX, everything in shared memory (factor, triangular inverse, X^T X);
//   * m  > 64: blocked right-looking Cholesky over the whole batch at once --
//       diagonal block: one CTA per matrix in shared memory (also yields inv(L_kk)),
//       panel:  L[i,k] = A[i,k] inv(L_kk)^T   (batched GEMM),
//       update: A[i,j] -= L[i,k] L[j,k]^T      (batched GEMM, lower tiles only),
//     then the blocked triangular inverse X = L^-1 (two batched GEMMs per block row) and
//     A^-1 = X^T X (batched GEMM with the transposed-A loader).
//   Every matrix gets info[b] = 0, or the 1-based index of the first non-positive pivot (the
//   result then contains NaN, like tf.linalg.cholesky failing -- the drivers' NaN-retry applies).
#include "gemm.cuh"
#include "gp.cuh"

namespace tes {

constexpr int CH_NB = 64;
constexpr int CH_LD = CH_NB + 1;
constexpr size_t CH_SMEM = (CH_NB * CH_LD + 5 * CH_NB) * sizeof(double) + 16;   // 36 KB: 6 CTAs per SM

// One CTA per matrix.  Loads the nb x nb diagonal block at (k0,k0) of A (ld = m), factors it, writes L (lower,
// upper zeroed) to Lout at the same position and inv(L) to Dinv[b][blk] (CH_NB x CH_NB, ld CH_NB, zero padded).
// If ainv != NULL (m <= 64 path) also writes inv(L)^T inv(L) there.
//
// Register-tiled right-looking factorisation: the 256 threads form a 16 x 16 grid and thread (ti, tj) owns the 4 x 4
// cyclic sub-matrix {(ti + 16 a, tj + 16 b)} in registers (the trailing matrix shrinks evenly over the threads).
// Step k: the owners of column k have published its (unscaled) entries c; with r = rsqrt(c_k) every thread applies
// a_ij -= (c_i r)(c_j r) to its registers, the column owners keep L_ik = c_i r, and the owners of column k+1 publish
// their freshly updated entries -- ONE barrier per step, no serial section (the previous version ran sqrt on thread 0
// between three barriers per step and inverted L with one thread per column: 0.44 ms for 1024 64 x 64 matrices).
// The triangular inverse X = L^-1 runs the same way: X starts as I, step k publishes row k scaled by 1/L_kk = r_k and
// every thread applies x_ij -= L_ik x_kj.  Blocks narrower than 64 are padded with the identity.
__global__ void __launch_bounds__(256)
    potrf_diag_kernel(const double* __restrict__ A, int64_t m, int64_t k0, int nb, double* __restrict__ Lout,
                      double* __restrict__ Dinv, int64_t dinv_batch_stride, int* __restrict__ info,
                      double* __restrict__ ainv) {
    extern __shared__ __align__(16) double chsm[];
    double (*a)[CH_LD] = reinterpret_cast<double (*)[CH_LD]>(chsm);
    double* colbuf = reinterpret_cast<double*>(a + CH_NB);   // [2][64]
    double* rowbuf = colbuf + 2 * CH_NB;                      // [2][64]
    double* rdiag = rowbuf + 2 * CH_NB;                       // [64]  1 / L_kk
    int& bad = *reinterpret_cast<int*>(rdiag + CH_NB);
    const int64_t b = blockIdx.x;
    const double* Ab = A + b * m * m;
    const int t = threadIdx.x;
    const int ti = t >> 4, tj = t & 15;
    if (t == 0) bad = 0;
    if (t < CH_NB) rdiag[t] = 1.0;                 // identity padding (k >= nb)
    for (int e = t; e < CH_NB * CH_NB; e += 256) {
        int i = e / CH_NB, j = e % CH_NB;
        a[i][j] = (i < nb && j < nb && j <= i) ? Ab[(k0 + i) * m + k0 + j] : ((i == j && i >= nb) ? 1.0 : 0.0);
    }
    __syncthreads();
    double r[4][4];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) r[p][q] = a[ti + 16 * p][tj + 16 * q];
    if (tj == 0) {
#pragma unroll
        for (int p = 0; p < 4; ++p) colbuf[ti + 16 * p] = r[p][0];
    }
    __syncthreads();
    // (steps k >= nb only touch the identity padding: column k = e_k, r = 1, no update -- skipped)
#pragma unroll 1
    for (int k = 0; k < nb; ++k) {
        const double* cb = colbuf + (k & 1) * CH_NB;
        const double dkk = cb[k];
        const double rs = rsqrt(dkk);
        if (t == 0) {
            if (!(dkk > 0.0) && bad == 0 && k < nb) bad = k + 1;
            rdiag[k] = rs;
        }
        double ci[4], cj[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            ci[p] = cb[ti + 16 * p] * rs;
            cj[p] = cb[tj + 16 * p] * rs;
        }
        const int k1 = k + 1;
        double* nb_ = colbuf + (k1 & 1) * CH_NB;
        // only blocks p >= q can hold lower-triangle entries (ti, tj < 16): the other 6 are pruned at compile time
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int j = tj + 16 * q;
            const bool jgt = j > k;
#pragma unroll
            for (int p = q; p < 4; ++p) {
                if (p > q) {
                    if (jgt) r[p][q] = fma(-ci[p], cj[q], r[p][q]);
                } else {
                    if (jgt && ti >= tj) r[p][q] = fma(-ci[p], cj[q], r[p][q]);
                }
            }
            if (j == k) {   // column k is final: L_ik = c_i r (i > k), L_kk = d r
#pragma unroll
                for (int p = q; p < 4; ++p) {
                    const int i = ti + 16 * p;
                    r[p][q] = i > k ? ci[p] : (i == k ? dkk * rs : r[p][q]);
                }
            }
            if (j == k1) {  // publish column k+1 (entries above the diagonal are never read)
#pragma unroll
                for (int p = q; p < 4; ++p) nb_[ti + 16 * p] = r[p][q];
            }
        }
        __syncthreads();
    }
    // L (lower) to shared memory; X = I in registers
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int i = ti + 16 * p, j = tj + 16 * q;
            a[i][j] = j <= i ? r[p][q] : 0.0;
            r[p][q] = i == j ? 1.0 : 0.0;
        }
    if (ti == 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q) rowbuf[tj + 16 * q] = r[0][q] * rdiag[0];
    }
    __syncthreads();
    {
        double* Lb = Lout + b * m * m;
        for (int e = t; e < nb * nb; e += 256) {
            int i = e / nb, j = e % nb;
            Lb[(k0 + i) * m + k0 + j] = (j <= i) ? a[i][j] : 0.0;
        }
    }
#pragma unroll 1
    for (int k = 0; k < nb; ++k) {
        const double* rb = rowbuf + (k & 1) * CH_NB;
        double xk[4], lk[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            xk[p] = rb[tj + 16 * p];
            lk[p] = a[ti + 16 * p][k];
        }
        const int k1 = k + 1;
        double* nr = rowbuf + (k1 & 1) * CH_NB;
        const double rs1 = rdiag[k1 < CH_NB ? k1 : 0];
        // X is lower-triangular: blocks p < q stay zero and are pruned at compile time
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int i = ti + 16 * p;
            const bool below = i > k;
#pragma unroll
            for (int q = 0; q <= p; ++q)
                if (below) r[p][q] = fma(-lk[p], xk[q], r[p][q]);
            if (i == k) {
#pragma unroll
                for (int q = 0; q <= p; ++q) r[p][q] = xk[q];
            }
            if (i == k1) {  // publish row k+1 scaled by 1 / L_{k+1,k+1}
#pragma unroll
                for (int q = 0; q < 4; ++q) nr[tj + 16 * q] = q <= p ? r[p][q] * rs1 : 0.0;
            }
        }
        __syncthreads();
    }
    // X overwrites L in shared memory (the last step's barrier has passed: nobody reads L any more)
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int i = ti + 16 * p, j = tj + 16 * q;
            a[i][j] = (j <= i && i < nb) ? r[p][q] : 0.0;
        }
    __syncthreads();
    if (Dinv) {
        double* Db = Dinv + b * dinv_batch_stride;
        for (int e = t; e < CH_NB * CH_NB; e += 256) Db[e] = a[e / CH_NB][e % CH_NB];
    }
    if (ainv) {
        double* Ib = ainv + b * m * m;
        for (int e = t; e < nb * nb; e += 256) {
            int i = e / nb, j = e % nb;
            int lo = i > j ? i : j;
            double s = 0.0;
            for (int k = lo; k < nb; ++k) s = fma(a[k][i], a[k][j], s);
            Ib[i * m + j] = s;
        }
    }
    if (t == 0 && info && bad != 0) {
        if (info[b] == 0) info[b] = (int)k0 + bad;
    }
}

__global__ void zero_upper_kernel(double* __restrict__ L, int64_t m, int64_t batch) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= batch * m * m) return;
    int64_t r = (e / m) % m, c = e % m;
    if (c > r) L[e] = 0.0;
}

// X[b][i0+i][i0+j] = Dinv[b][blk][i][j] for every diagonal block
__global__ void copy_diag_inv_kernel(const double* __restrict__ Dinv, int64_t dstride, double* __restrict__ X,
                                     int64_t m, int64_t batch) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= batch * m * CH_NB) return;
    int64_t b = e / (m * CH_NB);
    int64_t r = (e / CH_NB) % m;
    int j = (int)(e % CH_NB);
    int64_t blk = r / CH_NB;
    int i = (int)(r % CH_NB);
    int64_t c = blk * CH_NB + j;
    if (c < m && j <= i) X[b * m * m + r * m + c] = Dinv[b * dstride + blk * CH_NB * CH_NB + i * CH_NB + j];
}

// A[b][i][j] = A[b][j][i] for j > i (32 x 32 tiles through shared memory: coalesced on both sides)
__global__ void mirror_lower_kernel(double* __restrict__ A, int64_t m) {
    __shared__ double tile[32][33];
    const int64_t bi = blockIdx.y, bj = blockIdx.x;   // destination tile (rows bi, cols bj), strictly-upper elements
    if (bj < bi) return;
    double* Ab = A + (int64_t)blockIdx.z * m * m;
    for (int r = threadIdx.y; r < 32; r += 8) {        // source tile (rows bj, cols bi)
        const int64_t sr = bj * 32 + r, sc = bi * 32 + threadIdx.x;
        tile[r][threadIdx.x] = (sr < m && sc < m) ? Ab[sr * m + sc] : 0.0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int64_t dr = bi * 32 + r, dc = bj * 32 + threadIdx.x;
        if (dr < m && dc < m && dc > dr) Ab[dr * m + dc] = tile[threadIdx.x][r];
    }
}

struct CholPlan {
    int64_t nblk;
    int64_t off_work, off_dinv, off_x, off_s, total;
};

static CholPlan chol_plan(int64_t batch, int64_t m, int want_inverse) {
    CholPlan p;
    p.nblk = (m + CH_NB - 1) / CH_NB;
    int64_t o = 0;
    auto take = [&](int64_t bytes) {
        int64_t r = o;
        o += align_up(bytes, 256);
        return r;
    };
    const bool blocked = m > CH_NB;
    p.off_work = take(blocked ? batch * m * m * 8 : 256);                         // trailing-update copy of A
    p.off_dinv = take(batch * p.nblk * CH_NB * CH_NB * 8);                        // inv of diagonal blocks
    p.off_x = take(blocked && want_inverse ? batch * m * m * 8 : 256);            // X = L^-1
    p.off_s = take(blocked && want_inverse ? batch * (m / 2 + CH_NB) * (m / 2 + CH_NB) * 8 : 256);   // L21 X11 of the recursion
    p.total = o;
    return p;
}

// factor (and optionally invert).  A (batch,m,m) is not modified.  L (batch,m,m) receives the factor
// (may be workspace when only the inverse is wanted).
static int chol_run(const double* A, int64_t batch, int64_t m, double* L, double* Ainv, int* info, void* ws,
                    cudaStream_t st) {
    const int want_inverse = Ainv != nullptr;
    CholPlan pl = chol_plan(batch, m, want_inverse);
    char* base = (char*)ws;
    double* work = (double*)(base + pl.off_work);
    double* Dinv = (double*)(base + pl.off_dinv);
    double* X = (double*)(base + pl.off_x);
    double* S = (double*)(base + pl.off_s);
    const int64_t dstride = pl.nblk * CH_NB * CH_NB;
    if (info) TES_CUDA_OK(cudaMemsetAsync(info, 0, sizeof(int) * batch, st));
    TES_CUDA_OK(cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CH_SMEM));
    if (m <= CH_NB) {
        potrf_diag_kernel<<<(unsigned)batch, 256, CH_SMEM, st>>>(A, m, 0, (int)m, L, Dinv, dstride, info, Ainv);
        TES_LAUNCH_OK();
        return 0;
    }
    TES_CUDA_OK(cudaMemcpyAsync(work, A, sizeof(double) * batch * m * m, cudaMemcpyDeviceToDevice, st));
    const int64_t mm = m * m;
    int rc;
    // LEFT-looking: block column k is updated by ONE GEMM with K = k0 (all previous block columns) right before it is
    // factored.  The right-looking form (a rank-64 update of the whole trailing matrix after every panel) runs the DMMA
    // GEMM with K = 64 -- four slices per tile, prologue and epilogue exposed: ~40 % of the pipe; here K grows to m - 64.
    for (int64_t kb = 0; kb < pl.nblk; ++kb) {
        const int64_t k0 = kb * CH_NB;
        const int nb = (int)((m - k0 < CH_NB) ? (m - k0) : CH_NB);
        if (kb > 0) {
            // work[k0:, k0:k0+nb] -= L[k0:, :k0] L[k0:k0+nb, :k0]^T
            GemmArgs u{};
            u.A = L + k0 * m; u.lda = m; u.transa = 0; u.batchA = mm;
            u.B = L + k0 * m; u.sbk = 1; u.sbn = m; u.batchB = mm;      // B'[k][c] = L[k0 + c][k]
            u.C = work + k0 * m + k0; u.ldc = m; u.batchC = mm;
            u.M = m - k0; u.N = nb; u.K = k0; u.alpha = -1.0; u.beta = 1.0; u.tri = 0; u.batch = (int)batch;
            if ((rc = launch_gemm_ex(u, st))) return rc;
        }
        potrf_diag_kernel<<<(unsigned)batch, 256, CH_SMEM, st>>>(work, m, k0, nb, L, Dinv + kb * CH_NB * CH_NB, dstride,
                                                           info, nullptr);
        TES_LAUNCH_OK();
        const int64_t rest = m - k0 - nb;
        if (rest <= 0) break;
        // panel: L[k0+nb:, k0:k0+nb] = work[k0+nb:, k0:k0+nb] * inv(L_kk)^T
        GemmArgs g{};
        g.A = work + (k0 + nb) * m + k0; g.lda = m; g.transa = 0; g.batchA = mm;
        g.B = Dinv + kb * CH_NB * CH_NB; g.sbk = 1; g.sbn = CH_NB; g.batchB = dstride;  // B'[k][c] = Dinv[c][k]
        g.C = L + (k0 + nb) * m + k0; g.ldc = m; g.batchC = mm;
        g.M = rest; g.N = nb; g.K = nb; g.alpha = 1.0; g.beta = 0.0; g.tri = 0; g.batch = (int)batch;
        if ((rc = launch_gemm_ex(g, st))) return rc;
    }
    {
        int64_t tot = batch * mm;
        zero_upper_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(L, m, batch);
        TES_LAUNCH_OK();
    }
    if (!want_inverse) return 0;
    // X = L^-1 by block rows: X_ii = inv(L_ii); X[i, :i] = -inv(L_ii) * (L[i, :i] * X[:i, :i])
    TES_CUDA_OK(cudaMemsetAsync(X, 0, sizeof(double) * batch * mm, st));
    {
        int64_t tot = batch * m * CH_NB;
        copy_diag_inv_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(Dinv, dstride, X, m, batch);
        TES_LAUNCH_OK();
    }
    // RECURSIVE halving: for a diagonal range [r0, r0 + n) split at s (a multiple of 64) the off-diagonal block of the
    // inverse is X21 = -X22 (L21 X11) -- two GEMMs of (n - s) x s x s and (n - s) x s x (n - s) with both triangular
    // factors skipped from their zero side, after the two halves themselves.  Square-ish GEMMs with K up to m / 2
    // instead of the block-row sweep whose GEMMs had M = 64 rows (half of every 128-row tile idle) .
    struct Range { int64_t r0, n; int phase; };
    {
        Range stack[64];
        int sp = 0;
        stack[sp++] = {0, m, 0};
        while (sp > 0) {
            Range r = stack[--sp];
            if (r.n <= CH_NB) continue;                      // a diagonal block: already inv(L_ii)
            const int64_t nblk_r = (r.n + CH_NB - 1) / CH_NB;
            const int64_t s_ = (nblk_r / 2) * CH_NB;         // first half: whole 64-blocks
            if (r.phase == 0) {
                if (sp + 3 > 64) return -3;
                stack[sp++] = {r.r0, r.n, 1};                // combine after both halves
                stack[sp++] = {r.r0 + s_, r.n - s_, 0};
                stack[sp++] = {r.r0, s_, 0};
                continue;
            }
            const int64_t a0 = r.r0, b0 = r.r0 + s_, nb2 = r.n - s_;
            // S = L21 X11        (nb2 x s_, K = s_; X11 lower-triangular: rows k < col contribute nothing)
            GemmArgs g{};
            g.A = L + b0 * m + a0; g.lda = m; g.transa = 0; g.batchA = mm;
            g.B = X + a0 * m + a0; g.sbk = m; g.sbn = 1; g.batchB = mm;
            g.C = S; g.ldc = s_; g.batchC = (m / 2 + CH_NB) * (m / 2 + CH_NB);
            g.M = nb2; g.N = s_; g.K = s_; g.alpha = 1.0; g.beta = 0.0; g.tri = 0; g.batch = (int)batch;
            g.kskip = 2;
            if ((rc = launch_gemm_ex(g, st))) return rc;
            // X21 = -X22 S       (nb2 x s_, K = nb2)
            GemmArgs h{};
            h.A = X + b0 * m + b0; h.lda = m; h.transa = 0; h.batchA = mm;
            h.B = S; h.sbk = s_; h.sbn = 1; h.batchB = (m / 2 + CH_NB) * (m / 2 + CH_NB);
            h.C = X + b0 * m + a0; h.ldc = m; h.batchC = mm;
            h.M = nb2; h.N = s_; h.K = nb2; h.alpha = -1.0; h.beta = 0.0; h.tri = 0; h.batch = (int)batch;
            h.kskip = 3;     // X22 is lower-triangular
            if ((rc = launch_gemm_ex(h, st))) return rc;
        }
    }
    // A^-1 = X^T X: lower tiles only, k from max(row0, col0) (X lower-triangular), then mirrored (X^T X is
    // bit-symmetric: the same products in the same order) -- m^3/3 flops instead of 2 m^3
    GemmArgs f{};
    f.A = X; f.lda = m; f.transa = 1; f.batchA = mm;
    f.B = X; f.sbk = m; f.sbn = 1; f.batchB = mm;
    f.C = Ainv; f.ldc = m; f.batchC = mm;
    f.M = m; f.N = m; f.K = m; f.alpha = 1.0; f.beta = 0.0; f.tri = 1; f.kskip = 1; f.batch = (int)batch;
    if ((rc = launch_gemm_ex(f, st))) return rc;
    dim3 mg((unsigned)((m + 31) / 32), (unsigned)((m + 31) / 32), (unsigned)batch);
    mirror_lower_kernel<<<mg, dim3(32, 8), 0, st>>>(Ainv, m);
    TES_LAUNCH_OK();
    return 0;
}

}  // namespace tes

using namespace tes;

extern "C" int64_t tes_batched_chol_workspace_bytes(int64_t batch, int64_t m, int want_inverse) {
    if (batch < 0 || m < 1) return -1;
    // when only the inverse is wanted the factor itself also lives in the workspace
    return chol_plan(batch, m, want_inverse).total + (want_inverse ? align_up(batch * m * m * 8, 256) : 0);
}

extern "C" int tes_batched_potrf_f64(const double* A, int64_t batch, int64_t m, double* L, int* info, void* ws,
                                     int64_t ws_bytes, void* stream) {
    if (batch < 0 || (batch > 0 && !A)) return -1;
    if (m < 1 || m > 32768) return -3;
    if (batch > 0 && !L) return -4;
    if (batch > 65535 * 32) return -2;
    if (!ws || ws_bytes < tes_batched_chol_workspace_bytes(batch, m, 0)) return -100;
    if (batch == 0) return 0;
    return chol_run(A, batch, m, L, nullptr, info, ws, (cudaStream_t)stream);
}

extern "C" int tes_batched_chol2inv_f64(const double* A, int64_t batch, int64_t m, double* Ainv, int* info, void* ws,
                                        int64_t ws_bytes, void* stream) {
    if (batch < 0 || (batch > 0 && !A)) return -1;
    if (m < 1 || m > 32768) return -3;
    if (batch > 0 && !Ainv) return -4;
    if (!ws || ws_bytes < tes_batched_chol_workspace_bytes(batch, m, 1)) return -100;
    if (batch == 0) return 0;
    CholPlan pl = chol_plan(batch, m, 1);
    double* L = (double*)((char*)ws + pl.total);
    return chol_run(A, batch, m, L, Ainv, info, ws, (cudaStream_t)stream);
}
