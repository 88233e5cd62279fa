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
constexpr size_t CH_SMEM = 2 * CH_NB * CH_LD * sizeof(double) + 16;

// One CTA per matrix.  Loads the nb x nb diagonal block at (k0,k0) of A (ld = m), factors it in
// shared memory, writes L (lower, upper zeroed) to Lout at the same position and inv(L) to
// Dinv[b][blk] (CH_NB x CH_NB, ld CH_NB, zero padded).  If ainv != NULL (m <= 64 path) also writes
// inv(L)^T inv(L) there.
__global__ void __launch_bounds__(256)
    potrf_diag_kernel(const double* __restrict__ A, int64_t m, int64_t k0, int nb, double* __restrict__ Lout,
                      double* __restrict__ Dinv, int64_t dinv_batch_stride, int* __restrict__ info,
                      double* __restrict__ ainv) {
    extern __shared__ __align__(16) double chsm[];
    double (*a)[CH_LD] = reinterpret_cast<double (*)[CH_LD]>(chsm);
    double (*x)[CH_LD] = a + CH_NB;
    int& bad = *reinterpret_cast<int*>(x + CH_NB);
    const int64_t b = blockIdx.x;
    const double* Ab = A + b * m * m;
    const int t = threadIdx.x;
    if (t == 0) bad = 0;
    for (int e = t; e < CH_NB * CH_NB; e += 256) {
        int i = e / CH_NB, j = e % CH_NB;
        a[i][j] = (i < nb && j < nb && j <= i) ? Ab[(k0 + i) * m + k0 + j] : 0.0;
        x[i][j] = 0.0;
    }
    __syncthreads();
    for (int k = 0; k < nb; ++k) {
        if (t == 0) {
            double dkk = a[k][k];
            if (!(dkk > 0.0) && bad == 0) bad = k + 1;
            a[k][k] = sqrt(dkk);
        }
        __syncthreads();
        const double inv = 1.0 / a[k][k];
        for (int i = k + 1 + t; i < nb; i += 256) a[i][k] *= inv;
        __syncthreads();
        // trailing update of the lower triangle: pairs (i, j), k < j <= i < nb
        const int r = nb - k - 1;
        for (int e = t; e < r * r; e += 256) {
            int i = k + 1 + e / r, j = k + 1 + e % r;
            if (j <= i) a[i][j] = fma(-a[i][k], a[j][k], a[i][j]);
        }
        __syncthreads();
    }
    // triangular inverse, one thread per column
    if (t < nb) {
        const int j = t;
        x[j][j] = 1.0 / a[j][j];
        for (int i = j + 1; i < nb; ++i) {
            double s = 0.0;
            for (int k = j; k < i; ++k) s = fma(a[i][k], x[k][j], s);
            x[i][j] = -s / a[i][i];
        }
    }
    __syncthreads();
    double* Lb = Lout + b * m * m;
    for (int e = t; e < nb * nb; e += 256) {
        int i = e / nb, j = e % nb;
        Lb[(k0 + i) * m + k0 + j] = (j <= i) ? a[i][j] : 0.0;
    }
    if (Dinv) {
        double* Db = Dinv + b * dinv_batch_stride;
        for (int e = t; e < CH_NB * CH_NB; e += 256) Db[e] = x[e / CH_NB][e % CH_NB];
    }
    if (ainv) {
        double* Ib = ainv + b * m * m;
        for (int e = t; e < nb * nb; e += 256) {
            int i = e / nb, j = e % nb;
            int lo = i > j ? i : j;
            double s = 0.0;
            for (int k = lo; k < nb; ++k) s = fma(x[k][i], x[k][j], s);
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
    p.off_s = take(blocked && want_inverse ? batch * CH_NB * m * 8 : 256);        // block-row temporary
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
    for (int64_t kb = 0; kb < pl.nblk; ++kb) {
        const int64_t k0 = kb * CH_NB;
        const int nb = (int)((m - k0 < CH_NB) ? (m - k0) : CH_NB);
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
        // update: work[k0+nb:, k0+nb:] -= P P^T   (lower tiles only)
        GemmArgs u{};
        u.A = L + (k0 + nb) * m + k0; u.lda = m; u.transa = 0; u.batchA = mm;
        u.B = L + (k0 + nb) * m + k0; u.sbk = 1; u.sbn = m; u.batchB = mm;  // B'[k][c] = P[c][k]
        u.C = work + (k0 + nb) * m + (k0 + nb); u.ldc = m; u.batchC = mm;
        u.M = rest; u.N = rest; u.K = nb; u.alpha = -1.0; u.beta = 1.0; u.tri = 1; u.batch = (int)batch;
        if ((rc = launch_gemm_ex(u, st))) return rc;
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
    for (int64_t ib = 1; ib < pl.nblk; ++ib) {
        const int64_t i0 = ib * CH_NB;
        const int nb = (int)((m - i0 < CH_NB) ? (m - i0) : CH_NB);
        GemmArgs g{};
        g.A = L + i0 * m; g.lda = m; g.transa = 0; g.batchA = mm;
        g.B = X; g.sbk = m; g.sbn = 1; g.batchB = mm;
        g.C = S; g.ldc = m; g.batchC = CH_NB * m;
        g.M = nb; g.N = i0; g.K = i0; g.alpha = 1.0; g.beta = 0.0; g.tri = 0; g.batch = (int)batch;
        if ((rc = launch_gemm_ex(g, st))) return rc;
        GemmArgs h{};
        h.A = Dinv + ib * CH_NB * CH_NB; h.lda = CH_NB; h.transa = 0; h.batchA = dstride;
        h.B = S; h.sbk = m; h.sbn = 1; h.batchB = CH_NB * m;
        h.C = X + i0 * m; h.ldc = m; h.batchC = mm;
        h.M = nb; h.N = i0; h.K = nb; h.alpha = -1.0; h.beta = 0.0; h.tri = 0; h.batch = (int)batch;
        if ((rc = launch_gemm_ex(h, st))) return rc;
    }
    // A^-1 = X^T X
    GemmArgs f{};
    f.A = X; f.lda = m; f.transa = 1; f.batchA = mm;
    f.B = X; f.sbk = m; f.sbn = 1; f.batchB = mm;
    f.C = Ainv; f.ldc = m; f.batchC = mm;
    f.M = m; f.N = m; f.K = m; f.alpha = 1.0; f.beta = 0.0; f.tri = 0; f.batch = (int)batch;
    return launch_gemm_ex(f, st);
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
