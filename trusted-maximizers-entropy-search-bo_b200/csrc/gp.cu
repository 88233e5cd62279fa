// gp.cu -- Stage B of the TES hot path: SE-ARD kernels and the posterior statistics every TES
// criterion starts from.
//
//   tes_se_ard_cross_f64    utils.computeKnm   (utils.py:703-719)
//   tes_pnk_f64             get_pNK_test_obs   (criteria/evaluate_mp.py:12-53, the K + noise part)
//   tes_posterior_stats_f64 the common head of get_queried_f_stat_given_data_maxidx
//                           (evaluate_mp.py:76-90) and get_queried_f_stat_given_test_samples
//                           (evaluate_sample_mp.py:77-124): M = k invpNK[:, :T], b = k invpNK[:, T:] Y,
//                           Kq = sigma - rowsum((k invpNK) o k)
//   tes_gp_mean_var_f64     utils.compute_mean_var_f, fullcov=False (utils.py:786-815)
//   tes_gemm_f64            plain fp64 GEMM (the contraction primitive; exposed for tests)
#include "gemm.cuh"
#include "gp.cuh"

namespace tes {

// k[i][c] = sigma * exp(-0.5 * (Qbar[c] + Q[i] - 2 xs_i . xbs_c)), xs = x*sqrt(l) -- the reference's
// expansion (not the difference form), so rounding noise near duplicates matches its behaviour.
// If noise_from >= 0 and X == Xb (Gram use), sigma0 is added on the diagonal for c >= noise_from.
__global__ void se_ard_kernel(const double* __restrict__ x, int64_t N, const double* __restrict__ xb, int64_t m,
                              int d, const double* __restrict__ l, double sigma, int64_t noise_from, double sigma0,
                              double* __restrict__ out, int64_t ldo) {
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t i = (int64_t)blockIdx.y * blockDim.y + threadIdx.y;
    if (c >= m || i >= N) return;
    double q = 0.0, qb = 0.0, dot = 0.0;
    for (int k = 0; k < d; ++k) {
        double sl = sqrt(__ldg(l + k));
        double a = __dmul_rn(__ldg(x + i * d + k), sl);
        double bb = __dmul_rn(__ldg(xb + c * d + k), sl);
        q = __dadd_rn(q, __dmul_rn(a, a));
        qb = __dadd_rn(qb, __dmul_rn(bb, bb));
        dot = fma(a, bb, dot);
    }
    double dist = __dadd_rn(__dadd_rn(qb, q), -2.0 * dot);
    double v = sigma * exp(-0.5 * dist);
    if (noise_from >= 0 && i == c && c >= noise_from) v += sigma0;
    out[i * ldo + c] = v;
}

// Tiled form for d <= SE_MAXD: a CTA owns 128 (or 32) rows x 32 columns (a thread keeps its column's scaled coordinates in
// registers, the row's come as shared-memory broadcasts); sqrt(l), the scaled coordinates and the squared norms
// of its rows and columns are formed ONCE in shared memory (the kernel above recomputes d square roots and both norms
// for every element), then every element is d FMAs and one exp.  Same operations in the same order per element, so
// the values are the bits of se_ard_kernel (tests/test_stageB_gpu.py compares both).
constexpr int SE_MAXD = 16;
template <int SE_ROWS>
__global__ void __launch_bounds__(256)
    se_ard_tile_kernel(const double* __restrict__ x, int64_t N, const double* __restrict__ xb, int64_t m, int d,
                       const double* __restrict__ l, double sigma, int64_t noise_from, double sigma0,
                       double* __restrict__ out, int64_t ldo) {
    __shared__ double s_sl[SE_MAXD], s_b[32][SE_MAXD + 1], s_a[SE_ROWS][SE_MAXD + 1], s_qb[32], s_q[SE_ROWS];
    const int tx = threadIdx.x, ty = threadIdx.y, t = ty * 32 + tx;
    const int64_t c0 = (int64_t)blockIdx.x * 32, i0 = (int64_t)blockIdx.y * SE_ROWS;
    if (t < d) s_sl[t] = sqrt(__ldg(l + t));
    __syncthreads();
    for (int e = t; e < 32 * d; e += 256) {
        const int cc = e / d, k = e % d;
        const int64_t col = c0 + cc;
        s_b[cc][k] = col < m ? __dmul_rn(__ldg(xb + col * d + k), s_sl[k]) : 0.0;
    }
    for (int e = t; e < SE_ROWS * d; e += 256) {
        const int rr = e / d, k = e % d;
        const int64_t row = i0 + rr;
        s_a[rr][k] = row < N ? __dmul_rn(__ldg(x + row * d + k), s_sl[k]) : 0.0;
    }
    __syncthreads();
    if (t < 32) {
        double qb = 0.0;
        for (int k = 0; k < d; ++k) qb = __dadd_rn(qb, __dmul_rn(s_b[t][k], s_b[t][k]));
        s_qb[t] = qb;
    } else if (t < 32 + SE_ROWS) {
        double q = 0.0;
        for (int k = 0; k < d; ++k) q = __dadd_rn(q, __dmul_rn(s_a[t - 32][k], s_a[t - 32][k]));
        s_q[t - 32] = q;
    }
    __syncthreads();
    const int64_t c = c0 + tx;
    if (c >= m) return;
    const double qb = s_qb[tx];
    double bk[SE_MAXD];
#pragma unroll
    for (int k = 0; k < SE_MAXD; ++k) bk[k] = k < d ? s_b[tx][k] : 0.0;
#pragma unroll 2
    for (int r = ty; r < SE_ROWS; r += 8) {
        const int64_t i = i0 + r;
        if (i >= N) break;
        double dot = 0.0;
#pragma unroll
        for (int k = 0; k < SE_MAXD; ++k)
            if (k < d) dot = fma(s_a[r][k], bk[k], dot);
        const double dist = __dadd_rn(__dadd_rn(qb, s_q[r]), -2.0 * dot);
        double v = sigma * exp(-0.5 * dist);
        if (noise_from >= 0 && i == c && c >= noise_from) v += sigma0;
        out[i * ldo + c] = v;
    }
}

// TES_SE_ARD_TILED=0 keeps the per-element kernel (comparison path of the tests)
int launch_se_ard(const double* x, int64_t N, const double* xb, int64_t m, int d, const double* l,
                         double sigma, int64_t noise_from, double sigma0, double* out, int64_t ldo,
                         cudaStream_t st) {
    if (N == 0 || m == 0) return 0;
    const char* env = getenv("TES_SE_ARD_TILED");
    const bool tiled = d <= SE_MAXD && !(env && env[0] == '0');
    // 128-row tiles amortise the per-CTA set-up on the large chunks; small problems keep 32 rows (more CTAs)
    const bool big = tiled && ((N + 127) / 128) * ((m + 31) / 32) >= 4 * 148;
    const int rpb = tiled ? (big ? 128 : 32) : 8;              // rows per CTA
    dim3 block(32, 8);
    const int64_t rows_per = 65535LL * rpb;
    for (int64_t r0 = 0; r0 < N; r0 += rows_per) {             // grid.y <= 65535: split rows
        const int64_t nr = (N - r0 < rows_per) ? (N - r0) : rows_per;
        dim3 g2((unsigned)((m + 31) / 32), (unsigned)((nr + rpb - 1) / rpb));
        const int64_t nf = noise_from >= 0 ? noise_from - r0 : -1;
        if (tiled && big)
            se_ard_tile_kernel<128><<<g2, block, 0, st>>>(x + r0 * d, nr, xb, m, d, l, sigma, nf, sigma0, out + r0 * ldo, ldo);
        else if (tiled)
            se_ard_tile_kernel<32><<<g2, block, 0, st>>>(x + r0 * d, nr, xb, m, d, l, sigma, nf, sigma0, out + r0 * ldo, ldo);
        else
            se_ard_kernel<<<g2, block, 0, st>>>(x + r0 * d, nr, xb, m, d, l, sigma, nf, sigma0, out + r0 * ldo, ldo);
    }
    TES_LAUNCH_OK();
    return 0;
}

// ---- GEMM kernel: C = alpha * op(A) * B' + beta * C, batched over blockIdx.z ----------------------
// B' is addressed as B[k*sbk + col*sbn] so transposed B operands need no copy; op(A) = A (row-major,
// lda) or A^T (TRANSA: element (row,k) at A[k*lda + row]).  tri != 0 skips CTA tiles strictly above
// the diagonal (col0 > row0 + BM - 1), for symmetric updates whose upper part is never read.
template <class AProv>
__device__ __forceinline__ void gemm_body(const AProv& ap, const double* __restrict__ B, int64_t sbk, int64_t sbn,
                                          double* __restrict__ C, int64_t ldc, int64_t M, int64_t N, int64_t K,
                                          double alpha, double beta, int tri, int kskip = 0) {
    extern __shared__ __align__(16) unsigned char gemm_smem_raw[];
    GemmSmem& sm = *reinterpret_cast<GemmSmem*>(gemm_smem_raw);
    const int64_t row0 = (int64_t)blockIdx.x * GEMM_BM;
    const int64_t col0 = (int64_t)blockIdx.y * GEMM_BN;
    if (tri && col0 > row0 + GEMM_BM - 1) return;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    if (sbn == 1) {
        const int64_t kbeg = kskip == 1 ? (row0 > col0 ? row0 : col0) : (kskip == 2 ? col0 : 0);
        // kskip = 3: A is lower-triangular -- columns k > row contribute nothing to this row tile
        const int64_t kend = (kskip == 3 && row0 + GEMM_BM < K) ? row0 + GEMM_BM : K;
        gemm_mainloop(ap, B, sbk, kend, N, row0, col0, sm, acc, kbeg);
    } else {
        double ra[AProv::NREG], rb[4];
        const int64_t nk = (K + GEMM_BK - 1) / GEMM_BK;
        auto loadb = [&](int64_t k0) {
            const int kk = t >> 4, c = t & 15;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                int64_t k = k0 + kk, col = col0 + c + 16 * q;
                rb[q] = (k < K && col < N) ? __ldg(B + k * sbk + col * sbn) : 0.0;
            }
        };
        if (nk > 0) {
            ap.load(row0, 0, t, ra);
            loadb(0);
            ap.store(sm.a[0], t, ra);
            gemm_store_b(sm.b[0], t, rb);
        }
        __syncthreads();
        for (int64_t kt = 0; kt < nk; ++kt) {
            const int cur = (int)(kt & 1);
            if (kt + 1 < nk) {
                ap.load(row0, (kt + 1) * GEMM_BK, t, ra);
                loadb((kt + 1) * GEMM_BK);
            }
            gemm_mac_steps<0, GEMM_BK / 2>(sm.a[cur], sm.b[cur], warp, lane, acc);
            if (kt + 1 < nk) {
                ap.store(sm.a[cur ^ 1], t, ra);
                gemm_store_b(sm.b[cur ^ 1], t, rb);
            }
            gemm_mac_steps<GEMM_BK / 2, GEMM_BK>(sm.a[cur], sm.b[cur], warp, lane, acc);
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int64_t row = row0 + gemm_row_of(warp, lane, i);
        if (row >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                int64_t col = col0 + gemm_col_of(warp, lane, j, e);
                if (col < N) {
                    double v = alpha * acc[i][j][e];
                    if (beta != 0.0) v = fma(beta, C[row * ldc + col], v);
                    C[row * ldc + col] = v;
                }
            }
    }
}

template <bool TRANSA>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
    gemm_kernel(GemmArgs g) {
    const int64_t bz = blockIdx.z;
    const double* A = g.A + bz * g.batchA;
    const double* B = g.B + bz * g.batchB;
    double* C = g.C + bz * g.batchC;
    if (TRANSA) {
        ATrans ap{A, g.lda, g.M, g.K};
        gemm_body(ap, B, g.sbk, g.sbn, C, g.ldc, g.M, g.N, g.K, g.alpha, g.beta, g.tri, g.kskip);
    } else {
        APlain ap{A, g.lda, g.M, g.K};
        gemm_body(ap, B, g.sbk, g.sbn, C, g.ldc, g.M, g.N, g.K, g.alpha, g.beta, g.tri, g.kskip);
    }
}

int launch_gemm_ex(const GemmArgs& g, cudaStream_t st) {
    if (g.M <= 0 || g.N <= 0 || g.batch <= 0) return 0;
    dim3 grid((unsigned)((g.M + GEMM_BM - 1) / GEMM_BM), (unsigned)((g.N + GEMM_BN - 1) / GEMM_BN), (unsigned)g.batch);
    if (g.transa) {
        TES_CUDA_OK(cudaFuncSetAttribute(gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(GemmSmem)));
        TES_CUDA_OK(cudaFuncSetAttribute(gemm_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        gemm_kernel<true><<<grid, GEMM_THREADS, sizeof(GemmSmem), st>>>(g);
    } else {
        TES_CUDA_OK(cudaFuncSetAttribute(gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(GemmSmem)));
        TES_CUDA_OK(cudaFuncSetAttribute(gemm_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        gemm_kernel<false><<<grid, GEMM_THREADS, sizeof(GemmSmem), st>>>(g);
    }
    TES_LAUNCH_OK();
    return 0;
}

int launch_gemm(const double* A, int64_t lda, const double* B, int64_t sbk, int64_t sbn, double* C, int64_t ldc,
                int64_t M, int64_t N, int64_t K, cudaStream_t st) {
    GemmArgs g{};
    g.A = A; g.lda = lda; g.transa = 0; g.batchA = 0;
    g.B = B; g.sbk = sbk; g.sbn = sbn; g.batchB = 0;
    g.C = C; g.ldc = ldc; g.batchC = 0;
    g.M = M; g.N = N; g.K = K; g.alpha = 1.0; g.beta = 0.0; g.tri = 0; g.batch = 1;
    return launch_gemm_ex(g, st);
}

// out[i] = base - sum_c A[i][c]*B[i][c]  (optionally clipped below), one warp per row
__global__ void rowdot_kernel(const double* __restrict__ A, int64_t lda, const double* __restrict__ B, int64_t ldb,
                              int64_t N, int64_t m, double base, double clip_min, double* __restrict__ out) {
    int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (row >= N) return;
    double s = 0.0;
    for (int64_t c = lane; c < m; c += 32) s = fma(A[row * lda + c], B[row * ldb + c], s);
    s = warp_sum(s);
    if (lane == 0) {
        double v = base - s;
        if (v < clip_min) v = clip_min;  // NaN compares false and propagates, like tf.clip_by_value
        out[row] = v;
    }
}

// Bext = [invpNK | w] (m x (m+1)), w = invpNK[:, T:] * Y
__global__ void build_bext_kernel(const double* __restrict__ invpNK, const double* __restrict__ Y, int64_t m,
                                  int64_t T, double* __restrict__ Bext) {
    int64_t r = (int64_t)blockIdx.x;
    int lane = threadIdx.x;
    for (int64_t c = lane; c < m; c += blockDim.x) Bext[r * (m + 1) + c] = invpNK[r * m + c];
    double s = 0.0;
    for (int64_t c = T + lane; c < m; c += blockDim.x) s = fma(invpNK[r * m + c], Y[c - T], s);
    s = warp_sum(s);
    if (lane == 0) Bext[r * (m + 1) + m] = s;
}

__global__ void concat_rows_kernel(const double* __restrict__ a, int64_t na, const double* __restrict__ b,
                                   int64_t nb, int d, double* __restrict__ out) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (na + nb) * d) return;
    out[e] = e < na * d ? a[e] : b[e - na * d];
}

__global__ void extract_mb_kernel(const double* __restrict__ G, int64_t N, int64_t m, int64_t T,
                                  double* __restrict__ outM, double* __restrict__ outb) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N * (T + 1)) return;
    int64_t i = e / (T + 1), c = e % (T + 1);
    if (c < T)
        outM[i * T + c] = G[i * (m + 1) + c];
    else
        outb[i] = G[i * (m + 1) + m];
}

int launch_rowdot(const double* A, int64_t lda, const double* B, int64_t ldb, int64_t N, int64_t m, double base,
                  double clip_min, double* out, cudaStream_t st) {
    if (N == 0) return 0;
    rowdot_kernel<<<(unsigned)((N * 32 + 255) / 256), 256, 0, st>>>(A, lda, B, ldb, N, m, base, clip_min, out);
    TES_LAUNCH_OK();
    return 0;
}

int launch_concat_rows(const double* a, int64_t na, const double* b, int64_t nb, int d, double* out,
                       cudaStream_t st) {
    if ((na + nb) * d == 0) return 0;
    concat_rows_kernel<<<(unsigned)(((na + nb) * d + 255) / 256), 256, 0, st>>>(a, na, b, nb, d, out);
    TES_LAUNCH_OK();
    return 0;
}

int launch_build_bext(const double* invpNK, const double* Y, int64_t m, int64_t T, double* Bext, cudaStream_t st) {
    if (m == 0) return 0;
    build_bext_kernel<<<(unsigned)m, 32, 0, st>>>(invpNK, Y, m, T, Bext);
    TES_LAUNCH_OK();
    return 0;
}

}  // namespace tes

using namespace tes;

extern "C" int tes_se_ard_cross_f64(const double* x, int64_t N, const double* xb, int64_t m, int64_t d,
                                    const double* l, double sigma, double* out, void* stream) {
    if (N < 0 || (N > 0 && !x)) return -1;
    if (m < 0 || (m > 0 && !xb)) return -3;
    if (d < 1 || d > 64) return -5;
    if (!l) return -6;
    if (!out && N * m > 0) return -8;
    return launch_se_ard(x, N, xb, m, (int)d, l, sigma, -1, 0.0, out, m, (cudaStream_t)stream);
}

extern "C" int tes_gemm_f64(const double* A, int64_t lda, const double* B, int64_t ldb, int transb, double* C,
                            int64_t ldc, int64_t M, int64_t N, int64_t K, void* stream) {
    if (!A || !B || !C) return -1;
    if (M < 0 || N < 0 || K < 0) return -8;
    return launch_gemm(A, lda, B, transb ? 1 : ldb, transb ? ldb : 1, C, ldc, M, N, K, (cudaStream_t)stream);
}

// pNK = K([test_xs; X]) + sigma0 * diag(0_T, I_n)   (m x m, m = T + n); xto_out (m x d) receives the
// concatenated inputs (may be NULL).
extern "C" int64_t tes_pnk_workspace_bytes(int64_t T, int64_t n, int64_t d) { return align_up((T + n) * d * 8, 256); }

extern "C" int tes_pnk_f64(const double* test_xs, int64_t T, const double* X, int64_t n, int64_t d, const double* l,
                           double sigma, double sigma0, double* out_pNK, void* ws, int64_t ws_bytes, void* stream) {
    if (T < 0 || (T > 0 && !test_xs)) return -1;
    if (n < 0 || (n > 0 && !X)) return -3;
    if (d < 1 || d > 64) return -5;
    if (!l) return -6;
    if (!out_pNK) return -9;
    int64_t m = T + n;
    if (!ws || ws_bytes < tes_pnk_workspace_bytes(T, n, d)) return -100;
    cudaStream_t st = (cudaStream_t)stream;
    double* xto = (double*)ws;
    if (m == 0) return 0;
    concat_rows_kernel<<<(unsigned)((m * d + 255) / 256), 256, 0, st>>>(test_xs, T, X, n, (int)d, xto);
    TES_LAUNCH_OK();
    return launch_se_ard(xto, m, xto, m, (int)d, l, sigma, T, sigma0, out_pNK, m, st);
}

// ---- posterior statistics ------------------------------------------------------------------------

extern "C" int64_t tes_posterior_stats_workspace_bytes(int64_t N, int64_t T, int64_t n, int64_t d) {
    int64_t m = T + n;
    int64_t nc = N < PS_CHUNK ? N : PS_CHUNK;
    if (nc < 1) nc = 1;
    return align_up(m * d * 8, 256) + align_up(m * (m + 1) * 8, 256) + align_up(nc * m * 8, 256) +
           align_up(nc * (m + 1) * 8, 256);
}

extern "C" int tes_posterior_stats_f64(const double* x, int64_t N, int64_t d, const double* test_xs, int64_t T,
                                       const double* X, int64_t n, const double* Y, const double* l, double sigma,
                                       const double* invpNK, double* out_M, double* out_b, double* out_Kq, void* ws,
                                       int64_t ws_bytes, void* stream) {
    if (N < 0 || (N > 0 && !x)) return -1;
    if (d < 1 || d > 64) return -3;
    if (T < 0 || (T > 0 && !test_xs)) return -4;
    if (n < 0 || (n > 0 && (!X || !Y))) return -6;
    if (!l) return -9;
    if (!invpNK) return -11;
    if (N > 0 && ((T > 0 && !out_M) || !out_b || !out_Kq)) return -12;
    if (!ws || ws_bytes < tes_posterior_stats_workspace_bytes(N, T, n, d)) return -100;
    if (N == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t m = T + n;
    char* p = (char*)ws;
    double* xto = (double*)p;
    p += align_up(m * d * 8, 256);
    double* Bext = (double*)p;
    p += align_up(m * (m + 1) * 8, 256);
    const int64_t ncmax = N < PS_CHUNK ? N : PS_CHUNK;
    double* Kc = (double*)p;
    p += align_up(ncmax * m * 8, 256);
    double* Gc = (double*)p;
    concat_rows_kernel<<<(unsigned)((m * d + 255) / 256), 256, 0, st>>>(test_xs, T, X, n, (int)d, xto);
    TES_LAUNCH_OK();
    build_bext_kernel<<<(unsigned)m, 32, 0, st>>>(invpNK, Y, m, T, Bext);
    TES_LAUNCH_OK();
    for (int64_t c0 = 0; c0 < N; c0 += PS_CHUNK) {
        int64_t nc = (N - c0 < PS_CHUNK) ? (N - c0) : PS_CHUNK;
        int rc = launch_se_ard(x + c0 * d, nc, xto, m, (int)d, l, sigma, -1, 0.0, Kc, m, st);
        if (rc) return rc;
        rc = launch_gemm(Kc, m, Bext, m + 1, 1, Gc, m + 1, nc, m + 1, m, st);
        if (rc) return rc;
        rowdot_kernel<<<(unsigned)((nc * 32 + 255) / 256), 256, 0, st>>>(Gc, m + 1, Kc, m, nc, m, sigma,
                                                                         -INFINITY, out_Kq + c0);
        TES_LAUNCH_OK();
        extract_mb_kernel<<<(unsigned)((nc * (T + 1) + 255) / 256), 256, 0, st>>>(Gc, nc, m, T, out_M + c0 * T,
                                                                                  out_b + c0);
        TES_LAUNCH_OK();
    }
    return 0;
}

// GP posterior mean and clipped marginal variance (utils.compute_mean_var_f, fullcov=False).
extern "C" int64_t tes_gp_mean_var_workspace_bytes(int64_t N, int64_t n) {
    int64_t nc = N < PS_CHUNK ? N : PS_CHUNK;
    if (nc < 1) nc = 1;
    return 2 * align_up(nc * n * 8, 256) + align_up(n * 8, 256);
}

namespace tes {
__global__ void matvec_kernel(const double* __restrict__ A, int64_t rows, int64_t cols, const double* __restrict__ v,
                              double* __restrict__ out) {
    int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (row >= rows) return;
    double s = 0.0;
    for (int64_t c = lane; c < cols; c += 32) s = fma(A[row * cols + c], v[c], s);
    s = warp_sum(s);
    if (lane == 0) out[row] = s;
}
}  // namespace tes

extern "C" int tes_gp_mean_var_f64(const double* x, int64_t N, int64_t d, const double* X, int64_t n,
                                   const double* Y, const double* l, double sigma, const double* NKinv,
                                   double* out_mean, double* out_var, void* ws, int64_t ws_bytes, void* stream) {
    if (N < 0 || (N > 0 && !x)) return -1;
    if (d < 1 || d > 64) return -3;
    if (n < 1 || !X) return -4;
    if (!Y) return -6;
    if (!l) return -7;
    if (!NKinv) return -9;
    if (N > 0 && (!out_mean || !out_var)) return -10;
    if (!ws || ws_bytes < tes_gp_mean_var_workspace_bytes(N, n)) return -100;
    if (N == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t ncmax = N < PS_CHUNK ? N : PS_CHUNK;
    char* p = (char*)ws;
    double* Kc = (double*)p;
    p += align_up(ncmax * n * 8, 256);
    double* Gc = (double*)p;
    p += align_up(ncmax * n * 8, 256);
    double* alpha = (double*)p;
    matvec_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, st>>>(NKinv, n, n, Y, alpha);
    TES_LAUNCH_OK();
    for (int64_t c0 = 0; c0 < N; c0 += PS_CHUNK) {
        int64_t nc = (N - c0 < PS_CHUNK) ? (N - c0) : PS_CHUNK;
        int rc = launch_se_ard(x + c0 * d, nc, X, n, (int)d, l, sigma, -1, 0.0, Kc, n, st);
        if (rc) return rc;
        rc = launch_gemm(Kc, n, NKinv, n, 1, Gc, n, nc, n, n, st);
        if (rc) return rc;
        rowdot_kernel<<<(unsigned)((nc * 32 + 255) / 256), 256, 0, st>>>(Gc, n, Kc, n, nc, n, sigma, 1e-100,
                                                                         out_var + c0);
        TES_LAUNCH_OK();
        matvec_kernel<<<(unsigned)((nc * 32 + 255) / 256), 256, 0, st>>>(Kc, nc, n, alpha, out_mean + c0);
        TES_LAUNCH_OK();
    }
    return 0;
}
