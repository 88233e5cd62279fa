// rffdraw.cu -- the random-Fourier-feature posterior weight draw on the device.
//
// Reference: optfunc.draw_random_init_weights_features_np (optfunc.py:303-385), host numpy there: for every function
// sample j, W_j = randn * sqrt(l), b_j = 2 pi rand, Z_j = sqrt(2 sigma/F) cos(W_j X^T + b_j) (F x n) and
//   n <  F:  Sigma_j = Z_j^T Z_j + sigma0 I (n x n),  theta_j = eps - Z V r V^T Z^T eps + Z Sigma^-1 y,
//            (e, V) = eig(Sigma_j),  r = 1 / (sqrt(e) (sqrt(e) + sqrt(sigma0)))          (:348-374)
//   n >= F:  Sigma_j = inv(Z_j Z_j^T / sigma0 + I) (F x F),  theta_j = Sigma Z y / sigma0 + chol(Sigma) eps   (:377-383)
// The three random draws (randn W, rand b, randn eps) are inputs, in the reference's draw order.
//
// Device plan (S matrices at once): Zt (S, n, F) is materialised once (features contiguous, so Zt Zt^T, Zt eps and
// Zt^T t are all coalesced), the S Gram matrices come from ONE batched fp64 GEMM launch, the eigen-decompositions
// from the batched Jacobi kernel of maxstats.cu (matrix functions V f(e) V^T do not depend on the eigenvector order
// or sign, so any correct solver reproduces the reference), and one CTA per sample finishes theta.
#include "gp.cuh"
#include "tes_common.cuh"

extern "C" int64_t tes_sym_eig_workspace_bytes(int64_t batch, int64_t n);
extern "C" int tes_sym_eig_f64(const double* A, int64_t batch, int64_t n, double* evals, double* evecs, int* sweeps,
                               void* ws, int64_t ws_bytes, void* stream);

namespace tes {

__global__ void rff_wb_kernel(const double* __restrict__ randn_W, const double* __restrict__ rand_b,
                              const double* __restrict__ l, int64_t SF, int d, double* __restrict__ W,
                              double* __restrict__ b) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= SF * d) return;
    W[e] = __dmul_rn(randn_W[e], __dsqrt_rn(l[e % d]));                    // optfunc.py:329
    if (e < SF) b[e] = __dmul_rn(__dmul_rn(2.0, 3.141592653589793), rand_b[e]);  // optfunc.py:334
}

// Zt[j][i][f] = scale * cos(W[j][f] . X[i] + b[j][f]); thread = (j, f), X staged in shared memory
__global__ void __launch_bounds__(256) rff_zt_kernel(const double* __restrict__ W, const double* __restrict__ b,
                                                     const double* __restrict__ X, int64_t S, int64_t F, int n, int d,
                                                     double scale, double* __restrict__ Zt) {
    extern __shared__ double sX[];
    for (int e = threadIdx.x; e < n * d; e += blockDim.x) sX[e] = X[e];
    __syncthreads();
    const int64_t j = blockIdx.y;
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const double* w = W + (j * F + f) * d;
    const double bb = b[j * F + f];
    double* out = Zt + j * n * F + f;
    for (int i = 0; i < n; ++i) {
        double h = 0.0;
        for (int k = 0; k < d; ++k) h = fma(w[k], sX[i * d + k], h);
        out[(int64_t)i * F] = scale * cos(h + bb);
    }
}

__global__ void add_diag_kernel(double* __restrict__ A, int64_t batch, int n, double v) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= batch * n) return;
    A[(e / n) * n * n + (e % n) * (int64_t)(n + 1)] += v;
}

// n < F branch, one CTA per sample.  smem: u[n], w[n], t[n]
__global__ void __launch_bounds__(256) rff_theta_kernel(const double* __restrict__ Zt, const double* __restrict__ evals,
                                                        const double* __restrict__ evecs, const double* __restrict__ Y,
                                                        const double* __restrict__ noise, int64_t F, int n,
                                                        double sigma0, double* __restrict__ theta) {
    extern __shared__ double sm[];
    double* u = sm;
    double* w = sm + n;
    double* t = sm + 2 * n;
    const int64_t j = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const double* Z = Zt + j * n * F;
    const double* eps = noise + j * F;
    const double* V = evecs + j * n * n;
    const double* e = evals + j * n;
    // u = Z^T eps   (optfunc.py:373 innermost product)
    for (int i = warp; i < n; i += nwarps) {
        double s = 0.0;
        for (int64_t f = lane; f < F; f += 32) s = fma(Z[(int64_t)i * F + f], eps[f], s);
        s = warp_sum(s);
        if (lane == 0) u[i] = s;
    }
    __syncthreads();
    // w = (V^T y) / e - r (V^T u)
    const double s0 = sqrt(sigma0);
    for (int k = tid; k < n; k += blockDim.x) {
        double a = 0.0, c = 0.0;
        for (int i = 0; i < n; ++i) {
            const double v = V[(int64_t)i * n + k];
            a = fma(v, Y[i], a);
            c = fma(v, u[i], c);
        }
        const double se = sqrt(e[k]);
        const double r = 1.0 / (se * (se + s0));
        w[k] = a / e[k] - r * c;
    }
    __syncthreads();
    // t = V w
    for (int i = warp; i < n; i += nwarps) {
        double s = 0.0;
        for (int k = lane; k < n; k += 32) s = fma(V[(int64_t)i * n + k], w[k], s);
        s = warp_sum(s);
        if (lane == 0) t[i] = s;
    }
    __syncthreads();
    // theta = eps + Z t
    for (int64_t f = tid; f < F; f += blockDim.x) {
        double s = 0.0;
        for (int i = 0; i < n; ++i) s = fma(Z[(int64_t)i * F + f], t[i], s);
        theta[j * F + f] = eps[f] + s;
    }
}

// n >= F branch: q = Z y (F), one CTA per sample
__global__ void __launch_bounds__(256) rff_zy_kernel(const double* __restrict__ Zt, const double* __restrict__ Y,
                                                     int64_t F, int n, double* __restrict__ q) {
    const int64_t j = blockIdx.x;
    const double* Z = Zt + j * n * F;
    for (int64_t f = threadIdx.x; f < F; f += blockDim.x) {
        double s = 0.0;
        for (int i = 0; i < n; ++i) s = fma(Z[(int64_t)i * F + f], Y[i], s);
        q[j * F + f] = s;
    }
}
// theta = Sigma q / sigma0 + L eps   (Sigma symmetric F x F, L lower Cholesky factor of Sigma), warp per row
__global__ void __launch_bounds__(256) rff_theta_big_kernel(const double* __restrict__ Sigma,
                                                            const double* __restrict__ L, const double* __restrict__ q,
                                                            const double* __restrict__ noise, int64_t F, double sigma0,
                                                            double* __restrict__ theta) {
    const int64_t j = blockIdx.y;
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= F) return;
    const double* Sj = Sigma + (j * F + r) * F;
    const double* Lj = L + (j * F + r) * F;
    double a = 0.0, c = 0.0;
    for (int64_t k = lane; k < F; k += 32) {
        a = fma(Sj[k], q[j * F + k], a);
        if (k <= r) c = fma(Lj[k], noise[j * F + k], c);
    }
    a = warp_sum(a);
    c = warp_sum(c);
    if (lane == 0) theta[j * F + r] = a / sigma0 + c;
}

}  // namespace tes

using namespace tes;

extern "C" int64_t tes_rff_draw_workspace_bytes(int64_t S, int64_t F, int64_t n, int64_t d) {
    if (S < 1 || F < 1 || n < 1 || d < 1) return -1;
    const int64_t zt = align_up(S * n * F * 8, 256);
    if (n < F) {
        const int64_t eig_ws = tes_sym_eig_workspace_bytes(S, n);
        if (eig_ws < 0) return -1;
        return zt + align_up(S * n * n * 8, 256) * 2 + align_up(S * n * 8, 256) + eig_ws;
    }
    const int64_t chol = tes_batched_chol_workspace_bytes(S, F, 1);
    if (chol < 0) return -1;
    return zt + 3 * align_up(S * F * F * 8, 256) + align_up(S * F * 8, 256) + align_up(S * 4, 256) + chol;
}

extern "C" int tes_rff_draw_f64(const double* X, const double* Y, int64_t n, int64_t d, const double* l, double sigma,
                                double sigma0, const double* randn_W, const double* rand_b, const double* randn_noise,
                                int64_t S, int64_t F, double* theta, double* W, double* b, void* ws, int64_t ws_bytes,
                                void* stream) {
    if (!X) return -1;
    if (!Y) return -2;
    if (n < 1 || n > 4096) return -3;
    if (d < 1 || d > 64) return -4;
    if (!l) return -5;
    if (!(sigma0 > 0.0)) return -7;
    if (!randn_W) return -8;
    if (!rand_b) return -9;
    if (!randn_noise) return -10;
    if (S < 1 || S > 65535) return -11;
    if (F < 1) return -12;
    if (!theta) return -13;
    if (!W) return -14;
    if (!b) return -15;
    const int64_t need = tes_rff_draw_workspace_bytes(S, F, n, d);
    if (need < 0) return -3;
    if (!ws || ws_bytes < need) return -100;
    if ((size_t)(n * d) * 8 > 200 * 1024) return -3;
    cudaStream_t st = (cudaStream_t)stream;
    char* base = (char*)ws;
    double* Zt = (double*)base;
    base += align_up(S * n * F * 8, 256);
    const double scale = sqrt(2.0 * sigma / (double)F);

    rff_wb_kernel<<<(unsigned)((S * F * d + 255) / 256), 256, 0, st>>>(randn_W, rand_b, l, S * F, (int)d, W, b);
    TES_LAUNCH_OK();
    const size_t smx = (size_t)n * d * 8;
    if (smx > 48 * 1024)
        TES_CUDA_OK(cudaFuncSetAttribute(rff_zt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smx));
    rff_zt_kernel<<<dim3((unsigned)((F + 255) / 256), (unsigned)S), 256, smx, st>>>(W, b, X, S, F, (int)n, (int)d, scale,
                                                                                   Zt);
    TES_LAUNCH_OK();
    if (n < F) {
        double* Sig = (double*)base;
        base += align_up(S * n * n * 8, 256);
        double* evecs = (double*)base;
        base += align_up(S * n * n * 8, 256);
        double* evals = (double*)base;
        base += align_up(S * n * 8, 256);
        GemmArgs g{};
        g.A = Zt; g.lda = F; g.transa = 0; g.batchA = n * F;
        g.B = Zt; g.sbk = 1; g.sbn = F; g.batchB = n * F;
        g.C = Sig; g.ldc = n; g.batchC = n * n;
        g.M = n; g.N = n; g.K = F; g.alpha = 1.0; g.beta = 0.0; g.tri = 0; g.batch = (int)S;
        int rc = launch_gemm_ex(g, st);
        if (rc) return rc;
        add_diag_kernel<<<(unsigned)((S * n + 255) / 256), 256, 0, st>>>(Sig, S, (int)n, sigma0);
        TES_LAUNCH_OK();
        const int64_t eig_ws = tes_sym_eig_workspace_bytes(S, n);
        rc = tes_sym_eig_f64(Sig, S, n, evals, evecs, nullptr, base, eig_ws, stream);
        if (rc) return rc;
        rff_theta_kernel<<<(unsigned)S, 256, 3 * n * sizeof(double), st>>>(Zt, evals, evecs, Y, randn_noise, F, (int)n,
                                                                          sigma0, theta);
        TES_LAUNCH_OK();
        return 0;
    }
    // n >= F
    double* Mx = (double*)base;
    base += align_up(S * F * F * 8, 256);
    double* Sig = (double*)base;
    base += align_up(S * F * F * 8, 256);
    double* Lc = (double*)base;
    base += align_up(S * F * F * 8, 256);
    double* q = (double*)base;
    base += align_up(S * F * 8, 256);
    int* info = (int*)base;
    base += align_up(S * 4, 256);
    const int64_t chol_ws = tes_batched_chol_workspace_bytes(S, F, 1);
    GemmArgs g{};
    g.A = Zt; g.lda = F; g.transa = 1; g.batchA = n * F;   // op(A) = Zt^T = Z (F x n)
    g.B = Zt; g.sbk = F; g.sbn = 1; g.batchB = n * F;      // B'[k][col] = Zt[k][col]
    g.C = Mx; g.ldc = F; g.batchC = F * F;
    g.M = F; g.N = F; g.K = n; g.alpha = 1.0 / sigma0; g.beta = 0.0; g.tri = 0; g.batch = (int)S;
    int rc = launch_gemm_ex(g, st);
    if (rc) return rc;
    add_diag_kernel<<<(unsigned)((S * F + 255) / 256), 256, 0, st>>>(Mx, S, (int)F, 1.0);
    TES_LAUNCH_OK();
    rc = tes_batched_chol2inv_f64(Mx, S, F, Sig, info, base, chol_ws, stream);
    if (rc) return rc;
    rc = tes_batched_potrf_f64(Sig, S, F, Lc, info, base, chol_ws, stream);
    if (rc) return rc;
    rff_zy_kernel<<<(unsigned)S, 256, 0, st>>>(Zt, Y, F, (int)n, q);
    TES_LAUNCH_OK();
    rff_theta_big_kernel<<<dim3((unsigned)((F + 7) / 8), (unsigned)S), 256, 0, st>>>(Sig, Lc, q, randn_noise, F, sigma0,
                                                                                    theta);
    TES_LAUNCH_OK();
    return 0;
}
