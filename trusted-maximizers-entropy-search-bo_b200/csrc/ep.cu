// ep.cu -- batched expectation propagation for the maximizer-conditioned posteriors on the device.
//
// Reference: ep.approximate_EP_np (ep.py:73-204) with its helpers (ep.py:8-70), host numpy/scipy there, called once
// per trusted maximizer by utils.get_testidxs_stats(mode='ep') (utils.py:1820-1825): EP for N(mu, Sigma) truncated
// to {f_j >= f_i for all i}, T-1 half-space sites c_i = e_j - e_i, ONE site updated per outer iteration, the Gaussian
// approximation refreshed by a full T x T inverse after every update, stop at mean-square change <= resolution or
// max_niter.  Here: one CTA per maximizer (all T of them concurrently), the T x T system matrix lives in shared
// memory and is inverted in place by Gauss-Jordan elimination with partial pivoting (the same pivoting rule as the
// LU behind np.linalg.inv); Sigma^-1 is computed once by the same routine and shared through L2.
#include <cstdlib>

#include "tes_common.cuh"

namespace tes {

constexpr int EP_THREADS = 512;

// In-place inverse of the n x n row-major matrix a (shared or global memory), Gauss-Jordan with partial pivoting.
// piv: n ints of shared scratch; red_v / red_i: one slot per warp.  All threads of the CTA must call.
__device__ void gj_inverse(double* a, int n, int* piv, double* red_v, int* red_i) {
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    for (int k = 0; k < n; ++k) {
        // pivot search in column k, rows k..n-1 (first maximum of |a_ik|, like LAPACK's idamax)
        double bv = -1.0;
        int bi = n;
        for (int i = k + tid; i < n; i += nt) {
            const double v = fabs(a[(int64_t)i * n + k]);
            if (v > bv || !(v == v)) {  // a NaN pivot candidate wins: the NaN then propagates like in LAPACK
                bv = v == v ? v : pos_inf();
                bi = i;
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, off);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
            if (ov > bv || (ov == bv && oi < bi)) {
                bv = ov;
                bi = oi;
            }
        }
        if (lane == 0) {
            red_v[warp] = bv;
            red_i[warp] = bi;
        }
        __syncthreads();
        if (tid == 0) {
            double v = red_v[0];
            int i = red_i[0];
            for (int w = 1; w < nwarps; ++w)
                if (red_v[w] > v || (red_v[w] == v && red_i[w] < i)) {
                    v = red_v[w];
                    i = red_i[w];
                }
            piv[k] = i < n ? i : k;
        }
        __syncthreads();
        const int p = piv[k];
        if (p != k)
            for (int c = tid; c < n; c += nt) {
                const double t = a[(int64_t)k * n + c];
                a[(int64_t)k * n + c] = a[(int64_t)p * n + c];
                a[(int64_t)p * n + c] = t;
            }
        __syncthreads();
        const double pv = a[(int64_t)k * n + k];
        __syncthreads();
        const double ipv = 1.0 / pv;
        for (int c = tid; c < n; c += nt) a[(int64_t)k * n + c] = (c == k ? 1.0 : a[(int64_t)k * n + c]) * ipv;
        __syncthreads();
        // eliminate column k from every other row: thread = (row stripe, column stripe) over the whole matrix
        for (int e = tid; e < n * n; e += nt) {
            const int i = e / n, c = e - i * n;
            if (i == k) continue;
            const double f = a[(int64_t)i * n + k];
            if (c == k) continue;  // handled after the barrier (it is the multiplier the other columns still read)
            a[e] = fma(-f, a[(int64_t)k * n + c], a[e]);
        }
        __syncthreads();
        for (int i = tid; i < n; i += nt)
            if (i != k) a[(int64_t)i * n + k] = -a[(int64_t)i * n + k] * a[(int64_t)k * n + k];
        __syncthreads();
    }
    // undo the row interchanges as column interchanges, in reverse order
    for (int k = n - 1; k >= 0; --k) {
        const int p = piv[k];
        if (p != k)
            for (int r = tid; r < n; r += nt) {
                const double t = a[(int64_t)r * n + k];
                a[(int64_t)r * n + k] = a[(int64_t)r * n + p];
                a[(int64_t)r * n + p] = t;
            }
        __syncthreads();
    }
}

// inv_cov = inverse(cov) for each hyper-parameter sample (grid.x = nhyp); h = inv_cov * mean
__global__ void __launch_bounds__(EP_THREADS) ep_prepare_kernel(const double* __restrict__ cov,
                                                                const double* __restrict__ mean, int T,
                                                                double* __restrict__ inv_cov, double* __restrict__ h,
                                                                double* __restrict__ gscratch, int in_smem) {
    extern __shared__ double sm[];
    __shared__ int piv[2048];
    __shared__ double red_v[32];
    __shared__ int red_i[32];
    const int64_t b = blockIdx.x;
    double* a = in_smem ? sm : gscratch + b * T * T;
    for (int e = threadIdx.x; e < T * T; e += blockDim.x) a[e] = cov[b * T * T + e];
    __syncthreads();
    gj_inverse(a, T, piv, red_v, red_i);
    for (int e = threadIdx.x; e < T * T; e += blockDim.x) inv_cov[b * T * T + e] = a[e];
    for (int i = threadIdx.x; i < T; i += blockDim.x) {
        double s = 0.0;
        for (int k = 0; k < T; ++k) s = fma(a[(int64_t)i * T + k], mean[b * T + k], s);
        h[b * T + i] = s;
    }
}

__device__ __forceinline__ double norm_pdf(double x) { return exp(-0.5 * x * x) * 0.3989422804014327; }
__device__ __forceinline__ double norm_cdf(double x) { return 0.5 * erfc(-x * 0.7071067811865476); }

// site column c of maximizer j refers to the test point o(c) = c < j ? c : c + 1
__device__ __forceinline__ int ep_other(int c, int j) { return c < j ? c : c + 1; }

// M = inv_cov + C diag(1/site_var) C^T, rhs = h + C (site_mean / site_var); then cov = M^-1 (in place), mean = cov rhs
__device__ void ep_refresh(double* M, int T, int j, const double* __restrict__ inv_cov, const double* __restrict__ h,
                           const double* site_m, const double* site_v, double* rhs, double* out_mean, int* piv,
                           double* red_v, int* red_i, double* scal) {
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int e = tid; e < T * T; e += nt) M[e] = inv_cov[e];
    for (int i = tid; i < T; i += nt) rhs[i] = h[i];
    __syncthreads();
    // the reference forms csp = cs / sqrt(site_vars) and csp csp^T: a negative site variance gives NaN (ep.py:17-21)
    if (tid == 0) {
        double sw = 0.0, sr = 0.0;
        for (int c = 0; c < T - 1; ++c) {
            const double w = 1.0 / sqrt(site_v[c]);
            sw += w * w;
            sr += site_m[c] / site_v[c];
        }
        scal[0] = sw;
        scal[1] = sr;
    }
    for (int c = tid; c < T - 1; c += nt) {
        const int o = ep_other(c, j);
        const double w = 1.0 / sqrt(site_v[c]);
        const double ww = w * w;
        M[(int64_t)o * T + o] += ww;
        M[(int64_t)j * T + o] -= ww;
        M[(int64_t)o * T + j] -= ww;
        rhs[o] -= site_m[c] / site_v[c];
    }
    __syncthreads();
    if (tid == 0) {
        M[(int64_t)j * T + j] += scal[0];
        rhs[j] += scal[1];
    }
    __syncthreads();
    gj_inverse(M, T, piv, red_v, red_i);
    for (int i = tid; i < T; i += nt) {
        double s = 0.0;
        for (int k = 0; k < T; ++k) s = fma(M[(int64_t)i * T + k], rhs[k], s);
        out_mean[i] = s;
    }
    __syncthreads();
}

// grid = (T maximizers, nhyp).  Outputs post_mean (nhyp, T, T), post_cov (nhyp, T, T, T), niter (nhyp, T).
__global__ void __launch_bounds__(EP_THREADS) ep_kernel(const double* __restrict__ mean_all,
                                                        const double* __restrict__ cov_all,
                                                        const double* __restrict__ inv_cov_all,
                                                        const double* __restrict__ h_all, int T, double resolution,
                                                        int max_niter, double* __restrict__ post_mean,
                                                        double* __restrict__ post_cov, int* __restrict__ niter,
                                                        double* __restrict__ gscratch, int in_smem) {
    extern __shared__ double sm[];
    __shared__ int piv[2048];
    __shared__ double red_v[32];
    __shared__ int red_i[32];
    __shared__ double scal[4];
    __shared__ int s_uidx, s_count, s_stop;
    const int j = blockIdx.x;
    const int64_t hy = blockIdx.y;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    const double* mu = mean_all + hy * T;
    const double* Sg = cov_all + hy * T * T;
    const double* inv_cov = inv_cov_all + hy * T * T;
    const double* h = h_all + hy * T;
    double* a_cov = post_cov + (hy * T + j) * (int64_t)T * T;   // current approximation (state + output)
    double* a_mean_out = post_mean + (hy * T + j) * (int64_t)T;
    // shared layout: [M (T*T) if in_smem] a_mean[T] n_mean[T] rhs[T] site_m[T] site_v[T] new_m[T] new_v[T]
    double* M = in_smem ? sm : gscratch + (hy * gridDim.x + j) * (int64_t)T * T;
    double* vec = in_smem ? sm + (int64_t)T * T : sm;
    double *a_mean = vec, *n_mean = vec + T, *rhs = vec + 2 * T, *site_m = vec + 3 * T, *site_v = vec + 4 * T,
           *new_m = vec + 5 * T, *new_v = vec + 6 * T;

    // site initialisation (ep.py:107-108): c^T mu and diag(c^T Sigma c)
    for (int c = tid; c < T - 1; c += nt) {
        const int o = ep_other(c, j);
        site_m[c] = mu[j] - mu[o];
        site_v[c] = (Sg[(int64_t)j * T + j] - Sg[(int64_t)o * T + j]) - (Sg[(int64_t)j * T + o] - Sg[(int64_t)o * T + o]);
    }
    __syncthreads();
    ep_refresh(M, T, j, inv_cov, h, site_m, site_v, rhs, a_mean, piv, red_v, red_i, scal);
    for (int e = tid; e < T * T; e += nt) a_cov[e] = M[e];
    if (tid == 0) {
        s_count = 0;
        s_stop = 0;
    }
    __syncthreads();
    double diff = 1e10;
    while (diff > resolution && s_count < max_niter) {
        __syncthreads();
        // cavity (ep.py:31-49), truncated-normal moments (:52-60), site update (:63-70), all sites at once
        for (int c = tid; c < T - 1; c += nt) {
            const int o = ep_other(c, j);
            const double ctSc = (a_cov[(int64_t)j * T + j] - a_cov[(int64_t)o * T + j]) -
                                (a_cov[(int64_t)j * T + o] - a_cov[(int64_t)o * T + o]);
            const double ctm = a_mean[j] - a_mean[o];
            const double cav_v = 1.0 / (1.0 / ctSc - 1.0 / site_v[c]);
            const double cav_m = cav_v * (ctm / ctSc - site_m[c] / site_v[c]);
            const double sd = sqrt(cav_v);
            const double beta = cav_m / sd;
            const double poc = norm_pdf(beta) / norm_cdf(beta);
            const double cf_m = cav_m + sd * poc;
            const double cf_v = -cav_m * sd * poc - cf_m * cf_m + 2.0 * cav_m * cf_m - cav_m * cav_m + cav_v;
            const double nv = 1.0 / (1.0 / cf_v - 1.0 / cav_v);
            new_v[c] = nv;
            new_m[c] = nv * (cf_m / cf_v - cav_m / cav_v);
        }
        __syncthreads();
        if (tid == 0) {
            int count = s_count + 1;
            int uidx = count % (T - 1);
            const int last = count >= 1 ? count - 1 : T - 2;
            int guard = 0;
            auto bad = [&](int u) {
                const double v = new_v[u];
                return v <= 0.0 || v != v || isinf(v);
            };
            // ep.py:162-172.  The reference's loop has no exit when every site is invalid and count-1 >= T-1; a
            // device kernel must terminate, so after one full cycle without a valid site the EP stops and returns
            // the current approximation.
            while (bad(uidx) && uidx != last) {
                ++count;
                uidx = count % (T - 1);
                if (++guard >= T - 1 && last >= T - 1) {
                    s_stop = 1;
                    break;
                }
            }
            s_count = count;
            s_uidx = uidx;
            if (!s_stop) {
                site_m[uidx] = new_m[uidx];
                site_v[uidx] = new_v[uidx];
            }
        }
        __syncthreads();
        if (s_stop) break;
        ep_refresh(M, T, j, inv_cov, h, site_m, site_v, rhs, n_mean, piv, red_v, red_i, scal);
        // converge_diff = mean over the (T, T) broadcast of dmean_i^2 + dcov_ik^2 (ep.py:190-193); NaN check
        double part = 0.0;
        int nan = 0;
        for (int e = tid; e < T * T; e += nt) {
            const int i = e / T;
            const double dm = n_mean[i] - a_mean[i];
            const double dc = M[e] - a_cov[e];
            part += dm * dm + dc * dc;
            nan |= (M[e] != M[e]) || (n_mean[i] != n_mean[i]);
        }
        part = warp_sum(part);
        nan = __any_sync(0xffffffffu, nan);
        if (lane == 0) {
            red_v[warp] = part;
            red_i[warp] = nan;
        }
        __syncthreads();
        double tot = 0.0;
        int anynan = 0;
        for (int w = 0; w < nwarps; ++w) {
            tot += red_v[w];
            anynan |= red_i[w];
        }
        diff = tot / ((double)T * (double)T);
        __syncthreads();
        if (anynan) break;  // ep.py:195-199: keep the previous approximation
        for (int e = tid; e < T * T; e += nt) a_cov[e] = M[e];
        for (int i = tid; i < T; i += nt) a_mean[i] = n_mean[i];
        __syncthreads();
    }
    __syncthreads();
    for (int i = tid; i < T; i += nt) a_mean_out[i] = a_mean[i];
    if (tid == 0 && niter) niter[hy * T + j] = s_count;
}

// ---- T <= 32: one WARP per maximizer, the system matrix in registers ------------------------------------------------
// The CTA-per-maximizer kernel above spends its time in block barriers: ~8 per Gauss-Jordan step, T steps per inverse,
// one inverse per EP iteration (the reference recomputes the full T x T inverse after every single-site update,
// ep.py:182-185) -- 79 us per iteration at T = 30, 7.9 ms for the 100 iterations of BASELINE config C2.  Here lane c
// holds COLUMN c of the matrix (col[i], i = row; the step and row loops are fully unrolled, so every register index
// is static): the pivot search runs inside lane k, rows are exchanged by predicated moves, the multipliers a[i][k]
// come from lane k by shuffle -- no barrier at all, only warp-synchronous code.  Pivot rule, operation order and
// rounding of every element are those of gj_inverse / ep_kernel, so the outputs are the same bits.
constexpr int EPW = 32;
constexpr int EPW_LD = EPW + 1;
constexpr int EPW_DOUBLES = 2 * EPW * EPW_LD + 8 * EPW;      // per warp: M, a_cov, 8 vectors

__device__ __forceinline__ void gj_inverse_warp(double (&col)[EPW], int n, int lane, double* __restrict__ Mw) {
    int mypiv = lane;                 // lane k keeps the pivot row of step k
    const double INF = pos_inf();
    // the step loop is NOT unrolled (the body is ~500 instructions: 32 unrolled steps overflow the instruction cache
    // and ran 15x slower); row k / row p of the lane's column are reached through predicated moves over the
    // statically indexed registers
#pragma unroll 1
    for (int k = 0; k < n; ++k) {
        // pivot search in column k, rows k..n-1: first maximum of |a_ik|, a NaN candidate counts as +inf
        double bv = -1.0, rk = 0.0;
        int bi = k;
#pragma unroll
        for (int i = 0; i < EPW; ++i) {
            if (i == k) rk = col[i];
            if (i >= k && i < n) {
                const double v = fabs(col[i]);
                const double key = v == v ? v : INF;
                if (key > bv) {
                    bv = key;
                    bi = i;
                }
            }
        }
        const int p = __shfl_sync(0xffffffffu, bi, k);
        if (lane == k) mypiv = p;
        if (p != k) {                 // exchange rows k and p of this column
            double rp = 0.0;
#pragma unroll
            for (int i = 0; i < EPW; ++i)
                if (i == p) {
                    rp = col[i];
                    col[i] = rk;
                }
            rk = rp;
        }
        const double pv = __shfl_sync(0xffffffffu, rk, k);      // a[k][k] lives in lane k
        const double ipv = 1.0 / pv;
        const double rkn = (lane == k ? 1.0 : rk) * ipv;        // the normalised row k
#pragma unroll
        for (int i = 0; i < EPW; ++i) {
            const double f = __shfl_sync(0xffffffffu, col[i], k);              // a[i][k] before this step
            if (i == k) col[i] = rkn;
            else if (i < n) col[i] = lane == k ? -f * rkn : fma(-f, rkn, col[i]);
        }
    }
    // undo the row interchanges as column interchanges, in reverse order: track where this lane's column ends up
    int pos = lane;
    for (int k = n - 1; k >= 0; --k) {
        const int p = __shfl_sync(0xffffffffu, mypiv, k);
        if (p != k) pos = pos == k ? p : (pos == p ? k : pos);
    }
    if (lane < n) {
#pragma unroll
        for (int i = 0; i < EPW; ++i)
            if (i < n) Mw[i * EPW_LD + pos] = col[i];
    }
    __syncwarp();
}

// M = inv_cov + C diag(1/site_var) C^T (column `lane` built in registers), rhs = h + C (site_mean / site_var);
// Mw = M^-1, out_mean = Mw rhs -- ep_refresh for one warp
__device__ __forceinline__ void ep_refresh_warp(double* __restrict__ Mw, int T, int j, const double* __restrict__ inv_cov,
                                                const double* __restrict__ h, const double* site_m, const double* site_v,
                                                double* rhs, double* wwv, double* out_mean, int lane) {
    if (lane < T - 1) {
        const double w = 1.0 / sqrt(site_v[lane]);     // the reference forms cs / sqrt(site_vars): negative -> NaN
        wwv[lane] = w * w;
    }
    if (lane < T) rhs[lane] = h[lane];
    __syncwarp();
    double sw = 0.0, sr = 0.0;
    if (lane == 0)
        for (int c = 0; c < T - 1; ++c) {
            const double w = 1.0 / sqrt(site_v[c]);
            sw += w * w;
            sr += site_m[c] / site_v[c];
        }
    sw = __shfl_sync(0xffffffffu, sw, 0);
    sr = __shfl_sync(0xffffffffu, sr, 0);
    if (lane < T - 1) rhs[ep_other(lane, j)] -= site_m[lane] / site_v[lane];
    __syncwarp();
    if (lane == 0) rhs[j] += sr;
    double col[EPW];
#pragma unroll
    for (int i = 0; i < EPW; ++i) col[i] = (i < T && lane < T) ? inv_cov[i * T + lane] : 0.0;
    if (lane < T) {
        if (lane == j) {
#pragma unroll
            for (int i = 0; i < EPW; ++i)
                if (i < T) {
                    if (i == j) col[i] += sw;
                    else col[i] -= wwv[i < j ? i : i - 1];
                }
        } else {
            const double ww = wwv[lane < j ? lane : lane - 1];
#pragma unroll
            for (int i = 0; i < EPW; ++i) {
                if (i == lane) col[i] += ww;
                if (i == j) col[i] -= ww;
            }
        }
    }
    __syncwarp();
    gj_inverse_warp(col, T, lane, Mw);
    if (lane < T) {
        double s = 0.0;
        for (int k = 0; k < T; ++k) s = fma(Mw[lane * EPW_LD + k], rhs[k], s);
        out_mean[lane] = s;
    }
    __syncwarp();
}

__global__ void __launch_bounds__(128) ep_warp_kernel(const double* __restrict__ mean_all,
                                                      const double* __restrict__ cov_all,
                                                      const double* __restrict__ inv_cov_all,
                                                      const double* __restrict__ h_all, int T, int64_t nhyp,
                                                      double resolution, int max_niter, double* __restrict__ post_mean,
                                                      double* __restrict__ post_cov, int* __restrict__ niter) {
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t gid = (int64_t)blockIdx.x * (blockDim.x >> 5) + wid;
    if (gid >= (int64_t)T * nhyp) return;          // whole warps leave: no block-wide barrier below
    const int j = (int)(gid % T);
    const int64_t hy = gid / T;
    const double* mu = mean_all + hy * T;
    const double* Sg = cov_all + hy * T * T;
    const double* inv_cov = inv_cov_all + hy * T * T;
    const double* h = h_all + hy * T;
    double* Mw = sm + (size_t)wid * EPW_DOUBLES;
    double* Cw = Mw + EPW * EPW_LD;                // current approximation (a_cov)
    double* vec = Cw + EPW * EPW_LD;
    double *a_mean = vec, *n_mean = vec + EPW, *rhs = vec + 2 * EPW, *site_m = vec + 3 * EPW, *site_v = vec + 4 * EPW,
           *new_m = vec + 5 * EPW, *new_v = vec + 6 * EPW, *wwv = vec + 7 * EPW;

    if (lane < T - 1) {      // site initialisation (ep.py:107-108)
        const int o = ep_other(lane, j);
        site_m[lane] = mu[j] - mu[o];
        site_v[lane] = (Sg[(int64_t)j * T + j] - Sg[(int64_t)o * T + j]) - (Sg[(int64_t)j * T + o] - Sg[(int64_t)o * T + o]);
    }
    __syncwarp();
    ep_refresh_warp(Mw, T, j, inv_cov, h, site_m, site_v, rhs, wwv, a_mean, lane);
    for (int e = lane; e < T * T; e += 32) Cw[(e / T) * EPW_LD + e % T] = Mw[(e / T) * EPW_LD + e % T];
    __syncwarp();
    int count = 0, stop = 0;
    double diff = 1e10;
    while (diff > resolution && count < max_niter) {
        if (lane < T - 1) {      // cavity (ep.py:31-49), truncated-normal moments (:52-60), site proposal (:63-70)
            const int c = lane, o = ep_other(c, j);
            const double ctSc = (Cw[j * EPW_LD + j] - Cw[o * EPW_LD + j]) - (Cw[j * EPW_LD + o] - Cw[o * EPW_LD + o]);
            const double ctm = a_mean[j] - a_mean[o];
            const double cav_v = 1.0 / (1.0 / ctSc - 1.0 / site_v[c]);
            const double cav_m = cav_v * (ctm / ctSc - site_m[c] / site_v[c]);
            const double sd = sqrt(cav_v);
            const double beta = cav_m / sd;
            const double poc = norm_pdf(beta) / norm_cdf(beta);
            const double cf_m = cav_m + sd * poc;
            const double cf_v = -cav_m * sd * poc - cf_m * cf_m + 2.0 * cav_m * cf_m - cav_m * cav_m + cav_v;
            const double nv = 1.0 / (1.0 / cf_v - 1.0 / cav_v);
            new_v[c] = nv;
            new_m[c] = nv * (cf_m / cf_v - cav_m / cav_v);
        }
        __syncwarp();
        if (lane == 0) {         // ep.py:162-172 (see ep_kernel for the termination guard)
            int cnt = count + 1;
            int uidx = cnt % (T - 1);
            const int last = cnt >= 1 ? cnt - 1 : T - 2;
            int guard = 0;
            auto bad = [&](int u) {
                const double v = new_v[u];
                return v <= 0.0 || v != v || isinf(v);
            };
            while (bad(uidx) && uidx != last) {
                ++cnt;
                uidx = cnt % (T - 1);
                if (++guard >= T - 1 && last >= T - 1) {
                    stop = 1;
                    break;
                }
            }
            count = cnt;
            if (!stop) {
                site_m[uidx] = new_m[uidx];
                site_v[uidx] = new_v[uidx];
            }
        }
        count = __shfl_sync(0xffffffffu, count, 0);
        stop = __shfl_sync(0xffffffffu, stop, 0);
        __syncwarp();
        if (stop) break;
        ep_refresh_warp(Mw, T, j, inv_cov, h, site_m, site_v, rhs, wwv, n_mean, lane);
        // converge_diff = mean over the (T, T) broadcast of dmean_i^2 + dcov_ik^2 (ep.py:190-193); NaN check
        double part = 0.0;
        int nan = 0;
        for (int e = lane; e < T * T; e += 32) {
            const int i = e / T, k = e - i * T;
            const double dm = n_mean[i] - a_mean[i];
            const double mv = Mw[i * EPW_LD + k];
            const double dc = mv - Cw[i * EPW_LD + k];
            part += dm * dm + dc * dc;
            nan |= (mv != mv) || (n_mean[i] != n_mean[i]);
        }
        part = warp_sum(part);
        nan = __any_sync(0xffffffffu, nan);
        diff = part / ((double)T * (double)T);
        if (nan) break;          // ep.py:195-199: keep the previous approximation
        for (int e = lane; e < T * T; e += 32) Cw[(e / T) * EPW_LD + e % T] = Mw[(e / T) * EPW_LD + e % T];
        if (lane < T) a_mean[lane] = n_mean[lane];
        __syncwarp();
    }
    __syncwarp();
    double* a_cov_out = post_cov + (hy * T + j) * (int64_t)T * T;
    for (int e = lane; e < T * T; e += 32) a_cov_out[e] = Cw[(e / T) * EPW_LD + e % T];
    if (lane < T) post_mean[(hy * T + j) * (int64_t)T + lane] = a_mean[lane];
    if (lane == 0 && niter) niter[hy * T + j] = count;
}


// ---- T <= 32: FOUR warps per maximizer, ping-pong matrices in shared memory -----------------------------------------
// The one-warp kernel above is a single dependent instruction stream: ~500 instructions per Gauss-Jordan step (64-bit
// shuffles for the multipliers, predicated register selects for rows k / p), 47 us per EP iteration at T = 30.  Here
// a CTA of 128 threads owns one maximizer: lane = column, warp w owns the rows w, w + 4, ...; a step READS the matrix
// from one shared buffer and WRITES every row of the next one (row exchange folded into the read), so ONE block barrier
// per step suffices.  Every warp finds the pivot redundantly (lane i looks at a[i][k]; two REDUX.MAX on the halves of
// |a| as an integer key + a ballot = first maximum, NaN counted as +inf: the rule of gj_inverse), every lane divides
// 1 / a[lane][k] while the reduction is in flight and the winner's quotient is broadcast.  Per-element arithmetic is
// that of gj_inverse.  The constant Sigma^-1 and h = Sigma^-1 mu are computed by the same routine inside the kernel
// (no prepare launch); 1/site_var terms are cached per site (one site changes per iteration).
constexpr int EPQ_LD = 33;
constexpr int EPQ_MAT = 32 * EPQ_LD;

// inverse of the n x n matrix in cur (n <= 32); returns the buffer (cur or nxt) that holds it.  All 128 threads call.
// The step is branch-free: every thread updates its 8 rows of the full 32 x 32 array (the padding rows / columns
// beyond n are inert: they never enter the pivot search and nothing in the n x n block reads them), row k is written
// afterwards by its owner, the column-k entries -f * rkn come out of the same fma with a zero addend.
__device__ __forceinline__ double* gj_inverse_quad(double* cur, double* nxt, int n, int lane, int w) {
    int Q = lane;                 // where column `lane` of the working matrix ends up (row interchanges undone as
                                  // column interchanges, composed forward: Q_k = t_0 o t_1 o ... o t_k)
    const int lrow = lane * EPQ_LD, wrow = w * EPQ_LD;
#pragma unroll 1
    for (int k = 0; k < n; ++k) {
        const double ck = cur[lrow + k];                                   // a[lane][k]
        const bool inr = lane < n && lane >= k;
        const double av = fabs(ck);
        const unsigned long long key = av == av ? (unsigned long long)__double_as_longlong(av) : 0x7ff0000000000000ULL;
        const unsigned hi = inr ? (unsigned)(key >> 32) : 0u, lo = inr ? (unsigned)key : 0u;
        const double ipc = 1.0 / ck;
        const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
        const unsigned ml = __reduce_max_sync(0xffffffffu, (inr && hi == mh) ? lo : 0u);
        const unsigned bal = __ballot_sync(0xffffffffu, inr && hi == mh && lo == ml);
        const int p = __ffs(bal) - 1;                                      // pivot row (>= k)
        const int koff = k * EPQ_LD, poff = p * EPQ_LD;
        const double rp = cur[poff + lane];
        const double ipv = __shfl_sync(0xffffffffu, ipc, p);
        const double rkn = (lane == k ? 1.0 : rp) * ipv;                   // the normalised row k
        const double nrkn = -rkn;
        const int qk = __shfl_sync(0xffffffffu, Q, k), qp = __shfl_sync(0xffffffffu, Q, p);
        Q = lane == k ? qp : (lane == p ? qk : Q);
        double f[8], v[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int ioff = wrow + 4 * r * EPQ_LD;
            const int soff = ioff == poff ? koff : ioff;                   // rows k and p exchanged
            f[r] = cur[soff + k];
            v[r] = cur[soff + lane];
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) nxt[wrow + 4 * r * EPQ_LD + lane] = fma(f[r], nrkn, lane == k ? 0.0 : v[r]);
        if ((k & 3) == w) nxt[koff + lane] = rkn;
        __syncthreads();
        double* t = cur;
        cur = nxt;
        nxt = t;
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) nxt[wrow + 4 * r * EPQ_LD + Q] = cur[wrow + 4 * r * EPQ_LD + lane];
    __syncthreads();
    return nxt;
}

__global__ void __launch_bounds__(128) ep_quad_kernel(const double* __restrict__ mean_all,
                                                      const double* __restrict__ cov_all, int T, double resolution,
                                                      int max_niter, double* __restrict__ post_mean,
                                                      double* __restrict__ post_cov, int* __restrict__ niter) {
    __shared__ double buf[4][EPQ_MAT];
    __shared__ double a_mean[32], n_mean[32], rhs[32], hv[32], site_m[32], site_v[32], new_m[32], new_v[32], wwv[32],
        qv[32], red[4], scal[2];
    __shared__ int ctl[2], nanw[4];        // count, stop; NaN flag of each warp
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int j = (int)(blockIdx.x % (unsigned)T);
    const int64_t hy = blockIdx.x / (unsigned)T;
    const double* mu = mean_all + hy * T;
    const double* Sg = cov_all + hy * T * T;

    // Sigma^-1 and h = Sigma^-1 mu (the padding of all four buffers is zero and stays zero)
    for (int e = tid; e < 4 * EPQ_MAT; e += 128) (&buf[0][0])[e] = 0.0;
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int i = w + 4 * r;
        if (i < T && lane < T) buf[0][i * EPQ_LD + lane] = Sg[(int64_t)i * T + lane];
    }
    if (tid < T - 1) {            // site initialisation (ep.py:107-108): c^T mu and diag(c^T Sigma c)
        const int o = ep_other(tid, j);
        const double sm = mu[j] - mu[o];
        const double sv = (Sg[(int64_t)j * T + j] - Sg[(int64_t)o * T + j]) - (Sg[(int64_t)j * T + o] - Sg[(int64_t)o * T + o]);
        site_m[tid] = sm;
        site_v[tid] = sv;
        const double iw = 1.0 / sqrt(sv);          // the reference forms cs / sqrt(site_vars): negative -> NaN
        wwv[tid] = iw * iw;
        qv[tid] = sm / sv;
    }
    if (tid == 0) ctl[0] = ctl[1] = 0;
    __syncthreads();
    double* IC = gj_inverse_quad(buf[0], buf[1], T, lane, w);
    double* f0 = IC == buf[0] ? buf[1] : buf[0];
    double* f1 = buf[2];
    double* f2 = buf[3];
    if (tid < T) {
        double s = 0.0;
        for (int k = 0; k < T; ++k) s = fma(IC[tid * EPQ_LD + k], mu[k], s);
        hv[tid] = s;
    }

    // system matrix M = Sigma^-1 + C diag(1/site_var) C^T into F, rhs = h + C (site_mean / site_var); then F^-1 and
    // the mean; returns the buffer with the inverse, *other = the second buffer used
    auto refresh = [&](double* F, double* G2, double* out_mean, double** other) -> double* {
        if (tid == 0) {
            double sw = 0.0;
            for (int c = 0; c < T - 1; ++c) sw += wwv[c];
            scal[0] = sw;
        } else if (tid == 32) {
            double sr = 0.0;
            for (int c = 0; c < T - 1; ++c) sr += qv[c];
            scal[1] = sr;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int i = w + 4 * r;
            if (i < T && lane < T) {
                double v = IC[i * EPQ_LD + lane];
                if (lane == j) {
                    if (i == j) v += scal[0];
                    else v -= wwv[i < j ? i : i - 1];
                } else {
                    const double ww = wwv[lane < j ? lane : lane - 1];
                    if (i == lane) v += ww;
                    if (i == j) v -= ww;
                }
                F[i * EPQ_LD + lane] = v;
            }
        }
        if (tid < T) rhs[tid] = tid == j ? hv[j] + scal[1] : hv[tid] - qv[tid < j ? tid : tid - 1];
        __syncthreads();
        double* R = gj_inverse_quad(F, G2, T, lane, w);
        *other = R == F ? G2 : F;
        if (tid < T) {
            double s = 0.0;
            for (int k = 0; k < T; ++k) s = fma(R[tid * EPQ_LD + k], rhs[k], s);
            out_mean[tid] = s;
        }
        __syncthreads();
        return R;
    };
    __syncthreads();
    double* other;
    double* Cw = refresh(f0, f1, a_mean, &other);     // current approximation
    f0 = other;                                        // free buffers: f0, f2
    f1 = f2;
    double diff = 1e10;
    int count = 0;
    while (diff > resolution && count < max_niter) {
        if (tid < T - 1) {       // cavity (ep.py:31-49), truncated-normal moments (:52-60), site proposal (:63-70)
            const int c = tid, o = ep_other(c, j);
            const double ctSc = (Cw[j * EPQ_LD + j] - Cw[o * EPQ_LD + j]) - (Cw[j * EPQ_LD + o] - Cw[o * EPQ_LD + o]);
            const double ctm = a_mean[j] - a_mean[o];
            const double cav_v = 1.0 / (1.0 / ctSc - 1.0 / site_v[c]);
            const double cav_m = cav_v * (ctm / ctSc - site_m[c] / site_v[c]);
            const double sd = sqrt(cav_v);
            const double beta = cav_m / sd;
            const double poc = norm_pdf(beta) / norm_cdf(beta);
            const double cf_m = cav_m + sd * poc;
            const double cf_v = -cav_m * sd * poc - cf_m * cf_m + 2.0 * cav_m * cf_m - cav_m * cav_m + cav_v;
            const double nv = 1.0 / (1.0 / cf_v - 1.0 / cav_v);
            new_v[c] = nv;
            new_m[c] = nv * (cf_m / cf_v - cav_m / cav_v);
        }
        __syncthreads();
        if (tid == 0) {          // ep.py:162-172 (see ep_kernel for the termination guard)
            int cnt = count + 1;
            int uidx = cnt % (T - 1);
            const int last = cnt >= 1 ? cnt - 1 : T - 2;
            int guard = 0, stop = 0;
            auto bad = [&](int u) {
                const double v = new_v[u];
                return v <= 0.0 || v != v || isinf(v);
            };
            while (bad(uidx) && uidx != last) {
                ++cnt;
                uidx = cnt % (T - 1);
                if (++guard >= T - 1 && last >= T - 1) {
                    stop = 1;
                    break;
                }
            }
            ctl[0] = cnt;
            ctl[1] = stop;
            if (!stop) {
                const double sm = new_m[uidx], sv = new_v[uidx];
                site_m[uidx] = sm;
                site_v[uidx] = sv;
                const double iw = 1.0 / sqrt(sv);
                wwv[uidx] = iw * iw;
                qv[uidx] = sm / sv;
            }
        }
        __syncthreads();
        count = ctl[0];
        if (ctl[1]) break;
        double* R = refresh(f0, f1, n_mean, &other);
        // converge_diff = mean over the (T, T) broadcast of dmean_i^2 + dcov_ik^2 (ep.py:190-193); NaN check
        double part = 0.0;
        int nan = 0;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int i = w + 4 * r;
            if (i < T && lane < T) {
                const double dm = n_mean[i] - a_mean[i];
                const double mv = R[i * EPQ_LD + lane];
                const double dc = mv - Cw[i * EPQ_LD + lane];
                part += dm * dm + dc * dc;
                nan |= (mv != mv) || (n_mean[i] != n_mean[i]);
            }
        }
        part = warp_sum(part);
        nan = __any_sync(0xffffffffu, nan);
        if (lane == 0) {
            red[w] = part;
            nanw[w] = nan;
        }
        __syncthreads();
        diff = (((red[0] + red[1]) + red[2]) + red[3]) / ((double)T * (double)T);
        const int anynan = nanw[0] | nanw[1] | nanw[2] | nanw[3];
        __syncthreads();
        if (anynan) break;       // ep.py:195-199: keep the previous approximation
        if (tid < T) a_mean[tid] = n_mean[tid];
        f0 = other;              // the inverse becomes the approximation, the old approximation a free buffer
        f1 = Cw;
        Cw = R;
        __syncthreads();
    }
    __syncthreads();
    double* a_cov_out = post_cov + (hy * T + j) * (int64_t)T * T;
    for (int e = tid; e < T * T; e += 128) a_cov_out[e] = Cw[(e / T) * EPQ_LD + e % T];
    if (tid < T) post_mean[(hy * T + j) * (int64_t)T + tid] = a_mean[tid];
    if (tid == 0 && niter) niter[hy * T + j] = count;
}

}  // namespace tes

using namespace tes;

constexpr int EP_SMEM_MAX_T = 160;

namespace tes {
// out[b] = inverse(A[b]) by Gauss-Jordan with partial pivoting (the pivot rule of the LU behind np.linalg.inv /
// tf.linalg.inv), one CTA per matrix, in place in `out` (global memory; small matrices stay in L1/L2)
__global__ void __launch_bounds__(EP_THREADS) gj_inverse_kernel(const double* __restrict__ A, int n,
                                                                double* __restrict__ out) {
    extern __shared__ int gj_piv[];
    __shared__ double red_v[32];
    __shared__ int red_i[32];
    const int64_t b = blockIdx.x;
    double* a = out + b * n * n;
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) a[e] = A[b * n * n + e];
    __syncthreads();
    gj_inverse(a, n, gj_piv, red_v, red_i);
}
}  // namespace tes

extern "C" int tes_batched_gj_inverse_f64(const double* A, int64_t batch, int64_t n, double* out, void* stream) {
    if (batch < 0) return -2;
    if (n < 1 || n > 8192) return -3;
    if (batch == 0) return 0;
    if (!A) return -1;
    if (!out) return -4;
    tes::gj_inverse_kernel<<<(unsigned)batch, tes::EP_THREADS, (size_t)n * sizeof(int), (cudaStream_t)stream>>>(
        A, (int)n, out);
    TES_LAUNCH_OK();
    return 0;
}

extern "C" int64_t tes_ep_moments_workspace_bytes(int64_t nhyp, int64_t T) {
    if (nhyp < 1 || T < 2 || T > 2048) return -1;
    int64_t b = align_up(nhyp * T * T * 8, 256) + align_up(nhyp * T * 8, 256);
    if (T > EP_SMEM_MAX_T) b += align_up(nhyp * T * T * T * 8, 256);
    return b;
}

// Batched EP: mean (nhyp, T), cov (nhyp, T, T) -> post_mean (nhyp, T, T), post_cov (nhyp, T, T, T): row j of the
// outputs is the EP approximation of N(mean, cov) conditioned on test point j being the maximizer.
extern "C" int tes_ep_moments_f64(const double* mean, const double* cov, int64_t nhyp, int64_t T, double resolution,
                                  int max_niter, double* post_mean, double* post_cov, int* niter, void* ws,
                                  int64_t ws_bytes, void* stream) {
    if (!mean) return -1;
    if (!cov) return -2;
    if (nhyp < 1 || nhyp > 65535) return -3;
    if (T < 2 || T > 2048) return -4;
    if (max_niter < 0) return -6;
    if (!post_mean) return -7;
    if (!post_cov) return -8;
    const int64_t need = tes_ep_moments_workspace_bytes(nhyp, T);
    if (!ws || ws_bytes < need) return -100;
    cudaStream_t st = (cudaStream_t)stream;
    char* base = (char*)ws;
    double* inv_cov = (double*)base;
    base += align_up(nhyp * T * T * 8, 256);
    double* h = (double*)base;
    base += align_up(nhyp * T * 8, 256);
    double* gscratch = (double*)base;
    const int in_smem = T <= EP_SMEM_MAX_T;
    const size_t sm_prep = in_smem ? (size_t)T * T * 8 : 0;
    const size_t sm_main = (in_smem ? (size_t)T * T * 8 : 0) + (size_t)7 * T * 8;
    if (sm_prep > 48 * 1024)
        TES_CUDA_OK(cudaFuncSetAttribute(ep_prepare_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_prep));
    if (sm_main > 48 * 1024)
        TES_CUDA_OK(cudaFuncSetAttribute(ep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_main));
    // developer knob: 0 = four-warp kernel for T <= 32, 3 = one-warp kernel, 1 / 2 = CTA kernel with 512 / 128 threads
    const char* ep_env = getenv("TES_EP_MODE");
    const int ep_mode = ep_env ? atoi(ep_env) : 0;
    if (T <= 32 && ep_mode == 0) {
        ep_quad_kernel<<<(unsigned)(T * nhyp), 128, 0, st>>>(mean, cov, (int)T, resolution, max_niter, post_mean, post_cov,
                                                            niter);
        TES_LAUNCH_OK();
        return 0;
    }
    // the prepare step reuses the first nhyp slots of the main kernel's scratch
    ep_prepare_kernel<<<(unsigned)nhyp, EP_THREADS, sm_prep, st>>>(cov, mean, (int)T, inv_cov, h, gscratch, in_smem);
    TES_LAUNCH_OK();
    if (ep_mode == 2) {
        ep_kernel<<<dim3((unsigned)T, (unsigned)nhyp), 128, sm_main, st>>>(mean, cov, inv_cov, h, (int)T, resolution, max_niter,
                                                                          post_mean, post_cov, niter, gscratch, in_smem);
        TES_LAUNCH_OK();
        return 0;
    }
    if (T <= EPW && ep_mode == 3) {
        // one warp per maximizer, the system matrix in registers (no block barriers); two warps per CTA so that the
        // maximizers spread over the SMs
        const int wpb = 2;
        const size_t smw = (size_t)wpb * EPW_DOUBLES * sizeof(double);
        const int64_t nwarp = T * nhyp;
        ep_warp_kernel<<<(unsigned)((nwarp + wpb - 1) / wpb), 32 * wpb, smw, st>>>(mean, cov, inv_cov, h, (int)T, nhyp,
                                                                                  resolution, max_niter, post_mean,
                                                                                  post_cov, niter);
        TES_LAUNCH_OK();
        return 0;
    }
    ep_kernel<<<dim3((unsigned)T, (unsigned)nhyp), EP_THREADS, sm_main, st>>>(mean, cov, inv_cov, h, (int)T, resolution,
                                                                             max_niter, post_mean, post_cov, niter,
                                                                             gscratch, in_smem);
    TES_LAUNCH_OK();
    return 0;
}
