// criteria.cu -- Stage D of the TES hot path: the Monte-Carlo entropy criteria.
//
//   tes_ep_acq_f64        TES_ep, q=1   criteria/evaluate_mp.py:56-295 (deterministic + stochastic),
//                                       criteria/evaluate_mp_lite.py:116-225 (closed-form "lite")
//   tes_sp_acq_f64        TES_sp, q>=1  criteria/evaluate_sample_mp.py:54-336 (q=1) and
//                                       criteria/evaluate_batch_sample_mp.py:54-382 (batch q)
//   tes_batch_ep_acq_f64  TES_ep batch  criteria/evaluate_batch_mp.py:145-331
//
// One hyper-parameter sample per call (the reference loops `for i in range(nhyp)` and averages).
// The caller passes only the maximizers with max_probs > 0 (evaluate_mp.py:169-173).
#include <cstdlib>

#include "gemm.cuh"
#include "gp.cuh"
#include "sp2.cuh"

namespace tes {

constexpr double HALF_LOG_2PI = 0.91893853320467274178;
// utils.evaluate_norm_entropy (utils.py:653-654) writes tf.cast(0.5 * tf.log(2*np.pi), dtype): TF turns the
// bare Python float into a float32 tensor, so the entropy constant is float32(0.5 * logf(6.2831855f)) widened to
// fp64 -- 4.4e-8 above 0.5*log(2*pi) (SURVEY quirk list, Q12; pinned by tests/golden/criteria.npz).  TFP's
// Normal.log_prob uses the fp64 constant (HALF_LOG_2PI).
constexpr double NORM_ENTROPY_CONST = 0.918938577175140380859375;

// ---- TES_ep query moments: var[x][j] = Kq[x] + M[x]^T Sigma_j M[x]  (evaluate_mp.py:103-107) ----
// The quadratic form is a GEMM over the index pairs (a <= b) of the symmetric Sigma_j:
//     var[x][j] - Kq[x] = sum_{a <= b} (M[x,a] M[x,b]) * Bpack[(a,b)][j],   Bpack = Sigma_j[a][b] (+ Sigma_j[b][a] if a < b)
// with the A operand (the products) generated on the fly.  The K axis is cut into 16-wide SLICES described by a table:
//   * off-diagonal pair of 16-blocks (pa < pb): 16 slices, slice i holds a = 16 pa + i, b = 16 pb + kk (kk = 0..15);
//   * diagonal pair (pa = pb): only the 136 pairs a <= b are enumerated, in 9 slices instead of 16: slice 0 is row 0
//     (16 entries), slice i = 1..7 packs row i (16 - i entries, b = i + kk) with row 16 - i (i entries, b = kk), slice 8
//     is row 8 (8 entries, the rest zero).  At T = 128 this is K = 8 * 144 + 28 * 256 = 8320 instead of 9216.
__global__ void quad_slice_table_kernel(int nb, QSlice* table) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int64_t n = 0;
        for (int pa = 0; pa < nb; ++pa)
            for (int pb = pa; pb < nb; ++pb) {
                const int ns = pa == pb ? 9 : 16;
                for (int i = 0; i < ns; ++i) table[n++] = QSlice{pa, pb, i, pa == pb};
            }
    }
}

// Bpack[krow][j] = Sigma_j[a][a] (a = b) or Sigma_j[a][b] + Sigma_j[b][a] (a < b)
__global__ void quad_pack_kernel(const double* __restrict__ cov, int64_t Sp, int64_t T, const QSlice* __restrict__ table,
                                 int64_t nrows, double* __restrict__ Bpack) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nrows * Sp) return;
    int64_t krow = e / Sp, j = e % Sp;
    const QSlice q = table[krow >> 4];
    int64_t a, b;
    double v = 0.0;
    if (quad_slice_ab(q, (int)(krow & 15), a, b) && a < T && b < T) {
        const double* c = cov + j * T * T;
        v = c[a * T + b];
        if (a != b) v += c[b * T + a];
    }
    Bpack[e] = v;
}

struct AQuad {
    const double* G;  // rows x ldg, the first T columns are M
    int64_t ldg, M, T;
    const QSlice* table;
    // load() only ISSUES the 16 global loads of the next slice (raw M[x,a], M[x,b]); the products are formed in
    // store(), in the middle of the current slice's DMMA stream (gemm_mainloop) -- a product in load() would stall the
    // warp on the L2 latency before it can issue its tensor work.
    static constexpr int NREG = 16;
    __device__ __forceinline__ void load(int64_t row0, int64_t k0, int t, double regs[16]) const {
        const int kk = t & 15;
        const int r = t >> 4;
        const int4 qv = __ldg(reinterpret_cast<const int4*>(table) + (k0 >> 4));
        const QSlice q{qv.x, qv.y, qv.z, qv.w};
        int64_t a = 0, b = 0;
        const bool ok = quad_slice_ab(q, kk, a, b) && a < T && b < T;
#pragma unroll
        for (int qq = 0; qq < 8; ++qq) {
            int64_t row = row0 + r + 16 * qq;
            const bool in = ok && row < M;
            regs[qq] = in ? __ldg(G + row * ldg + a) : 0.0;
            regs[8 + qq] = in ? __ldg(G + row * ldg + b) : 0.0;
        }
    }
    __device__ __forceinline__ void store(double (*as)[GEMM_LDA_S], int t, const double regs[16]) const {
        const int kk = t & 15;
        const int r = t >> 4;
#pragma unroll
        for (int q = 0; q < 8; ++q) as[kk][r + 16 * q] = __dmul_rn(regs[q], regs[8 + q]);
    }
};

__global__ void __launch_bounds__(GEMM_THREADS, 2)
    quadform_kernel(const double* __restrict__ G, int64_t ldg, int64_t rows, int64_t T, const QSlice* __restrict__ table,
                    int64_t Kdim, const double* __restrict__ Bpack, int64_t Sp, const double* __restrict__ Kq,
                    double* __restrict__ var) {
    extern __shared__ __align__(16) unsigned char gemm_smem_raw[];
    GemmSmem& sm = *reinterpret_cast<GemmSmem*>(gemm_smem_raw);
    const int64_t row0 = (int64_t)blockIdx.x * GEMM_BM;
    const int64_t col0 = (int64_t)blockIdx.y * GEMM_BN;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    AQuad ap{G, ldg, rows, T, table};
    gemm_mainloop(ap, Bpack, Sp, Kdim, Sp, row0, col0, sm, acc);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int64_t row = row0 + gemm_row_of(warp, lane, i);
        if (row >= rows) continue;
        const double kq = Kq[row];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                int64_t col = col0 + gemm_col_of(warp, lane, j, e);
                if (col < Sp) var[row * Sp + col] = kq + acc[i][j][e];
            }
    }
}

__global__ void add_col_kernel(double* __restrict__ A, int64_t rows, int64_t cols, const double* __restrict__ v,
                               int64_t ldv) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= rows * cols) return;
    A[e] += v[(e / cols) * ldv];
}

// ---- TES_ep deterministic (evaluate_mp.py:199-219) ----------------------------------------------
// acq = H[y] - mean_j( p_j * H[y | j] ),  H(s) = 0.5 log(2 pi) + log(s) + 0.5  (utils.py:653)
__global__ void ep_det_kernel(const double* __restrict__ var, const double* __restrict__ varf,
                              const double* __restrict__ p, int64_t rows, int64_t Sp, double sigma0,
                              double* __restrict__ out) {
    int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (row >= rows) return;
    double s = 0.0;
    for (int64_t j = lane; j < Sp; j += 32) {
        double sd = sqrt(var[row * Sp + j] + sigma0);  // no clipping: negative -> NaN
        s += ((NORM_ENTROPY_CONST + log(sd)) + 0.5) * p[j];
    }
    s = warp_sum(s);
    if (lane == 0) {
        double ent = (NORM_ENTROPY_CONST + log(sqrt(varf[row] + sigma0))) + 0.5;
        out[row] = ent - s / (double)Sp;  // reduce_MEAN over j
    }
}

// ---- TES_ep lite (evaluate_mp_lite.py:187-220) ----------------------------------------------------
// sum_{i,j} p_i p_j [ 0.5 log(v_i/v_j) + 0.5 (v_j + (m_j - m_i)^2)/v_i - 0.5 ],  v = var + sigma0
__global__ void ep_lite_kernel(const double* __restrict__ mean, const double* __restrict__ var,
                               const double* __restrict__ p, int64_t rows, int64_t Sp, double sigma0,
                               double* __restrict__ out) {
    int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const double* mr = mean + row * Sp;
    const double* vr = var + row * Sp;
    double s = 0.0;
    for (int64_t e = lane; e < Sp * Sp; e += 32) {
        int64_t i = e / Sp, j = e % Sp;
        double vi = vr[i] + sigma0, vj = vr[j] + sigma0;
        double diff = mr[j] - mr[i];
        double mips = 0.5 * log(vi / vj) + 0.5 * (vj + diff * diff) / vi - 0.5;
        s += p[i] * p[j] * mips;
    }
    s = warp_sum(s);
    if (lane == 0) out[row] = s;
}

// ---- TES_ep stochastic (evaluate_mp.py:248-295) -----------------------------------------------------
// per (it, iy, j): y = mu_j + s_j z;  cond term lp_j(y);  marginal LSE_j'( lp_j'(y) + log p_j' ).
// acq = (1/(I*Y)) sum_{it,iy,j} p_j * ( lp_j(y) - LSE_j' )          (= ent - cond, averaged)
// One CTA per candidate; z either host-supplied (layout (I, Y, Sp, Nz) with row stride Nz) or
// Philox (tes_spec.h, stream TES_STREAM_EP_Y, element (iy*Sp + j)*n_global + x_global).
// exp(x) for x in (-700, 700): x = (16 n + i) ln2/16 + r, e^x = 2^n 2^(i/16) P5(r) (16-entry table, degree-5 Taylor,
// truncation 1.4e-13 relative): 10 fp64-pipe instructions instead of ~25 for exp()
__device__ __forceinline__ double exp_fast16(double x, const double* __restrict__ tab) {
    const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52
    double t = fma(x, 0x1.71547652b82fep+4, MAGIC);
    const int k = __double2loint(t);          // rint(x * 16/ln2)
    const double kf = t - MAGIC;
    double r = fma(kf, -0x1.62e42fefa0000p-5, x);
    r = fma(kf, -0x1.cf79abc9e3b3ap-44, r);
    double p = fma(r, 1.0 / 120.0, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    double v = tab[k & 15] * p;
    v = __hiloint2double(__double2hiint(v) + ((k >> 4) << 20), __double2loint(v));
    return x > -700.0 ? v : (x != x ? x : 0.0);
}

// ONE pass per (draw, component j'): the log-sum-exp is shifted by the draw's OWN component (the j' = j term, which is
// then exp(0) = 1: the sum can neither underflow nor -- density ratios of the mixture's components are bounded by
// e^(z^2/2) s_j / s_j' p_j' / p_j << e^700 -- overflow), so no max search is needed; mathematically the reference's
// max-shifted tf.reduce_logsumexp, different by rounding only.  With w' = (y - mu_j') / (sqrt(2) s_j') a term is
// exp(cst_j' - w'^2 - shift): two FMAs, one subtraction and the table-driven exp -- 14 fp64-pipe instructions per
// (draw, component) pair instead of ~42 (two passes, libm exp).  Two draws per thread share the component loads.
__global__ void __launch_bounds__(128)
    ep_stoch_kernel(const double* __restrict__ mean, const double* __restrict__ var, const double* __restrict__ p,
                    int64_t rows, int Sp, double sigma0, int niter, int ny, const double* __restrict__ z,
                    int64_t z_n, int64_t z_x0, uint64_t seed, int64_t n_global, int64_t x_global0,
                    double* __restrict__ out) {
    // n_global < 0: SHARED draws -- every candidate consumes the same normals (common random numbers for the
    // finite-difference gradients of the criterion refinement, utils_for_continuous.py:62-87)
    const int64_t kg_n = n_global < 0 ? 1 : n_global, kg_x0 = n_global < 0 ? 0 : x_global0, kg_rmul = n_global < 0 ? 0 : 1;
    extern __shared__ double sh[];
    double* mu = sh;              // [Sp]
    double* sdv = mu + Sp;        // [Sp] std
    double* isd = sdv + Sp;       // [Sp] 1 / std
    double* own = isd + Sp;       // [Sp] -(0.5 log 2pi + log std)
    double3* cmp = reinterpret_cast<double3*>(own + Sp);   // [Sp] (1 / (sqrt2 std), mu / (sqrt2 std), cst)
    double* tab = reinterpret_cast<double*>(cmp + Sp);     // [16] 2^(i/16)
    double* red = tab + 16;                                 // [4]
    const int64_t row = blockIdx.x;
    if (row >= rows) return;
    for (int j = threadIdx.x; j < Sp; j += blockDim.x) {
        const double sd = sqrt(var[row * Sp + j] + sigma0);   // no clipping: negative -> NaN
        const double m = mean[row * Sp + j];
        const double id = 1.0 / sd;
        mu[j] = m;
        sdv[j] = sd;
        isd[j] = id;
        const double o = -(HALF_LOG_2PI + log(sd));
        own[j] = o;
        cmp[j] = make_double3(id * 0.70710678118654752440, m * id * 0.70710678118654752440, o + log(p[j]));
    }
    if (threadIdx.x < 16) tab[threadIdx.x] = exp2((double)threadIdx.x * 0.0625);
    __syncthreads();
    double acc = 0.0;
    const int64_t items = (int64_t)niter * ny * Sp;
    const int64_t half = (items + 1) / 2;                    // thread handles items e and e + half together
    for (int64_t e0 = threadIdx.x; e0 < half; e0 += blockDim.x) {
        double y[2], lpc[2], sh_[2], se[2], pj[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int64_t it_ = e0 + u * half;
            const bool live = it_ < items;
            const int64_t ic = live ? it_ : 0;
            const int j = (int)(ic % Sp);
            const int iy = (int)((ic / Sp) % ny);
            const int it = (int)(ic / ((int64_t)Sp * ny));
            double zz;
            if (z)
                zz = z[(((int64_t)it * ny + iy) * Sp + j) * z_n + z_x0 + row];
            else
                zz = normal_elem(seed, (uint32_t)it, TES_STREAM_EP_Y,
                                 (uint64_t)(((int64_t)iy * Sp + j) * kg_n + kg_x0 + row * kg_rmul));
            y[u] = fma(sdv[j], zz, mu[j]);
            const double uu = (y[u] - mu[j]) * isd[j];
            lpc[u] = fma(-0.5 * uu, uu, own[j]);           // log N(y; mu_j, s_j)   (evaluate_mp.py:268)
            const double3 c = cmp[j];
            const double w = fma(y[u], c.x, -c.y);
            sh_[u] = fma(-w, w, c.z);                      // the own component's term: the shift
            se[u] = 0.0;
            pj[u] = live ? p[j] : 0.0;
        }
        for (int jp = 0; jp < Sp; ++jp) {
            const double3 c = cmp[jp];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const double w = fma(y[u], c.x, -c.y);
                se[u] += exp_fast16(fma(-w, w, c.z) - sh_[u], tab);
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const double lse = log(se[u]) + sh_[u];
            if (pj[u] != 0.0) acc += pj[u] * (lpc[u] - lse);
        }
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < (blockDim.x >> 5); ++w) s += red[w];
        out[row] = s / ((double)niter * (double)ny);
    }
}

// ---- TES_sp (evaluate_sample_mp.py:250-336; batch: evaluate_batch_sample_mp.py:294-382) -------------
// One CTA per query (a query = Q inputs).  Shared memory: mean[j][s][q] for all maximizers j and
// joint samples s, then per warp a y buffer of K*Q doubles.
//   lmix[iy,j]     = LSE_{s valid for j}  logN(y_{j,s}; mean[j,s], L)
//   lmarg[iy,j]    = LSE_{j'} ( log p_j' + LSE_{s valid for SOURCE j} logN(y_{j,s}; mean[j',s], L) )   (Q3)
//   acq = (1/(I*Y)) sum p_j (lmix - lmarg)
template <int Q>
__global__ void __launch_bounds__(256)
    sp_kernel(const double* __restrict__ Mq,   // (rows*Q, T)   k invpNK[:, :T]
              const double* __restrict__ bq,   // (rows*Q)      k invpNK[:, T:] Y
              const double* __restrict__ Linv, // (rows, Q, Q)  inverse of chol(cov_y), lower
              const double* __restrict__ Lmat, // (rows, Q, Q)  chol(cov_y), lower
              const double* __restrict__ P,    // (Sp, T, K)
              const unsigned char* __restrict__ mask,  // (Sp, K)
              const double* __restrict__ p, int64_t rows, int Sp, int T, int K, int niter, int ny,
              const double* __restrict__ z, int64_t z_n, int64_t z_x0, uint64_t seed, int64_t n_global,
              int64_t x_global0, double* __restrict__ mean_g,  // optional global scratch (rows? no: gridDim.x) x Sp*K*Q
              int mean_in_smem, double* __restrict__ out) {
    const int64_t kg_n = n_global < 0 ? 1 : n_global, kg_x0 = n_global < 0 ? 0 : x_global0, kg_rmul = n_global < 0 ? 0 : 1;
    extern __shared__ double sh[];
    const int nwarps = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* red = sh;                       // [8]
    double* lq = red + 8;                   // [Q*Q] Linv, [Q*Q] L, [Q] b, pad to 2*Q*Q+Q (+1 logdet)
    double* ybuf = lq + 2 * Q * Q + Q + 1;  // [nwarps][K*Q]
    double* mean_s = mean_in_smem ? (ybuf + (size_t)nwarps * K * Q) : (mean_g + (size_t)blockIdx.x * Sp * K * Q);
    for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
        __syncthreads();
        if (threadIdx.x < Q * Q) {
            lq[threadIdx.x] = Linv[row * Q * Q + threadIdx.x];
            lq[Q * Q + threadIdx.x] = Lmat[row * Q * Q + threadIdx.x];
        }
        if (threadIdx.x < Q) lq[2 * Q * Q + threadIdx.x] = bq[row * Q + threadIdx.x];
        if (threadIdx.x == 0) {
            double ld = 0.0;
            for (int a = 0; a < Q; ++a) ld += log(Lmat[row * Q * Q + a * Q + a]);
            lq[2 * Q * Q + Q] = ld;
        }
        __syncthreads();
        // mean[j][s][q] = sum_t M[row*Q+q][t] * P[j][t][s] + b[q]
        for (int e = threadIdx.x; e < Sp * K; e += blockDim.x) {
            const int j = e / K, s = e % K;
            double a[Q];
#pragma unroll
            for (int q = 0; q < Q; ++q) a[q] = 0.0;
            const double* pj = P + ((size_t)j * T) * K + s;
            for (int t = 0; t < T; ++t) {
                const double pv = pj[(size_t)t * K];
#pragma unroll
                for (int q = 0; q < Q; ++q) a[q] = fma(__ldg(Mq + (row * Q + q) * T + t), pv, a[q]);
            }
#pragma unroll
            for (int q = 0; q < Q; ++q) mean_s[((size_t)j * K + s) * Q + q] = a[q] + lq[2 * Q * Q + q];
        }
        __syncthreads();
        const double logdet = lq[2 * Q * Q + Q];
        const double cnorm = -logdet - 0.5 * Q * 1.8378770664093454836;  // - sum log L_ii - q/2 log(2 pi)
        double acc = 0.0;
        double* yb = ybuf + (size_t)warp * K * Q;
        const int64_t items = (int64_t)niter * ny * Sp;
        for (int64_t it_ = warp; it_ < items; it_ += nwarps) {
            const int j = (int)(it_ % Sp);
            const int iy = (int)((it_ / Sp) % ny);
            const int it = (int)(it_ / ((int64_t)Sp * ny));
            const unsigned char* mj = mask + (size_t)j * K;
            // y_s = mean[j][s] + L z_s for every s (invalid s are never read)
            __syncwarp();
            for (int s = lane; s < K; s += 32) {
                double zz[Q];
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    if (z)
                        zz[q] = z[(((((int64_t)it * ny + iy) * Sp + j) * z_n + z_x0 + row) * K + s) * Q + q];
                    else
                        zz[q] = normal_elem(seed, (uint32_t)it, Q == 1 ? TES_STREAM_SP_Y : TES_STREAM_BSP_Y,
                                            (uint64_t)((((int64_t)iy * Sp + j) * kg_n + kg_x0 + row * kg_rmul) * K + s) * Q + q);
                }
#pragma unroll
                for (int a = 0; a < Q; ++a) {
                    double v = mean_s[((size_t)j * K + s) * Q + a];
#pragma unroll
                    for (int b2 = 0; b2 <= a; ++b2) v = fma(lq[Q * Q + a * Q + b2], zz[b2], v);
                    yb[s * Q + a] = v;
                }
            }
            __syncwarp();
            // running log-sum-exp over j' of (log p_j' + LSE_s ...), and the j'==j inner LSE for lmix
            double lmix = 0.0;
            double out_mx = neg_inf(), out_se = 0.0;  // online LSE over j' (Sp terms, cheap)
            bool out_nan = false;
            for (int jp = 0; jp < Sp; ++jp) {
                const double* mp_ = mean_s + (size_t)jp * K * Q;
                // pass 1: min of u^2 over valid s
                double umin = pos_inf();
                for (int s = lane; s < K; s += 32) {
                    if (!mj[s]) continue;
                    double u2 = 0.0;
                    double dlt[Q];
#pragma unroll
                    for (int q = 0; q < Q; ++q) dlt[q] = yb[s * Q + q] - mp_[s * Q + q];
#pragma unroll
                    for (int a = 0; a < Q; ++a) {
                        double w = 0.0;
#pragma unroll
                        for (int b2 = 0; b2 <= a; ++b2) w = fma(lq[a * Q + b2], dlt[b2], w);
                        u2 = fma(w, w, u2);
                    }
                    umin = u2 < umin ? u2 : umin;
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    double o = __shfl_xor_sync(0xffffffffu, umin, off);
                    umin = o < umin ? o : umin;
                }
                // shift = max log-prob if finite else 0 (tf.reduce_logsumexp)
                const double mxlp = fma(-0.5, umin, cnorm);
                const double shift = (mxlp > neg_inf() && mxlp < pos_inf()) ? mxlp : 0.0;
                double se = 0.0;
                for (int s = lane; s < K; s += 32) {
                    if (!mj[s]) continue;
                    double u2 = 0.0;
                    double dlt[Q];
#pragma unroll
                    for (int q = 0; q < Q; ++q) dlt[q] = yb[s * Q + q] - mp_[s * Q + q];
#pragma unroll
                    for (int a = 0; a < Q; ++a) {
                        double w = 0.0;
#pragma unroll
                        for (int b2 = 0; b2 <= a; ++b2) w = fma(lq[a * Q + b2], dlt[b2], w);
                        u2 = fma(w, w, u2);
                    }
                    se += exp(fma(-0.5, u2, cnorm) - shift);
                }
                se = warp_sum(se);
                const double lse_s = log(se) + shift;  // (Y, j, j') entry
                if (jp == j) lmix = lse_s;
                const double term = lse_s + log(p[jp]);
                // online LSE over j' with the same non-finite-max rule applied at the end
                if (term != term) {
                    out_nan = true;
                } else if (term > out_mx) {
                    out_se = out_se * exp(out_mx - term) + 1.0;  // out_mx == -inf: 0 * 0 + 1
                    out_mx = term;
                } else if (term > neg_inf()) {
                    out_se += exp(term - out_mx);
                }
            }
            double lmarg;
            if (out_nan)
                lmarg = __longlong_as_double(0x7ff8000000000000LL);
            else if (out_mx > neg_inf() && out_mx < pos_inf())
                lmarg = log(out_se) + out_mx;
            else
                lmarg = out_mx;  // all -inf (or +inf): propagate
            acc += p[j] * (lmix - lmarg);
        }
        // acc identical across lanes of a warp; reduce over warps
        if (lane == 0) red[warp] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
            for (int w = 0; w < nwarps; ++w) s += red[w];
            out[row] = s / ((double)niter * (double)ny);
        }
    }
}

// cov_y = k_x - k invpNK k^T + sigma0 I (q x q), its Cholesky factor and inverse, per query
// (evaluate_batch_sample_mp.py:152, :251, :306-310).
// One WARP per query: lanes stride the m = T + n columns (coalesced reads of the q rows of G and of Kc),
// each lane keeps the q(q+1)/2 partial dot products in registers, butterfly reductions leave the full q x q matrix in
// every lane, which then runs the (tiny, branch-uniform) Cholesky + triangular inverse redundantly; lane e writes
// entries e, e+32 of the outputs.  (A thread-per-query version read 2 q m doubles per thread with a row-sized stride:
// 143 ms at C4; this one is bandwidth-bound on the L2-resident chunk.)
template <int Q>
__global__ void __launch_bounds__(256)
    sp_cov_warp_kernel(const double* __restrict__ x, int d, const double* __restrict__ l, double sigma, double sigma0,
                       const double* __restrict__ G, int64_t ldg, const double* __restrict__ Kc, int64_t ldk, int64_t m,
                       int64_t rows, double* __restrict__ Lmat, double* __restrict__ Linv, int* __restrict__ info) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    double c[Q][Q];
#pragma unroll
    for (int a = 0; a < Q; ++a)
#pragma unroll
        for (int b2 = 0; b2 < Q; ++b2) c[a][b2] = 0.0;
    for (int64_t t = lane; t < m; t += 32) {
        double ga[Q], kb[Q];
#pragma unroll
        for (int a = 0; a < Q; ++a) {
            ga[a] = G[(row * Q + a) * ldg + t];
            kb[a] = Kc[(row * Q + a) * ldk + t];
        }
#pragma unroll
        for (int a = 0; a < Q; ++a)
#pragma unroll
            for (int b2 = 0; b2 <= a; ++b2) c[a][b2] = fma(ga[a], kb[b2], c[a][b2]);
    }
#pragma unroll
    for (int a = 0; a < Q; ++a)
#pragma unroll
        for (int b2 = 0; b2 <= a; ++b2) {
            const double s = warp_sum(c[a][b2]);
            double qa = 0.0, qb = 0.0, dot = 0.0;   // Kxx via the reference's expansion (utils.py:740-762)
            for (int k = 0; k < d; ++k) {
                const double sl = sqrt(l[k]);
                const double xa = x[(row * Q + a) * d + k] * sl, xb = x[(row * Q + b2) * d + k] * sl;
                qa = fma(xa, xa, qa);
                qb = fma(xb, xb, qb);
                dot = fma(xa, xb, dot);
            }
            const double kxx = sigma * exp(-0.5 * ((qa + qb) - 2.0 * dot));
            c[a][b2] = kxx - s + (a == b2 ? sigma0 : 0.0);
        }
    bool bad = false;
#pragma unroll
    for (int a = 0; a < Q; ++a) {
#pragma unroll
        for (int b2 = 0; b2 <= a; ++b2) {
            double s = c[a][b2];
#pragma unroll
            for (int k = 0; k < b2; ++k) s -= c[a][k] * c[b2][k];
            if (a == b2) {
                if (!(s > 0.0)) bad = true;
                c[a][a] = sqrt(s);
            } else {
                c[a][b2] = s / c[b2][b2];
            }
        }
    }
    double inv[Q][Q];
#pragma unroll
    for (int a = 0; a < Q; ++a)
#pragma unroll
        for (int b2 = 0; b2 < Q; ++b2) inv[a][b2] = 0.0;
#pragma unroll
    for (int col = 0; col < Q; ++col) {
        inv[col][col] = 1.0 / c[col][col];
#pragma unroll
        for (int a = col + 1; a < Q; ++a) {
            double s = 0.0;
#pragma unroll
            for (int k = col; k < a; ++k) s -= c[a][k] * inv[k][col];
            inv[a][col] = s / c[a][a];
        }
    }
#pragma unroll
    for (int a = 0; a < Q; ++a)
#pragma unroll
        for (int b2 = 0; b2 < Q; ++b2) {
            const int e = a * Q + b2;
            if ((e & 31) == lane) {
                Lmat[row * Q * Q + e] = b2 <= a ? c[a][b2] : 0.0;
                Linv[row * Q * Q + e] = inv[a][b2];
            }
        }
    if (bad && info && lane == 0) atomicAdd(info, 1);
}

__global__ void batch_ep_kernel(double* out, int64_t n, const double* __restrict__ p, int64_t Sp,
                                const double* __restrict__ x, int64_t qd) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    double s = 0.0;
    for (int64_t j = 0; j < Sp; ++j) s += p[j];
    // evaluate_batch_mp.py:302-327: ent - cond == -(sum p) * log(sum p) for every finite input.  A
    // non-finite coordinate of the query reaches the result through k_x_xto -> cov -> eigh -> y -> log_prob
    // (:88-141, :236-241, :296-302; inf gives inf - inf in the squared-distance expansion), so the reference
    // returns NaN for that query (the drivers retry on NaN, bo_batch.py:1080-1081).
    bool finite = true;
    for (int64_t c = 0; c < qd; ++c) finite = finite && isfinite(x[e * qd + c]);
    out[e] = finite ? -s * log(s) : __longlong_as_double(0x7ff8000000000000LL);
}

struct EpPlan {
    int64_t m, nb, nbp, Kdim, nc;
    int64_t off_xto, off_bext, off_kc, off_gc, off_table, off_bpack, off_mean, off_var, off_kq, off_gobs, off_varf,
        total;
};

// The screen's chunk: two 128-row tiles per SM -- its fp64 GEMMs fill whole waves (N = T + n + 1 = 190: 888 CTAs = 3
// waves of 2 per SM instead of 1.5) and the per-chunk launches halve; set by getenv TES_SCREEN_CHUNK_TILES (1 or 2)
static const int64_t PS_SCREEN_CHUNK = [] {
    const char* e = getenv("TES_SCREEN_CHUNK_TILES");
    return (e && e[0] == '1') ? PS_CHUNK : 2 * PS_CHUNK;
}();
static EpPlan ep_plan(int64_t N, int64_t T, int64_t n, int64_t d, int64_t Sp, int64_t chunk = PS_CHUNK) {
    EpPlan p;
    p.m = T + n;
    p.nb = (T + 15) / 16;
    p.nbp = p.nb * (p.nb + 1) / 2;
    p.Kdim = quad_num_slices(p.nb) * 16;
    p.nc = N < chunk ? (N < 1 ? 1 : N) : chunk;
    int64_t o = 0;
    auto take = [&](int64_t bytes) {
        int64_t r = o;
        o += align_up(bytes, 256);
        return r;
    };
    p.off_xto = take(p.m * d * 8);
    p.off_bext = take(p.m * (p.m + 1) * 8);
    p.off_kc = take(p.nc * p.m * 8);
    p.off_gc = take(p.nc * (p.m + 1) * 8);
    p.off_table = take((quad_num_slices(p.nb) + 1) * (int64_t)sizeof(QSlice));
    p.off_bpack = take(p.Kdim * Sp * 8);
    p.off_mean = take(p.nc * Sp * 8);
    p.off_var = take(p.nc * Sp * 8);
    p.off_kq = take(p.nc * 8);
    p.off_gobs = take(p.nc * (n > 0 ? n : 1) * 8);
    p.off_varf = take(p.nc * 8);
    p.total = o;
    return p;
}

}  // namespace tes

using namespace tes;

extern "C" int64_t tes_ep_acq_workspace_bytes(int64_t N, int64_t T, int64_t n, int64_t d, int64_t Sp) {
    if (N < 0 || T < 1 || n < 0 || d < 1 || Sp < 1) return -1;
    return ep_plan(N, T, n, d, Sp).total;
}

extern "C" int tes_ep_acq_f64(int mode, const double* x, int64_t N, int64_t d, const double* test_xs, int64_t T,
                              const double* X, int64_t n, const double* Y, const double* l, double sigma,
                              double sigma0, const double* invpNK, const double* invKobs, const double* p,
                              int64_t Sp, const double* post_mean, const double* post_cov, int niter, int nysample,
                              const double* z, uint64_t seed, int64_t n_global, int64_t x_offset, double* out_acq,
                              double* out_mean, double* out_var, void* ws, int64_t ws_bytes, void* stream) {
    if (mode < 0 || mode > 3) return -1;
    if (N < 0 || (N > 0 && !x)) return -2;
    if (d < 1 || d > 64) return -4;
    if (T < 1 || !test_xs) return -5;
    if (n < 0 || (n > 0 && (!X || !Y))) return -7;
    if (!l) return -10;
    if (!invpNK) return -13;
    if (mode == 0 && (!invKobs || n < 1)) return -14;
    if (!p) return -15;
    if (Sp < 1 || Sp > (1 << 20)) return -16;
    if (!post_mean) return -17;
    if (!post_cov) return -18;
    if (mode == 1 && (niter < 1 || nysample < 1)) return -19;
    if (mode != 3 && N > 0 && !out_acq) return -25;
    EpPlan pl = ep_plan(N, T, n, d, Sp);
    if (!ws || ws_bytes < pl.total) return -100;
    if (N == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    char* base = (char*)ws;
    double* xto = (double*)(base + pl.off_xto);
    double* Bext = (double*)(base + pl.off_bext);
    double* Kc = (double*)(base + pl.off_kc);
    double* Gc = (double*)(base + pl.off_gc);
    QSlice* table = (QSlice*)(base + pl.off_table);
    double* Bpack = (double*)(base + pl.off_bpack);
    double* meanc = (double*)(base + pl.off_mean);
    double* varc = (double*)(base + pl.off_var);
    double* Kq = (double*)(base + pl.off_kq);
    double* Gobs = (double*)(base + pl.off_gobs);
    double* varf = (double*)(base + pl.off_varf);
    const int64_t m = pl.m;
    int rc;
    if ((rc = launch_concat_rows(test_xs, T, X, n, (int)d, xto, st))) return rc;
    if ((rc = launch_build_bext(invpNK, Y, m, T, Bext, st))) return rc;
    quad_slice_table_kernel<<<1, 32, 0, st>>>((int)pl.nb, table);
    TES_LAUNCH_OK();
    {
        int64_t tot = pl.Kdim * Sp;
        quad_pack_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(post_cov, Sp, T, table, pl.Kdim, Bpack);
        TES_LAUNCH_OK();
    }
    if (mode == 1) {
        size_t smem = (size_t)(7 * Sp + 16 + 8) * sizeof(double);
        if (smem > 48 * 1024) {
            if (smem > 200 * 1024) return -16;
            TES_CUDA_OK(cudaFuncSetAttribute(ep_stoch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        }
    }
    for (int64_t c0 = 0; c0 < N; c0 += PS_CHUNK) {
        const int64_t nc = (N - c0 < PS_CHUNK) ? (N - c0) : PS_CHUNK;
        if ((rc = launch_se_ard(x + c0 * d, nc, xto, m, (int)d, l, sigma, -1, 0.0, Kc, m, st))) return rc;
        if ((rc = launch_gemm(Kc, m, Bext, m + 1, 1, Gc, m + 1, nc, m + 1, m, st))) return rc;
        if ((rc = launch_rowdot(Gc, m + 1, Kc, m, nc, m, sigma, -INFINITY, Kq, st))) return rc;
        // mean[x][j] = M mu_j^T + b
        double* mean_dst = out_mean ? out_mean + c0 * Sp : meanc;
        double* var_dst = out_var ? out_var + c0 * Sp : varc;
        if ((rc = launch_gemm(Gc, m + 1, post_mean, 1, T, mean_dst, Sp, nc, Sp, T, st))) return rc;
        add_col_kernel<<<(unsigned)((nc * Sp + 255) / 256), 256, 0, st>>>(mean_dst, nc, Sp, Gc + m, m + 1);
        TES_LAUNCH_OK();
        {
            dim3 grid((unsigned)((nc + GEMM_BM - 1) / GEMM_BM), (unsigned)((Sp + GEMM_BN - 1) / GEMM_BN));
            TES_CUDA_OK(cudaFuncSetAttribute(quadform_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)sizeof(GemmSmem)));
            TES_CUDA_OK(cudaFuncSetAttribute(quadform_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            quadform_kernel<<<grid, GEMM_THREADS, sizeof(GemmSmem), st>>>(Gc, m + 1, nc, T, table, pl.Kdim, Bpack, Sp, Kq, var_dst);
            TES_LAUNCH_OK();
        }
        if (mode == 0) {
            // H[y] from the observation-only posterior (utils.compute_mean_var_f, clip 1e-100)
            if ((rc = launch_gemm(Kc + T, m, invKobs, n, 1, Gobs, n, nc, n, n, st))) return rc;
            if ((rc = launch_rowdot(Gobs, n, Kc + T, m, nc, n, sigma, 1e-100, varf, st))) return rc;
            ep_det_kernel<<<(unsigned)((nc * 32 + 255) / 256), 256, 0, st>>>(var_dst, varf, p, nc, Sp, sigma0,
                                                                            out_acq + c0);
            TES_LAUNCH_OK();
        } else if (mode == 2) {
            ep_lite_kernel<<<(unsigned)((nc * 32 + 255) / 256), 256, 0, st>>>(mean_dst, var_dst, p, nc, Sp, sigma0,
                                                                             out_acq + c0);
            TES_LAUNCH_OK();
        } else if (mode == 1) {
            size_t smem = (size_t)(7 * Sp + 16 + 8) * sizeof(double);
            ep_stoch_kernel<<<(unsigned)nc, 128, smem, st>>>(mean_dst, var_dst, p, nc, (int)Sp, sigma0, niter,
                                                             nysample, z, N, c0, seed, n_global, x_offset + c0,
                                                             out_acq + c0);
            TES_LAUNCH_OK();
        }
    }
    return 0;
}

// ---- TES_ep deterministic, tensor-core SCREEN ------------------------------------------------------------------
// acq~(x) and a proven bound eps(x) on |acq~(x) - acq(x)|: the quadratic forms run as a 3xFP16 tcgen05 GEMM
// (rffprod.cu), everything else as in tes_ep_acq_f64.  The caller keeps the candidates with acq~ + eps >= max(acq~ - eps)
// and evaluates only those with the fp64 path, which leaves the arg-max and its value unchanged.
//   |q~_j - q_j| <= delta_j = QS_ETA * h_j^2 + QS_ABS * Ks * rowmax(x)^2 * colmax_j,
//   h_j = sum_a |M[x,a]| sqrt(Sigma_j[a][a])   (uses |Sigma_ab| <= sqrt(S_aa S_bb): CHECKED for every (a, b, j) by
//   screen_diag_kernel -- EP-mode covariances need not be positive semi-definite; a column j that violates it, or has a
//   negative / NaN diagonal entry, gets h_j = +inf, i.e. eps = +inf for every candidate: no pruning)
//   QS_ETA = 1.5 * (2.5 * 2^-20 [hi/lo fp16 operands, dropped lo*lo] + 4 * 2^-24 * 96 [fp32 segments of 256 k]
//                   + 64 * 2^-24 [fp32 running sum of <= 64 segments in the fused kernel; the sum of |terms| <= h^2])
//   QS_ABS = 1.5 * (4096 + 8) * 2^-25 / (4096 * 8): the fp16 lo halves are SUBNORMAL below 2^-14, so an operand carries an
//   absolute error of up to 2^-25 in scaled units (rows scaled to products <= 4096 = QS_AMAX, columns to <= 8 = QS_BMAX)
//   on top of the relative one; per term <= (|a| + |b|) 2^-25, summed over the Ks padded terms and unscaled by
//   rowmax^2 / 4096 * colmax / 8.  (Entries of Sigma_j or M far below colmax / rowmax are where it matters.)
//   |1/2 log(v~ + s0) - 1/2 log(v + s0)| <= 1/2 log(w / (w - delta)),  w = v~ + s0  (infinite when w <= delta or NaN)
namespace tes {
constexpr double QS_ETA = 4.6e-5;
constexpr double QS_ABS = 1.5 * 4104.0 * 2.98023223876953125e-08 / 32768.0;
// D[a][j] = sqrt(Sigma_j[a][a]); one CTA per maximizer j checks Sigma_ab^2 <= Sigma_aa Sigma_bb for every pair (the
// inequality the h-bound rests on) and a non-negative diagonal, else the whole column of D is +inf.  Df (optional):
// the same factor as floats ROUNDED UP, T x 128 with the columns j >= Sp zero (operand of ep_det_screen_h_kernel).
__global__ void __launch_bounds__(256) screen_diag_kernel(const double* __restrict__ cov, int64_t Sp, int64_t T,
                                                          double* __restrict__ D, float* __restrict__ Df) {
    __shared__ int bad;
    const int64_t j = blockIdx.x;
    if (j >= Sp) {                                     // zero padding columns of Df (grid = 128 in that mode)
        for (int64_t a = threadIdx.x; a < T; a += blockDim.x) Df[a * 128 + j] = 0.f;
        return;
    }
    const double* c = cov + j * T * T;
    if (threadIdx.x == 0) bad = 0;
    __syncthreads();
    int mybad = 0;
    for (int64_t e = threadIdx.x; e < T * T; e += blockDim.x) {
        const int64_t a = e / T, b = e % T;
        const double sab = c[e], saa = c[a * T + a], sbb = c[b * T + b];
        // (written so that NaN anywhere counts as a violation)
        if (!(saa >= 0.0) || !(sbb >= 0.0) || !(sab * sab <= saa * sbb * (1.0 + 1e-12))) mybad = 1;
    }
    if (mybad) bad = 1;
    __syncthreads();
    const int isbad = bad;
    for (int64_t a = threadIdx.x; a < T; a += blockDim.x) {
        const double v = isbad ? pos_inf() : sqrt(c[a * T + a]);
        D[a * Sp + j] = v;
        if (Df) Df[a * 128 + j] = __double2float_ru(v * (1.0 + 1e-15));   // >= the exact root (sqrt is within 1/2 ulp)
    }
}
__global__ void abs_rows_kernel(const double* __restrict__ G, int64_t ldg, int64_t rows, int64_t T,
                                double* __restrict__ out) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= rows * T) return;
    out[e] = fabs(G[(e / T) * ldg + e % T]);
}
// H[x][j] = sum_a |M[x,a]| sqrt(Sigma_j[a][a]) comes from a GEMM (abs_rows_kernel + launch_gemm)
__global__ void ep_det_screen_kernel(const double* __restrict__ Qt, const double* __restrict__ Kq,
                                     const double* __restrict__ H, const double* __restrict__ varf,
                                     const double* __restrict__ p, int64_t rows, int64_t Sp, double sigma0,
                                     const double* __restrict__ rowmax, const double* __restrict__ colmax, double kabs,
                                     double* __restrict__ out_acq, double* __restrict__ out_eps) {
    int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (row >= rows) return;
    double s = 0.0, er = 0.0;
    const double kq = Kq[row];
    const double rm = rowmax[row];
    const double ra = kabs * rm * rm;                    // QS_ABS * Ks * rowmax^2
    for (int64_t j = lane; j < Sp; j += 32) {
        const double h = H[row * Sp + j];
        const double w = (kq + Qt[row * Sp + j]) + sigma0;
        const double delta = (QS_ETA * h * h + ra * colmax[j]) + 1e-300;
        s += ((NORM_ENTROPY_CONST + log(sqrt(w))) + 0.5) * p[j];
        const double e1 = (w > delta) ? 0.5 * log(w / (w - delta)) : pos_inf();   // NaN w -> false -> +inf
        er += e1 * p[j];
    }
    s = warp_sum(s);
    er = warp_sum(er);
    if (lane == 0) {
        double ent = (NORM_ENTROPY_CONST + log(sqrt(varf[row] + sigma0))) + 0.5;
        const double a = ent - s / (double)Sp;
        out_acq[row] = a;
        out_eps[row] = er / (double)Sp + 1e-13 * fabs(a);   // + rounding slack of the two fp64 evaluations
    }
}

// The same, with the bound H computed HERE on the fp32 pipe instead of by abs_rows_kernel + an fp64 GEMM: H only has to
// be an UPPER bound of sum_a |M[x,a]| sqrt(Sigma_j[a][a]).  |M| and the roots are rounded UP to floats, the T-term
// fp32 sum of non-negative products loses at most (T + 1) 2^-24 of its value (and 1.4e-45 per product that underflows):
// H = fl32(sum) * (1 + 2e-5) + 1e-40 >= the exact sum for T <= 256.  Overflow gives +inf, i.e. eps = +inf (no pruning).
// A warp owns 4 candidate rows (their |M| as [a][4] floats in shared memory: one broadcast LDS.128 per a) and lane l
// the maximizer columns 4l .. 4l+3 (Df row a: one conflict-free LDS.128): 16 FFMA per two shared loads.
constexpr int EDS_WARPS = 16;
constexpr int EDS_ROWS = 4;
__global__ void __launch_bounds__(EDS_WARPS * 32, 2) ep_det_screen_h_kernel(
    const double* __restrict__ Qt, const double* __restrict__ G, int64_t ldg, const double* __restrict__ Kc, int64_t ldk,
    const double* __restrict__ Gobs, int T, int m, int n, double sigma, const float* __restrict__ Df,
    const double* __restrict__ p, int64_t rows, int Sp, double sigma0, const double* __restrict__ rowmax,
    const double* __restrict__ colmax, double kabs, double* __restrict__ out_acq, double* __restrict__ out_eps) {
    extern __shared__ __align__(16) float eds_sm[];
    float* Ds = eds_sm;                                            // [T][128]
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float* Ms = eds_sm + (size_t)T * 128 + (size_t)wid * T * EDS_ROWS;   // [T][4]
    for (int e = threadIdx.x; e < T * 32; e += blockDim.x)
        reinterpret_cast<float4*>(Ds)[e] = reinterpret_cast<const float4*>(Df)[e];
    double pj[4], cm[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const int j = 4 * lane + c;
        pj[c] = j < Sp ? p[j] : 0.0;
        cm[c] = j < Sp ? colmax[j] : 0.0;
    }
    __syncthreads();
    const int64_t ngroups = (rows + EDS_ROWS - 1) / EDS_ROWS;
    for (int64_t g = (int64_t)blockIdx.x * EDS_WARPS + wid; g < ngroups; g += (int64_t)gridDim.x * EDS_WARPS) {
        const int64_t r0 = g * EDS_ROWS;
        // one pass over the rows of G = K [invpNK | w] and K: the |M| tile, and the two row dots of the criterion
        //   kq = sigma - sum_c G[x][c] K[x][c]  (c < m),  varf = max(sigma - sum_c Gobs[x][c] K[x][T + c], 1e-100)
        // in the summation order of rowdot_kernel (lane-strided fma chain, butterfly) -> the bits of the fp64 criterion
        double kq4[EDS_ROWS], vf4[EDS_ROWS];
        {
            int64_t rr[EDS_ROWS];
#pragma unroll
            for (int r = 0; r < EDS_ROWS; ++r) rr[r] = r0 + r < rows ? r0 + r : rows - 1;
            double sk[EDS_ROWS] = {0.0, 0.0, 0.0, 0.0}, sv[EDS_ROWS] = {0.0, 0.0, 0.0, 0.0};
            for (int c = lane; c < m; c += 32) {
                double g[EDS_ROWS];
#pragma unroll
                for (int r = 0; r < EDS_ROWS; ++r) {
                    g[r] = G[rr[r] * ldg + c];
                    sk[r] = fma(g[r], Kc[rr[r] * ldk + c], sk[r]);
                }
                if (c < T)
                    *reinterpret_cast<float4*>(Ms + c * EDS_ROWS) =
                        make_float4(__double2float_ru(fabs(g[0])), __double2float_ru(fabs(g[1])),
                                    __double2float_ru(fabs(g[2])), __double2float_ru(fabs(g[3])));
            }
            for (int c = lane; c < n; c += 32) {
#pragma unroll
                for (int r = 0; r < EDS_ROWS; ++r) sv[r] = fma(Gobs[rr[r] * n + c], Kc[rr[r] * ldk + T + c], sv[r]);
            }
#pragma unroll
            for (int r = 0; r < EDS_ROWS; ++r) {
                kq4[r] = sigma - warp_sum(sk[r]);
                double v = sigma - warp_sum(sv[r]);
                if (v < 1e-100) v = 1e-100;
                vf4[r] = v;
            }
        }
        __syncwarp();
        float acc[EDS_ROWS][4];
#pragma unroll
        for (int r = 0; r < EDS_ROWS; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
#pragma unroll 4
        for (int a = 0; a < T; ++a) {
            const float4 mv = *reinterpret_cast<const float4*>(Ms + a * EDS_ROWS);
            const float4 dv = *reinterpret_cast<const float4*>(Ds + a * 128 + 4 * lane);
            const float mr[4] = {mv.x, mv.y, mv.z, mv.w};
            const float dc[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
            for (int r = 0; r < EDS_ROWS; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(mr[r], dc[c], acc[r][c]);
        }
        __syncwarp();
#pragma unroll
        for (int r = 0; r < EDS_ROWS; ++r) {
            const int64_t row = r0 + r;
            if (row >= rows) break;                    // uniform over the warp
            const double kq = kq4[r];
            const double rm = rowmax[row];
            const double ra = kabs * rm * rm;
            // screen arithmetic (cheaper than the criterion's own sequence, covered by the bound):
            //   log(sqrt(w)) as 0.5 log(w)   -- differs by rounding only: 1e-13 (|ent| + sum |term| p / Sp) of slack
            //   1/2 log(w / (w - delta)) <= 1/2 delta / (w - delta)   (-log(1 - t) <= t / (1 - t), t = delta / w)
            double s = 0.0, er = 0.0, sabs = 0.0;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int j = 4 * lane + c;
                if (j < Sp) {
                    const double h = (double)acc[r][c] * (1.0 + 2e-5) + 1e-40;
                    const double w = (kq + Qt[row * Sp + j]) + sigma0;
                    const double delta = (QS_ETA * h * h + ra * cm[c]) + 1e-300;
                    const double term = ((NORM_ENTROPY_CONST + 0.5 * log(w)) + 0.5) * pj[c];
                    s += term;
                    sabs += fabs(term);
                    const double e1 = (w > delta) ? 0.5 * delta / (w - delta) : pos_inf();   // NaN w -> +inf
                    er += e1 * pj[c];
                }
            }
            s = warp_sum(s);
            er = warp_sum(er);
            sabs = warp_sum(sabs);
            if (lane == 0) {
                const double ent = (NORM_ENTROPY_CONST + 0.5 * log(vf4[r] + sigma0)) + 0.5;
                const double a = ent - s / (double)Sp;
                out_acq[row] = a;
                out_eps[row] = er / (double)Sp * (1.0 + 1e-12) + 1e-13 * (fabs(ent) + sabs / (double)Sp);
            }
        }
    }
}
}  // namespace tes

extern "C" int64_t tes_ep_acq_screen_workspace_bytes(int64_t N, int64_t T, int64_t n, int64_t d, int64_t Sp) {
    if (N < 0 || T < 1 || n < 1 || d < 1 || Sp < 1 || Sp > 128) return -1;
    EpPlan pl = ep_plan(N, T, n, d, Sp, PS_SCREEN_CHUNK);
    const int64_t Ks = quad_screen_kdim(T);
    return pl.total + quad_screen_a_bytes(pl.nc, Ks) + quad_screen_b_bytes(Ks) + align_up(pl.nc * 8, 256) +
           align_up(T * Sp * 8, 256) + align_up(pl.nc * T * 8, 256) + align_up(Ks * Sp * 8, 256) +
           align_up(T * 128 * 4, 256) + (T <= 128 ? quad_tri_b_bytes(T) + 256 * 8 : 0);
}

extern "C" int tes_ep_acq_screen_f64(const double* x, int64_t N, int64_t d, const double* test_xs, int64_t T,
                                     const double* X, int64_t n, const double* Y, const double* l, double sigma,
                                     double sigma0, const double* invpNK, const double* invKobs, const double* p,
                                     int64_t Sp, const double* post_cov, double* out_acq, double* out_eps, void* ws,
                                     int64_t ws_bytes, void* stream) {
    if (N < 0 || (N > 0 && !x)) return -1;
    if (d < 1 || d > 64) return -3;
    if (T < 1 || !test_xs) return -4;
    if (n < 1 || !X || !Y) return -6;
    if (!l) return -9;
    if (!invpNK) return -12;
    if (!invKobs) return -13;
    if (!p) return -14;
    if (Sp < 1 || Sp > 128) return -15;
    if (!post_cov) return -16;
    if (N > 0 && (!out_acq || !out_eps)) return -17;
    EpPlan pl = ep_plan(N, T, n, d, Sp, PS_SCREEN_CHUNK);
    if (!ws || ws_bytes < tes_ep_acq_screen_workspace_bytes(N, T, n, d, Sp)) return -100;
    if (N == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    char* base = (char*)ws;
    double* xto = (double*)(base + pl.off_xto);
    double* Bext = (double*)(base + pl.off_bext);
    double* Kc = (double*)(base + pl.off_kc);
    double* Gc = (double*)(base + pl.off_gc);
    double* varc = (double*)(base + pl.off_var);
    double* Kq = (double*)(base + pl.off_kq);
    double* Gobs = (double*)(base + pl.off_gobs);
    double* varf = (double*)(base + pl.off_varf);
    char* extra = base + pl.total;
    void* Atc = extra;
    const int64_t Ks = quad_screen_kdim(T);      // the screen's own K axis (run order), not the DMMA slice table's
    extra += quad_screen_a_bytes(pl.nc, Ks);
    void* Btc = extra;
    double* colmax = (double*)(extra + align_up((quad_screen_kp(Ks) / 64) * (int64_t)(2 * 8 * 128 * 16), 256));
    extra += quad_screen_b_bytes(Ks);
    double* rowmax = (double*)extra;
    extra += align_up(pl.nc * 8, 256);
    double* Dm = (double*)extra;
    extra += align_up(T * Sp * 8, 256);
    double* absM = (double*)extra;
    extra += align_up(pl.nc * T * 8, 256);
    double* Bscr = (double*)extra;                   // Ks x Sp
    extra += align_up(Ks * Sp * 8, 256);
    float* Df = (float*)extra;                       // T x 128 floats, rounded up
    extra += align_up(T * 128 * 4, 256);
    void* Btri = extra;                              // triangular variant: row blocks of the packed covariances
    double* colmax_tri = (double*)(extra + (T <= 128 ? quad_tri_b_bytes(T) : 0));     // [colmax (128) | colscale (128)]
    // quadratic forms: triangular GEMM (A = the tile of M, no generated operand) for T <= 128; TES_QUAD_TRI=0: the
    // index-pair GEMM with the generated / streamed A operand
    const char* tmode = std::getenv("TES_QUAD_TRI");
    const bool tri = T <= 128 && !(tmode && tmode[0] == '0');
    double* Hb = (double*)(base + pl.off_mean);      // nc x Sp (the mean buffer is unused in this mode)
    // bound H on the fp32 pipe inside the criterion kernel whenever its operands fit in shared memory
    const size_t eds_smem = ((size_t)T * 128 + (size_t)EDS_WARPS * T * EDS_ROWS) * sizeof(float);
    const char* hmode = std::getenv("TES_SCREEN_H");
    const bool fuse_h = eds_smem <= 200 * 1024 && !(hmode && hmode[0] == '0');
    if (fuse_h)
        TES_CUDA_OK(cudaFuncSetAttribute(ep_det_screen_h_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)eds_smem));
    const int64_t m = pl.m;
    int rc;
    if ((rc = launch_concat_rows(test_xs, T, X, n, (int)d, xto, st))) return rc;
    if ((rc = launch_build_bext(invpNK, Y, m, T, Bext, st))) return rc;
    if (tri) {
        if ((rc = launch_quad_tri_build_b(post_cov, Sp, T, Btri, colmax_tri, st))) return rc;
        colmax = colmax_tri;
    } else {
        if ((rc = launch_quad_screen_ab(post_cov, Sp, T, colmax, Bscr, st))) return rc;
        if ((rc = launch_quad_screen_build_b(Bscr, Ks, Sp, Btc, colmax, st))) return rc;
    }
    screen_diag_kernel<<<fuse_h ? 128u : (unsigned)Sp, 256, 0, st>>>(post_cov, Sp, T, Dm, fuse_h ? Df : nullptr);
    TES_LAUNCH_OK();
    // absolute (fp16-subnormal) term of the bound: per (a, b) term (|A| + |B|) 2^-25 in scaled units -- index-pair GEMM:
    // A = products <= 4096, B <= 8; triangular GEMM: A = M~ <= 64, B <= 8 inside the MMA, times M~_a <= 64 in the
    // epilogue, whose own floor adds 64 * 8: 64 * 80 per term -- over the T (T + 1) / 2 terms, unscaled by 4096 * 8
    const double kabs = tri ? 1.5 * 64.0 * 80.0 * 2.98023223876953125e-08 / 32768.0 * (double)quad_tri_terms(T)
                            : QS_ABS * (double)quad_screen_kp(Ks);
    for (int64_t c0 = 0; c0 < N; c0 += PS_SCREEN_CHUNK) {
        const int64_t nc = (N - c0 < PS_SCREEN_CHUNK) ? (N - c0) : PS_SCREEN_CHUNK;
        if ((rc = launch_se_ard(x + c0 * d, nc, xto, m, (int)d, l, sigma, -1, 0.0, Kc, m, st))) return rc;
        if ((rc = launch_gemm(Kc, m, Bext, m + 1, 1, Gc, m + 1, nc, m + 1, m, st))) return rc;
        if (!fuse_h && (rc = launch_rowdot(Gc, m + 1, Kc, m, nc, m, sigma, -INFINITY, Kq, st))) return rc;
        if (tri) {
            if ((rc = launch_quad_screen_tri(Gc, m + 1, nc, T, Btri, colmax, Sp, rowmax, varc, st))) return rc;
        } else if ((rc = launch_quad_screen(Gc, m + 1, nc, T, Ks, Btc, colmax, Sp, Atc, rowmax, varc, st))) {
            return rc;
        }
        if ((rc = launch_gemm(Kc + T, m, invKobs, n, 1, Gobs, n, nc, n, n, st))) return rc;
        if (!fuse_h && (rc = launch_rowdot(Gobs, n, Kc + T, m, nc, n, sigma, 1e-100, varf, st))) return rc;
        if (fuse_h) {
            const int64_t ngroups = (nc + EDS_ROWS - 1) / EDS_ROWS;
            int64_t grid = (ngroups + EDS_WARPS - 1) / EDS_WARPS;
            const int64_t cap = 148 * (int64_t)(eds_smem <= 110 * 1024 ? 2 : 1);
            if (grid > cap) grid = cap;
            ep_det_screen_h_kernel<<<(unsigned)grid, EDS_WARPS * 32, eds_smem, st>>>(
                varc, Gc, m + 1, Kc, m, Gobs, (int)T, (int)m, (int)n, sigma, Df, p, nc, (int)Sp, sigma0, rowmax, colmax,
                kabs, out_acq + c0, out_eps + c0);
            TES_LAUNCH_OK();
            continue;
        }
        abs_rows_kernel<<<(unsigned)((nc * T + 255) / 256), 256, 0, st>>>(Gc, m + 1, nc, T, absM);
        TES_LAUNCH_OK();
        if ((rc = launch_gemm(absM, T, Dm, Sp, 1, Hb, Sp, nc, Sp, T, st))) return rc;
        ep_det_screen_kernel<<<(unsigned)((nc * 32 + 255) / 256), 256, 0, st>>>(varc, Kq, Hb, varf, p, nc, Sp, sigma0,
                                                                               rowmax, colmax, kabs, out_acq + c0,
                                                                               out_eps + c0);
        TES_LAUNCH_OK();
    }
    return 0;
}

// ---- TES_sp ------------------------------------------------------------------------------------------
namespace tes {
struct SpPlan {
    int64_t m, nc, rows_c;
    int64_t off_xto, off_bext, off_kc, off_gc, off_M, off_b, off_L, off_Li, off_meang, total;
    int nwarps, mean_in_smem;
    size_t smem;
    int64_t grid;
};
static SpPlan sp_plan(int64_t nx, int q, int64_t T, int64_t n, int64_t d, int64_t Sp, int64_t K) {
    SpPlan p;
    p.m = T + n;
    int64_t rows = nx * q;
    p.nc = rows < PS_CHUNK ? (rows < 1 ? 1 : rows) : (PS_CHUNK / q) * q;  // whole queries per chunk
    p.rows_c = p.nc / q;
    if (p.rows_c < 1) p.rows_c = 1;
    // shared memory: red[8] + lq + ybuf[nwarps][K*q] (+ mean[Sp*K*q] if it fits)
    const size_t fixed = (8 + 2 * q * q + q + 1) * sizeof(double);
    const size_t mean_b = (size_t)Sp * K * q * sizeof(double);
    const size_t budget = 200 * 1024;
    p.nwarps = 8;
    p.mean_in_smem = 1;
    while (p.nwarps > 1 && fixed + mean_b + (size_t)p.nwarps * K * q * 8 > budget) p.nwarps >>= 1;
    if (fixed + mean_b + (size_t)p.nwarps * K * q * 8 > budget) {
        p.mean_in_smem = 0;
        p.nwarps = 8;
        while (p.nwarps > 1 && fixed + (size_t)p.nwarps * K * q * 8 > budget) p.nwarps >>= 1;
    }
    p.smem = fixed + (p.mean_in_smem ? mean_b : 0) + (size_t)p.nwarps * K * q * 8;
    p.grid = p.rows_c < 148 * 8 ? p.rows_c : 148 * 8;
    int64_t o = 0;
    auto take = [&](int64_t bytes) {
        int64_t r = o;
        o += align_up(bytes, 256);
        return r;
    };
    p.off_xto = take(p.m * d * 8);
    p.off_bext = take(p.m * (p.m + 1) * 8);
    p.off_kc = take(p.nc * p.m * 8);
    p.off_gc = take(p.nc * (p.m + 1) * 8);
    p.off_M = take(p.nc * (T > 0 ? T : 1) * 8);
    p.off_b = take(p.nc * 8);
    p.off_L = take(p.rows_c * q * q * 8);
    p.off_Li = take(p.rows_c * q * q * 8);
    p.off_meang = take(p.mean_in_smem ? 256 : p.grid * Sp * K * q * 8);
    p.total = o + 256;  // + info word
    return p;
}
__global__ void extract_mb2_kernel(const double* __restrict__ G, int64_t N, int64_t m, int64_t T,
                                   double* __restrict__ outM, double* __restrict__ outb) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N * (T + 1)) return;
    int64_t i = e / (T + 1), c = e % (T + 1);
    if (c < T)
        outM[i * T + c] = G[i * (m + 1) + c];
    else
        outb[i] = G[i * (m + 1) + m];
}

template <int Q>
static int run_sp_chunk(const SpPlan& pl, const double* xq, int64_t rows, int d, const double* l, double sigma,
                        double sigma0, const double* xto, const double* Bext, double* Kc, double* Gc, double* Mq,
                        double* bq, double* Lmat, double* Linv, double* meang, int* info, int64_t T, const double* P,
                        const unsigned char* mask, const double* p, int Sp, int K, int niter, int ny, const double* z,
                        int64_t z_n, int64_t z_x0, uint64_t seed, int64_t n_global, int64_t x_global0, double* out,
                        cudaStream_t st) {
    const int64_t m = pl.m;
    const int64_t nr = rows * Q;
    int rc;
    if ((rc = launch_se_ard(xq, nr, xto, m, d, l, sigma, -1, 0.0, Kc, m, st))) return rc;
    if ((rc = launch_gemm(Kc, m, Bext, m + 1, 1, Gc, m + 1, nr, m + 1, m, st))) return rc;
    extract_mb2_kernel<<<(unsigned)((nr * (T + 1) + 255) / 256), 256, 0, st>>>(Gc, nr, m, T, Mq, bq);
    TES_LAUNCH_OK();
    sp_cov_warp_kernel<Q><<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(xq, d, l, sigma, sigma0, Gc, m + 1, Kc, m, m, rows,
                                                                      Lmat, Linv, info);
    TES_LAUNCH_OK();
    // whitened kernel (sp2.cuh) whenever nu fits in shared memory; TES_SP_METHOD=1 forces the direct kernel
    const char* force = std::getenv("TES_SP_METHOD");
    int nw2 = 0;
    if (!(force && force[0] == '1') && K <= 65535) {
        for (int nw = 16; nw >= 4 && !nw2; nw >>= 1) {
            Sp2Smem L2 = sp2_layout(Q, Sp, K, nw);
            if (L2.rounds <= 4 && L2.bytes <= 216 * 1024) nw2 = nw;
        }
    }
    if (nw2) {
        Sp2Smem L2 = sp2_layout(Q, Sp, K, nw2);
        auto k2 = sp2_kernel<Q>;
        TES_CUDA_OK(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L2.bytes));
        int64_t grid2 = rows < 148 * 4 ? rows : 148 * 4;
        k2<<<(unsigned)grid2, nw2 * 32, L2.bytes, st>>>(Mq, bq, Linv, Lmat, P, mask, p, rows, Sp, (int)T, K, niter, ny, z,
                                                        z_n, z_x0, seed, n_global, x_global0, out);
        TES_LAUNCH_OK();
        return 0;
    }
    auto kern = sp_kernel<Q>;
    if (pl.smem > 48 * 1024)
        TES_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    int64_t grid = rows < pl.grid ? rows : pl.grid;
    kern<<<(unsigned)grid, pl.nwarps * 32, pl.smem, st>>>(Mq, bq, Linv, Lmat, P, mask, p, rows, Sp, (int)T, K, niter, ny,
                                                          z, z_n, z_x0, seed, n_global, x_global0, meang,
                                                          pl.mean_in_smem, out);
    TES_LAUNCH_OK();
    return 0;
}
}  // namespace tes

extern "C" int64_t tes_sp_acq_workspace_bytes(int64_t nx, int q, int64_t T, int64_t n, int64_t d, int64_t Sp,
                                              int64_t K) {
    if (nx < 0 || q < 1 || q > 8 || T < 1 || n < 0 || d < 1 || Sp < 1 || K < 1) return -1;
    return sp_plan(nx, q, T, n, d, Sp, K).total;
}

extern "C" int tes_sp_acq_f64(const double* x, int64_t nx, int q, int64_t d, const double* test_xs, int64_t T,
                              const double* X, int64_t n, const double* Y, const double* l, double sigma,
                              double sigma0, const double* invpNK, const double* p, int64_t Sp,
                              const double* post_samples, const unsigned char* post_mask, int64_t K, int niter,
                              int nysample, const double* z, uint64_t seed, int64_t n_global, int64_t x_offset,
                              double* out_acq, void* ws, int64_t ws_bytes, void* stream) {
    if (nx < 0 || (nx > 0 && !x)) return -1;
    if (q < 1 || q > 8) return -3;
    if (d < 1 || d > 64) return -4;
    if (T < 1 || !test_xs) return -5;
    if (n < 0 || (n > 0 && (!X || !Y))) return -7;
    if (!l) return -10;
    if (!invpNK) return -13;
    if (!p) return -14;
    if (Sp < 1) return -15;
    if (!post_samples) return -16;
    if (!post_mask) return -17;
    if (K < 1) return -18;
    if (niter < 1) return -19;
    if (nysample < 1) return -20;
    if (nx > 0 && !out_acq) return -25;
    SpPlan pl = sp_plan(nx, q, T, n, d, Sp, K);
    if (!ws || ws_bytes < pl.total) return -100;
    if (pl.smem > 220 * 1024) return -18;
    if (nx == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    char* base = (char*)ws;
    double* xto = (double*)(base + pl.off_xto);
    double* Bext = (double*)(base + pl.off_bext);
    double* Kc = (double*)(base + pl.off_kc);
    double* Gc = (double*)(base + pl.off_gc);
    double* Mq = (double*)(base + pl.off_M);
    double* bq = (double*)(base + pl.off_b);
    double* Lmat = (double*)(base + pl.off_L);
    double* Linv = (double*)(base + pl.off_Li);
    double* meang = (double*)(base + pl.off_meang);
    int* info = (int*)(base + pl.total - 256);
    int rc;
    if ((rc = launch_concat_rows(test_xs, T, X, n, (int)d, xto, st))) return rc;
    if ((rc = launch_build_bext(invpNK, Y, pl.m, T, Bext, st))) return rc;
    TES_CUDA_OK(cudaMemsetAsync(info, 0, sizeof(int), st));
    for (int64_t r0 = 0; r0 < nx; r0 += pl.rows_c) {
        const int64_t rows = (nx - r0 < pl.rows_c) ? (nx - r0) : pl.rows_c;
        const double* xq = x + r0 * q * d;
#define TES_SP_CASE(QQ)                                                                                               \
    case QQ:                                                                                                          \
        rc = run_sp_chunk<QQ>(pl, xq, rows, (int)d, l, sigma, sigma0, xto, Bext, Kc, Gc, Mq, bq, Lmat, Linv, meang,   \
                              info, T, post_samples, post_mask, p, (int)Sp, (int)K, niter, nysample, z, nx, r0, seed, \
                              n_global, x_offset + r0, out_acq + r0, st);                                            \
        break;
        switch (q) {
            TES_SP_CASE(1) TES_SP_CASE(2) TES_SP_CASE(3) TES_SP_CASE(4) TES_SP_CASE(5) TES_SP_CASE(6) TES_SP_CASE(7)
            TES_SP_CASE(8)
        }
#undef TES_SP_CASE
        if (rc) return rc;
    }
    return 0;
}

extern "C" int tes_batch_ep_acq_f64(const double* p, int64_t Sp, const double* x, int64_t nx, int64_t q, int64_t d,
                                    double* out_acq, void* stream) {
    if (!p) return -1;
    if (Sp < 1) return -2;
    if (nx < 0) return -4;
    if (nx > 0 && !x) return -3;
    if (q < 1) return -5;
    if (d < 1) return -6;
    if (nx > 0 && !out_acq) return -7;
    if (nx == 0) return 0;
    batch_ep_kernel<<<(unsigned)((nx + 255) / 256), 256, 0, (cudaStream_t)stream>>>(out_acq, nx, p, Sp, x, q * d);
    TES_LAUNCH_OK();
    return 0;
}
