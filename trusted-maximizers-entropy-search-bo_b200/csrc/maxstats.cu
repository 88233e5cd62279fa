// maxstats.cu -- stage C on the device: de-duplication of the trusted maximizers, joint posterior samples at the
// test points, arg-max buckets, empirical moments / padded sample tensors, and the batched symmetric eigen-solver
// (one-sided Jacobi, one CTA per matrix) that supplies the colouring factor and the RFF posterior weight draw.
//
// Reference (host numpy there): utils.py:1554-1581 (remove_duplicates_np), utils.py:1654-1729
// (sample_multivariate_normal_maxidx_np), utils.py:1733-1833 (get_testidxs_stats),
// empirical_approximation.py:86-107, optfunc.py:355-374 (np.linalg.eig of the n x n feature Gram).
#include "tes_common.cuh"

namespace tes {

// ---- counter-based normal draws (tes_spec.h) ---------------------------------------------------------------
__global__ void normal_fill_kernel(uint64_t seed, uint32_t iteration, uint32_t stream_id, uint64_t first,
                                   int64_t count, double* __restrict__ out) {
    // thread = one Philox call = the two components of a Box-Muller pair
    const uint64_t p0 = first >> 1;
    const uint64_t p1 = (first + (uint64_t)count + 1) >> 1;
    const uint64_t pair = p0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pair >= p1) return;
    double z0, z1;
    normal_pair(seed, iteration, stream_id, pair, z0, z1);
    const uint64_t e0 = 2 * pair, e1 = 2 * pair + 1;
    if (e0 >= first && e0 < first + (uint64_t)count) out[e0 - first] = z0;
    if (e1 >= first && e1 < first + (uint64_t)count) out[e1 - first] = z1;
}

// uniform draws on (0,1) from the same counters: element e -> pair e >> 1, component e & 1, u = (k + 0.5) 2^-53 with
// k the top 53 bits of (x0 | x1 << 32) resp. (x2 | x3 << 32)  (the u1 / u2 of the Box-Muller pair in tes_spec.h)
__global__ void uniform_fill_kernel(uint64_t seed, uint32_t iteration, uint32_t stream_id, uint64_t first,
                                    int64_t count, double* __restrict__ out) {
    const uint64_t p0 = first >> 1;
    const uint64_t p1 = (first + (uint64_t)count + 1) >> 1;
    const uint64_t pair = p0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pair >= p1) return;
    uint32_t c0 = (uint32_t)pair, c1 = (uint32_t)(pair >> 32), c2 = iteration, c3 = stream_id;
    philox4x32_10(c0, c1, c2, c3, (uint32_t)seed, (uint32_t)(seed >> 32));
    const uint64_t a = ((uint64_t)c1 << 32) | c0, b = ((uint64_t)c3 << 32) | c2;
    const uint64_t e0 = 2 * pair, e1 = 2 * pair + 1;
    if (e0 >= first && e0 < first + (uint64_t)count) out[e0 - first] = ((double)(a >> 11) + 0.5) * 0x1.0p-53;
    if (e1 >= first && e1 < first + (uint64_t)count) out[e1 - first] = ((double)(b >> 11) + 0.5) * 0x1.0p-53;
}

// ---- de-duplication (utils.py:1554-1573) ---------------------------------------------------------------------
// np.sum of a contiguous run of n doubles (numpy's pairwise summation, n <= 128 branch), applied to the squared
// differences of two rows: identical operation order => identical distance bits => identical keep / drop decisions.
__device__ __forceinline__ double np_sqdist(const double* __restrict__ a, const double* __restrict__ b, int n) {
    auto sq = [&](int k) {
        const double df = __dsub_rn(a[k], b[k]);
        return __dmul_rn(df, df);
    };
    if (n < 8) {
        double r = 0.0;
        for (int k = 0; k < n; ++k) r = __dadd_rn(r, sq(k));
        return r;
    }
    double r[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = sq(k);
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) r[k] = __dadd_rn(r[k], sq(i + k));
    }
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                           __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __dadd_rn(res, sq(i));
    return res;
}

// Walking the rows in order, every not-yet-flagged row flags all LATER rows within `resolution` (Euclidean).
// The walk is sequential by definition; one CTA, the inner scan over later rows is parallel.
// stage != 0: the rows are copied to shared memory first (each of the up to S leader iterations then reads them at
// shared-memory instead of L2 latency: 0.77 -> ~0.2 ms for the 1024 maximizers of C3).
__global__ void __launch_bounds__(1024) dedup_kernel(const double* __restrict__ xs_g, int S, int d, double resolution,
                                                     int* __restrict__ mask, int64_t* __restrict__ leader, int stage) {
    extern __shared__ __align__(8) unsigned char dd_smem[];
    double* xs_s = reinterpret_cast<double*>(dd_smem);
    int* smask = reinterpret_cast<int*>(dd_smem + (stage ? (size_t)S * d * sizeof(double) : 0));
    if (stage)
        for (int e = threadIdx.x; e < S * d; e += blockDim.x) xs_s[e] = xs_g[e];
    const double* xs = stage ? xs_s : xs_g;
    for (int j = threadIdx.x; j < S; j += blockDim.x) {
        smask[j] = 0;
        leader[j] = j;
    }
    __syncthreads();
    for (int i = 0; i < S - 1; ++i) {
        if (smask[i]) continue;  // uniform: written before the barrier that ended an earlier iteration
        const double* xi = xs + (int64_t)i * d;
        for (int j = i + 1 + threadIdx.x; j < S; j += blockDim.x) {
            if (smask[j]) continue;
            const double dist = __dsqrt_rn(np_sqdist(xs + (int64_t)j * d, xi, d));
            if (dist <= resolution) {
                smask[j] = 1;
                leader[j] = i;
            }
        }
        __syncthreads();
    }
    for (int j = threadIdx.x; j < S; j += blockDim.x) mask[j] = smask[j];
}

// ---- batched symmetric eigen-decomposition -----------------------------------------------------------------------
// One-sided (Hestenes) Jacobi applied to the columns of the symmetric matrix A itself: plane rotations from the right
// make the columns of G = A R mutually orthogonal; then G = V diag(lambda) with V orthogonal, i.e. the normalised
// columns are eigenvectors and the Rayleigh quotients v^T A v the eigenvalues.  Only ONE n x n array is needed, so a
// matrix up to n = 160 lives entirely in shared memory (larger ones in a global scratch that stays in L2).  Pairs of a
// round-robin tournament step are disjoint: one warp per pair, (n-1) steps per sweep.
constexpr int EIG_MAX_SWEEPS = 40;

// One tournament step = one block barrier.  Eight lanes own a pair (four pairs per warp at a time): they form the
// three inner products (|g_p|^2, |g_q|^2, g_p . g_q) with a 3-round butterfly, derive the rotation and apply it.
// The rotation angle only has to be ROUGHLY right (an error of 1e-7 leaves a residual 1e-7 g, gone in the next
// sweep), so t = tan(theta) is formed in fp32 on exponent-normalised inputs (MUFU square root / reciprocal); what has
// to be exact is orthogonality, so c = (1 + t^2)^-1/2 comes from an fp32 seed and two fp64 Newton steps and s = c t:
// c^2 + s^2 = 1 to rounding whatever t is.  This replaces the dependent fp64 sqrt / division / rsqrt sequences
// (~1700 cycles of latency per step, issued redundantly by every warp) that bound the first versions: ncu of the
// one-warp-per-pair kernel showed 1200 warp-instructions per warp and step (index divisions, 64-bit shuffles) at 4.4 us
// per step for n = 64.  The pair schedule is the round-robin tournament, advanced without integer divisions.
__device__ __forceinline__ void jacobi_rotation(double a, double b, double g, double tol2, double& c, double& sn) {
    c = 1.0;
    sn = 0.0;
    if (!(g != 0.0 && g * g > tol2 * (a * b))) return;
    const double tau = b - a, gam = 2.0 * g;
    // t = sign(zeta) / (|zeta| + sqrt(1 + zeta^2)), zeta = tau / gam  ==  sign(tau) gam / (|tau| + hypot(tau, gam))
    const int et = (__double2hiint(tau) >> 20) & 0x7ff, eg = (__double2hiint(gam) >> 20) & 0x7ff;
    const int em = et > eg ? et : eg;
    double t;
    if (em > 64 && em < 1980) {
        const double sc = __hiloint2double((2046 - em) << 20, 0);          // 2^(1023 - em): max(|tau|, |gam|) -> [1, 2)
        const float ft = (float)(tau * sc), fg = (float)(gam * sc);
        const float h = sqrtf(fmaf(ft, ft, fg * fg));
        const float tf = __fdividef(fg, fabsf(ft) + h);
        t = (double)(ft < 0.0f ? -tf : tf);
    } else {
        const double h = sqrt(fma(tau, tau, gam * gam));
        t = (tau < 0.0 ? -gam : gam) / (fabs(tau) + h);
    }
    const double x = fma(t, t, 1.0);
    double r = (double)rsqrtf((float)x);
    r = r * fma(-0.5 * x, r * r, 1.5);
    r = r * fma(-0.5 * x, r * r, 1.5);
    c = r;
    sn = r * t;
}

template <bool SMEM, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) sym_eig_kernel(const double* __restrict__ A, int n, double* __restrict__ evals,
                               double* __restrict__ evecs, int* __restrict__ sweeps_out, double* __restrict__ gws,
                               const double* __restrict__ Lfac, const int* __restrict__ linfo) {
    extern __shared__ double sm_g[];
    __shared__ int s_rot;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    const int64_t mat = blockIdx.x;
    const double* Ab = A + mat * n * n;
    const int ld = SMEM ? n + 1 : n;                      // odd leading dimension: columns start in different banks
    double* G = SMEM ? sm_g : gws + mat * n * n;
    const int m = n + (n & 1), npair = m / 2, mm1 = m - 1;
    // Start: column p of G = row p of A (A is symmetric) -- or, when a Cholesky factor A = L L^T is supplied (and the
    // factorisation of this matrix succeeded), column p of G = column p of L: the rotations then orthogonalise the
    // columns of L (G = L R, G^T G diagonal), whose normalised columns are the eigenvectors of L L^T = A all the same,
    // but L is far closer to orthogonal columns than A (cond(L) = sqrt(cond(A))): 6-7 sweeps instead of 13 on the Gram
    // matrices of the weight draw (Veselic-Hari).  The eigenvalues are Rayleigh quotients against A either way.
    if (Lfac && linfo[mat] == 0) {
        const double* Lb = Lfac + mat * n * n;
        for (int e = tid; e < n * n; e += nt) G[(e % n) * ld + (e / n)] = Lb[e];
    } else {
        for (int e = tid; e < n * n; e += nt) G[(e / n) * ld + (e % n)] = Ab[e];
    }
    __syncthreads();
    const double tol = 2.2e-16 * sqrt((double)n);
    const double tol2 = tol * tol;
    const int grp = lane >> 3, sub = lane & 7;            // pair slot of the warp, lane inside the pair
    const int kstride = nwarps * 4;
    int sweep = 0;
    for (; sweep < EIG_MAX_SWEEPS && n > 1; ++sweep) {
        if (tid == 0) s_rot = 0;
        __syncthreads();
        for (int s = 0; s < mm1; ++s) {
            for (int k0 = warp * 4; k0 < npair; k0 += kstride) {
                const int k = k0 + grp;
                // round robin: pair 0 = (m - 1, s); pair k = ((s + k) mod (m - 1), (s - k) mod (m - 1))
                int p = s + k, q = s - k;
                if (p >= mm1) p -= mm1;
                if (q < 0) q += mm1;
                if (k == 0) p = mm1;
                const bool live = k < npair && p < n && q < n;   // an odd n pads with an idle player
                double* gp = G + (live ? p : 0) * ld;
                double* gq = G + (live ? q : 0) * ld;
                double a = 0.0, b = 0.0, g = 0.0;
                if (MAXT == 256) {
                    // n <= 64 (the instance with <= 256 threads): a lane's 8 elements of both columns stay in registers between the inner products and the
                    // rotation (one shared-memory read per element and step instead of two); same operation order
                    double xr[8], yr[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int i = sub + 8 * e;
                        const bool in = live && i < n;
                        xr[e] = in ? gp[i] : 0.0;
                        yr[e] = in ? gq[i] : 0.0;
                    }
                    double a1 = 0.0, b1 = 0.0, g1 = 0.0;
#pragma unroll
                    for (int e = 0; e < 8; e += 2) {
                        a = fma(xr[e], xr[e], a);
                        b = fma(yr[e], yr[e], b);
                        g = fma(xr[e], yr[e], g);
                        a1 = fma(xr[e + 1], xr[e + 1], a1);
                        b1 = fma(yr[e + 1], yr[e + 1], b1);
                        g1 = fma(xr[e + 1], yr[e + 1], g1);
                    }
                    a += a1;
                    b += b1;
                    g += g1;
#pragma unroll
                    for (int off = 4; off > 0; off >>= 1) {
                        a += __shfl_xor_sync(0xffffffffu, a, off);
                        b += __shfl_xor_sync(0xffffffffu, b, off);
                        g += __shfl_xor_sync(0xffffffffu, g, off);
                    }
                    double c, sn;
                    jacobi_rotation(a, b, g, tol2, c, sn);
                    if (live && sn != 0.0) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const int i = sub + 8 * e;
                            if (i < n) {
                                gp[i] = c * xr[e] - sn * yr[e];
                                gq[i] = sn * xr[e] + c * yr[e];
                            }
                        }
                        if (sub == 0) s_rot = 1;
                    }
                    continue;
                }
                if (live) {
                    double a1 = 0.0, b1 = 0.0, g1 = 0.0;
                    int i = sub;
                    for (; i + 8 < n; i += 16) {
                        const double x = gp[i], y = gq[i], x1 = gp[i + 8], y1 = gq[i + 8];
                        a = fma(x, x, a);
                        b = fma(y, y, b);
                        g = fma(x, y, g);
                        a1 = fma(x1, x1, a1);
                        b1 = fma(y1, y1, b1);
                        g1 = fma(x1, y1, g1);
                    }
                    if (i < n) {
                        const double x = gp[i], y = gq[i];
                        a = fma(x, x, a);
                        b = fma(y, y, b);
                        g = fma(x, y, g);
                    }
                    a += a1;
                    b += b1;
                    g += g1;
                }
#pragma unroll
                for (int off = 4; off > 0; off >>= 1) {
                    a += __shfl_xor_sync(0xffffffffu, a, off);
                    b += __shfl_xor_sync(0xffffffffu, b, off);
                    g += __shfl_xor_sync(0xffffffffu, g, off);
                }
                // every lane of the group holds the same (a, b, g) and derives the same rotation
                double c, sn;
                jacobi_rotation(a, b, g, tol2, c, sn);
                if (live && sn != 0.0) {
                    for (int i = sub; i < n; i += 8) {
                        const double x = gp[i], y = gq[i];
                        gp[i] = c * x - sn * y;
                        gq[i] = sn * x + c * y;
                    }
                    if (sub == 0) s_rot = 1;
                }
            }
            __syncthreads();
        }
        const int again = s_rot;
        __syncthreads();
        if (!again) {
            ++sweep;
            break;
        }
    }
    if (tid == 0 && sweeps_out) sweeps_out[mat] = sweep;
    // normalise the columns
    for (int k = warp; k < n; k += nwarps) {
        double* gk = G + k * ld;
        double a = 0.0;
        for (int i = lane; i < n; i += 32) a = fma(gk[i], gk[i], a);
        a = warp_sum(a);
        const double inv = a > 0.0 ? 1.0 / sqrt(a) : 0.0;
        for (int i = lane; i < n; i += 32) gk[i] *= inv;
    }
    __syncthreads();
    // eigenvalues: Rayleigh quotients against the ORIGINAL matrix; eigenvectors out as columns of a row-major (n, n)
    for (int k = warp; k < n; k += nwarps) {
        const double* vk = G + k * ld;
        double acc = 0.0;
        for (int i = 0; i < n; ++i) {
            double r = 0.0;
            for (int j = lane; j < n; j += 32) r = fma(Ab[(int64_t)i * n + j], vk[j], r);
            acc = fma(vk[i], r, acc);  // every lane accumulates its partial of (A v)_i * v_i
        }
        acc = warp_sum(acc);
        if (lane == 0) evals[mat * n + k] = acc;
    }
    double* Vb = evecs + mat * n * n;
    for (int e = tid; e < n * n; e += nt) {
        const int i = e / n, k = e % n;
        Vb[e] = G[k * ld + i];
    }
}

// ---- joint-sample statistics (utils.py:1654-1729) ------------------------------------------------------------------
// column means of x (ns, T): block (32 cols, 32 row-lanes), fixed-order tree over the row lanes
__global__ void __launch_bounds__(1024) col_mean_kernel(const double* __restrict__ x, int64_t ns, int T,
                                                        double* __restrict__ mean) {
    __shared__ double part[32][33];
    const int t = blockIdx.x * 32 + threadIdx.x;
    double a = 0.0;
    if (t < T)
        for (int64_t r = threadIdx.y; r < ns; r += 32) a += x[r * T + t];
    part[threadIdx.y][threadIdx.x] = a;
    __syncthreads();
    if (threadIdx.y == 0 && t < T) {
        double s = 0.0;
        for (int y = 0; y < 32; ++y) s += part[y][threadIdx.x];
        mean[t] = s / (double)ns;
    }
}

// partial centred Gram matrices: gram[split][a][b] = sum over the rows of the split of (x[r][a]-m[a])(x[r][b]-m[b])
constexpr int GRAM_AT = 4, GRAM_SPLIT = 8;
__global__ void __launch_bounds__(256) gram_kernel(const double* __restrict__ x, int64_t ns, int T,
                                                   const double* __restrict__ mean, double* __restrict__ gram) {
    const int a0 = blockIdx.x * GRAM_AT, split = blockIdx.y;
    const int64_t r0 = ns * split / GRAM_SPLIT, r1 = ns * (split + 1) / GRAM_SPLIT;
    double ma[GRAM_AT];
#pragma unroll
    for (int u = 0; u < GRAM_AT; ++u) ma[u] = a0 + u < T ? mean[a0 + u] : 0.0;
    for (int b = threadIdx.x; b < T; b += blockDim.x) {
        const double mb = mean[b];
        double acc[GRAM_AT] = {0.0, 0.0, 0.0, 0.0};
        for (int64_t r = r0; r < r1; ++r) {
            const double xb = x[r * T + b] - mb;
#pragma unroll
            for (int u = 0; u < GRAM_AT; ++u)
                if (a0 + u < T) acc[u] = fma(x[r * T + a0 + u] - ma[u], xb, acc[u]);
        }
#pragma unroll
        for (int u = 0; u < GRAM_AT; ++u)
            if (a0 + u < T) gram[((int64_t)split * T + a0 + u) * T + b] = acc[u];
    }
}

// np.corrcoef semantics: c = G/(ns-1); c /= sd[:,None]; c /= sd[None,:]; clip to [-1,1]; dimension i is suppressed
// when some j < i has c[i][j] > threshold (utils.py:1688-1694)
__global__ void corr_flag_kernel(const double* __restrict__ gram, int64_t ns, int T, double threshold,
                                 int* __restrict__ masked) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T) return;
    auto G = [&](int a, int b) {
        double s = 0.0;
        for (int sp = 0; sp < GRAM_SPLIT; ++sp) s += gram[((int64_t)sp * T + a) * T + b];
        return s / (double)(ns - 1);
    };
    const double sdi = sqrt(G(i, i));
    int flag = 0;
    for (int j = 0; j < i; ++j) {
        double c = G(i, j) / sdi / sqrt(G(j, j));
        c = fmin(fmax(c, -1.0), 1.0);
        if (c > threshold) flag = 1;
    }
    masked[i] = flag;
}

// arg-max of every sample row over the non-suppressed dimensions (first maximum wins) + bucket counts
__global__ void __launch_bounds__(256) row_argmax_kernel(const double* __restrict__ x, int64_t ns, int T,
                                                         const int* __restrict__ masked,
                                                         int64_t* __restrict__ maxidx, int* __restrict__ counts) {
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= ns) return;
    double bv = neg_inf();
    int64_t bi = lane < T ? lane : (int64_t)1 << 40;
    for (int t = lane; t < T; t += 32) {
        const double v = masked[t] ? neg_inf() : x[r * T + t];
        if (v > bv) {
            bv = v;
            bi = t;
        }
    }
    warp_argmax(bv, bi);
    if (lane == 0) {
        maxidx[r] = bi;
        atomicAdd(&counts[bi], 1);
    }
}

// per-bucket mean over the kept columns: mean[j][c] = sum_{r in bucket j} x[order[r]][cols[c]] / n_j (row order)
__global__ void __launch_bounds__(256) bucket_mean_kernel(const double* __restrict__ x, int T,
                                                          const int64_t* __restrict__ order,
                                                          const int64_t* __restrict__ starts,
                                                          const int64_t* __restrict__ cols, int nk,
                                                          double* __restrict__ mean) {
    const int j = blockIdx.x;
    const int64_t r0 = starts[j], r1 = starts[j + 1];
    for (int c = threadIdx.x; c < nk; c += blockDim.x) {
        const int64_t col = cols[c];
        double s = 0.0;
        for (int64_t r = r0; r < r1; ++r) s += x[order[r] * T + col];
        mean[(int64_t)j * nk + c] = s / (double)(r1 - r0);
    }
}

// per-bucket covariance: cov[j][a][b] = sum_r (x_a - m_a)(x_b - m_b) / (n_j - 1)  (empirical_approximation.py:100-105
// with 0/1 weights = the bucket rows)
constexpr int BCOV_AT = 4;
__global__ void __launch_bounds__(128) bucket_cov_kernel(const double* __restrict__ x, int T,
                                                         const int64_t* __restrict__ order,
                                                         const int64_t* __restrict__ starts,
                                                         const int64_t* __restrict__ cols, int nk,
                                                         const double* __restrict__ mean, double* __restrict__ cov) {
    const int j = blockIdx.x, a0 = blockIdx.y * BCOV_AT;
    const int64_t r0 = starts[j], r1 = starts[j + 1];
    const double* mj = mean + (int64_t)j * nk;
    double ma[BCOV_AT];
    int64_t ca[BCOV_AT];
#pragma unroll
    for (int u = 0; u < BCOV_AT; ++u) {
        ma[u] = a0 + u < nk ? mj[a0 + u] : 0.0;
        ca[u] = a0 + u < nk ? cols[a0 + u] : 0;
    }
    const double denom = (double)(r1 - r0) - 1.0;
    for (int b = threadIdx.x; b < nk; b += blockDim.x) {
        const int64_t cb = cols[b];
        const double mb = mj[b];
        double acc[BCOV_AT] = {0.0, 0.0, 0.0, 0.0};
        for (int64_t r = r0; r < r1; ++r) {
            const double* row = x + order[r] * T;
            const double xb = row[cb] - mb;
#pragma unroll
            for (int u = 0; u < BCOV_AT; ++u) acc[u] = fma(row[ca[u]] - ma[u], xb, acc[u]);
        }
#pragma unroll
        for (int u = 0; u < BCOV_AT; ++u)
            if (a0 + u < nk) cov[((int64_t)j * nk + a0 + u) * nk + b] = acc[u] / denom;
    }
}

// padded per-maximizer sample tensor (utils.py:1716-1727): out[j][t][s] = s < n_j ? x[order[r0+s]][cols[t]]
// : prior_mean[cols[t]]; mask[j][s] = s < n_j
__global__ void __launch_bounds__(256) bucket_samples_kernel(const double* __restrict__ x, int T,
                                                             const int64_t* __restrict__ order,
                                                             const int64_t* __restrict__ starts,
                                                             const int64_t* __restrict__ cols, int nk, int K,
                                                             const double* __restrict__ prior_mean,
                                                             double* __restrict__ out, uint8_t* __restrict__ mask) {
    const int j = blockIdx.x, t = blockIdx.y;
    const int64_t r0 = starts[j];
    const int nj = (int)(starts[j + 1] - r0);
    const int64_t col = cols[t];
    const double pm = prior_mean[col];
    for (int s = threadIdx.x; s < K; s += blockDim.x) {
        out[((int64_t)j * nk + t) * K + s] = s < nj ? x[order[r0 + s] * T + col] : pm;
        if (t == 0) mask[(int64_t)j * K + s] = s < nj ? 1 : 0;
    }
}

__global__ void add_scalar_kernel(double* __restrict__ x, int64_t n, const double* __restrict__ vec, int64_t nv) {
    // x += mean(vec)  (utils.py:1675 quirk Q11: the AVERAGE of the prior means is added to every coordinate)
    __shared__ double s_m;
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int64_t i = 0; i < nv; ++i) s += vec[i];
        s_m = s / (double)nv;
    }
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] += s_m;
}

// colouring factor of utils.py:1668-1670: A = V diag(sqrt(clip(e, 0, inf))), row-major (n, n)
__global__ void color_factor_kernel(const double* __restrict__ evals, const double* __restrict__ evecs, int n,
                                    double* __restrict__ out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n * n) return;
    const double ev = evals[e % n];
    out[e] = evecs[e] * sqrt(ev > 0.0 ? ev : 0.0);
}

}  // namespace tes

using namespace tes;

extern "C" int tes_normal_fill_f64(uint64_t seed, uint32_t iteration, uint32_t stream_id, uint64_t first_elem,
                                   int64_t count, double* out, void* stream) {
    if (count < 0) return -5;
    if (count == 0) return 0;
    if (!out) return -6;
    const uint64_t npairs = ((first_elem + (uint64_t)count + 1) >> 1) - (first_elem >> 1);
    normal_fill_kernel<<<(unsigned)((npairs + 255) / 256), 256, 0, (cudaStream_t)stream>>>(seed, iteration, stream_id,
                                                                                          first_elem, count, out);
    TES_LAUNCH_OK();
    return 0;
}

extern "C" int tes_uniform_fill_f64(uint64_t seed, uint32_t iteration, uint32_t stream_id, uint64_t first_elem,
                                    int64_t count, double* out, void* stream) {
    if (count < 0) return -5;
    if (count == 0) return 0;
    if (!out) return -6;
    const uint64_t npairs = ((first_elem + (uint64_t)count + 1) >> 1) - (first_elem >> 1);
    uniform_fill_kernel<<<(unsigned)((npairs + 255) / 256), 256, 0, (cudaStream_t)stream>>>(seed, iteration, stream_id,
                                                                                           first_elem, count, out);
    TES_LAUNCH_OK();
    return 0;
}

extern "C" int tes_dedup_f64(const double* xs, int64_t S, int64_t d, double resolution, int* mask, int64_t* leader,
                             void* stream) {
    if (S < 0 || S > 49152) return -2;
    if (d < 1 || d > 128) return -3;
    if (S == 0) return 0;
    if (!xs) return -1;
    if (!mask) return -5;
    if (!leader) return -6;
    const size_t rows_b = (size_t)S * d * sizeof(double);
    const int stage = rows_b + (size_t)S * sizeof(int) <= 200 * 1024;
    const size_t smem = (size_t)S * sizeof(int) + (stage ? rows_b : 0);
    if (smem > 48 * 1024)
        TES_CUDA_OK(cudaFuncSetAttribute(dedup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int threads = S >= 1024 ? 1024 : (int)((S + 31) / 32 * 32);
    dedup_kernel<<<1, threads, smem, (cudaStream_t)stream>>>(xs, (int)S, (int)d, resolution, mask, leader, stage);
    TES_LAUNCH_OK();
    return 0;
}

constexpr int EIG_SMEM_MAX_N = 160;

extern "C" int64_t tes_batched_chol_workspace_bytes(int64_t batch, int64_t m, int want_inverse);
extern "C" int tes_batched_potrf_f64(const double* A, int64_t batch, int64_t m, double* L, int* info, void* ws,
                                     int64_t ws_bytes, void* stream);
static int64_t eig_scratch_bytes(int64_t batch, int64_t n) {
    return n <= EIG_SMEM_MAX_N ? 256 : align_up(batch * n * n * (int64_t)sizeof(double), 256);
}
// [Jacobi scratch (n > 160)] [L (batch n n)] [info (batch)] [workspace of the batched Cholesky factorisation]
extern "C" int64_t tes_sym_eig_workspace_bytes(int64_t batch, int64_t n) {
    if (batch < 0 || n < 1 || n > 2048) return -1;
    return eig_scratch_bytes(batch, n) + align_up(batch * n * n * (int64_t)sizeof(double), 256) +
           align_up(batch * (int64_t)sizeof(int), 256) + align_up(tes_batched_chol_workspace_bytes(batch, n, 0), 256);
}

extern "C" int tes_sym_eig_f64(const double* A, int64_t batch, int64_t n, double* evals, double* evecs, int* sweeps,
                               void* ws, int64_t ws_bytes, void* stream) {
    if (batch < 0) return -2;
    if (n < 1 || n > 2048) return -3;
    if (batch == 0) return 0;
    if (!A) return -1;
    if (!evals) return -4;
    if (!evecs) return -5;
    const int in_smem = n <= EIG_SMEM_MAX_N;
    if (!in_smem && (!ws || ws_bytes < eig_scratch_bytes(batch, n))) return -100;
    // Cholesky pre-factor (whenever the caller's workspace has room for it; TES_EIG_PREFACTOR=0 disables): matrices
    // that are not positive definite (info != 0) start from A itself
    const double* Lfac = nullptr;
    const int* linfo = nullptr;
    {
        const char* env = getenv("TES_EIG_PREFACTOR");
        if (ws && ws_bytes >= tes_sym_eig_workspace_bytes(batch, n) && !(env && env[0] == '0') && n >= 2) {
            char* p = (char*)ws + eig_scratch_bytes(batch, n);
            double* L = (double*)p;
            p += align_up(batch * n * n * (int64_t)sizeof(double), 256);
            int* info = (int*)p;
            p += align_up(batch * (int64_t)sizeof(int), 256);
            const int64_t cw = tes_batched_chol_workspace_bytes(batch, n, 0);
            const int rc = tes_batched_potrf_f64(A, batch, n, L, info, p, cw, stream);
            if (rc) return rc;
            Lfac = L;
            linfo = info;
        }
    }
    const int64_t npair = (n + 1) / 2;
    const size_t smem = in_smem ? (size_t)n * (n + 1) * sizeof(double) : 0;
    if (smem > 48 * 1024)
        TES_CUDA_OK(cudaFuncSetAttribute(sym_eig_kernel<true, 1024, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // four pairs per warp at a time (eight lanes each); n = 64: 256 threads, several matrices per SM hide each other's
    // barrier and fp64 latency
    int threads = 32 * (int)((npair + 3) / 4);
    threads = threads < 128 ? 128 : (threads > 1024 ? 1024 : threads);
    if (in_smem && n <= 64)     // <= 32 pairs: 256 threads, a lane's elements of both columns in registers, 5 CTAs per SM
        sym_eig_kernel<true, 256, 3><<<(unsigned)batch, threads, smem, (cudaStream_t)stream>>>(A, (int)n, evals, evecs, sweeps,
                                                                                              (double*)ws, Lfac, linfo);
    else if (in_smem)
        sym_eig_kernel<true, 1024, 1><<<(unsigned)batch, threads, smem, (cudaStream_t)stream>>>(A, (int)n, evals, evecs, sweeps,
                                                                                               (double*)ws, Lfac, linfo);
    else
        sym_eig_kernel<false, 1024, 1><<<(unsigned)batch, threads, smem, (cudaStream_t)stream>>>(A, (int)n, evals, evecs, sweeps,
                                                                                                (double*)ws, Lfac, linfo);
    TES_LAUNCH_OK();
    return 0;
}

extern "C" int tes_color_factor_f64(const double* evals, const double* evecs, int64_t n, double* out, void* stream) {
    if (n < 1 || n > 2048) return -3;
    if (!evals) return -1;
    if (!evecs) return -2;
    if (!out) return -4;
    color_factor_kernel<<<(unsigned)((n * n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(evals, evecs, (int)n, out);
    TES_LAUNCH_OK();
    return 0;
}

extern "C" int64_t tes_joint_maxidx_workspace_bytes(int64_t ns, int64_t T) {
    if (ns < 1 || T < 1) return -1;
    return align_up(T * 8, 256) + align_up((int64_t)GRAM_SPLIT * T * T * 8, 256);
}

// samples (ns, T) row-major, in place: += mean(prior_mean) (Q11); then the correlated-dimension mask, the arg-max of
// every row and the bucket counts.  masked (T) int32, maxidx (ns) int64, counts (T) int32.
extern "C" int tes_joint_maxidx_f64(double* samples, int64_t ns, int64_t T, const double* prior_mean,
                                    int remove_correlated, int* masked, int64_t* maxidx, int* counts, void* ws,
                                    int64_t ws_bytes, void* stream) {
    if (ns < 1) return -2;
    if (T < 1 || T > (1 << 20)) return -3;
    if (!samples) return -1;
    if (!masked) return -6;
    if (!maxidx) return -7;
    if (!counts) return -8;
    if (!ws || ws_bytes < tes_joint_maxidx_workspace_bytes(ns, T)) return -100;
    cudaStream_t st = (cudaStream_t)stream;
    double* cmean = (double*)ws;
    double* gram = (double*)((char*)ws + align_up(T * 8, 256));
    if (prior_mean) {
        add_scalar_kernel<<<(unsigned)((ns * T + 255) / 256), 256, 0, st>>>(samples, ns * T, prior_mean, T);
        TES_LAUNCH_OK();
    }
    TES_CUDA_OK(cudaMemsetAsync(masked, 0, T * sizeof(int), st));
    TES_CUDA_OK(cudaMemsetAsync(counts, 0, T * sizeof(int), st));
    if (remove_correlated && ns > 1) {
        col_mean_kernel<<<(unsigned)((T + 31) / 32), dim3(32, 32), 0, st>>>(samples, ns, (int)T, cmean);
        TES_LAUNCH_OK();
        gram_kernel<<<dim3((unsigned)((T + GRAM_AT - 1) / GRAM_AT), GRAM_SPLIT), 256, 0, st>>>(samples, ns, (int)T,
                                                                                                 cmean, gram);
        TES_LAUNCH_OK();
        corr_flag_kernel<<<(unsigned)((T + 127) / 128), 128, 0, st>>>(gram, ns, (int)T, 1.0 - 1e-6, masked);
        TES_LAUNCH_OK();
    }
    row_argmax_kernel<<<(unsigned)((ns + 7) / 8), 256, 0, st>>>(samples, ns, (int)T, masked, maxidx, counts);
    TES_LAUNCH_OK();
    return 0;
}

// Bucket statistics.  samples (ns, T); order (ns) = sample rows sorted (stably) by bucket; starts (nb + 1) = bucket
// boundaries in `order`; cols (nk) = kept test-point columns.  mean (nb, nk), cov (nb, nk, nk).
extern "C" int tes_bucket_moments_f64(const double* samples, int64_t ns, int64_t T, const int64_t* order,
                                      const int64_t* starts, int64_t nb, const int64_t* cols, int64_t nk,
                                      double* mean, double* cov, void* stream) {
    if (ns < 1) return -2;
    if (T < 1) return -3;
    if (nb < 0) return -6;
    if (nk < 1) return -8;
    if (nb == 0) return 0;
    if (!samples || !order || !starts || !cols) return -1;
    if (!mean) return -9;
    if (!cov) return -10;
    cudaStream_t st = (cudaStream_t)stream;
    bucket_mean_kernel<<<(unsigned)nb, 256, 0, st>>>(samples, (int)T, order, starts, cols, (int)nk, mean);
    TES_LAUNCH_OK();
    bucket_cov_kernel<<<dim3((unsigned)nb, (unsigned)((nk + BCOV_AT - 1) / BCOV_AT)), 128, 0, st>>>(
        samples, (int)T, order, starts, cols, (int)nk, mean, cov);
    TES_LAUNCH_OK();
    return 0;
}

// Padded per-maximizer sample tensor of mode 'sample': out (nb, nk, K), mask (nb, K) bytes.
extern "C" int tes_bucket_samples_f64(const double* samples, int64_t ns, int64_t T, const int64_t* order,
                                      const int64_t* starts, int64_t nb, const int64_t* cols, int64_t nk, int64_t K,
                                      const double* prior_mean, double* out, uint8_t* mask, void* stream) {
    if (ns < 1) return -2;
    if (T < 1) return -3;
    if (nb < 0) return -6;
    if (nk < 1 || nk > 65535) return -8;
    if (K < 1) return -9;
    if (nb == 0) return 0;
    if (!samples || !order || !starts || !cols || !prior_mean) return -1;
    if (!out) return -11;
    if (!mask) return -12;
    bucket_samples_kernel<<<dim3((unsigned)nb, (unsigned)nk), 256, 0, (cudaStream_t)stream>>>(
        samples, (int)T, order, starts, cols, (int)nk, (int)K, prior_mean, out, mask);
    TES_LAUNCH_OK();
    return 0;
}
