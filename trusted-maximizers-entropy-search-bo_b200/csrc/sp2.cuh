// sp2.cuh -- TES_sp Monte-Carlo estimator, whitened formulation (criteria/evaluate_sample_mp.py:246-336 for q = 1,
// criteria/evaluate_batch_sample_mp.py:294-382 for batch q).
//
// The reference evaluates, for every source draw y = mu_{j,s} + L z_s, the q-variate log-density under every
// component (j', same s):  -1/2 |L^-1 (y - mu_{j',s})|^2 - sum log L_ii - q/2 log 2pi.
// With nu_{j,s} = L^-1 mu_{j,s} (one triangular product per joint sample and query, shared by all I*Y draws):
//     L^-1 (y - mu_{j',s}) = nu_{j,s} + z_s - nu_{j',s}
// so one log-density costs q subtractions + q FMAs instead of a q x q triangular solve (q = 1: this is TFP's own
// (x/s - mu/s)^2 form), y is never formed and L is not needed in the hot loop.
//
// Log-sum-exp over the joint samples s: every log-density is <= cnorm = -sum log L_ii - q/2 log 2pi, and the
// (j' = j, same s) term has |z_s|^2 ~ chi^2_q, so the shift cnorm is always safe: LSE = cnorm + log sum exp(-u2/2).
// A component whose terms all underflow yields -inf where the max-shifted reference yields ~ -1e3: either way it adds
// exp(.) = 0 to the outer LSE over j' (which always contains the finite j' = j term).  The outer LSE keeps the
// reference's rule (shift = max if finite else 0, NaN propagates).
//
// exp on (-700, 0] is table driven: x = (16 n + i) ln2/16 + r, |r| <= ln2/32, e^x = 2^n * 2^(i/16) * P5(r) with a
// 16-entry table (128 B: one entry per shared-memory bank pair, conflict-free for any access pattern) and a degree-5
// Taylor polynomial (truncation 1.4e-13 relative) -- 10 fp64-pipe instructions instead of ~25 for exp().
//
// Work decomposition: one CTA per query (grid-stride), nu[q][s][j'] in shared memory (j' fastest: lanes read
// consecutive words).  A warp owns work items (iteration, source maximizer j, block of R*NG y-draws): lanes are split
// into NG groups of G lanes (G = power of two >= min(Sp, 32)); lane (g, jl) accumulates sum_s exp(.) for component
// j' = jl (+ G per round) and the R draws iy = (block*R + k)*NG + g.  nu is loaded once per (s, j') and reused for the R
// draws held in registers.  Only the valid joint samples of the SOURCE j are visited (Q3) and only their normals are
// generated (counter-based: two normals per Philox call when q is even).
#pragma once
#include "tes_common.cuh"

namespace tes {

constexpr int SP2_R = 4;        // y-draws per lane held in registers
constexpr int SP2_CHUNK = 16;   // (valid joint samples) x NG per a-buffer refill

__device__ __forceinline__ double exp_neg_fast(double x, const double* __restrict__ tab) {
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

struct Sp2Smem {  // offsets in doubles (computed identically on host and device)
    int lq, logp, tab, acc_w, a_w, nu, vlist;
    int SpS, rounds, G, NG, CB;
    size_t bytes;
};

__host__ __device__ inline Sp2Smem sp2_layout(int Q, int Sp, int K, int nwarps) {
    Sp2Smem L;
    L.G = 32;
    while (L.G > 4 && (L.G >> 1) >= Sp) L.G >>= 1;
    L.NG = 32 / L.G;
    L.CB = SP2_CHUNK / L.NG;
    L.rounds = (Sp + L.G - 1) / L.G;
    L.SpS = Sp | 1;
    int o = 16;                        // red[16]
    L.lq = o;      o += Q * Q + Q + 2; // Linv, b, cnorm
    o = (o + 1) & ~1;
    L.logp = o;    o += 2 * Sp;        // log p, p
    o = (o + 1) & ~1;
    L.tab = o;     o += 16;
    L.acc_w = o;   o += nwarps * L.rounds * 32 * SP2_R;
    L.a_w = o;     o += nwarps * SP2_R * L.NG * L.CB * Q;
    L.nu = o;      o += Q * K * L.SpS;
    o = (o + 1) & ~1;
    L.vlist = o;                       // uint16 [Sp][K] + int nvalid[Sp]
    L.bytes = (size_t)o * 8 + (((size_t)Sp * K * 2 + 7) & ~(size_t)7) + (size_t)Sp * 4 + 8;
    return L;
}

template <int Q>
__global__ void __launch_bounds__(512)
    sp2_kernel(const double* __restrict__ Mq,    // (rows*Q, T)   k invpNK[:, :T]
               const double* __restrict__ bq,    // (rows*Q)      k invpNK[:, T:] Y
               const double* __restrict__ Linv,  // (rows, Q, Q)  inverse of chol(cov_y), lower
               const double* __restrict__ Lmat,  // (rows, Q, Q)  chol(cov_y) (diagonal -> log-determinant)
               const double* __restrict__ P,     // (Sp, T, K)
               const unsigned char* __restrict__ mask,  // (Sp, K)
               const double* __restrict__ p, int64_t rows, int Sp, int T, int K, int niter, int ny,
               const double* __restrict__ z, int64_t z_n, int64_t z_x0, uint64_t seed, int64_t n_global,
               int64_t x_global0, double* __restrict__ out) {
    // n_global < 0: SHARED draws (every query consumes the same normals; common random numbers)
    const int64_t kg_n = n_global < 0 ? 1 : n_global, kg_x0 = n_global < 0 ? 0 : x_global0, kg_rmul = n_global < 0 ? 0 : 1;
    extern __shared__ __align__(16) double sh[];
    const int nwarps = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const Sp2Smem L = sp2_layout(Q, Sp, K, nwarps);
    double* red = sh;
    double* lq = sh + L.lq;            // [Q*Q] Linv, [Q] b, [1] cnorm
    double* logp = sh + L.logp;        // [Sp] log p, [Sp] p
    double* tab = sh + L.tab;
    double* accw = sh + L.acc_w + (size_t)warp * L.rounds * 32 * SP2_R;
    double* aw = sh + L.a_w + (size_t)warp * SP2_R * L.NG * L.CB * Q;
    double* nu = sh + L.nu;            // [(q*K + s)*SpS + j]
    unsigned short* vlist = reinterpret_cast<unsigned short*>(sh + L.vlist);
    int* nvalid = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(vlist) + (((size_t)Sp * K * 2 + 7) & ~(size_t)7));
    const int G = L.G, NG = L.NG, CB = L.CB, SpS = L.SpS, rounds = L.rounds;
    const int g = lane / G, jl = lane % G;
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (g * G));

    // ---- per-CTA constants: table, log p, compacted lists of valid joint samples per source maximizer
    if (threadIdx.x < 16) {
        const double t16[16] = {0x1.0000000000000p+0, 0x1.0b5586cf9890fp+0, 0x1.172b83c7d517bp+0, 0x1.2387a6e756238p+0,
                                0x1.306fe0a31b715p+0, 0x1.3dea64c123422p+0, 0x1.4bfdad5362a27p+0, 0x1.5ab07dd485429p+0,
                                0x1.6a09e667f3bcdp+0, 0x1.7a11473eb0187p+0, 0x1.8ace5422aa0dbp+0, 0x1.9c49182a3f090p+0,
                                0x1.ae89f995ad3adp+0, 0x1.c199bdd85529cp+0, 0x1.d5818dcfba487p+0, 0x1.ea4afa2a490dap+0};
        tab[threadIdx.x] = t16[threadIdx.x];
    }
    for (int j = threadIdx.x; j < Sp; j += blockDim.x) {
        logp[j] = log(p[j]);
        logp[Sp + j] = p[j];
        int c = 0;
        for (int s = 0; s < K; ++s)
            if (mask[(size_t)j * K + s]) vlist[(size_t)j * K + c++] = (unsigned short)s;
        nvalid[j] = c;
    }

    const int iyb = SP2_R * NG;                         // y-draws per work item
    const int nblk = (ny + iyb - 1) / iyb;
    const int64_t items = (int64_t)niter * Sp * nblk;

    for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
        __syncthreads();
        if (threadIdx.x < Q * Q) lq[threadIdx.x] = Linv[row * Q * Q + threadIdx.x];
        if (threadIdx.x < Q) lq[Q * Q + threadIdx.x] = bq[row * Q + threadIdx.x];
        if (threadIdx.x == 32) {
            double ld = 0.0;
            for (int a = 0; a < Q; ++a) ld += log(Lmat[row * Q * Q + a * Q + a]);
            lq[Q * Q + Q] = -ld - 0.5 * Q * 1.8378770664093454836;  // - sum log L_ii - q/2 log(2 pi)
        }
        __syncthreads();
        // nu[q][s][j] = (Linv (M P_j[:, s] + b))_q
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
            for (int q = 0; q < Q; ++q) a[q] += lq[Q * Q + q];
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                double w = 0.0;
#pragma unroll
                for (int b2 = 0; b2 <= q; ++b2) w = fma(lq[q * Q + b2], a[b2], w);
                nu[((size_t)q * K + s) * SpS + j] = w;
            }
        }
        __syncthreads();
        const double cnorm = lq[Q * Q + Q];
        double acc = 0.0;
        for (int64_t item = warp; item < items; item += nwarps) {
            const int blk = (int)(item % nblk);
            const int j = (int)((item / nblk) % Sp);
            const int it = (int)(item / ((int64_t)nblk * Sp));
            const int nv = nvalid[j];
            const unsigned short* vl = vlist + (size_t)j * K;
            for (int e = lane; e < rounds * 32 * SP2_R; e += 32) accw[e] = 0.0;
            for (int c0 = 0; c0 < nv; c0 += CB) {
                const int cn = (nv - c0 < CB) ? (nv - c0) : CB;
                __syncwarp();
                // a[il][c][q] = nu[q][s][j] + z  for the iyb draws of this item and the cn samples of this chunk
                if (Q % 2 == 0 && !z) {
                    const int ntask = iyb * cn * (Q / 2);
                    for (int tk = lane; tk < ntask; tk += 32) {
                        const int qp = tk % (Q / 2);
                        const int c = (tk / (Q / 2)) % cn;
                        const int il = tk / ((Q / 2) * cn);
                        const int iy = blk * iyb + il;
                        const int s = vl[c0 + c];
                        double z0 = 0.0, z1 = 0.0;
                        if (iy < ny) {
                            const uint64_t e0 = (uint64_t)((((int64_t)iy * Sp + j) * kg_n + kg_x0 + row * kg_rmul) * K + s) * Q + 2 * qp;
                            normal_pair(seed, (uint32_t)it, TES_STREAM_BSP_Y, e0 >> 1, z0, z1);
                        }
                        double* dst = aw + ((size_t)il * CB + c) * Q + 2 * qp;
                        dst[0] = nu[((size_t)(2 * qp) * K + s) * SpS + j] + z0;
                        dst[1] = nu[((size_t)(2 * qp + 1) * K + s) * SpS + j] + z1;
                    }
                } else {
                    const int ntask = iyb * cn * Q;
                    for (int tk = lane; tk < ntask; tk += 32) {
                        const int q = tk % Q;
                        const int c = (tk / Q) % cn;
                        const int il = tk / (Q * cn);
                        const int iy = blk * iyb + il;
                        const int s = vl[c0 + c];
                        double zz = 0.0;
                        if (iy < ny) {
                            if (z)
                                zz = z[(((((int64_t)it * ny + iy) * Sp + j) * z_n + z_x0 + row) * K + s) * Q + q];
                            else
                                zz = normal_elem(seed, (uint32_t)it, Q == 1 ? TES_STREAM_SP_Y : TES_STREAM_BSP_Y,
                                                 (uint64_t)((((int64_t)iy * Sp + j) * kg_n + kg_x0 + row * kg_rmul) * K + s) * Q + q);
                        }
                        aw[((size_t)il * CB + c) * Q + q] = nu[((size_t)q * K + s) * SpS + j] + zz;
                    }
                }
                __syncwarp();
                for (int rd = 0; rd < rounds; ++rd) {
                    int jp = rd * G + jl;
                    jp = jp < Sp ? jp : Sp - 1;  // clamped lanes compute a duplicate that is never used
                    double sum[SP2_R];
#pragma unroll
                    for (int k = 0; k < SP2_R; ++k) sum[k] = accw[(rd * 32 + lane) * SP2_R + k];
                    for (int c = 0; c < cn; ++c) {
                        const int s = vl[c0 + c];
                        double nv_[Q];
#pragma unroll
                        for (int q = 0; q < Q; ++q) nv_[q] = nu[((size_t)q * K + s) * SpS + jp];
#pragma unroll
                        for (int k = 0; k < SP2_R; ++k) {
                            const double* ak = aw + ((size_t)(k * NG + g) * CB + c) * Q;
                            double u2 = 0.0;
#pragma unroll
                            for (int q = 0; q < Q; ++q) {
                                const double dlt = ak[q] - nv_[q];
                                u2 = fma(dlt, dlt, u2);
                            }
                            sum[k] += exp_neg_fast(-0.5 * u2, tab);
                        }
                    }
#pragma unroll
                    for (int k = 0; k < SP2_R; ++k) accw[(rd * 32 + lane) * SP2_R + k] = sum[k];
                }
            }
            __syncwarp();
            // ---- finish the item: LSE over s -> (iy, j, j'), then the LSE over j' with the reference's rules
#pragma unroll
            for (int k = 0; k < SP2_R; ++k) {
                const int iy = blk * iyb + k * NG + g;
                double mx = neg_inf(), lmix = neg_inf();
                bool anynan = false;
                double term_r[4];  // rounds <= 4 (host-checked)
                for (int rd = 0; rd < rounds; ++rd) {
                    const int jp = rd * G + jl;
                    double term = neg_inf();
                    if (jp < Sp) {
                        const double lse = cnorm + log(accw[(rd * 32 + lane) * SP2_R + k]);
                        if (jp == j && lse == lse) lmix = lse;
                        term = lse + logp[jp];
                        if (term != term) anynan = true;
                        else mx = term > mx ? term : mx;
                    }
                    term_r[rd] = term;
                }
                for (int off = G >> 1; off > 0; off >>= 1) {
                    const double o = __shfl_xor_sync(0xffffffffu, mx, off);
                    mx = o > mx ? o : mx;
                    const double o2 = __shfl_xor_sync(0xffffffffu, lmix, off);
                    lmix = o2 > lmix ? o2 : lmix;
                }
                anynan = (__ballot_sync(0xffffffffu, anynan) & gmask) != 0u;
                // lmix may be NaN (sum NaN): carry it separately, max-reduce drops NaN
                bool mixnan = false;
                for (int rd = 0; rd < rounds; ++rd)
                    if (rd * G + jl == j && jl + rd * G < Sp && term_r[rd] != term_r[rd]) mixnan = true;
                mixnan = (__ballot_sync(0xffffffffu, mixnan) & gmask) != 0u;
                const bool fin = mx > neg_inf() && mx < pos_inf();
                double se = 0.0;
                for (int rd = 0; rd < rounds; ++rd) {
                    const double term = term_r[rd];
                    if (term == term && term > neg_inf() && fin) se += exp(term - mx);
                }
                for (int off = G >> 1; off > 0; off >>= 1) se += __shfl_xor_sync(0xffffffffu, se, off);
                double lmarg;
                if (anynan) lmarg = __longlong_as_double(0x7ff8000000000000LL);
                else if (fin) lmarg = log(se) + mx;
                else lmarg = mx;
                if (mixnan) lmix = __longlong_as_double(0x7ff8000000000000LL);
                if (jl == 0 && iy < ny) acc += logp[Sp + j] * (lmix - lmarg);
            }
        }
        acc = warp_sum(acc);
        __syncthreads();
        if (lane == 0) red[warp] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
            for (int w = 0; w < nwarps; ++w) s += red[w];
            out[row] = s / ((double)niter * (double)ny);
        }
    }
}

}  // namespace tes
