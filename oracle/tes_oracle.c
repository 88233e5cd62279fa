/*
 * tes_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of Stage A of the reference's TES hot path -- evaluating the S
 * random-Fourier-feature posterior function samples on all candidates and taking each sample's
 * top-k / arg-max (utils_for_continuous.py:172-187; numpy twin optfunc.py:389-402) -- following
 * the operation sequence fixed in include/tes_spec.h, plus the counter-based normal generator
 * the stochastic criteria use at scale.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product never does.
 *
 * Pinning: tests/test_oracle_golden.py checks these functions against the reference's own
 * optfunc.make_function_sample_np outputs (tests/golden/rff.npz): values to 1e-12 absolute and
 * identical arg-max indices.
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -ffp-contract=off -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/tes_spec.h"

#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
#define TES_CLONES __attribute__((target_clones("fma", "default")))
#else
#define TES_CLONES
#endif

static inline double cospi_spec(double h) {
    double n = rint(h); /* round-to-nearest-even (default rounding mode) */
    double r = h - n;
    double s = r * r;
    double p = TES_COSPI_C8;
    p = __builtin_fma(p, s, TES_COSPI_C7);
    p = __builtin_fma(p, s, TES_COSPI_C6);
    p = __builtin_fma(p, s, TES_COSPI_C5);
    p = __builtin_fma(p, s, TES_COSPI_C4);
    p = __builtin_fma(p, s, TES_COSPI_C3);
    p = __builtin_fma(p, s, TES_COSPI_C2);
    p = __builtin_fma(p, s, TES_COSPI_C1);
    p = __builtin_fma(p, s, TES_COSPI_C0);
    /* n odd?  |h| < 2^51 assumed (else every double is an even integer anyway) */
    long long ni = (long long)n;
    return (ni & 1) ? -p : p;
}

double tes_oracle_cospi(double h) { return cospi_spec(h); }

/* (value desc, index asc) insertion into a sorted list of length k */
static inline void topk_insert(double* vals, int64_t* idxs, int k, double v, int64_t i) {
    int pos = k;
    while (pos > 0 && (v > vals[pos - 1] || (v == vals[pos - 1] && i < idxs[pos - 1]))) pos--;
    if (pos >= k) return;
    for (int t = k - 1; t > pos; --t) {
        vals[t] = vals[t - 1];
        idxs[t] = idxs[t - 1];
    }
    vals[pos] = v;
    idxs[pos] = i;
}

/* f[j*N + i] for all samples j and candidates i.  W: (S or 1, F, d), b: (S or 1, F), theta: (S, F). */
TES_CLONES
int tes_oracle_rff_eval(const double* cand, int64_t N, int64_t d, const double* W, const double* b,
                        const double* theta, int64_t S, int64_t F, int w_shared, double sigma,
                        double* out) {
    if (d > 64) return -3;
    const double scale = sqrt(2.0 * sigma / (double)F);
    const int64_t SW = w_shared ? 1 : S;
    double* wp = (double*)malloc(sizeof(double) * (size_t)(SW * F * d));
    double* bp = (double*)malloc(sizeof(double) * (size_t)(SW * F));
    if (!wp || !bp) return -100;
    for (int64_t t = 0; t < SW * F * d; ++t) wp[t] = W[t] * TES_INV_PI;
    for (int64_t t = 0; t < SW * F; ++t) bp[t] = b[t] * TES_INV_PI;
#pragma omp parallel for schedule(static) collapse(2)
    for (int64_t j = 0; j < S; ++j) {
        for (int64_t i = 0; i < N; ++i) {
            const double* wj = wp + (w_shared ? 0 : j) * F * d;
            const double* bj = bp + (w_shared ? 0 : j) * F;
            const double* tj = theta + j * F;
            const double* x = cand + i * d;
            double acc = 0.0;
            for (int64_t f = 0; f < F; ++f) {
                double h = bj[f];
                for (int64_t k = 0; k < d; ++k) h = __builtin_fma(wj[f * d + k], x[k], h);
                acc = __builtin_fma(tj[f], cospi_spec(h), acc);
            }
            out[j * N + i] = scale * acc;
        }
    }
    free(wp);
    free(bp);
    return 0;
}

/* top-k (value desc, index asc) per sample; idx_offset is added to the reported indices.
 * Slots beyond N are (-inf, -1). */
TES_CLONES
int tes_oracle_rff_topk(const double* cand, int64_t N, int64_t d, const double* W, const double* b,
                        const double* theta, int64_t S, int64_t F, int w_shared, double sigma, int k,
                        int64_t idx_offset, int64_t* out_idx, double* out_val) {
    if (d > 64) return -3;
    if (k < 1 || k > 64) return -11;
    const double scale = sqrt(2.0 * sigma / (double)F);
    const int64_t SW = w_shared ? 1 : S;
    double* wp = (double*)malloc(sizeof(double) * (size_t)(SW * F * d));
    double* bp = (double*)malloc(sizeof(double) * (size_t)(SW * F));
    if (!wp || !bp) return -100;
    for (int64_t t = 0; t < SW * F * d; ++t) wp[t] = W[t] * TES_INV_PI;
    for (int64_t t = 0; t < SW * F; ++t) bp[t] = b[t] * TES_INV_PI;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t j = 0; j < S; ++j) {
        const double* wj = wp + (w_shared ? 0 : j) * F * d;
        const double* bj = bp + (w_shared ? 0 : j) * F;
        const double* tj = theta + j * F;
        double vals[64];
        int64_t idxs[64];
        for (int t = 0; t < k; ++t) {
            vals[t] = -INFINITY;
            idxs[t] = -1;
        }
        for (int64_t i = 0; i < N; ++i) {
            const double* x = cand + i * d;
            double acc = 0.0;
            for (int64_t f = 0; f < F; ++f) {
                double h = bj[f];
                for (int64_t kk = 0; kk < d; ++kk) h = __builtin_fma(wj[f * d + kk], x[kk], h);
                acc = __builtin_fma(tj[f], cospi_spec(h), acc);
            }
            double v = scale * acc;
            if (v > vals[k - 1]) topk_insert(vals, idxs, k, v, i + idx_offset);
        }
        for (int t = 0; t < k; ++t) {
            out_val[j * k + t] = vals[t];
            out_idx[j * k + t] = idxs[t];
        }
    }
    free(wp);
    free(bp);
    return 0;
}

/* ---- Philox4x32-10 + Box-Muller, per include/tes_spec.h --------------------------------------*/
static inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)TES_PHILOX_M0 * c[0];
        uint64_t p1 = (uint64_t)TES_PHILOX_M1 * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += TES_PHILOX_W0;
        k1 += TES_PHILOX_W1;
    }
}

static inline void normal_pair(uint64_t seed, uint32_t iteration, uint32_t stream, uint64_t pair,
                               double* z0, double* z1) {
    uint32_t c[4] = {(uint32_t)pair, (uint32_t)(pair >> 32), iteration, stream};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    uint64_t a = ((uint64_t)c[1] << 32) | c[0];
    uint64_t b = ((uint64_t)c[3] << 32) | c[2];
    double u1 = ((double)(a >> 11) + 0.5) * 0x1.0p-53;
    double u2 = ((double)(b >> 11) + 0.5) * 0x1.0p-53;
    double rad = sqrt(-2.0 * log(u1));
    double ang = 6.283185307179586476925 * u2;
    *z0 = rad * cos(ang);
    *z1 = rad * sin(ang);
}

/* out[t] = z(first_elem + t), t in [0,count) */
int tes_oracle_normals(uint64_t seed, uint32_t iteration, uint32_t stream, uint64_t first_elem,
                       int64_t count, double* out) {
#pragma omp parallel for schedule(static)
    for (int64_t t = 0; t < count; ++t) {
        uint64_t e = first_elem + (uint64_t)t;
        double z0, z1;
        normal_pair(seed, iteration, stream, e >> 1, &z0, &z1);
        out[t] = (e & 1) ? z1 : z0;
    }
    return 0;
}

int tes_oracle_philox_raw(uint64_t seed, uint32_t iteration, uint32_t stream, uint64_t pair,
                          uint32_t out[4]) {
    uint32_t c[4] = {(uint32_t)pair, (uint32_t)(pair >> 32), iteration, stream};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    memcpy(out, c, sizeof(c));
    return 0;
}
