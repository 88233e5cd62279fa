// rffopt.cu -- gradient refinement of the trusted maximizers on the device.
//
// Reference: utils_for_continuous.sample_xmaxs_fmaxs, second half (utils_for_continuous.py:201-232): the top-k
// initialisers of every function sample become a TF variable xs (S, k, d) with a clip-to-[xmin, xmax] constraint;
// tf.train.AdamOptimizer() (TF-1 defaults lr = 1e-3, beta1 = 0.9, beta2 = 0.999, epsilon = 1e-8) minimises
// -sum_{j,i} f_j(xs[j,i]) for `ntrain` sess.run(train) calls (one Adam step each); then max_f = max_i f_j(xs[j,i]),
// max_x = the arg-max row (first maximum).  The objective is separable, so every (sample, start point) is an
// independent Adam trajectory:
//   f_j(x) = scale sum_f theta_f cos(w_f.x + b_f),   d(-f_j)/dx = scale sum_f theta_f sin(w_f.x + b_f) w_f,
//   t += 1; lr_t = lr sqrt(1 - beta2^t) / (1 - beta1^t); m = beta1 m + (1-beta1) g; v = beta2 v + (1-beta2) g^2;
//   x = clip(x - lr_t m / (sqrt(v) + epsilon)).
// One warp per trajectory: the lanes split the features, the gradient is a warp all-reduce, every lane carries the
// (identical) Adam state in registers.  All `ntrain` steps run inside ONE launch (the reference pays a sess.run
// round trip per step).
#include "tes_common.cuh"

namespace tes {

constexpr int OPT_MAX_D = 16;
constexpr int OPT_WARPS = 4;

template <int D>
__global__ void __launch_bounds__(32 * OPT_WARPS)
    rff_adam_kernel(const double* __restrict__ x0, const double* __restrict__ W, const double* __restrict__ b,
                    const double* __restrict__ theta, int64_t S, int64_t F, int k, int w_shared, double scale,
                    const double* __restrict__ xmin, const double* __restrict__ xmax, int ntrain, double lr, double beta1,
                    double beta2, double eps, double* __restrict__ out_x, double* __restrict__ out_f) {
    const int lane = threadIdx.x & 31;
    const int64_t traj = (int64_t)blockIdx.x * OPT_WARPS + (threadIdx.x >> 5);
    if (traj >= S * k) return;
    const int64_t j = traj / k;
    const int64_t jw = w_shared ? 0 : j;
    const double* Wj = W + jw * F * D;
    const double* bj = b + jw * F;
    const double* tj = theta + j * F;
    double x[D], m[D], v[D], lo[D], hi[D];
#pragma unroll
    for (int c = 0; c < D; ++c) {
        x[c] = x0[traj * D + c];
        m[c] = 0.0;
        v[c] = 0.0;
        lo[c] = xmin[c];
        hi[c] = xmax[c];
    }
    double b1t = 1.0, b2t = 1.0;
    for (int step = 0; step < ntrain; ++step) {
        double g[D];
#pragma unroll
        for (int c = 0; c < D; ++c) g[c] = 0.0;
        for (int64_t f = lane; f < F; f += 32) {
            double w[D];
            double h = bj[f];
#pragma unroll
            for (int c = 0; c < D; ++c) {
                w[c] = Wj[f * D + c];
                h = fma(w[c], x[c], h);
            }
            const double t = tj[f] * sin(h);
#pragma unroll
            for (int c = 0; c < D; ++c) g[c] = fma(t, w[c], g[c]);
        }
        b1t *= beta1;
        b2t *= beta2;
        const double lr_t = lr * sqrt(1.0 - b2t) / (1.0 - b1t);
#pragma unroll
        for (int c = 0; c < D; ++c) {
            const double gc = scale * warp_sum(g[c]);   // butterfly all-reduce: identical in every lane
            m[c] = beta1 * m[c] + (1.0 - beta1) * gc;
            v[c] = beta2 * v[c] + (1.0 - beta2) * gc * gc;
            double xn = x[c] - lr_t * m[c] / (sqrt(v[c]) + eps);
            xn = fmin(fmax(xn, lo[c]), hi[c]);          // the variable's clip_by_value constraint
            x[c] = xn;
        }
    }
    double acc = 0.0;
    for (int64_t f = lane; f < F; f += 32) {
        double h = bj[f];
#pragma unroll
        for (int c = 0; c < D; ++c) h = fma(Wj[f * D + c], x[c], h);
        acc = fma(tj[f], cos(h), acc);
    }
    acc = scale * warp_sum(acc);
    if (lane == 0) {
        out_f[traj] = acc;
#pragma unroll
        for (int c = 0; c < D; ++c) out_x[traj * D + c] = x[c];
    }
}

}  // namespace tes

using namespace tes;

// x0 (S, k, d) start points (the top-k initialisers), W (S or 1, F, d), b (S or 1, F), theta (S, F), xmin / xmax (d)
// device arrays.  out_x (S, k, d) refined points, out_f (S, k) their function-sample values.
extern "C" int tes_rff_adam_refine_f64(const double* x0, int64_t S, int k, int64_t d, const double* W, const double* b,
                                       const double* theta, int64_t F, int w_shared, double sigma, const double* xmin,
                                       const double* xmax, int ntrain, double lr, double beta1, double beta2,
                                       double epsilon, double* out_x, double* out_f, void* stream) {
    if (!x0) return -1;
    if (S < 1) return -2;
    if (k < 1) return -3;
    if (d < 1 || d > OPT_MAX_D) return -4;
    if (!W) return -5;
    if (!b) return -6;
    if (!theta) return -7;
    if (F < 1) return -8;
    if (!xmin) return -11;
    if (!xmax) return -12;
    if (ntrain < 0) return -13;
    if (!out_x) return -18;
    if (!out_f) return -19;
    const double scale = sqrt(2.0 * sigma / (double)F);
    const unsigned grid = (unsigned)((S * k + OPT_WARPS - 1) / OPT_WARPS);
    cudaStream_t st = (cudaStream_t)stream;
    switch (d) {
#define TES_CASE(DD)                                                                                                  \
    case DD:                                                                                                          \
        rff_adam_kernel<DD><<<grid, 32 * OPT_WARPS, 0, st>>>(x0, W, b, theta, S, F, k, w_shared, scale, xmin, xmax,    \
                                                            ntrain, lr, beta1, beta2, epsilon, out_x, out_f);         \
        break;
        TES_CASE(1) TES_CASE(2) TES_CASE(3) TES_CASE(4) TES_CASE(5) TES_CASE(6) TES_CASE(7) TES_CASE(8) TES_CASE(9)
        TES_CASE(10) TES_CASE(11) TES_CASE(12) TES_CASE(13) TES_CASE(14) TES_CASE(15) TES_CASE(16)
#undef TES_CASE
    }
    TES_LAUNCH_OK();
    return 0;
}
