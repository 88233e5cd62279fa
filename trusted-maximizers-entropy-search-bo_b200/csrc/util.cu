// util.cu -- ABI version, error strings and the fp64-FMA burn used as the roofline denominator.
#include "tes_common.cuh"

#include <atomic>
static std::atomic<long long> g_launches{0};
extern "C" void tes_count_launch_(void) { g_launches.fetch_add(1, std::memory_order_relaxed); }
extern "C" int64_t tes_launch_count(void) { return (int64_t)g_launches.load(std::memory_order_relaxed); }

extern "C" int tes_abi_version(void) { return TES_B200_ABI_VERSION; }

extern "C" const char* tes_error_string(int code) {
    if (code == 0) return "ok";
    if (code == -100) return "workspace missing or too small";
    if (code < 0) return "invalid argument (position = -code)";
    return cudaGetErrorString((cudaError_t)code);
}

namespace tes {
// 8 independent DFMA chains per thread, 256 threads per CTA: enough ILP x TLP to saturate the
// 64-lane/SM fp64 pipe.  flops = blocks*256*iters*8*2.
__global__ void __launch_bounds__(256) fp64_fma_burn_kernel(int64_t iters, double* sink) {
    double a0 = 1.0 + threadIdx.x * 1e-9, a1 = a0 + 1e-3, a2 = a0 + 2e-3, a3 = a0 + 3e-3;
    double a4 = a0 + 4e-3, a5 = a0 + 5e-3, a6 = a0 + 6e-3, a7 = a0 + 7e-3;
    const double m = 0.999999999, c = 1e-9;
    for (int64_t i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    sink[(int64_t)blockIdx.x * 256 + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}
// packed fp32 (FFMA2) burn: 8 independent chains of f32x2 FMAs per thread.  flops = blocks*256*iters*8*2*2.
__global__ void __launch_bounds__(256) fp32_fma_burn_kernel(int64_t iters, float* sink) {
    unsigned long long a[8];
    const float base = 1.0f + threadIdx.x * 1e-6f;
    for (int q = 0; q < 8; ++q) {
        float v = base + q * 1e-3f;
        a[q] = ((unsigned long long)__float_as_uint(v) << 32) | __float_as_uint(v + 1e-4f);
    }
    const unsigned long long m = ((unsigned long long)__float_as_uint(0.99999f) << 32) | __float_as_uint(0.99999f);
    const unsigned long long c = ((unsigned long long)__float_as_uint(1e-5f) << 32) | __float_as_uint(1e-5f);
    for (int64_t i = 0; i < iters; ++i) {
#pragma unroll
        for (int q = 0; q < 8; ++q) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[q]) : "l"(m), "l"(c));
    }
    float s = 0.f;
    for (int q = 0; q < 8; ++q) s += __uint_as_float((unsigned)a[q]) + __uint_as_float((unsigned)(a[q] >> 32));
    sink[(int64_t)blockIdx.x * 256 + threadIdx.x] = s;
}
// fp64 tensor-core (DMMA m8n8k4) burn: 4 independent accumulator fragments per warp.
// flops = blocks*8 warps*iters*4*(8*8*4*2).
__global__ void __launch_bounds__(256) fp64_mma_burn_kernel(int64_t iters, double* sink) {
    double a = 1.0 + threadIdx.x * 1e-9, b = 0.999999 + threadIdx.x * 1e-10;
    double c[4][2];
    for (int q = 0; q < 4; ++q) c[q][0] = c[q][1] = 1e-3 * q;
    for (int64_t i = 0; i < iters; ++i) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[q][0]), "+d"(c[q][1])
                         : "d"(a), "d"(b));
    }
    double s = 0.0;
    for (int q = 0; q < 4; ++q) s += c[q][0] + c[q][1];
    sink[(int64_t)blockIdx.x * 256 + threadIdx.x] = s;
}
// MUFU.COS burn: 8 independent __cosf chains per thread (FMUL.RZ + MUFU.COS each).  ops = blocks*256*iters*8.
__global__ void __launch_bounds__(256) mufu_cos_burn_kernel(int64_t iters, float* sink) {
    float a[8];
    for (int q = 0; q < 8; ++q) a[q] = 0.1f * (q + 1) + threadIdx.x * 1e-4f;
    for (int64_t i = 0; i < iters; ++i) {
#pragma unroll
        for (int q = 0; q < 8; ++q) a[q] = __cosf(a[q]);
    }
    float s = 0.f;
    for (int q = 0; q < 8; ++q) s += a[q];
    sink[(int64_t)blockIdx.x * 256 + threadIdx.x] = s;
}
}  // namespace tes

extern "C" int tes_mufu_cos_burn(int blocks, int64_t iters, float* sink, double* ops, void* stream) {
    if (blocks < 1) return -1;
    if (iters < 1) return -2;
    if (!sink) return -3;
    tes::mufu_cos_burn_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, sink);
    TES_LAUNCH_OK();
    if (ops) *ops = (double)blocks * 256.0 * (double)iters * 8.0;
    return 0;
}

extern "C" int tes_fp64_mma_burn(int blocks, int64_t iters, double* sink, double* flops, void* stream) {
    if (blocks < 1) return -1;
    if (iters < 1) return -2;
    if (!sink) return -3;
    tes::fp64_mma_burn_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, sink);
    TES_LAUNCH_OK();
    if (flops) *flops = (double)blocks * 8.0 * (double)iters * 4.0 * 512.0;
    return 0;
}

extern "C" int tes_fp32_fma_burn(int blocks, int64_t iters, float* sink, double* flops, void* stream) {
    if (blocks < 1) return -1;
    if (iters < 1) return -2;
    if (!sink) return -3;
    tes::fp32_fma_burn_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, sink);
    TES_LAUNCH_OK();
    if (flops) *flops = (double)blocks * 256.0 * (double)iters * 32.0;
    return 0;
}

extern "C" int tes_fp64_fma_burn(int blocks, int64_t iters, double* sink, double* flops, void* stream) {
    if (blocks < 1) return -1;
    if (iters < 1) return -2;
    if (!sink) return -3;
    tes::fp64_fma_burn_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, sink);
    TES_LAUNCH_OK();
    if (flops) *flops = (double)blocks * 256.0 * (double)iters * 16.0;
    return 0;
}
