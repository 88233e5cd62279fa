// tes_common.cuh -- device helpers shared by the sm_100a kernels of libtes_b200.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/tes_b200.h"
#include "../../include/tes_spec.h"

#define TES_CUDA_OK(call)                          \
    do {                                           \
        cudaError_t e__ = (call);                  \
        if (e__ != cudaSuccess) return (int)e__;   \
    } while (0)

// every kernel launch of the library is followed by TES_LAUNCH_OK(): it checks the launch and bumps
// the process-wide launch counter (a statistic, read by tes_launch_count(); not program state)
extern "C" void tes_count_launch_(void);
#define TES_LAUNCH_OK()                                  \
    do {                                                 \
        cudaError_t e__ = cudaGetLastError();            \
        if (e__ != cudaSuccess) return (int)e__;         \
        tes_count_launch_();                             \
    } while (0)


namespace tes {

// polynomial of tes_cospi in constant memory: DFMA reads c[bank][off] operands directly
__constant__ double kCospiC[TES_COSPI_NCOEF] = {TES_COSPI_C0, TES_COSPI_C1, TES_COSPI_C2, TES_COSPI_C3, TES_COSPI_C4,
                                               TES_COSPI_C5, TES_COSPI_C6, TES_COSPI_C7, TES_COSPI_C8};

static inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

__device__ __forceinline__ double neg_inf() { return __longlong_as_double(0xfff0000000000000LL); }
__device__ __forceinline__ double pos_inf() { return __longlong_as_double(0x7ff0000000000000LL); }

// cos(pi*h) per include/tes_spec.h.  rint(h) by the 1.5*2^52 magic constant (identical to
// round-to-nearest-even for |h| < 2^51); parity of the integer is the low mantissa bit.
__device__ __forceinline__ double cospi_spec(double h) {
    const double MAGIC = 6755399441055744.0;  // 0x1.8p52
    double t = __dadd_rn(h, MAGIC);
    int lo = __double2loint(t);
    double n = __dsub_rn(t, MAGIC);
    double r = __dsub_rn(h, n);
    double s = __dmul_rn(r, r);
    double p = kCospiC[8];
    p = fma(p, s, kCospiC[7]);
    p = fma(p, s, kCospiC[6]);
    p = fma(p, s, kCospiC[5]);
    p = fma(p, s, kCospiC[4]);
    p = fma(p, s, kCospiC[3]);
    p = fma(p, s, kCospiC[2]);
    p = fma(p, s, kCospiC[1]);
    p = fma(p, s, kCospiC[0]);
    int hi = __double2hiint(p) ^ (lo << 31);
    return __hiloint2double(hi, __double2loint(p));
}

// ordering rule: larger value first, ties -> lower index
__device__ __forceinline__ bool better(double va, int64_t ia, double vb, int64_t ib) {
    return va > vb || (va == vb && ia < ib);
}

__device__ __forceinline__ void warp_argmax(double& v, int64_t& i) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        double ov = __shfl_xor_sync(0xffffffffu, v, off);
        int64_t oi = __shfl_xor_sync(0xffffffffu, i, off);
        if (better(ov, oi, v, i)) {
            v = ov;
            i = oi;
        }
    }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// ---- mbarrier + 1-D bulk async copy (TMA engine; SASS: UBLKCP) ---------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
// same wait for service warps that have slack (producer / MMA issuer / commit): back off with nanosleep instead of
// spinning so that they do not take issue slots from the compute warps of their scheduler
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity, unsigned ns) {
    uint32_t ok;
    for (;;) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (ok) break;
        __nanosleep(ns);
    }
}
// global -> shared bulk copy, completion counted in bytes on `bar`.  16-byte aligned, size % 16 == 0.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- Philox4x32-10 + Box-Muller per include/tes_spec.h -----------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t k0,
                                              uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(TES_PHILOX_M0, c0), lo0 = TES_PHILOX_M0 * c0;
        uint32_t hi1 = __umulhi(TES_PHILOX_M1, c2), lo1 = TES_PHILOX_M1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0;
        c1 = lo1;
        c2 = n2;
        c3 = lo0;
        k0 += TES_PHILOX_W0;
        k1 += TES_PHILOX_W1;
    }
}

__device__ __forceinline__ void normal_pair(uint64_t seed, uint32_t iteration, uint32_t stream, uint64_t pair,
                                            double& z0, double& z1) {
    uint32_t c0 = (uint32_t)pair, c1 = (uint32_t)(pair >> 32), c2 = iteration, c3 = stream;
    philox4x32_10(c0, c1, c2, c3, (uint32_t)seed, (uint32_t)(seed >> 32));
    uint64_t a = ((uint64_t)c1 << 32) | c0;
    uint64_t b = ((uint64_t)c3 << 32) | c2;
    double u1 = ((double)(a >> 11) + 0.5) * 0x1.0p-53;
    double u2 = ((double)(b >> 11) + 0.5) * 0x1.0p-53;
    double rad = sqrt(-2.0 * log(u1));
    double sn, cs;
    sincos(6.283185307179586476925 * u2, &sn, &cs);
    z0 = rad * cs;
    z1 = rad * sn;
}

__device__ __forceinline__ double normal_elem(uint64_t seed, uint32_t iteration, uint32_t stream, uint64_t e) {
    double z0, z1;
    normal_pair(seed, iteration, stream, e >> 1, z0, z1);
    return (e & 1) ? z1 : z0;
}

}  // namespace tes
