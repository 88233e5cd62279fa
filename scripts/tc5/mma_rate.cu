// Micro-benchmark: issue rate of tcgen05.mma kind::f16 (cta_group::1, M = 128, K = 16) for N = 64 / 128 / 256 on
// operands resident in shared memory (no loads at all): what one SM's tensor pipe sustains per instruction shape.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/tc5/mma_rate scripts/tc5/mma_rate.cu && scripts/tc5/mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__host__ __device__ constexpr uint32_t idesc_f16(int m, int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }

template <int N>
__global__ void __launch_bounds__(128, 1) rate_kernel(int iters, unsigned long long* cycles) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    for (int i = threadIdx.x; i < (128 * 64 * 2 + N * 64 * 2) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        const uint32_t sa = smem_u32(sm), sb = sa + 128 * 64 * 2;
        const uint32_t idesc = idesc_f16(128, N);
        const unsigned long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {     // a 64-k slab: 4 k16 steps, K-major no-swizzle core matrices
                const uint64_t da = tc_desc(sa + ks * 2 * 128 * 16, 128 * 16, 128);
                const uint64_t db = tc_desc(sb + ks * 2 * N * 16, N * 16, 128);
                asm volatile(
                    "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                    "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem + (uint32_t)((it & 1) * 256)),
                    "l"(da), "l"(db), "r"(idesc), "r"(1u)
                    : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}\n"
                         : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
        cycles[blockIdx.x] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
    }
}

template <int N>
void run(int iters) {
    unsigned long long* d;
    cudaMalloc(&d, 148 * 8);
    const size_t smem = 128 * 64 * 2 + N * 64 * 2 + 1024;
    cudaFuncSetAttribute(rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int rep = 0; rep < 2; ++rep) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        rate_kernel<N><<<148, 128, smem>>>(iters, d);
        cudaEventRecord(e1);
        cudaError_t err = cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        unsigned long long h[148];
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        const double flops = 148.0 * iters * 4.0 * 2.0 * 128 * N * 16;
        if (rep) printf("N=%3d: %s  %.3f ms  %.1f TFLOP/s chip  (%.1f clk per M128 N%d K16 mma, SM 0)\n", N, cudaGetErrorString(err), ms,
                        flops / (ms * 1e-3) * 1e-12, (double)h[0] / (iters * 4.0), N);
    }
    cudaFree(d);
}

int main() {
    run<64>(20000);
    run<128>(20000);
    run<256>(20000);
    return 0;
}
