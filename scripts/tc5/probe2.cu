// Standalone probes for the tcgen05 screen design (not part of the library):
//  (1) tcgen05.ld throughput per SM with 4/8/16 reader warps (32x32b.x32), all 148 SMs busy;
//  (2) throughput of a MUFU / FMA-pipe-polynomial cosine mix on register inputs (P of every 8 pairs polynomial).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe2 probe2.cu ; run: ./probe2
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define LD32(v, taddr)                                                                                                  \
    asm volatile(                                                                                                       \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19," \
        "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                                                      \
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),   \
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),        \
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),       \
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                     \
        : "r"(taddr))

__global__ void __launch_bounds__(512) tmem_ld_kernel(int iters, uint32_t* sink, long long* cycles) {
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    uint32_t acc = 0;
    const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        uint32_t v[32], w[32];
        const uint32_t col = ((it * 64) + (warp >> 2) * 128) & 511 & ~63;
        LD32(v, base + col);
        LD32(w, base + col + 32);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int q = 0; q < 32; ++q) acc ^= v[q] + w[q];
    }
    long long t1 = clock64();
    __syncthreads();
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * blockDim.x + tid] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

// ---------------- cosine mix ------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) {
    return ((uint64_t)__float_as_uint(hi) << 32) | (uint64_t)__float_as_uint(lo);
}
__device__ __forceinline__ float f2_lo(uint64_t v) { return __uint_as_float((uint32_t)v); }
__device__ __forceinline__ float f2_hi(uint64_t v) { return __uint_as_float((uint32_t)(v >> 32)); }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t fmul2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
// arguments in HALF-TURNS g (cos(pi g)): t = g + magic, n = t - magic, r = g - n, poly in r^2, sign from parity
__device__ __forceinline__ uint64_t cospi_poly2(uint64_t g) {
    const uint64_t MAGIC2 = f2_pack(12582912.0f, 12582912.0f);
    const uint64_t NMAGIC2 = f2_pack(-12582912.0f, -12582912.0f);
    const uint64_t NONE2 = f2_pack(-1.0f, -1.0f);
    const uint64_t t = fadd2(g, MAGIC2);
    const uint64_t n = fadd2(t, NMAGIC2);
    const uint64_t r = ffma2(n, NONE2, g);
    const uint64_t s2 = fmul2(r, r);
    uint64_t p = ffma2(f2_pack(0.2196f, 0.2196f), s2, f2_pack(-1.3326f, -1.3326f));
    p = ffma2(p, s2, f2_pack(4.0586f, 4.0586f));
    p = ffma2(p, s2, f2_pack(-4.9348f, -4.9348f));
    p = ffma2(p, s2, f2_pack(1.0f, 1.0f));
    const uint32_t lo = (uint32_t)p ^ ((uint32_t)t << 31);
    const uint32_t hi = (uint32_t)(p >> 32) ^ ((uint32_t)(t >> 32) << 31);
    return ((uint64_t)hi << 32) | lo;
}

template <int NPOLY>  // of every 8 pairs, NPOLY go through the polynomial
__global__ void __launch_bounds__(512) cosmix_kernel(int iters, float* sink) {
    const int tid = threadIdx.x;
    uint64_t acc[8];
    uint64_t h[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        acc[q] = 0ull;
        h[q] = f2_pack(0.001f * tid + q, 0.002f * tid - q);
    }
    const uint64_t step = f2_pack(0.37f, 0.41f), one = f2_pack(1.0f, 1.0f);
    const uint64_t th = f2_pack(0.5f, 0.25f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            h[q] = ffma2(h[q], one, step);  // stand-in for the TMEM load
            uint64_t c;
            if (q < NPOLY)
                c = cospi_poly2(h[q]);
            else
                c = f2_pack(__cosf(f2_lo(h[q])), __cosf(f2_hi(h[q])));
            acc[q] = ffma2(th, c, acc[q]);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) s += f2_lo(acc[q]) + f2_hi(acc[q]);
    sink[blockIdx.x * blockDim.x + tid] = s;
}

template <int NPOLY>
void run_mix(float* sink) {
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cosmix_kernel<NPOLY><<<148, 512>>>(100, sink);
    cudaEventRecord(e0);
    cosmix_kernel<NPOLY><<<148, 512>>>(iters, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double vals = 148.0 * 512 * 16.0 * iters;
    printf("cos mix NPOLY=%d/8: %.3f ms, %.3f Tcos/s\n", NPOLY, ms, vals / ms * 1e-9);
}

int main() {
    uint32_t* sink;
    long long* cyc;
    cudaMalloc(&sink, 148 * 512 * 4 * 2);
    cudaMalloc(&cyc, 148 * 8);
    for (int threads : {128, 256, 512}) {
        const int iters = 20000;
        tmem_ld_kernel<<<148, threads>>>(100, sink, cyc);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0);
        tmem_ld_kernel<<<148, threads>>>(iters, sink, cyc);
        cudaEventRecord(e1);
        cudaError_t e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) { printf("tmem_ld error %s\n", cudaGetErrorString(e)); return 1; }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        long long c;
        cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        double bytes_per_cta = (double)iters * (threads / 32) * 64 * 32 * 4;
        printf("tmem ld, %d warps: %.3f ms, %lld cycles, %.1f B/clk/SM, %.2f T values/s chip\n", threads / 32, ms, c,
               bytes_per_cta / (double)c, bytes_per_cta / 4 * 148 / ms * 1e-9);
    }
    run_mix<0>((float*)sink);
    run_mix<1>((float*)sink);
    run_mix<2>((float*)sink);
    run_mix<3>((float*)sink);
    run_mix<4>((float*)sink);
    run_mix<5>((float*)sink);
    run_mix<6>((float*)sink);
    run_mix<8>((float*)sink);
    return 0;
}
