// Standalone probe: D(128 x N) = A(128 x K) * B(N x K)^T with tcgen05.mma kind::tf32, operands in shared memory in
// the K-major / no-swizzle canonical layout [K/4 chunks][rows][4 floats]; accumulator read back with tcgen05.ld.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o probe probe.cu ; run: ./probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (sm_100)
    return d;                // layout_type = 0 (no swizzle), base_offset = 0
}

template <int N, int KC>  // KC = number of 16-byte K chunks (K = 4*KC), must be even
__global__ void __launch_bounds__(128) probe_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                    float* __restrict__ D, int variant) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* sA = reinterpret_cast<float*>(smem);                 // [KC][128][4]
    float* sB = sA + KC * 128 * 4;                              // [KC][N][4]
    uint64_t* bar = reinterpret_cast<uint64_t*>(sB + KC * N * 4);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    // operands: A is (128 x K) row-major in global, B is (N x K) row-major
    const int K = 4 * KC;
    for (int e = tid; e < 128 * K; e += 128) {
        int r = e / K, k = e % K;
        sA[((k >> 2) * 128 + r) * 4 + (k & 3)] = A[e];
    }
    for (int e = tid; e < N * K; e += 128) {
        int r = e / K, k = e % K;
        sB[((k >> 2) * N + r) * 4 + (k & 3)] = B[e];
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(N));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    if (tid == 0) {
        // instruction descriptor: c=F32 (1<<4), a=b=TF32 (2<<7, 2<<10), K-major both, N>>3 at 17, M>>4 at 24
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        uint32_t lboA = 128 * 16, sboA = 128, lboB = N * 16, sboB = 128;
        if (variant == 1) { uint32_t t = lboA; lboA = sboA; sboA = t; t = lboB; lboB = sboB; sboB = t; }
        for (int ks = 0; ks < KC / 2; ++ks) {
            const uint64_t da = make_desc(smem_u32(sA) + ks * 2 * 128 * 16, lboA, sboA);
            const uint64_t db = make_desc(smem_u32(sB) + ks * 2 * N * 16, lboB, sboB);
            const uint32_t acc = ks > 0;
            asm volatile(
                "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem),
                "l"(da), "l"(db), "r"(idesc), "r"(acc));
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)));
    }
    // everyone waits for the MMA to finish
    {
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}\n"
                         : "=r"(ok)
                         : "r"(smem_u32(bar)), "r"(0u));
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // each warp reads its 32 lanes, 32 columns at a time
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,"
            "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
              "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
              "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int q = 0; q < 32; ++q) D[(size_t)tid * N + c0 + q] = __uint_as_float(v[q]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(N));
}

static float tf32r(float x) { uint32_t u; memcpy(&u, &x, 4); u = (u + 0x1000u) & 0xFFFFE000u; float y; memcpy(&y, &u, 4); return y; }
#include <string.h>

template <int N, int KC>
int run(int variant) {
    const int K = 4 * KC;
    float *hA = (float*)malloc(128 * K * 4), *hB = (float*)malloc(N * K * 4), *hD = (float*)malloc(128 * N * 4);
    srand(1);
    for (int i = 0; i < 128 * K; ++i) hA[i] = tf32r((float)rand() / RAND_MAX - 0.5f);
    for (int i = 0; i < N * K; ++i) hB[i] = tf32r((float)rand() / RAND_MAX - 0.5f);
    float *dA, *dB, *dD;
    cudaMalloc(&dA, 128 * K * 4); cudaMalloc(&dB, N * K * 4); cudaMalloc(&dD, 128 * N * 4);
    cudaMemcpy(dA, hA, 128 * K * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, N * K * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, 128 * N * 4);
    size_t smem = (size_t)KC * (128 + N) * 16 + 64;
    cudaFuncSetAttribute(probe_kernel<N, KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe_kernel<N, KC><<<1, 128, smem>>>(dA, dB, dD, variant);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d K=%d variant=%d: CUDA error %s\n", N, K, variant, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(hD, dD, 128 * N * 4, cudaMemcpyDeviceToHost);
    double worst = 0; int bad = 0;
    for (int i = 0; i < 128; ++i) for (int j = 0; j < N; ++j) {
        double ref = 0; for (int k = 0; k < K; ++k) ref += (double)hA[i * K + k] * hB[j * K + k];
        double err = fabs(ref - hD[i * N + j]); if (err > worst) worst = err; if (err > 1e-4) ++bad;
    }
    printf("N=%d K=%d variant=%d: max err %.3e, bad %d / %d  (D[0][0]=%g D[5][7]=%g)\n", N, K, variant, worst, bad, 128 * N, hD[0], hD[5 * N + 7]);
    return bad != 0;
}

int main() {
    int rc = 0;
    rc |= run<64, 2>(0);
    rc |= run<256, 6>(0);
    rc |= run<256, 8>(0);
    rc |= run<128, 2>(0);
    rc |= run<256, 2>(0);
    return rc;
}
