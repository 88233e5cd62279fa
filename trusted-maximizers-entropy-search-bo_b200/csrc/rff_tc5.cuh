// rff_tc5.cuh -- stage-A screen on the 5th-generation tensor cores (tcgen05 + TMEM), included by rff.cu.
//
// Same screen-and-refine scheme and the same 3xTF32 split as rff_screen_mma_kernel (x = x_hi + x_lo, w = w_hi + w_lo,
// h ~ x_hi w_hi + x_hi w_lo + x_lo w_hi, bias as two K columns against a constant 1), but the d-term inner products
// leave the issue path entirely:
//   * warp 0 (one lane)  : producer -- streams 64-feature B tiles [K/4 chunks][64 rows][4 floats] + theta[64] through
//                          a 6-stage shared-memory ring with 1-D bulk async copies (TMA engine) + mbarriers;
//   * warp 1 (one lane)  : issues tcgen05.mma kind::tf32, M = 128 candidates x N = 64 features x K = 8 per
//                          instruction, A (candidates) and B (features) K-major / no-swizzle in shared memory,
//                          accumulators in TMEM: 2 stages x 4 candidate tiles x 64 columns = all 512 columns;
//                          tcgen05.commit -> mbarrier tells the epilogue an accumulator stage is complete;
//   * warps 2..17        : epilogue -- thread = one candidate (TMEM lane); tcgen05.ld 32 columns at a time, cosine on
//                          the MUFU (FMUL.RZ + MUFU.COS) for most pairs and on the FMA pipe (packed polynomial) for
//                          NPOLY of every 8 pairs so that both pipes run concurrently, packed FFMA2 with theta, fp32
//                          partial sums flushed to fp64 every 2 feature tiles; the finished screen value goes to a
//                          double-buffered shared array;
//   * warp 18            : commit -- screens each sample's 512 values against the running k-th best and keeps the
//                          candidates within 2 B_j (same rule as the other screens), off the epilogue's critical path.
// A candidate tile (512 rows) is converted fp64 -> (hi, lo) tf32 once and stays in shared memory while the whole
// sample group streams past it.
#pragma once
// (included inside namespace tes, after the helpers of the mma.sync screen)

constexpr int TC5_EPI_WARPS = 16;
constexpr int TC5_EPI_THREADS = 32 * TC5_EPI_WARPS;
constexpr int TC5_THREADS = 64 + TC5_EPI_THREADS + 32;
constexpr int TC5_POLY_WARPS = 4;           // warp-specialised variant: one FMA-pipe cosine warp per TMEM lane quarter
constexpr int TC5_THREADS_SPEC = TC5_THREADS + 32 * TC5_POLY_WARPS;
constexpr int TC5_MT = 4;                   // candidate tiles of 128 rows per CTA tile
constexpr int TC5_TILE = 128 * TC5_MT;      // candidates per CTA tile
constexpr int TC5_NF = 64;                  // features per MMA tile = TMEM columns per candidate tile
constexpr int TC5_NSTAGE = 8;               // B ring stages (power of two: cheap ring index)
constexpr int TC5_MAX_KC = 8;               // K <= 32 (3d + 2 <= 32 <=> d <= 10)
#ifndef TC5_SLEEP_NS
#define TC5_SLEEP_NS 64
#endif
constexpr int TC5_PC = 0;                   // default: mixed epilogue (set to 16/24 for the warp-specialised variant)
constexpr int TC5_NPOLY = 1;                // pairs of every 8 whose cosine runs on the FMA pipe

__host__ __device__ constexpr int tc5_kc(int d) { return (((3 * d + 2 + 3) / 4) + 1) & ~1; }  // 16-byte K chunks, even
__host__ __device__ constexpr int tc5_stage_floats(int kc) { return kc * TC5_NF * 4 + TC5_NF; }

// ---- PTX wrappers ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc5_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc5_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc5_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// shared-memory matrix descriptor, K-major, no swizzle: core matrix = 8 rows x 16 bytes (128 contiguous bytes);
// LBO = byte distance between core matrices adjacent in K, SBO = between core matrices adjacent in M/N
__device__ __forceinline__ uint64_t tc5_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version of sm_100
    return d;
}
// D[tmem] (+)= A[smem] * B[smem]^T, kind::tf32, issued by one thread
__device__ __forceinline__ void tc5_mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
}
// instruction descriptor: D = f32, A = B = tf32, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr uint32_t tc5_idesc(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
#define TC5_LD32(v, taddr)                                                                                            \
    asm volatile(                                                                                                     \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,"  \
        "%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                                                \
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),             \
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),       \
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),     \
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])      \
        : "r"(taddr))
#define TC5_LD16(v, taddr)                                                                                            \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, "  \
                 "[%16];"                                                                                             \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),    \
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),           \
                   "=r"(v[15])                                                                                        \
                 : "r"(taddr))
#define TC5_LD8(v, taddr)                                                                                             \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"                            \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])     \
                 : "r"(taddr))
__device__ __forceinline__ void tc5_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- operand values (runtime d) --------------------------------------------------------------------------------
// B side (feature): K columns [w_hi (d) | w_lo (d) | w_hi (d) | b_hi | b_lo | 0...], radians
__device__ __forceinline__ float tc5_bval(const double* __restrict__ w, double b, int d, int k) {
    if (k < 3 * d) {
        const int kk = k < d ? k : (k < 2 * d ? k - d : k - 2 * d);
        const float wf = __double2float_rn(w[kk]);
        const float hi = __uint_as_float(to_tf32(wf));
        if (k >= d && k < 2 * d) return __uint_as_float(to_tf32(wf - hi));
        return hi;
    }
    if (k == 3 * d || k == 3 * d + 1) {
        const float bf = __double2float_rn(b);
        const float hi = __uint_as_float(to_tf32(bf));
        return k == 3 * d ? hi : __uint_as_float(to_tf32(bf - hi));
    }
    return 0.0f;
}
// A side (candidate): K columns [x_hi (d) | x_hi (d) | x_lo (d) | 1 | 1 | 0...]; written element-wise into the
// [chunk][row][4] layout of one 128-row tile (`tile_base` = float index of the tile, `rows` = rows per chunk)
__device__ __forceinline__ void tc5_store_a_row(float* sA, int rows, int row, const double* __restrict__ x, int d,
                                                bool valid) {
    auto at = [&](int k) -> float& { return sA[((k >> 2) * rows + row) * 4 + (k & 3)]; };
    for (int kk = 0; kk < d; ++kk) {
        const float xf = valid ? __double2float_rn(__ldg(x + kk)) : 0.0f;
        const float hi = __uint_as_float(to_tf32(xf));
        const float lo = __uint_as_float(to_tf32(xf - hi));
        at(kk) = hi;
        at(d + kk) = hi;
        at(2 * d + kk) = lo;
    }
    at(3 * d) = 1.0f;
    at(3 * d + 1) = 1.0f;
}

// feature tiles in operand order: featt[(j * nft + ft) * stage_floats]: [KC chunks][64 rows][4 floats], theta[64]
template <int NF>
__global__ void rff_pack_tc5_kernel(const double* __restrict__ W, const double* __restrict__ b,
                                    const double* __restrict__ theta, int64_t S, int64_t F, int d, int w_shared,
                                    int kc, float* __restrict__ featt) {
    const int64_t nft = (F + NF - 1) / NF;
    const int64_t t_ = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t_ >= S * nft * NF) return;
    const int row = (int)(t_ % NF);
    const int64_t tile = t_ / NF, ft = tile % nft, j = tile / nft;
    const int64_t jw = w_shared ? 0 : j;
    const int64_t f = ft * NF + row;
    float* base = featt + tile * (kc * NF * 4 + NF);
    for (int c = 0; c < kc; ++c) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (f < F) {
            const double* w = W + (jw * F + f) * d;
            const double bb = b[jw * F + f];
            v.x = tc5_bval(w, bb, d, 4 * c + 0);
            v.y = tc5_bval(w, bb, d, 4 * c + 1);
            v.z = tc5_bval(w, bb, d, 4 * c + 2);
            v.w = tc5_bval(w, bb, d, 4 * c + 3);
        }
        *reinterpret_cast<float4*>(base + (c * NF + row) * 4) = v;
    }
    base[kc * NF * 4 + row] = f < F ? __double2float_rn(theta[j * F + f]) : 0.0f;
}

// 32 accumulator columns of one candidate: cosine (MUFU for most pairs, FMA-pipe polynomial for NPOLY of every 8
// pairs), packed multiply-add with theta into 4 packed fp32 accumulators
template <int NPOLY, bool REDUCE>
__device__ __forceinline__ void tc5_consume32(const uint32_t (&v)[32], uint32_t th_addr, uint64_t (&acc)[4]) {
    const uint64_t INV2PI2 = f2_pack(0.15915494f, 0.15915494f);
    const uint64_t MAGIC2 = f2_pack(12582912.0f, 12582912.0f);
    const uint64_t NMAGIC2 = f2_pack(-12582912.0f, -12582912.0f);
    const uint64_t NTWOPI2 = f2_pack(-6.2831855f, -6.2831855f);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        // theta: one 16-byte shared-memory load per 4 features, same address in every lane (broadcast)
        uint64_t t01, t23;
        asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(t01), "=l"(t23) : "r"(th_addr + 16u * q));
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int p = 2 * q + e;
            uint64_t h2 = ((uint64_t)v[2 * p + 1] << 32) | (uint64_t)v[2 * p];
            if (REDUCE) h2 = ffma2(fadd2(ffma2(h2, INV2PI2, MAGIC2), NMAGIC2), NTWOPI2, h2);
            uint64_t c2;
            if ((p & 7) < NPOLY)
                c2 = cos_poly2(h2);
            else
                c2 = f2_pack(__cosf(f2_lo(h2)), __cosf(f2_hi(h2)));
            const uint64_t th2 = e == 0 ? t01 : t23;
            acc[p & 3] = ffma2(th2, c2, acc[p & 3]);
        }
    }
}

// NC accumulator columns (NC % 4 == 0) of one candidate, all on the MUFU (POLY = false) or all as the packed FMA-pipe
// polynomial (POLY = true; it reduces the argument itself), multiply-add with theta into NA packed accumulators
template <int NC, bool POLY, bool REDUCE, int NA>
__device__ __forceinline__ void tc5_consume_n(const uint32_t* v, uint32_t th_addr, uint64_t (&acc)[NA]) {
    const uint64_t INV2PI2 = f2_pack(0.15915494f, 0.15915494f);
    const uint64_t MAGIC2 = f2_pack(12582912.0f, 12582912.0f);
    const uint64_t NMAGIC2 = f2_pack(-12582912.0f, -12582912.0f);
    const uint64_t NTWOPI2 = f2_pack(-6.2831855f, -6.2831855f);
#pragma unroll
    for (int q = 0; q < NC / 4; ++q) {
        uint64_t t01, t23;
        asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(t01), "=l"(t23) : "r"(th_addr + 16u * q));
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int p = 2 * q + e;
            uint64_t h2 = ((uint64_t)v[2 * p + 1] << 32) | (uint64_t)v[2 * p];
            uint64_t c2;
            if (POLY) {
                c2 = cos_poly2(h2);
            } else {
                if (REDUCE) h2 = ffma2(fadd2(ffma2(h2, INV2PI2, MAGIC2), NMAGIC2), NTWOPI2, h2);
                c2 = f2_pack(__cosf(f2_lo(h2)), __cosf(f2_hi(h2)));
            }
            acc[p % NA] = ffma2(e == 0 ? t01 : t23, c2, acc[p % NA]);
        }
    }
}
__device__ __forceinline__ double tc5_flush2(uint64_t (&acc)[2]) {
    const uint64_t s = fadd2(acc[0], acc[1]);
    const float r = f2_lo(s) + f2_hi(s);
    acc[0] = acc[1] = 0ull;
    return (double)r;
}

__device__ __forceinline__ double tc5_flush(uint64_t (&acc)[4]) {
    // tree sum of the 8 fp32 lanes (3 roundings), one conversion, one DADD per flush
    const uint64_t s01 = fadd2(acc[0], acc[1]), s23 = fadd2(acc[2], acc[3]);
    const uint64_t s = fadd2(s01, s23);
    const float r = f2_lo(s) + f2_hi(s);
    acc[0] = acc[1] = acc[2] = acc[3] = 0ull;
    return (double)r;
}

// One warp screens the TC5_TILE values u[] (candidates base .. base + TILE, -inf where invalid) of sample jj: keep every
// candidate within 2 B_j of max(k-th largest of the 32 lane maxima, running k-th best of the chunk).
template <int TILE = TC5_TILE>
__device__ __forceinline__ void screen_commit_warp(const ScreenCtx& cx, int jj, const double* __restrict__ u,
                                                   const double* __restrict__ u2, int64_t base, int lane) {
    constexpr int PER = TILE / 32;
    const int k = cx.k;
    double v[PER];
    double lmax = neg_inf();
#pragma unroll
    for (int q = 0; q < PER; ++q) {
        v[q] = u[q * 32 + lane];
        if (u2) v[q] += u2[q * 32 + lane];  // part of the sum computed by the FMA-pipe cosine warps
        lmax = fmax(lmax, v[q]);
    }
    // k-th largest (with multiplicity) of the lane maxima: a lower bound of the tile's k-th largest value
    double lk = neg_inf();
    bool taken = false;
    for (int r = 0; r < k; ++r) {
        double m = taken ? neg_inf() : lmax;
        int64_t who = lane;
        warp_argmax(m, who);
        if (who == lane) taken = true;
        lk = m;
    }
    const double thr = fmax(lk, cx.tk_val[jj * k + k - 1]) - 2.0 * cx.bound[cx.j0 + jj];
    int n = 0;
#pragma unroll
    for (int q = 0; q < PER; ++q) {
        const int64_t ci = base + q * 32 + lane;
        const bool keep = ci < cx.c1 && v[q] >= thr;
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const int pos = k + n + __popc(m & ((1u << lane) - 1u));
            cx.ap_val[pos] = v[q];
            cx.ap_idx[pos] = ci + cx.idx_offset;
        }
        n += __popc(m);
    }
    if (n > 0) {
        if (lane < k) {
            cx.ap_val[lane] = cx.tk_val[jj * k + lane];
            cx.ap_idx[lane] = cx.tk_idx[jj * k + lane];
        }
        if (lane == 0) *cx.ap_count = n;
        __syncwarp();
        screen_fold_warp(cx, jj, lane);
        __syncwarp();
    }
}

// PC = 0: every epilogue warp mixes MUFU and polynomial cosines (NPOLY of every 8 pairs).  PC > 0 (warp-specialised):
// the 16 epilogue warps take columns [0, 64 - PC) of their candidate tile on the MUFU only, and 4 extra warps (one per
// TMEM lane quarter) take columns [64 - PC, 64) of all 4 candidate tiles with the polynomial only.  A warp issues in
// order: in the mixed scheme a warp stalled on the full MUFU queue cannot reach its FMA-pipe work; dedicated warps can.
template <int NPOLY, int PC>
__global__ void __launch_bounds__(PC > 0 ? TC5_THREADS_SPEC : TC5_THREADS, 1)
    rff_screen_tc5_kernel(const double* __restrict__ cand, int64_t N, int d, const float* __restrict__ featt, int kc,
                          int S, int F, double scale, int k, int sg, int64_t chunk, int64_t idx_offset,
                          const double* __restrict__ bound, const int* __restrict__ mode,
                          double* __restrict__ slot_val, int64_t* __restrict__ slot_idx, int slot_stride,
                          int* __restrict__ flags) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int stage_floats = tc5_stage_floats(kc);
    float* sA = reinterpret_cast<float*>(smem_raw);                       // [MT][kc][128][4]
    float* sB = sA + TC5_MT * kc * 128 * 4;                               // [NSTAGE][stage_floats]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + TC5_NSTAGE * stage_floats);
    uint64_t* full_b = bars;                                             // [NSTAGE] TMA -> MMA, epilogue
    uint64_t* empty_b = bars + TC5_NSTAGE;                               // [NSTAGE] epilogue -> TMA
    uint64_t* tmem_full = bars + 2 * TC5_NSTAGE;                         // [2] MMA -> epilogue
    uint64_t* tmem_empty = tmem_full + 2;                                // [2] epilogue -> MMA
    uint64_t* a_ready = tmem_empty + 2;                                  // [1] epilogue -> MMA
    uint64_t* u_full = a_ready + 1;                                      // [2] epilogue -> commit warp
    uint64_t* u_empty = u_full + 2;                                      // [2] commit warp -> epilogue
    double* ubuf = reinterpret_cast<double*>(bars + 2 * TC5_NSTAGE + 10);  // [2][TILE] screen values of a sample
    double* ubuf_p = ubuf + 2 * TC5_TILE;                                  // [2][TILE] polynomial-warp part of the sums
    double* tk_val = ubuf_p + 2 * TC5_TILE;
    int64_t* tk_idx = reinterpret_cast<int64_t*>(tk_val + sg * k);
    double* ap_val = reinterpret_cast<double*>(tk_idx + sg * k);          // [TILE + 16]
    int64_t* ap_idx = reinterpret_cast<int64_t*>(ap_val + TC5_TILE + 16);
    int* slot_cnt = reinterpret_cast<int*>(ap_idx + TC5_TILE + 16);
    int* ap_count = slot_cnt + sg;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ap_count + 1);

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int nsg = (S + sg - 1) / sg;
    // consecutive CTAs share the sample group and differ in the candidate chunk: the ~150 CTAs in flight then stream
    // the SAME few feature tiles, which stay L2 hits instead of cycling the whole S x F feature set through DRAM
    const int nchunks_ = (int)((N + chunk - 1) / chunk);
    const int64_t chunk_id = blockIdx.x % nchunks_;
    const int gq = blockIdx.x / nchunks_;
    const int j0 = gq * sg;
    const int nj = min(sg, S - j0);
    const int64_t c0 = chunk_id * chunk;
    const int64_t c1 = min(N, c0 + chunk);
    const int ntiles = (int)((c1 - c0 + TC5_TILE - 1) / TC5_TILE);
    const int nft = (F + TC5_NF - 1) / TC5_NF;
    const uint32_t stage_bytes = (uint32_t)(stage_floats * sizeof(float));

    constexpr int NTHREADS = PC > 0 ? TC5_THREADS_SPEC : TC5_THREADS;
    constexpr int NARRIVE = TC5_EPI_WARPS + (PC > 0 ? TC5_POLY_WARPS : 0);   // warps that read an accumulator stage / theta
    constexpr int CM = TC5_NF - (PC > 0 ? PC : 0);                                       // MUFU columns per candidate tile
    for (int e = tid; e < sg * k; e += NTHREADS) {
        tk_val[e] = neg_inf();
        tk_idx[e] = -1;
    }
    for (int e = tid; e < nj * SCR_CAP; e += NTHREADS) {
        int64_t o = (chunk_id * S + j0 + e / SCR_CAP) * slot_stride + e % SCR_CAP;
        slot_idx[o] = -1;
        slot_val[o] = neg_inf();
    }
    for (int e = tid; e < nj; e += NTHREADS) {
        slot_cnt[e] = 0;
        flags[chunk_id * S + j0 + e] = 0;
    }
    for (int e = tid; e < TC5_MT * kc * 128 * 4; e += NTHREADS) sA[e] = 0.0f;  // K padding stays zero
    if (tid == 0) {
        *ap_count = 0;
        for (int s = 0; s < TC5_NSTAGE; ++s) {
            mbar_init(&full_b[s], 1);
            mbar_init(&empty_b[s], NARRIVE);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full[a], 1);
            mbar_init(&tmem_empty[a], NARRIVE);
        }
        mbar_init(a_ready, TC5_EPI_WARPS);
        for (int a = 0; a < 2; ++a) {
            mbar_init(&u_full[a], NARRIVE);
            mbar_init(&u_empty[a], 1);
        }
        mbar_fence_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc5_fence_before();
    __syncthreads();
    tc5_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ---------------- producer ----------------
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = 0; tile < ntiles; ++tile)
                for (int jj = 0; jj < nj; ++jj) {
                    const float* src = featt + (int64_t)(j0 + jj) * nft * stage_floats;
                    for (int ft = 0; ft < nft; ++ft, ++it) {
                        const uint32_t s = it % TC5_NSTAGE, ph = (it / TC5_NSTAGE) & 1u;
                        mbar_wait_relaxed(&empty_b[s], ph ^ 1u, TC5_SLEEP_NS);
                        mbar_expect_tx(&full_b[s], stage_bytes);
                        bulk_g2s(sB + s * stage_floats, src + (int64_t)ft * stage_floats, stage_bytes, &full_b[s]);
                    }
                }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ---------------- MMA issuer ----------------
        if (lane == 0) {
            const uint32_t idesc = tc5_idesc(128, TC5_NF);
            const uint32_t a_addr = smem_u32(sA), b_addr = smem_u32(sB);
            uint32_t it = 0;
            for (int tile = 0; tile < ntiles; ++tile) {
                mbar_wait_relaxed(a_ready, (uint32_t)tile & 1u, TC5_SLEEP_NS);
                tc5_fence_after();
                for (int jj = 0; jj < nj; ++jj)
                    for (int ft = 0; ft < nft; ++ft, ++it) {
                        const uint32_t s = it % TC5_NSTAGE, ph = (it / TC5_NSTAGE) & 1u;
                        const uint32_t a = it & 1u, pa = (it >> 1) & 1u;
                        mbar_wait_relaxed(&full_b[s], ph, TC5_SLEEP_NS);
                        mbar_wait_relaxed(&tmem_empty[a], pa ^ 1u, TC5_SLEEP_NS);
                        tc5_fence_after();
#pragma unroll
                        for (int mt = 0; mt < TC5_MT; ++mt)
                            for (int ks = 0; ks < kc / 2; ++ks) {
                                const uint64_t da = tc5_desc(a_addr + (uint32_t)((mt * kc + 2 * ks) * 128 * 16), 128 * 16, 128);
                                const uint64_t db = tc5_desc(b_addr + s * stage_bytes + (uint32_t)(2 * ks * TC5_NF * 16),
                                                             TC5_NF * 16, 128);
                                tc5_mma_tf32(tmem + a * (TC5_MT * TC5_NF) + mt * TC5_NF, da, db, idesc, ks > 0);
                            }
                        tc5_commit(&tmem_full[a]);
                    }
            }
        }
        __syncwarp();
    } else if (warp == 2 + TC5_EPI_WARPS) {
        // ---------------- commit warp ----------------
        const ScreenCtx cx{tk_val, tk_idx, ap_val, ap_idx, nullptr, slot_cnt, ap_count, bound, slot_val, slot_idx, flags,
                           k, j0, S, slot_stride, chunk_id, c1, idx_offset};
        uint32_t sidx = 0;
        for (int tile = 0; tile < ntiles; ++tile)
            for (int jj = 0; jj < nj; ++jj, ++sidx) {
                const uint32_t ub = sidx & 1u, up = (sidx >> 1) & 1u;
                mbar_wait_relaxed(&u_full[ub], up, TC5_SLEEP_NS);
                screen_commit_warp(cx, jj, ubuf + ub * TC5_TILE, PC > 0 ? ubuf_p + ub * TC5_TILE : nullptr,
                                   c0 + (int64_t)tile * TC5_TILE, lane);
                __syncwarp();
                if (lane == 0) mbar_arrive(&u_empty[ub]);
            }
    } else if (PC > 0 && warp > 2 + TC5_EPI_WARPS) {
        // ---------------- FMA-pipe cosine warps: thread = TMEM lane of its quarter, all 4 candidate tiles -----------
        constexpr int PCC = PC > 0 ? PC : 16;
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t t_lane = tmem + ((uint32_t)(quarter * 32) << 16) + CM;
        uint32_t sidx = 0, it = 0;
        for (int tile = 0; tile < ntiles; ++tile) {
            for (int jj = 0; jj < nj; ++jj, ++sidx) {
                double dsum[TC5_MT] = {0.0, 0.0, 0.0, 0.0};
                uint64_t acc[TC5_MT][2];
#pragma unroll
                for (int mt = 0; mt < TC5_MT; ++mt) acc[mt][0] = acc[mt][1] = 0ull;
                for (int ft = 0; ft < nft; ++ft, ++it) {
                    const uint32_t s = it % TC5_NSTAGE, ph = (it / TC5_NSTAGE) & 1u;
                    const uint32_t a = it & 1u, pa = (it >> 1) & 1u;
                    const uint32_t th = smem_u32(sB + s * stage_floats + kc * TC5_NF * 4) + CM * 4u;
                    mbar_wait(&full_b[s], ph);
                    mbar_wait(&tmem_full[a], pa);
                    tc5_fence_after();
                    const uint32_t taddr = t_lane + a * (TC5_MT * TC5_NF);
#pragma unroll
                    for (int mp = 0; mp < TC5_MT; mp += 2) {
                        uint32_t va[PCC], vb[PCC];
                        if (PCC == 8) {
                            TC5_LD8(va, taddr + mp * TC5_NF);
                            TC5_LD8(vb, taddr + (mp + 1) * TC5_NF);
                        } else if (PCC == 16) {
                            TC5_LD16(va, taddr + mp * TC5_NF);
                            TC5_LD16(vb, taddr + (mp + 1) * TC5_NF);
                        } else if (PCC == 24) {
                            TC5_LD16(va, taddr + mp * TC5_NF);
                            TC5_LD8((va + 16), taddr + mp * TC5_NF + 16);
                            TC5_LD16(vb, taddr + (mp + 1) * TC5_NF);
                            TC5_LD8((vb + 16), taddr + (mp + 1) * TC5_NF + 16);
                        } else {
                            TC5_LD32(va, taddr + mp * TC5_NF);
                            TC5_LD32(vb, taddr + (mp + 1) * TC5_NF);
                        }
                        tc5_wait_ld();
                        tc5_consume_n<PCC, true, false, 2>(va, th, acc[mp]);
                        tc5_consume_n<PCC, true, false, 2>(vb, th, acc[mp + 1]);
                    }
                    tc5_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(&tmem_empty[a]);
                        mbar_arrive(&empty_b[s]);
                    }
                    if (ft & 1) {
#pragma unroll
                        for (int mt = 0; mt < TC5_MT; ++mt) dsum[mt] += tc5_flush2(acc[mt]);
                    }
                }
                if (nft & 1) {
#pragma unroll
                    for (int mt = 0; mt < TC5_MT; ++mt) dsum[mt] += tc5_flush2(acc[mt]);
                }
                const uint32_t ub = sidx & 1u, up = (sidx >> 1) & 1u;
                mbar_wait(&u_empty[ub], up ^ 1u);
#pragma unroll
                for (int mt = 0; mt < TC5_MT; ++mt) ubuf_p[ub * TC5_TILE + mt * 128 + row] = scale * dsum[mt];
                __syncwarp();
                if (lane == 0) mbar_arrive(&u_full[ub]);
            }
        }
    } else {
        // ---------------- epilogue: thread = candidate ----------------
        const int ew = warp - 2;
        const int mt = ew >> 2, quarter = warp & 3;  // a warp may only touch TMEM lanes 32*(warp%4) .. +31
        const int row = quarter * 32 + lane;
        uint32_t sidx = 0;
        const uint32_t t_lane = tmem + ((uint32_t)(quarter * 32) << 16) + mt * TC5_NF;
        uint32_t it = 0;
        for (int tile = 0; tile < ntiles; ++tile) {
            const int64_t ci = c0 + (int64_t)tile * TC5_TILE + mt * 128 + row;
            const bool valid = ci < c1;
            // every MMA that read the previous tile's A has completed (its last tmem_full was consumed)
            tc5_store_a_row(sA + mt * kc * 128 * 4, 128, row, cand + (valid ? ci : 0) * d, d, valid);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_ready);
            for (int jj = 0; jj < nj; ++jj, ++sidx) {
                const int red = mode[j0 + jj];
                double dsum = 0.0;
                uint64_t acc[4] = {0ull, 0ull, 0ull, 0ull};
                for (int ft = 0; ft < nft; ++ft, ++it) {
                    const uint32_t s = it % TC5_NSTAGE, ph = (it / TC5_NSTAGE) & 1u;
                    const uint32_t a = it & 1u, pa = (it >> 1) & 1u;
                    const uint32_t th = smem_u32(sB + s * stage_floats + kc * TC5_NF * 4);
                    mbar_wait(&full_b[s], ph);     // theta of this tile has landed (already true when the MMA ran)
                    mbar_wait(&tmem_full[a], pa);
                    tc5_fence_after();
                    const uint32_t taddr = t_lane + a * (TC5_MT * TC5_NF);
                    if (PC < 0) {
                        // time-alternating roles: this warp evaluates 1 of every -PC feature tiles entirely as the
                        // FMA-pipe polynomial, staggered by candidate tile so that on every scheduler one warp is on
                        // the FMA pipe while the others keep the MUFU queue full
                        const bool poly_stage = ((it + (uint32_t)mt) % (uint32_t)(PC < 0 ? -PC : 1)) == 0u;
#pragma unroll
                        for (int cb = 0; cb < 2; ++cb) {
                            uint32_t v[32];
                            TC5_LD32(v, taddr + cb * 32);
                            tc5_wait_ld();
                            if (poly_stage)
                                tc5_consume_n<32, true, false, 4>(v, th + cb * 128u, acc);
                            else if (red)
                                tc5_consume32<NPOLY, true>(v, th + cb * 128u, acc);
                            else
                                tc5_consume32<NPOLY, false>(v, th + cb * 128u, acc);
                        }
                    } else if (PC == 0) {
#pragma unroll
                        for (int cb = 0; cb < 2; ++cb) {
                            uint32_t v[32];
                            TC5_LD32(v, taddr + cb * 32);
                            tc5_wait_ld();
                            if (red)
                                tc5_consume32<NPOLY, true>(v, th + cb * 128u, acc);
                            else
                                tc5_consume32<NPOLY, false>(v, th + cb * 128u, acc);
                        }
                    } else {
                        // columns [0, CM) on the MUFU only: 32 + (CM - 32)
                        uint32_t v[32];
                        TC5_LD32(v, taddr);
                        tc5_wait_ld();
                        if (red)
                            tc5_consume_n<32, false, true, 4>(v, th, acc);
                        else
                            tc5_consume_n<32, false, false, 4>(v, th, acc);
                        constexpr int R = CM > 32 ? CM - 32 : 8;
                        if (CM > 32) {
                            if (R == 8) {
                                TC5_LD8(v, taddr + 32);
                            } else if (R == 16) {
                                TC5_LD16(v, taddr + 32);
                            } else {
                                TC5_LD16(v, taddr + 32);
                                TC5_LD8((v + 16), taddr + 48);
                            }
                            tc5_wait_ld();
                            if (red)
                                tc5_consume_n<R, false, true, 4>(v, th + 128u, acc);
                            else
                                tc5_consume_n<R, false, false, 4>(v, th + 128u, acc);
                        }
                    }
                    tc5_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(&tmem_empty[a]);
                        mbar_arrive(&empty_b[s]);
                    }
                    if (ft & 1) dsum += tc5_flush(acc);
                }
                if (nft & 1) dsum += tc5_flush(acc);
                const uint32_t ub = sidx & 1u, up = (sidx >> 1) & 1u;
                mbar_wait(&u_empty[ub], up ^ 1u);
                ubuf[ub * TC5_TILE + mt * 128 + row] = valid ? scale * dsum : neg_inf();
                __syncwarp();
                if (lane == 0) mbar_arrive(&u_full[ub]);
            }
        }
    }
    tc5_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

// test probe: max |h_tc5 - h_fp64| / A over 128 candidates x (64 * nt) features through the same operand
// construction and tcgen05 path as the screen kernel (x (128,d), w (64*nt,d), b (64*nt), fp64)
__global__ void __launch_bounds__(128) tc5_arg_probe_kernel(const double* __restrict__ x, const double* __restrict__ w,
                                                            const double* __restrict__ b, int d, int nt, double* out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int kc = tc5_kc(d);
    float* sA = reinterpret_cast<float*>(smem_raw);   // [kc][128][4]
    float* sB = sA + kc * 128 * 4;                    // [kc][64][4]
    uint64_t* bar = reinterpret_cast<uint64_t*>(sB + kc * TC5_NF * 4);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int e = tid; e < kc * 128 * 4; e += 128) sA[e] = 0.0f;
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(64));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc5_fence_before();
    __syncthreads();
    tc5_fence_after();
    const uint32_t tmem = *tmem_slot;
    tc5_store_a_row(sA, 128, tid, x + (int64_t)tid * d, d, true);
    double worst = 0.0;
    for (int t = 0; t < nt; ++t) {
        if (tid < TC5_NF) {
            const int64_t f = (int64_t)t * TC5_NF + tid;
            for (int k = 0; k < 4 * kc; ++k) sB[((k >> 2) * TC5_NF + tid) * 4 + (k & 3)] = tc5_bval(w + f * d, b[f], d, k);
        }
        fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0) {
            tc5_fence_after();
            for (int ks = 0; ks < kc / 2; ++ks)
                tc5_mma_tf32(tmem, tc5_desc(smem_u32(sA) + 2 * ks * 128 * 16, 128 * 16, 128),
                             tc5_desc(smem_u32(sB) + 2 * ks * TC5_NF * 16, TC5_NF * 16, 128), tc5_idesc(128, TC5_NF),
                             ks > 0);
            tc5_commit(bar);
        }
        mbar_wait(bar, (uint32_t)t & 1u);
        tc5_fence_after();
#pragma unroll
        for (int cb = 0; cb < 2; ++cb) {
            uint32_t v[32];
            TC5_LD32(v, tmem + ((uint32_t)(warp * 32) << 16) + cb * 32);
            tc5_wait_ld();
#pragma unroll
            for (int q = 0; q < 32; ++q) {
                const int64_t f = (int64_t)t * TC5_NF + cb * 32 + q;
                double h = b[f], A = fabs(b[f]);
                for (int kk = 0; kk < d; ++kk) {
                    h = fma(w[f * d + kk], x[(int64_t)tid * d + kk], h);
                    A += fabs(w[f * d + kk] * x[(int64_t)tid * d + kk]);
                }
                worst = fmax(worst, fabs((double)__uint_as_float(v[q]) - h) / A);
            }
        }
        tc5_fence_before();
        __syncthreads();  // the accumulator and sB are reused by the next tile
        tc5_fence_after();
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) worst = fmax(worst, __shfl_xor_sync(0xffffffffu, worst, off));
    if (lane == 0) atomicMax((unsigned long long*)out, (unsigned long long)__double_as_longlong(worst));
    tc5_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(64));
}

