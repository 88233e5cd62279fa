// rff_tc5t.cuh -- stage-A screen on tcgen05 + TMEM, TRANSPOSED accumulator layout (included by rff.cu after rff_tc5.cuh).
//
// Same screen-and-refine scheme, operand split and service warps as rff_screen_tc5_kernel, but the MMA is issued with
// the FEATURES as the M operand and the CANDIDATES as the N operand:  D[128 features][256 candidates].  A TMEM lane
// (= an epilogue thread) is then one feature and its 32x32b load yields that feature's argument for 32 candidates:
//   * theta is a per-thread scalar, loaded ONCE per feature tile -- the per-element shared-memory theta loads of the
//     candidate-major kernel (16 LDS.128 per warp and tile, through the same MIO queue that feeds the MUFU) are gone;
//   * every thread accumulates theta_f cos(h) for its 64 candidates over the F/128 feature tiles of a sample (8 terms
//     at F = 1024: no periodic fp64 flush) and the sum over the features is a warp transpose-reduce (31 packed
//     shuffles per warp and SAMPLE) followed by a 4-way add over the lane quarters in the commit warp.
// Accumulator stages: 2 x 256 TMEM columns; the candidate tile of a CTA is 256 rows.
#pragma once

constexpr int T5T_NF = 128;     // features per MMA tile = TMEM lanes
constexpr int T5T_TILE = 256;   // candidates per CTA tile = TMEM columns of one accumulator stage
constexpr int T5T_NSTAGE = 6;   // feature ring stages
constexpr int T5T_THREADS = TC5_THREADS;  // producer + MMA + 16 epilogue warps + commit warp

__host__ __device__ constexpr int t5t_stage_floats(int kc) { return kc * T5T_NF * 4 + T5T_NF; }

// cosine of 16 arguments (8 packed pairs = 16 candidates of one feature), theta-weighted accumulation
template <int NPOLY, bool REDUCE>
__device__ __forceinline__ void t5t_consume16(const uint32_t (&v)[16], uint64_t th2, uint64_t* acc) {
    const uint64_t INV2PI2 = f2_pack(0.15915494f, 0.15915494f);
    const uint64_t MAGIC2 = f2_pack(12582912.0f, 12582912.0f);
    const uint64_t NMAGIC2 = f2_pack(-12582912.0f, -12582912.0f);
    const uint64_t NTWOPI2 = f2_pack(-6.2831855f, -6.2831855f);
#pragma unroll
    for (int p = 0; p < 8; ++p) {
        uint64_t h2 = ((uint64_t)v[2 * p + 1] << 32) | (uint64_t)v[2 * p];
        uint64_t c2;
        if (p < NPOLY) {
            c2 = cos_poly2(h2);
        } else {
            if (REDUCE) h2 = ffma2(fadd2(ffma2(h2, INV2PI2, MAGIC2), NMAGIC2), NTWOPI2, h2);
            c2 = f2_pack(__cosf(f2_lo(h2)), __cosf(f2_hi(h2)));
        }
        acc[p] = ffma2(th2, c2, acc[p]);
    }
}

__device__ __forceinline__ uint64_t shfl_xor_u64(uint64_t v, int m) {
    const uint32_t lo = __shfl_xor_sync(0xffffffffu, (uint32_t)v, m);
    const uint32_t hi = __shfl_xor_sync(0xffffffffu, (uint32_t)(v >> 32), m);
    return ((uint64_t)hi << 32) | lo;
}

// Sum acc[0..31] (packed pairs) over the 32 lanes of the warp: afterwards lane l holds in acc[0] the total of packed
// element l (bit i of the element index is decided by bit i of the lane at the step with stride 2^i).
__device__ __forceinline__ void t5t_transpose_reduce(uint64_t (&acc)[32], int lane) {
#pragma unroll
    for (int h = 16; h >= 1; h >>= 1) {
        const bool up = (lane & h) != 0;
#pragma unroll
        for (int i = 0; i < h; ++i) {
            const uint64_t keep = up ? acc[i + h] : acc[i];
            const uint64_t send = up ? acc[i] : acc[i + h];
            acc[i] = fadd2(keep, shfl_xor_u64(send, h));
        }
    }
}

template <int NPOLY>
__global__ void __launch_bounds__(T5T_THREADS, 1)
    rff_screen_tc5t_kernel(const double* __restrict__ cand, int64_t N, int d, const float* __restrict__ featt, int kc,
                           int S, int F, double scale, int k, int sg, int64_t chunk, int64_t idx_offset,
                           const double* __restrict__ bound, const int* __restrict__ mode,
                           double* __restrict__ slot_val, int64_t* __restrict__ slot_idx, int slot_stride,
                           int* __restrict__ flags) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int stage_floats = t5t_stage_floats(kc);
    float* sC = reinterpret_cast<float*>(smem_raw);                        // [kc][256][4] candidates (N operand)
    float* sF = sC + kc * T5T_TILE * 4;                                    // [NSTAGE][stage_floats] features + theta
    uint64_t* bars = reinterpret_cast<uint64_t*>(sF + T5T_NSTAGE * stage_floats);
    uint64_t* full_b = bars;                                              // [NSTAGE] TMA -> MMA, epilogue
    uint64_t* empty_b = bars + T5T_NSTAGE;                                // [NSTAGE] epilogue -> TMA
    uint64_t* tmem_full = bars + 2 * T5T_NSTAGE;                          // [2] MMA -> epilogue
    uint64_t* tmem_empty = tmem_full + 2;                                 // [2] epilogue -> MMA
    uint64_t* a_ready = tmem_empty + 2;                                   // [1] epilogue -> MMA
    uint64_t* u_full = a_ready + 1;                                       // [2] epilogue -> commit warp
    uint64_t* u_empty = u_full + 2;                                       // [2] commit warp -> epilogue
    float* upart = reinterpret_cast<float*>(bars + 2 * T5T_NSTAGE + 10);    // [2][4 quarters][TILE] partial sums
    double* ubuf = reinterpret_cast<double*>(upart + 2 * 4 * T5T_TILE);    // [TILE] screen values of a sample
    double* tk_val = ubuf + T5T_TILE;
    int64_t* tk_idx = reinterpret_cast<int64_t*>(tk_val + sg * k);
    double* ap_val = reinterpret_cast<double*>(tk_idx + sg * k);           // [TILE + 16]
    int64_t* ap_idx = reinterpret_cast<int64_t*>(ap_val + T5T_TILE + 16);
    int* slot_cnt = reinterpret_cast<int*>(ap_idx + T5T_TILE + 16);
    int* ap_count = slot_cnt + sg;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ap_count + 1);

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int nsg = (S + sg - 1) / sg;
    const int nchunks_ = (int)((N + chunk - 1) / chunk);
    const int64_t chunk_id = blockIdx.x % nchunks_;   // sample-group-major grid: feature tiles stay L2 hits
    const int gq = blockIdx.x / nchunks_;
    (void)nsg;
    const int j0 = gq * sg;
    const int nj = min(sg, S - j0);
    const int64_t c0 = chunk_id * chunk;
    const int64_t c1 = min(N, c0 + chunk);
    const int ntiles = (int)((c1 - c0 + T5T_TILE - 1) / T5T_TILE);
    const int nft = (F + T5T_NF - 1) / T5T_NF;
    const uint32_t stage_bytes = (uint32_t)(stage_floats * sizeof(float));

    for (int e = tid; e < sg * k; e += T5T_THREADS) {
        tk_val[e] = neg_inf();
        tk_idx[e] = -1;
    }
    for (int e = tid; e < nj * SCR_CAP; e += T5T_THREADS) {
        int64_t o = (chunk_id * S + j0 + e / SCR_CAP) * slot_stride + e % SCR_CAP;
        slot_idx[o] = -1;
        slot_val[o] = neg_inf();
    }
    for (int e = tid; e < nj; e += T5T_THREADS) {
        slot_cnt[e] = 0;
        flags[chunk_id * S + j0 + e] = 0;
    }
    for (int e = tid; e < kc * T5T_TILE * 4; e += T5T_THREADS) sC[e] = 0.0f;  // K padding stays zero
    if (tid == 0) {
        *ap_count = 0;
        for (int s = 0; s < T5T_NSTAGE; ++s) {
            mbar_init(&full_b[s], 1);
            mbar_init(&empty_b[s], TC5_EPI_WARPS);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full[a], 1);
            mbar_init(&tmem_empty[a], TC5_EPI_WARPS);
        }
        mbar_init(a_ready, TC5_EPI_WARPS);
        for (int a = 0; a < 2; ++a) {
            mbar_init(&u_full[a], TC5_EPI_WARPS);
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
                        const uint32_t s = it % T5T_NSTAGE, ph = (it / T5T_NSTAGE) & 1u;
                        mbar_wait_relaxed(&empty_b[s], ph ^ 1u, TC5_SLEEP_NS);
                        mbar_expect_tx(&full_b[s], stage_bytes);
                        bulk_g2s(sF + s * stage_floats, src + (int64_t)ft * stage_floats, stage_bytes, &full_b[s]);
                    }
                }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ---------------- MMA issuer: D[128 features][256 candidates] ----------------
        if (lane == 0) {
            const uint32_t idesc = tc5_idesc(T5T_NF, T5T_TILE);
            const uint32_t c_addr = smem_u32(sC), f_addr = smem_u32(sF);
            uint32_t it = 0;
            for (int tile = 0; tile < ntiles; ++tile) {
                mbar_wait_relaxed(a_ready, (uint32_t)tile & 1u, TC5_SLEEP_NS);
                tc5_fence_after();
                for (int jj = 0; jj < nj; ++jj)
                    for (int ft = 0; ft < nft; ++ft, ++it) {
                        const uint32_t s = it % T5T_NSTAGE, ph = (it / T5T_NSTAGE) & 1u;
                        const uint32_t a = it & 1u, pa = (it >> 1) & 1u;
                        mbar_wait_relaxed(&full_b[s], ph, TC5_SLEEP_NS);
                        mbar_wait_relaxed(&tmem_empty[a], pa ^ 1u, TC5_SLEEP_NS);
                        tc5_fence_after();
                        for (int ks = 0; ks < kc / 2; ++ks) {
                            const uint64_t da = tc5_desc(f_addr + s * stage_bytes + (uint32_t)(2 * ks * T5T_NF * 16),
                                                         T5T_NF * 16, 128);
                            const uint64_t db = tc5_desc(c_addr + (uint32_t)(2 * ks * T5T_TILE * 16), T5T_TILE * 16, 128);
                            tc5_mma_tf32(tmem + a * T5T_TILE, da, db, idesc, ks > 0);
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
        for (int tile = 0; tile < ntiles; ++tile) {
            const int64_t base = c0 + (int64_t)tile * T5T_TILE;
            for (int jj = 0; jj < nj; ++jj, ++sidx) {
                const uint32_t ub = sidx & 1u, up = (sidx >> 1) & 1u;
                mbar_wait_relaxed(&u_full[ub], up, TC5_SLEEP_NS);
                const float* pp = upart + ub * 4 * T5T_TILE;
#pragma unroll
                for (int q = 0; q < T5T_TILE / 32; ++q) {
                    const int c = q * 32 + lane;
                    const double sum = ((double)pp[c] + (double)pp[T5T_TILE + c]) +
                                       ((double)pp[2 * T5T_TILE + c] + (double)pp[3 * T5T_TILE + c]);
                    ubuf[c] = base + c < c1 ? scale * sum : neg_inf();
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&u_empty[ub]);  // the partial sums have been read
                screen_commit_warp<T5T_TILE>(cx, jj, ubuf, nullptr, base, lane);
                __syncwarp();
            }
        }
    } else {
        // ---------------- epilogue: thread = feature (TMEM lane), 64 candidate columns ----------------
        const int ew = warp - 2;
        const int quarter = warp & 3;  // a warp may only touch TMEM lanes 32*(warp%4) .. +31
        const int cg = ew >> 2;        // column group: candidates cg*64 .. +63
        const int et = ew * 32 + lane;
        uint32_t sidx = 0, it = 0;
        const uint32_t t_lane = tmem + ((uint32_t)(quarter * 32) << 16) + cg * 64;
        for (int tile = 0; tile < ntiles; ++tile) {
            // every MMA that read the previous tile's candidates has completed (its last tmem_full was consumed)
            if (et < T5T_TILE) {
                const int64_t ci = c0 + (int64_t)tile * T5T_TILE + et;
                const bool valid = ci < c1;
                tc5_store_a_row(sC, T5T_TILE, et, cand + (valid ? ci : 0) * d, d, valid);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_ready);
            for (int jj = 0; jj < nj; ++jj, ++sidx) {
                const int red = mode[j0 + jj];
                uint64_t acc[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) acc[i] = 0ull;
                for (int ft = 0; ft < nft; ++ft, ++it) {
                    const uint32_t s = it % T5T_NSTAGE, ph = (it / T5T_NSTAGE) & 1u;
                    const uint32_t a = it & 1u, pa = (it >> 1) & 1u;
                    mbar_wait(&full_b[s], ph);  // theta of this tile has landed (already true when the MMA ran)
                    const float th = sF[s * stage_floats + kc * T5T_NF * 4 + quarter * 32 + lane];
                    const uint64_t th2 = f2_pack(th, th);
                    mbar_wait(&tmem_full[a], pa);
                    tc5_fence_after();
                    const uint32_t taddr = t_lane + a * T5T_TILE;
#pragma unroll
                    for (int cb = 0; cb < 4; ++cb) {  // 16 columns at a time: 64 accumulator registers leave room for 16
                        uint32_t v[16];
                        TC5_LD16(v, taddr + cb * 16);
                        tc5_wait_ld();
                        if (red)
                            t5t_consume16<NPOLY, true>(v, th2, acc + cb * 8);
                        else
                            t5t_consume16<NPOLY, false>(v, th2, acc + cb * 8);
                    }
                    tc5_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(&tmem_empty[a]);
                        mbar_arrive(&empty_b[s]);
                    }
                }
                // sum over the 32 features of this warp: lane l ends with candidates (2l, 2l+1) of the column group
                t5t_transpose_reduce(acc, lane);
                const uint32_t ub = sidx & 1u, up = (sidx >> 1) & 1u;
                mbar_wait(&u_empty[ub], up ^ 1u);
                *reinterpret_cast<float2*>(upart + (ub * 4 + quarter) * T5T_TILE + cg * 64 + 2 * lane) =
                    make_float2(f2_lo(acc[0]), f2_hi(acc[0]));
                __syncwarp();
                if (lane == 0) mbar_arrive(&u_full[ub]);
            }
        }
    }
    tc5_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}
