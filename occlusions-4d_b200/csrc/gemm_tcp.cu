// Persistent, warp-specialized tcgen05 dense layer:  C = post(pre(A) @ W^T + bias) [+ R]
// (same contract, packed-weight format and bf16x3 arithmetic as gemm_tc.cu).
//
// Why: gemm_tc.cu runs two CTAs per SM so that one CTA's epilogue can overlap the other's main
// loop -- but both start together and stay in lock-step: their main loops contend for the A-load
// path (27 k cycles each instead of 15 k alone) and then their epilogues run side by side (22 k
// instead of ~14 k), 25 k cycles per 128 x 208 tile per SM against an 8 k MMA floor
// (in-kernel stamps).  Here ONE CTA per SM walks a list of tiles with separate
// warps per role and two TMEM accumulators, so the epilogue of tile t really overlaps the main loop
// of tile t + 1:
//   warps 0-7   A producers (fp32 global -> bf16 hi/lo core-matrix images, full-sector loads)
//   warps 8-15  epilogue (TMEM lane quarter w & 3, column half (w - 8) >> 2; transposing staging tile)
//   warp  16    TMEM allocation; lane 0 streams the packed weight slabs (cp.async.bulk)
//   warp  17    lane 0 issues tcgen05.mma and commits
// Barriers: full[s] / empty[s] per shared-memory stage (3 stages, phases run on across tiles),
// acc_full[a] / acc_empty[a] per accumulator (a = tile counter & 1).
// Tile order: n-tile fastest, so the CTAs working on the n-tiles of one row block run at the same
// time and the second read of A hits in L2.
// Status (round 1): parity green (every dense-layer and model test with O4D_TC_PERSIST=1), but SLOWER than
// gemm_tc.cu: dense family 27.8 vs 22.0 ms per step.  Stamps: per tile the epilogue warps wait 10-12 k cycles
// for the accumulator and then take 17-21 k themselves -- with the main loop's A loads and the epilogue's
// residual loads / stores in flight together, both are slower than in the lock-step arrangement, which
// (accidentally) keeps the two traffic types apart.  ncu's stall samples put the epilogue's time on the
// residual loads (issued one 32-column block ahead, far less than a memory latency).  Next: preload R into
// the free accumulator with tcgen05.st a whole tile ahead, so the epilogue has no loads at all.  Opt-in.
#include "o4d_common.cuh"
#include "tc_helpers.cuh"

namespace o4d {
namespace tcp {

using namespace tch;

constexpr int BK = 32;
constexpr int STAGES = 3;
constexpr int PROD_WARPS = 8;
constexpr int EPI_WARPS = 8;
constexpr int THREADS = (PROD_WARPS + EPI_WARPS + 2) * 32;
constexpr int A_HALF_BYTES = BM * BK * 2;                  // one bf16 image of the A slab (8 KB)
constexpr int BN_MAX = 256;
constexpr int STAGE_BYTES = 2 * A_HALF_BYTES + 2 * BN_MAX * BK * 2;   // 48 KB
constexpr int SLD = 36;                                    // epilogue staging row pitch (floats)
constexpr int STG_BYTES = EPI_WARPS * 32 * SLD * 4;        // 36 KB
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STG_BYTES + 256;

struct PackMeta {
    int n, k, bn, ntiles, kchunks;
};
// identical to tc::pack_meta (gemm_tc.cu): the packed weights are shared between the two kernels
__host__ __device__ inline PackMeta pack_meta(int n, int k) {
    PackMeta m;
    m.n = n;
    m.k = k;
    int tiles = (n + BN_MAX - 1) / BN_MAX;
    int bn = (n + tiles - 1) / tiles;
    bn = (bn + 15) / 16 * 16;
    m.bn = bn;
    m.ntiles = (n + bn - 1) / bn;
    m.kchunks = (k + BK - 1) / BK;
    return m;
}

__device__ long long g_dbg_tcp[16];

__global__ void __launch_bounds__(THREADS, 1)
linear_tcp_kernel(const float* __restrict__ A, int64_t rows, int k, int64_t lda, const __nv_bfloat16* __restrict__ Wp,
                  PackMeta m, const float* __restrict__ bias, const float* R, int64_t ldr, float* C, int64_t ldc,
                  int flags, int split, RowGather g, int total_tiles) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + STG_BYTES);
    // full[3], empty[3], acc_full[2], acc_empty[2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bn = m.bn;
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[STAGES]);
    const uint32_t accf0 = smem_u32(&bars[2 * STAGES]), acce0 = smem_u32(&bars[2 * STAGES + 2]);
    const uint32_t smem_base = smem_u32(smem);

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full0 + 8 * s, PROD_WARPS + 1);   // producer warps + the weight-copy thread
            mbar_init(empty0 + 8 * s, 1);               // one tcgen05.commit
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(accf0 + 8 * a, 1);                // one tcgen05.commit
            mbar_init(acce0 + 8 * a, EPI_WARPS);        // every epilogue warp has drained the accumulator
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == PROD_WARPS + EPI_WARPS) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int nchunks = m.kchunks;
    const uint32_t b_half_bytes = (uint32_t)bn * BK * 2;
    // tiles of this CTA: T = blockIdx.x, + gridDim.x, ...;  T -> (row tile T / ntiles, n tile T % ntiles)
    const int my_tiles = (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int k1c = (k + BK - 1) / BK;              // chunks fed by A; the rest (if any) by the K-concatenated g.a2

    if (warp < PROD_WARPS) {
        // ------------------------------------------------------------ A producers
        const bool relu_in = flags & O4D_RELU_IN;
        const int rr = lane >> 2, pq = lane & 3;
        constexpr int GPW = (BM / 8) / PROD_WARPS;      // 8-row groups per producer warp (2)
        const int total_chunks = my_tiles * nchunks;
        // chunk gidx of this CTA's flattened (tile, chunk) sequence -> registers
        auto load_chunk = [&](int gidx, float (&v)[GPW][8]) {
            const int ti = gidx / nchunks, c = gidx - ti * nchunks;
            const int T = (int)blockIdx.x + ti * (int)gridDim.x;
            const int64_t row0 = (int64_t)(T / m.ntiles) * BM;
            const bool second = c >= k1c;
            const float* Ab = second ? g.a2 : A;
            const int64_t ldab = second ? g.lda2 : lda;
            const int kk = second ? g.k2 : k;
            const int gk0 = (second ? c - k1c : c) * BK + pq * 4, gk1 = gk0 + 16;
#pragma unroll
            for (int gi = 0; gi < GPW; ++gi) {
                const int64_t grow = row0 + (warp * GPW + gi) * 8 + rr;
#pragma unroll
                for (int i = 0; i < 8; ++i) v[gi][i] = 0.f;
                if (grow < rows) {
                    const float* src = Ab + grow * ldab + gk0;
                    if (gk1 + 4 <= kk && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
                        const float4 p0 = *reinterpret_cast<const float4*>(src);
                        const float4 p1 = *reinterpret_cast<const float4*>(src + 16);
                        v[gi][0] = p0.x; v[gi][1] = p0.y; v[gi][2] = p0.z; v[gi][3] = p0.w;
                        v[gi][4] = p1.x; v[gi][5] = p1.y; v[gi][6] = p1.z; v[gi][7] = p1.w;
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            if (gk0 + i < kk) v[gi][i] = src[i];
                            if (gk1 + i < kk) v[gi][4 + i] = src[16 + i];
                        }
                    }
                }
            }
        };
        float cur[GPW][8], nxt[GPW][8];
        if (total_chunks > 0) load_chunk(0, cur);
        for (int gidx = 0; gidx < total_chunks; ++gidx) {
            const int s = gidx % STAGES;
            const uint32_t ph = (uint32_t)(gidx / STAGES) & 1u;
            if (gidx + 1 < total_chunks) load_chunk(gidx + 1, nxt);
            mbar_wait(empty0 + 8 * s, ph ^ 1u);
            uint8_t* a_hi = smem + s * STAGE_BYTES;
            uint8_t* a_lo = a_hi + A_HALF_BYTES;
            const bool relu_c = relu_in && (gidx % nchunks) < k1c;
#pragma unroll
            for (int gi = 0; gi < GPW; ++gi) {
                const int rg = warp * GPW + gi;               // 8-row group inside the 128-row tile
                __align__(16) __nv_bfloat16 h[8];
                __align__(16) __nv_bfloat16 l[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float x = relu_c ? fmaxf(cur[gi][i], 0.f) : cur[gi][i];
                    split_bf16(x, h[i], l[i]);
                }
                const int off = (pq >> 1) * (BM * 16) + rg * 128 + rr * 16 + (pq & 1) * 8;
                *reinterpret_cast<uint2*>(a_hi + off) = *reinterpret_cast<const uint2*>(h);
                *reinterpret_cast<uint2*>(a_lo + off) = *reinterpret_cast<const uint2*>(l);
                *reinterpret_cast<uint2*>(a_hi + off + 2 * (BM * 16)) = *reinterpret_cast<const uint2*>(h + 4);
                *reinterpret_cast<uint2*>(a_lo + off + 2 * (BM * 16)) = *reinterpret_cast<const uint2*>(l + 4);
            }
            fence_proxy_async_smem();   // generic-proxy stores -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(full0 + 8 * s);
#pragma unroll
            for (int gi = 0; gi < GPW; ++gi)
#pragma unroll
                for (int i = 0; i < 8; ++i) cur[gi][i] = nxt[gi][i];
        }
    } else if (warp < PROD_WARPS + EPI_WARPS) {
        // ------------------------------------------------------------ epilogue warps
        const int ew = warp - PROD_WARPS;
        const int quarter = warp & 3, chalf = ew >> 2;      // TMEM lanes 32 * (warp % 4) .. + 31 belong to this warp
        const bool relu_out = flags & O4D_RELU_OUT;
        const int nblk = (bn + 31) / 32;
        const int c_begin = chalf ? ((nblk + 1) / 2) * 32 : 0;
        const int c_end = chalf ? bn : min(bn, ((nblk + 1) / 2) * 32);
        float* stg = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES) + ew * (32 * SLD);
        const bool vec_ok = !g.qa && (m.n % 4 == 0) && (ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0) &&
                            (!R || ((ldr % 4 == 0) && ((reinterpret_cast<uintptr_t>(R) & 15) == 0))) &&
                            (!bias || ((reinterpret_cast<uintptr_t>(bias) & 15) == 0));
        const bool dbg = O4D_STAMPS && blockIdx.x == gridDim.x / 2 && ew == 0 && lane == 0;
        for (int ti = 0; ti < my_tiles; ++ti) {
            const int T = (int)blockIdx.x + ti * (int)gridDim.x;
            const int64_t row0 = (int64_t)(T / m.ntiles) * BM;
            const int tile_n = T % m.ntiles;
            const int acc = ti & 1;
            if (dbg && ti == 1) g_dbg_tcp[0] = clock64();
            mbar_wait(accf0 + 8 * acc, (uint32_t)(ti >> 1) & 1u);
            tc_fence_after();
            if (dbg && ti == 1) g_dbg_tcp[1] = clock64();
            const int64_t warp_row0 = row0 + quarter * 32;
            const int col_base = tile_n * bn;
            const uint32_t taddr_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * 256);
            const int rows_here = (int)min((int64_t)32, rows - warp_row0);     // may be <= 0 for a ragged last tile
            if (vec_ok) {
                // write phase: lane -> (row-in-group rr, 4 columns c4): one instruction covers 4 rows x 128 B
                const int rr = lane >> 3, c4 = (lane & 7) * 4;
                float4 resn[8];
                auto load_res = [&](int c0) {
                    const int gc = col_base + c0 + c4;
                    const bool ok = R && c4 < min(32, bn - c0) && gc < m.n;
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int row = it * 4 + rr;
                        resn[it] = (ok && row < rows_here) ? *reinterpret_cast<const float4*>(R + (warp_row0 + row) * ldr + gc)
                                                           : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                };
                if (c_begin < c_end) load_res(c_begin);
                for (int c0 = c_begin; c0 < c_end; c0 += 32) {
                    const int width = min(32, bn - c0);
                    float4 res[8];
#pragma unroll
                    for (int it = 0; it < 8; ++it) res[it] = resn[it];
                    if (c0 + 32 < c_end) load_res(c0 + 32);
                    float v[32];
                    tmem_ld16(taddr_row + (uint32_t)c0, v);
                    if (width > 16) tmem_ld16(taddr_row + (uint32_t)(c0 + 16), v + 16);
#pragma unroll
                    for (int i = 0; i < 32; i += 4)
                        if (i < width) *reinterpret_cast<float4*>(stg + lane * SLD + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                    __syncwarp();
                    const int gc = col_base + c0 + c4;
                    if (c4 < width && gc < m.n) {
                        const float4 bv = bias ? *reinterpret_cast<const float4*>(bias + gc) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int it = 0; it < 8; ++it) {
                            const int row = it * 4 + rr;
                            if (row < rows_here) {
                                float4 x = *reinterpret_cast<const float4*>(stg + row * SLD + c4);
                                x.x += bv.x; x.y += bv.y; x.z += bv.z; x.w += bv.w;
                                if (relu_out) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f); }
                                x.x += res[it].x; x.y += res[it].y; x.z += res[it].z; x.w += res[it].w;
                                *reinterpret_cast<float4*>(C + (warp_row0 + row) * ldc + gc) = x;
                            }
                        }
                    }
                    __syncwarp();
                }
            } else {
                // generic path (unaligned views, n % 4 != 0, row-gather epilogue): one row per thread
                const int64_t grow = warp_row0 + lane;
                const float* gq = nullptr;
                const float* gkp = nullptr;
                if (g.qa && grow < rows) {
                    const int64_t ar = g.row_offset + grow;
                    gq = g.qa + (ar / g.knbr) * m.n;
                    gkp = g.ka + (int64_t)g.nbr[ar] * m.n;
                }
                for (int c0 = c_begin; c0 < c_end; c0 += 16) {
                    float v[16];
                    tmem_ld16(taddr_row + (uint32_t)c0, v);
                    if (grow < rows) {
                        const int gc0 = col_base + c0;
                        for (int i = 0; i < 16 && gc0 + i < m.n; ++i) {
                            float x = v[i] + (bias ? bias[gc0 + i] : 0.f);
                            if (gq) x += gq[gc0 + i] - gkp[gc0 + i];
                            if (relu_out) x = fmaxf(x, 0.f);
                            if (R) x += R[grow * ldr + gc0 + i];
                            C[grow * ldc + gc0 + i] = x;
                        }
                    }
                }
            }
            // accumulator drained: the MMA warp may overwrite it (tile ti + 2)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acce0 + 8 * acc);
            if (dbg && ti == 1) g_dbg_tcp[2] = clock64();
        }
    } else if (warp == PROD_WARPS + EPI_WARPS) {
        // ------------------------------------------------------------ weight slabs via the TMA engine
        if (lane == 0) {
            int gidx = 0;
            for (int ti = 0; ti < my_tiles; ++ti) {
                const int T = (int)blockIdx.x + ti * (int)gridDim.x;
                const int tile_n = T % m.ntiles;
                const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(Wp) + (size_t)tile_n * nchunks * 2 * b_half_bytes;
                for (int c = 0; c < nchunks; ++c, ++gidx) {
                    const int s = gidx % STAGES;
                    const uint32_t ph = (uint32_t)(gidx / STAGES) & 1u;
                    mbar_wait(empty0 + 8 * s, ph ^ 1u);
                    const uint32_t dst = smem_base + s * STAGE_BYTES + 2 * A_HALF_BYTES;
                    mbar_arrive_expect_tx(full0 + 8 * s, 2 * b_half_bytes);
                    bulk_g2s(dst, wsrc + (size_t)c * 2 * b_half_bytes, 2 * b_half_bytes, full0 + 8 * s);
                }
            }
        }
    } else {
        // ------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            const uint32_t idesc = umma_idesc(bn);
            const uint32_t lbo_a = BM * 16, lbo_b = (uint32_t)bn * 16;
            int gidx = 0;
            for (int ti = 0; ti < my_tiles; ++ti) {
                const int acc = ti & 1;
                // the epilogue warps have drained this accumulator (tile ti - 2)
                mbar_wait(acce0 + 8 * acc, ((uint32_t)(ti >> 1) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t dcol = tmem_base + (uint32_t)(acc * 256);
                for (int c = 0; c < nchunks; ++c, ++gidx) {
                    const int s = gidx % STAGES;
                    const uint32_t ph = (uint32_t)(gidx / STAGES) & 1u;
                    mbar_wait(full0 + 8 * s, ph);
                    tc_fence_after();
                    const uint32_t a_hi = smem_base + s * STAGE_BYTES;
                    const uint32_t a_lo = a_hi + A_HALF_BYTES;
                    const uint32_t b_hi = a_hi + 2 * A_HALF_BYTES;
                    const uint32_t b_lo = b_hi + b_half_bytes;
#pragma unroll
                    for (int ks = 0; ks < BK / 16; ++ks) {
                        const uint64_t da_hi = umma_desc(a_hi + ks * 2 * lbo_a, lbo_a, 128);
                        const uint64_t db_hi = umma_desc(b_hi + ks * 2 * lbo_b, lbo_b, 128);
                        umma_f16(dcol, da_hi, db_hi, idesc, (c | ks) ? 1u : 0u);
                        if (split) {
                            const uint64_t da_lo = umma_desc(a_lo + ks * 2 * lbo_a, lbo_a, 128);
                            const uint64_t db_lo = umma_desc(b_lo + ks * 2 * lbo_b, lbo_b, 128);
                            umma_f16(dcol, da_lo, db_hi, idesc, 1u);
                            umma_f16(dcol, da_hi, db_lo, idesc, 1u);
                        }
                    }
                    umma_commit(empty0 + 8 * s);          // frees the stage when these MMAs retire
                }
                umma_commit(accf0 + 8 * acc);             // accumulator complete -> epilogue warps
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == PROD_WARPS + EPI_WARPS) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

}  // namespace tcp

// Same packed-weight format as gemm_tc.cu (tc_pack_launch).
int linear_tcp_packed_launch(const float* A, int64_t rows, int64_t k, int64_t lda, const void* packed, int64_t n,
                             const float* bias, const float* R, int64_t ldr, float* C, int64_t ldc, int flags,
                             int precision, cudaStream_t st, const RowGather* gp) {
    if (rows == 0) return 0;
    RowGather g;
    if (gp) g = *gp;
    O4D_SMEM_ATTR(tcp::linear_tcp_kernel, tcp::SMEM_BYTES);
    const int64_t ktot = g.a2 ? cdiv(k, tcp::BK) * tcp::BK + g.k2 : k;   // K-concatenated second operand
    tcp::PackMeta m = tcp::pack_meta((int)n, (int)ktot);
    const int64_t total = cdiv(rows, tch::BM) * m.ntiles;
    O4D_REQUIRE(total < (1LL << 30), "linear: too many tiles");
    // one CTA per SM; an even grid keeps the n-tiles of a row block (n fastest) on CTAs that run side by side
    int grid = 148;
    if (total < grid) grid = (int)total;
    tcp::linear_tcp_kernel<<<grid, tcp::THREADS, tcp::SMEM_BYTES, st>>>(A, rows, (int)k, lda, (const __nv_bfloat16*)packed, m, bias, R,
                                                                        ldr, C, ldc, flags, precision == 1 ? 1 : 0, g, (int)total);
    O4D_LAUNCH_CHECK();
    return 0;
}

}  // namespace o4d

extern "C" int o4d_debug_read_tcp(long long* out16) {
    return (int)cudaMemcpyFromSymbol(out16, o4d::tcp::g_dbg_tcp, sizeof(long long) * 16);
}
