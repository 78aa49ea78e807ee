// Fused multi-layer MLP kernel of the implicit decoder ("chain"): ONE persistent launch runs a whole
// sequence of dense layers -- lin_in, every ResnetBlockFC (fc_0 -> ReLU -> fc_1 + residual), the lin_z
// additions folded into the producing layers, the composite Qa projection of the next cross-attention block,
// lin_out -- for every 128-row tile of queries (model/implicit.py:93-101, 403-418, 441-443).  All of these
// are row-local, so a CTA carries ITS row tiles through the whole sequence without any grid-wide barrier.
//
// What moved compared with the per-layer kernel (gemm_tc.cu), and why:
//   * Activations between layers are exchanged as "activation images": per (128-row tile, 32-column chunk) one
//     16 KB block [bf16 hi 8 KB][bf16 lo 8 KB], each [k/8][row/8][8 rows][8 values] -- byte for byte the K-major
//     no-swizzle core-matrix layout the UMMA descriptor of the A operand describes.  The epilogue of the
//     producing layer writes them (ReLU and the hi/lo split applied once, by the producer); the consuming layer's
//     A operand is then ONE cp.async.bulk per K chunk issued by one thread.  The old main loop had eight producer
//     warps doing LDG fp32 -> split -> STS for every chunk of every n-tile and was bound by outstanding L1 sector
//     requests (36 k -> 27 k cycles per tile with full-sector loads, 15 k without the loads).
//   * The hidden activations h of a residual block never exist as fp32 anywhere: fc_0's epilogue emits only the
//     image of relu(h).  The residual stream x stays fp32 in HBM (it is the `penult` output and the exact
//     residual operand) next to its image.
//   * One CTA per SM, roles by warp: warp 8 streams A-image chunks and packed weight slabs (TMA engine), warp 9
//     issues tcgen05.mma, warps 0-7 run epilogues.  Two TMEM accumulators (columns 0 and 256): the epilogue of
//     item i overlaps the main loop of item i + 1.  An item is (layer, row tile, n-tile); a CTA interleaves its
//     row tiles inside a layer, so the layer-to-layer dependency of a tile (its images must be complete before the
//     next layer loads them) is normally satisfied two items earlier and costs nothing.
//   * bf16x3 (hi*hi + lo*hi + hi*lo, fp32 accumulate) as everywhere else; `split` = 0 issues hi*hi only.
#include "o4d_common.cuh"
#include "tc_helpers.cuh"
#include "mlp_chain.cuh"
#include <stdlib.h>

namespace o4d {
namespace mc {

using namespace tch;

constexpr int BK = 32;
constexpr int STAGES = 4;
constexpr int EPI_WARPS = 16;                            // four warps per TMEM lane quarter (column quarters)
constexpr int THREADS = (EPI_WARPS + 2) * 32;
constexpr int A_HALF = BM * BK * 2;                      // 8 KB: one bf16 image half of a chunk
constexpr int W_BYTES_MAX = 2 * BN_MAX * BK * 2;         // hi + lo slab of a 208-column n-tile
constexpr int STAGE_BYTES = IMG_CHUNK_BYTES + W_BYTES_MAX;
constexpr int SLD = 20;                                  // epilogue staging row pitch (floats): 32 rows x 16 columns per warp
constexpr int STG_BYTES = EPI_WARPS * 32 * SLD * 4;
constexpr int OFF_STG = STAGES * STAGE_BYTES;
constexpr int OFF_BAR = OFF_STG + STG_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 256;
static_assert(IMG_CHUNK_BYTES == 2 * A_HALF, "image chunk = hi + lo halves");
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

__device__ __forceinline__ void st_release_shared(uint32_t addr, int v) {
    asm volatile("st.release.cta.shared::cta.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_shared(uint32_t addr) {
    int v;
    asm volatile("ld.acquire.cta.shared::cta.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t* r) {      // no wait: pair with tmem_ld_wait()
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// byte offset of element (tile-row trow, column gc) inside an image with `cpt` chunks per tile (hi half)
__device__ __forceinline__ size_t img_off(int64_t tile, int cpt, int trow, int gc) {
    return ((size_t)tile * cpt + (gc >> 5)) * IMG_CHUNK_BYTES + (size_t)(((gc & 31) >> 3) * (BM * 16) + (trow >> 3) * 128 +
                                                                          (trow & 7) * 16 + (gc & 7) * 2);
}

__device__ __forceinline__ uint32_t pack2(__nv_bfloat16 a, __nv_bfloat16 b) {
    return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

// Cycle counters of one CTA (O4D_STAMPS builds only, o4d_debug_read_chain): which wait bounds the pipeline.
// [0] MMA thread total, [1] its wait for a free accumulator (epilogue-bound), [2] its wait for a full stage (load-bound),
// [3] producer total wait for an empty stage (MMA-bound), [4] producer wait for the tile's previous layer (dependency),
// [5] epilogue warp 0 wait for a full accumulator, [6] epilogue warp 0 busy
__device__ long long g_dbg_chain[8];
#define CHAIN_T0() (O4D_STAMPS ? clock64() : 0LL)

// PAIR = true: two CTAs of a cluster (one TPC) work as a pair (tcgen05 cta_group::2).  Each owns one 128-row tile of a
// tile pair and its own A-image stream, but loads only HALF of every weight slab; the leader's MMA thread issues
// M = 256 instructions that read B half from each CTA's shared memory.  Weight bytes per SM halve: 42.6 -> 29.3 KB
// per (128 x 208 x 32) chunk = 624 tensor cycles, i.e. 68 -> 47 B/clk per SM against the ~43 B/clk per SM the L2 can
// deliver chip-wide (the single-CTA kernel sits at 0.59 of the tensor peak for exactly this reason).
//   full[s]      leader: own expect_tx + the peer's forwarder (peer warp 9 waits for the peer's own full[s], then arrives
//                remotely); peer: own expect_tx
//   empty[s], acc_full[a]   tcgen05.commit ... multicast -> the barrier at the same offset in both CTAs
//   acc_empty[a] leader only: 8 local + 8 remote epilogue-warp arrivals
template <bool PAIR>
__global__ void __launch_bounds__(THREADS, 1) mlp_chain_kernel(const Program P) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    // full[4], empty[4], acc_full[2], acc_empty[2]; then the TMEM slot and the per-warp progress counters
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
    int* done = reinterpret_cast<int*>(bars + 13);       // [EPI_WARPS] items finished by each epilogue warp
    static_assert(13 * 8 + EPI_WARPS * 4 <= 256, "barrier area");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[STAGES]);
    const uint32_t accf0 = smem_u32(&bars[2 * STAGES]), acce0 = smem_u32(&bars[2 * STAGES + 2]);
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t done0 = smem_u32(done);
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;   // 0 = leader (issues the pair's MMAs), 1 = peer

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full0 + 8 * s, (PAIR && rank == 0) ? 2 : 1);   // the producer's arrive.expect_tx (+ bytes) [+ the peer's forwarder]
            mbar_init(empty0 + 8 * s, 1);               // one tcgen05.commit
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(accf0 + 8 * a, 1);                // one tcgen05.commit
            mbar_init(acce0 + 8 * a, PAIR ? 2 * EPI_WARPS : EPI_WARPS);   // every epilogue warp (of both CTAs) has drained it
        }
        for (int w = 0; w < EPI_WARPS; ++w) done[w] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == EPI_WARPS + 1) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();      // both CTAs' mbarriers are initialised and TMEM allocated before any remote arrive / MMA
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // work units of this CTA (pair): unit, unit + nunits, ...; a unit is a row tile, or a PAIR of row tiles (2u, 2u + 1)
    // of which this CTA takes tile 2u + rank
    const int unit0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int nunits = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int tunits = PAIR ? (P.tiles + 1) >> 1 : P.tiles;
    const int ntl = (tunits - unit0 + nunits - 1) / nunits;
    auto tile_of = [&](int j) -> int64_t {
        const int64_t u = (int64_t)unit0 + (int64_t)j * nunits;
        return PAIR ? 2 * u + rank : u;
    };

    if (warp < EPI_WARPS) {
        // ================================================================== epilogue warps
        // Sixteen warps: TMEM lane quarter warp & 3, column part warp >> 2 (a quarter of the n-tile's 16-column blocks).
        // (Measured with in-kernel counters, eight warps and 32-column blocks: the epilogue warps were busy 90 % of the
        // kernel and the MMA thread spent 40 % of its time waiting for a drained accumulator -- 18.7 k cycles of epilogue
        // per 128 x 208 item against 8.1 - 13.7 k of MMA -- at an IPC of ~0.2: dependent tcgen05.ld -> STS -> LDS -> FADD ->
        // STG chains with two warps per scheduler.  Four warps per scheduler hide that latency.)
        const int quarter = warp & 3, cpart = warp >> 2;
        const uint32_t taddr_q = tmem_base + ((uint32_t)(quarter * 32) << 16);
        float* stg = reinterpret_cast<float*>(smem + OFF_STG) + warp * (32 * SLD);
        int item = 0;
        const bool dbg = O4D_STAMPS && blockIdx.x == gridDim.x / 2 && threadIdx.x == 0;
        long long t_wait = 0, t_busy = 0;
        for (int o = 0; o < P.nops; ++o) {
            const Op& op = P.op[o];
            const int bn = op.bn;
            const bool relu_img = op.img_relu != 0;
            const int nblk = bn / 16;                                   // bn is a multiple of 16
            const int c_begin = ((nblk * cpart) / 4) * 16;
            const int c_end = ((nblk * (cpart + 1)) / 4) * 16;
            const bool image_only = op.out == nullptr && op.res == nullptr;
            const bool vec_ok = (op.n % 4 == 0) && (!op.out || ((op.ldo % 4 == 0) && ((reinterpret_cast<uintptr_t>(op.out) & 15) == 0))) &&
                                (!op.res || ((op.ldr % 4 == 0) && ((reinterpret_cast<uintptr_t>(op.res) & 15) == 0))) &&
                                (!op.bias || ((reinterpret_cast<uintptr_t>(op.bias) & 15) == 0));
            for (int j = 0; j < ntl; ++j) {
                const int64_t tile = tile_of(j);
                const int64_t row0 = tile * BM + quarter * 32;          // first row of this warp
                const int rows_here = (int)max((int64_t)0, min((int64_t)32, P.rows - row0));
                for (int nh = 0; nh < op.ntiles; ++nh, ++item) {
                    const int acc = item & 1;
                    const int col_base = nh * bn;
                    // residual rows of the first 16-column block: requested BEFORE the wait for the accumulator (they do
                    // not depend on it), so their L2 latency hides behind the item's MMAs
                    const int rr = lane >> 2, c4 = (lane & 3) * 4;
                    float4 resn[4];
                    auto load_res = [&](int c0) {
                        const int gc = col_base + c0 + c4;
                        const bool ok = op.res != nullptr && gc < op.n && c0 < c_end;
#pragma unroll
                        for (int it = 0; it < 4; ++it) {
                            const int row = it * 8 + rr;
                            resn[it] = (ok && row < rows_here) ? *reinterpret_cast<const float4*>(op.res + (row0 + row) * op.ldr + gc)
                                                               : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    };
                    if (!image_only && vec_ok) load_res(c_begin);
                    const long long te0 = CHAIN_T0();
                    if (PAIR) mbar_wait_cl(accf0 + 8 * acc, (uint32_t)(item >> 1) & 1u);
                    else mbar_wait(accf0 + 8 * acc, (uint32_t)(item >> 1) & 1u);
                    tc_fence_after();
                    const long long te1 = CHAIN_T0();
                    const uint32_t taddr = taddr_q + (uint32_t)(acc * 256);
                    if (image_only) {
                        // ---- (E1) row per thread: + bias, ReLU, split, two 16-byte lines per 8 columns.  A warp
                        // instruction stores 32 rows x 16 B = 512 contiguous bytes of the image.
                        const int trow = quarter * 32 + lane;
                        uint8_t* img_row = op.img + (size_t)tile * op.img_cpt * IMG_CHUNK_BYTES + (trow >> 3) * 128 + (trow & 7) * 16;
                        for (int c0 = c_begin; c0 < c_end; c0 += 16) {
                            float v[16];
                            tmem_ld16(taddr + (uint32_t)c0, v);
                            const int gc0 = col_base + c0;
#pragma unroll
                            for (int h8 = 0; h8 < 2; ++h8) {
                                const int gc = gc0 + h8 * 8;
                                if (gc >= op.n) continue;               // n % 32 == 0 for image outputs: whole groups only
                                float b8[8];
                                if (op.bias) {
                                    const float4 b0 = __ldg(reinterpret_cast<const float4*>(op.bias + gc));
                                    const float4 b1 = __ldg(reinterpret_cast<const float4*>(op.bias + gc + 4));
                                    b8[0] = b0.x; b8[1] = b0.y; b8[2] = b0.z; b8[3] = b0.w;
                                    b8[4] = b1.x; b8[5] = b1.y; b8[6] = b1.z; b8[7] = b1.w;
                                } else {
#pragma unroll
                                    for (int e = 0; e < 8; ++e) b8[e] = 0.f;
                                }
                                uint32_t hi[4], lo[4];
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    float x0 = v[h8 * 8 + 2 * e] + b8[2 * e], x1 = v[h8 * 8 + 2 * e + 1] + b8[2 * e + 1];
                                    if (relu_img) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); }
                                    split_bf16x2(x0, x1, hi[e], lo[e]);
                                }
                                uint8_t* dst = img_row + (size_t)(gc >> 5) * IMG_CHUNK_BYTES + ((gc & 31) >> 3) * (BM * 16);
                                *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                                *reinterpret_cast<uint4*>(dst + A_HALF) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                            }
                        }
                    } else if (vec_ok) {
                        // ---- (E2) fp32 output and / or residual: 32 x 16 blocks transposed through a padded staging tile:
                        // in the write phase a lane owns (row rr of an 8-row group, 4 columns), so residual loads and
                        // output stores are 64 contiguous bytes per row (whole sectors), and the image (if any) is
                        // written from the same mapping: 4 columns = half a core-matrix line, 8 consecutive rows = 128 B.
                        uint8_t* img_rows = op.img ? op.img + (size_t)tile * op.img_cpt * IMG_CHUNK_BYTES + (quarter * 4) * 128 + rr * 16
                                                   : nullptr;      // + it * 128: row quarter*32 + it*8 + rr
                        uint32_t vn[16];                                 // accumulator block in flight (tcgen05.ld issued, not awaited)
                        if (c_begin < c_end) tmem_ld16_issue(taddr + (uint32_t)c_begin, vn);
                        for (int c0 = c_begin; c0 < c_end; c0 += 16) {
                            const int gc = col_base + c0 + c4;
                            const bool col_ok = gc < op.n;
                            float4 res[4];
#pragma unroll
                            for (int it = 0; it < 4; ++it) res[it] = resn[it];
                            load_res(c0 + 16);                          // next block's residual rows: in flight during this block
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 16; i += 4)
                                *reinterpret_cast<float4*>(stg + lane * SLD + i) =
                                    make_float4(__uint_as_float(vn[i]), __uint_as_float(vn[i + 1]), __uint_as_float(vn[i + 2]), __uint_as_float(vn[i + 3]));
                            // the next block's accumulator columns travel while this block is transposed and written out
                            if (c0 + 16 < c_end) tmem_ld16_issue(taddr + (uint32_t)(c0 + 16), vn);
                            __syncwarp();
                            if (col_ok) {
                                const float4 bv = op.bias ? *reinterpret_cast<const float4*>(op.bias + gc) : make_float4(0.f, 0.f, 0.f, 0.f);
                                const size_t img_col = (size_t)(gc >> 5) * IMG_CHUNK_BYTES + ((gc & 31) >> 3) * (BM * 16) + (gc & 7) * 2;
#pragma unroll
                                for (int it = 0; it < 4; ++it) {
                                    const int row = it * 8 + rr;
                                    float4 x = *reinterpret_cast<const float4*>(stg + row * SLD + c4);
                                    x.x += bv.x + res[it].x; x.y += bv.y + res[it].y; x.z += bv.z + res[it].z; x.w += bv.w + res[it].w;
                                    if (op.out && row < rows_here) *reinterpret_cast<float4*>(op.out + (row0 + row) * op.ldo + gc) = x;
                                    if (op.img) {
                                        if (relu_img) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f); }
                                        uint32_t h01, l01, h23, l23;
                                        split_bf16x2(x.x, x.y, h01, l01);
                                        split_bf16x2(x.z, x.w, h23, l23);
                                        uint8_t* dst = img_rows + it * 128 + img_col;
                                        *reinterpret_cast<uint2*>(dst) = make_uint2(h01, h23);
                                        *reinterpret_cast<uint2*>(dst + A_HALF) = make_uint2(l01, l23);
                                    }
                                }
                            }
                            __syncwarp();
                        }
                    } else {
                        // ---- (E3) narrow / unaligned fp32 outputs (lin_out: 9 .. 33 columns): one row per thread
                        const int64_t grow = row0 + lane;
                        for (int c0 = c_begin; c0 < c_end; c0 += 16) {
                            float v[16];
                            tmem_ld16(taddr + (uint32_t)c0, v);
                            if (lane < rows_here && op.out) {
                                const int gc0 = col_base + c0;
                                for (int i = 0; i < 16 && gc0 + i < op.n; ++i) {
                                    float x = v[i] + (op.bias ? op.bias[gc0 + i] : 0.f);
                                    if (op.res) x += op.res[grow * op.ldr + gc0 + i];
                                    op.out[grow * op.ldo + gc0 + i] = x;
                                }
                            }
                        }
                    }
                    // accumulator drained -> the MMA issuer may overwrite it (item + 2)
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if (PAIR && rank != 0) mbar_arrive_cluster(acce0 + 8 * acc, 0);   // the leader's MMA thread waits for both CTAs
                        else mbar_arrive(acce0 + 8 * acc);
                    }
                    // this warp's part of the item is in global memory: order it before the TMA engine's (async proxy)
                    // reads of this CTA's producer: generic stores -> fence.proxy.async (every writing lane) -> warp
                    // barrier -> release store of the progress flag; the producer acquires the flag, then issues the copy.
                    // (A gpu-scope __threadfence() in front of the proxy fence showed up as 17 % of all warp stall samples
                    // in ncu -- "membar" -- and is not needed: writer and reader are the same CTA.)
                    if (P.fence_gpu) __threadfence();
                    fence_proxy_async_all();
                    __syncwarp();
                    if (lane == 0) st_release_shared(done0 + 4 * warp, item + 1);
                    if (dbg) { t_wait += te1 - te0; t_busy += clock64() - te1; }
                }
            }
        }
        if (dbg) { g_dbg_chain[5] = t_wait; g_dbg_chain[6] = t_busy; }
    } else if (warp == EPI_WARPS) {
        // ================================================================== producer: A-image chunks + weight slabs
        if (lane == 0) {
            int g = 0;            // running stage counter
            int items_before = 0; // items of all earlier ops
            const bool dbg = O4D_STAMPS && blockIdx.x == gridDim.x / 2;
            long long t_empty = 0, t_dep = 0;
            int prev_nt = 0;      // n-tiles per row tile of the previous op
            for (int o = 0; o < P.nops; ++o) {
                const Op& op = P.op[o];
                // bytes of this CTA's share of a weight slab: the whole [hi][lo] slab, or (pair format) half `rank` of it
                const uint32_t w_bytes = (uint32_t)(2 * op.bn * BK * 2) / (PAIR ? 2u : 1u);
                const int nch = op.k1c + op.k2c;
                for (int j = 0; j < ntl; ++j) {
                    const int64_t tile = tile_of(j);
                    const long long td0 = CHAIN_T0();
                    if (o > 0) {
                        // the images this op reads were written by the previous op's epilogues of THIS tile: wait
                        // until every epilogue warp has finished its last item of (o - 1, j)
                        const int need = items_before - (ntl - 1 - j) * prev_nt;
                        for (int w = 0; w < EPI_WARPS; ++w) {
                            if (ld_acquire_shared(done0 + 4 * w) >= need) continue;
                            const long long t0 = clock64();
                            while (ld_acquire_shared(done0 + 4 * w) < need) {
                                __nanosleep(64);
                                if (clock64() - t0 > 4000000000LL) __trap();
                            }
                        }
                        fence_proxy_async_all();
                    }
                    if (dbg) t_dep += clock64() - td0;
                    const uint8_t* a1 = op.a1 + (size_t)tile * op.k1c * IMG_CHUNK_BYTES;
                    const uint8_t* a2 = op.a2 ? op.a2 + (size_t)tile * op.k2c * IMG_CHUNK_BYTES : nullptr;
                    for (int nh = 0; nh < op.ntiles; ++nh) {
                        // packed per (n-tile, k-chunk): [hi slab][lo slab], or in pair format [half 0: hi, lo][half 1: hi, lo]
                        const size_t w_stride = PAIR ? 2 * (size_t)w_bytes : (size_t)w_bytes;
                        const uint8_t* wsrc = op.w + (size_t)nh * nch * w_stride + (PAIR ? (size_t)rank * w_bytes : 0);
                        for (int c = 0; c < nch; ++c, ++g) {
                            const int s = g % STAGES;
                            const uint32_t ph = (uint32_t)(g / STAGES) & 1u;
                            const long long tp0 = CHAIN_T0();
                            if (PAIR) mbar_wait_cl(empty0 + 8 * s, ph ^ 1u);
                            else mbar_wait(empty0 + 8 * s, ph ^ 1u);
                            if (dbg) t_empty += clock64() - tp0;
                            const uint32_t dst = smem_base + s * STAGE_BYTES;
                            mbar_arrive_expect_tx(full0 + 8 * s, IMG_CHUNK_BYTES + w_bytes);
                            const uint8_t* asrc = c < op.k1c ? a1 + (size_t)c * IMG_CHUNK_BYTES : a2 + (size_t)(c - op.k1c) * IMG_CHUNK_BYTES;
                            bulk_g2s(dst, asrc, IMG_CHUNK_BYTES, full0 + 8 * s);
                            bulk_g2s(dst + IMG_CHUNK_BYTES, wsrc + (size_t)c * w_stride, w_bytes, full0 + 8 * s);
                        }
                    }
                }
                items_before += ntl * op.ntiles;
                prev_nt = op.ntiles;
            }
            if (dbg) { g_dbg_chain[3] = t_empty; g_dbg_chain[4] = t_dep; }
        }
    } else {
        // ================================================================== MMA issuer (pair: the leader's; the peer's warp forwards)
        if (lane == 0 && rank == 0) {
            int g = 0, item = 0;
            const uint32_t lbo_a = BM * 16;
            const bool dbg = O4D_STAMPS && blockIdx.x == gridDim.x / 2;
            const long long tm_start = CHAIN_T0();
            long long t_acc = 0, t_full = 0;
            for (int o = 0; o < P.nops; ++o) {
                const Op& op = P.op[o];
                const int bn = op.bn;
                const int hb = PAIR ? bn / 2 : bn;                       // B rows in THIS CTA's shared memory
                const uint32_t idesc = PAIR ? umma_idesc_pair(bn) : umma_idesc(bn);
                const uint32_t lbo_b = (uint32_t)hb * 16;
                const uint32_t b_half = (uint32_t)hb * BK * 2;
                const int nch = op.k1c + op.k2c;
                for (int j = 0; j < ntl; ++j) {
                    for (int nh = 0; nh < op.ntiles; ++nh, ++item) {
                        const int acc = item & 1;
                        // accumulator free?  (its previous user was item - 2)
                        const long long ta0 = CHAIN_T0();
                        if (PAIR) mbar_wait_cl(acce0 + 8 * acc, ((uint32_t)(item >> 1) & 1u) ^ 1u);
                        else mbar_wait(acce0 + 8 * acc, ((uint32_t)(item >> 1) & 1u) ^ 1u);
                        tc_fence_after();
                        if (dbg) t_acc += clock64() - ta0;
                        const uint32_t dcol = tmem_base + (uint32_t)(acc * 256);
                        for (int c = 0; c < nch; ++c, ++g) {
                            const int s = g % STAGES;
                            const uint32_t ph = (uint32_t)(g / STAGES) & 1u;
                            const long long tf0 = CHAIN_T0();
                            if (PAIR) mbar_wait_cl(full0 + 8 * s, ph);
                            else mbar_wait(full0 + 8 * s, ph);
                            tc_fence_after();
                            if (dbg) t_full += clock64() - tf0;
                            const uint32_t a_hi = smem_base + s * STAGE_BYTES;
                            const uint32_t a_lo = a_hi + A_HALF;
                            const uint32_t b_hi = a_hi + IMG_CHUNK_BYTES;
                            const uint32_t b_lo = b_hi + b_half;
#pragma unroll
                            for (int ks = 0; ks < BK / 16; ++ks) {
                                const uint64_t da_hi = umma_desc(a_hi + ks * 2 * lbo_a, lbo_a, 128);
                                const uint64_t db_hi = umma_desc(b_hi + ks * 2 * lbo_b, lbo_b, 128);
                                const uint64_t da_lo = umma_desc(a_lo + ks * 2 * lbo_a, lbo_a, 128);
                                const uint64_t db_lo = umma_desc(b_lo + ks * 2 * lbo_b, lbo_b, 128);
                                if (PAIR) {
                                    umma_f16_pair(dcol, da_hi, db_hi, idesc, (c | ks) ? 1u : 0u);
                                    if (P.split) {
                                        umma_f16_pair(dcol, da_lo, db_hi, idesc, 1u);
                                        umma_f16_pair(dcol, da_hi, db_lo, idesc, 1u);
                                    }
                                } else {
                                    umma_f16(dcol, da_hi, db_hi, idesc, (c | ks) ? 1u : 0u);
                                    if (P.split) {
                                        umma_f16(dcol, da_lo, db_hi, idesc, 1u);
                                        umma_f16(dcol, da_hi, db_lo, idesc, 1u);
                                    }
                                }
                            }
                            if (PAIR) umma_commit_pair(empty0 + 8 * s);   // frees the stage (in both CTAs) when these MMAs retire
                            else umma_commit(empty0 + 8 * s);
                        }
                        if (PAIR) umma_commit_pair(accf0 + 8 * acc);      // accumulator complete -> epilogues (of both CTAs)
                        else umma_commit(accf0 + 8 * acc);
                    }
                }
            }
            if (dbg) { g_dbg_chain[0] = clock64() - tm_start; g_dbg_chain[1] = t_acc; g_dbg_chain[2] = t_full; g_dbg_chain[7] = item; }
        } else if (PAIR && lane == 0) {
            // peer: tell the leader's MMA thread when THIS CTA's share of a stage (its A rows, its half of the weight
            // slab) has landed.  A stage cannot be refilled before the leader has consumed it, so no phase is skipped.
            int g = 0;
            for (int o = 0; o < P.nops; ++o) {
                const int total = ntl * P.op[o].ntiles * (P.op[o].k1c + P.op[o].k2c);
                for (int i = 0; i < total; ++i, ++g) {
                    const int s = g % STAGES;
                    mbar_wait(full0 + 8 * s, (uint32_t)(g / STAGES) & 1u);
                    mbar_arrive_cluster(full0 + 8 * s, 0);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();      // the peer's TMEM / shared memory stay valid until the leader's last MMA has retired
    if (warp == EPI_WARPS + 1) {
        tc_fence_after();
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// fp32 (rows, cols) row-major -> activation image with ceil(cols / 32) chunks per tile; padding rows / columns = 0.
// thread = (row, 8-column group): 4 consecutive lanes read one 128-byte line of a row, and the 8 rows of a warp
// write 128 contiguous bytes per core-matrix column group.
__global__ void image_from_f32_kernel(const float* __restrict__ src, int64_t ld, int64_t rows, int cols, int relu,
                                      uint8_t* __restrict__ img, int cpt, int64_t tiles) {
    const int64_t units = tiles * cpt * (BM * 4);          // (tile, chunk, row, kc)
    for (int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u < units; u += (int64_t)gridDim.x * blockDim.x) {
        const int kc = (int)(u & 3);
        const int trow = (int)((u >> 2) % BM);
        const int64_t tc = u / (BM * 4);
        const int chunk = (int)(tc % cpt);
        const int64_t tile = tc / cpt;
        const int64_t grow = tile * BM + trow;
        const int gc = chunk * 32 + kc * 8;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = 0.f;
        if (grow < rows) {
            const float* p = src + grow * ld + gc;
            if (gc + 8 <= cols && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
                const float4 a = *reinterpret_cast<const float4*>(p);
                const float4 b = *reinterpret_cast<const float4*>(p + 4);
                v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e)
                    if (gc + e < cols) v[e] = p[e];
            }
        }
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float x0 = v[2 * e], x1 = v[2 * e + 1];
            if (relu) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); }
            split_bf16x2(x0, x1, hi[e], lo[e]);
        }
        uint8_t* dst = img + img_off(tile, cpt, trow, gc);
        *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(dst + A_HALF) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
}

}  // namespace mc

// O4D_CHAIN_PAIR=1 selects the CTA-pair (cta_group::2) instantiation.  Measured on the B200 (profiles/r2_c_bench_*):
// parity green, but SLOWER than the single-CTA kernel -- dense family 18.9 vs 15.2 ms per step -- although it moves a
// third fewer bytes from L2: the leader's MMA thread now waits for the slower of two TMA streams plus a remote
// mbarrier arrive per stage (the peer's forwarder), and both CTAs' epilogues gate the accumulator hand-back.  The
// single-CTA kernel's 0.59 of tensor peak is therefore not simply the L2 -> SM byte rate.  Default off.  The
// packed-weight format differs between the two, so packing and launching consult the same switch.
bool mlp_chain_pair() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("O4D_CHAIN_PAIR");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}

int mlp_chain_pack_launch(const float* W, int64_t n, int64_t k, int64_t ldw, void* packed, cudaStream_t st) {
    return mlp_chain_pair() ? tc_pack_pair_launch(W, n, k, ldw, packed, st) : tc_pack_launch(W, n, k, ldw, packed, st);
}

// images are sized for an EVEN number of row tiles: the pair kernel always processes tiles two at a time
size_t act_image_bytes(int64_t rows, int cols) {
    return (size_t)(cdiv(cdiv(rows, mc::BM), 2) * 2) * (size_t)cdiv(cols, 32) * mc::IMG_CHUNK_BYTES;
}

int act_image_launch(const float* src, int64_t ld, int64_t rows, int cols, int relu, void* img, cudaStream_t st) {
    if (rows <= 0) return 0;
    const int64_t tiles = cdiv(rows, mc::BM);
    const int cpt = (int)cdiv(cols, 32);
    const int64_t units = tiles * cpt * (mc::BM * 4);
    int64_t blocks = cdiv(units, 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    ProfScope prof(PROF_MISC, 0.0, st);
    mc::image_from_f32_kernel<<<(unsigned)blocks, 256, 0, st>>>(src, ld, rows, cols, relu, (uint8_t*)img, cpt, tiles);
    O4D_LAUNCH_CHECK();
    return 0;
}

// tc::pack_meta's tiling (gemm_tc.cu): the packed weights are shared with the per-layer kernel
void mlp_chain_tiling(int64_t n, int* bn_out, int* ntiles_out) {
    const int tiles = (int)cdiv(n, 256);
    int bn = (int)cdiv(n, tiles);
    bn = (bn + 15) / 16 * 16;
    *bn_out = bn;
    *ntiles_out = (int)cdiv(n, bn);
}

bool mlp_chain_layer_ok(int64_t k_total, int64_t n) {
    if (n < 4 || n > 65536 || k_total < 32 || k_total > 65536) return false;
    int bn, nt;
    mlp_chain_tiling(n, &bn, &nt);
    return bn <= mc::BN_MAX;
}

int mlp_chain_launch(mc::Program& prog, cudaStream_t st) {
    if (prog.rows <= 0 || prog.nops <= 0) return 0;
    O4D_REQUIRE(prog.nops <= mc::MAX_OPS, "mlp chain: too many layers in one launch (%d)", prog.nops);
    prog.tiles = (int)cdiv(prog.rows, mc::BM);
    double flops = 0.0;
    for (int o = 0; o < prog.nops; ++o) {
        mc::Op& op = prog.op[o];
        O4D_REQUIRE(op.a1 && op.w && op.k1c >= 1 && op.bn >= 16 && op.bn <= mc::BN_MAX && op.bn % 16 == 0 && op.ntiles >= 1,
                    "mlp chain: bad layer %d", o);
        O4D_REQUIRE(!op.img || (op.n % 32 == 0 && op.img_cpt * 32 == op.n && (!op.bias || ((uintptr_t)op.bias & 15) == 0)),
                    "mlp chain: an image output needs n %% 32 == 0 and a 16-byte aligned bias");
        O4D_REQUIRE(op.img || op.out, "mlp chain: layer %d has no output", o);
        O4D_REQUIRE(!op.img || ((!op.out || (op.ldo % 4 == 0 && ((uintptr_t)op.out & 15) == 0)) &&
                                (!op.res || (op.ldr % 4 == 0 && ((uintptr_t)op.res & 15) == 0))),
                    "mlp chain: layer %d writes an image next to an unaligned fp32 output", o);
        op.n_img = op.img ? op.img_cpt * 32 : 0;
        flops += 2.0 * (double)prog.rows * op.k_alg * op.n;
    }
    {
        static int fg = -1;
        if (fg < 0) {
            const char* e = getenv("O4D_CHAIN_FENCE_GPU");      // 1 = gpu-scope fence before the proxy fence (A/B, paranoia)
            fg = (e && e[0] == '1') ? 1 : 0;
        }
        prog.fence_gpu = fg;
    }
    ProfScope prof(PROF_LINEAR, flops, st);
    if (mlp_chain_pair()) {
        for (int o = 0; o < prog.nops; ++o)
            O4D_REQUIRE(prog.op[o].bn % 16 == 0, "mlp chain: pair kernel needs n-tiles in multiples of 16");
        O4D_SMEM_ATTR(mc::mlp_chain_kernel<true>, mc::SMEM_BYTES);
        const int tile_pairs = (prog.tiles + 1) / 2;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(2 * (tile_pairs < 74 ? tile_pairs : 74)), 1, 1);
        cfg.blockDim = dim3(mc::THREADS, 1, 1);
        cfg.dynamicSmemBytes = mc::SMEM_BYTES;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        O4D_CUDA(cudaLaunchKernelEx(&cfg, mc::mlp_chain_kernel<true>, prog));
        count_launch();
        return 0;
    }
    O4D_SMEM_ATTR(mc::mlp_chain_kernel<false>, mc::SMEM_BYTES);
    int grid = 148;
    if (prog.tiles < grid) grid = prog.tiles;
    mc::mlp_chain_kernel<false><<<grid, mc::THREADS, mc::SMEM_BYTES, st>>>(prog);
    O4D_LAUNCH_CHECK();
    return 0;
}

// ---- ResnetBlockFC as ONE launch (C ABI, include/o4d.h) ------------------------------------------------------------
struct ResblockWs {
    uint8_t *img_x, *img_h, *w0p, *w1p;
    size_t bytes;
};
static ResblockWs resblock_ws(int64_t rows, int d, int dh, void* base, size_t cap, bool* ok) {
    Arena a(base, cap);
    ResblockWs w;
    w.img_x = a.get<uint8_t>(act_image_bytes(rows, d));
    w.img_h = a.get<uint8_t>(act_image_bytes(rows, dh));
    w.w0p = a.get<uint8_t>(tc_pack_bytes(dh, d));
    w.w1p = a.get<uint8_t>(tc_pack_bytes(d, dh));
    w.bytes = a.off;
    if (ok) *ok = a.ok;
    return w;
}

static bool resblock_ok(int64_t rows, int d, int dh) {
    return rows >= 1 && d >= 32 && dh >= 32 && d % 32 == 0 && dh % 32 == 0 && mlp_chain_layer_ok(d, dh) && mlp_chain_layer_ok(dh, d);
}

}  // namespace o4d

extern "C" size_t o4d_resblock_workspace_bytes(int64_t rows, int d, int d_hidden) {
    if (!o4d::resblock_ok(rows, d, d_hidden)) return 0;
    return o4d::resblock_ws(rows, d, d_hidden, nullptr, 0, nullptr).bytes;
}

extern "C" int o4d_resblock_forward_f32(const float* x, int64_t rows, int d, int64_t ldx, const float* w0, const float* b0,
                                        int d_hidden, const float* w1, const float* b1, float* out, int64_t ldo,
                                        int precision, void* workspace, size_t workspace_bytes, void* stream) {
    using namespace o4d;
    cudaStream_t st = (cudaStream_t)stream;
    if (rows == 0) return 0;
    O4D_REQUIRE(x && w0 && w1 && out && workspace, "resblock: null pointer");
    O4D_REQUIRE(precision == 1 || precision == 2, "resblock: the fused kernel is the tcgen05 path (precision 1 or 2)");
    if (!resblock_ok(rows, d, d_hidden)) {
        set_error("resblock: widths must be multiples of 32 within the chain kernel's tile limits (d=%d, d_hidden=%d)", d, d_hidden);
        return O4D_E_UNSUPPORTED;
    }
    O4D_REQUIRE(ldx >= d && ldo >= d && ldx % 4 == 0 && ldo % 4 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)out & 15) == 0 &&
                    (!b0 || ((uintptr_t)b0 & 15) == 0) && (!b1 || ((uintptr_t)b1 & 15) == 0),
                "resblock: rows and biases must be 16-byte aligned");
    bool ok = false;
    ResblockWs w = resblock_ws(rows, d, d_hidden, workspace, workspace_bytes, &ok);
    if (!ok) {
        set_error("resblock: workspace too small (%zu < %zu)", workspace_bytes, w.bytes);
        return O4D_E_WORKSPACE;
    }
    O4D_TRY(mlp_chain_pack_launch(w0, d_hidden, d, d, w.w0p, st));
    O4D_TRY(mlp_chain_pack_launch(w1, d, d_hidden, d_hidden, w.w1p, st));
    O4D_TRY(act_image_launch(x, ldx, rows, d, 1, w.img_x, st));                 // implicit.py:93  act(x)
    mc::Program prog;
    prog.nops = 2;
    prog.rows = rows;
    prog.split = precision == 1 ? 1 : 0;
    prog.tiles = 0;
    mc::Op& f0 = prog.op[0];                                                    // net = fc_0(act(x)); image of act(net)
    f0 = mc::Op{};
    f0.a1 = w.img_x; f0.k1c = d / 32; f0.w = w.w0p; f0.bias = b0; f0.img = w.img_h; f0.img_cpt = d_hidden / 32; f0.img_relu = 1;
    f0.n = d_hidden; f0.k_alg = d;
    mlp_chain_tiling(d_hidden, &f0.bn, &f0.ntiles);
    mc::Op& f1 = prog.op[1];                                                    // x + fc_1(act(net))   (implicit.py:94-101)
    f1 = mc::Op{};
    f1.a1 = w.img_h; f1.k1c = d_hidden / 32; f1.w = w.w1p; f1.bias = b1; f1.out = out; f1.ldo = ldo; f1.res = x; f1.ldr = ldx;
    f1.n = d; f1.k_alg = d_hidden;
    mlp_chain_tiling(d, &f1.bn, &f1.ntiles);
    return mlp_chain_launch(prog, st);
}

extern "C" int o4d_debug_read_chain(long long* out8) {
    return (int)cudaMemcpyFromSymbol(out8, o4d::mc::g_dbg_chain, sizeof(long long) * 8);
}
