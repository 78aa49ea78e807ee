// tcgen05 dense layer:  C = post(pre(A) @ W^T + bias) [+ R]   with fp32 tensors in HBM.
//
// The reference runs every nn.Linear as a cuBLAS fp32 SGEMM.  Here the contraction runs on
// the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM) while keeping
// fp32-grade results: both operands are split on the fly into bf16 hi + bf16 lo
// (x = hi + lo + O(2^-17 |x|)) and three MMAs accumulate hi*hi + lo*hi + hi*lo into the
// same fp32 TMEM accumulator ("bf16x3", precision 1).  precision 2 issues only hi*hi.
//
// Per CTA: one 128 x BN output tile (BN <= 256 TMEM columns), K walked in chunks of 32.
//   warps 0-7  A producers: coalesced fp32 loads (optionally ReLU), hi/lo split, 16-byte
//              stores into the K-major no-swizzle core-matrix layout the UMMA descriptor
//              describes; after the main loop the same warps run the epilogue
//              (tcgen05.ld 32x32b -> bias / ReLU / residual -> global): warp w owns TMEM lane
//              quarter w & 3 and column half w >> 2.  (In-kernel cycle stamps showed the main loop
//              producer-bound -- identical with one MMA pass instead of three -- and the epilogue as
//              long as 60 % of it, so both run on eight warps instead of four.)
//   warp 8     allocates TMEM; lane 0 streams the pre-packed weight tiles (hi+lo image of a
//              BN x 32 slab, already in shared-memory layout) with cp.async.bulk (TMA engine,
//              mbarrier complete_tx).
//   warp 9     lane 0 issues tcgen05.mma and commits to the stage's "empty" mbarrier.
// Two CTAs are resident per SM (2 x ~97 KB smem, 2 x 256 TMEM columns), so one CTA's
// epilogue overlaps the other's main loop.
#include "o4d_common.cuh"
#include <cuda_bf16.h>
#include "tc_helpers.cuh"
#include <stdlib.h>

namespace o4d {
namespace tc {

constexpr int BM = 128;
constexpr int BK = 32;
constexpr int STAGES = 2;
constexpr int PROD_WARPS = 8;
constexpr int THREADS = (PROD_WARPS + 2) * 32;
constexpr int A_HALF_BYTES = BM * BK * 2;          // one bf16 image of the A slab (8 KB)
constexpr int BN_MAX = 256;
constexpr int STAGE_BYTES = 2 * A_HALF_BYTES + 2 * BN_MAX * BK * 2;  // 48 KB
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t a, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t a) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t a, uint32_t tx) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(tx) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t a, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug traps (launch error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t a, uint32_t parity) {
    if (mbar_try_wait(a, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(a, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(mbar)
                 : "memory");
}

// K-major, no-swizzle shared-memory matrix descriptor: 8 x 16 B core matrices; LBO = byte
// distance between the two core matrices of one K=16 step, SBO = distance between 8-row groups.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
    return d;                // layout_type 0 = no swizzle, base_offset 0
}

// kind::f16 instruction descriptor: D fp32, A/B bf16, both K-major, M = 128.
__device__ __forceinline__ uint32_t umma_idesc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}


__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(x);
    lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

__device__ long long g_dbg_tc[16];   // cycle stamps of one CTA (diagnostics)

struct PackMeta {
    int n, k;      // logical weight shape
    int bn;        // tile width (multiple of 16, <= 256)
    int ntiles;    // ceil(n / bn)
    int kchunks;   // ceil(k / 32)
};

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

__host__ __device__ inline size_t pack_bytes(const PackMeta& m) {
    return (size_t)m.ntiles * m.kchunks * 2 * m.bn * BK * 2;
}

// W (n, k) fp32 row-major (ldw) -> per (n-tile, k-chunk): [hi image][lo image], each
// [kc = 4][row-group = bn/8][8 rows][8 bf16] -- exactly the shared-memory image of the slab.
__global__ void pack_weight_kernel(const float* __restrict__ W, int64_t ldw, PackMeta m, __nv_bfloat16* __restrict__ out) {
    const int64_t slab_elems = (int64_t)m.bn * BK;
    const int64_t total = (int64_t)m.ntiles * m.kchunks * slab_elems;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t slab = e / slab_elems;
        const int within = (int)(e % slab_elems);
        const int t = (int)(slab / m.kchunks), c = (int)(slab % m.kchunks);
        const int kc = within / (m.bn * 8);
        const int rem = within % (m.bn * 8);
        const int row = rem / 8, el = rem % 8;      // row = rg*8 + r8
        const int gn = t * m.bn + row, gk = c * BK + kc * 8 + el;
        float v = (gn < m.n && gk < m.k) ? W[(int64_t)gn * ldw + gk] : 0.f;
        __nv_bfloat16 hi, lo;
        split_bf16(v, hi, lo);
        out[slab * 2 * slab_elems + within] = hi;
        out[slab * 2 * slab_elems + slab_elems + within] = lo;
    }
}

// Pair format (mlp_chain.cu, cta_group::2): per (n-tile, k-chunk) [half 0: hi, lo][half 1: hi, lo], half h = rows
// [h * bn/2, (h+1) * bn/2) of the tile, each image [kc = 4][row-group][8 rows][8 bf16]: exactly what CTA h of a pair
// copies into its shared memory.  Same total size as the single-CTA format.
__global__ void pack_weight_pair_kernel(const float* __restrict__ W, int64_t ldw, PackMeta m, __nv_bfloat16* __restrict__ out) {
    const int64_t slab_elems = (int64_t)m.bn * BK;
    const int64_t total = (int64_t)m.ntiles * m.kchunks * slab_elems;
    const int hb = m.bn / 2;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t slab = e / slab_elems;
        const int within = (int)(e % slab_elems);
        const int t = (int)(slab / m.kchunks), c = (int)(slab % m.kchunks);
        const int half = within / (hb * BK);
        const int w2 = within % (hb * BK);
        const int kc = w2 / (hb * 8);
        const int rem = w2 % (hb * 8);
        const int row = rem / 8, el = rem % 8;
        const int gn = t * m.bn + half * hb + row, gk = c * BK + kc * 8 + el;
        float v = (gn < m.n && gk < m.k) ? W[(int64_t)gn * ldw + gk] : 0.f;
        __nv_bfloat16 hi, lo;
        split_bf16(v, hi, lo);
        __nv_bfloat16* dst = out + slab * 2 * slab_elems + (int64_t)half * 2 * hb * BK;
        dst[w2] = hi;
        dst[hb * BK + w2] = lo;
    }
}

__global__ void __launch_bounds__(THREADS, 2)
linear_tc_kernel(const float* __restrict__ A, int64_t rows, int k, int64_t lda, const __nv_bfloat16* __restrict__ Wp,
                 PackMeta m, const float* __restrict__ bias, const float* R, int64_t ldr, float* C, int64_t ldc,
                 int flags, int split, RowGather g) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);  // full[2], empty[2], accum
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // 1-D grid, column tile fastest: the CTAs that share a 128-row slab of A are adjacent in launch order, so A comes
    // from HBM once and from L2 for the other column tiles (with the column tile in blockIdx.y they were a whole grid
    // row apart -- 400 MB for the attention MLP of a training frame -- and A was re-read from HBM per column tile).
    const int64_t row0 = (int64_t)(blockIdx.x / m.ntiles) * BM;
    const int tile_n = blockIdx.x % m.ntiles;
    const int bn = m.bn;
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[STAGES]), accum_bar = smem_u32(&bars[2 * STAGES]);
    const uint32_t smem_base = smem_u32(smem);

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full0 + 8 * s, PROD_WARPS + 1);   // producer warps + the weight-copy thread
            mbar_init(empty0 + 8 * s, 1);  // one tcgen05.commit
        }
        mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == PROD_WARPS) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int nchunks = m.kchunks;
    const uint32_t b_half_bytes = (uint32_t)bn * BK * 2;

    if (warp < PROD_WARPS) {
        // ------------------------------------------------------------ A producers
        // The fp32 -> bf16 hi/lo conversion needs the data in registers, so the global loads of
        // chunk c+1 are issued BEFORE chunk c is converted and stored: their latency hides behind
        // the stage wait, the conversion and the shared-memory stores of the current chunk.
        const bool relu_in = flags & O4D_RELU_IN;
        // lane -> (row rr of the 8-row group, quarter p): floats [4p, 4p+4) and [16+4p, 16+4p+4) of the row's
        // 32-float chunk, so every LDG.128 of the warp covers 8 rows x 64 contiguous bytes = 16 FULL sectors.
        // (The earlier mapping -- 8 consecutive floats per lane -- used half of each 32-byte sector per
        // instruction; an ablation showed the main loop at 36 k cycles with and 15 k without the A loads,
        // independent of warps / CTAs per SM, i.e. bound by outstanding L1 sector requests.)
        const int rr = lane >> 2, pq = lane & 3;
        constexpr int GPW = (BM / 8) / PROD_WARPS;      // 8-row groups per producer warp (2)
        const int k1c = (k + BK - 1) / BK;              // chunks fed by A; the rest (if any) by the K-concatenated g.a2
        auto load_chunk = [&](int c, float (&v)[GPW][8]) {
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
        const bool dbg = O4D_STAMPS && (blockIdx.x == (gridDim.x / 2 / m.ntiles) * m.ntiles) && threadIdx.x == 0;
        if (dbg) g_dbg_tc[0] = clock64();
        load_chunk(0, cur);
        for (int c = 0; c < nchunks; ++c) {
            const int s = c % STAGES;
            const uint32_t ph = (uint32_t)(c / STAGES) & 1u;
            if (c + 1 < nchunks) load_chunk(c + 1, nxt);
            mbar_wait(empty0 + 8 * s, ph ^ 1u);
            uint8_t* a_hi = smem + s * STAGE_BYTES;
            uint8_t* a_lo = a_hi + A_HALF_BYTES;
#pragma unroll
            for (int g = 0; g < GPW; ++g) {
                const int rg = warp * GPW + g;                // 8-row group inside the 128-row tile
                __align__(16) uint32_t h[4];
                __align__(16) uint32_t l[4];
                const bool relu_c = relu_in && c < k1c;
#pragma unroll
                for (int i = 0; i < 4; ++i) {   // packed conversions (full rate), see tch::split_bf16x2
                    const float x0 = relu_c ? fmaxf(cur[g][2 * i], 0.f) : cur[g][2 * i];
                    const float x1 = relu_c ? fmaxf(cur[g][2 * i + 1], 0.f) : cur[g][2 * i + 1];
                    if (split) {
                        tch::split_bf16x2(x0, x1, h[i], l[i]);
                    } else {                    // single-pass bf16 (precision 2): no low-order image at all
                        const __nv_bfloat162 hb = __floats2bfloat162_rn(x0, x1);
                        h[i] = *reinterpret_cast<const uint32_t*>(&hb);
                        l[i] = 0u;
                    }
                }
                // [kc][row group][row][16 B]: floats 4p..4p+3 are half (p & 1) of core-matrix line kc = p >> 1,
                // floats 16+4p.. the same half of line kc + 2
                const int off = (pq >> 1) * (BM * 16) + rg * 128 + rr * 16 + (pq & 1) * 8;
                *reinterpret_cast<uint2*>(a_hi + off) = *reinterpret_cast<const uint2*>(h);
                *reinterpret_cast<uint2*>(a_hi + off + 2 * (BM * 16)) = *reinterpret_cast<const uint2*>(h + 2);
                if (split) {
                    *reinterpret_cast<uint2*>(a_lo + off) = *reinterpret_cast<const uint2*>(l);
                    *reinterpret_cast<uint2*>(a_lo + off + 2 * (BM * 16)) = *reinterpret_cast<const uint2*>(l + 2);
                }
            }
            fence_proxy_async_smem();   // generic-proxy stores -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(full0 + 8 * s);
#pragma unroll
            for (int g = 0; g < GPW; ++g)
#pragma unroll
                for (int i = 0; i < 8; ++i) cur[g][i] = nxt[g][i];
        }
        // ------------------------------------------------------------ epilogue
        // TMEM hands every thread one ROW (32x32b shape); storing C that way touches 32 cache
        // lines per instruction and made the epilogue 2x longer than the main loop.  Each warp
        // instead transposes 32 x 32 blocks through a padded shared-memory tile (the pipeline
        // stages are idle by now): in the write phase a lane owns one COLUMN, so every load of
        // the residual and every store of C is one contiguous 128-byte row segment, and the loads
        // of eight rows are issued back to back before they are consumed.
        if (dbg) g_dbg_tc[1] = clock64();
        mbar_wait(accum_bar, 0);
        tc_fence_after();
        if (dbg) g_dbg_tc[2] = clock64();
        const bool relu_out = flags & O4D_RELU_OUT;
        const bool mask_res = flags & O4D_MASK_RES;         // R is a ReLU mask, not an addend
        const int quarter = warp & 3, chalf = warp >> 2;
        const int64_t warp_row0 = row0 + quarter * 32;
        const int col_base = tile_n * bn;
        const uint32_t taddr_row = tmem_base + ((uint32_t)(quarter * 32) << 16);
        // column range of this warp: 32-column blocks [0, nblk/2 rounded up) or the rest
        const int nblk = (bn + 31) / 32;
        const int c_begin = chalf ? ((nblk + 1) / 2) * 32 : 0;
        const int c_end = chalf ? bn : min(bn, ((nblk + 1) / 2) * 32);
        constexpr int SLD = 36;                            // staging row pitch (floats): 16-byte aligned rows, conflict-free
        float* stg = reinterpret_cast<float*>(smem) + warp * (32 * SLD);
        const int rows_here = (int)min((int64_t)32, rows - warp_row0);     // may be <= 0 for a ragged last tile
        // the row-gather term Qa[row / K] - Ka[nbr[row]] rides in the registers of the residual (never both at once here)
        const bool gather = g.qa != nullptr;
        const bool vec_ok = (m.n % 4 == 0) && (ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0) &&
                            (!R || ((ldr % 4 == 0) && ((reinterpret_cast<uintptr_t>(R) & 15) == 0))) &&
                            (!bias || ((reinterpret_cast<uintptr_t>(bias) & 15) == 0)) &&
                            (!gather || (!R && ((reinterpret_cast<uintptr_t>(g.qa) & 15) == 0) &&
                                         ((reinterpret_cast<uintptr_t>(g.ka) & 15) == 0)));
        if (vec_ok) {
            // write phase: lane -> (row-in-group rr, 4 columns c4): one instruction covers 4 rows x 128 B
            const int rr = lane >> 3, c4 = (lane & 7) * 4;
            float4 resn[8];
            auto load_res = [&](int c0) {
                const int gc = col_base + c0 + c4;
                const bool ok = (R || gather) && c4 < min(32, bn - c0) && gc < m.n;
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int row = it * 4 + rr;
                    resn[it] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (ok && row < rows_here) {
                        if (gather) {   // (row -> Qa / Ka rows recomputed per column block: keeping them would spill)
                            const int64_t ar = g.row_offset + warp_row0 + row;
                            const float4 a = __ldg(reinterpret_cast<const float4*>(g.qa + (ar / g.knbr) * m.n + gc));
                            const float4 b = __ldg(reinterpret_cast<const float4*>(g.ka + g.neighbour(ar) * m.n + gc));
                            resn[it] = make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
                        } else {
                            resn[it] = *reinterpret_cast<const float4*>(R + (warp_row0 + row) * ldr + gc);
                        }
                    }
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
                            if (gather) { x.x += res[it].x; x.y += res[it].y; x.z += res[it].z; x.w += res[it].w; }
                            if (relu_out) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f); }
                            if (gather) {
                                // already added in front of the activation
                            } else if (mask_res) {
                                x.x = res[it].x > 0.f ? x.x : 0.f; x.y = res[it].y > 0.f ? x.y : 0.f;
                                x.z = res[it].z > 0.f ? x.z : 0.f; x.w = res[it].w > 0.f ? x.w : 0.f;
                            } else {
                                x.x += res[it].x; x.y += res[it].y; x.z += res[it].z; x.w += res[it].w;
                            }
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
            const float* gk = nullptr;
            if (g.qa && grow < rows) {
                const int64_t ar = g.row_offset + grow;
                gq = g.qa + (ar / g.knbr) * m.n;
                gk = g.ka + g.neighbour(ar) * m.n;
            }
            for (int c0 = c_begin; c0 < c_end; c0 += 16) {
                float v[16];
                tmem_ld16(taddr_row + (uint32_t)c0, v);
                if (grow < rows) {
                    const int gc0 = col_base + c0;
                    for (int i = 0; i < 16 && gc0 + i < m.n; ++i) {
                        float x = v[i] + (bias ? bias[gc0 + i] : 0.f);
                        if (gq) x += gq[gc0 + i] - gk[gc0 + i];
                        if (relu_out) x = fmaxf(x, 0.f);
                        if (R) x = mask_res ? (R[grow * ldr + gc0 + i] > 0.f ? x : 0.f) : x + R[grow * ldr + gc0 + i];
                        C[grow * ldc + gc0 + i] = x;
                    }
                }
            }
        }
        if (dbg) g_dbg_tc[3] = clock64();
        tc_fence_before();
    } else if (warp == PROD_WARPS) {
        // ------------------------------------------------------------ weight slabs via the TMA engine
        if (lane == 0) {
            const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(Wp) + (size_t)tile_n * nchunks * 2 * b_half_bytes;
            for (int c = 0; c < nchunks; ++c) {
                const int s = c % STAGES;
                const uint32_t ph = (uint32_t)(c / STAGES) & 1u;
                mbar_wait(empty0 + 8 * s, ph ^ 1u);
                const uint32_t dst = smem_base + s * STAGE_BYTES + 2 * A_HALF_BYTES;
                const uint32_t wbytes = split ? 2 * b_half_bytes : b_half_bytes;     // [hi][lo] per chunk: hi only at precision 2
                mbar_arrive_expect_tx(full0 + 8 * s, wbytes);
                bulk_g2s(dst, wsrc + (size_t)c * 2 * b_half_bytes, wbytes, full0 + 8 * s);
            }
        }
    } else {
        // ------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            const uint32_t idesc = umma_idesc(bn);
            const uint32_t lbo_a = BM * 16, lbo_b = (uint32_t)bn * 16;
            for (int c = 0; c < nchunks; ++c) {
                const int s = c % STAGES;
                const uint32_t ph = (uint32_t)(c / STAGES) & 1u;
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
                    umma_f16(tmem_base, da_hi, db_hi, idesc, (c | ks) ? 1u : 0u);
                    if (split) {
                        const uint64_t da_lo = umma_desc(a_lo + ks * 2 * lbo_a, lbo_a, 128);
                        const uint64_t db_lo = umma_desc(b_lo + ks * 2 * lbo_b, lbo_b, 128);
                        umma_f16(tmem_base, da_lo, db_hi, idesc, 1u);
                        umma_f16(tmem_base, da_hi, db_lo, idesc, 1u);
                    }
                }
                umma_commit(empty0 + 8 * s);          // frees the stage when these MMAs retire
            }
            umma_commit(accum_bar);                   // accumulator complete -> epilogue
        }
    }
    __syncthreads();
    if (warp == PROD_WARPS) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
    }
}

}  // namespace tc

size_t tc_pack_bytes(int64_t n, int64_t k) { return tc::pack_bytes(tc::pack_meta((int)n, (int)k)); }

int tc_pack_launch(const float* W, int64_t n, int64_t k, int64_t ldw, void* packed, cudaStream_t st) {
    tc::PackMeta m = tc::pack_meta((int)n, (int)k);
    const int64_t total = (int64_t)m.ntiles * m.kchunks * m.bn * tc::BK;
    int64_t blocks = cdiv(total, 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    tc::pack_weight_kernel<<<(unsigned)blocks, 256, 0, st>>>(W, ldw, m, (__nv_bfloat16*)packed);
    O4D_LAUNCH_CHECK();
    return 0;
}

int tc_pack_pair_launch(const float* W, int64_t n, int64_t k, int64_t ldw, void* packed, cudaStream_t st) {
    tc::PackMeta m = tc::pack_meta((int)n, (int)k);
    const int64_t total = (int64_t)m.ntiles * m.kchunks * m.bn * tc::BK;
    int64_t blocks = cdiv(total, 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    tc::pack_weight_pair_kernel<<<(unsigned)blocks, 256, 0, st>>>(W, ldw, m, (__nv_bfloat16*)packed);
    O4D_LAUNCH_CHECK();
    return 0;
}

// n >= 8: narrow outputs (lin_out, 9 .. 33 columns) run as one 16/32/48-column UMMA tile; the CUDA-core kernel
// took as long for 416 -> 9 as the tensor-core kernel for 416 -> 416 (69 us per 32768 rows).
bool tc_shape_ok(int64_t rows, int64_t k, int64_t n) { return rows >= 512 && k >= 32 && n >= 4 && k <= 65536 && n <= 65536; }

int linear_tc_packed_launch(const float* A, int64_t rows, int64_t k, int64_t lda, const void* packed, int64_t n,
                            const float* bias, const float* R, int64_t ldr, float* C, int64_t ldc, int flags,
                            int precision, cudaStream_t st, const RowGather* gp) {
    if (rows == 0) return 0;
    RowGather g;
    if (gp) g = *gp;
    O4D_SMEM_ATTR(tc::linear_tc_kernel, tc::SMEM_BYTES);
    // with a K-concatenated second operand the packed weight spans round32(k) + k2 columns
    const int64_t ktot = g.a2 ? cdiv(k, tc::BK) * tc::BK + g.k2 : k;
    tc::PackMeta m = tc::pack_meta((int)n, (int)ktot);
    dim3 grid((unsigned)(cdiv(rows, tc::BM) * m.ntiles));
    tc::linear_tc_kernel<<<grid, tc::THREADS, tc::SMEM_BYTES, st>>>(A, rows, (int)k, lda, (const __nv_bfloat16*)packed, m, bias, R,
                                                                    ldr, C, ldc, flags, precision == 1 ? 1 : 0, g);
    O4D_LAUNCH_CHECK();
    return 0;
}

// Un-packed entry: packs the weight into a stream-ordered temporary first (generic C-ABI
// calls; the decoder keeps its weights packed in the scene buffer instead).
int linear_tc_launch(const float* A, int64_t rows, int64_t k, int64_t lda, const float* W, const float* bias,
                     int64_t n, const float* R, int64_t ldr, float* C, int64_t ldc, int flags, int precision,
                     cudaStream_t st, const RowGather* g) {
    if (!tc_shape_ok(rows, k, n)) return O4D_E_UNSUPPORTED;
    void* packed = nullptr;
    O4D_CUDA(cudaMallocAsync(&packed, tc_pack_bytes(n, k), st));
    int rc = tc_pack_launch(W, n, k, k, packed, st);
    if (rc == 0) rc = linear_tc_packed_launch(A, rows, k, lda, packed, n, bias, R, ldr, C, ldc, flags, precision, st, g);
    cudaFreeAsync(packed, st);
    return rc;
}

}  // namespace o4d

extern "C" int o4d_has_tcgen05(void) { return 1; }
extern "C" int o4d_debug_read_tc(long long* out16) {
    return (int)cudaMemcpyFromSymbol(out16, o4d::tc::g_dbg_tc, sizeof(long long) * 16);
}
