// Fused vector-attention kernel (tcgen05 / TMEM), one 128-row tile of (query, neighbour)
// pairs per CTA.  It replaces, for one cross-attention layer of the decoder
// (model/point_transformer_layer.py:174-179 as restructured in attn.cu):
//
//     r      = relu(W_p1 (p_i - p2_j) + b_p1)                       CUDA cores, 32 wide
//     hidden = relu(Wc r + Qa_i - Ka_j)           K = 32 MMA  + gather epilogue, 2d wide
//     logits = W_a2 hidden                        K = 2d MMA, accumulated in TMEM (d columns)
//     delta  = W_p2 r + b_p2                      K = 32 MMA
//     out_i  = sum_j softmax_j(logits / sqrt(d)) * (V_j + delta)    per channel
//
// (b_a2 is constant over j and cancels in the per-channel softmax, so it is not added.)
// Nothing but Qa (n, 2d) in and out (n, d) out touches HBM: the reference's (n, k, d) and
// (n, k, 2d) intermediates live in TMEM / shared memory only.
//
// The hidden layer is walked in chunks of 32 units.  For chunk c the tensor core computes
// acc1 = r Wc[c]^T (128 x 32 fp32, TMEM); the eight row warps pull it into registers, add the
// staged Qa_i - Ka_j slice, apply ReLU, split to bf16 hi/lo and write the result back into
// the SAME tensor-memory columns, where MMA2 reads it as its A operand (tcgen05.mma with A in
// TMEM) and accumulates it against W_a2[:, c] into the (128 x d) logits accumulator.  The acc1 /
// hidden buffers form a ring of three.  Weights arrive as pre-packed shared-memory images
// through cp.async.bulk (two 53 KB W_a2 stages + a ring of four 4 KB Wc stages); the Qa / Ka / V
// slices through cooperative cp.async into two gather stages.  All contractions use the bf16x3
// split (hi*hi + lo*hi + hi*lo, fp32 accumulate) unless `split` is 0.
//
// Warps: 0-15 row warps = two groups of eight (TMEM lane quarter w & 3, column half (w >> 2) & 1, group w >> 3) that
// convert ALTERNATE hidden chunks, 16 weight stream, 17 MMA2 issuer (both join the softmax reduce of the epilogue),
// 18 MMA1 / delta issuer.
// TMEM: columns [0, d) logits, [416, 512) the three acc1 / hidden buffers.  One CTA per SM.
// How the kernel got here (what bound it at each step, measured): DESIGN.md section 4.
#include "o4d_common.cuh"
#include <cuda_bf16.h>
#include <stdlib.h>

namespace o4d {
namespace fa {

constexpr int BM = 128;
constexpr int HC = 32;          // hidden units per chunk == K of the second contraction per chunk
constexpr int GROUP_WARPS = 8;  // one conversion group: two warps per TMEM lane quarter
constexpr int ROW_GROUPS = 2;   // groups convert alternate hidden chunks (chunk c belongs to group c & 1)
constexpr int ROW_WARPS = GROUP_WARPS * ROW_GROUPS;
constexpr int ROW_THREADS = ROW_WARPS * 32;
constexpr int THREADS = (ROW_WARPS + 3) * 32;   // 16 row warps, weight-stream warp, MMA2 issuer, MMA1 / delta issuer
constexpr int R_BYTES = BM * 32 * 2;         // one bf16 image of a (128 x 32) A operand
constexpr int NBUF = 3;                      // acc1 (TMEM) / hidden A2 (smem) buffers: MMA1 runs two chunks ahead
constexpr int WC_STAGES = 4;                 // ring of Wc chunks (4 KB each)
constexpr int OFF_R = 0;                     // r hi, r lo
constexpr int OFF_A2 = 2 * R_BYTES;          // NBUF x (hi, lo)
constexpr int WC_BYTES = HC * 32 * 2;        // one bf16 image of a Wc chunk (32 x 32)
constexpr int OFF_WC = OFF_A2 + NBUF * 2 * R_BYTES;
constexpr int OFF_W = OFF_WC + WC_STAGES * 2 * WC_BYTES;   // two stages of W_a2 chunks
constexpr int ACC1_COL = 416;                // TMEM: [0, d) logits, [416, 512) three acc1 buffers
constexpr int STG_LD = 33;                   // padded row pitch (floats) of the epilogue staging tiles
// Gather staging: per hidden chunk the 32-float slices of Ka (one per tile row) and Qa (one per
// query of the tile) are copied into shared memory with cp.async by lanes that cooperate on a row
// (8 lanes x 16 B = one 128-byte line per row, 4 rows per warp instruction).  A row-per-thread
// gather straight from global costs one L1 wavefront per lane per 16 bytes; ncu showed the LSU
// data pipe at 74 % with the tensor pipe at 40 % (profiles/r1_i_*).  Row pitch 144 B: the
// row-per-thread LDS.128 reads that follow are bank-conflict free.
constexpr int G_PITCH = 144;
constexpr int G_ROWS = BM + 16;              // 128 Ka rows + up to 16 Qa rows (k >= 8)
constexpr int G_STAGE = G_ROWS * G_PITCH;    // 20,736 B
constexpr int G_STAGES = 2;                  // per group; group 1's pair lives in the (idle during the main loop) hidden-buffer region
constexpr int BAR_BYTES = 320;               // 31 mbarriers + the TMEM slot
constexpr int TAIL_BYTES = BAR_BYTES + BM * 4;   // then the tile's neighbour ids

__host__ __device__ inline int w2_bytes(int d) { return 2 * d * 32 * 2; }             // W_a2 chunk, hi + lo
__host__ __device__ inline int wstage_bytes(int d) { return 2 * WC_BYTES + w2_bytes(d); }  // packed chunk in HBM
__host__ __device__ inline int off_g(int d) { return OFF_W + 2 * w2_bytes(d); }
__host__ __device__ inline int smem_bytes(int d) { return off_g(d) + G_STAGES * G_STAGE + TAIL_BYTES; }

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
__device__ __forceinline__ void mbar_wait(uint32_t a, uint32_t parity) {
    if (mbar_try_wait(a, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(a, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();   // protocol bug: fail the launch, do not hang
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
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
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
// A operand from tensor memory (lane = row, two bf16 per 32-bit column), B from shared memory
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
// arrive on `mbar` (without raising its pending count) once every cp.async this thread issued so far has landed
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint32_t mbar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// 2^x for x <= 0 (softmax weights): single MUFU, flush-to-zero below 2^-126
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(x);
    lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// Two values at once: ONE packing convert (F2FP.BF16.F32.PACK_AB, full rate) per pair and image half instead of two
// scalar F2F conversions (quarter-rate conversion pipe); the bf16 -> fp32 back-conversion is a shift / a mask.
// Same round-to-nearest-even results as split_bf16.  Low 16 bits = a, high 16 bits = b.
__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t& hi2, uint32_t& lo2) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    hi2 = *reinterpret_cast<const uint32_t*>(&h);
    const float ha = __uint_as_float(hi2 << 16), hb = __uint_as_float(hi2 & 0xffff0000u);
    const __nv_bfloat162 l = __floats2bfloat162_rn(a - ha, b - hb);
    lo2 = *reinterpret_cast<const uint32_t*>(&l);
}

// 32 fp32 values of one row -> bf16 hi/lo images of a (128 x 32) K-major A operand.
__device__ __forceinline__ void store_a_row(uint8_t* hi_img, uint8_t* lo_img, int row, const float* v) {
#pragma unroll
    for (int kc = 0; kc < 4; ++kc) {
        __align__(16) __nv_bfloat16 h[8];
        __align__(16) __nv_bfloat16 l[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) split_bf16(v[kc * 8 + e], h[e], l[e]);
        const int off = kc * (BM * 16) + (row >> 3) * 128 + (row & 7) * 16;
        *reinterpret_cast<uint4*>(hi_img + off) = *reinterpret_cast<const uint4*>(h);
        *reinterpret_cast<uint4*>(lo_img + off) = *reinterpret_cast<const uint4*>(l);
    }
}

// 16 fp32 values (columns [16*half, 16*half+16) of a 32-wide K slab) of one row.
__device__ __forceinline__ void store_a_half_row(uint8_t* hi_img, uint8_t* lo_img, int row, int half, const float* v) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int kc = half * 2 + q;
        uint32_t h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) split_bf16x2(v[q * 8 + 2 * e], v[q * 8 + 2 * e + 1], h[e], l[e]);
        const int off = kc * (BM * 16) + (row >> 3) * 128 + (row & 7) * 16;
        *reinterpret_cast<uint4*>(hi_img + off) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(lo_img + off) = make_uint4(l[0], l[1], l[2], l[3]);
    }
}

__device__ long long g_dbg[16];   // cycle stamps of one tile (diagnostics, read by o4d_debug_read)

constexpr int EPI_WARPS = ROW_WARPS + 2;     // the weight-stream warp and the MMA2 warp join the softmax reduce
constexpr int EPI_THREADS = EPI_WARPS * 32;
static_assert(G_STAGES * G_STAGE <= NBUF * 2 * R_BYTES, "group 1's gather stages must fit the idle hidden-buffer region");
constexpr int EPI_GROUP = 2;                 // chunks staged per barrier pair

// Per-channel softmax over a query's k staged rows and the weighted sum of (V + delta):
// lane = channel inside the 32-channel chunk.  tile_l / tile_v: staging tiles [row][STG_LD].
template <int KT>
__device__ __forceinline__ void reduce_task(const float* tile_l, const float* tile_v, int q, int k, int lane,
                                            float* out_row, const float* bias) {
    constexpr int KV = KT ? KT : O4D_MAX_K;
    const float* lq = tile_l + (q * k) * STG_LD + lane;
    const float* vq = tile_v + (q * k) * STG_LD + lane;
    // Four independent accumulation chains (j = 0, 4, 8, ... / 1, 5, ... / ...): the serial max -> exp -> sum chain over
    // 14 neighbours ran at an IPC of ~0.4 with 4.5 warps per scheduler (in-kernel counters: 8 k cycles of reduce per tile).
    float lgt[KV], val[KV];
    float m4[4] = {-3.4e38f, -3.4e38f, -3.4e38f, -3.4e38f};
#pragma unroll
    for (int jj = 0; jj < KV; ++jj)
        if (jj < k) {
            lgt[jj] = lq[jj * STG_LD];
            val[jj] = vq[jj * STG_LD];
            m4[jj & 3] = fmaxf(m4[jj & 3], lgt[jj]);
        }
    const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
    float d4[4] = {0.f, 0.f, 0.f, 0.f}, n4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int jj = 0; jj < KV; ++jj)
        if (jj < k) {
            const float w = ex2_approx(lgt[jj] - mx);
            d4[jj & 3] += w;
            n4[jj & 3] = fmaf(w, val[jj], n4[jj & 3]);
        }
    const float den = (d4[0] + d4[1]) + (d4[2] + d4[3]);
    const float num = (n4[0] + n4[1]) + (n4[2] + n4[3]);
    // (Writing the bf16 hi / lo activation image of the next layer from here -- 4-byte stores scattered over the image --
    // was measured: +1.0 ms per step in this reduce against 0.9 ms for the separate conversion pass it saves.  Not kept.)
    out_row[lane] = __fdividef(num, den) + bias[lane];
}

struct Params {
    const float* pos; int64_t ldpos;       // query coordinates (n, ldpos)
    const float* pos2; int64_t ldpos2;     // key coordinates (m, ldpos2)
    const int32_t* nbr;                    // (n, k)
    const float* qa;                       // (n, 2d)
    const float* ka;                       // (m, 2d)
    const float* vtab;                     // (m, d)
    const float* wp1; const float* bp1;    // (32, 3), (32)
    const float* bp2;                      // (d)
    const uint8_t* wmain;                  // per hidden chunk: [Wc hi][Wc lo][W_a2 hi][W_a2 lo]
    const uint8_t* wp2;                    // [W_p2 hi][W_p2 lo]   (d rows x 32)
    float* out;                            // (n, d)
    int64_t n;
    int d, k, tq, split;
    float scale_log2;                      // log2(e) / sqrt(d): softmax evaluated with exp2
};

template <int KT>   // KT = neighbours per query when known at compile time (0: use p.k)
__global__ void __launch_bounds__(THREADS, 1) attn_fused_kernel(const Params p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int d = p.d;
    const int k = KT ? KT : p.k;
    const int w2 = w2_bytes(d);
    const int wpacked = wstage_bytes(d);
    const int NC = 2 * d / HC;              // hidden chunks
    const int ND = d / 32;                  // output-column chunks of the softmax epilogue
    const int ntile = d > 256 ? 2 : 1;      // MMA2 is issued per n-tile of dn <= 256 columns
    const int dn = d / ntile;
    const int OFF_G = off_g(d);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_G + G_STAGES * G_STAGE);
    enum { W_FULL = 0, W_EMPTY = 2, WC_FULL = 4, WC_EMPTY = 8, ACC1_FULL = 12, A2_FULL = 15, A2_EMPTY = 18,
           R_READY = 21, ACC2_FULL = 22, G_FULL = 23, G_EMPTY = 27, NBARS = 31 };   // G_*: [group][stage]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBARS);
    int32_t* s_j = reinterpret_cast<int32_t*>(smem + OFF_G + G_STAGES * G_STAGE + BAR_BYTES);   // neighbour id per tile row
    const uint32_t bar0 = smem_u32(bars);
#define BAR(i) (bar0 + 8u * (uint32_t)(i))
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem_base = smem_u32(smem);

    if (O4D_STAMPS && threadIdx.x == 0 && blockIdx.x == gridDim.x / 2) { g_dbg[4] = clock64(); g_dbg[9] = 0; }
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(BAR(W_FULL + i), 1); mbar_init(BAR(W_EMPTY + i), 1); }
        for (int i = 0; i < WC_STAGES; ++i) { mbar_init(BAR(WC_FULL + i), 1); mbar_init(BAR(WC_EMPTY + i), 1); }
        for (int i = 0; i < NBUF; ++i) {
            mbar_init(BAR(ACC1_FULL + i), 1);
            mbar_init(BAR(A2_FULL + i), GROUP_WARPS);     // a chunk is converted (or drained) by ONE group
            mbar_init(BAR(A2_EMPTY + i), 1);
        }
        mbar_init(BAR(R_READY), GROUP_WARPS);             // group 0 builds the r operand
        mbar_init(BAR(ACC2_FULL), 1);
        for (int i = 0; i < ROW_GROUPS * G_STAGES; ++i) {
            mbar_init(BAR(G_FULL + i), GROUP_WARPS * 32);   // one cp.async-completion arrival per thread of the group
            mbar_init(BAR(G_EMPTY + i), GROUP_WARPS);       // one arrival per warp of the group after its reads
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == ROW_WARPS) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // epilogue staging tiles (see the epilogue): 3 over the hidden-buffer region / Wc ring, 1 over W stage 1
    constexpr int TILE_B = BM * STG_LD * 4;
    auto tile_ptr = [&](int t) -> float* {
        return reinterpret_cast<float*>(t < 3 ? smem + OFF_A2 + t * TILE_B : smem + OFF_W + w2);
    };

    // ---- geometry of a row thread.  Two GROUPS of eight warps; inside a group two warps per TMEM lane quarter: thread
    // (r, half) owns tile row r and columns [16*half, 16*half+16) of every 32-column chunk its group converts; group g
    // converts the hidden chunks c = g, g + 2, ...  Three things bound the main loop at about the same level (in-kernel
    // stamps): the L2 -> SM stream of W_a2 (57 KB per chunk with hi + lo images), the tensor pipe
    // (1.25 k cycles of MMA2 per chunk with three passes) and the per-chunk chain wait -> tmem ld -> add / ReLU / split
    // -> tmem st -> arrive of ONE group (~1.2 k).  With the two-pass logits contraction (split 2: half the stream, two
    // thirds of the MMAs) the conversion chain is what is left, so two groups take alternate chunks.  (Splitting a
    // chunk's 32 columns over four warps instead does NOT work: the hi / lo images of a K=16 step interleave in
    // 8-column blocks, so a warp's stores would overwrite accumulator columns another warp has not read yet.)
    // The two reduce-only warps of the epilogue evaluate the same expressions (unused there).
    const bool is_row = warp < ROW_WARPS;
    const int r = threadIdx.x & (BM - 1);            // tile row == TMEM lane
    const int half = (threadIdx.x >> 7) & 1;
    const int grp = (threadIdx.x >> 8) & 1;          // conversion group
    const int gwarp = warp & (GROUP_WARPS - 1);      // warp index inside the group
    const int64_t q0 = (int64_t)blockIdx.x * p.tq;
    const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    // gather staging (see G_PITCH): a warp copies 16 Ka rows and up to 2 Qa rows per chunk of its group
    const int g_piece = lane & 7, g_sub = lane >> 3;
    const int g_rbase = (gwarp & 3) * 32 + (gwarp >> 2) * 16;
    int gj[4] = {0, 0, 0, 0};
    const bool dbg = O4D_STAMPS && (blockIdx.x == gridDim.x / 2) && threadIdx.x == 0;

    if (is_row) {
        // ================================================================== row warps: prologue + main loop
        const int qi = r / k, jn = r - qi * k;
        const int64_t i = q0 + qi;
        const bool valid = (qi < p.tq) && (i < p.n);
        const int j = valid ? p.nbr[i * k + jn] : 0;
        if (half == 0 && grp == 0) s_j[r] = j;
        if (grp == 0) {
            float rv[16];
            if (valid) {
                const float rx = p.pos[i * p.ldpos + 0] - p.pos2[(int64_t)j * p.ldpos2 + 0];
                const float ry = p.pos[i * p.ldpos + 1] - p.pos2[(int64_t)j * p.ldpos2 + 1];
                const float rz = p.pos[i * p.ldpos + 2] - p.pos2[(int64_t)j * p.ldpos2 + 2];
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const int t = half * 16 + e;
                    const float h = fmaf(p.wp1[t * 3 + 2], rz, fmaf(p.wp1[t * 3 + 1], ry, fmaf(p.wp1[t * 3 + 0], rx, p.bp1[t])));
                    rv[e] = fmaxf(h, 0.f);
                }
            } else {
#pragma unroll
                for (int e = 0; e < 16; ++e) rv[e] = 0.f;
            }
            store_a_half_row(smem + OFF_R, smem + OFF_R + R_BYTES, r, half, rv);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(R_READY));
        }
        __syncwarp();
        asm volatile("bar.sync 1, %0;" ::"n"(ROW_THREADS) : "memory");      // s_j complete
#pragma unroll
        for (int it = 0; it < 4; ++it) gj[it] = s_j[g_rbase + it * 4 + g_sub];
        const int g_qi = gwarp * 2 + g_sub;                             // Qa row handled by lanes 0-15
        const bool g_has_q = lane < 16 && g_qi < p.tq;
        const int64_t g_qrow = min(q0 + (int64_t)g_qi, p.n - 1);
        // gather stages of this group: group 0 behind the weight stages, group 1 over the hidden-buffer region (idle
        // until the epilogue turns it into staging tiles)
        const uint32_t g_region = grp == 0 ? (uint32_t)OFF_G : (uint32_t)OFF_A2;
        const int gbar = grp * G_STAGES;                                // this group's G_FULL / G_EMPTY barriers
        auto issue_gather = [&](int cn) {                               // cn: a chunk of THIS group, stage (cn >> 1) & 1
            const uint32_t gst = smem_base + g_region + ((cn >> 1) & 1) * G_STAGE;
#pragma unroll
            for (int it = 0; it < 4; ++it)
                cp_async16(gst + (g_rbase + it * 4 + g_sub) * G_PITCH + g_piece * 16,
                           p.ka + (int64_t)gj[it] * 2 * d + cn * HC + g_piece * 4);
            if (g_has_q)
                cp_async16(gst + (BM + g_qi) * G_PITCH + g_piece * 16, p.qa + g_qrow * 2 * d + cn * HC + g_piece * 4);
            cp_async_mbar_arrive_noinc(BAR(G_FULL + gbar + ((cn >> 1) & 1)));
        };
        if (grp < NC) issue_gather(grp);
        const int qi_rd = min(qi, p.tq - 1);                            // padding rows read a real query's slice
        const uint32_t rd_k = g_region + r * G_PITCH + half * 64;
        const uint32_t rd_q = g_region + (BM + qi_rd) * G_PITCH + half * 64;
        if (dbg) g_dbg[0] = clock64();
        long long s_top = 0, s_acc1 = 0, s_ld = 0, s_cvt = 0, s_st = 0, ts = 0;
        for (int c = grp; c < NC; c += ROW_GROUPS) {
            if (dbg) ts = clock64();
            const int b = c % NBUF;
            const uint32_t use = (uint32_t)(c / NBUF);
            const int gi = c >> 1;                       // running index of the chunk inside this group
            const int gs_idx = gi & 1;                   // gather stage of chunk c
            if (c + ROW_GROUPS < NC) {
                // the other stage of this group was last read for the group's previous chunk (gi - 1): every warp of
                // the group must have released it
                if (gi > 0) mbar_wait(BAR(G_EMPTY + gbar + (gs_idx ^ 1)), (uint32_t)((gi - 1) >> 1) & 1u);
                issue_gather(c + ROW_GROUPS);
            }
            if (dbg) { const long long t1 = clock64(); s_top += t1 - ts; ts = t1; }
            mbar_wait(BAR(ACC1_FULL + b), use & 1u);
            tc_fence_after();
            if (dbg) { const long long t1 = clock64(); s_acc1 += t1 - ts; ts = t1; }
            uint32_t acc[16];
            tmem_ld16_nowait(taddr + ACC1_COL + b * 32 + half * 16, acc);
            mbar_wait(BAR(G_FULL + gbar + gs_idx), (uint32_t)(gi >> 1) & 1u);
            float qk[16];
            {
                const uint8_t* gs = smem + gs_idx * G_STAGE;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float4 qv = *reinterpret_cast<const float4*>(gs + rd_q + 16 * e);
                    const float4 kv = *reinterpret_cast<const float4*>(gs + rd_k + 16 * e);
                    qk[4 * e] = qv.x - kv.x; qk[4 * e + 1] = qv.y - kv.y;
                    qk[4 * e + 2] = qv.z - kv.z; qk[4 * e + 3] = qv.w - kv.w;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(G_EMPTY + gbar + gs_idx));   // this warp's reads of the stage are done
            tmem_ld_wait();
            if (dbg) { const long long t1 = clock64(); s_ld += t1 - ts; ts = t1; }
            float h[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) h[e] = fmaxf(__uint_as_float(acc[e]) + qk[e], 0.f);
            // The hidden chunk goes back into the SAME tensor-memory columns acc1[b] came from, as the A
            // operand of MMA2 (bf16 hi / lo, two values per 32-bit column): this thread read columns
            // [16*half, 16*half+16) of the buffer and overwrites them with [hi (8 columns) | lo (8 columns)]
            // of its 16 hidden units = one K=16 step.  (With the hidden chunk in shared memory the loop was
            // shared-memory-bandwidth bound: every one of the 12 MMA2 instructions re-read a 4 KB A slab plus
            // 6.6 KB of weights; A from TMEM removed 49 KB of reads and 16 KB of stores per chunk.)
            // No "buffer empty" wait: ACC1_FULL(c) means MMA1(c) has retired, and the tensor pipe runs in
            // issue order, so MMA2(c - NBUF), the last reader of these columns, retired before it.
            {
                uint32_t whi[8], wlo[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) split_bf16x2(h[2 * e], h[2 * e + 1], whi[e], wlo[e]);
                const uint32_t ta = taddr + ACC1_COL + b * 32 + half * 16;
                if (dbg) { const long long t1 = clock64(); s_cvt += t1 - ts; ts = t1; }
                tmem_st8(ta, whi);
                tmem_st8(ta + 8, wlo);
                tmem_st_wait();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(A2_FULL + b));
            if (dbg) { const long long t1 = clock64(); s_st += t1 - ts; ts = t1; }
        }
        if (dbg) { g_dbg[1] = clock64(); g_dbg[10] = s_top; g_dbg[11] = s_acc1; g_dbg[12] = s_ld; g_dbg[13] = s_cvt; g_dbg[14] = s_st; }
    } else if (warp == ROW_WARPS) {
        // ================================================================== weight stream
        if (lane == 0) {
            auto load_wc = [&](int c) {
                const int s = c % WC_STAGES;
                const uint32_t use = (uint32_t)(c / WC_STAGES);
                mbar_wait(BAR(WC_EMPTY + s), (use & 1u) ^ 1u);
                mbar_arrive_expect_tx(BAR(WC_FULL + s), 2 * WC_BYTES);
                bulk_g2s(smem_base + OFF_WC + s * 2 * WC_BYTES, p.wmain + (size_t)c * wpacked, 2 * WC_BYTES, BAR(WC_FULL + s));
            };
            for (int c = 0; c < WC_STAGES && c < NC; ++c) load_wc(c);
            for (int c = 0; c < NC; ++c) {
                if (c + WC_STAGES < NC) load_wc(c + WC_STAGES);   // the Wc ring runs a full ring ahead (MMA1 leads MMA2 by NBUF chunks)
                const int s = c & 1;
                const uint32_t use = (uint32_t)(c >> 1);
                mbar_wait(BAR(W_EMPTY + s), (use & 1u) ^ 1u);
                // [W_a2 hi][W_a2 lo]; split 2 never reads the lo image: half the bytes of the dominant stream
                const uint32_t wbytes = p.split == 2 ? (uint32_t)(w2 / 2) : (uint32_t)w2;
                mbar_arrive_expect_tx(BAR(W_FULL + s), wbytes);
                bulk_g2s(smem_base + OFF_W + s * w2, p.wmain + (size_t)c * wpacked + 2 * WC_BYTES, wbytes, BAR(W_FULL + s));
            }
            // W_p2 image into W stage 0 for the delta contraction (NC is even: this is use NC/2 of stage 0)
            const uint32_t use = (uint32_t)(NC >> 1);
            mbar_wait(BAR(W_EMPTY + 0), (use & 1u) ^ 1u);
            mbar_arrive_expect_tx(BAR(W_FULL + 0), (uint32_t)(2 * d * 64));
            bulk_g2s(smem_base + OFF_W, p.wp2, (uint32_t)(2 * d * 64), BAR(W_FULL + 0));
        }
        __syncwarp();
    } else if (warp == ROW_WARPS + 1) {
        // ================================================================== MMA issuers
        // Two issuing threads.  With a single one the loop was bound by that thread: per chunk
        // ~1080 cycles inside the MMA2 issue + commits (back-pressured by the tensor pipe), ~420 in the
        // MMA1 issue + commits, ~460 in three mbarrier waits, all serial (in-kernel stamps: 51.7 k of the
        // 48.7 k loop).  Warp 17 issues MMA2 only; warp 18 issues MMA1 (and the delta contractions of the
        // epilogue) and runs ahead.  Cross-thread ordering is explicit: MMA1(c) overwrites acc1[c % 3],
        // which MMA2(c - 3) read as its A operand, so warp 18 waits on A2_EMPTY (committed by warp 17).
        // Shared-memory descriptors are loop invariant per (buffer, k-step, hi/lo, n-tile); the
        // addresses differ only in the 14-bit start field, so every descriptor is a constant
        // base plus a small per-buffer offset.
        if (lane == 0) {
            const uint32_t idescN = umma_idesc(dn);
            const uint32_t lbo_b = (uint32_t)d * 16;
            const int split = p.split;
            const uint64_t dw0 = umma_desc(smem_base + OFF_W, lbo_b, 128);         // W stage 0, hi, n-tile 0, ks 0
            // logits += hidden[c % NBUF] . W_a2[c]^T  (K = 32, N = dn per n-tile); hidden (A) sits in tensor
            // memory: per K=16 step [hi: 8 columns | lo: 8 columns] inside the acc1 buffer's 32 columns
            auto mma2 = [&](int c) {
                const uint32_t abase = tmem_base + ACC1_COL + (c % NBUF) * 32;
                const uint64_t wbase = dw0 + (uint64_t)(((c & 1) * w2) >> 4);
                for (int t = 0; t < ntile; ++t) {
                    const uint32_t dcol = tmem_base + t * dn;
                    const uint64_t wt = wbase + (uint64_t)(((t * dn / 8) * 128) >> 4);
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) {
                        const uint32_t a_hi = abase + ks * 16;
                        const uint32_t a_lo = a_hi + 8;
                        const uint64_t w_hi = wt + (uint64_t)((ks * 2 * lbo_b) >> 4);
                        const uint64_t w_lo = w_hi + (uint64_t)((d * 64) >> 4);
                        umma_f16_ts(dcol, a_hi, w_hi, idescN, (c | ks) ? 1u : 0u);
                        if (split) {
                            if (split != 3) umma_f16_ts(dcol, a_lo, w_hi, idescN, 1u);   // split 3: hidden rounded to bf16
                            if (split != 2) umma_f16_ts(dcol, a_hi, w_lo, idescN, 1u);   // split 2: W_a2 rounded to bf16
                        }
                    }
                }
            };
            // (Tried, not kept -- 39.1 k instead of 34.2 k cycles per tile: four hi-only W_a2 stages in the two-pass mode
            // together with awaiting chunk c + 1's operands between the two n-tiles of chunk c.)
            for (int c = 0; c < NC; ++c) {
                mbar_wait(BAR(A2_FULL + c % NBUF), (uint32_t)(c / NBUF) & 1u);
                mbar_wait(BAR(W_FULL + (c & 1)), (uint32_t)(c >> 1) & 1u);
                tc_fence_after();
                mma2(c);
                umma_commit(BAR(W_EMPTY + (c & 1)));
                umma_commit(BAR(A2_EMPTY + c % NBUF));          // acc1[c % 3] may be overwritten by MMA1(c + 3)
            }
            umma_commit(BAR(ACC2_FULL));
        }
        __syncwarp();
    } else {
        if (lane == 0) {
            const uint32_t idesc32 = umma_idesc(32);
            const uint32_t lbo_a = BM * 16, lbo_b = (uint32_t)d * 16;
            const int split = p.split;
            uint64_t dr_hi[2], dr_lo[2];
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                dr_hi[ks] = umma_desc(smem_base + OFF_R + ks * 2 * lbo_a, lbo_a, 128);
                dr_lo[ks] = umma_desc(smem_base + OFF_R + R_BYTES + ks * 2 * lbo_a, lbo_a, 128);
            }
            const uint64_t dc0 = umma_desc(smem_base + OFF_WC, 512, 128);          // Wc ring slot 0, hi, ks 0
            // acc1[c % NBUF] = r . Wc[c]^T  (K = 32, N = 32)
            auto mma1 = [&](int c) {
                const uint32_t dcol = tmem_base + ACC1_COL + (c % NBUF) * 32;
                const uint64_t base = dc0 + (uint64_t)(((c % WC_STAGES) * 2 * WC_BYTES) >> 4);
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    const uint64_t b_hi = base + (uint64_t)((ks * 2 * 512) >> 4);
                    const uint64_t b_lo = b_hi + (uint64_t)(WC_BYTES >> 4);
                    umma_f16(dcol, dr_hi[ks], b_hi, idesc32, ks ? 1u : 0u);
                    if (split) {
                        umma_f16(dcol, dr_lo[ks], b_hi, idesc32, 1u);
                        umma_f16(dcol, dr_hi[ks], b_lo, idesc32, 1u);
                    }
                }
            };
            mbar_wait(BAR(R_READY), 0);
            tc_fence_after();
            // MMA1(c) overwrites acc1[c % 3], which MMA2(c - 3) read as its A operand: wait for its completion.
            // (Tried, measured, not kept -- profiles/r2_c_*: (i) ordering MMA1(c) only behind the ISSUE of MMA2(c - 3)
            // through a flag, relying on the in-order tensor pipe: correct, same speed, because the MMA2 thread queues
            // MMA2(c - 2) before this thread wakes up; (ii) ONE thread issuing MMA2(c), MMA1(c + 3) in program order:
            // correct, but the ~50 cycles of issue per tcgen05 instruction and the waits then serialise, 44.5 k instead
            // of 34 k cycles per tile.)
            for (int c = 0; c < NC; ++c) {
                mbar_wait(BAR(WC_FULL + c % WC_STAGES), (uint32_t)(c / WC_STAGES) & 1u);
                if (c >= NBUF) mbar_wait(BAR(A2_EMPTY + c % NBUF), (uint32_t)(c / NBUF - 1) & 1u);   // MMA2(c - 3) retired
                tc_fence_after();
                mma1(c);
                umma_commit(BAR(ACC1_FULL + c % NBUF));
                umma_commit(BAR(WC_EMPTY + c % WC_STAGES));
            }
            // delta chunks: acc1[g % 3] = r . W_p2[cc*32 .. +32]^T   (g = NC + cc continues the buffer ring).
            // This thread runs ahead of the MMA2 issuer, and a parity wait only distinguishes adjacent phases
            // of W_FULL[0]: first wait until every MMA2 has retired (so all W_a2 phases of the stage are
            // over), then for the W_p2 image.
            mbar_wait(BAR(ACC2_FULL), 0);
            mbar_wait(BAR(W_FULL + 0), (uint32_t)(NC >> 1) & 1u);
            tc_fence_after();
            const uint64_t dp_hi0 = umma_desc(smem_base + OFF_W, lbo_b, 128);
            const uint64_t dp_hi1 = umma_desc(smem_base + OFF_W + 2 * lbo_b, lbo_b, 128);
            const uint64_t dp_lo0 = umma_desc(smem_base + OFF_W + d * 64, lbo_b, 128);
            const uint64_t dp_lo1 = umma_desc(smem_base + OFF_W + d * 64 + 2 * lbo_b, lbo_b, 128);
            for (int cc = 0; cc < ND; ++cc) {
                const int g = NC + cc;
                const int b = g % NBUF;
                const uint32_t use = (uint32_t)(g / NBUF);
                if (g - NBUF < NC) {
                    // the buffer's previous user was a main-loop chunk: MMA2(g - 3) read it as A
                    if (g >= NBUF) mbar_wait(BAR(A2_EMPTY + b), (uint32_t)(g / NBUF - 1) & 1u);
                } else if (use > 0) {                            // a delta chunk: its acc1[b] was drained by the row warps
                    mbar_wait(BAR(A2_FULL + b), (use - 1u) & 1u);
                }
                tc_fence_after();
                const uint32_t dcol = tmem_base + ACC1_COL + b * 32;
                const uint64_t adv = (uint64_t)(cc * 512 >> 4);  // 32 rows further down the W_p2 image
                umma_f16(dcol, dr_hi[0], dp_hi0 + adv, idesc32, 0u);
                if (split) {
                    umma_f16(dcol, dr_lo[0], dp_hi0 + adv, idesc32, 1u);
                    umma_f16(dcol, dr_hi[0], dp_lo0 + adv, idesc32, 1u);
                }
                umma_f16(dcol, dr_hi[1], dp_hi1 + adv, idesc32, 1u);
                if (split) {
                    umma_f16(dcol, dr_lo[1], dp_hi1 + adv, idesc32, 1u);
                    umma_f16(dcol, dr_hi[1], dp_lo1 + adv, idesc32, 1u);
                }
                umma_commit(BAR(ACC1_FULL + b));
            }
        }
    }

    if (warp <= ROW_WARPS + 1) {
        // ================================================================== softmax / aggregation epilogue
        // 16 row warps stage, 18 warps (+ weight-stream warp, + MMA2 warp: both idle by now) reduce.  All 18 run THIS
        // code, so every warp reaches the named barrier through the same (aligned) instruction.
        // Staging tiles [row][33] of one 32-channel chunk (padding: conflict-free row-wise stores and column-wise
        // loads): logits * scale and V + delta, 16.5 KB each.  The per-channel softmax over a query's k rows is a
        // reduction ACROSS TMEM lanes, so it goes through shared memory.  Chunks are processed in groups of two
        // (4 tiles: 3 over the idle hidden-buffer region / Wc ring, 1 over the idle W stage 1; W_p2 sits in W stage 0);
        // conversion group g stages chunk u = g of a group (its 256 threads cover the 128 x 32 block), and the
        // 2 x tq (chunk, query) tasks are dealt round-robin to the 18 warps: ONE round at k = 14.
        // The V_j slices of a group are fetched one group ahead into the two group-0 gather stages by the same
        // cooperative cp.async pattern as Ka in the main loop: a row-per-thread gather from global costs one L1
        // wavefront per lane (~1000 LSU cycles per chunk), and adding V in the reduce phase instead costs ~9 address
        // instructions per load (measured: no gain).  b_p2 is added once per output instead of once per pair: a
        // channel's softmax weights sum to one.
        constexpr int GROUP = EPI_GROUP;
        static_assert(GROUP == ROW_GROUPS, "one chunk of a group per conversion group");
        auto issue_v = [&](int cc, int slot) {
            const uint32_t gst = smem_base + OFF_G + slot * G_STAGE;
#pragma unroll
            for (int it = 0; it < 4; ++it)
                cp_async16(gst + (g_rbase + it * 4 + g_sub) * G_PITCH + g_piece * 16,
                           p.vtab + (int64_t)gj[it] * d + cc * 32 + g_piece * 4);
        };
        long long t_wait = 0, t_stage = 0, t_bar = 0, t_red = 0, t0 = 0;
        if (is_row) {
            mbar_wait(BAR(ACC2_FULL), 0);
            tc_fence_after();
            if (dbg) g_dbg[2] = clock64();
            // every warp of both groups has consumed its last Ka/Qa stage (its G_EMPTY arrival precedes this barrier):
            // from here on the group-0 stages hold V slices and the hidden-buffer region (group 1's stages) holds tiles
            __syncwarp();
            asm volatile("bar.sync 1, %0;" ::"n"(ROW_THREADS) : "memory");
            if (grp < ND) issue_v(grp, grp);
            cp_async_commit();
        }
        for (int g0 = 0; g0 < ND; g0 += GROUP) {
            const int gn = min(GROUP, ND - g0);
            if (dbg) t0 = clock64();
            if (is_row) cp_async_wait_all();
            __syncwarp();
            asm volatile("bar.sync 2, %0;" ::"n"(EPI_THREADS) : "memory");      // V slices landed; previous group's tiles are free
            if (dbg) t_bar += clock64() - t0;
            if (is_row && grp < gn) {
                const int u = grp;
                const int cc = g0 + u;
                const int g = NC + cc;
                const int b = g % NBUF;
                const uint32_t use = (uint32_t)(g / NBUF);
                if (dbg) t0 = clock64();
                mbar_wait(BAR(ACC1_FULL + b), use & 1u);
                tc_fence_after();
                if (dbg) { const long long t1 = clock64(); t_wait += t1 - t0; t0 = t1; }
                uint32_t dl[16], lg[16];
                tmem_ld16_nowait(taddr + ACC1_COL + b * 32 + half * 16, dl);
                tmem_ld16_nowait(taddr + cc * 32 + half * 16, lg);
                float4 vv[4];
                {
                    const uint8_t* vs = smem + OFF_G + u * G_STAGE + r * G_PITCH + half * 64;
#pragma unroll
                    for (int e = 0; e < 4; ++e) vv[e] = *reinterpret_cast<const float4*>(vs + 16 * e);
                }
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(BAR(A2_FULL + b));      // acc1[b] drained
                float* sl = tile_ptr(2 * u) + r * STG_LD + half * 16;
                float* sv = tile_ptr(2 * u + 1) + r * STG_LD + half * 16;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    sl[4 * e + 0] = __uint_as_float(lg[4 * e + 0]) * p.scale_log2;
                    sl[4 * e + 1] = __uint_as_float(lg[4 * e + 1]) * p.scale_log2;
                    sl[4 * e + 2] = __uint_as_float(lg[4 * e + 2]) * p.scale_log2;
                    sl[4 * e + 3] = __uint_as_float(lg[4 * e + 3]) * p.scale_log2;
                    sv[4 * e + 0] = __uint_as_float(dl[4 * e + 0]) + vv[e].x;
                    sv[4 * e + 1] = __uint_as_float(dl[4 * e + 1]) + vv[e].y;
                    sv[4 * e + 2] = __uint_as_float(dl[4 * e + 2]) + vv[e].z;
                    sv[4 * e + 3] = __uint_as_float(dl[4 * e + 3]) + vv[e].w;
                }
                if (dbg) t_stage += clock64() - t0;
            }
            if (dbg) t0 = clock64();
            __syncwarp();
            asm volatile("bar.sync 2, %0;" ::"n"(EPI_THREADS) : "memory");      // the group's tiles are complete, V stages consumed
            if (dbg) { const long long t1 = clock64(); t_bar += t1 - t0; t0 = t1; }
            if (is_row) {                                                       // next group's V slices: in flight during the reduce
                if (g0 + GROUP + grp < ND) issue_v(g0 + GROUP + grp, grp);
                cp_async_commit();
            }
            if (dbg) { const long long t1 = clock64(); g_dbg[9] += t1 - t0; t0 = t1; }
            // task = (chunk u, query q)
            const int ntask = gn * p.tq;
            for (int task = warp; task < ntask; task += EPI_WARPS) {
                const int u = task >= p.tq ? 1 : 0, q = task - u * p.tq;
                const int64_t gi = q0 + q;
                if (gi < p.n)
                    reduce_task<KT>(tile_ptr(2 * u), tile_ptr(2 * u + 1), q, k, lane, p.out + gi * d + (g0 + u) * 32,
                                    p.bp2 + (g0 + u) * 32);
            }
            if (dbg) t_red += clock64() - t0;
        }
        if (dbg) { g_dbg[5] = t_wait; g_dbg[6] = t_stage; g_dbg[7] = t_bar; g_dbg[8] = t_red; }
        if (dbg) g_dbg[3] = clock64();
        if (is_row) tc_fence_before();
    }
#undef BAR
    __syncthreads();
    if (warp == ROW_WARPS) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---- weight images ---------------------------------------------------------------------------
// element (row, kk) of a K-major no-swizzle image with `rows` rows and 32 k-values:
//   [kc = kk/8][row-group = row/8][row%8][kk%8]
__device__ __forceinline__ int img_index(int rows, int row, int kk) {
    return (kk >> 3) * (rows * 8) + (row >> 3) * 64 + (row & 7) * 8 + (kk & 7);
}

__global__ void fused_pack_kernel(const float* __restrict__ wc, const float* __restrict__ wa2, const float* __restrict__ wp2,
                                  int d, __nv_bfloat16* __restrict__ out_main, __nv_bfloat16* __restrict__ out_wp2) {
    const int NC = 2 * d / HC;
    const int stage_elems = wstage_bytes(d) / 2;
    const int64_t total_main = (int64_t)NC * (HC * 32 + d * 32);
    const int64_t total = total_main + (int64_t)d * 32;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        __nv_bfloat16 hi, lo;
        if (e < total_main) {
            const int c = (int)(e / (HC * 32 + d * 32));
            const int w = (int)(e % (HC * 32 + d * 32));
            __nv_bfloat16* stage = out_main + (int64_t)c * stage_elems;
            if (w < HC * 32) {            // Wc chunk: rows = hidden units c*32.., k = r index
                const int row = w / 32, kk = w % 32;
                split_bf16(wc[(int64_t)(c * HC + row) * 32 + kk], hi, lo);
                const int idx = img_index(HC, row, kk);
                stage[idx] = hi;
                stage[HC * 32 + idx] = lo;
            } else {                      // W_a2 chunk: rows = output channels, k = hidden units c*32..
                const int w2 = w - HC * 32;
                const int row = w2 / 32, kk = w2 % 32;
                split_bf16(wa2[(int64_t)row * (2 * d) + c * HC + kk], hi, lo);
                const int idx = img_index(d, row, kk);
                stage[2 * HC * 32 + idx] = hi;
                stage[2 * HC * 32 + d * 32 + idx] = lo;
            }
        } else {
            const int w = (int)(e - total_main);
            const int row = w / 32, kk = w % 32;
            split_bf16(wp2[(int64_t)row * 32 + kk], hi, lo);
            const int idx = img_index(d, row, kk);
            out_wp2[idx] = hi;
            out_wp2[d * 32 + idx] = lo;
        }
    }
}

}  // namespace fa

// Which of the three bf16x3 products the LOGITS contraction (hidden -> logits, 81 % of the decoder's flops) issues:
// 1 = all three, 2 = hidden_hi W_hi + hidden_lo W_hi (W_a2 rounded to bf16, the hidden layer keeps its 16-bit split),
// 3 = drop hidden_lo W_hi instead (diagnostics only: 5e-2 .. 2e-1 on the released checkpoints).
// Default 2.  Measured on the released checkpoints against the fp64 arbitration run (profiles/r2_a_ckpt_error.json,
// r2_c_ckpt_error.json): GREATER 6.1e-4 (mode 1) vs 6.3e-4 (mode 2), CARLA 9.7e-4 vs 8.8e-4 -- indistinguishable,
// while the reference's own fp32 forward sits at 2.3e-4 / 1.6e-4.  Why the weight rounding is harmless here and the
// hidden rounding is not: the rounding error of W_a2 is the SAME perturbation for all k neighbours of a query, and the
// per-channel softmax over neighbours only sees differences between their logits; rounding the hidden activations
// perturbs every (query, neighbour) pair independently.  What it buys: one third of the kernel's MMA work and half of
// its weight stream (the main loop is bound by the L2 -> SM stream of W_a2: 57 KB per 32-unit chunk per SM, in-kernel
// stamps in DESIGN.md section 4).  O4D_FUSED_LOGIT_PASSES=3 (or o4d_debug_set_fused_passes(1)) restores all three.
static int fused_mode_default() {
    const char* e = getenv("O4D_FUSED_LOGIT_PASSES");
    return (e && e[0] == '3') ? 1 : 2;
}
static std::atomic<int> g_fused_mma2_mode{0};   // 0 = not set: environment / default

bool attn_fused_supported(int d, int k) {
    if (d % 32 != 0 || d < 288 || d > 416 || k < 8 || k > O4D_MAX_K) return false;  // k >= 8: at most 16 queries per tile
    const int dn = d > 256 ? d / 2 : d;
    return dn % 16 == 0 && dn <= 256 && fa::smem_bytes(d) <= 227 * 1024;
}

size_t attn_fused_pack_bytes(int d) {
    const int NC = 2 * d / fa::HC;
    return align_up((size_t)NC * fa::wstage_bytes(d), 256) + align_up((size_t)2 * d * 64, 256);
}

int attn_fused_pack_launch(const float* wc, const float* wa2, const float* wp2, int d, void* packed, cudaStream_t st) {
    const int NC = 2 * d / fa::HC;
    __nv_bfloat16* main_img = (__nv_bfloat16*)packed;
    __nv_bfloat16* wp2_img = (__nv_bfloat16*)((char*)packed + align_up((size_t)NC * fa::wstage_bytes(d), 256));
    fa::fused_pack_kernel<<<148 * 4, 256, 0, st>>>(wc, wa2, wp2, d, main_img, wp2_img);
    O4D_LAUNCH_CHECK();
    return 0;
}

int attn_fused_launch(const PtBlockParams& P, const AttnTables& T, const float* qa, const float* pos, int64_t ldpos,
                      const float* pos2, int64_t ldpos2, const int32_t* nbr, int64_t n, int d, int k, float* out,
                      int precision, cudaStream_t st) {
    if (n == 0) return 0;
    O4D_REQUIRE(attn_fused_supported(d, k) && T.fused, "fused attention: unsupported shape or missing weights");
    O4D_SMEM_ATTR(fa::attn_fused_kernel<0>, 227 * 1024);
    O4D_SMEM_ATTR(fa::attn_fused_kernel<12>, 227 * 1024);
    O4D_SMEM_ATTR(fa::attn_fused_kernel<14>, 227 * 1024);
    O4D_SMEM_ATTR(fa::attn_fused_kernel<16>, 227 * 1024);
    const int NC = 2 * d / fa::HC;
    fa::Params p;
    p.pos = pos; p.ldpos = ldpos; p.pos2 = pos2; p.ldpos2 = ldpos2; p.nbr = nbr;
    p.qa = qa; p.ka = T.ka; p.vtab = T.vtab;
    p.wp1 = P.wp1; p.bp1 = P.bp1; p.bp2 = P.bp2;
    p.wmain = (const uint8_t*)T.fused;
    p.wp2 = (const uint8_t*)T.fused + align_up((size_t)NC * fa::wstage_bytes(d), 256);
    p.out = out;
    p.n = n; p.d = d; p.k = k; p.tq = fa::BM / k; {
        int mode = g_fused_mma2_mode.load(std::memory_order_relaxed);
        if (mode == 0) mode = fused_mode_default();
        p.split = (precision == 1) ? mode : 0;
    }
    p.scale_log2 = (float)(1.4426950408889634 / sqrt((double)d));
    const int64_t tiles = cdiv(n, p.tq);
    // algorithmic flops of what this launch replaces (reference formulation): per pair 2*(3*32 + 32*d) +
    // 2*d*2d*2, softmax/aggregate ~6d
    ProfScope prof(PROF_FUSED, (double)n * k * (2.0 * (3 * 32 + 32.0 * d) + 8.0 * d * d + 6.0 * d), st);
    const unsigned grid = (unsigned)tiles;
    const size_t smem = fa::smem_bytes(d);
    switch (k) {
        case 12: fa::attn_fused_kernel<12><<<grid, fa::THREADS, smem, st>>>(p); break;
        case 14: fa::attn_fused_kernel<14><<<grid, fa::THREADS, smem, st>>>(p); break;
        case 16: fa::attn_fused_kernel<16><<<grid, fa::THREADS, smem, st>>>(p); break;
        default: fa::attn_fused_kernel<0><<<grid, fa::THREADS, smem, st>>>(p); break;
    }
    O4D_LAUNCH_CHECK();
    return 0;
}

}  // namespace o4d

extern "C" void o4d_debug_set_fused_passes(int mode) { o4d::g_fused_mma2_mode.store(mode >= 1 && mode <= 3 ? mode : 0); }
extern "C" int o4d_debug_read(long long* out16) {
    return (int)cudaMemcpyFromSymbol(out16, o4d::fa::g_dbg, sizeof(long long) * 16);
}
