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
// acc1 = r Wc[c]^T (128 x 32, TMEM), the four row warps pull it into registers, add the
// gathered Qa_i - Ka_j slice, apply ReLU, split to bf16 hi/lo and store it as the next
// A operand; MMA2 accumulates it against W_a2[:, c] into the (128 x d) logits accumulator.
// acc1 and the A2 buffer are double buffered so chunk c+1's first MMA and epilogue overlap
// chunk c's second MMA.  Weights arrive as pre-packed shared-memory images through
// cp.async.bulk (two 56 KB stages).  All contractions use the bf16x3 split (hi*hi + lo*hi +
// hi*lo, fp32 accumulate) unless `split` is 0.
//
// TMEM: columns [0, d) logits, [448, 512) the two acc1 buffers.  One CTA per SM.
#include "o4d_common.cuh"
#include <cuda_bf16.h>

namespace o4d {
namespace fa {

constexpr int BM = 128;
constexpr int HC = 32;          // hidden units per chunk == K of the second contraction per chunk
constexpr int THREADS = 192;
constexpr int R_BYTES = BM * 32 * 2;         // one bf16 image of a (128 x 32) A operand
constexpr int OFF_R = 0;                     // r hi, r lo
constexpr int OFF_A2 = 2 * R_BYTES;          // 2 buffers x (hi, lo)
constexpr int OFF_W = OFF_A2 + 4 * R_BYTES;  // weight stages
constexpr int ACC1_COL = 448;
constexpr int WC_BYTES = HC * 32 * 2;        // one bf16 image of a Wc chunk (32 x 32)
constexpr int STG_LD = 33;                   // padded row stride of the epilogue staging tiles

__host__ __device__ inline int wstage_bytes(int d) { return 2 * WC_BYTES + 2 * d * 32 * 2; }
__host__ __device__ inline int smem_bytes(int d) { return OFF_W + 2 * wstage_bytes(d) + 256; }

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
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(x);
    lo = __float2bfloat16_rn(x - __bfloat162float(hi));
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
    float inv_sqrt_d;
};

__global__ void __launch_bounds__(THREADS, 1) attn_fused_kernel(const Params p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int d = p.d, k = p.k;
    const int wstage = wstage_bytes(d);
    const int NC = 2 * d / HC;              // hidden chunks
    const int ND = d / 32;                  // output-column chunks of the softmax epilogue
    const int ntile = d > 256 ? 2 : 1;      // MMA2 is issued per n-tile of dn <= 256 columns
    const int dn = d / ntile;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_W + 2 * wstage);
    // barrier indices
    enum { W_FULL = 0, W_EMPTY = 2, ACC1_FULL = 4, A2_FULL = 6, A2_EMPTY = 8, R_READY = 10, ACC2_FULL = 11, NBARS = 12 };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
    const uint32_t bar0 = smem_u32(bars);
#define BAR(i) (bar0 + 8u * (uint32_t)(i))
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem_base = smem_u32(smem);

    if (threadIdx.x == 0) {
        mbar_init(BAR(W_FULL + 0), 1); mbar_init(BAR(W_FULL + 1), 1);
        mbar_init(BAR(W_EMPTY + 0), 1); mbar_init(BAR(W_EMPTY + 1), 1);
        mbar_init(BAR(ACC1_FULL + 0), 1); mbar_init(BAR(ACC1_FULL + 1), 1);
        mbar_init(BAR(A2_FULL + 0), 4); mbar_init(BAR(A2_FULL + 1), 4);
        mbar_init(BAR(A2_EMPTY + 0), 1); mbar_init(BAR(A2_EMPTY + 1), 1);
        mbar_init(BAR(R_READY), 4);
        mbar_init(BAR(ACC2_FULL), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
        // ================================================================== row warps
        const int r = threadIdx.x;                       // tile row == TMEM lane
        const int qi = r / k, jn = r - qi * k;
        const int64_t i = (int64_t)blockIdx.x * p.tq + qi;
        const bool valid = (qi < p.tq) && (i < p.n);
        const int j = valid ? p.nbr[i * k + jn] : 0;
        {
            float rv[32];
            if (valid) {
                const float rx = p.pos[i * p.ldpos + 0] - p.pos2[(int64_t)j * p.ldpos2 + 0];
                const float ry = p.pos[i * p.ldpos + 1] - p.pos2[(int64_t)j * p.ldpos2 + 1];
                const float rz = p.pos[i * p.ldpos + 2] - p.pos2[(int64_t)j * p.ldpos2 + 2];
#pragma unroll
                for (int t = 0; t < 32; ++t) {
                    const float h = fmaf(p.wp1[t * 3 + 2], rz, fmaf(p.wp1[t * 3 + 1], ry, fmaf(p.wp1[t * 3 + 0], rx, p.bp1[t])));
                    rv[t] = fmaxf(h, 0.f);
                }
            } else {
#pragma unroll
                for (int t = 0; t < 32; ++t) rv[t] = 0.f;
            }
            store_a_row(smem + OFF_R, smem + OFF_R + R_BYTES, r, rv);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(R_READY));
        }
        const float* qrow = p.qa + (valid ? i : 0) * 2 * d;
        const float* krow = p.ka + (int64_t)j * 2 * d;
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        float qk[32];
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
            const float4 a = *reinterpret_cast<const float4*>(qrow + e);
            const float4 b = *reinterpret_cast<const float4*>(krow + e);
            qk[e] = a.x - b.x; qk[e + 1] = a.y - b.y; qk[e + 2] = a.z - b.z; qk[e + 3] = a.w - b.w;
        }
        for (int c = 0; c < NC; ++c) {
            const int b = c & 1;
            const uint32_t use = (uint32_t)(c >> 1);
            mbar_wait(BAR(ACC1_FULL + b), use & 1u);
            tc_fence_after();
            uint32_t acc[32];
            tmem_ld16_nowait(tmem_base + lane_base + ACC1_COL + b * 32, acc);
            tmem_ld16_nowait(tmem_base + lane_base + ACC1_COL + b * 32 + 16, acc + 16);
            tmem_ld_wait();
            float h[32];
#pragma unroll
            for (int e = 0; e < 32; ++e) h[e] = valid ? fmaxf(__uint_as_float(acc[e]) + qk[e], 0.f) : 0.f;
            if (c + 1 < NC) {   // gather the next chunk's Qa_i - Ka_j slice while this one is stored
#pragma unroll
                for (int e = 0; e < 32; e += 4) {
                    const float4 a = *reinterpret_cast<const float4*>(qrow + (c + 1) * HC + e);
                    const float4 bb = *reinterpret_cast<const float4*>(krow + (c + 1) * HC + e);
                    qk[e] = a.x - bb.x; qk[e + 1] = a.y - bb.y; qk[e + 2] = a.z - bb.z; qk[e + 3] = a.w - bb.w;
                }
            }
            mbar_wait(BAR(A2_EMPTY + b), (use & 1u) ^ 1u);     // MMA2 of chunk c-2 has released the buffer
            uint8_t* a2 = smem + OFF_A2 + b * 2 * R_BYTES;
            store_a_row(a2, a2 + R_BYTES, r, h);
            fence_proxy_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(A2_FULL + b));
        }
        // ------------------------------------------------ softmax / aggregation epilogue
        mbar_wait(BAR(ACC2_FULL), 0);
        tc_fence_after();
        float* stg_l = reinterpret_cast<float*>(smem + OFF_W + wstage);          // stage 1 is free now
        float* stg_v = stg_l + BM * STG_LD;
        const float* vrow = p.vtab + (int64_t)j * d;
        const int items = p.tq * 32;
        for (int cc = 0; cc < ND; ++cc) {
            const int g = NC + cc;
            const int b = g & 1;
            const uint32_t use = (uint32_t)(g >> 1);
            mbar_wait(BAR(ACC1_FULL + b), use & 1u);
            tc_fence_after();
            uint32_t dl[32], lg[32];
            tmem_ld16_nowait(tmem_base + lane_base + ACC1_COL + b * 32, dl);
            tmem_ld16_nowait(tmem_base + lane_base + ACC1_COL + b * 32 + 16, dl + 16);
            tmem_ld16_nowait(tmem_base + lane_base + cc * 32, lg);
            tmem_ld16_nowait(tmem_base + lane_base + cc * 32 + 16, lg + 16);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(A2_FULL + b));      // acc1[b] drained
#pragma unroll
            for (int e = 0; e < 32; ++e) {
                const int ch = cc * 32 + e;
                const float val = valid ? (__uint_as_float(dl[e]) + p.bp2[ch] + vrow[ch]) : 0.f;
                stg_l[r * STG_LD + e] = __uint_as_float(lg[e]);
                stg_v[r * STG_LD + e] = val;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            for (int it = r; it < items; it += BM) {
                const int q = it >> 5, ch = it & 31;
                const int64_t gi = (int64_t)blockIdx.x * p.tq + q;
                if (gi < p.n) {
                    float mx = -3.4e38f;
                    for (int jj = 0; jj < k; ++jj) mx = fmaxf(mx, stg_l[(q * k + jj) * STG_LD + ch] * p.inv_sqrt_d);
                    float den = 0.f, num = 0.f;
                    for (int jj = 0; jj < k; ++jj) {
                        const float w = expf(stg_l[(q * k + jj) * STG_LD + ch] * p.inv_sqrt_d - mx);
                        den += w;
                        num = fmaf(w, stg_v[(q * k + jj) * STG_LD + ch], num);
                    }
                    p.out[gi * d + cc * 32 + ch] = num / den;
                }
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
        }
        tc_fence_before();
    } else if (warp == 4) {
        // ================================================================== weight stream
        if (lane == 0) {
            for (int c = 0; c < NC; ++c) {
                const int s = c & 1;
                const uint32_t use = (uint32_t)(c >> 1);
                mbar_wait(BAR(W_EMPTY + s), (use & 1u) ^ 1u);
                mbar_arrive_expect_tx(BAR(W_FULL + s), (uint32_t)wstage);
                bulk_g2s(smem_base + OFF_W + s * wstage, p.wmain + (size_t)c * wstage, (uint32_t)wstage, BAR(W_FULL + s));
            }
            // W_p2 image into stage 0 for the delta contraction (NC is even, so this is use NC/2 of stage 0)
            const uint32_t use = (uint32_t)(NC >> 1);
            mbar_wait(BAR(W_EMPTY + 0), (use & 1u) ^ 1u);
            mbar_arrive_expect_tx(BAR(W_FULL + 0), (uint32_t)(2 * d * 64));
            bulk_g2s(smem_base + OFF_W, p.wp2, (uint32_t)(2 * d * 64), BAR(W_FULL + 0));
        }
    } else {
        // ================================================================== MMA issuer
        if (lane == 0) {
            const uint32_t idesc32 = umma_idesc(32), idescN = umma_idesc(dn);
            const uint32_t r_hi = smem_base + OFF_R, r_lo = r_hi + R_BYTES;
            const uint32_t lbo_a = BM * 16;
            const int split = p.split;
            // acc1[buf] = r . B^T with B a (32 x 32) K-major image pair at b_hi / b_lo (row pitch lbo_b)
            auto mma_k32_n32 = [&](uint32_t b_hi, uint32_t b_lo, uint32_t lbo_b, int buf) {
                const uint32_t dcol = tmem_base + ACC1_COL + buf * 32;
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    const uint64_t da_hi = umma_desc(r_hi + ks * 2 * lbo_a, lbo_a, 128);
                    const uint64_t db_hi = umma_desc(b_hi + ks * 2 * lbo_b, lbo_b, 128);
                    umma_f16(dcol, da_hi, db_hi, idesc32, ks ? 1u : 0u);
                    if (split) {
                        const uint64_t da_lo = umma_desc(r_lo + ks * 2 * lbo_a, lbo_a, 128);
                        const uint64_t db_lo = umma_desc(b_lo + ks * 2 * lbo_b, lbo_b, 128);
                        umma_f16(dcol, da_lo, db_hi, idesc32, 1u);
                        umma_f16(dcol, da_hi, db_lo, idesc32, 1u);
                    }
                }
            };
            mbar_wait(BAR(R_READY), 0);
            tc_fence_after();
            mbar_wait(BAR(W_FULL + 0), 0);
            tc_fence_after();
            mma_k32_n32(smem_base + OFF_W, smem_base + OFF_W + WC_BYTES, 32 * 16, 0);
            umma_commit(BAR(ACC1_FULL + 0));
            for (int c = 0; c < NC; ++c) {
                const int s = c & 1;
                const uint32_t use = (uint32_t)(c >> 1);
                if (c + 1 < NC) {
                    const int s1 = (c + 1) & 1;
                    mbar_wait(BAR(W_FULL + s1), (uint32_t)((c + 1) >> 1) & 1u);
                    tc_fence_after();
                    const uint32_t wb = smem_base + OFF_W + s1 * wstage;
                    mma_k32_n32(wb, wb + WC_BYTES, 32 * 16, s1);
                    umma_commit(BAR(ACC1_FULL + s1));
                }
                mbar_wait(BAR(A2_FULL + s), use & 1u);
                tc_fence_after();
                const uint32_t a_hi = smem_base + OFF_A2 + s * 2 * R_BYTES, a_lo = a_hi + R_BYTES;
                const uint32_t w_hi = smem_base + OFF_W + s * wstage + 2 * WC_BYTES, w_lo = w_hi + d * 64;
                const uint32_t lbo_b = (uint32_t)d * 16;
                for (int t = 0; t < ntile; ++t) {
                    const uint32_t dcol = tmem_base + t * dn;
                    const uint32_t boff = (uint32_t)(t * dn / 8) * 128;
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) {
                        const uint64_t da_hi = umma_desc(a_hi + ks * 2 * lbo_a, lbo_a, 128);
                        const uint64_t db_hi = umma_desc(w_hi + boff + ks * 2 * lbo_b, lbo_b, 128);
                        umma_f16(dcol, da_hi, db_hi, idescN, (c | ks) ? 1u : 0u);
                        if (split) {
                            const uint64_t da_lo = umma_desc(a_lo + ks * 2 * lbo_a, lbo_a, 128);
                            const uint64_t db_lo = umma_desc(w_lo + boff + ks * 2 * lbo_b, lbo_b, 128);
                            umma_f16(dcol, da_lo, db_hi, idescN, 1u);
                            umma_f16(dcol, da_hi, db_lo, idescN, 1u);
                        }
                    }
                }
                umma_commit(BAR(W_EMPTY + s));
                umma_commit(BAR(A2_EMPTY + s));
            }
            umma_commit(BAR(ACC2_FULL));
            // delta chunks: acc1[g & 1] = r . W_p2[cc*32 .. +32]^T
            mbar_wait(BAR(W_FULL + 0), (uint32_t)(NC >> 1) & 1u);
            tc_fence_after();
            const uint32_t p_hi = smem_base + OFF_W, p_lo = p_hi + d * 64;
            for (int cc = 0; cc < ND; ++cc) {
                const int g = NC + cc;
                const int b = g & 1;
                const uint32_t use = (uint32_t)(g >> 1);
                mbar_wait(BAR(A2_FULL + b), (use - 1u) & 1u);   // previous contents of acc1[b] were drained
                tc_fence_after();
                mma_k32_n32(p_hi + cc * 512, p_lo + cc * 512, (uint32_t)d * 16, b);
                umma_commit(BAR(ACC1_FULL + b));
            }
        }
    }
#undef BAR
    __syncthreads();
    if (warp == 4) {
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

bool attn_fused_supported(int d, int k) {
    if (d % 32 != 0 || d < 256 || d > 416 || k < 1 || k > O4D_MAX_K) return false;
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
    static bool attr_done = false;
    if (!attr_done) {
        O4D_CUDA(cudaFuncSetAttribute(fa::attn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_done = true;
    }
    const int NC = 2 * d / fa::HC;
    fa::Params p;
    p.pos = pos; p.ldpos = ldpos; p.pos2 = pos2; p.ldpos2 = ldpos2; p.nbr = nbr;
    p.qa = qa; p.ka = T.ka; p.vtab = T.vtab;
    p.wp1 = P.wp1; p.bp1 = P.bp1; p.bp2 = P.bp2;
    p.wmain = (const uint8_t*)T.fused;
    p.wp2 = (const uint8_t*)T.fused + align_up((size_t)NC * fa::wstage_bytes(d), 256);
    p.out = out;
    p.n = n; p.d = d; p.k = k; p.tq = fa::BM / k; p.split = (precision == 1) ? 1 : 0;
    p.inv_sqrt_d = (float)(1.0 / sqrt((double)d));
    const int64_t tiles = cdiv(n, p.tq);
    // algorithmic flops of what this launch replaces (reference formulation): per pair 2*(3*32 + 32*d) +
    // 2*d*2d*2, softmax/aggregate ~6d
    ProfScope prof(PROF_FUSED, (double)n * k * (2.0 * (3 * 32 + 32.0 * d) + 8.0 * d * d + 6.0 * d), st);
    fa::attn_fused_kernel<<<(unsigned)tiles, fa::THREADS, fa::smem_bytes(d), st>>>(p);
    O4D_LAUNCH_CHECK();
    return 0;
}

}  // namespace o4d
