// tcgen05 weight gradient:  part[s][p][q] = sum over the rows r of split s of dY[r, p] * pre(A[r, q])
//
// (dW = dY^T pre(A) of a dense layer Y = pre(A) W^T + b; in the reference this is the second
// SGEMM of autograd's AddmmBackward.)  The contraction runs over ROWS, which is the slow
// dimension of both row-major operands, so neither can be bulk-copied into a K-major UMMA
// operand: eight producer warps read the fp32 tiles along their contiguous dimension
// (warp-coalesced 128-byte rows), split every value into bf16 hi + lo (same bf16x3 scheme as
// gemm_tc.cu: hi*hi + lo*hi + hi*lo in one fp32 TMEM accumulator) and store eight consecutive
// rows of one column as one 16-byte core-matrix line -- the transpose costs nothing extra.
//
//   CTA tile   : 128 (p, TMEM lanes) x BQ <= 256 (q, TMEM columns), rows walked 32 at a time
//   warps 0-7  : producers (both operands), then the epilogue (TMEM -> fp32 partial tile)
//   warp 8     : TMEM allocation;   warp 9 : lane 0 issues tcgen05.mma / commits
//   grid       : (p-tiles * q-tiles, row splits); tiles of one split are adjacent in launch order
//                so the operand rows they share are read from HBM once and hit in L2 afterwards.
// Two CTAs per SM (2 x 96 KB smem, 2 x 256 TMEM columns).  Partials are reduced in a fixed order
// by reduce_partials_kernel (train.cu): deterministic.
#include "o4d_common.cuh"
#include <cuda_bf16.h>
#include "tc_helpers.cuh"

namespace o4d {
namespace wg {

constexpr int BM = 128;
constexpr int BK = 32;
constexpr int STAGES = 2;
constexpr int PROD_WARPS = 8;
constexpr int THREADS = (PROD_WARPS + 2) * 32;
constexpr int BQ_MAX = 256;
constexpr int A_HALF = BM * BK * 2;                       // one bf16 image of the dY^T slab (8 KB)
constexpr int STAGE_BYTES = 2 * A_HALF + 2 * BQ_MAX * BK * 2;   // 48 KB
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t a, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t a) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory");
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
// Eight values of one operand column -> one 16-byte core-matrix line per image half.  Packed conversions
// (tch::split_bf16x2: F2FP.BF16.F32.PACK_AB, full rate) -- the scalar F2F form ran on the quarter-rate conversion
// pipe and made the producers, not the tensor pipe, the bound of this kernel (96 F2F per thread and 32-row chunk).
__device__ __forceinline__ void split8(const float* x, uint4& hi, uint4& lo) {
    tch::split_bf16x2(x[0], x[1], hi.x, lo.x);
    tch::split_bf16x2(x[2], x[3], hi.y, lo.y);
    tch::split_bf16x2(x[4], x[5], hi.z, lo.z);
    tch::split_bf16x2(x[6], x[7], hi.w, lo.w);
}

struct Tiling {
    int n, k;       // dW shape (p extent, q extent)
    int ptiles;     // ceil(n / 128)
    int bq, qtiles; // q tile width (multiple of 16, <= 256) and count
};

__host__ __device__ inline Tiling make_tiling(int n, int k) {
    Tiling t;
    t.n = n;
    t.k = k;
    t.ptiles = (n + BM - 1) / BM;
    int qt = (k + BQ_MAX - 1) / BQ_MAX;
    int bq = (k + qt - 1) / qt;
    bq = (bq + 15) / 16 * 16;
    t.bq = bq;
    t.qtiles = (k + bq - 1) / bq;
    return t;
}

// One 8-row x 1-column strip of a row-major fp32 matrix -> registers (zero outside the matrix).
template <bool RELU>
__device__ __forceinline__ void load_strip(const float* __restrict__ X, int64_t ld, int64_t r0, int64_t r_hi, int col,
                                           int ncols, float (&v)[8]) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int64_t r = r0 + e;
        float x = 0.f;
        if (col < ncols && r < r_hi) {
            x = X[r * ld + col];
            if (RELU) x = fmaxf(x, 0.f);
        }
        v[e] = x;
    }
}

template <bool RELU>
__global__ void __launch_bounds__(THREADS, 2)
wgrad_tc_kernel(const float* __restrict__ dY, int64_t lddy, const float* __restrict__ A, int64_t lda, int64_t rows,
                Tiling tl, int64_t rows_per_split, int split3, float* __restrict__ part) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);   // full[2], empty[2], accum
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x;
    const int pt = tile / tl.qtiles, qt = tile % tl.qtiles;
    const int p0 = pt * BM, q0 = qt * tl.bq;
    const int bq = tl.bq;
    const int64_t r_lo = (int64_t)blockIdx.y * rows_per_split;
    const int64_t r_hi = min(rows, r_lo + rows_per_split);
    const int nchunks = r_hi > r_lo ? (int)((r_hi - r_lo + BK - 1) / BK) : 0;
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[STAGES]), accum_bar = smem_u32(&bars[2 * STAGES]);
    const uint32_t smem_base = smem_u32(smem);

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full0 + 8 * s, PROD_WARPS);
            mbar_init(empty0 + 8 * s, 1);
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
    const uint32_t b_half = (uint32_t)bq * BK * 2;

    if (warp < PROD_WARPS) {
        // ------------------------------------------------------------ producers
        // thread -> column (t & 127) of each operand and the k-core pair (t >> 7): rows
        // [16 * half, 16 * half + 16) of the 32-row chunk, i.e. kc = 2 * half, 2 * half + 1.
        const int t = threadIdx.x;
        const int col = t & (BM - 1);
        const int half = t >> 7;
        for (int c = 0; c < nchunks; ++c) {
            const int s = c % STAGES;
            const uint32_t ph = (uint32_t)(c / STAGES) & 1u;
            const int64_t r0 = r_lo + (int64_t)c * BK + half * 16;
            float ya[2][8], a0[2][8], a1[2][8];
            const bool second = col + BM < bq;
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
                load_strip<false>(dY, lddy, r0 + kk * 8, r_hi, p0 + col, tl.n, ya[kk]);
                load_strip<RELU>(A, lda, r0 + kk * 8, r_hi, (col < bq) ? q0 + col : tl.k, tl.k, a0[kk]);
                if (second) load_strip<RELU>(A, lda, r0 + kk * 8, r_hi, q0 + col + BM, tl.k, a1[kk]);
            }
            mbar_wait(empty0 + 8 * s, ph ^ 1u);
            uint8_t* a_hi = smem + s * STAGE_BYTES;
            uint8_t* a_lo = a_hi + A_HALF;
            uint8_t* b_hi = a_hi + 2 * A_HALF;
            uint8_t* b_lo = b_hi + b_half;
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
                const int kc = half * 2 + kk;
                uint4 hi, lo;
                split8(ya[kk], hi, lo);
                const int offa = kc * (BM * 16) + (col >> 3) * 128 + (col & 7) * 16;
                *reinterpret_cast<uint4*>(a_hi + offa) = hi;
                *reinterpret_cast<uint4*>(a_lo + offa) = lo;
                if (col < bq) {
                    split8(a0[kk], hi, lo);
                    const int offb = kc * (bq * 16) + (col >> 3) * 128 + (col & 7) * 16;
                    *reinterpret_cast<uint4*>(b_hi + offb) = hi;
                    *reinterpret_cast<uint4*>(b_lo + offb) = lo;
                }
                if (second) {
                    split8(a1[kk], hi, lo);
                    const int c2 = col + BM;
                    const int offb = kc * (bq * 16) + (c2 >> 3) * 128 + (c2 & 7) * 16;
                    *reinterpret_cast<uint4*>(b_hi + offb) = hi;
                    *reinterpret_cast<uint4*>(b_lo + offb) = lo;
                }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(full0 + 8 * s);
        }
        // ------------------------------------------------------------ epilogue
        // warp w: TMEM lane quarter (w & 3), column half (w >> 2); one output row per thread
        float* dst = part + (int64_t)blockIdx.y * tl.n * tl.k;
        const int m = (warp & 3) * 32 + lane;
        const int gp = p0 + m;
        const int chalf = (bq / 2 + 15) / 16 * 16;            // first half rounded to the 16-column load width
        const int c_lo = (warp >> 2) ? chalf : 0;
        const int c_hi = (warp >> 2) ? bq : min(chalf, bq);
        if (nchunks > 0) {
            mbar_wait(accum_bar, 0);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
            for (int c0 = c_lo; c0 < c_hi; c0 += 16) {
                float v[16];
                tmem_ld16(taddr + (uint32_t)c0, v);
                if (gp < tl.n) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int gq = q0 + c0 + i;
                        if (c0 + i < bq && gq < tl.k) dst[(int64_t)gp * tl.k + gq] = v[i];
                    }
                }
            }
            tc_fence_before();
        } else if (gp < tl.n) {
            for (int c0 = c_lo; c0 < c_hi; ++c0)
                if (q0 + c0 < tl.k) dst[(int64_t)gp * tl.k + q0 + c0] = 0.f;
        }
    } else if (warp == PROD_WARPS + 1) {
        // ------------------------------------------------------------ MMA issuer
        if (lane == 0 && nchunks > 0) {
            const uint32_t idesc = umma_idesc(bq);
            const uint32_t lbo_a = BM * 16, lbo_b = (uint32_t)bq * 16;
            for (int c = 0; c < nchunks; ++c) {
                const int s = c % STAGES;
                const uint32_t ph = (uint32_t)(c / STAGES) & 1u;
                mbar_wait(full0 + 8 * s, ph);
                tc_fence_after();
                const uint32_t a_hi = smem_base + s * STAGE_BYTES;
                const uint32_t a_lo = a_hi + A_HALF;
                const uint32_t b_hi = a_hi + 2 * A_HALF;
                const uint32_t b_lo = b_hi + b_half;
#pragma unroll
                for (int ks = 0; ks < BK / 16; ++ks) {
                    const uint64_t da_hi = umma_desc(a_hi + ks * 2 * lbo_a, lbo_a, 128);
                    const uint64_t db_hi = umma_desc(b_hi + ks * 2 * lbo_b, lbo_b, 128);
                    umma_f16(tmem_base, da_hi, db_hi, idesc, (c | ks) ? 1u : 0u);
                    if (split3) {
                        const uint64_t da_lo = umma_desc(a_lo + ks * 2 * lbo_a, lbo_a, 128);
                        const uint64_t db_lo = umma_desc(b_lo + ks * 2 * lbo_b, lbo_b, 128);
                        umma_f16(tmem_base, da_lo, db_hi, idesc, 1u);
                        umma_f16(tmem_base, da_hi, db_lo, idesc, 1u);
                    }
                }
                umma_commit(empty0 + 8 * s);
            }
            umma_commit(accum_bar);
        }
    }
    __syncthreads();
    if (warp == PROD_WARPS) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
    }
}

}  // namespace wg

bool wgrad_tc_ok(int64_t rows, int64_t n, int64_t k) { return rows >= 4096 && n >= 32 && k >= 32 && n <= 8192 && k <= 8192; }

int wgrad_tc_splits(int64_t rows, int64_t n, int64_t k) {
    wg::Tiling t = wg::make_tiling((int)n, (int)k);
    const int64_t tiles = (int64_t)t.ptiles * t.qtiles;
    int64_t s = cdiv(148 * 2 * 2, tiles);            // about two waves of 2 CTAs per SM
    const int64_t max_by_rows = cdiv(rows, 1024);    // at least 32 chunks per CTA
    if (s > max_by_rows) s = max_by_rows;
    if (s > 256) s = 256;
    if (s < 1) s = 1;
    return (int)s;
}

// part must hold splits * n * k floats (wgrad_tc_splits).
int wgrad_tc_launch(const float* dY, int64_t lddy, const float* A, int64_t lda, int64_t rows, int64_t n, int64_t k,
                    bool relu_a, int precision, float* part, int splits, cudaStream_t st) {
    O4D_SMEM_ATTR(wg::wgrad_tc_kernel<true>, wg::SMEM_BYTES);
    O4D_SMEM_ATTR(wg::wgrad_tc_kernel<false>, wg::SMEM_BYTES);
    wg::Tiling t = wg::make_tiling((int)n, (int)k);
    const int64_t rps = cdiv(cdiv(rows, splits), wg::BK) * wg::BK;
    dim3 grid((unsigned)(t.ptiles * t.qtiles), (unsigned)splits);
    const int split3 = precision == 2 ? 0 : 1;
    if (relu_a)
        wg::wgrad_tc_kernel<true><<<grid, wg::THREADS, wg::SMEM_BYTES, st>>>(dY, lddy, A, lda, rows, t, rps, split3, part);
    else
        wg::wgrad_tc_kernel<false><<<grid, wg::THREADS, wg::SMEM_BYTES, st>>>(dY, lddy, A, lda, rows, t, rps, split3, part);
    O4D_LAUNCH_CHECK();
    return 0;
}

}  // namespace o4d
