// tcgen05 weight gradient:  part[s][p][q] = sum over the rows r of split s of dY[r, p] * pre(A[r, q])
// and (optionally, fused) the bias gradient  bpart[s][p] = sum over the same rows of dY[r, p].
//
// (dW = dY^T pre(A) of a dense layer Y = pre(A) W^T + b; in the reference this is the second SGEMM of autograd's
// AddmmBackward, db the column sum next to it.)  The contraction runs over ROWS, the slow dimension of both row-major
// operands, so the fp32 tiles cannot be bulk-copied straight into a K-major UMMA operand.  Round 1 had the converting
// warps load them from global memory into registers: 48 loads per thread and 32-row chunk that the compiler issues in
// dependent batches, two CTAs per SM, and ncu showed the result -- issue slots 24 % busy, long-scoreboard stalls 12.8
// of 19 cycles per instruction, 18,000 cycles per chunk against 620 of tensor work, 40 algorithmic TFLOP/s on the
// (240,842 x 416)^T (240,842 x 832) gradient.  This version separates the latency from the arithmetic:
//
//   warp 16 (loader)      : per 32-row chunk TWO tiled TMA loads (cp.async.bulk.tensor.2d through a tensor map per
//                           operand: a 32 x 128 box of dY, a 32 x BQ box of A) into an fp32 staging stage; rows past the
//                           end and columns past n / k arrive as zeros; completion by mbarrier transaction bytes; NS
//                           (2-4) stages in flight, so ~100 KB per SM are outstanding instead of one register file's
//                           worth.  (One cp.async.bulk per ROW was tried first: UBLKCP is a uniform-datapath instruction,
//                           the 64 copies of a chunk issue one after the other, 4,500 cycles per chunk.)
//   warps 0-15 (convert)  : read the staging stage along columns (conflict-free), split every value into bf16 hi + lo
//                           (packed conversions; same bf16x3 scheme as gemm_tc.cu: hi*hi + lo*hi + hi*lo into one fp32
//                           TMEM accumulator) and store eight consecutive ROWS of one column as one 16-byte core-matrix
//                           line of the UMMA stage -- the transpose costs nothing extra; they also carry the running
//                           column sums of dY for the bias gradient, and run the epilogue (TMEM -> fp32 partial tile)
//   warp 17 (MMA)         : lane 0 issues tcgen05.mma / commits
//   CTA tile              : 128 (p, TMEM lanes) x BQ <= 256 (q, TMEM columns); grid (p-tiles * q-tiles, row splits);
//                           tiles of one split are adjacent in launch order, so the operand rows they share come from
//                           HBM once and from L2 afterwards.  One CTA per SM (shared memory), at most two waves.
// Measured (tools/time_wgrad.py, stamps build): 1,580 cycles per chunk, the converters busy, the loader waiting on them,
// the MMA thread idle 46 % of the time -- 0.74 ms for the (240,842 x 416)^T (240,842 x 832) gradient (220 algorithmic
// TFLOP/s; round 1: 4.2 ms).  Tried and rejected: staging as 32-column boxes so that every converter read is base +
// immediate (234 instead of 293 instructions per chunk and warp) -- 11 tensor-map loads per chunk instead of 2, and one
// thread needs ~200 cycles to issue one (2,220 cycles per chunk, loader-bound); with four loader warps 1,690 cycles,
// still behind the two-box version.  The next lever is less conversion work per MMA (a CTA pair sharing the A tile).
// Partials are reduced in a fixed order by reduce_partials_kernel (train.cu): deterministic.
// Requirements (wgrad_tc_ok / wgrad_tc_aligned, else the caller falls back to the fp32 kernel): 16-byte aligned operand
// pointers, leading dimensions multiples of 4 floats (tensor-map strides), k a multiple of 4 (vector stores).
#include "o4d_common.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include "tc_helpers.cuh"

namespace o4d {
namespace wg {

using namespace tch;

constexpr int BK = 32;                  // rows (contraction) per chunk
constexpr int CONV_WARPS = 16;
constexpr int CONV_THREADS = CONV_WARPS * 32;
constexpr int THREADS = (CONV_WARPS + 2) * 32;
constexpr int BQ_MAX = 256;
constexpr int NU = 2;                   // UMMA operand stages
constexpr int NS_MAX = 4;               // fp32 staging stages
constexpr int MAX_ITEMS = (4 * (BM + BQ_MAX) + CONV_THREADS - 1) / CONV_THREADS;   // (8-row group, column) items per thread
constexpr int BAR_BYTES = 256;
constexpr int SMEM_LIMIT = 227 * 1024;
constexpr uint32_t A_HALF = BM * BK * 2;   // one bf16 image of the dY^T slab (8 KB)

struct Tiling {
    int n, k;       // dW shape (p extent, q extent)
    int ptiles;     // ceil(n / 128)
    int bq, qtiles; // q tile width (multiple of 16, <= 256) and count
    int ns;         // staging stages that fit
};

__host__ __device__ inline int stage_bytes(int bq) { return BK * (BM + bq) * 4; }   // fp32 staging == bf16 hi + lo images

inline Tiling make_tiling(int n, int k) {
    Tiling t;
    t.n = n;
    t.k = k;
    t.ptiles = (n + BM - 1) / BM;
    int qt = (k + BQ_MAX - 1) / BQ_MAX;
    int bq = (k + qt - 1) / qt;
    bq = (bq + 15) / 16 * 16;
    t.bq = bq;
    t.qtiles = (k + bq - 1) / bq;
    int ns = (SMEM_LIMIT - BAR_BYTES) / stage_bytes(bq) - NU;
    t.ns = ns > NS_MAX ? NS_MAX : ns;
    return t;
}

// Cycle counters of one CTA (O4D_STAMPS builds only, o4d_debug_read_wgrad): which wait bounds the pipeline.
//   [0] loader: total  [1] loader: wait for an empty staging stage   [2] converter warp 0: wait staging full
//   [3] converter warp 0: wait UMMA stage empty  [4] converter warp 0: total loop  [5] MMA: wait UMMA stage full
//   [6] MMA: total  [7] chunks
__device__ long long g_dbg_wg[8];

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c_inner, int c_outer, uint32_t mbar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
                 "l"(tm), "r"(mbar), "r"(c_inner), "r"(c_outer)
                 : "memory");
}

template <bool RELU>
__global__ void __launch_bounds__(THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tm_y, const __grid_constant__ CUtensorMap tm_a, int64_t rows, Tiling tl,
                int64_t rows_per_split, int split3, float* __restrict__ part, float* __restrict__ bpart) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bq = tl.bq, W = BM + bq, NS = tl.ns;
    const int sbytes = stage_bytes(bq);
    uint8_t* ustage0 = smem + NS * sbytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (NS + NU) * sbytes);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
    const uint32_t sfull0 = smem_u32(&bars[0]), sempty0 = smem_u32(&bars[NS_MAX]);
    const uint32_t ufull0 = smem_u32(&bars[2 * NS_MAX]), uempty0 = smem_u32(&bars[2 * NS_MAX + NU]);
    const uint32_t accum_bar = smem_u32(&bars[2 * NS_MAX + 2 * NU]);
    const int tile = blockIdx.x;
    const int pt = tile / tl.qtiles, qt = tile % tl.qtiles;
    const int p0 = pt * BM, q0 = qt * bq;
    const int64_t r_lo = (int64_t)blockIdx.y * rows_per_split;
    const int64_t r_hi = min(rows, r_lo + rows_per_split);
    const int nchunks = r_hi > r_lo ? (int)((r_hi - r_lo + BK - 1) / BK) : 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NS_MAX; ++s) {
            mbar_init(sfull0 + 8 * s, 1);
            mbar_init(sempty0 + 8 * s, CONV_WARPS);
        }
        for (int u = 0; u < NU; ++u) {
            mbar_init(ufull0 + 8 * u, CONV_WARPS);
            mbar_init(uempty0 + 8 * u, 1);
        }
        mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == CONV_WARPS) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t b_half = (uint32_t)bq * BK * 2;

    if (warp < CONV_WARPS) {
        // ------------------------------------------------------------ converters
        // item = (8-row group kc, column col of the staging row): eight floats down a column -> one core-matrix line
        const int t = threadIdx.x;
        int src[MAX_ITEMS], dst[MAX_ITEMS], pitch[MAX_ITEMS];
        bool is_y[MAX_ITEMS];
        float bsum[MAX_ITEMS];
#pragma unroll
        for (int i = 0; i < MAX_ITEMS; ++i) {
            const int item = t + i * CONV_THREADS;
            const int kc = item / W, col = item - kc * W;
            const bool in = item < 4 * W;
            const int c2 = col - BM;
            is_y[i] = col < BM;
            pitch[i] = is_y[i] ? BM : bq;                       // staging stage = [32][128] box of dY, then [32][bq] box of A
            src[i] = !in ? 0 : is_y[i] ? kc * 8 * BM + col : BK * BM + kc * 8 * bq + c2;
            dst[i] = !in ? -1
                         : is_y[i] ? kc * (BM * 16) + (col >> 3) * 128 + (col & 7) * 16
                                   : (int)(2 * A_HALF) + kc * (bq * 16) + (c2 >> 3) * 128 + (c2 & 7) * 16;
            bsum[i] = 0.f;
        }
        const bool want_bias = bpart != nullptr && qt == 0;
        const bool dbg = O4D_STAMPS && blockIdx.x == 0 && blockIdx.y == gridDim.y / 2 && t == 0;
        long long t_sfull = 0, t_uempty = 0, t_start = dbg ? clock64() : 0;
        for (int c = 0; c < nchunks; ++c) {
            const int s = c % NS, u = c % NU;
            const uint32_t sph = (uint32_t)(c / NS) & 1u, uph = (uint32_t)(c / NU) & 1u;
            long long tw = dbg ? clock64() : 0;
            mbar_wait(sfull0 + 8 * s, sph);
            if (dbg) t_sfull += clock64() - tw;
            const float* stg = reinterpret_cast<const float*>(smem + s * sbytes);
            float v[MAX_ITEMS][8];
#pragma unroll
            for (int i = 0; i < MAX_ITEMS; ++i) {
                if (dst[i] >= 0) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float x = stg[src[i] + e * pitch[i]];       // out-of-range rows / columns were zero-filled
                        v[i][e] = (RELU && !is_y[i]) ? fmaxf(x, 0.f) : x;
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(sempty0 + 8 * s);        // the staging values are in registers
            if (dbg) tw = clock64();
            mbar_wait(uempty0 + 8 * u, uph ^ 1u);
            if (dbg) t_uempty += clock64() - tw;
            uint8_t* ust = ustage0 + u * sbytes;
#pragma unroll
            for (int i = 0; i < MAX_ITEMS; ++i) {
                if (dst[i] >= 0) {
                    uint4 hi, lo;
                    if (split3) {
                        split_bf16x2(v[i][0], v[i][1], hi.x, lo.x);
                        split_bf16x2(v[i][2], v[i][3], hi.y, lo.y);
                        split_bf16x2(v[i][4], v[i][5], hi.z, lo.z);
                        split_bf16x2(v[i][6], v[i][7], hi.w, lo.w);
                        *reinterpret_cast<uint4*>(ust + dst[i] + (is_y[i] ? A_HALF : b_half)) = lo;
                    } else {                    // single-pass bf16 (precision 2): no low-order image at all
                        const __nv_bfloat162 h0 = __floats2bfloat162_rn(v[i][0], v[i][1]), h1 = __floats2bfloat162_rn(v[i][2], v[i][3]);
                        const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[i][4], v[i][5]), h3 = __floats2bfloat162_rn(v[i][6], v[i][7]);
                        hi.x = *reinterpret_cast<const uint32_t*>(&h0); hi.y = *reinterpret_cast<const uint32_t*>(&h1);
                        hi.z = *reinterpret_cast<const uint32_t*>(&h2); hi.w = *reinterpret_cast<const uint32_t*>(&h3);
                    }
                    *reinterpret_cast<uint4*>(ust + dst[i]) = hi;
                    if (want_bias && is_y[i])
                        bsum[i] += ((v[i][0] + v[i][1]) + (v[i][2] + v[i][3])) + ((v[i][4] + v[i][5]) + (v[i][6] + v[i][7]));
                }
            }
            fence_proxy_async_smem();   // generic-proxy stores -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(ufull0 + 8 * u);
        }
        if (dbg) { g_dbg_wg[2] = t_sfull; g_dbg_wg[3] = t_uempty; g_dbg_wg[4] = clock64() - t_start; g_dbg_wg[7] = nchunks; }
        // ------------------------------------------------------------ bias partial: fixed-order sum of the 4 row groups
        if (want_bias) {
            float* red = reinterpret_cast<float*>(smem);        // staging stage 0: every copy into it has been consumed
            named_bar_sync(1, CONV_THREADS);
#pragma unroll
            for (int i = 0; i < MAX_ITEMS; ++i) {
                const int item = t + i * CONV_THREADS;
                if (dst[i] >= 0 && is_y[i]) red[(item / W) * BM + (item % W)] = bsum[i];
            }
            named_bar_sync(1, CONV_THREADS);
            if (t < BM && p0 + t < tl.n)
                bpart[(int64_t)blockIdx.y * tl.n + p0 + t] = (red[t] + red[BM + t]) + (red[2 * BM + t] + red[3 * BM + t]);
        }
        // ------------------------------------------------------------ epilogue
        // warp w: TMEM lane quarter (w & 3), column quarter (w >> 2); one output row per thread
        float* dstp = part + (int64_t)blockIdx.y * tl.n * tl.k;
        const int m = (warp & 3) * 32 + lane;
        const int gp = p0 + m;
        const int cq = ((bq + 15) / 16 + 3) / 4 * 16;
        const int c_lo = (warp >> 2) * cq;
        const int c_hi = min(c_lo + cq, bq);
        if (nchunks > 0) {
            mbar_wait(accum_bar, 0);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
            for (int c0 = c_lo; c0 < c_hi; c0 += 16) {
                float o[16];
                tmem_ld16(taddr + (uint32_t)c0, o);
                if (gp < tl.n) {
                    float* row = dstp + (int64_t)gp * tl.k + q0 + c0;
                    if (q0 + c0 + 16 <= tl.k) {
#pragma unroll
                        for (int i = 0; i < 16; i += 4)
                            *reinterpret_cast<float4*>(row + i) = make_float4(o[i], o[i + 1], o[i + 2], o[i + 3]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (q0 + c0 + i < tl.k) row[i] = o[i];
                    }
                }
            }
            tc_fence_before();
        } else if (gp < tl.n) {
            for (int c0 = c_lo; c0 < c_hi; ++c0)
                if (q0 + c0 < tl.k) dstp[(int64_t)gp * tl.k + q0 + c0] = 0.f;
        }
    } else if (warp == CONV_WARPS) {
        // ------------------------------------------------------------ loader: two tiled TMA loads per chunk
        const bool dbg = O4D_STAMPS && blockIdx.x == 0 && blockIdx.y == gridDim.y / 2 && lane == 0;
        long long t_wait = 0, t_start = dbg ? clock64() : 0;
        if (lane == 0) {
            for (int c = 0; c < nchunks; ++c) {
                const int s = c % NS;
                const uint32_t ph = (uint32_t)(c / NS) & 1u;
                const long long tw = dbg ? clock64() : 0;
                mbar_wait(sempty0 + 8 * s, ph ^ 1u);
                if (dbg) t_wait += clock64() - tw;
                const int r0 = (int)(r_lo + (int64_t)c * BK);
                const uint32_t dsts = smem_u32(smem + s * sbytes);
                mbar_arrive_expect_tx(sfull0 + 8 * s, (uint32_t)sbytes);
                tma_load_2d(dsts, &tm_y, p0, r0, sfull0 + 8 * s);
                tma_load_2d(dsts + BK * BM * 4, &tm_a, q0, r0, sfull0 + 8 * s);
            }
        }
        if (dbg) { g_dbg_wg[0] = clock64() - t_start; g_dbg_wg[1] = t_wait; }
    } else {
        // ------------------------------------------------------------ MMA issuer
        if (lane == 0 && nchunks > 0) {
            const uint32_t idesc = umma_idesc(bq);
            const uint32_t lbo_a = BM * 16, lbo_b = (uint32_t)bq * 16;
            const uint32_t ubase = smem_u32(ustage0);
            const bool dbg = O4D_STAMPS && blockIdx.x == 0 && blockIdx.y == gridDim.y / 2;
            long long t_wait = 0, t_start = dbg ? clock64() : 0;
            for (int c = 0; c < nchunks; ++c) {
                const int u = c % NU;
                const uint32_t ph = (uint32_t)(c / NU) & 1u;
                const long long tw = dbg ? clock64() : 0;
                mbar_wait(ufull0 + 8 * u, ph);
                if (dbg) t_wait += clock64() - tw;
                tc_fence_after();
                const uint32_t a_hi = ubase + u * sbytes;
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
                umma_commit(uempty0 + 8 * u);
            }
            umma_commit(accum_bar);
            if (dbg) { g_dbg_wg[5] = t_wait; g_dbg_wg[6] = clock64() - t_start; }
        }
    }
    __syncthreads();
    if (warp == CONV_WARPS) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
    }
}

}  // namespace wg

bool wgrad_tc_ok(int64_t rows, int64_t n, int64_t k) {
    return rows >= 1024 && n >= 32 && k >= 32 && n <= 8192 && k <= 8192 && k % 4 == 0;
}

// Tensor maps need a 16-byte aligned base and row strides that are multiples of 16 bytes.
bool wgrad_tc_aligned(const float* dY, int64_t lddy, const float* A, int64_t lda) {
    return ((uintptr_t)dY % 16 == 0) && ((uintptr_t)A % 16 == 0) && lddy % 4 == 0 && lda % 4 == 0;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda).
typedef CUresult (*TensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TensorMapEncodeFn tensor_map_encoder() {
    static TensorMapEncodeFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (TensorMapEncodeFn)p;
    }();
    return fn;
}

// fp32 row-major (rows, cols) matrix with leading dimension ld; box = 32 rows x box_cols columns, zero fill outside.
static int make_row_map(CUtensorMap* tm, const float* X, int64_t rows, int64_t cols, int64_t ld, int box_cols) {
    TensorMapEncodeFn enc = tensor_map_encoder();
    O4D_REQUIRE(enc != nullptr, "weight gradient: cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)wg::BK};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(X), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    O4D_REQUIRE(r == CUDA_SUCCESS, "weight gradient: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return 0;
}

int wgrad_tc_splits(int64_t rows, int64_t n, int64_t k) {
    wg::Tiling t = wg::make_tiling((int)n, (int)k);
    const int64_t tiles = (int64_t)t.ptiles * t.qtiles;
    int64_t s = (148 * 2) / tiles;                   // at most two full waves of one CTA per SM (no third, ragged wave)
    const int64_t max_by_rows = cdiv(rows, 1024);    // at least 32 chunks per CTA
    if (s > max_by_rows) s = max_by_rows;
    if (s > 256) s = 256;
    if (s < 1) s = 1;
    return (int)s;
}

// part must hold splits * n * k floats (wgrad_tc_splits); bpart (optional) splits * n floats: column sums of dY.
int wgrad_tc_launch(const float* dY, int64_t lddy, const float* A, int64_t lda, int64_t rows, int64_t n, int64_t k,
                    bool relu_a, int precision, float* part, float* bpart, int splits, cudaStream_t st) {
    wg::Tiling t = wg::make_tiling((int)n, (int)k);
    O4D_REQUIRE(t.ns >= 2, "weight gradient: tile does not fit shared memory");
    const int smem_bytes = (t.ns + wg::NU) * wg::stage_bytes(t.bq) + wg::BAR_BYTES;
    O4D_SMEM_ATTR(wg::wgrad_tc_kernel<true>, wg::SMEM_LIMIT);
    O4D_SMEM_ATTR(wg::wgrad_tc_kernel<false>, wg::SMEM_LIMIT);
    const int64_t rps = cdiv(cdiv(rows, splits), wg::BK) * wg::BK;
    dim3 grid((unsigned)(t.ptiles * t.qtiles), (unsigned)splits);
    const int split3 = precision == 2 ? 0 : 1;
    CUtensorMap tm_y, tm_a;
    O4D_TRY(make_row_map(&tm_y, dY, rows, n, lddy, wg::BM));
    O4D_TRY(make_row_map(&tm_a, A, rows, k, lda, t.bq));
    if (relu_a)
        wg::wgrad_tc_kernel<true><<<grid, wg::THREADS, smem_bytes, st>>>(tm_y, tm_a, rows, t, rps, split3, part, bpart);
    else
        wg::wgrad_tc_kernel<false><<<grid, wg::THREADS, smem_bytes, st>>>(tm_y, tm_a, rows, t, rps, split3, part, bpart);
    O4D_LAUNCH_CHECK();
    return 0;
}

}  // namespace o4d

extern "C" int o4d_debug_read_wgrad(long long* out8) {
    return (int)cudaMemcpyFromSymbol(out8, o4d::wg::g_dbg_wg, sizeof(long long) * 8);
}
