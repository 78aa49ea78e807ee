// fp32 CUDA-core dense layer:  C = post(pre(A) @ W^T + bias) [+ R]
//
// Replaces torch.nn.Linear (cuBLAS SGEMM in the reference) wherever bit-for-bit fp32
// FMA arithmetic is wanted (precision 0) and for shapes the tcgen05 path does not take
// (tiny n/k, unaligned views such as the abstract-feature slice with ld = 3 + E).
// Register-tiled 16x16 threads, (16*TM) x (16*TN) output tile, BK-deep shared-memory
// slabs stored k-major so the inner product reads broadcast/contiguous float4s.
#include "o4d_common.cuh"

namespace o4d {

constexpr int GS_BK = 16;

template <int TM, int TN>
__global__ void __launch_bounds__(256)
linear_simt_kernel(const float* __restrict__ A, int64_t rows, int k, int64_t lda,
                   const float* __restrict__ W, int64_t ldw, const float* __restrict__ bias, int n,
                   const float* R, int64_t ldr, float* C, int64_t ldc, int flags, RowGather g) {
    constexpr int BM = 16 * TM, BN = 16 * TN;
    __shared__ __align__(16) float As[GS_BK][BM + 4];
    __shared__ __align__(16) float Ws[GS_BK][BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;
    const int64_t row0 = (int64_t)blockIdx.x * BM;
    const int col0 = blockIdx.y * BN;
    const bool relu_in = flags & O4D_RELU_IN;

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < k; k0 += GS_BK) {
        // A slab: BM rows x BK, consecutive threads walk k first (contiguous in memory).
        for (int e = tid; e < BM * GS_BK; e += 256) {
            int kk = e % GS_BK, r = e / GS_BK;
            int64_t gr = row0 + r;
            float v = 0.f;
            if (gr < rows && k0 + kk < k) {
                v = A[gr * lda + k0 + kk];
                if (relu_in) v = fmaxf(v, 0.f);
            }
            As[kk][r] = v;
        }
        for (int e = tid; e < BN * GS_BK; e += 256) {
            int kk = e % GS_BK, c = e / GS_BK;
            int gc = col0 + c;
            float v = 0.f;
            if (gc < n && k0 + kk < k) v = W[(int64_t)gc * ldw + k0 + kk];
            Ws[kk][c] = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < GS_BK; ++kk) {
            float a[TM], w[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = As[kk][ty + 16 * i];
#pragma unroll
            for (int j = 0; j < TN; ++j) w[j] = Ws[kk][tx + 16 * j];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
        }
        __syncthreads();
    }

    const bool relu_out = flags & O4D_RELU_OUT;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int64_t gr = row0 + ty + 16 * i;
        if (gr >= rows) continue;
        const float* gq = nullptr;
        const float* gk = nullptr;
        if (g.qa) {
            const int64_t ar = g.row_offset + gr;
            gq = g.qa + (ar / g.knbr) * n;
            gk = g.ka + g.neighbour(ar) * n;
        }
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int gc = col0 + tx + 16 * j;
            if (gc >= n) continue;
            float v = acc[i][j];
            if (bias) v += bias[gc];
            if (gq) v += gq[gc] - gk[gc];
            if (relu_out) v = fmaxf(v, 0.f);
            if (R) v = (flags & O4D_MASK_RES) ? (R[gr * ldr + gc] > 0.f ? v : 0.f) : v + R[gr * ldr + gc];
            C[gr * ldc + gc] = v;
        }
    }
}

int linear_simt_launch(const float* A, int64_t rows, int64_t k, int64_t lda, const float* W,
                       int64_t ldw, const float* bias, int64_t n, const float* R, int64_t ldr,
                       float* C, int64_t ldc, int flags, cudaStream_t st, const RowGather* gp = nullptr) {
    if (rows == 0) return 0;
    RowGather g;
    if (gp) g = *gp;
    if (n > 96 && rows >= 8192) {
        dim3 grid((unsigned)cdiv(rows, 128), (unsigned)cdiv(n, 128));
        linear_simt_kernel<8, 8><<<grid, 256, 0, st>>>(A, rows, (int)k, lda, W, ldw, bias, (int)n, R, ldr, C, ldc, flags, g);
    } else if (n > 48) {
        dim3 grid((unsigned)cdiv(rows, 128), (unsigned)cdiv(n, 64));
        linear_simt_kernel<8, 4><<<grid, 256, 0, st>>>(A, rows, (int)k, lda, W, ldw, bias, (int)n, R, ldr, C, ldc, flags, g);
    } else if (rows >= 4096) {
        dim3 grid((unsigned)cdiv(rows, 128), (unsigned)cdiv(n, 16));
        linear_simt_kernel<8, 1><<<grid, 256, 0, st>>>(A, rows, (int)k, lda, W, ldw, bias, (int)n, R, ldr, C, ldc, flags, g);
    } else {
        dim3 grid((unsigned)cdiv(rows, 16), (unsigned)cdiv(n, 16));
        linear_simt_kernel<1, 1><<<grid, 256, 0, st>>>(A, rows, (int)k, lda, W, ldw, bias, (int)n, R, ldr, C, ldc, flags, g);
    }
    O4D_LAUNCH_CHECK();
    return 0;
}

int linear_ldw_launch(const float* A, int64_t rows, int64_t k, int64_t lda, const float* W,
                      int64_t ldw, const float* bias, int64_t n, const float* R, int64_t ldr,
                      float* C, int64_t ldc, int flags, int precision, cudaStream_t st);

int linear_launch(const float* A, int64_t rows, int64_t k, int64_t lda, const float* W,
                  const float* bias, int64_t n, const float* R, int64_t ldr, float* C, int64_t ldc,
                  int flags, int precision, cudaStream_t st) {
    return linear_ldw_launch(A, rows, k, lda, W, k, bias, n, R, ldr, C, ldc, flags, precision, st);
}

// W with an explicit leading dimension (column slices of a wider weight, e.g. the local
// half of lin_z).
int linear_ldw_launch(const float* A, int64_t rows, int64_t k, int64_t lda, const float* W,
                      int64_t ldw, const float* bias, int64_t n, const float* R, int64_t ldr,
                      float* C, int64_t ldc, int flags, int precision, cudaStream_t st) {
    O4D_REQUIRE(A && W && C, "linear: null pointer");
    O4D_REQUIRE(rows >= 0 && k >= 1 && n >= 1, "linear: bad shape rows=%lld k=%lld n=%lld",
                (long long)rows, (long long)k, (long long)n);
    O4D_REQUIRE(lda >= k && ldw >= k && ldc >= n && (!R || ldr >= n), "linear: bad leading dimension");
    O4D_REQUIRE(precision >= 0 && precision <= 2, "linear: precision %d not in {0,1,2}", precision);
    return linear_ps_launch(nullptr, A, rows, k, lda, W, ldw, bias, n, R, ldr, C, ldc, flags, precision, st);
}

int linear_ps_launch(const PackedSet* ps, const float* A, int64_t rows, int64_t k, int64_t lda, const float* W,
                     int64_t ldw, const float* bias, int64_t n, const float* R, int64_t ldr, float* C,
                     int64_t ldc, int flags, int precision, cudaStream_t st, const RowGather* g) {
    ProfScope prof(PROF_LINEAR, 2.0 * (double)rows * (double)(k + (g ? g->k2 : 0)) * (double)n, st);
    if (precision != 0 && tc_shape_ok(rows, k, n)) {
        const void* packed = (ps && !ps->pair) ? ps->find(W) : nullptr;
        if (packed)
            return linear_tc_packed_launch(A, rows, k, lda, packed, n, bias, R, ldr, C, ldc, flags, precision, st, g);
    }
    if (g && g->a2) {
        set_error("linear: a K-concatenated second operand needs the tcgen05 path with a pre-packed weight");
        return O4D_E_UNSUPPORTED;
    }
    if (precision != 0 && tc_shape_ok(rows, k, n)) {
        if (ldw == k) {
            int rc = linear_tc_launch(A, rows, k, lda, W, bias, n, R, ldr, C, ldc, flags, precision, st, g);
            if (rc != O4D_E_UNSUPPORTED) return rc;
        }
    }
    return linear_simt_launch(A, rows, k, lda, W, ldw, bias, n, R, ldr, C, ldc, flags, st, g);
}

}  // namespace o4d

extern "C" int o4d_linear_f32(const float* A, int64_t rows, int64_t k, int64_t lda, const float* W,
                              const float* bias, int64_t n, const float* R, int64_t ldr, float* C,
                              int64_t ldc, int flags, int precision, void* stream) {
    return o4d::linear_launch(A, rows, k, lda, W, bias, n, R, ldr, C, ldc, flags, precision,
                              (cudaStream_t)stream);
}

// ---- caller-owned packed weights (no allocation, no re-packing inside the compute call) ----------------------------
extern "C" size_t o4d_linear_pack_bytes(int64_t rows, int64_t k, int64_t n, int precision) {
    if (precision == 0 || !o4d::tc_shape_ok(rows, k, n)) return 0;      // this shape runs on the CUDA-core kernel: nothing to pack
    return o4d::tc_pack_bytes(n, k);
}

extern "C" int o4d_linear_pack_f32(const float* W, int64_t n, int64_t k, int64_t ldw, void* packed, void* stream) {
    using namespace o4d;
    O4D_REQUIRE(W && packed && n >= 1 && k >= 1 && ldw >= k, "linear pack: bad argument");
    return tc_pack_launch(W, n, k, ldw, packed, (cudaStream_t)stream);
}

extern "C" int o4d_linear_packed_f32(const float* A, int64_t rows, int64_t k, int64_t lda, const void* packed,
                                     const float* bias, int64_t n, const float* R, int64_t ldr, float* C, int64_t ldc,
                                     int flags, int precision, void* stream) {
    using namespace o4d;
    O4D_REQUIRE(A && packed && C, "linear (packed): null pointer");
    O4D_REQUIRE(rows >= 0 && k >= 1 && n >= 1 && lda >= k && ldc >= n && (!R || ldr >= n), "linear (packed): bad shape");
    O4D_REQUIRE(precision == 1 || precision == 2, "linear (packed): the packed form is the tcgen05 path (precision 1 or 2)");
    if (!tc_shape_ok(rows, k, n)) {
        set_error("linear (packed): shape (%lld, %lld, %lld) is not on the tensor-core path (o4d_linear_pack_bytes == 0)",
                  (long long)rows, (long long)k, (long long)n);
        return O4D_E_UNSUPPORTED;
    }
    if (rows == 0) return 0;
    ProfScope prof(PROF_LINEAR, 2.0 * (double)rows * (double)k * (double)n, (cudaStream_t)stream);
    return linear_tc_packed_launch(A, rows, k, lda, packed, n, bias, R, ldr, C, ldc, flags, precision, (cudaStream_t)stream);
}
