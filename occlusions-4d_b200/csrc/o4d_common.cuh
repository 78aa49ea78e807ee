// Shared helpers for libo4d.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>
#include "../../include/o4d.h"

namespace o4d {

void set_error(const char* fmt, ...);

#define O4D_REQUIRE(cond, ...)                         \
    do {                                               \
        if (!(cond)) {                                 \
            ::o4d::set_error(__VA_ARGS__);             \
            return O4D_E_ARG;                          \
        }                                              \
    } while (0)

#define O4D_CUDA(expr)                                                              \
    do {                                                                            \
        cudaError_t e__ = (expr);                                                   \
        if (e__ != cudaSuccess) {                                                   \
            ::o4d::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                             __FILE__, __LINE__);                                   \
            return (int)e__;                                                        \
        }                                                                           \
    } while (0)

// In-kernel cycle stamps (clock64 into __device__ globals, read by o4d_debug_read*) are a bottleneck-hunting
// tool: compiled OUT of release builds.  `make STAMPS=1` (-DO4D_STAMPS=1) turns them on for tools/stamps_*.py.
#ifndef O4D_STAMPS
#define O4D_STAMPS 0
#endif

void count_launch();  // diagnostics: kernels launched by this library (o4d_launch_count)

// Opt-in kernel-family timer (o4d_profile_enable): CUDA events on the launching stream.
enum { PROF_LINEAR = 0, PROF_KNN = 1, PROF_FPS = 2, PROF_ATTN_GLUE = 3, PROF_MISC = 4, PROF_FUSED = 5 };
struct ProfScope {
    ProfScope(int family, double flops, cudaStream_t st);
    ~ProfScope();
    int rec_;
    cudaStream_t st_;
};

#define O4D_LAUNCH_CHECK()                                                          \
    do {                                                                            \
        ::o4d::count_launch();                                                      \
        cudaError_t e__ = cudaGetLastError();                                       \
        if (e__ != cudaSuccess) {                                                   \
            ::o4d::set_error("kernel launch failed: %s (%s:%d)",                    \
                             cudaGetErrorString(e__), __FILE__, __LINE__);          \
            return (int)e__;                                                        \
        }                                                                           \
    } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute: set it once per (call site, device).
// A process-wide flag would leave the second GPU a process uses (nn.DataParallel in train.py:305) on the 48 KB
// default.  One byte per device, written after the call succeeded; racing threads at worst both set it.
#define O4D_SMEM_ATTR(kernel, bytes)                                                               \
    do {                                                                                           \
        static std::atomic<unsigned char> done__[64];                                              \
        int dev__ = 0;                                                                             \
        O4D_CUDA(cudaGetDevice(&dev__));                                                           \
        const bool slot__ = dev__ >= 0 && dev__ < 64;                                              \
        if (!slot__ || !done__[dev__].load(std::memory_order_acquire)) {                           \
            O4D_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
            if (slot__) done__[dev__].store(1, std::memory_order_release);                         \
        }                                                                                          \
    } while (0)

// Internal epilogue flag of the dense-layer kernels (not part of include/o4d.h): the "residual" operand R is a ReLU
// mask -- C = R > 0 ? result : 0 -- the backward of a ReLU in front of the layer, fused into its input-gradient GEMM.
#define O4D_MASK_RES 8

#define O4D_TRY(expr)                  \
    do {                               \
        int rc__ = (expr);             \
        if (rc__ != 0) return rc__;    \
    } while (0)

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Bump allocator over a caller-owned workspace.  With base == nullptr it only measures.
struct Arena {
    char* base;
    size_t cap;
    size_t off;
    bool ok;
    Arena(void* b, size_t c) : base((char*)b), cap(c), off(0), ok(true) {}
    template <typename T>
    T* get(size_t count) {
        size_t bytes = align_up(count * sizeof(T), 256);
        size_t start = off;
        off += bytes;
        if (base == nullptr) return nullptr;
        if (off > cap) {
            ok = false;
            return nullptr;
        }
        return (T*)(base + start);
    }
};

// Optional row-dependent epilogue term of a dense layer: before the output ReLU,
//   C[r, c] += qa[(row_offset + r) / knbr, c] - ka[nbr[row_offset + r], c]
// (rows are (query, neighbour) pairs; qa / ka have the layer's n columns).
struct RowGather {
    __host__ __device__ int64_t neighbour(int64_t r) const { return nbr ? (int64_t)nbr[r] : nbr64[r]; }
    const float* qa = nullptr;
    const float* ka = nullptr;
    const int32_t* nbr = nullptr;
    const int64_t* nbr64 = nullptr;   // the same list as int64 (training path); exactly one of nbr / nbr64 is set
    int knbr = 1;
    int64_t row_offset = 0;
    // Optional SECOND A operand, concatenated along K behind the first one (tcgen05 path with a
    // pre-packed weight only):  C = post(pre(A) W[:, :k1p]^T + A2 W[:, k1p:]^T + bias) [+ R],
    // k1p = k rounded up to 32.  O4D_RELU_IN applies to A only.  Used to fold every lin_z of the
    // decoder (x += W_z f_local) into the layer that produces x.
    const float* a2 = nullptr;
    int64_t lda2 = 0;
    int k2 = 0;
};

// ---- internal launchers shared between translation units (all async on `st`) ----
int knn_launch(const float* query, int64_t nq, int64_t ldq, const float* ref, int64_t m,
               int64_t ldr, int k, int sqrt_dist, int32_t* idx32, int64_t* idx64, float* dist,
               cudaStream_t st);
// one scan -> the k nearest by squared distance and the k2 <= k nearest by Euclidean distance (+ distances)
int knn_dual_launch(const float* query, int64_t nq, int64_t ldq, const float* ref, int64_t m, int64_t ldr, int k,
                    int32_t* idx32, int k2, int32_t* idx2, float* dist2, cudaStream_t st);
int fps_launch(const float* xyz, int64_t n, int64_t ld, int64_t n_out, int64_t start,
               int32_t* idx_sorted32, int64_t* idx_sorted64, int64_t* order64, void* ws,
               size_t ws_bytes, cudaStream_t st);
int linear_launch(const float* A, int64_t rows, int64_t k, int64_t lda, const float* W,
                  const float* bias, int64_t n, const float* R, int64_t ldr, float* C,
                  int64_t ldc, int flags, int precision, cudaStream_t st);
int linear_ldw_launch(const float* A, int64_t rows, int64_t k, int64_t lda, const float* W,
                      int64_t ldw, const float* bias, int64_t n, const float* R, int64_t ldr,
                      float* C, int64_t ldc, int flags, int precision, cudaStream_t st);
// tcgen05 path (gemm_tc.cu); returns O4D_E_UNSUPPORTED when the shape cannot use it.
int linear_tc_launch(const float* A, int64_t rows, int64_t k, int64_t lda, const float* W,
                     const float* bias, int64_t n, const float* R, int64_t ldr, float* C,
                     int64_t ldc, int flags, int precision, cudaStream_t st, const RowGather* g = nullptr);

// Pre-packed (bf16 hi/lo, shared-memory image) weights for the tcgen05 path, looked up by the
// fp32 weight pointer they were packed from.
size_t tc_pack_bytes(int64_t n, int64_t k);
int tc_pack_launch(const float* W, int64_t n, int64_t k, int64_t ldw, void* packed, cudaStream_t st);
// same bytes, pair format of the fused multi-layer kernel (cta_group::2: each CTA of a pair copies half of a slab)
int tc_pack_pair_launch(const float* W, int64_t n, int64_t k, int64_t ldw, void* packed, cudaStream_t st);
bool tc_shape_ok(int64_t rows, int64_t k, int64_t n);
int linear_tc_packed_launch(const float* A, int64_t rows, int64_t k, int64_t lda, const void* packed, int64_t n,
                            const float* bias, const float* R, int64_t ldr, float* C, int64_t ldc, int flags,
                            int precision, cudaStream_t st, const RowGather* g = nullptr);
struct PackedSet {
    static constexpr int CAP = 96;
    const float* key[CAP];
    const void* packed[CAP];
    int count = 0;
    bool pair = false;   // packed in the pair format of the fused multi-layer kernel: not readable by the per-layer kernel
    void add(const float* w, const void* p) {
        if (count < CAP) { key[count] = w; packed[count] = p; ++count; }
    }
    const void* find(const float* w) const {
        for (int i = 0; i < count; ++i)
            if (key[i] == w) return packed[i];
        return nullptr;
    }
};
// linear_ldw_launch that prefers a pre-packed weight from `ps` (may be null).
int linear_ps_launch(const PackedSet* ps, const float* A, int64_t rows, int64_t k, int64_t lda, const float* W,
                     int64_t ldw, const float* bias, int64_t n, const float* R, int64_t ldr, float* C,
                     int64_t ldc, int flags, int precision, cudaStream_t st, const RowGather* g = nullptr);

struct PtBlockParams {
    const float *w1, *b1, *wq, *wk, *wv, *wp1, *bp1, *wp2, *bp2, *wa1, *ba1, *wa2, *ba2, *w3, *b3;
    const PackedSet* ps = nullptr;  // optional pre-packed tcgen05 weights
    static PtBlockParams from(const float* const* p) {
        PtBlockParams r;
        r.w1 = p[0]; r.b1 = p[1]; r.wq = p[2]; r.wk = p[3]; r.wv = p[4];
        r.wp1 = p[5]; r.bp1 = p[6]; r.wp2 = p[7]; r.bp2 = p[8];
        r.wa1 = p[9]; r.ba1 = p[10]; r.wa2 = p[11]; r.ba2 = p[12]; r.w3 = p[13]; r.b3 = p[14];
        return r;
    }
};

// Attention core shared by the encoder (self) and decoder (cross):
//   out = x_res + W3 . attn(q, ktab, vtab, pos, pos2, nbr) + b3      (n rows, d channels)
// q (n,d) already projected; ktab/vtab (m,d); nbr (n,k) int32.
// Per key cloud: V table (m,d), Ka = K W_a1^T (m,2d), Wc = W_a1 W_p2 (2d,32), cvec = W_a1 b_p2 + b_a1.
struct AttnTables {
    const float* vtab;
    const float* ka;
    const float* wc;
    const float* cvec;
    const void* fused = nullptr;   // packed weight images of the fused tcgen05 kernel (attn_fused.cu) or null
    // optional composite of everything between the block input x and Qa (decoder):
    //   Qa = (W_a1 W_q W_1) x + (W_a1 W_q b_1 + cvec);  when set, attn_core takes x instead of q
    const float* wqa = nullptr;
    const float* bqa = nullptr;
    // optional K-concatenated layer3 (decoder): [W_3 | W_z,local of the next block], b_3 + zg, second operand
    const float* w3cat = nullptr;
    const float* b3cat = nullptr;
    const float* cat_a2 = nullptr;
    int64_t cat_lda2 = 0;
    int cat_k2 = 0;
};
bool attn_fused_supported(int d, int k);
size_t attn_fused_pack_bytes(int d);
int attn_fused_pack_launch(const float* wc, const float* wa2, const float* wp2, int d, void* packed, cudaStream_t st);
int attn_fused_launch(const PtBlockParams& P, const AttnTables& T, const float* qa, const float* pos, int64_t ldpos,
                      const float* pos2, int64_t ldpos2, const int32_t* nbr, int64_t n, int d, int k, float* out,
                      int precision, cudaStream_t st);
size_t attn_tables_bytes(int64_t m, int d);
int attn_tables_launch(const PtBlockParams& P, const float* ktab, const float* vtab, int64_t m, int d, void* buf,
                       size_t buf_bytes, AttnTables* out, cudaStream_t st, bool weights_too = true);
int matmul_nn_launch(const float* A, int lda, const float* B, int ldb, const float* addvec, float* C, int p, int q, int r,
                     cudaStream_t st);
size_t attn_core_workspace_bytes(int64_t n, int d, int k);
int attn_core_launch(const PtBlockParams& P, const float* q, const AttnTables& T, const float* pos, int64_t ldpos,
                     const float* pos2, int64_t ldpos2, const int32_t* nbr, int64_t n, int d, int k,
                     const float* x_res, float* out, int precision, void* ws, size_t ws_bytes, cudaStream_t st);

size_t pt_block_ws(int64_t n, int64_t m, int d, int k, bool self_mode);
int pt_block_launch(const float* const* p, const float* x, int64_t n, int d, const float* pos,
                    int64_t ldpos, const float* x2, int64_t m, int d2, int64_t ldx2, const float* pos2,
                    int64_t ldpos2, int k, int precision, float* z, int64_t* knn_idx_out, void* ws,
                    size_t ws_bytes, cudaStream_t st);
size_t down_ws(int64_t n, int d_in, int d_out, int factor, int k);
int down_launch(const float* const* p, const float* x, int64_t n, int d_in, const float* pos,
                int64_t ldpos, int d_out, int factor, int k, int norm, int64_t start_idx, int precision,
                float* z, float* pos_out, int64_t* fps_idx_out, void* ws, size_t ws_bytes,
                cudaStream_t st);

// misc elementwise / gather kernels (misc.cu)
int posenc_launch(const float* q, int64_t n, int d_in, int n_freq, float* out, cudaStream_t st);
// the same features as the activation image of the fused multi-layer kernel (no fp32 copy)
int posenc_image_launch(const float* q, int64_t n, int d_in, int n_freq, void* img, cudaStream_t st);
// img != nullptr: the blend is written as an activation image (ceil(e / 32) chunks per 128-row tile) instead of fp32 rows
int local_blend_launch(const int32_t* idx, const float* dist, const float* feat, int64_t ldfeat,
                       int64_t n, int k, int e, float* out, int64_t ldout, cudaStream_t st, void* img = nullptr);
int gather_max_launch(const float* y, int64_t ldy, const int32_t* nbr, int64_t n_out, int k, int d,
                      float* z, cudaStream_t st);
int layernorm_relu_launch(float* y, int64_t rows, int d, const float* gamma, const float* beta,
                          float eps, cudaStream_t st);
int gather_rows_launch(const float* src, int64_t ldsrc, const int32_t* idx, int64_t n, int d,
                       float* dst, int64_t lddst, cudaStream_t st);
int col_mean_launch(const float* x, int64_t rows, int d, float* out, cudaStream_t st);
int copy2d_launch(const float* src, int64_t ldsrc, int64_t rows, int cols, float* dst,
                  int64_t lddst, cudaStream_t st);
int fill_col_launch(float* dst, int64_t ld, int64_t rows, int col, float value, cudaStream_t st);
int widen_idx_launch(const int32_t* in, int64_t count, int64_t* out, cudaStream_t st);

}  // namespace o4d
