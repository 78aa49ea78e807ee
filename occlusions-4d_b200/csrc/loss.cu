// Implicit loss heads fused over one frame's decoder output.
//
// Replaces  MyLosses.implicit_density_loss  loss.py:50-64    BCE-with-logits on column 0, mean over all rows
//           MyLosses.implicit_color_loss    loss.py:66-154   L1 on RGB | hue CE + sat / val L1 | 9-bin CE, over
//                                                            solid rows whose colour is available
//           MyLosses.implicit_segm_loss     loss.py:156-173  CE on the last semantic_classes columns, rows with tag >= 0
//           MyLosses.implicit_track_loss    loss.py:175-194  BCE-with-logits on the track column, solid rows with a label
//           utils.rgb_to_hsv                utils/utils.py:169-191
//
// The reference runs ~20 small kernels per head (masks, boolean-mask indexing with a host sync each, HSV
// conversion, CE, mean) and reads the (n, g) output four times.  Here: one pass over output + target that
// accumulates every head's sum and count (per-thread fp64, fixed-order block and grid reduction ->
// run-to-run deterministic), a one-block finalise that turns them into the four scalar losses on the device,
// and one pass that writes d loss / d output for all heads at once (each row read once, written once).
// HBM-bound: algorithmic bytes per row = 4 * (g + 6) forward, 4 * (2 g + 6) backward.
#include "o4d_common.cuh"
#include "loss_core.cuh"

namespace o4d {
namespace lossk {

constexpr int THREADS = 256;
__global__ void __launch_bounds__(THREADS)
loss_partial_kernel(const float* __restrict__ out, int64_t ldo, const float* __restrict__ tgt, int64_t ldt,
                    int64_t n, Head hd, double* __restrict__ partial) {
    double acc[NSTAT];
#pragma unroll
    for (int s = 0; s < NSTAT; ++s) acc[s] = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * THREADS) {
        const float* o = out + i * ldo;
        const float* t = tgt + i * ldt;
        row_accumulate(hd, o, t, acc);
    }
    // fixed-order block reduction: lanes by shuffle, warps through shared memory
    __shared__ double warp_acc[THREADS / 32][NSTAT];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int s = 0; s < NSTAT; ++s) {
        double v = acc[s];
        for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
        if (lane == 0) warp_acc[warp][s] = v;
    }
    __syncthreads();
    if (threadIdx.x < NSTAT) {
        double v = 0.0;
        for (int w = 0; w < THREADS / 32; ++w) v += warp_acc[w][threadIdx.x];
        partial[(int64_t)blockIdx.x * NSTAT + threadIdx.x] = v;
    }
}

// stats = sum of the per-block partials in block order; losses4 = (rgb, dens, segm, track) as loss.py returns them.
__global__ void loss_finalize_kernel(const double* __restrict__ partial, int blocks, Head hd,
                                     double* __restrict__ stats, float* __restrict__ losses4) {
    __shared__ double st[NSTAT];
    if (threadIdx.x < NSTAT) {
        double v = 0.0;
        for (int b = 0; b < blocks; ++b) v += partial[(int64_t)b * NSTAT + threadIdx.x];
        st[threadIdx.x] = v;
        stats[threadIdx.x] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        finalize_losses(hd, st, losses4);
    }
}

// d(sum_h w_h * loss_h) / d output, one thread per row; w = dlosses4 (rgb, dens, segm, track).
__global__ void __launch_bounds__(THREADS)
loss_backward_kernel(const float* __restrict__ out, int64_t ldo, const float* __restrict__ tgt, int64_t ldt,
                     int64_t n, Head hd, const double* __restrict__ stats, const float* __restrict__ w,
                     float* __restrict__ dout, int64_t lddo) {
    const int64_t i = (int64_t)blockIdx.x * THREADS + threadIdx.x;
    if (i >= n) return;
    const float* o = out + i * ldo;
    const float* t = tgt + i * ldt;
    float* d = dout + i * lddo;
    row_backward(hd, o, t, stats, w, d);
}

static int blocks_for(int64_t n) {
    const int64_t want = cdiv(n > 0 ? n : 1, THREADS);
    return (int)(want < 148 * 4 ? want : 148 * 4);
}

static int check_head(const char* who, int64_t n, int g, int64_t ldo, int64_t ldt, int color_mode,
                      int semantic_classes, int track_idx) {
    O4D_REQUIRE(n >= 0 && g >= 1, "%s: need n >= 0, g >= 1", who);
    O4D_REQUIRE(ldo >= g && ldt >= 6, "%s: bad leading dimension (target rows have 6 columns)", who);
    O4D_REQUIRE(color_mode >= O4D_COLOR_RGB && color_mode <= O4D_COLOR_BINS, "%s: unknown color_mode %d", who, color_mode);
    const int color_cols = color_mode == O4D_COLOR_RGB ? 3 : (color_mode == O4D_COLOR_HSV ? 14 : 9);
    O4D_REQUIRE(g >= 1 + color_cols, "%s: g=%d too narrow for the colour head (%d columns)", who, g, color_cols);
    O4D_REQUIRE(semantic_classes >= 0 && semantic_classes <= g, "%s: bad semantic_classes %d", who, semantic_classes);
    O4D_REQUIRE(track_idx < g, "%s: track_idx %d outside the %d output columns", who, track_idx, g);
    return 0;
}

}  // namespace lossk
}  // namespace o4d

extern "C" size_t o4d_implicit_loss_workspace_bytes(int64_t n) {
    return (size_t)o4d::lossk::blocks_for(n) * o4d::lossk::NSTAT * sizeof(double);
}

extern "C" int o4d_implicit_loss_forward_f32(const float* output, int64_t n, int g, int64_t ldo,
                                             const float* target, int64_t ldt, int color_mode,
                                             int semantic_classes, int track_idx, float* losses4_out,
                                             double* stats_out, void* ws, size_t ws_bytes, void* stream) {
    using namespace o4d;
    O4D_TRY(lossk::check_head("o4d_implicit_loss_forward_f32", n, g, ldo, ldt, color_mode, semantic_classes, track_idx));
    O4D_REQUIRE(losses4_out && stats_out, "o4d_implicit_loss_forward_f32: null output");
    O4D_REQUIRE(n == 0 || (output && target), "o4d_implicit_loss_forward_f32: null input");
    const int blocks = lossk::blocks_for(n);
    O4D_REQUIRE(ws && ws_bytes >= o4d_implicit_loss_workspace_bytes(n),
                "o4d_implicit_loss_forward_f32: workspace too small");
    lossk::Head hd{g, color_mode, semantic_classes, track_idx};
    cudaStream_t st = (cudaStream_t)stream;
    lossk::loss_partial_kernel<<<blocks, lossk::THREADS, 0, st>>>(output, ldo, target, ldt, n, hd, (double*)ws);
    O4D_LAUNCH_CHECK();
    lossk::loss_finalize_kernel<<<1, 32, 0, st>>>((const double*)ws, blocks, hd, stats_out, losses4_out);
    O4D_LAUNCH_CHECK();
    return 0;
}

extern "C" int o4d_implicit_loss_backward_f32(const float* output, int64_t n, int g, int64_t ldo,
                                              const float* target, int64_t ldt, int color_mode,
                                              int semantic_classes, int track_idx, const double* stats,
                                              const float* dlosses4, float* doutput, int64_t lddo, void* stream) {
    using namespace o4d;
    O4D_TRY(lossk::check_head("o4d_implicit_loss_backward_f32", n, g, ldo, ldt, color_mode, semantic_classes, track_idx));
    O4D_REQUIRE(lddo >= g, "o4d_implicit_loss_backward_f32: bad leading dimension");
    if (n == 0) return 0;
    O4D_REQUIRE(output && target && stats && dlosses4 && doutput, "o4d_implicit_loss_backward_f32: null pointer");
    lossk::Head hd{g, color_mode, semantic_classes, track_idx};
    lossk::loss_backward_kernel<<<(unsigned)cdiv(n, lossk::THREADS), lossk::THREADS, 0, (cudaStream_t)stream>>>(
        output, ldo, target, ldt, n, hd, stats, dlosses4, doutput, lddo);
    O4D_LAUNCH_CHECK();
    return 0;
}
