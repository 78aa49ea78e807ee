// TEST INFRASTRUCTURE: compiles the per-row arithmetic of the loss-head kernels (csrc/loss_core.cuh, the very
// source the CUDA kernels include) for the host, so tests/test_host_core.py can check it against the reference's
// golden losses and gradients without a GPU.  Serial loops stand in for the grid; the reduction order differs
// from the device (fp64 sums: far below the test tolerance).
#include "../../occlusions-4d_b200/csrc/loss_core.cuh"

using namespace o4d::lossk;

extern "C" void host_loss_forward(const float* out, long n, int g, const float* tgt, int color_mode,
                                  int semantic_classes, int track_idx, float* losses4, double* stats) {
    Head hd{g, color_mode, semantic_classes, track_idx};
    for (int s = 0; s < NSTAT; ++s) stats[s] = 0.0;
    for (long i = 0; i < n; ++i) row_accumulate(hd, out + i * g, tgt + i * 6, stats);
    finalize_losses(hd, stats, losses4);
}

extern "C" void host_loss_backward(const float* out, long n, int g, const float* tgt, int color_mode,
                                   int semantic_classes, int track_idx, const double* stats, const float* w,
                                   float* dout) {
    Head hd{g, color_mode, semantic_classes, track_idx};
    for (long i = 0; i < n; ++i) row_backward(hd, out + i * g, tgt + i * 6, stats, w, dout + i * g);
}
