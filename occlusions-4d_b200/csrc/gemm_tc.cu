// tcgen05 dense layer (placeholder until the tensor-core path lands: every shape is
// declined, so callers fall through to the CUDA-core kernel).
#include "o4d_common.cuh"

namespace o4d {
int linear_tc_launch(const float*, int64_t, int64_t, int64_t, const float*, const float*, int64_t,
                     const float*, int64_t, float*, int64_t, int, int, cudaStream_t) {
    return O4D_E_UNSUPPORTED;
}
}  // namespace o4d

extern "C" int o4d_has_tcgen05(void) { return 0; }
