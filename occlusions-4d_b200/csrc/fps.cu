// Farthest point sampling, one cloud per launch.
//
// Replaces torch_cluster.fps + torch.sort at model/modules.py:133-135 (un-vendored CUDA
// extension; algorithm restated from its published semantics, see oracle/cluster_ops.py).
//
// The selection loop is a serial chain of n_out dependent arg-max steps, so the whole
// cloud lives on ONE SM: coordinates and running min-distances sit in registers
// (P points per thread), the current centre is broadcast through shared memory and the
// arg-max is a shuffle + one shared-memory hop.  Two block barriers per step.
// Tie rule: first (lowest-index) maximum, like torch.argmax / the upstream kernel.
// The sorted index list the caller wants is produced in the same launch by a
// counting pass (indices may repeat when the cloud holds duplicates, e.g. zero padding).
#include "o4d_common.cuh"
#include <math_constants.h>
#include <stdlib.h>

namespace o4d {

constexpr int FPS_THREADS = 1024;

__device__ __forceinline__ float sqdist3(float ax, float ay, float az, float bx, float by, float bz) {
    float dx = ax - bx, dy = ay - by, dz = az - bz;
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

__device__ __forceinline__ void argmax_combine(float& v, int& i, float ov, int oi) {
    if (ov > v || (ov == v && oi < i)) {
        v = ov;
        i = oi;
    }
}

// P = points per thread held in registers (n <= P * FPS_THREADS).
template <int P>
__global__ void __launch_bounds__(FPS_THREADS, 1)
fps_kernel(const float* __restrict__ xyz, int n, int64_t ld, int n_out, int start,
           int32_t* __restrict__ counts,  // (n) zero-initialised scratch
           int32_t* __restrict__ sorted32, int64_t* __restrict__ sorted64,
           int64_t* __restrict__ order64) {
    __shared__ float s_val[32];
    __shared__ int s_idx[32];
    __shared__ float s_c[3];
    __shared__ int s_cur;
    __shared__ int s_scan[FPS_THREADS / 32];
    __shared__ int s_running;

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    float px[P], py[P], pz[P], md[P];
#pragma unroll
    for (int p = 0; p < P; ++p) {
        int j = tid + p * FPS_THREADS;
        if (j < n) {
            px[p] = xyz[(int64_t)j * ld + 0];
            py[p] = xyz[(int64_t)j * ld + 1];
            pz[p] = xyz[(int64_t)j * ld + 2];
            md[p] = CUDART_INF_F;
        } else {
            px[p] = py[p] = pz[p] = 0.f;
            md[p] = -1.f;  // never wins the arg-max (real distances are >= 0)
        }
    }
    if (tid == 0) {
        s_cur = start;
        s_c[0] = xyz[(int64_t)start * ld + 0];
        s_c[1] = xyz[(int64_t)start * ld + 1];
        s_c[2] = xyz[(int64_t)start * ld + 2];
    }
    __syncthreads();

    for (int it = 0; it < n_out; ++it) {
        const int cur = s_cur;
        const float cx = s_c[0], cy = s_c[1], cz = s_c[2];
        if (tid == 0) {
            if (order64) order64[it] = cur;
            atomicAdd(&counts[cur], 1);
        }
        float bv = -1.f;  // padded slots carry -1 and can never win (strict >)
        int bi = 0x7fffffff;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            float d = sqdist3(px[p], py[p], pz[p], cx, cy, cz);
            float mm = fminf(md[p], d);
            // padded slots keep md = -1 (fminf(-1, d) = -1).
            md[p] = mm;
            if (mm > bv) {  // ascending index within the thread: strict keeps the first
                bv = mm;
                bi = tid + p * FPS_THREADS;
            }
        }
        // warp arg-max in two REDUX instructions (running-min distances are >= 0, padding -1: their bit patterns order like
        // signed integers; ties resolve to the lowest index)
        {
            const int vb = __float_as_int(bv);
            const int vmax = __reduce_max_sync(0xffffffffu, vb);
            bi = __reduce_min_sync(0xffffffffu, vb == vmax ? bi : 0x7fffffff);
            bv = __int_as_float(vmax);
        }
        if (lane == 0) {
            s_val[warp] = bv;
            s_idx[warp] = bi;
        }
        __syncthreads();
        if (warp == 0) {
            float v = s_val[lane];
            int i = s_idx[lane];
            {
                const int vb = __float_as_int(v);
                const int vmax = __reduce_max_sync(0xffffffffu, vb);
                i = __reduce_min_sync(0xffffffffu, vb == vmax ? i : 0x7fffffff);
                v = __int_as_float(vmax);
            }
            if (lane == 0) {
                if (i >= n) i = 0;              // all-NaN cloud: the sentinel index must not reach the loads below
                s_cur = i;
                s_c[0] = xyz[(int64_t)i * ld + 0];
                s_c[1] = xyz[(int64_t)i * ld + 1];
                s_c[2] = xyz[(int64_t)i * ld + 2];
            }
        }
        __syncthreads();
    }

    // counting pass -> ascending index list (with multiplicity).
    __threadfence_block();
    if (tid == 0) s_running = 0;
    __syncthreads();
    for (int base = 0; base < n; base += FPS_THREADS) {
        int j = base + tid;
        int c = (j < n) ? counts[j] : 0;
        // block exclusive scan of c
        int incl = c;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += t;
        }
        if (lane == 31) s_scan[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = s_scan[lane];
            int wi = w;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, wi, off);
                if (lane >= off) wi += t;
            }
            s_scan[lane] = wi - w;  // exclusive prefix of warp totals
        }
        __syncthreads();
        int pos = s_running + s_scan[warp] + incl - c;
        for (int r = 0; r < c; ++r) {
            if (sorted32) sorted32[pos + r] = j;
            if (sorted64) sorted64[pos + r] = j;
        }
        __syncthreads();
        if (tid == FPS_THREADS - 1) s_running = pos + c;
        __syncthreads();
    }
}

// Fallback for clouds too large for the register-resident kernel: distances in global
// scratch (L2 resident), same tie rule.  Slow path; the reference caps n at 65536
// (args.py:106) and the released configurations use 14336.
__global__ void __launch_bounds__(FPS_THREADS, 1)
fps_big_kernel(const float* __restrict__ xyz, int n, int64_t ld, int n_out, int start,
               float* __restrict__ mind, int32_t* __restrict__ counts) {
    __shared__ float s_val[32];
    __shared__ int s_idx[32];
    __shared__ int s_cur;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int j = tid; j < n; j += FPS_THREADS) mind[j] = CUDART_INF_F;
    if (tid == 0) s_cur = start;
    __syncthreads();
    for (int it = 0; it < n_out; ++it) {
        const int cur = s_cur;
        const float cx = xyz[(int64_t)cur * ld], cy = xyz[(int64_t)cur * ld + 1], cz = xyz[(int64_t)cur * ld + 2];
        if (tid == 0) atomicAdd(&counts[cur], 1);
        float bv = -2.f;
        int bi = 0x7fffffff;
        for (int j = tid; j < n; j += FPS_THREADS) {
            float d = sqdist3(xyz[(int64_t)j * ld], xyz[(int64_t)j * ld + 1], xyz[(int64_t)j * ld + 2], cx, cy, cz);
            float mm = fminf(mind[j], d);
            mind[j] = mm;
            if (mm > bv) { bv = mm; bi = j; }
        }
        // warp arg-max in two REDUX instructions (running-min distances are >= 0, padding -1: their bit patterns order like
        // signed integers; ties resolve to the lowest index)
        {
            const int vb = __float_as_int(bv);
            const int vmax = __reduce_max_sync(0xffffffffu, vb);
            bi = __reduce_min_sync(0xffffffffu, vb == vmax ? bi : 0x7fffffff);
            bv = __int_as_float(vmax);
        }
        if (lane == 0) { s_val[warp] = bv; s_idx[warp] = bi; }
        __syncthreads();
        if (warp == 0) {
            float v = s_val[lane];
            int i = s_idx[lane];
            {
                const int vb = __float_as_int(v);
                const int vmax = __reduce_max_sync(0xffffffffu, vb);
                i = __reduce_min_sync(0xffffffffu, vb == vmax ? i : 0x7fffffff);
                v = __int_as_float(vmax);
            }
            if (lane == 0) s_cur = i < n ? i : 0;      // (all-NaN cloud: stay inside the arrays)
        }
        __syncthreads();
    }
}

// single-block exclusive scan of counts -> sorted index list (used by the fallback).
__global__ void __launch_bounds__(FPS_THREADS, 1)
fps_emit_sorted_kernel(const int32_t* __restrict__ counts, int n, int32_t* __restrict__ sorted32,
                       int64_t* __restrict__ sorted64) {
    __shared__ int s_scan[FPS_THREADS / 32];
    __shared__ int s_running;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_running = 0;
    __syncthreads();
    for (int base = 0; base < n; base += FPS_THREADS) {
        int j = base + tid;
        int c = (j < n) ? counts[j] : 0;
        int incl = c;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += t;
        }
        if (lane == 31) s_scan[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = s_scan[lane], wi = w;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, wi, off);
                if (lane >= off) wi += t;
            }
            s_scan[lane] = wi - w;
        }
        __syncthreads();
        int pos = s_running + s_scan[warp] + incl - c;
        for (int r = 0; r < c; ++r) {
            if (sorted32) sorted32[pos + r] = j;
            if (sorted64) sorted64[pos + r] = j;
        }
        __syncthreads();
        if (tid == FPS_THREADS - 1) s_running = pos + c;
        __syncthreads();
    }
}

// 8-CTA cluster version (fps_cluster.cu); O4D_E_UNSUPPORTED when the cloud does not fit it.
int fps_cluster_launch(const float* xyz, int64_t n, int64_t ld, int64_t n_out, int64_t start, int32_t* counts,
                       int64_t* order64, cudaStream_t st);

static bool fps_cluster_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("O4D_FPS_CLUSTER");      // 0 = single-SM kernel only (A/B timing)
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

static size_t fps_ws_bytes(int64_t n) {
    return align_up((size_t)n * sizeof(int32_t), 256) + align_up((size_t)n * sizeof(float), 256);
}

int fps_launch(const float* xyz, int64_t n, int64_t ld, int64_t n_out, int64_t start,
               int32_t* sorted32, int64_t* sorted64, int64_t* order64, void* ws, size_t ws_bytes,
               cudaStream_t st) {
    O4D_REQUIRE(xyz && (sorted32 || sorted64), "fps: null pointer");
    O4D_REQUIRE(n >= 1 && n < (1 << 30) && ld >= 3, "fps: bad cloud size n=%lld ld=%lld", (long long)n,
                (long long)ld);
    O4D_REQUIRE(n_out >= 1 && n_out <= n, "fps: need 1 <= n_out <= n (n_out=%lld, n=%lld)",
                (long long)n_out, (long long)n);
    O4D_REQUIRE(start >= 0 && start < n, "fps: start index outside the cloud");
    if (ws_bytes < fps_ws_bytes(n) || ws == nullptr) {
        set_error("fps: workspace too small (%zu < %zu)", ws_bytes, fps_ws_bytes(n));
        return O4D_E_WORKSPACE;
    }
    ProfScope prof(PROF_FPS, 10.0 * (double)n * (double)n_out, st);
    int32_t* counts = (int32_t*)ws;
    float* mind = (float*)((char*)ws + align_up((size_t)n * sizeof(int32_t), 256));
    O4D_CUDA(cudaMemsetAsync(counts, 0, (size_t)n * sizeof(int32_t), st));
    if (fps_cluster_enabled()) {
        const int rc = fps_cluster_launch(xyz, n, ld, n_out, start, counts, order64, st);
        if (rc == 0) {
            fps_emit_sorted_kernel<<<1, FPS_THREADS, 0, st>>>(counts, (int)n, sorted32, sorted64);
            O4D_LAUNCH_CHECK();
            return 0;
        }
        if (rc != O4D_E_UNSUPPORTED) return rc;
    }
    const int ppt = (int)cdiv(n, FPS_THREADS);
#define O4D_FPS_CASE(PV)                                                                        \
    fps_kernel<PV><<<1, FPS_THREADS, 0, st>>>(xyz, (int)n, ld, (int)n_out, (int)start, counts,  \
                                              sorted32, sorted64, order64)
    if (ppt <= 1) O4D_FPS_CASE(1);
    else if (ppt <= 2) O4D_FPS_CASE(2);
    else if (ppt <= 4) O4D_FPS_CASE(4);
    else if (ppt <= 8) O4D_FPS_CASE(8);
    else if (ppt <= 16) O4D_FPS_CASE(16);
    else {
        O4D_REQUIRE(order64 == nullptr, "fps: selection order output unsupported for n > %d",
                    16 * FPS_THREADS);
        fps_big_kernel<<<1, FPS_THREADS, 0, st>>>(xyz, (int)n, ld, (int)n_out, (int)start, mind, counts);
        O4D_LAUNCH_CHECK();
        fps_emit_sorted_kernel<<<1, FPS_THREADS, 0, st>>>(counts, (int)n, sorted32, sorted64);
    }
#undef O4D_FPS_CASE
    O4D_LAUNCH_CHECK();
    return 0;
}

}  // namespace o4d

extern "C" size_t o4d_fps_workspace_bytes(int64_t n, int64_t n_out) {
    (void)n_out;
    return n > 0 ? o4d::fps_ws_bytes(n) : 0;
}

extern "C" int o4d_fps_f32(const float* xyz, int64_t n, int64_t ld, int64_t n_out, int64_t start_idx,
                           int64_t* idx_sorted_out, int64_t* order_out, void* workspace,
                           size_t workspace_bytes, void* stream) {
    return o4d::fps_launch(xyz, n, ld, n_out, start_idx, nullptr, idx_sorted_out, order_out, workspace,
                           workspace_bytes, (cudaStream_t)stream);
}
