// Training-time query sampler, device pieces.
//
// Replaces  filter_air_solid_gap     utils/geometry.py:1164-1196  (sliced (slice, N, 3) difference tensors,
//                                    torch.linalg.norm, topk(1), torch.minimum over slices, boolean-mask
//                                    indexing: a dynamic shape and a host sync per call)
//           select_safely            utils/geometry.py:1095-1105  (first num_select rows, the kept rows repeated
//                                    by doubling when there are too few)
//           filter_pcl_bounds_torch  utils/geometry.py:175-188    (cuboid mask + boolean-mask indexing)
//
// One call = 1-NN distance of every candidate against the whole target cloud (the K = 1 + distance variant of
// the brute-force kNN kernel, knn.cu: no slicing is needed because no distance matrix is materialised), an
// order-preserving compaction of the candidates farther than the radius, and the gather of the kept rows --
// either the n' kept rows (boolean-mask semantics) or exactly num_select rows, row j = kept[j mod n']
// (select_safely's doubling is periodic repetition).  The kept count stays on the device: the fixed-size
// form needs no host synchronisation at all.
//
// Everything here is index / byte work on tens of thousands of rows: the distance stage is CUDA-core bound
// (n x m pair evaluations), the compaction and gather stages are latency bound (one block; a few tiles).
#include "o4d_common.cuh"

namespace o4d {
namespace smp {

constexpr int SCAN_THREADS = 1024;

// dist2 = squared 1-NN distance; the Euclidean distance is rooted once per candidate instead of once per pair
// (sqrt is monotone, so sqrt(min d2) is the minimum of the rooted distances).
struct FartherThan {
    const float* dist2;
    float radius;
    __device__ __forceinline__ bool operator()(int64_t i) const { return __fsqrt_rn(dist2[i]) > radius; }
};

struct InsideCuboid {
    const float* pcl;
    int64_t ld;
    float lo[3], hi[3];
    __device__ __forceinline__ bool operator()(int64_t i) const {
        const float* p = pcl + i * ld;
        bool ok = true;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v = p[c];
            ok = ok && (lo[c] <= v) && (v <= hi[c]);
        }
        return ok;
    }
};

// Ordered list of the row indices that satisfy `pred` and their count; one block walks the rows in tiles of
// SCAN_THREADS (ballot + warp totals), so the output order is the input order.
template <class Pred>
__global__ void __launch_bounds__(SCAN_THREADS)
compact_index_kernel(Pred pred, int64_t n, int32_t* __restrict__ kept, int32_t* __restrict__ count_out) {
    __shared__ int warp_total[SCAN_THREADS / 32];
    __shared__ int running;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) running = 0;
    __syncthreads();
    for (int64_t base = 0; base < n; base += SCAN_THREADS) {
        const int64_t i = base + threadIdx.x;
        const bool keep = (i < n) && pred(i);
        const unsigned ballot = __ballot_sync(0xffffffffu, keep);
        const int within = __popc(ballot & ((1u << lane) - 1u));
        if (lane == 0) warp_total[warp] = __popc(ballot);
        __syncthreads();
        int before = 0;
        for (int w = 0; w < warp; ++w) before += warp_total[w];
        const int off = running;
        if (keep) kept[off + before + within] = (int32_t)i;
        __syncthreads();
        if (threadIdx.x == SCAN_THREADS - 1) running = off + before + __popc(ballot);
        __syncthreads();
    }
    if (threadIdx.x == 0) *count_out = running;
}

// out[j, :] = src[kept[j mod n'], :], dist_out[j] = sqrt(dist2[kept[j mod n']]) for j < rows_out, where
// rows_out = num_select (wrap mode) or n' (mask mode).  n' = 0 in wrap mode writes zeros.
__global__ void gather_kept_kernel(const float* __restrict__ src, int64_t lds, int d,
                                   const float* __restrict__ dist, const int32_t* __restrict__ kept,
                                   const int32_t* __restrict__ count, int64_t num_select,
                                   float* __restrict__ out, int64_t ldo, float* __restrict__ dist_out,
                                   int64_t capacity) {
    const int np = *count;
    const int64_t rows_out = num_select > 0 ? num_select : (int64_t)np;
    const int cols = d + 1;  // the extra column is the distance
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < capacity * cols;
         t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = t / cols;
        const int c = (int)(t - j * cols);
        if (j >= rows_out) break;
        if (np == 0) {
            if (c < d) out[j * ldo + c] = 0.f;
            else if (dist_out) dist_out[j] = 0.f;
            continue;
        }
        const int64_t s = kept[j % np];
        if (c < d) out[j * ldo + c] = src[s * lds + c];
        else if (dist_out && dist) dist_out[j] = __fsqrt_rn(dist[s]);
    }
}

static int gather_launch(const float* src, int64_t lds, int d, const float* dist, const int32_t* kept,
                         const int32_t* count, int64_t num_select, float* out, int64_t ldo, float* dist_out,
                         int64_t capacity, cudaStream_t st) {
    if (capacity == 0) return 0;
    const int64_t work = capacity * (d + 1);
    const unsigned blocks = (unsigned)(cdiv(work, 256) < 148 * 8 ? cdiv(work, 256) : 148 * 8);
    gather_kept_kernel<<<blocks, 256, 0, st>>>(src, lds, d, dist, kept, count, num_select, out, ldo, dist_out,
                                                capacity);
    O4D_LAUNCH_CHECK();
    return 0;
}

}  // namespace smp
}  // namespace o4d

extern "C" size_t o4d_filter_workspace_bytes(int64_t n) {
    o4d::Arena a(nullptr, 0);
    a.get<float>((size_t)(n > 0 ? n : 1));
    a.get<int32_t>((size_t)(n > 0 ? n : 1));
    return a.off;
}

extern "C" int o4d_filter_air_solid_gap_f32(const float* cand, int64_t n, int d, int64_t ldc,
                                            const float* target, int64_t m, int64_t ldt, float radius,
                                            int64_t num_select, float* out, int64_t ldo, float* dist_out,
                                            int32_t* count_out, void* ws, size_t ws_bytes, void* stream) {
    using namespace o4d;
    cudaStream_t st = (cudaStream_t)stream;
    O4D_REQUIRE(n >= 0 && m >= 1 && d >= 3, "o4d_filter_air_solid_gap_f32: need n >= 0, m >= 1, d >= 3");
    O4D_REQUIRE(n < (int64_t)0x7fffffff, "o4d_filter_air_solid_gap_f32: n must be < 2^31");
    O4D_REQUIRE(ldc >= d && ldo >= d && ldt >= 3, "o4d_filter_air_solid_gap_f32: bad leading dimension");
    O4D_REQUIRE(num_select >= 0, "o4d_filter_air_solid_gap_f32: negative num_select");
    O4D_REQUIRE(count_out, "o4d_filter_air_solid_gap_f32: count_out is required");
    if (n == 0) {
        O4D_CUDA(cudaMemsetAsync(count_out, 0, sizeof(int32_t), st));
        if (num_select > 0) {
            O4D_REQUIRE(out, "o4d_filter_air_solid_gap_f32: null output");
            O4D_CUDA(cudaMemset2DAsync(out, ldo * sizeof(float), 0, d * sizeof(float), num_select, st));
            if (dist_out) O4D_CUDA(cudaMemsetAsync(dist_out, 0, num_select * sizeof(float), st));
        }
        return 0;
    }
    O4D_REQUIRE(cand && target && out, "o4d_filter_air_solid_gap_f32: null pointer");
    Arena a(ws, ws_bytes);
    float* dist = a.get<float>((size_t)n);
    int32_t* kept = a.get<int32_t>((size_t)n);
    O4D_REQUIRE(ws && a.ok, "o4d_filter_air_solid_gap_f32: workspace too small (%zu < %zu)", ws_bytes, a.off);
    // squared 1-NN distance to the whole target cloud (min over slices of the reference == global min).
    O4D_TRY(knn_launch(cand, n, ldc, target, m, ldt, 1, 0, nullptr, nullptr, dist, st));
    smp::FartherThan pred{dist, radius};
    smp::compact_index_kernel<<<1, smp::SCAN_THREADS, 0, st>>>(pred, n, kept, count_out);
    O4D_LAUNCH_CHECK();
    const int64_t capacity = num_select > 0 ? num_select : n;
    return smp::gather_launch(cand, ldc, d, dist, kept, count_out, num_select, out, ldo, dist_out, capacity, st);
}

extern "C" int o4d_filter_bounds_f32(const float* pcl, int64_t n, int d, int64_t ld, const float* lo3_host,
                                     const float* hi3_host, float* out, int64_t ldo, int32_t* count_out,
                                     void* ws, size_t ws_bytes, void* stream) {
    using namespace o4d;
    cudaStream_t st = (cudaStream_t)stream;
    O4D_REQUIRE(n >= 0 && d >= 3 && n < (int64_t)0x7fffffff, "o4d_filter_bounds_f32: need 0 <= n < 2^31, d >= 3");
    O4D_REQUIRE(ld >= d && ldo >= d, "o4d_filter_bounds_f32: bad leading dimension");
    O4D_REQUIRE(lo3_host && hi3_host && count_out, "o4d_filter_bounds_f32: null pointer");
    if (n == 0) {
        O4D_CUDA(cudaMemsetAsync(count_out, 0, sizeof(int32_t), st));
        return 0;
    }
    O4D_REQUIRE(pcl && out, "o4d_filter_bounds_f32: null pointer");
    Arena a(ws, ws_bytes);
    a.get<float>((size_t)n);  // same layout as o4d_filter_workspace_bytes
    int32_t* kept = a.get<int32_t>((size_t)n);
    O4D_REQUIRE(ws && a.ok, "o4d_filter_bounds_f32: workspace too small (%zu < %zu)", ws_bytes, a.off);
    smp::InsideCuboid pred;
    pred.pcl = pcl;
    pred.ld = ld;
    for (int c = 0; c < 3; ++c) {
        pred.lo[c] = lo3_host[c];
        pred.hi[c] = hi3_host[c];
    }
    smp::compact_index_kernel<<<1, smp::SCAN_THREADS, 0, st>>>(pred, n, kept, count_out);
    O4D_LAUNCH_CHECK();
    return smp::gather_launch(pcl, ld, d, nullptr, kept, count_out, 0, out, ldo, nullptr, n, st);
}
