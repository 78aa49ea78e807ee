// Brute-force 3-D kNN without materialising the (nq, m) distance matrix.
//
// Replaces  kNN_torch        model/point_transformer_layer.py:76-99  (square_distance + argsort)
//           my_knn_torch     utils/geometry.py:458-503               (norm + topk)
//           torch_cluster.knn at model/modules.py:142-146
//
// Layout: reference points are staged through shared memory as SoA tiles (x[], y[], z[]);
// every query is served by S adjacent lanes of a warp ("sub-lanes"), each scanning the
// tile with stride S and keeping its own ascending top-KMAX list in registers; the S lists
// are merged with warp shuffles.  Ordering is ascending (distance, index), distance =
// fp32 ((dx*dx + dy*dy) + dz*dz) with every operation rounded (no FMA contraction),
// optionally square-rooted (sqrt is monotone but merges neighbouring values, so the
// Euclidean ordering is evaluated on the rooted value to honour the index tie-break).
#include "o4d_common.cuh"
#include <stdlib.h>
#include <math_constants.h>

namespace o4d {

constexpr int KNN_THREADS = 256;
constexpr int KNN_TILE = 2048;  // reference points per shared-memory tile (24 KB)

template <int KMAX>
struct TopK {
    float d[KMAX];
    int i[KMAX];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int t = 0; t < KMAX; ++t) {
            d[t] = CUDART_INF_F;
            i[t] = 0x7fffffff;
        }
    }
    // candidates arrive in ascending index order within one thread, so a strict
    // comparison keeps the lower index ahead of an equal distance.
    __device__ __forceinline__ void push(float dist, int idx) {
        if (dist < d[KMAX - 1]) {
            d[KMAX - 1] = dist;
            i[KMAX - 1] = idx;
#pragma unroll
            for (int t = KMAX - 1; t > 0; --t) {
                if (d[t] < d[t - 1]) {
                    float td = d[t]; d[t] = d[t - 1]; d[t - 1] = td;
                    int ti = i[t]; i[t] = i[t - 1]; i[t - 1] = ti;
                }
            }
        }
    }
    __device__ __forceinline__ void pop_front() {
#pragma unroll
        for (int t = 0; t < KMAX - 1; ++t) {
            d[t] = d[t + 1];
            i[t] = i[t + 1];
        }
        d[KMAX - 1] = CUDART_INF_F;
        i[KMAX - 1] = 0x7fffffff;
    }
};

// Second result of the same scan (decoder: the K_l = 8 Euclidean neighbours of implicit.py:328 next to the K_c = 14
// squared-distance neighbours of point_transformer_layer.py:167, same query, same cloud): the k2 nearest under the key
// (sqrt(d2) correctly rounded, index).  sqrt is monotone, so they are among the k nearest under (d2, index) unless the
// rounded roots of the k2-th and the k-th coincide; inside the list only runs of equal roots need re-ordering by index
// (two different d2 can round to the same root).  Both rare cases are handled exactly (run fix-up, rescan).
struct KnnSecond {
    int k2 = 0;
    int32_t* idx32 = nullptr;     // (nq, k2)
    float* dist = nullptr;        // (nq, k2) Euclidean distances
};

template <int KMAX, int S, bool SQRT>
__global__ void __launch_bounds__(KNN_THREADS)
knn_kernel(const float* __restrict__ query, int64_t nq, int64_t ldq,
           const float* __restrict__ ref, int m, int64_t ldr, int k,
           int32_t* __restrict__ idx32, int64_t* __restrict__ idx64, float* __restrict__ dist_out, KnnSecond second) {
    __shared__ float sx[KNN_TILE], sy[KNN_TILE], sz[KNN_TILE];
    constexpr int QPB = KNN_THREADS / S;  // queries per block
    const int sub = threadIdx.x % S;
    const int64_t qi = (int64_t)blockIdx.x * QPB + threadIdx.x / S;
    const bool live = qi < nq;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (live) {
        qx = query[qi * ldq + 0];
        qy = query[qi * ldq + 1];
        qz = query[qi * ldq + 2];
    }
    TopK<KMAX> top;
    top.init();

    for (int base = 0; base < m; base += KNN_TILE) {
        const int cnt = min(KNN_TILE, m - base);
        __syncthreads();
        for (int t = threadIdx.x; t < cnt; t += KNN_THREADS) {
            const float* r = ref + (int64_t)(base + t) * ldr;
            sx[t] = r[0];
            sy[t] = r[1];
            sz[t] = r[2];
        }
        __syncthreads();
        // Two phases per block of 32 candidates of a thread (warp-uniform trip counts).  Phase 1 only FILTERS against the
        // thread's current K-th best (a bit per candidate, no divergence); phase 2 inserts the survivors, lowest index
        // first, recomputing their distance.  With unordered queries nearly every candidate used to send the whole warp
        // down the ~60-instruction sorted insertion (some lane inserts); now the warp runs it max-over-lanes(survivors)
        // times per block.  Results are unchanged: the filter bound is never tighter than the one push() applies, and the
        // insertion order within a thread is still ascending.
        auto dist = [&](int t) {
            const float dx = qx - sx[t], dy = qy - sy[t], dz = qz - sz[t];
            float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            if (SQRT) d2 = __fsqrt_rn(d2);
            return d2;
        };
        for (int blk = 0; blk * (S * 32) < cnt; ++blk) {
            const int t0 = blk * (S * 32) + sub;
            const float tau = top.d[KMAX - 1];
            unsigned pend = 0;
#pragma unroll
            for (int b = 0; b < 32; ++b) {
                const int t = t0 + S * b;
                if (live && t < cnt && dist(t) < tau) pend |= 1u << b;
            }
            while (__any_sync(0xffffffffu, pend != 0)) {
                if (pend) {
                    const int b = __ffs(pend) - 1;
                    pend &= pend - 1;
                    const int t = t0 + S * b;
                    top.push(dist(t), base + t);
                }
            }
        }
    }

    // merge the S sub-lane lists: k rounds of "global minimum of the heads".
    const unsigned full = 0xffffffffu;
    float md[KMAX];
    int mi[KMAX];
#pragma unroll
    for (int t = 0; t < KMAX; ++t) { md[t] = CUDART_INF_F; mi[t] = 0x7fffffff; }
#pragma unroll
    for (int r = 0; r < KMAX; ++r) {
        if (r >= k) break;
        float bd = top.d[0];
        int bi = top.i[0];
        if (S > 1) {
#pragma unroll
            for (int off = S / 2; off > 0; off >>= 1) {
                float od = __shfl_xor_sync(full, bd, off);
                int oi = __shfl_xor_sync(full, bi, off);
                if (od < bd || (od == bd && oi < bi)) {
                    bd = od;
                    bi = oi;
                }
            }
            if (top.i[0] == bi && top.d[0] == bd) top.pop_front();
        } else {
            top.pop_front();
        }
        // fewer than k finite distances (NaN / Inf coordinates): the slot keeps the sentinel -- emit a valid row index (0)
        // so that the gathers downstream stay inside their tables; the distance stays +Inf
        if (bi >= m) bi = 0;
        md[r] = bd;
        mi[r] = bi;
        if (live && sub == 0) {
            if (idx32) idx32[qi * k + r] = bi;
            if (idx64) idx64[qi * k + r] = (int64_t)bi;
            if (dist_out) dist_out[qi * k + r] = bd;
        }
    }
    if (!SQRT && second.k2 > 0 && live && sub == 0) {
        const int k2 = second.k2;
        float sr[KMAX];
#pragma unroll
        for (int t = 0; t < KMAX; ++t) sr[t] = t < k ? __fsqrt_rn(md[t]) : CUDART_INF_F;
        if (sr[k2 - 1] == sr[k - 1]) {
            // a point outside the list may tie with the k2-th root: exact rescan under the (root, index) key
            TopK<KMAX> t2;
            t2.init();
            for (int j = 0; j < m; ++j) {
                const float* r = ref + (int64_t)j * ldr;
                const float dx = qx - r[0], dy = qy - r[1], dz = qz - r[2];
                t2.push(__fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz))), j);
            }
#pragma unroll
            for (int t = 0; t < KMAX; ++t) { sr[t] = t2.d[t]; mi[t] = t2.i[t]; }
        } else {
            // inside the list: equal roots in ascending index (the list is ordered by (d2, index))
            bool dirty = false;
#pragma unroll
            for (int t = 0; t + 1 < KMAX; ++t) dirty = dirty || (t + 1 < k && sr[t] == sr[t + 1] && mi[t] > mi[t + 1]);
            while (dirty) {
                dirty = false;
#pragma unroll
                for (int t = 0; t + 1 < KMAX; ++t)
                    if (t + 1 < k && sr[t] == sr[t + 1] && mi[t] > mi[t + 1]) {
                        const int ti = mi[t]; mi[t] = mi[t + 1]; mi[t + 1] = ti;
                        dirty = true;
                    }
            }
        }
#pragma unroll
        for (int t = 0; t < KMAX; ++t)
            if (t < k2) {
                if (second.idx32) second.idx32[qi * k2 + t] = mi[t] < m ? mi[t] : 0;      // (sentinel: see above)
                if (second.dist) second.dist[qi * k2 + t] = sr[t];
            }
    }
}

template <int KMAX, bool SQRT>
static int knn_dispatch_s(int S, dim3 grid_unused, const float* query, int64_t nq, int64_t ldq,
                          const float* ref, int m, int64_t ldr, int k, int32_t* idx32,
                          int64_t* idx64, float* dist, cudaStream_t st, KnnSecond second = KnnSecond()) {
    (void)grid_unused;
#define O4D_KNN_CASE(SV)                                                                         \
    case SV: {                                                                                   \
        int64_t blocks = cdiv(nq, KNN_THREADS / SV);                                             \
        knn_kernel<KMAX, SV, SQRT><<<(unsigned)blocks, KNN_THREADS, 0, st>>>(                    \
            query, nq, ldq, ref, m, ldr, k, idx32, idx64, dist, second);                         \
        break;                                                                                   \
    }
    switch (S) {
        O4D_KNN_CASE(1)
        O4D_KNN_CASE(4)
        O4D_KNN_CASE(8)
        O4D_KNN_CASE(16)
        O4D_KNN_CASE(32)
        default:
            set_error("knn: bad sub-lane count %d", S);
            return O4D_E_ARG;
    }
#undef O4D_KNN_CASE
    O4D_LAUNCH_CHECK();
    return 0;
}

// Sub-lanes per query: enough threads to cover the machine (148 SMs x 2048 resident threads) a few times.
// O4D_KNN_S overrides the choice (tuning: tools/time_knn.py).
static int knn_sub_lanes(int64_t nq, int64_t m) {
    static const int forced = [] {
        const char* e = getenv("O4D_KNN_S");
        const int v = e ? atoi(e) : 0;
        return (v == 1 || v == 4 || v == 8 || v == 16 || v == 32) ? v : 0;
    }();
    if (forced) return forced;
    // measured on a B200 (tools/time_knn.py, unordered points, profiles/r2_h_knn_sub_lanes.txt): more sub-lanes pay off
    // only while the queries alone cannot fill the machine -- 14336 x 14336, K = 16: 361 us with 4 sub-lanes, 673 us with
    // 32; 4779 x 14336: 228 us with 8 (4: 307, 32: 277); 1593 x 4779: 68 us with 16; 531 x 531: 22 us with 32
    int S = 1;
    if (nq < 600000 && m >= 128) S = 4;
    if (m >= 1024) {
        if (nq < 8192) S = 8;
        if (nq < 3000) S = 16;
        if (nq < 1024) S = 32;
    }
    return S;
}

// One scan, two neighbour lists: the k nearest by squared distance (idx32) and the k2 <= k nearest by Euclidean distance
// with their distances (idx2, dist2); see KnnSecond.
int knn_dual_launch(const float* query, int64_t nq, int64_t ldq, const float* ref, int64_t m, int64_t ldr, int k,
                    int32_t* idx32, int k2, int32_t* idx2, float* dist2, cudaStream_t st) {
    // k2 < k: the k-th root bounds every point outside the list, which is what makes the in-list selection exact
    O4D_REQUIRE(k >= 9 && k <= O4D_MAX_K && k2 >= 1 && k2 < k && m >= k && m < (int64_t)0x7fffffff,
                "knn (two lists): need 1 <= k2 < k, 9 <= k <= %d <= m", O4D_MAX_K);
    if (nq == 0) return 0;
    O4D_REQUIRE(query && ref && idx32 && idx2 && dist2 && ldq >= 3 && ldr >= 3, "knn (two lists): bad argument");
    ProfScope prof(PROF_KNN, 8.0 * (double)nq * (double)m, st);
    const int S = knn_sub_lanes(nq, m);
    KnnSecond second;
    second.k2 = k2;
    second.idx32 = idx2;
    second.dist = dist2;
    dim3 g;
    return knn_dispatch_s<16, false>(S, g, query, nq, ldq, ref, (int)m, ldr, k, idx32, nullptr, nullptr, st, second);
}

int knn_launch(const float* query, int64_t nq, int64_t ldq, const float* ref, int64_t m,
               int64_t ldr, int k, int sqrt_dist, int32_t* idx32, int64_t* idx64, float* dist,
               cudaStream_t st) {
    O4D_REQUIRE(nq >= 0, "knn: negative query count");
    O4D_REQUIRE(k >= 1 && k <= O4D_MAX_K, "knn: k=%d outside [1,%d]", k, O4D_MAX_K);
    O4D_REQUIRE(m >= k && m < (int64_t)0x7fffffff, "knn: need k <= m < 2^31 (m=%lld, k=%d)",
                (long long)m, k);
    if (nq == 0) return 0;
    O4D_REQUIRE(query && ref, "knn: null input");
    O4D_REQUIRE(idx32 || idx64 || dist, "knn: no output requested");
    O4D_REQUIRE(ldq >= 3 && ldr >= 3, "knn: leading dimensions must be >= 3");
    ProfScope prof(PROF_KNN, 8.0 * (double)nq * (double)m, st);
    const int S = knn_sub_lanes(nq, m);
    dim3 g;
    if (k == 1) {  // nearest neighbour only (the sampler's air / solid gap filter): a single register pair
        return sqrt_dist ? knn_dispatch_s<1, true>(S, g, query, nq, ldq, ref, (int)m, ldr, k, idx32, idx64, dist, st)
                         : knn_dispatch_s<1, false>(S, g, query, nq, ldq, ref, (int)m, ldr, k, idx32, idx64, dist, st);
    }
    if (k <= 8) {
        return sqrt_dist ? knn_dispatch_s<8, true>(S, g, query, nq, ldq, ref, (int)m, ldr, k, idx32, idx64, dist, st)
                         : knn_dispatch_s<8, false>(S, g, query, nq, ldq, ref, (int)m, ldr, k, idx32, idx64, dist, st);
    }
    return sqrt_dist ? knn_dispatch_s<16, true>(S, g, query, nq, ldq, ref, (int)m, ldr, k, idx32, idx64, dist, st)
                     : knn_dispatch_s<16, false>(S, g, query, nq, ldq, ref, (int)m, ldr, k, idx32, idx64, dist, st);
}

}  // namespace o4d

extern "C" int o4d_knn_f32(const float* query, int64_t nq, int64_t ldq, const float* ref, int64_t m,
                           int64_t ldr, int k, int sqrt_dist, int64_t* idx_out, float* dist_out,
                           void* stream) {
    O4D_REQUIRE(nq == 0 || idx_out || dist_out, "o4d_knn_f32: no output requested");
    return o4d::knn_launch(query, nq, ldq, ref, m, ldr, k, sqrt_dist, nullptr, idx_out, dist_out,
                           (cudaStream_t)stream);
}

extern "C" size_t o4d_knn_two_lists_workspace_bytes(int64_t nq, int k, int k2) {
    if (nq < 0 || k < 1 || k2 < 1) return 0;
    return o4d::align_up((size_t)nq * k * sizeof(int32_t), 256) + o4d::align_up((size_t)nq * k2 * sizeof(int32_t), 256);
}

extern "C" int o4d_knn_two_lists_f32(const float* query, int64_t nq, int64_t ldq, const float* ref, int64_t m, int64_t ldr,
                                     int k, int k2, int64_t* idx_out, int64_t* idx2_out, float* dist2_out, void* workspace,
                                     size_t workspace_bytes, void* stream) {
    using namespace o4d;
    O4D_REQUIRE(nq == 0 || (idx_out && idx2_out && dist2_out && workspace), "knn (two lists): null pointer");
    if (nq == 0) return 0;
    Arena a(workspace, workspace_bytes);
    int32_t* i1 = a.get<int32_t>((size_t)nq * k);
    int32_t* i2 = a.get<int32_t>((size_t)nq * k2);
    if (!a.ok) {
        set_error("knn (two lists): workspace too small (%zu < %zu)", workspace_bytes, a.off);
        return O4D_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    O4D_TRY(knn_dual_launch(query, nq, ldq, ref, m, ldr, k, i1, k2, i2, dist2_out, st));
    O4D_TRY(widen_idx_launch(i1, nq * k, idx_out, st));
    return widen_idx_launch(i2, nq * k2, idx2_out, st);
}
