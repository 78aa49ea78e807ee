// Farthest point sampling on a thread-block CLUSTER (8 CTAs = 8 SMs of one GPC).
//
// The selection loop of FPS is a serial chain of dependent arg-max steps (6,903 for the
// 14336 -> 4779 -> 1593 -> 531 pyramid); fps.cu runs it on one SM, where every pick costs
// ~3.6 k cycles (1.84 us at N = 14336: 16 points per thread of distance update + two block
// barriers + a global load of the winner's coordinates).  Here the cloud is spread over the
// 8 CTAs of a cluster: each thread owns <= P points (P = 4 at N = 14336), every CTA keeps the
// WHOLE cloud's coordinates in its own shared memory (172 KB, SoA) so the next centre is a
// local broadcast read.  Per pick: warp shuffles + one block barrier give the CTA's candidate,
// lanes 0-7 of warp 0 store it into the shared memory of the 8 CTAs (one 8-byte
// distributed-shared-memory store each), one hardware cluster barrier (arrive.release /
// wait.acquire) publishes them, and every thread reduces the 8 candidates itself (no second
// block barrier).  Candidate slots are double buffered by pick parity.
// Round 2: the cluster barrier gave way to tagged candidate words that the receivers poll (0.99 -> 0.67 us per pick), then
// the block barrier and the second-level reduce went too: every WARP publishes its candidate to all CTAs and every warp
// reduces all of them (fps_cluster_warp_kernel, 0.57 us per pick at N = 14336, 0.47 at N = 4779; profiles/r2_h_fps_timing.txt).
// Measured per pick at N = 14336 on a B200 in round 1: single SM 1.85 us; barrier version 0.99 us; replacing
// the cluster barrier by remote mbarrier arrives was SLOWER (8 arrivals per CTA: 1.17 us; one per
// warp, 128 per CTA and no block barrier: 1.36 us).  Tie rule unchanged: first (lowest-index)
// maximum.
#include "o4d_common.cuh"
#include <math_constants.h>
#include <stdlib.h>

namespace o4d {
namespace fc {

constexpr int THREADS = 512;
constexpr int WARPS = THREADS / 32;

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
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t map_remote(const void* local_smem_ptr, uint32_t rank) {
    const uint32_t la = (uint32_t)__cvta_generic_to_shared(local_smem_ptr);
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(rank));
    return ra;
}
__device__ __forceinline__ void st_remote_u64_addr(uint32_t ra, uint64_t v) {
    asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(ra), "l"(v) : "memory");
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint64_t ld_volatile_shared_u64(const void* p) {
    uint64_t v;
    asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
    return v;
}
// Candidate word: [63:32] running-min distance bits, [31:18] pick tag ((pick + 1) mod 2^14), [17:0] point index.
constexpr uint32_t IDX_BITS = 18, IDX_MASK = (1u << IDX_BITS) - 1u, TAG_MASK = 0x3fffu;

// CTAS = cluster size (launch attribute), P = points per thread; points are dealt out round-robin over all
// CTAS * THREADS threads of the cluster.
// POLL (round 2, default): no cluster barrier inside the pick loop.  Every candidate word carries the tag of its pick;
// a thread spins on the CTAS slots of its own shared memory until all of them show the current tag.  The exchange then
// costs one remote 8-byte store plus its flight time instead of a hardware cluster barrier of 8 x 512 threads
// (arrive.release / wait.acquire), which was most of the 0.99 us per pick.  Slots are double buffered by pick parity:
// a CTA can write slot parity p of pick i + 2 only after it has seen every candidate of pick i + 1, which every CTA
// published after its last read of pick i -- so no slot is overwritten while still being read, and a stale word (pick
// i - 2, same parity) never matches the tag.
template <int CTAS, int P, bool POLL>
__global__ void __launch_bounds__(THREADS, 1)
fps_cluster_kernel(const float* __restrict__ xyz, int n, int64_t ld, int n_out, int start,
                   int32_t* __restrict__ counts,   // (n) zero-initialised
                   int64_t* __restrict__ order64) {
    extern __shared__ float s_xyz[];               // sx[n], sy[n], sz[n]
    constexpr int STRIDE = CTAS * THREADS;
    constexpr int NCAND = CTAS;                    // one candidate per CTA of the cluster
    __shared__ float s_val[WARPS];
    __shared__ int s_idx[WARPS];
    __shared__ __align__(8) uint64_t s_cand[2][NCAND];  // [pick parity][source CTA] = (value bits << 32) | index
    float* sx = s_xyz;
    float* sy = s_xyz + n;
    float* sz = s_xyz + 2 * n;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t rank = cluster_ctarank();

    if (tid < 2 * NCAND) (&s_cand[0][0])[tid] = 0ull;
    for (int j = tid; j < n; j += THREADS) {
        sx[j] = xyz[(int64_t)j * ld + 0];
        sy[j] = xyz[(int64_t)j * ld + 1];
        sz[j] = xyz[(int64_t)j * ld + 2];
    }
    __syncthreads();
    float px[P], py[P], pz[P], md[P];
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const int j = p * STRIDE + (int)rank * THREADS + tid;     // ascending in p: strict '>' keeps the first maximum
        if (j < n) {
            px[p] = sx[j]; py[p] = sy[j]; pz[p] = sz[j];
            md[p] = CUDART_INF_F;
        } else {
            px[p] = py[p] = pz[p] = 0.f;
            md[p] = -1.f;                                          // never wins (real distances are >= 0)
        }
    }
    // lane l < 8 of warp 0 hands this CTA's candidate to CTA l
    uint32_t r_cand[2] = {0, 0};
    if (warp == 0 && lane < CTAS) {
#pragma unroll
        for (int b = 0; b < 2; ++b) r_cand[b] = map_remote(&s_cand[b][rank], (uint32_t)lane);
    }
    // all CTAs of the cluster are running before any remote access
    cluster_arrive();
    cluster_wait();

    int cur = start;
    for (int it = 0; it < n_out; ++it) {
        const float cx = sx[cur], cy = sy[cur], cz = sz[cur];
        if (rank == 0 && tid == 0) {
            if (order64) order64[it] = cur;
            atomicAdd(&counts[cur], 1);
        }
        float bv = -1.f;
        int bi = 0x7fffffff;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const float dd = sqdist3(px[p], py[p], pz[p], cx, cy, cz);
            const float mm = fminf(md[p], dd);                     // padded slots keep -1
            md[p] = mm;
            if (mm > bv) {
                bv = mm;
                bi = p * STRIDE + (int)rank * THREADS + tid;
            }
        }
        // warp arg-max in two REDUX instructions instead of five shuffle / compare rounds: running-min distances are >= 0
        // (padding -1), so their bit patterns order like signed integers; ties resolve to the lowest index
        {
            const int vb = __float_as_int(bv);
            const int vmax = __reduce_max_sync(0xffffffffu, vb);
            bi = __reduce_min_sync(0xffffffffu, vb == vmax ? bi : 0x7fffffff);
            bv = __int_as_float(vmax);
        }
        const int par = it & 1;
        if (lane == 0) {
            s_val[warp] = bv;
            s_idx[warp] = bi;
        }
        __syncthreads();
        if (warp == 0) {
            float v = lane < WARPS ? s_val[lane] : -1.f;      // (-1: the padding value, orders below every real distance)
            int i = lane < WARPS ? s_idx[lane] : 0x7fffffff;
            {
                const int vb = __float_as_int(v);
                const int vmax = __reduce_max_sync(0xffffffffu, vb);
                i = __reduce_min_sync(0xffffffffu, vb == vmax ? i : 0x7fffffff);
                v = __int_as_float(vmax);
            }
            // (a CTA without points publishes value -1 and the sentinel index: masked, so that it cannot spill into the tag)
            const uint32_t tagged = POLL ? ((((uint32_t)it + 1u) & TAG_MASK) << IDX_BITS) | ((uint32_t)i & IDX_MASK) : (uint32_t)i;
            if (lane < CTAS) st_remote_u64_addr(r_cand[par], ((uint64_t)__float_as_uint(v) << 32) | tagged);
        }
        float gv = -2.f;
        int gi = 0x7fffffff;
        if (POLL) {
            const uint32_t tag = ((uint32_t)it + 1u) & TAG_MASK;
            uint64_t e[NCAND];
            bool ok;
            long long t0 = 0;
            do {
                ok = true;
#pragma unroll
                for (int c = 0; c < NCAND; ++c) {
                    e[c] = ld_volatile_shared_u64(&s_cand[par][c]);
                    ok = ok && ((((uint32_t)e[c]) >> IDX_BITS) == tag);
                }
                if (!ok) {                                   // bounded: a protocol bug fails the launch instead of hanging
                    if (t0 == 0) t0 = clock64();
                    else if (clock64() - t0 > 2000000000LL) __trap();
                }
            } while (!ok);
#pragma unroll
            for (int c = 0; c < NCAND; ++c)
                argmax_combine(gv, gi, __uint_as_float((uint32_t)(e[c] >> 32)), (int)((uint32_t)e[c] & IDX_MASK));
        } else {
            cluster_arrive();       // release: the remote stores above are visible after the matching wait
            cluster_wait();
#pragma unroll
            for (int c = 0; c < NCAND; ++c) {
                const uint64_t e = s_cand[par][c];
                argmax_combine(gv, gi, __uint_as_float((uint32_t)(e >> 32)), (int)(uint32_t)(e & 0xffffffffu));
            }
        }
        cur = gi < n ? gi : 0;      // all-NaN clouds keep the sentinel: stay inside the coordinate arrays
        // s_val / s_idx are rewritten only after the next distance update and s_cand[parity] only two
        // picks later (see POLL above / separated from these reads by a cluster barrier).
    }
    // no CTA may exit while a peer can still store into its shared memory
    cluster_arrive();
    cluster_wait();
}


// WARP variant: no block barrier and no second-level reduce inside the pick loop.  Every warp publishes its own
// candidate (tagged as above) to all CTAS shared memories -- lanes 0..CTAS-1 store one 8-byte word each -- and every warp
// polls all CTAS * WARPS slots itself (lane l reads slots l, l + 32, ...), then a REDUX arg-max over the lanes gives the
// pick.  The slot-reuse argument of POLL holds at warp granularity: a warp can write parity p of pick i + 2 only after it
// has seen the pick-(i + 1) candidate of EVERY warp of the cluster, and a warp publishes pick i + 1 only after its own
// last read of pick i.
template <int CTAS, int P, int T>
__global__ void __launch_bounds__(T, 1)
fps_cluster_warp_kernel(const float* __restrict__ xyz, int n, int64_t ld, int n_out, int start,
                        int32_t* __restrict__ counts, int64_t* __restrict__ order64) {
    extern __shared__ float s_xyz[];               // sx[n], sy[n], sz[n]
    constexpr int STRIDE = CTAS * T;
    constexpr int NW = T / 32;                     // warps per CTA
    constexpr int NSLOT = CTAS * NW;            // one candidate per warp of the cluster
    constexpr int SPL = (NSLOT + 31) / 32;         // slots per lane
    __shared__ __align__(8) uint64_t s_cand[2][NSLOT];
    float* sx = s_xyz;
    float* sy = s_xyz + n;
    float* sz = s_xyz + 2 * n;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t rank = cluster_ctarank();

    for (int j = tid; j < 2 * NSLOT; j += T) (&s_cand[0][0])[j] = 0ull;
    for (int j = tid; j < n; j += T) {
        sx[j] = xyz[(int64_t)j * ld + 0];
        sy[j] = xyz[(int64_t)j * ld + 1];
        sz[j] = xyz[(int64_t)j * ld + 2];
    }
    __syncthreads();
    float px[P], py[P], pz[P], md[P];
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const int j = p * STRIDE + (int)rank * T + tid;     // ascending in p: strict '>' keeps the first maximum
        if (j < n) {
            px[p] = sx[j]; py[p] = sy[j]; pz[p] = sz[j];
            md[p] = CUDART_INF_F;
        } else {
            px[p] = py[p] = pz[p] = 0.f;
            md[p] = -1.f;                                          // never wins (real distances are >= 0)
        }
    }
    // lane l < CTAS hands this warp's candidate to CTA l
    uint32_t r_cand[2] = {0, 0};
    if (lane < CTAS) {
#pragma unroll
        for (int b = 0; b < 2; ++b) r_cand[b] = map_remote(&s_cand[b][rank * NW + warp], (uint32_t)lane);
    }
    cluster_arrive();       // all CTAs are running (and have cleared their slots) before any remote access
    cluster_wait();

    int cur = start;
    for (int it = 0; it < n_out; ++it) {
        const float cx = sx[cur], cy = sy[cur], cz = sz[cur];
        if (rank == 0 && tid == 0) {
            if (order64) order64[it] = cur;
            atomicAdd(&counts[cur], 1);
        }
        float bv = -1.f;
        int bi = 0x7fffffff;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const float dd = sqdist3(px[p], py[p], pz[p], cx, cy, cz);
            const float mm = fminf(md[p], dd);                     // padded slots keep -1
            md[p] = mm;
            if (mm > bv) {
                bv = mm;
                bi = p * STRIDE + (int)rank * T + tid;
            }
        }
        {
            const int vb = __float_as_int(bv);
            const int vmax = __reduce_max_sync(0xffffffffu, vb);
            bi = __reduce_min_sync(0xffffffffu, vb == vmax ? bi : 0x7fffffff);
            bv = __int_as_float(vmax);
        }
        const int par = it & 1;
        const uint32_t tag = ((uint32_t)it + 1u) & TAG_MASK;
        if (lane < CTAS)
            st_remote_u64_addr(r_cand[par], ((uint64_t)__float_as_uint(bv) << 32) | (tag << IDX_BITS) | ((uint32_t)bi & IDX_MASK));
        uint64_t e[SPL];
        long long t0 = 0;
        for (;;) {
            bool ok = true;
#pragma unroll
            for (int c = 0; c < SPL; ++c) {
                const int slot = lane + 32 * c;
                if (slot < NSLOT) {
                    e[c] = ld_volatile_shared_u64(&s_cand[par][slot]);
                    ok = ok && ((((uint32_t)e[c]) >> IDX_BITS) == tag);
                } else {
                    e[c] = ((uint64_t)0xBF800000u << 32) | IDX_MASK;      // value -1, sentinel index: never wins
                }
            }
            if (__all_sync(0xffffffffu, ok)) break;
            if (t0 == 0) t0 = clock64();                         // bounded: a protocol bug fails the launch instead of hanging
            else if (clock64() - t0 > 2000000000LL) __trap();
        }
        float gv = -2.f;
        int gi = 0x7fffffff;
#pragma unroll
        for (int c = 0; c < SPL; ++c)
            argmax_combine(gv, gi, __uint_as_float((uint32_t)(e[c] >> 32)), (int)((uint32_t)e[c] & IDX_MASK));
        {
            const int vb = __float_as_int(fmaxf(gv, -1.f));      // (-2 would order above -1 as a signed integer)
            const int vmax = __reduce_max_sync(0xffffffffu, vb);
            gi = __reduce_min_sync(0xffffffffu, vb == vmax ? gi : 0x7fffffff);
        }
        cur = gi < n ? gi : 0;      // all-NaN clouds keep the sentinel: stay inside the coordinate arrays
    }
    // no CTA may exit while a peer can still store into its shared memory
    cluster_arrive();
    cluster_wait();
}

}  // namespace fc

// Returns O4D_E_UNSUPPORTED when the cloud does not fit this kernel (the caller falls back).
template <int CTAS, int P, bool POLL>
static int fps_cluster_launch_t(const float* xyz, int64_t n, int64_t ld, int64_t n_out, int64_t start, int32_t* counts,
                                int64_t* order64, size_t smem, cudaStream_t st) {
    O4D_SMEM_ATTR((fc::fps_cluster_kernel<CTAS, P, POLL>), 200 * 1024);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CTAS, 1, 1);
    cfg.blockDim = dim3(fc::THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CTAS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    O4D_CUDA(cudaLaunchKernelEx(&cfg, fc::fps_cluster_kernel<CTAS, P, POLL>, xyz, (int)n, ld, (int)n_out, (int)start, counts, order64));
    count_launch();
    return 0;
}

template <int CTAS, int P, int T>
static int fps_cluster_warp_launch_t(const float* xyz, int64_t n, int64_t ld, int64_t n_out, int64_t start, int32_t* counts,
                                     int64_t* order64, size_t smem, cudaStream_t st) {
    O4D_SMEM_ATTR((fc::fps_cluster_warp_kernel<CTAS, P, T>), 200 * 1024);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CTAS, 1, 1);
    cfg.blockDim = dim3(T, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CTAS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    O4D_CUDA(cudaLaunchKernelEx(&cfg, fc::fps_cluster_warp_kernel<CTAS, P, T>, xyz, (int)n, ld, (int)n_out, (int)start, counts, order64));
    count_launch();
    return 0;
}

int fps_cluster_launch(const float* xyz, int64_t n, int64_t ld, int64_t n_out, int64_t start, int32_t* counts,
                       int64_t* order64, cudaStream_t st) {
    const size_t smem = (size_t)3 * n * sizeof(float);
    static const int64_t min_n = getenv("O4D_FPS_CLUSTER_MIN") ? atoll(getenv("O4D_FPS_CLUSTER_MIN")) : 1024;   // below: the single-SM kernel
    if (n <= min_n || smem > 200 * 1024) return O4D_E_UNSUPPORTED;
    static int ctas_env = -1;
    if (ctas_env < 0) {
        const char* e = getenv("O4D_FPS_CTAS");          // force a cluster size: 2, 4 or 8
        ctas_env = e ? atoi(e) : 0;
        if (ctas_env != 2 && ctas_env != 4 && ctas_env != 8) ctas_env = 0;
    }
    // measured per pick with the polled exchange (profiles/r2_g_fps_timing.txt): N = 14336: 0.89 us on 8 CTAs, 0.87 on 4,
    // 0.99 on 2; N = 4779: 0.85 / 0.76 / 0.79 -> 8 CTAs above 8192 points, 4 below
    const int ctas = ctas_env ? ctas_env : (n > 8192 ? 8 : 4);
    static int poll = -1;
    if (poll < 0) {
        // default: per-warp candidates, no block barrier (fps_cluster_warp_kernel).  A/B timing: "cta" = one candidate per
        // CTA behind a block barrier, polled; "barrier" = the same with a hardware cluster barrier per pick
        const char* e = getenv("O4D_FPS_SYNC");
        poll = (e && e[0] == 'b') ? 0 : (e && e[0] == 'c') ? 1 : 2;
    }
    if (n > (int64_t)fc::IDX_MASK) poll = 0;             // the tagged candidate word holds 18 index bits
    static int threads_env = -1;
    if (threads_env < 0) {
        // threads per CTA of the per-warp exchange: every pick waits for the slowest warp of the cluster, so FEWER warps
        // are faster until the distance update of P points per thread takes over (profiles/r2_h_fps_timing.txt:
        // 512 threads 0.57 us per pick at N = 14336, 256: 0.44, 128: 0.39, 64: 0.50; a 16-CTA (non-portable) cluster of
        // 128 threads: 0.42 -- more candidate words to wait for than distance work saved).  A/B timing: 256, 512
        const char* e = getenv("O4D_FPS_THREADS");
        threads_env = e ? atoi(e) : 128;
    }
    if (poll == 2 && threads_env == 256) {
        const int ppt2 = (int)cdiv(n, (int64_t)ctas * 256);
#define O4D_FW(C, PV) return fps_cluster_warp_launch_t<C, PV, 256>(xyz, n, ld, n_out, start, counts, order64, smem, st)
        if (ctas == 8) {
            if (ppt2 <= 2) O4D_FW(8, 2);
            if (ppt2 <= 4) O4D_FW(8, 4);
            if (ppt2 <= 8) O4D_FW(8, 8);
            if (ppt2 <= 10) O4D_FW(8, 10);
        } else if (ctas == 4) {
            if (ppt2 <= 4) O4D_FW(4, 4);
            if (ppt2 <= 8) O4D_FW(4, 8);
            if (ppt2 <= 16) O4D_FW(4, 16);
            if (ppt2 <= 18) O4D_FW(4, 18);
        }
#undef O4D_FW
    }
    if (poll == 2 && threads_env == 128) {
        const int ppt2 = (int)cdiv(n, (int64_t)ctas * 128);
#define O4D_FW(C, PV) return fps_cluster_warp_launch_t<C, PV, 128>(xyz, n, ld, n_out, start, counts, order64, smem, st)
        if (ctas == 8) {
            if (ppt2 <= 5) O4D_FW(8, 5);
            if (ppt2 <= 10) O4D_FW(8, 10);
            if (ppt2 <= 14) O4D_FW(8, 14);
            if (ppt2 <= 20) O4D_FW(8, 20);
        } else if (ctas == 4) {
            if (ppt2 <= 10) O4D_FW(4, 10);
            if (ppt2 <= 16) O4D_FW(4, 16);
            if (ppt2 <= 28) O4D_FW(4, 28);
            if (ppt2 <= 36) O4D_FW(4, 36);
        }
#undef O4D_FW
    }
    const int ppt = (int)cdiv(n, (int64_t)ctas * fc::THREADS);
#define O4D_FC(C, PV)                                                                                                \
    {                                                                                                                \
        if (poll == 2) return fps_cluster_warp_launch_t<C, PV, fc::THREADS>(xyz, n, ld, n_out, start, counts, order64, smem, st); \
        return poll ? fps_cluster_launch_t<C, PV, true>(xyz, n, ld, n_out, start, counts, order64, smem, st)        \
                    : fps_cluster_launch_t<C, PV, false>(xyz, n, ld, n_out, start, counts, order64, smem, st);      \
    }
    if (ctas == 8) {
        if (ppt <= 1) O4D_FC(8, 1);
        if (ppt <= 2) O4D_FC(8, 2);
        if (ppt <= 4) O4D_FC(8, 4);
        if (ppt <= 5) O4D_FC(8, 5);
    } else if (ctas == 4) {
        if (ppt <= 2) O4D_FC(4, 2);
        if (ppt <= 4) O4D_FC(4, 4);
        if (ppt <= 8) O4D_FC(4, 8);
        if (ppt <= 9) O4D_FC(4, 9);
    } else {
        if (ppt <= 4) O4D_FC(2, 4);
        if (ppt <= 8) O4D_FC(2, 8);
        if (ppt <= 16) O4D_FC(2, 16);
        if (ppt <= 17) O4D_FC(2, 17);
    }
#undef O4D_FC
    return O4D_E_UNSUPPORTED;
}

}  // namespace o4d
