#!/usr/bin/env python
"""Generates csrc/gemm_tc2.cu (the cta_group::2 dense layer) from csrc/gemm_tc.cu by textual
transformation, so the two kernels share producers / epilogue by construction.
Usage: python tools/gen_tc2.py   (from the repo root; every replacement must match exactly once)"""
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CS = os.path.join(ROOT, 'occlusions-4d_b200', 'csrc')
src = open(os.path.join(CS, 'gemm_tc.cu')).read()


def rep(body, old, new, count=1):
    assert body.count(old) >= 1, 'pattern not found: ' + old[:60]
    return body.replace(old, new) if count == 0 else body.replace(old, new, count)


a = src.index('namespace o4d {\nnamespace tc {')
b = src.index('}  // namespace tc')
body = src[a:b]
body = rep(body, 'namespace tc {', 'namespace tc2 {')
body = rep(body, 'constexpr int STAGES = 2;', 'constexpr int STAGES = 3;')
body = rep(body, 'constexpr int STAGE_BYTES = 2 * A_HALF_BYTES + 2 * BN_MAX * BK * 2;  // 48 KB',
           'constexpr int STAGE_BYTES = 2 * A_HALF_BYTES + 2 * (BN_MAX / 2) * BK * 2;  // 32 KB: own A rows + this CTA\'s half of the B rows')
body = rep(body, '__device__ __forceinline__ void fence_proxy_async_smem()', '''__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t local_addr, uint32_t rank) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_addr), "r"(rank));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cl(uint32_t a, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\\n\\t.reg .pred p;\\n\\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\\n\\t"
        "selp.u32 %0, 1, 0, p;\\n\\t}"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cl(uint32_t a, uint32_t parity) {
    if (mbar_try_wait_cl(a, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait_cl(a, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ void fence_proxy_async_smem()''')
body = rep(body, '((uint32_t)(BM >> 4) << 24);', '((uint32_t)((2 * BM) >> 4) << 24);')
body = rep(body, '"tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\\n\\t}"', '"tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\\n\\t}"')
body = rep(body, '''__device__ __forceinline__ void umma_commit(uint32_t mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}''', '''// arrives on the mbarrier at this offset in BOTH CTAs of the pair once the MMAs issued so far have retired
__device__ __forceinline__ void umma_commit(uint32_t mbar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(mbar),
                 "h"((uint16_t)3)
                 : "memory");
}''')
body = rep(body, '__device__ long long g_dbg_tc[16];', '__device__ long long g_dbg_tc2[16];')
body = rep(body, 'g_dbg_tc[', 'g_dbg_tc2[', 0)
pa = body.index('// W (n, k) fp32 row-major (ldw) -> per (n-tile, k-chunk)')
pb = body.index('__global__ void __launch_bounds__(THREADS, 2)')
body = body[:pa] + '''// W (n, k) fp32 row-major (ldw) -> per (n-tile, k-chunk): [half 0 hi][half 0 lo][half 1 hi][half 1 lo],
// half h = rows [h * bn/2, (h+1) * bn/2) of the tile, each image [kc = 4][row-group][8 rows][8 bf16]:
// exactly what CTA h of the pair copies into its shared memory.
__global__ void pack_weight_pair_kernel(const float* __restrict__ W, int64_t ldw, PackMeta m, __nv_bfloat16* __restrict__ out) {
    const int64_t slab_elems = (int64_t)m.bn * BK;
    const int64_t total = (int64_t)m.ntiles * m.kchunks * slab_elems;
    const int hb = m.bn / 2;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t slab = e / slab_elems;
        const int within = (int)(e % slab_elems);
        const int t = (int)(slab / m.kchunks), c = (int)(slab % m.kchunks);
        const int half = within / (hb * BK);
        const int w2 = within % (hb * BK);
        const int kc = w2 / (hb * 8);
        const int rem = w2 % (hb * 8);
        const int row = rem / 8, el = rem % 8;
        const int gn = t * m.bn + half * hb + row, gk = c * BK + kc * 8 + el;
        float v = (gn < m.n && gk < m.k) ? W[(int64_t)gn * ldw + gk] : 0.f;
        __nv_bfloat16 hi, lo;
        split_bf16(v, hi, lo);
        __nv_bfloat16* dst = out + slab * 2 * slab_elems + (int64_t)half * 2 * hb * BK;
        dst[w2] = hi;
        dst[hb * BK + w2] = lo;
    }
}

''' + body[pb:]
body = rep(body, 'linear_tc_kernel(', 'linear_tc2_kernel(')
body = rep(body, '''    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row0 = (int64_t)blockIdx.x * BM;''', '''    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();          // 0 = leader (issues the MMAs of the pair), 1 = peer
    const int64_t row0 = (int64_t)blockIdx.x * BM;''')
body = rep(body, '''            mbar_init(full0 + 8 * s, PROD_WARPS + 1);   // producer warps + the weight-copy thread''', '''            // leader: its producer warps + its weight copy + the peer's producer warps (whose warp 0 also
            // vouches for the peer's weight half); peer: only its own weight copy lands here
            mbar_init(full0 + 8 * s, rank == 0 ? 2 * PROD_WARPS + 1 : 1);''')
body = rep(body, '''        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();''', '''        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();      // both CTAs' mbarriers are initialised and TMEM allocated before any remote arrive / MMA
    tc_fence_after();''')
body = rep(body, '''    const uint32_t b_half_bytes = (uint32_t)bn * BK * 2;''', '''    const int hb = bn / 2;                                   // B rows held by each CTA of the pair
    const uint32_t b_half_bytes = (uint32_t)hb * BK * 2;     // one bf16 image (hi or lo) of this CTA's B rows''')
body = rep(body, '''            fence_proxy_async_smem();   // generic-proxy stores -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(full0 + 8 * s);''', '''            fence_proxy_async_smem();   // generic-proxy stores -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) {
                if (rank == 0) {
                    mbar_arrive(full0 + 8 * s);
                } else {
                    // the peer's warp 0 also waits for the peer's own weight half before vouching for the stage
                    if (warp == 0) mbar_wait(full0 + 8 * s, ph);
                    mbar_arrive_cluster(full0 + 8 * s, 0);
                }
            }''')
body = rep(body, '''        mbar_wait(accum_bar, 0);
        tc_fence_after();''', '''        mbar_wait_cl(accum_bar, 0);
        tc_fence_after();''')
body = rep(body, '''            const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(Wp) + (size_t)tile_n * nchunks * 2 * b_half_bytes;
            for (int c = 0; c < nchunks; ++c) {
                const int s = c % STAGES;
                const uint32_t ph = (uint32_t)(c / STAGES) & 1u;
                mbar_wait(empty0 + 8 * s, ph ^ 1u);
                const uint32_t dst = smem_base + s * STAGE_BYTES + 2 * A_HALF_BYTES;
                mbar_arrive_expect_tx(full0 + 8 * s, 2 * b_half_bytes);
                bulk_g2s(dst, wsrc + (size_t)c * 2 * b_half_bytes, 2 * b_half_bytes, full0 + 8 * s);
            }''', '''            // packed per (tile, chunk): [half 0 hi][half 0 lo][half 1 hi][half 1 lo]; this CTA takes half `rank`
            const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(Wp) + (size_t)tile_n * nchunks * 4 * b_half_bytes +
                                  (size_t)rank * 2 * b_half_bytes;
            for (int c = 0; c < nchunks; ++c) {
                const int s = c % STAGES;
                const uint32_t ph = (uint32_t)(c / STAGES) & 1u;
                mbar_wait_cl(empty0 + 8 * s, ph ^ 1u);
                const uint32_t dst = smem_base + s * STAGE_BYTES + 2 * A_HALF_BYTES;
                mbar_arrive_expect_tx(full0 + 8 * s, 2 * b_half_bytes);
                bulk_g2s(dst, wsrc + (size_t)c * 4 * b_half_bytes, 2 * b_half_bytes, full0 + 8 * s);
            }''')
body = rep(body, '''        if (lane == 0) {
            const uint32_t idesc = umma_idesc(bn);
            const uint32_t lbo_a = BM * 16, lbo_b = (uint32_t)bn * 16;
            for (int c = 0; c < nchunks; ++c) {
                const int s = c % STAGES;
                const uint32_t ph = (uint32_t)(c / STAGES) & 1u;
                mbar_wait(full0 + 8 * s, ph);''', '''        if (lane == 0 && rank == 0) {
            const uint32_t idesc = umma_idesc(bn);
            const uint32_t lbo_a = BM * 16, lbo_b = (uint32_t)hb * 16;
            for (int c = 0; c < nchunks; ++c) {
                const int s = c % STAGES;
                const uint32_t ph = (uint32_t)(c / STAGES) & 1u;
                mbar_wait_cl(full0 + 8 * s, ph);''')
body = rep(body, '''    __syncthreads();
    if (warp == PROD_WARPS) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
    }
}''', '''    __syncthreads();
    cluster_sync_all();      // the peer's TMEM / shared memory stay valid until the leader's last MMA has retired
    if (warp == PROD_WARPS) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
    }
}''')
body = rep(body, '''            if (c + 1 < nchunks) load_chunk(c + 1, nxt);
            mbar_wait(empty0 + 8 * s, ph ^ 1u);''', '''            if (c + 1 < nchunks) load_chunk(c + 1, nxt);
            mbar_wait_cl(empty0 + 8 * s, ph ^ 1u);''')
old = open(os.path.join(CS, 'gemm_tc2.cu')).read()
header = old[:old.index('namespace o4d {\nnamespace tc2 {')]
tail = old[old.index('}  // namespace tc2'):]
open(os.path.join(CS, 'gemm_tc2.cu'), 'w').write(header + body + tail)
print('gemm_tc2.cu regenerated')
