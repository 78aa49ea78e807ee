// Point-transformer vector attention (self and cross).
//
// Replaces PointTransformerLayer.forward, model/point_transformer_layer.py:148-183, and the
// surrounding PointTransformerBlock, model/modules.py:45-67.  Per query row i with
// neighbours j = nbr[i, 0..k):
//     delta_ij = W_p2 relu(W_p1 (p_i - p2_j) + b_p1) + b_p2                (:174)
//     a_ij     = W_a2 relu(W_a1 (q_i - K_j + delta_ij) + b_a1) + b_a2       (:176)
//     w_ij     = softmax_j(a_ij / sqrt(d))          per channel            (:177)
//     out_i    = sum_j w_ij * (V_j + delta_ij)                              (:179)
// K = to_k(x2), V = to_v(x2) are computed ONCE per cloud as (m, d) tables and gathered
// (the reference gathers after the projection too, :171-172).
//
// This file holds the gather / positional-MLP / softmax-aggregate kernels; the two wide
// contractions of the attention MLP go through linear_launch (tcgen05 or CUDA-core).
#include "o4d_common.cuh"

namespace o4d {

constexpr int POS_HID = 32;  // pos_mlp_hidden_dim, hard-coded at modules.py:38

// One warp per (query, neighbour) row: lane t owns hidden unit t of the positional MLP,
// then lanes stride over the d output channels.  W_p2 is staged transposed ([t][c]) in
// shared memory so the channel loop is conflict-free.
//   delta (rows, d)   a1 (rows, d) = q_i - K_j + delta
__global__ void __launch_bounds__(256)
attn_prep_kernel(const float* __restrict__ q, const float* __restrict__ ktab,
                 const float* __restrict__ pos, int64_t ldpos, const float* __restrict__ pos2,
                 int64_t ldpos2, const int32_t* __restrict__ nbr, int64_t i0, int64_t n_rows, int d,
                 int k, const float* __restrict__ wp1, const float* __restrict__ bp1,
                 const float* __restrict__ wp2, const float* __restrict__ bp2,
                 float* __restrict__ delta, float* __restrict__ a1) {
    extern __shared__ float s_wp2t[];  // [POS_HID][d]
    for (int e = threadIdx.x; e < d * POS_HID; e += blockDim.x) {
        int c = e / POS_HID, t = e % POS_HID;
        s_wp2t[t * d + c] = wp2[e];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const float w0 = wp1[lane * 3 + 0], w1 = wp1[lane * 3 + 1], w2 = wp1[lane * 3 + 2];
    const float b = bp1[lane];
    for (int64_t row = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < n_rows;
         row += (int64_t)gridDim.x * warps_per_block) {
        const int64_t i = i0 + row / k;  // absolute query index
        const int j = nbr[i0 * k + row];
        const float rx = pos[i * ldpos + 0] - pos2[(int64_t)j * ldpos2 + 0];
        const float ry = pos[i * ldpos + 1] - pos2[(int64_t)j * ldpos2 + 1];
        const float rz = pos[i * ldpos + 2] - pos2[(int64_t)j * ldpos2 + 2];
        float h = fmaf(w2, rz, fmaf(w1, ry, fmaf(w0, rx, b)));
        h = fmaxf(h, 0.f);
        const float* qi = q + i * d;
        const float* kj = ktab + (int64_t)j * d;
        for (int c0 = 0; c0 < d; c0 += 32) {
            const int c = c0 + lane;
            const bool ok = c < d;
            float acc = ok ? bp2[c] : 0.f;
#pragma unroll
            for (int t = 0; t < POS_HID; ++t) {
                float ht = __shfl_sync(0xffffffffu, h, t);
                if (ok) acc = fmaf(s_wp2t[t * d + c], ht, acc);
            }
            if (ok) {
                delta[row * d + c] = acc;
                a1[row * d + c] = qi[c] - kj[c] + acc;
            }
        }
    }
}

// One thread per (query, channel): softmax over the k neighbour logits, weighted sum of
// (V_j + delta_ij).  logits (rows, d) with rows = n_q * k.
template <int KMAX>
__global__ void __launch_bounds__(256)
attn_softmax_agg_kernel(const float* __restrict__ logits, const float* __restrict__ delta,
                        const float* __restrict__ vtab, const int32_t* __restrict__ nbr, int64_t i0,
                        int64_t n_q, int d, int k, float inv_sqrt_d, float* __restrict__ agg) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_q * d) return;
    const int64_t il = e / d;  // query index local to this chunk
    const int c = (int)(e % d);
    const int32_t* nb = nbr + (i0 + il) * k;
    float a[KMAX];
    float mx = -3.4e38f;
#pragma unroll
    for (int j = 0; j < KMAX; ++j) {
        if (j < k) {
            a[j] = logits[(il * k + j) * d + c] * inv_sqrt_d;
            mx = fmaxf(mx, a[j]);
        }
    }
    float den = 0.f, num = 0.f;
#pragma unroll
    for (int j = 0; j < KMAX; ++j) {
        if (j < k) {
            float w = expf(a[j] - mx);
            den += w;
            float val = vtab[(int64_t)nb[j] * d + c] + delta[(il * k + j) * d + c];
            num = fmaf(w, val, num);
        }
    }
    agg[(i0 + il) * d + c] = num / den;
}

static int64_t attn_chunk_queries(int64_t n, int d, int k) {
    // bound the (rows, 2d) hidden activation to ~1 GiB
    int64_t per_q = (int64_t)k * d * 4 * 4;  // delta + a1 + hidden(2d)
    int64_t q = ((int64_t)1 << 30) / per_q;
    q = q / 128 * 128;
    if (q < 128) q = 128;
    return q < n ? q : n;
}

size_t attn_core_workspace_bytes(int64_t n, int d, int k) {
    Arena a(nullptr, 0);
    int64_t cq = attn_chunk_queries(n, d, k);
    a.get<float>((size_t)cq * k * d);      // delta
    a.get<float>((size_t)cq * k * d);      // a1 / logits
    a.get<float>((size_t)cq * k * 2 * d);  // hidden
    a.get<float>((size_t)n * d);           // agg
    return a.off;
}

int attn_core_launch(const PtBlockParams& P, const float* q, const float* ktab, const float* vtab,
                     const float* pos, int64_t ldpos, const float* pos2, int64_t ldpos2,
                     const int32_t* nbr, int64_t n, int d, int k, const float* x_res, float* out,
                     int precision, void* ws, size_t ws_bytes, cudaStream_t st) {
    O4D_REQUIRE(k >= 1 && k <= O4D_MAX_K, "attention: k=%d outside [1,%d]", k, O4D_MAX_K);
    O4D_REQUIRE(d >= 1 && (size_t)d * POS_HID * 4 <= 200 * 1024, "attention: width %d unsupported", d);
    if (n == 0) return 0;
    Arena a(ws, ws_bytes);
    const int64_t cq = attn_chunk_queries(n, d, k);
    float* delta = a.get<float>((size_t)cq * k * d);
    float* a1 = a.get<float>((size_t)cq * k * d);
    float* hid = a.get<float>((size_t)cq * k * 2 * d);
    float* agg = a.get<float>((size_t)n * d);
    if (P.w3 == nullptr) agg = out;  // bare PointTransformerLayer: no layer3 / residual
    if (!a.ok) {
        set_error("attention: workspace too small (%zu < %zu)", ws_bytes, a.off);
        return O4D_E_WORKSPACE;
    }
    const size_t smem = (size_t)d * POS_HID * sizeof(float);
    if (smem > 48 * 1024) {
        O4D_CUDA(cudaFuncSetAttribute(attn_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    const float inv_sqrt_d = (float)(1.0 / sqrt((double)d));
    for (int64_t i0 = 0; i0 < n; i0 += cq) {
        const int64_t nq = (n - i0 < cq) ? (n - i0) : cq;
        const int64_t rows = nq * k;
        int64_t blocks = cdiv(rows, 8);
        if (blocks > 148 * 8) blocks = 148 * 8;
        {
            ProfScope prof(PROF_ATTN_GLUE, 2.0 * (double)rows * (3 + d) * POS_HID, st);
            attn_prep_kernel<<<(unsigned)blocks, 256, smem, st>>>(q, ktab, pos, ldpos, pos2, ldpos2, nbr, i0, rows, d,
                                                                  k, P.wp1, P.bp1, P.wp2, P.bp2, delta, a1);
            O4D_LAUNCH_CHECK();
        }
        O4D_TRY(linear_ps_launch(P.ps, a1, rows, d, d, P.wa1, d, P.ba1, 2 * d, nullptr, 0, hid, 2 * d, O4D_RELU_OUT, precision, st));
        O4D_TRY(linear_ps_launch(P.ps, hid, rows, 2 * d, 2 * d, P.wa2, 2 * d, P.ba2, d, nullptr, 0, a1, d, 0, precision, st));
        {
            ProfScope prof(PROF_ATTN_GLUE, 6.0 * (double)rows * d, st);
            attn_softmax_agg_kernel<O4D_MAX_K><<<(unsigned)cdiv(nq * d, 256), 256, 0, st>>>(
                a1, delta, vtab, nbr, i0, nq, d, k, inv_sqrt_d, agg);
            O4D_LAUNCH_CHECK();
        }
    }
    if (P.w3 == nullptr) return 0;
    // z = x + W3 agg + b3   (modules.py:64-65)
    O4D_TRY(linear_ps_launch(P.ps, agg, n, d, d, P.w3, d, P.b3, d, x_res, d, out, d, 0, precision, st));
    return 0;
}

size_t pt_block_ws(int64_t n, int64_t m, int d, int k, bool self_mode) {
    Arena a(nullptr, 0);
    a.get<float>((size_t)n * d);                    // y = layer1(x)
    a.get<float>((size_t)n * d);                    // q
    a.get<float>((size_t)(self_mode ? n : m) * d);  // K table
    a.get<float>((size_t)(self_mode ? n : m) * d);  // V table
    a.get<int32_t>((size_t)n * k);                  // neighbours
    return a.off + attn_core_workspace_bytes(n, d, k);
}

int pt_block_launch(const float* const* p, const float* x, int64_t n, int d, const float* pos,
                    int64_t ldpos, const float* x2, int64_t m, int d2, int64_t ldx2, const float* pos2,
                    int64_t ldpos2, int k, int precision, float* z, int64_t* knn_idx_out, void* ws,
                    size_t ws_bytes, cudaStream_t st) {
    O4D_REQUIRE(p && x && pos && z, "pt_block: null pointer");
    const bool self_mode = (x2 == nullptr);
    if (self_mode) {
        m = n; d2 = d; pos2 = pos; ldpos2 = ldpos;
    } else {
        O4D_REQUIRE(pos2 && m >= 1 && d2 >= 1 && ldx2 >= d2, "pt_block: bad cross-attention inputs");
    }
    O4D_REQUIRE(k <= m, "pt_block: k=%d exceeds the number of key points %lld", k, (long long)m);
    PtBlockParams P = PtBlockParams::from(p);
    Arena a(ws, ws_bytes);
    float* y = a.get<float>((size_t)n * d);
    float* q = a.get<float>((size_t)n * d);
    float* ktab = a.get<float>((size_t)m * d);
    float* vtab = a.get<float>((size_t)m * d);
    int32_t* nbr = a.get<int32_t>((size_t)n * k);
    if (!a.ok) {
        set_error("pt_block: workspace too small (%zu)", ws_bytes);
        return O4D_E_WORKSPACE;
    }
    O4D_TRY(linear_launch(x, n, d, d, P.w1, P.b1, d, nullptr, 0, y, d, 0, precision, st));       // modules.py:61
    O4D_TRY(linear_launch(y, n, d, d, P.wq, nullptr, d, nullptr, 0, q, d, 0, precision, st));    // :170
    const float* src2 = self_mode ? y : x2;
    const int64_t ld2 = self_mode ? d : ldx2;
    O4D_TRY(linear_launch(src2, m, d2, ld2, P.wk, nullptr, d, nullptr, 0, ktab, d, 0, precision, st));  // :171
    O4D_TRY(linear_launch(src2, m, d2, ld2, P.wv, nullptr, d, nullptr, 0, vtab, d, 0, precision, st));  // :172
    O4D_TRY(knn_launch(pos, n, ldpos, pos2, m, ldpos2, k, 0, nbr, knn_idx_out, nullptr, st));           // :167
    return attn_core_launch(P, q, ktab, vtab, pos, ldpos, pos2, ldpos2, nbr, n, d, k, x, z, precision,
                            (char*)ws + a.off, ws_bytes - a.off, st);
}

// Bare PointTransformerLayer.forward (point_transformer_layer.py:148-183): no layer1/layer3.
int pt_layer_launch(const float* const* p, const float* x, int64_t n, int d, const float* pos,
                    int64_t ldpos, const float* x2, int64_t m, int d2, int64_t ldx2, const float* pos2,
                    int64_t ldpos2, int k, int precision, float* out, int64_t* knn_idx_out, void* ws,
                    size_t ws_bytes, cudaStream_t st) {
    O4D_REQUIRE(p && x && pos && out, "pt_layer: null pointer");
    const bool self_mode = (x2 == nullptr);
    if (self_mode) {
        m = n; d2 = d; pos2 = pos; ldpos2 = ldpos; x2 = x; ldx2 = d;
    } else {
        O4D_REQUIRE(pos2 && m >= 1 && d2 >= 1 && ldx2 >= d2, "pt_layer: bad cross-attention inputs");
    }
    O4D_REQUIRE(k <= m, "pt_layer: k=%d exceeds the number of key points %lld", k, (long long)m);
    PtBlockParams P;
    P.w1 = P.b1 = P.w3 = P.b3 = nullptr;
    P.wq = p[0]; P.wk = p[1]; P.wv = p[2]; P.wp1 = p[3]; P.bp1 = p[4]; P.wp2 = p[5]; P.bp2 = p[6];
    P.wa1 = p[7]; P.ba1 = p[8]; P.wa2 = p[9]; P.ba2 = p[10];
    Arena a(ws, ws_bytes);
    a.get<float>((size_t)n * d);  // (slot kept so the layout matches pt_block_ws)
    float* q = a.get<float>((size_t)n * d);
    float* ktab = a.get<float>((size_t)m * d);
    float* vtab = a.get<float>((size_t)m * d);
    int32_t* nbr = a.get<int32_t>((size_t)n * k);
    if (!a.ok) {
        set_error("pt_layer: workspace too small (%zu)", ws_bytes);
        return O4D_E_WORKSPACE;
    }
    O4D_TRY(linear_launch(x, n, d, d, P.wq, nullptr, d, nullptr, 0, q, d, 0, precision, st));
    O4D_TRY(linear_launch(x2, m, d2, ldx2, P.wk, nullptr, d, nullptr, 0, ktab, d, 0, precision, st));
    O4D_TRY(linear_launch(x2, m, d2, ldx2, P.wv, nullptr, d, nullptr, 0, vtab, d, 0, precision, st));
    O4D_TRY(knn_launch(pos, n, ldpos, pos2, m, ldpos2, k, 0, nbr, knn_idx_out, nullptr, st));
    return attn_core_launch(P, q, ktab, vtab, pos, ldpos, pos2, ldpos2, nbr, n, d, k, nullptr, out, precision,
                            (char*)ws + a.off, ws_bytes - a.off, st);
}

}  // namespace o4d

extern "C" int o4d_pt_layer_forward(const float* const* p, const float* x, int64_t n, int d,
                                    const float* pos, int64_t ldpos, const float* x2, int64_t m, int d2,
                                    int64_t ldx2, const float* pos2, int64_t ldpos2, int k, int precision,
                                    float* out, int64_t* knn_idx_out, void* workspace,
                                    size_t workspace_bytes, void* stream) {
    return o4d::pt_layer_launch(p, x, n, d, pos, ldpos, x2, m, d2, ldx2, pos2, ldpos2, k, precision, out,
                                knn_idx_out, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" size_t o4d_pt_block_workspace_bytes(int64_t n, int64_t m, int d, int d2, int k) {
    (void)d2;
    if (n <= 0 || d <= 0 || k <= 0) return 0;
    return o4d::pt_block_ws(n, m > 0 ? m : n, d, k, m <= 0);
}

extern "C" int o4d_pt_block_forward(const float* const* p, const float* x, int64_t n, int d,
                                    const float* pos, int64_t ldpos, const float* x2, int64_t m, int d2,
                                    int64_t ldx2, const float* pos2, int64_t ldpos2, int k, int precision,
                                    float* z, int64_t* knn_idx_out, void* workspace,
                                    size_t workspace_bytes, void* stream) {
    return o4d::pt_block_launch(p, x, n, d, pos, ldpos, x2, m, d2, ldx2, pos2, ldpos2, k, precision, z,
                                knn_idx_out, workspace, workspace_bytes, (cudaStream_t)stream);
}
