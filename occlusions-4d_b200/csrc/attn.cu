// Point-transformer vector attention (self and cross).
//
// Replaces PointTransformerLayer.forward, model/point_transformer_layer.py:148-183, and the
// surrounding PointTransformerBlock, model/modules.py:45-67.  Per query row i with
// neighbours j = nbr[i, 0..k):
//     r_ij     = relu(W_p1 (p_i - p2_j) + b_p1)                    (32 wide)
//     delta_ij = W_p2 r_ij + b_p2                                               (:174)
//     a_ij     = W_a2 relu(W_a1 (q_i - K_j + delta_ij) + b_a1) + b_a2           (:176)
//     w_ij     = softmax_j(a_ij / sqrt(d))          per channel                (:177)
//     out_i    = sum_j w_ij * (V_j + delta_ij)                                  (:179)
//
// The first attention-MLP layer is linear in (q_i - K_j + delta_ij), so it is evaluated as
//     W_a1 q_i + (W_a1 b_p2 + b_a1)   -- once per query          "Qa"  (n, 2d)
//   - W_a1 K_j                        -- once per key point      "Ka"  (m, 2d), per cloud
//   + (W_a1 W_p2) r_ij                -- K = 32 contraction      "Wc"  (2d, 32)
// which cuts that layer's per-(query, neighbour) contraction from K = d to K = 32 (13x fewer
// MACs at d = 416); the reference's (n, k, d) gathers of q - k + delta never exist.  Same
// algebra, fp32 re-association only.  The two contractions that remain per pair
// (r -> 2d with the row-dependent Qa - Ka term in the epilogue, 2d -> d) and delta = W_p2 r
// go through the dense-layer kernels (tcgen05 or CUDA-core).
#include "o4d_common.cuh"

namespace o4d {

constexpr int POS_HID = 32;  // pos_mlp_hidden_dim, hard-coded at modules.py:38

// R[row, t] = relu(W_p1[t] . (p_i - p2_j) + b_p1[t]); one thread per element.
__global__ void __launch_bounds__(256)
attn_posrelu_kernel(const float* __restrict__ pos, int64_t ldpos, const float* __restrict__ pos2, int64_t ldpos2,
                    const int32_t* __restrict__ nbr, int64_t row_offset, int64_t n_rows, int k,
                    const float* __restrict__ wp1, const float* __restrict__ bp1, float* __restrict__ R) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_rows * POS_HID) return;
    const int64_t row = e / POS_HID;
    const int t = (int)(e % POS_HID);
    const int64_t ar = row_offset + row;
    const int64_t i = ar / k;
    const int j = nbr[ar];
    const float rx = pos[i * ldpos + 0] - pos2[(int64_t)j * ldpos2 + 0];
    const float ry = pos[i * ldpos + 1] - pos2[(int64_t)j * ldpos2 + 1];
    const float rz = pos[i * ldpos + 2] - pos2[(int64_t)j * ldpos2 + 2];
    const float h = fmaf(wp1[t * 3 + 2], rz, fmaf(wp1[t * 3 + 1], ry, fmaf(wp1[t * 3 + 0], rx, bp1[t])));
    R[e] = fmaxf(h, 0.f);
}

// C (p, r) = A (p, q) . B (q, r) [+ addvec (p)], fp64 accumulation.  Weight-only composites
// (a few hundred thousand outputs), computed once per (weights, cloud).
__global__ void matmul_nn_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                                 const float* __restrict__ addvec, float* __restrict__ C, int p, int q, int r) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= p * r) return;
    const int i = e / r, j = e % r;
    double acc = addvec ? (double)addvec[i] : 0.0;
    for (int c = 0; c < q; ++c) acc += (double)A[(int64_t)i * lda + c] * (double)B[(int64_t)c * ldb + j];
    C[e] = (float)acc;
}

// One thread per (query, channel): softmax over the k neighbour logits, weighted sum of
// (V_j + delta_ij).  logits / delta (rows, d) with rows = n_q * k, local to the chunk.
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

int matmul_nn_launch(const float* A, int lda, const float* B, int ldb, const float* addvec, float* C, int p, int q, int r,
                     cudaStream_t st) {
    ProfScope prof(PROF_MISC, 2.0 * p * q * r, st);
    matmul_nn_kernel<<<(unsigned)cdiv((int64_t)p * r, 256), 256, 0, st>>>(A, lda, B, ldb, addvec, C, p, q, r);
    O4D_LAUNCH_CHECK();
    return 0;
}

// ---- per-cloud tables: Ka = K W_a1^T, Wc = W_a1 W_p2, cvec = W_a1 b_p2 + b_a1 ------------------
size_t attn_tables_bytes(int64_t m, int d) {
    Arena a(nullptr, 0);
    a.get<float>((size_t)m * 2 * d);
    a.get<float>((size_t)2 * d * POS_HID);
    a.get<float>((size_t)2 * d);
    return a.off;
}

int attn_tables_launch(const PtBlockParams& P, const float* ktab, const float* vtab, int64_t m, int d, void* buf,
                       size_t buf_bytes, AttnTables* out, cudaStream_t st, bool weights_too) {
    Arena a(buf, buf_bytes);
    float* ka = a.get<float>((size_t)m * 2 * d);
    float* wc = a.get<float>((size_t)2 * d * POS_HID);
    float* cvec = a.get<float>((size_t)2 * d);
    if (!a.ok || !buf) {
        set_error("attention tables: buffer too small (%zu < %zu)", buf_bytes, a.off);
        return O4D_E_WORKSPACE;
    }
    if (weights_too) {     // Wc, cvec depend on the weights only; Ka (below) on the key cloud
        ProfScope prof(PROF_MISC, 2.0 * 2 * d * d * (POS_HID + 1), st);
        matmul_nn_kernel<<<(unsigned)cdiv(2 * d * POS_HID, 256), 256, 0, st>>>(P.wa1, d, P.wp2, POS_HID, nullptr, wc, 2 * d, d, POS_HID);
        O4D_LAUNCH_CHECK();
        matmul_nn_kernel<<<(unsigned)cdiv(2 * d, 256), 256, 0, st>>>(P.wa1, d, P.bp2, 1, P.ba1, cvec, 2 * d, d, 1);
        O4D_LAUNCH_CHECK();
    }
    O4D_TRY(linear_launch(ktab, m, d, d, P.wa1, nullptr, 2 * d, nullptr, 0, ka, 2 * d, 0, 0, st));
    out->vtab = vtab;
    out->ka = ka;
    out->wc = wc;
    out->cvec = cvec;
    out->fused = nullptr;
    return 0;
}

static int64_t attn_chunk_queries(int64_t n, int d, int k) {
    // bound the per-chunk activations (hidden 2d + logits d + delta d + r 32 floats per pair) to ~1 GiB
    int64_t per_q = (int64_t)k * (4 * d + POS_HID) * 4;
    int64_t q = ((int64_t)1 << 30) / per_q;
    q = q / 128 * 128;
    if (q < 128) q = 128;
    return q < n ? q : n;
}

size_t attn_core_workspace_bytes(int64_t n, int d, int k) {
    Arena a(nullptr, 0);
    int64_t cq = attn_chunk_queries(n, d, k);
    a.get<float>((size_t)cq * k * POS_HID);  // r
    a.get<float>((size_t)cq * k * 2 * d);    // hidden
    a.get<float>((size_t)cq * k * d);        // logits
    a.get<float>((size_t)cq * k * d);        // delta
    a.get<float>((size_t)n * 2 * d);         // Qa
    a.get<float>((size_t)n * d);             // agg
    return a.off;
}

int attn_core_launch(const PtBlockParams& P, const float* q, const AttnTables& T, const float* pos, int64_t ldpos,
                     const float* pos2, int64_t ldpos2, const int32_t* nbr, int64_t n, int d, int k,
                     const float* x_res, float* out, int precision, void* ws, size_t ws_bytes, cudaStream_t st) {
    O4D_REQUIRE(k >= 1 && k <= O4D_MAX_K, "attention: k=%d outside [1,%d]", k, O4D_MAX_K);
    O4D_REQUIRE(d >= 1, "attention: bad width %d", d);
    if (n == 0) return 0;
    Arena a(ws, ws_bytes);
    const int64_t cq = attn_chunk_queries(n, d, k);
    float* r = a.get<float>((size_t)cq * k * POS_HID);
    float* hid = a.get<float>((size_t)cq * k * 2 * d);
    float* logits = a.get<float>((size_t)cq * k * d);
    float* delta = a.get<float>((size_t)cq * k * d);
    float* qa = a.get<float>((size_t)n * 2 * d);
    float* agg = a.get<float>((size_t)n * d);
    if (P.w3 == nullptr) agg = out;  // bare PointTransformerLayer: no layer3 / residual
    if (!a.ok) {
        set_error("attention: workspace too small (%zu < %zu)", ws_bytes, a.off);
        return O4D_E_WORKSPACE;
    }
    // Qa = W_a1 q + (W_a1 b_p2 + b_a1): the per-query part of the first attention-MLP layer
    if (T.wqa)   // q is the block input x: layer1, to_q and W_a1 folded into one weight
        O4D_TRY(linear_ps_launch(P.ps, q, n, d, d, T.wqa, d, T.bqa, 2 * d, nullptr, 0, qa, 2 * d, 0, precision, st));
    else
        O4D_TRY(linear_ps_launch(P.ps, q, n, d, d, P.wa1, d, T.cvec, 2 * d, nullptr, 0, qa, 2 * d, 0, precision, st));
    const float inv_sqrt_d = (float)(1.0 / sqrt((double)d));
    const bool fused = precision != 0 && T.fused != nullptr && attn_fused_supported(d, k);
    if (fused)   // everything between Qa and the aggregated output in one tcgen05 kernel
        O4D_TRY(attn_fused_launch(P, T, qa, pos, ldpos, pos2, ldpos2, nbr, n, d, k, agg, precision, st));
    for (int64_t i0 = 0; i0 < n && !fused; i0 += cq) {
        const int64_t nq = (n - i0 < cq) ? (n - i0) : cq;
        const int64_t rows = nq * k;
        {
            ProfScope prof(PROF_ATTN_GLUE, 8.0 * (double)rows * POS_HID, st);
            attn_posrelu_kernel<<<(unsigned)cdiv(rows * POS_HID, 256), 256, 0, st>>>(pos, ldpos, pos2, ldpos2, nbr, i0 * k, rows, k,
                                                                                      P.wp1, P.bp1, r);
            O4D_LAUNCH_CHECK();
        }
        RowGather g;
        g.qa = qa;
        g.ka = T.ka;
        g.nbr = nbr;
        g.knbr = k;
        g.row_offset = i0 * k;
        // hidden = relu(Wc r + Qa_i - Ka_j)
        O4D_TRY(linear_ps_launch(P.ps, r, rows, POS_HID, POS_HID, T.wc, POS_HID, nullptr, 2 * d, nullptr, 0, hid, 2 * d,
                                 O4D_RELU_OUT, precision, st, &g));
        // logits = W_a2 hidden + b_a2
        O4D_TRY(linear_ps_launch(P.ps, hid, rows, 2 * d, 2 * d, P.wa2, 2 * d, P.ba2, d, nullptr, 0, logits, d, 0, precision, st));
        // delta = W_p2 r + b_p2
        O4D_TRY(linear_ps_launch(P.ps, r, rows, POS_HID, POS_HID, P.wp2, POS_HID, P.bp2, d, nullptr, 0, delta, d, 0, precision, st));
        {
            ProfScope prof(PROF_ATTN_GLUE, 6.0 * (double)rows * d, st);
            attn_softmax_agg_kernel<O4D_MAX_K><<<(unsigned)cdiv(nq * d, 256), 256, 0, st>>>(
                logits, delta, T.vtab, nbr, i0, nq, d, k, inv_sqrt_d, agg);
            O4D_LAUNCH_CHECK();
        }
    }
    if (P.w3 == nullptr) return 0;
    // z = x + W3 agg + b3   (modules.py:64-65)
    if (T.w3cat) {   // decoder: layer3 K-concatenated with the next block's lin_z (second operand = local features)
        RowGather cat;
        cat.a2 = T.cat_a2;
        cat.lda2 = T.cat_lda2;
        cat.k2 = T.cat_k2;
        const int64_t kc = (d + 31) / 32 * 32 + T.cat_k2;
        O4D_TRY(linear_ps_launch(P.ps, agg, n, d, d, T.w3cat, kc, T.b3cat, d, x_res, d, out, d, 0, precision, st, &cat));
        return 0;
    }
    O4D_TRY(linear_ps_launch(P.ps, agg, n, d, d, P.w3, d, P.b3, d, x_res, d, out, d, 0, precision, st));
    return 0;
}

size_t pt_block_ws(int64_t n, int64_t m, int d, int k, bool self_mode) {
    const int64_t mm = self_mode ? n : m;
    Arena a(nullptr, 0);
    a.get<float>((size_t)n * d);     // y = layer1(x)
    a.get<float>((size_t)n * d);     // q
    a.get<float>((size_t)mm * d);    // K table
    a.get<float>((size_t)mm * d);    // V table
    a.get<int32_t>((size_t)n * k);   // neighbours
    a.get<char>(attn_tables_bytes(mm, d));
    return a.off + attn_core_workspace_bytes(n, d, k);
}

// layer1/layer3 optional (null -> bare PointTransformerLayer).
static int pt_common_launch(const PtBlockParams& P, const float* x, int64_t n, int d, const float* pos, int64_t ldpos,
                            const float* x2, int64_t m, int d2, int64_t ldx2, const float* pos2, int64_t ldpos2, int k,
                            int precision, float* z, int64_t* knn_idx_out, void* ws, size_t ws_bytes, cudaStream_t st) {
    const bool self_mode = (x2 == nullptr);
    if (self_mode) {
        m = n; d2 = d; pos2 = pos; ldpos2 = ldpos;
    }
    O4D_REQUIRE(k <= m, "attention: k=%d exceeds the number of key points %lld", k, (long long)m);
    Arena a(ws, ws_bytes);
    float* y = a.get<float>((size_t)n * d);
    float* q = a.get<float>((size_t)n * d);
    float* ktab = a.get<float>((size_t)m * d);
    float* vtab = a.get<float>((size_t)m * d);
    int32_t* nbr = a.get<int32_t>((size_t)n * k);
    const size_t tb = attn_tables_bytes(m, d);
    char* tbuf = a.get<char>(tb);
    if (!a.ok) {
        set_error("attention: workspace too small (%zu)", ws_bytes);
        return O4D_E_WORKSPACE;
    }
    const float* xin = x;
    if (P.w1) {
        O4D_TRY(linear_launch(x, n, d, d, P.w1, P.b1, d, nullptr, 0, y, d, 0, precision, st));  // modules.py:61
        xin = y;
    }
    O4D_TRY(linear_launch(xin, n, d, d, P.wq, nullptr, d, nullptr, 0, q, d, 0, precision, st));      // :170
    const float* src2 = self_mode ? xin : x2;
    const int64_t ld2 = self_mode ? d : ldx2;
    O4D_TRY(linear_launch(src2, m, d2, ld2, P.wk, nullptr, d, nullptr, 0, ktab, d, 0, precision, st));  // :171
    O4D_TRY(linear_launch(src2, m, d2, ld2, P.wv, nullptr, d, nullptr, 0, vtab, d, 0, precision, st));  // :172
    O4D_TRY(knn_launch(pos, n, ldpos, pos2, m, ldpos2, k, 0, nbr, knn_idx_out, nullptr, st));           // :167
    AttnTables T;
    O4D_TRY(attn_tables_launch(P, ktab, vtab, m, d, tbuf, tb, &T, st));
    return attn_core_launch(P, q, T, pos, ldpos, pos2, ldpos2, nbr, n, d, k, x, z, precision, (char*)ws + a.off,
                            ws_bytes - a.off, st);
}

int pt_block_launch(const float* const* p, const float* x, int64_t n, int d, const float* pos,
                    int64_t ldpos, const float* x2, int64_t m, int d2, int64_t ldx2, const float* pos2,
                    int64_t ldpos2, int k, int precision, float* z, int64_t* knn_idx_out, void* ws,
                    size_t ws_bytes, cudaStream_t st) {
    O4D_REQUIRE(p && x && pos && z, "pt_block: null pointer");
    O4D_REQUIRE(x2 == nullptr || (pos2 && m >= 1 && d2 >= 1 && ldx2 >= d2), "pt_block: bad cross-attention inputs");
    PtBlockParams P = PtBlockParams::from(p);
    return pt_common_launch(P, x, n, d, pos, ldpos, x2, m, d2, ldx2, pos2, ldpos2, k, precision, z, knn_idx_out, ws,
                            ws_bytes, st);
}

// Bare PointTransformerLayer.forward (point_transformer_layer.py:148-183): no layer1/layer3.
int pt_layer_launch(const float* const* p, const float* x, int64_t n, int d, const float* pos,
                    int64_t ldpos, const float* x2, int64_t m, int d2, int64_t ldx2, const float* pos2,
                    int64_t ldpos2, int k, int precision, float* out, int64_t* knn_idx_out, void* ws,
                    size_t ws_bytes, cudaStream_t st) {
    O4D_REQUIRE(p && x && pos && out, "pt_layer: null pointer");
    O4D_REQUIRE(x2 == nullptr || (pos2 && m >= 1 && d2 >= 1 && ldx2 >= d2), "pt_layer: bad cross-attention inputs");
    PtBlockParams P;
    P.w1 = P.b1 = P.w3 = P.b3 = nullptr;
    P.wq = p[0]; P.wk = p[1]; P.wv = p[2]; P.wp1 = p[3]; P.bp1 = p[4]; P.wp2 = p[5]; P.bp2 = p[6];
    P.wa1 = p[7]; P.ba1 = p[8]; P.wa2 = p[9]; P.ba2 = p[10];
    return pt_common_launch(P, x, n, d, pos, ldpos, x2, m, d2, ldx2, pos2, ldpos2, k, precision, out, knn_idx_out, ws,
                            ws_bytes, st);
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
