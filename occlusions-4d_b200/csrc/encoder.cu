// Encoder orchestration: DownTransition (model/modules.py:70-163) and
// PointCompletionNetV3.forward (model/model.py:148-233) for one cloud.
// Everything is asynchronous on the caller's stream; scratch comes from the caller.
#include "o4d_common.cuh"

namespace o4d {

size_t down_ws(int64_t n, int d_in, int d_out, int factor, int k) {
    (void)d_in;
    const int64_t n_out = cdiv(n, factor);
    Arena a(nullptr, 0);
    a.get<int32_t>((size_t)n_out);            // sorted fps indices
    a.get<int32_t>((size_t)n_out * k);        // neighbours
    a.get<float>((size_t)n * d_out);          // y = relu(norm(W x + b)) on all rows
    a.get<char>(o4d_fps_workspace_bytes(n, n_out));
    return a.off;
}

int down_launch(const float* const* p, const float* x, int64_t n, int d_in, const float* pos,
                int64_t ldpos, int d_out, int factor, int k, int norm, int64_t start_idx, int precision,
                float* z, float* pos_out, int64_t* fps_idx_out, void* ws, size_t ws_bytes,
                cudaStream_t st) {
    O4D_REQUIRE(p && x && pos && z && pos_out, "down: null pointer");
    O4D_REQUIRE(factor >= 1 && n >= 1, "down: bad factor/size");
    if (norm != 0 && norm != 1) {
        set_error("down: norm type %d unsupported (only none / layer; the released configs never use batch norm)", norm);
        return O4D_E_UNSUPPORTED;
    }
    O4D_REQUIRE(k >= 1 && k <= n, "down: k=%d exceeds the cloud size %lld", k, (long long)n);
    const int64_t n_out = cdiv(n, factor);  // modules.py:126
    Arena a(ws, ws_bytes);
    int32_t* fidx = a.get<int32_t>((size_t)n_out);
    int32_t* nbr = a.get<int32_t>((size_t)n_out * k);
    float* y = a.get<float>((size_t)n * d_out);
    const size_t fps_bytes = o4d_fps_workspace_bytes(n, n_out);
    char* fps_ws = a.get<char>(fps_bytes);
    if (!a.ok) {
        set_error("down: workspace too small (%zu < %zu)", ws_bytes, a.off);
        return O4D_E_WORKSPACE;
    }
    // modules.py:133-135  fps + sort
    O4D_TRY(fps_launch(pos, n, ldpos, n_out, start_idx, fidx, fps_idx_out, nullptr, fps_ws, fps_bytes, st));
    // modules.py:137  p_sub = p[inds]
    O4D_TRY(gather_rows_launch(pos, ldpos, fidx, n_out, 3, pos_out, 3, st));
    // modules.py:142-146  k nearest originals of every kept point (set semantics)
    O4D_TRY(knn_launch(pos_out, n_out, 3, pos, n, ldpos, k, 0, nbr, nullptr, nullptr, st));
    // modules.py:152  mlp on ALL rows
    if (norm == 0) {
        O4D_TRY(linear_launch(x, n, d_in, d_in, p[0], p[1], d_out, nullptr, 0, y, d_out, O4D_RELU_OUT, precision, st));
    } else {
        O4D_REQUIRE(p[2] && p[3], "down: LayerNorm parameters missing");
        O4D_TRY(linear_launch(x, n, d_in, d_in, p[0], p[1], d_out, nullptr, 0, y, d_out, 0, precision, st));
        O4D_TRY(layernorm_relu_launch(y, n, d_out, p[2], p[3], 1e-5f, st));
    }
    // modules.py:156-158  local max pool
    return gather_max_launch(y, d_out, nbr, n_out, k, d_out, z, st);
}

static bool enc_cfg_ok(const o4d_encoder_config* c) {
    return c && c->d_in >= 3 && c->d_feat >= 1 && c->down_blocks >= 0 && c->down_blocks <= O4D_MAX_BLOCKS &&
           c->transition_factor >= 1 && c->pt_num_neighbors >= 1 && c->pt_num_neighbors <= O4D_MAX_K &&
           c->down_neighbors >= 1 && c->down_neighbors <= O4D_MAX_K && (c->norm == 0 || c->norm == 1) &&
           c->abstract_levels >= 1 && c->abstract_levels <= c->down_blocks + 1 && c->global_dim >= 1 &&
           c->precision >= 0 && c->precision <= 2;
}

static int enc_ws_plan(const o4d_encoder_config* c, int64_t n, Arena& a, float** xa, float** xb,
                       float** posa, float** posb, float** xavg, float** hid, char** sub, size_t* sub_bytes) {
    const int dmax = c->d_feat << c->down_blocks;
    // widest activation: level l has n_l * d_l elements; the down transition's y is n_l * 2 d_l.
    size_t act = 0, subb = 0;
    int64_t nl = n;
    int d = c->d_feat;
    for (int l = 0; l <= c->down_blocks; ++l) {
        act = act > (size_t)nl * d ? act : (size_t)nl * d;
        size_t w = pt_block_ws(nl, nl, d, c->pt_num_neighbors, true);
        subb = subb > w ? subb : w;
        if (l < c->down_blocks) {
            w = down_ws(nl, d, 2 * d, c->transition_factor, c->down_neighbors);
            subb = subb > w ? subb : w;
            nl = cdiv(nl, c->transition_factor);
            d *= 2;
        }
    }
    size_t pre = (size_t)n * c->d_feat;
    act = act > pre ? act : pre;
    *xa = a.get<float>(act);
    *xb = a.get<float>(act);
    *posa = a.get<float>((size_t)n * 3);
    *posb = a.get<float>((size_t)n * 3);
    *xavg = a.get<float>((size_t)dmax);
    *hid = a.get<float>((size_t)(c->global_dim > dmax ? c->global_dim : dmax));
    *sub = a.get<char>(subb);
    *sub_bytes = subb;
    return 0;
}

int encoder_launch(const o4d_encoder_config* c, const float* const* P, const float* pcl, int64_t n,
                   const int64_t* start_idx, float* abstract_out, float* global_out, float* const* level_pos_out,
                   void* ws, size_t ws_bytes, cudaStream_t st) {
    O4D_REQUIRE(enc_cfg_ok(c), "encoder: invalid configuration");
    O4D_REQUIRE(P && pcl && abstract_out && global_out, "encoder: null pointer");
    O4D_REQUIRE(n >= 1, "encoder: empty cloud");
    Arena a(ws, ws_bytes);
    float *xa, *xb, *posa, *posb, *xavg, *hid;
    char* sub;
    size_t sub_bytes;
    enc_ws_plan(c, n, a, &xa, &xb, &posa, &posb, &xavg, &hid, &sub, &sub_bytes);
    if (!a.ok || !ws) {
        set_error("encoder: workspace too small (%zu < %zu)", ws_bytes, a.off);
        return O4D_E_WORKSPACE;
    }
    const int prec = c->precision;
    const int dfin = c->d_feat << c->down_blocks;
    const int feat_w = 3 + dfin;  // abstract row width
    int pi = 0;
    const float* const* pre0 = P + pi; pi += 4;      // pre_mlp.0.{w,b}, pre_mlp.2.{w,b}
    const float* const* glob = P + pi; pi += 4;      // global_mlp.0.{w,b}, global_mlp.2.{w,b}
    const float* const* skipp = P + pi; pi += 2 * (c->abstract_levels - 1);

    // model.py:167-168  x0 = pre_mlp(pcl); pos0 = pcl[..., :3]
    O4D_TRY(linear_launch(pcl, n, c->d_in, c->d_in, pre0[0], pre0[1], c->d_feat, nullptr, 0, xb, c->d_feat,
                          O4D_RELU_OUT, prec, st));
    O4D_TRY(linear_launch(xb, n, c->d_feat, c->d_feat, pre0[2], pre0[3], c->d_feat, nullptr, 0, xa, c->d_feat, 0,
                          prec, st));
    O4D_TRY(copy2d_launch(pcl, c->d_in, n, 3, posa, 3, st));
    if (level_pos_out && level_pos_out[0]) O4D_TRY(copy2d_launch(posa, 3, n, 3, level_pos_out[0], 3, st));

    float* x = xa;
    float* xo = xb;
    float* pos = posa;
    float* poso = posb;
    int64_t nl = n;
    int d = c->d_feat;
    int64_t abs_row = 0;  // next free row of abstract_out (skip levels first, model.py:228)
    for (int l = 0; l < c->down_blocks; ++l) {
        // PointTransformerBlock (self attention)
        O4D_TRY(pt_block_launch(P + pi, x, nl, d, pos, 3, nullptr, 0, 0, 0, nullptr, 0, c->pt_num_neighbors, prec,
                                xo, nullptr, sub, sub_bytes, st));
        pi += O4D_PTBLOCK_NPARAMS;
        { float* t = x; x = xo; xo = t; }
        // DownTransition
        const float* dp[4] = {P[pi], P[pi + 1], c->norm ? P[pi + 2] : nullptr, c->norm ? P[pi + 3] : nullptr};
        pi += c->norm ? 4 : 2;
        O4D_TRY(down_launch(dp, x, nl, d, pos, 3, 2 * d, c->transition_factor, c->down_neighbors, c->norm,
                            start_idx ? start_idx[l] : 0, prec, xo, poso, nullptr, sub, sub_bytes, st));
        { float* t = x; x = xo; xo = t; }
        { float* t = pos; pos = poso; poso = t; }
        nl = cdiv(nl, c->transition_factor);
        d *= 2;
        if (level_pos_out && level_pos_out[l + 1]) O4D_TRY(copy2d_launch(pos, 3, nl, 3, level_pos_out[l + 1], 3, st));
        // model.py:202-207  external skip: the level whose width matches abstract_skip_mlps[j].in_features
        for (int j = 0; j < c->abstract_levels - 1; ++j) {
            if ((dfin >> (c->abstract_levels - 1 - j)) == d) {
                float* dst = abstract_out + abs_row * feat_w;
                O4D_TRY(copy2d_launch(pos, 3, nl, 3, dst, feat_w, st));
                O4D_TRY(linear_launch(x, nl, d, d, skipp[2 * j], skipp[2 * j + 1], dfin, nullptr, 0, dst + 3, feat_w, 0,
                                      prec, st));
                O4D_TRY(fill_col_launch(dst, feat_w, nl, feat_w - 1, (float)(j + 1), st));
                abs_row += nl;
            }
        }
    }
    // centre block
    O4D_TRY(pt_block_launch(P + pi, x, nl, d, pos, 3, nullptr, 0, 0, 0, nullptr, 0, c->pt_num_neighbors, prec, xo,
                            nullptr, sub, sub_bytes, st));
    pi += O4D_PTBLOCK_NPARAMS;
    { float* t = x; x = xo; xo = t; }
    // model.py:188-190  global embedding
    O4D_TRY(col_mean_launch(x, nl, d, xavg, st));
    O4D_TRY(linear_launch(xavg, 1, d, d, glob[0], glob[1], c->global_dim, nullptr, 0, hid, c->global_dim,
                          O4D_RELU_OUT, 0, st));
    O4D_TRY(linear_launch(hid, 1, c->global_dim, c->global_dim, glob[2], glob[3], c->global_dim, nullptr, 0,
                          global_out, c->global_dim, 0, 0, st));
    // model.py:220-228  pcl_out = cat([pos, x]); last channel := level id when levels > 1
    float* dst = abstract_out + abs_row * feat_w;
    O4D_TRY(copy2d_launch(pos, 3, nl, 3, dst, feat_w, st));
    O4D_TRY(copy2d_launch(x, d, nl, d, dst + 3, feat_w, st));
    if (c->abstract_levels > 1) O4D_TRY(fill_col_launch(dst, feat_w, nl, feat_w - 1, (float)c->abstract_levels, st));
    return 0;
}

}  // namespace o4d

extern "C" size_t o4d_down_workspace_bytes(int64_t n, int d_in, int d_out, int factor, int k) {
    if (n <= 0 || factor <= 0) return 0;
    return o4d::down_ws(n, d_in, d_out, factor, k);
}

extern "C" int o4d_down_forward(const float* const* p, const float* x, int64_t n, int d_in, const float* pos,
                                int64_t ldpos, int d_out, int factor, int k, int norm, int64_t start_idx,
                                int precision, float* z, float* pos_out, int64_t* fps_idx_out, void* workspace,
                                size_t workspace_bytes, void* stream) {
    return o4d::down_launch(p, x, n, d_in, pos, ldpos, d_out, factor, k, norm, start_idx, precision, z, pos_out,
                            fps_idx_out, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int o4d_encoder_num_params(const o4d_encoder_config* c) {
    if (!o4d::enc_cfg_ok(c)) return O4D_E_ARG;
    return 8 + 2 * (c->abstract_levels - 1) + (c->down_blocks + 1) * O4D_PTBLOCK_NPARAMS +
           c->down_blocks * (c->norm ? 4 : 2);
}

extern "C" int64_t o4d_encoder_num_abstract(const o4d_encoder_config* c, int64_t n) {
    if (!o4d::enc_cfg_ok(c) || n < 1) return O4D_E_ARG;
    const int dfin = c->d_feat << c->down_blocks;
    int64_t total = 0, nl = n;
    int d = c->d_feat;
    for (int l = 0; l < c->down_blocks; ++l) {
        nl = o4d::cdiv(nl, c->transition_factor);
        d *= 2;
        for (int j = 0; j < c->abstract_levels - 1; ++j)
            if ((dfin >> (c->abstract_levels - 1 - j)) == d) total += nl;
    }
    return total + nl;
}

extern "C" size_t o4d_encoder_workspace_bytes(const o4d_encoder_config* c, int64_t n) {
    if (!o4d::enc_cfg_ok(c) || n < 1) return 0;
    o4d::Arena a(nullptr, 0);
    float *xa, *xb, *pa, *pb, *av, *h;
    char* s;
    size_t sb;
    o4d::enc_ws_plan(c, n, a, &xa, &xb, &pa, &pb, &av, &h, &s, &sb);
    return a.off;
}

extern "C" int o4d_encoder_forward(const o4d_encoder_config* cfg, const float* const* params, const float* pcl,
                                   int64_t n, const int64_t* start_idx, float* abstract_out, float* global_out,
                                   float* const* level_pos_out, void* workspace, size_t workspace_bytes,
                                   void* stream) {
    return o4d::encoder_launch(cfg, params, pcl, n, start_idx, abstract_out, global_out, level_pos_out, workspace,
                               workspace_bytes, (cudaStream_t)stream);
}
