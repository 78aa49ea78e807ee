// Encoder orchestration: DownTransition (model/modules.py:70-163) and
// PointCompletionNetV3.forward (model/model.py:148-233) for one cloud.
// Everything is asynchronous on the caller's stream; scratch comes from the caller.
#include "o4d_common.cuh"

namespace o4d {

// ---- geometry half of a down transition (positions only: independent of the features) ----
// fidx (n_out) sorted FPS picks, pos_out (n_out, 3), nbr (n_out, k) nearest originals of every kept point.
static int down_geometry_launch(const float* pos, int64_t ldpos, int64_t n, int64_t n_out, int k, int64_t start_idx,
                                int32_t* fidx, int64_t* fps_idx_out, float* pos_out, int32_t* nbr, void* fps_ws,
                                size_t fps_bytes, cudaStream_t st) {
    // modules.py:133-135  fps + sort
    O4D_TRY(fps_launch(pos, n, ldpos, n_out, start_idx, fidx, fps_idx_out, nullptr, fps_ws, fps_bytes, st));
    // modules.py:137  p_sub = p[inds]
    O4D_TRY(gather_rows_launch(pos, ldpos, fidx, n_out, 3, pos_out, 3, st));
    // modules.py:142-146  k nearest originals of every kept point (set semantics)
    return knn_launch(pos_out, n_out, 3, pos, n, ldpos, k, 0, nbr, nullptr, nullptr, st);
}

// ---- feature half: y = relu([LN](W x + b)) on all rows, z = max over the neighbour rows ----
static int down_features_launch(const float* const* p, const float* x, int64_t n, int d_in, int d_out, int norm,
                                const int32_t* nbr, int64_t n_out, int k, int precision, float* y, float* z,
                                cudaStream_t st) {
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
    O4D_TRY(down_geometry_launch(pos, ldpos, n, n_out, k, start_idx, fidx, fps_idx_out, pos_out, nbr, fps_ws, fps_bytes, st));
    return down_features_launch(p, x, n, d_in, d_out, norm, nbr, n_out, k, precision, y, z, st);
}

static bool enc_cfg_ok(const o4d_encoder_config* c) {
    return c && c->d_in >= 3 && c->d_feat >= 1 && c->down_blocks >= 0 && c->down_blocks <= O4D_MAX_BLOCKS &&
           c->transition_factor >= 1 && c->pt_num_neighbors >= 1 && c->pt_num_neighbors <= O4D_MAX_K &&
           c->down_neighbors >= 1 && c->down_neighbors <= O4D_MAX_K && (c->norm == 0 || c->norm == 1) &&
           c->abstract_levels >= 1 && c->abstract_levels <= c->down_blocks + 1 && c->global_dim >= 1 &&
           c->precision >= 0 && c->precision <= 2;
}

// Workspace plan.  Geometry buffers (per level: coordinates, FPS picks, down-transition neighbours,
// FPS scratch) are separate from the feature buffers: the geometry chain runs on its own stream.
struct EncWs {
    float *xa, *xb, *xavg, *hid, *ydown;
    float* pos[O4D_MAX_BLOCKS + 1];      // coordinates at every level
    int32_t* fidx[O4D_MAX_BLOCKS];       // sorted FPS picks of every down transition
    int32_t* nbr[O4D_MAX_BLOCKS];        // (n_{l+1}, k_down) neighbours of every kept point
    char* fps_ws;
    size_t fps_bytes;
    char* sub;
    size_t sub_bytes;
};

static void enc_ws_plan(const o4d_encoder_config* c, int64_t n, Arena& a, EncWs* w) {
    const int dmax = c->d_feat << c->down_blocks;
    // widest activation: level l has n_l * d_l elements; the down transition's y is n_l * 2 d_l.
    size_t act = 0, subb = 0, ydown = 0;
    int64_t nl = n;
    int d = c->d_feat;
    for (int l = 0; l <= c->down_blocks; ++l) {
        act = act > (size_t)nl * d ? act : (size_t)nl * d;
        size_t ws = pt_block_ws(nl, nl, d, c->pt_num_neighbors, true);
        subb = subb > ws ? subb : ws;
        if (l < c->down_blocks) {
            ydown = ydown > (size_t)nl * 2 * d ? ydown : (size_t)nl * 2 * d;
            nl = cdiv(nl, c->transition_factor);
            d *= 2;
        }
    }
    size_t pre = (size_t)n * c->d_feat;
    act = act > pre ? act : pre;
    w->xa = a.get<float>(act);
    w->xb = a.get<float>(act);
    w->ydown = a.get<float>(ydown);
    w->xavg = a.get<float>((size_t)dmax);
    w->hid = a.get<float>((size_t)(c->global_dim > dmax ? c->global_dim : dmax));
    nl = n;
    for (int l = 0; l <= c->down_blocks; ++l) {
        w->pos[l] = a.get<float>((size_t)nl * 3);
        if (l < c->down_blocks) {
            const int64_t nn = cdiv(nl, c->transition_factor);
            w->fidx[l] = a.get<int32_t>((size_t)nn);
            w->nbr[l] = a.get<int32_t>((size_t)nn * c->down_neighbors);
            nl = nn;
        }
    }
    w->fps_bytes = o4d_fps_workspace_bytes(n, cdiv(n, c->transition_factor));
    w->fps_ws = a.get<char>(w->fps_bytes);
    w->sub = a.get<char>(subb);
    w->sub_bytes = subb;
}

// Side stream + events for the geometry chain, one set per (host thread, device): created on first
// use and kept (like the kernels' function attributes); nothing is shared between host threads.
struct GeoStream {
    cudaStream_t st = nullptr;
    cudaEvent_t fork = nullptr;
    cudaEvent_t done[O4D_MAX_BLOCKS] = {};
};

static int geo_stream(GeoStream** out) {
    constexpr int MAX_DEV = 64;
    static thread_local GeoStream cache[MAX_DEV];
    int dev = 0;
    O4D_CUDA(cudaGetDevice(&dev));
    O4D_REQUIRE(dev >= 0 && dev < MAX_DEV, "encoder: device index %d out of range", dev);
    GeoStream& g = cache[dev];
    if (g.st == nullptr) {
        O4D_CUDA(cudaStreamCreateWithFlags(&g.st, cudaStreamNonBlocking));
        O4D_CUDA(cudaEventCreateWithFlags(&g.fork, cudaEventDisableTiming));
        for (int i = 0; i < O4D_MAX_BLOCKS; ++i) O4D_CUDA(cudaEventCreateWithFlags(&g.done[i], cudaEventDisableTiming));
    }
    *out = &g;
    return 0;
}

int encoder_launch(const o4d_encoder_config* c, const float* const* P, const float* pcl, int64_t n,
                   const int64_t* start_idx, float* abstract_out, float* global_out, float* const* level_pos_out,
                   void* ws, size_t ws_bytes, cudaStream_t st) {
    O4D_REQUIRE(enc_cfg_ok(c), "encoder: invalid configuration");
    O4D_REQUIRE(P && pcl && abstract_out && global_out, "encoder: null pointer");
    O4D_REQUIRE(n >= 1, "encoder: empty cloud");
    Arena a(ws, ws_bytes);
    EncWs w;
    enc_ws_plan(c, n, a, &w);
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

    // ---- geometry chain (side stream).  The whole FPS / kNN pyramid depends on the input coordinates
    // only (modules.py:129-146 never looks at features), it is a serial chain that occupies 8 SMs
    // (fps_cluster.cu) and it is the longest thing in the encoder (6.6 of 10.9 ms at N = 14336), so it
    // runs concurrently with the feature path and each down transition waits for its level's event.
    GeoStream* geo = nullptr;
    O4D_TRY(geo_stream(&geo));
    O4D_TRY(copy2d_launch(pcl, c->d_in, n, 3, w.pos[0], 3, st));             // model.py:168  pos0 = pcl[..., :3]
    O4D_CUDA(cudaEventRecord(geo->fork, st));
    O4D_CUDA(cudaStreamWaitEvent(geo->st, geo->fork, 0));
    {
        int64_t nl = n;
        for (int l = 0; l < c->down_blocks; ++l) {
            const int64_t nn = cdiv(nl, c->transition_factor);               // modules.py:126
            O4D_REQUIRE(c->down_neighbors <= nl, "down: k=%d exceeds the cloud size %lld", c->down_neighbors, (long long)nl);
            O4D_TRY(down_geometry_launch(w.pos[l], 3, nl, nn, c->down_neighbors, start_idx ? start_idx[l] : 0, w.fidx[l],
                                         nullptr, w.pos[l + 1], w.nbr[l], w.fps_ws, w.fps_bytes, geo->st));
            O4D_CUDA(cudaEventRecord(geo->done[l], geo->st));
            nl = nn;
        }
    }

    // ---- feature path (caller's stream)
    // model.py:167  x0 = pre_mlp(pcl)
    O4D_TRY(linear_launch(pcl, n, c->d_in, c->d_in, pre0[0], pre0[1], c->d_feat, nullptr, 0, w.xb, c->d_feat,
                          O4D_RELU_OUT, prec, st));
    O4D_TRY(linear_launch(w.xb, n, c->d_feat, c->d_feat, pre0[2], pre0[3], c->d_feat, nullptr, 0, w.xa, c->d_feat, 0,
                          prec, st));
    if (level_pos_out && level_pos_out[0]) O4D_TRY(copy2d_launch(w.pos[0], 3, n, 3, level_pos_out[0], 3, st));

    float* x = w.xa;
    float* xo = w.xb;
    int64_t nl = n;
    int d = c->d_feat;
    int64_t abs_row = 0;  // next free row of abstract_out (skip levels first, model.py:228)
    for (int l = 0; l < c->down_blocks; ++l) {
        const float* pos = w.pos[l];
        // PointTransformerBlock (self attention)
        O4D_TRY(pt_block_launch(P + pi, x, nl, d, pos, 3, nullptr, 0, 0, 0, nullptr, 0, c->pt_num_neighbors, prec,
                                xo, nullptr, w.sub, w.sub_bytes, st));
        pi += O4D_PTBLOCK_NPARAMS;
        { float* t = x; x = xo; xo = t; }
        // DownTransition: geometry from the side stream, features here
        const float* dp[4] = {P[pi], P[pi + 1], c->norm ? P[pi + 2] : nullptr, c->norm ? P[pi + 3] : nullptr};
        pi += c->norm ? 4 : 2;
        const int64_t nn = cdiv(nl, c->transition_factor);
        O4D_CUDA(cudaStreamWaitEvent(st, geo->done[l], 0));
        O4D_TRY(down_features_launch(dp, x, nl, d, 2 * d, c->norm, w.nbr[l], nn, c->down_neighbors, prec, w.ydown, xo, st));
        { float* t = x; x = xo; xo = t; }
        nl = nn;
        d *= 2;
        pos = w.pos[l + 1];
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
    const float* pos = w.pos[c->down_blocks];
    // centre block
    O4D_TRY(pt_block_launch(P + pi, x, nl, d, pos, 3, nullptr, 0, 0, 0, nullptr, 0, c->pt_num_neighbors, prec, xo,
                            nullptr, w.sub, w.sub_bytes, st));
    pi += O4D_PTBLOCK_NPARAMS;
    { float* t = x; x = xo; xo = t; }
    // model.py:188-190  global embedding
    O4D_TRY(col_mean_launch(x, nl, d, w.xavg, st));
    O4D_TRY(linear_launch(w.xavg, 1, d, d, glob[0], glob[1], c->global_dim, nullptr, 0, w.hid, c->global_dim,
                          O4D_RELU_OUT, 0, st));
    O4D_TRY(linear_launch(w.hid, 1, c->global_dim, c->global_dim, glob[2], glob[3], c->global_dim, nullptr, 0,
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
    o4d::EncWs w;
    o4d::enc_ws_plan(c, n, a, &w);
    return a.off;
}

extern "C" int o4d_encoder_forward(const o4d_encoder_config* cfg, const float* const* params, const float* pcl,
                                   int64_t n, const int64_t* start_idx, float* abstract_out, float* global_out,
                                   float* const* level_pos_out, void* workspace, size_t workspace_bytes,
                                   void* stream) {
    return o4d::encoder_launch(cfg, params, pcl, n, start_idx, abstract_out, global_out, level_pos_out, workspace,
                               workspace_bytes, (cudaStream_t)stream);
}
