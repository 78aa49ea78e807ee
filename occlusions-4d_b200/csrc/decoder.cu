// Implicit decoder orchestration: LocalPclResnetFC.forward + do_forward_attention,
// model/implicit.py:271-445 (local_mode 'attention', ReLU, cross layers of type 'c', B = 1).
//
// Scene-constant work is hoisted into o4d_decoder_prepare_scene (the reference redoes it
// for every query mini-batch): packed abstract coordinates/features, the K/V tables
// to_k(x2), to_v(x2) of every cross layer (point_transformer_layer.py:171-172) and the
// global half of every lin_z (implicit.py:417: lin_z(cat[global, local]) =
// W[:, :Dg] g + b  +  W[:, Dg:] f_local; the first term is a per-scene vector).
#include "o4d_common.cuh"
#include "mlp_chain.cuh"
#include <stdlib.h>

namespace o4d {

__global__ void vec_add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] + b[i];
}
static int vec_add_launch(const float* a, const float* b, float* out, int n, cudaStream_t st) {
    vec_add_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(a, b, out, n);
    O4D_LAUNCH_CHECK();
    return 0;
}

static bool dec_cfg_ok(const o4d_decoder_config* c) {
    return c && c->d_in == 4 && c->d_hidden >= 1 && c->d_out >= 1 && c->d_latent == c->d_hidden &&
           c->d_latent_local >= 1 && c->d_latent_local < c->d_latent && c->n_blocks >= 1 &&
           c->n_blocks <= O4D_MAX_BLOCKS && c->pos_encoding_freqs >= 0 && c->pos_encoding_freqs <= 24 &&
           c->num_local_features >= 1 && c->num_local_features <= O4D_MAX_K && c->cross_attn_neighbors >= 1 &&
           c->cross_attn_neighbors <= O4D_MAX_K && c->cross_attn_layers >= 0 &&
           c->cross_attn_layers <= O4D_MAX_BLOCKS && c->precision >= 0 && c->precision <= 2;
}

struct DecParams {
    const float *lin_in_w, *lin_in_b, *lin_out_w, *lin_out_b;
    const float *fc0_w[O4D_MAX_BLOCKS], *fc0_b[O4D_MAX_BLOCKS], *fc1_w[O4D_MAX_BLOCKS], *fc1_b[O4D_MAX_BLOCKS];
    const float *z_w[O4D_MAX_BLOCKS], *z_b[O4D_MAX_BLOCKS];
    const float* const* pt[O4D_MAX_BLOCKS];
    int use_pt[O4D_MAX_BLOCKS];  // block -> cross layer index or -1 (implicit.py:265-269)
};

static void dec_unpack(const o4d_decoder_config* c, const float* const* P, DecParams* d) {
    int i = 0;
    d->lin_in_w = P[i++]; d->lin_in_b = P[i++];
    d->lin_out_w = P[i++]; d->lin_out_b = P[i++];
    for (int b = 0; b < c->n_blocks; ++b) {
        d->fc0_w[b] = P[i++]; d->fc0_b[b] = P[i++];
        d->fc1_w[b] = P[i++]; d->fc1_b[b] = P[i++];
    }
    for (int b = 0; b < c->n_blocks; ++b) {
        d->z_w[b] = P[i++]; d->z_b[b] = P[i++];
    }
    for (int b = 0; b < c->n_blocks; ++b) d->use_pt[b] = -1;
    for (int j = 0; j < c->cross_attn_layers; ++j) {
        d->pt[j] = P + i;
        i += O4D_PTBLOCK_NPARAMS;
        int at = ((j + 1) * c->n_blocks) / (c->cross_attn_layers + 1);
        if (at < c->n_blocks) d->use_pt[at] = j;  // later layers win a collision, like the dict at :269
    }
}

struct SceneView {
    float* abs_xyz;   // (m, 3)
    float* abs_feat;  // (m, E)
    float* zg;        // (n_blocks, H)   W_z[:, :Dg] g + b_z
    float* ktab[O4D_MAX_BLOCKS];  // (m, H) per cross layer
    float* vtab[O4D_MAX_BLOCKS];
    char* tables[O4D_MAX_BLOCKS];  // attention tables (Ka, Wc, cvec) per cross layer
    size_t tables_bytes;
    char* fused[O4D_MAX_BLOCKS];   // packed weights of the fused attention kernel (or null)
    float* wqa[O4D_MAX_BLOCKS];    // (2H, H) = W_a1 W_q W_1         per cross layer
    float* bqa[O4D_MAX_BLOCKS];    // (2H)    = W_a1 W_q b_1 + cvec
    float* t1[O4D_MAX_BLOCKS];     // (H, H)  = W_q W_1 (scratch kept for the lifetime of the scene)
    float* t1b[O4D_MAX_BLOCKS];    // (H)     = W_q b_1
    // lin_z folded into the layer that produces x (precision != 0):  for block i,
    //   wcat[i] (H, kcat[i]) = [W_pred | 0 pad to 32 | W_z[i][:, Dg:]],  bcat[i] = b_pred + zg[i]
    // with pred = lin_in (i = 0), layer3 of the cross layer after block i-1, or fc_1 of block i-1.
    float* wcat[O4D_MAX_BLOCKS];
    float* bcat[O4D_MAX_BLOCKS];
    int kcat[O4D_MAX_BLOCKS];
    size_t pack_off;  // byte offset of the pre-packed tcgen05 weights (precision != 0)
    size_t bytes;
};

static int round32(int v) { return (v + 31) / 32 * 32; }
static int pe_width(const o4d_decoder_config* c) {
    return c->pos_encoding_freqs > 0 ? c->d_in * (2 * c->pos_encoding_freqs + 1) : c->d_in;
}

static size_t packed_total_bytes(const o4d_decoder_config* c);
static bool dec_chain_possible(const o4d_decoder_config* c);

static SceneView scene_view(const o4d_decoder_config* c, int64_t m, void* base) {
    Arena a(base ? base : nullptr, base ? (size_t)-1 : 0);
    SceneView s;
    s.abs_xyz = a.get<float>((size_t)m * 3);
    s.abs_feat = a.get<float>((size_t)m * c->d_latent_local);
    s.zg = a.get<float>((size_t)c->n_blocks * c->d_hidden);
    for (int j = 0; j < c->cross_attn_layers; ++j) {
        s.ktab[j] = a.get<float>((size_t)m * c->d_hidden);
        s.vtab[j] = a.get<float>((size_t)m * c->d_hidden);
    }
    s.tables_bytes = attn_tables_bytes(m, c->d_hidden);
    for (int j = 0; j < c->cross_attn_layers; ++j) s.tables[j] = a.get<char>(s.tables_bytes);
    const bool use_fused = c->precision != 0 && attn_fused_supported(c->d_hidden, c->cross_attn_neighbors);
    for (int j = 0; j < c->cross_attn_layers; ++j)
        s.fused[j] = use_fused ? a.get<char>(attn_fused_pack_bytes(c->d_hidden)) : nullptr;
    if (use_fused && base == nullptr)
        for (int j = 0; j < c->cross_attn_layers; ++j) s.fused[j] = nullptr;
    for (int j = 0; j < c->cross_attn_layers; ++j) {
        s.wqa[j] = a.get<float>((size_t)2 * c->d_hidden * c->d_hidden);
        s.bqa[j] = a.get<float>((size_t)2 * c->d_hidden);
        s.t1[j] = a.get<float>((size_t)c->d_hidden * c->d_hidden);
        s.t1b[j] = a.get<float>((size_t)c->d_hidden);
    }
    for (int b = 0; b < c->n_blocks; ++b) {
        const int k1 = b == 0 ? pe_width(c) : c->d_hidden;
        s.kcat[b] = round32(k1) + c->d_latent_local;
        s.wcat[b] = a.get<float>((size_t)c->d_hidden * s.kcat[b]);
        s.bcat[b] = a.get<float>((size_t)c->d_hidden);
    }
    s.pack_off = a.off;
    a.get<char>(packed_total_bytes(c));
    s.bytes = a.off;
    return s;
}

// Weights the tcgen05 path consumes, in a fixed order: sizing, packing (prepare) and lookup
// (forward) all walk this list.  fn(weight, n, k, ldw).
static const float* tables_wc(const SceneView* s, int j, int64_t m, int d) {
    // second entry of the attn_tables_launch arena: Ka (m, 2d) first, then Wc
    return s && s->tables[j] ? (const float*)(s->tables[j] + align_up((size_t)m * 2 * d * sizeof(float), 256)) : nullptr;
}

template <typename F>
static void for_each_tc_weight(const o4d_decoder_config* c, const DecParams& d, const SceneView* s, int64_t m, F fn) {
    const int H = c->d_hidden, E = c->d_latent_local, Dg = c->d_latent - c->d_latent_local;
    const int pe_w = c->pos_encoding_freqs > 0 ? c->d_in * (2 * c->pos_encoding_freqs + 1) : c->d_in;
    fn(d.lin_in_w, H, pe_w, pe_w);
    fn(d.lin_out_w, c->d_out, H, H);
    for (int b = 0; b < c->n_blocks; ++b) {       // K-concatenated [pred | lin_z local] weights
        const int kc = round32(b == 0 ? pe_w : H) + E;
        fn(s ? s->wcat[b] : nullptr, H, kc, kc);
    }
    for (int b = 0; b < c->n_blocks; ++b) {
        fn(d.z_w[b] ? d.z_w[b] + Dg : nullptr, H, E, c->d_latent);  // local half of lin_z (column slice)
        fn(d.fc0_w[b], H, H, H);
        fn(d.fc1_w[b], H, H, H);
    }
    for (int j = 0; j < c->cross_attn_layers; ++j) {
        const float* const* p = d.pt[j];
        fn(s ? s->wqa[j] : nullptr, 2 * H, H, H);  // W_a1 W_q W_1: block input -> Qa
        fn(p ? p[11] : nullptr, H, 2 * H, 2 * H);  // attn_mlp.2
        fn(p ? p[13] : nullptr, H, H, H);          // layer3
        fn(p ? p[7] : nullptr, H, 32, 32);         // pos_mlp.2 (delta)
        fn(tables_wc(s, j, m, H), 2 * H, 32, 32);  // Wc = attn_mlp.0 . pos_mlp.2
    }
}

static size_t packed_total_bytes(const o4d_decoder_config* c) {
    if (c->precision == 0) return 0;
    DecParams d = {};
    size_t total = 0;
    for_each_tc_weight(c, d, nullptr, 0, [&](const float*, int n, int k, int) { total += align_up(tc_pack_bytes(n, k), 256); });
    return total;
}

static void packed_set(const o4d_decoder_config* c, const DecParams& d, const SceneView& s, int64_t m, const void* scene,
                       PackedSet* ps) {
    if (c->precision == 0) return;
    const char* p = (const char*)scene + s.pack_off;
    for_each_tc_weight(c, d, &s, m, [&](const float* w, int n, int k, int) {
        ps->add(w, p);
        p += align_up(tc_pack_bytes(n, k), 256);
    });
    ps->pair = dec_chain_possible(c) && mlp_chain_pair();
}

// weights_too = false (o4d_decoder_update_scene): `scene` already holds everything that depends on the weights only --
// composite Qa weights, Wc / cvec, the K-concatenated [W_pred | W_z,local] matrices, every packed tensor-core image --
// from an earlier prepare with the same configuration, parameters and m; only the scene-dependent parts are rewritten.
int decoder_prepare(const o4d_decoder_config* c, const float* const* P, const float* pcl_abstract, int64_t m,
                    int64_t ld, const float* feat_global, void* scene, size_t scene_bytes, cudaStream_t st,
                    bool weights_too = true) {
    O4D_REQUIRE(dec_cfg_ok(c), "decoder: invalid configuration");
    O4D_REQUIRE(P && pcl_abstract && feat_global && scene, "decoder prepare: null pointer");
    const int E = c->d_latent_local, H = c->d_hidden, Dg = c->d_latent - c->d_latent_local;
    O4D_REQUIRE(m >= 1 && ld >= 3 + E, "decoder prepare: abstract cloud must be (m, >= 3 + %d)", E);
    O4D_REQUIRE(m >= c->num_local_features && (c->cross_attn_layers == 0 || m >= c->cross_attn_neighbors),
                "decoder prepare: abstract cloud of %lld points is smaller than the neighbour count", (long long)m);
    SceneView s = scene_view(c, m, scene);
    if (scene_bytes < s.bytes) {
        set_error("decoder prepare: scene buffer too small (%zu < %zu)", scene_bytes, s.bytes);
        return O4D_E_WORKSPACE;
    }
    DecParams d;
    dec_unpack(c, P, &d);
    O4D_TRY(copy2d_launch(pcl_abstract, ld, m, 3, s.abs_xyz, 3, st));          // implicit.py:288
    O4D_TRY(copy2d_launch(pcl_abstract + 3, ld, m, E, s.abs_feat, E, st));     // implicit.py:289
    for (int b = 0; b < c->n_blocks; ++b) {
        O4D_TRY(linear_ldw_launch(feat_global, 1, Dg, Dg, d.z_w[b], c->d_latent, d.z_b[b], H, nullptr, 0,
                                  s.zg + (size_t)b * H, H, 0, 0, st));
    }
    for (int j = 0; j < c->cross_attn_layers; ++j) {
        PtBlockParams pp = PtBlockParams::from(d.pt[j]);
        O4D_TRY(linear_launch(s.abs_feat, m, E, E, pp.wk, nullptr, H, nullptr, 0, s.ktab[j], H, 0, 0, st));
        O4D_TRY(linear_launch(s.abs_feat, m, E, E, pp.wv, nullptr, H, nullptr, 0, s.vtab[j], H, 0, 0, st));
        AttnTables T;
        O4D_TRY(attn_tables_launch(pp, s.ktab[j], s.vtab[j], m, H, s.tables[j], s.tables_bytes, &T, st, weights_too));
        if (!weights_too) continue;
        if (s.fused[j]) O4D_TRY(attn_fused_pack_launch(T.wc, pp.wa2, pp.wp2, H, s.fused[j], st));
        // composite input weight of the attention block (fp64 accumulation, rounded once)
        O4D_TRY(matmul_nn_launch(pp.wq, H, pp.w1, H, nullptr, s.t1[j], H, H, H, st));
        O4D_TRY(matmul_nn_launch(pp.wa1, H, s.t1[j], H, nullptr, s.wqa[j], 2 * H, H, H, st));
        O4D_TRY(matmul_nn_launch(pp.wq, H, pp.b1, 1, nullptr, s.t1b[j], H, H, 1, st));
        O4D_TRY(matmul_nn_launch(pp.wa1, H, s.t1b[j], 1, T.cvec, s.bqa[j], 2 * H, H, 1, st));
    }
    if (c->precision != 0 && !weights_too) {
        // only b_pred + zg changes with the scene (zg = W_z[:, :Dg] g + b_z)
        for (int b = 0; b < c->n_blocks; ++b) {
            const float* bpred = b == 0 ? d.lin_in_b : (d.use_pt[b - 1] >= 0 ? d.pt[d.use_pt[b - 1]][14] : d.fc1_b[b - 1]);
            O4D_TRY(vec_add_launch(bpred, s.zg + (size_t)b * H, s.bcat[b], H, st));
        }
        return 0;
    }
    if (c->precision != 0) {
        // [W_pred | pad | W_z,local] and b_pred + zg for every block (see SceneView)
        for (int b = 0; b < c->n_blocks; ++b) {
            const float* wpred;
            const float* bpred;
            int k1;
            if (b == 0) {
                wpred = d.lin_in_w; bpred = d.lin_in_b; k1 = pe_width(c);
            } else if (d.use_pt[b - 1] >= 0) {
                wpred = d.pt[d.use_pt[b - 1]][13]; bpred = d.pt[d.use_pt[b - 1]][14]; k1 = H;
            } else {
                wpred = d.fc1_w[b - 1]; bpred = d.fc1_b[b - 1]; k1 = H;
            }
            const int kc = s.kcat[b];
            O4D_CUDA(cudaMemsetAsync(s.wcat[b], 0, (size_t)H * kc * sizeof(float), st));
            O4D_TRY(copy2d_launch(wpred, k1, H, k1, s.wcat[b], kc, st));
            O4D_TRY(copy2d_launch(d.z_w[b] + Dg, c->d_latent, H, E, s.wcat[b] + round32(k1), kc, st));
            // bcat = 1 * bpred + zg[b]  (a 1-row dense layer with the identity-free trick: copy then add)
            O4D_TRY(vec_add_launch(bpred, s.zg + (size_t)b * H, s.bcat[b], H, st));
        }
        // bf16 hi/lo shared-memory images of every weight the tcgen05 path reads
        char* p = (char*)scene + s.pack_off;
        int rc = 0;
        for_each_tc_weight(c, d, &s, m, [&](const float* w, int n, int k, int ldw) {
            // the fused multi-layer kernel is the only reader of these images whenever it can run (format: mlp_chain.cu)
            if (rc == 0) rc = dec_chain_possible(c) ? mlp_chain_pack_launch(w, n, k, ldw, p, st) : tc_pack_launch(w, n, k, ldw, p, st);
            p += align_up(tc_pack_bytes(n, k), 256);
        });
        O4D_TRY(rc);
    }
    return 0;
}

struct DecWs {
    int32_t *idx_l, *idx_c;
    float *dist_l, *f_loc, *pe, *x, *h, *y;
    char* sub;
    size_t sub_bytes, bytes;
    // activation images of the fused multi-layer path (mlp_chain.cu); null when that path cannot run
    uint8_t *img_pe, *img_floc, *img_x, *img_h, *img_y;
};

// The fused multi-layer MLP path (mlp_chain.cu) needs: tcgen05 precision, lin_z folded (packed K-concatenated
// weights), widths in whole 32-column image chunks, every layer inside the chain kernel's tile limits, and -- when
// there are cross-attention layers -- the fused attention kernel (it takes Qa and leaves the aggregate as fp32).
// O4D_MLP_CHAIN=0 forces the per-layer kernels (A/B timing).
static bool dec_chain_possible(const o4d_decoder_config* c) {
    static int env = -1;
    if (env < 0) {
        const char* e = getenv("O4D_MLP_CHAIN");
        env = (e && e[0] == '0') ? 0 : 1;
    }
    if (!env || c->precision == 0) return false;
    const int H = c->d_hidden, E = c->d_latent_local;
    const int in_w = c->pos_encoding_freqs > 0 ? c->d_in * (2 * c->pos_encoding_freqs + 1) : c->d_in;
    if (H % 32 != 0 || H < 32 || in_w < 32 || c->d_out < 4) return false;   // (tc_shape_ok of every layer of the path)
    if (c->cross_attn_layers > 0 && !attn_fused_supported(H, c->cross_attn_neighbors)) return false;
    return mlp_chain_layer_ok(H, H) && mlp_chain_layer_ok(H, 2 * H) && mlp_chain_layer_ok(H, c->d_out) &&
           mlp_chain_layer_ok((in_w + 31) / 32 * 32 + E, H) && mlp_chain_layer_ok(H + E, H);
}

static DecWs dec_ws(const o4d_decoder_config* c, int64_t nq, void* base, size_t cap) {
    Arena a(base, cap);
    DecWs w;
    const int H = c->d_hidden;
    const int pe_w = c->d_in * (2 * c->pos_encoding_freqs + 1);
    w.idx_l = a.get<int32_t>((size_t)nq * c->num_local_features);
    w.dist_l = a.get<float>((size_t)nq * c->num_local_features);
    w.idx_c = a.get<int32_t>((size_t)nq * c->cross_attn_neighbors);
    w.f_loc = a.get<float>((size_t)nq * c->d_latent_local);
    w.pe = a.get<float>((size_t)nq * pe_w);
    w.x = a.get<float>((size_t)nq * H);
    w.h = a.get<float>((size_t)nq * H);
    w.y = a.get<float>((size_t)nq * H);
    w.sub_bytes = c->cross_attn_layers > 0 ? attn_core_workspace_bytes(nq, H, c->cross_attn_neighbors) : 0;
    w.sub = a.get<char>(w.sub_bytes);
    w.img_pe = w.img_floc = w.img_x = w.img_h = w.img_y = nullptr;
    if (dec_chain_possible(c)) {
        const int in_w = c->pos_encoding_freqs > 0 ? pe_w : c->d_in;
        w.img_pe = a.get<uint8_t>(act_image_bytes(nq, in_w));
        w.img_floc = a.get<uint8_t>(act_image_bytes(nq, c->d_latent_local));
        w.img_x = a.get<uint8_t>(act_image_bytes(nq, H));
        w.img_h = a.get<uint8_t>(act_image_bytes(nq, H));
        w.img_y = a.get<uint8_t>(act_image_bytes(nq, H));
    }
    w.bytes = a.off;
    if (!a.ok) w.x = nullptr;
    return w;
}

// ---- fused multi-layer path --------------------------------------------------------------------------------------
// The whole dense part of do_forward_attention (implicit.py:403-443) as at most cross_attn_layers + 1 launches of the
// chain kernel (mlp_chain.cu), cut only where a cross-attention layer needs every query's Qa:
//   [lin_in + lin_z0] -> { fc_0 -> fc_1 (+ lin_z of the next block) }* -> Qa      | fused attention |
//   [layer3 + lin_z]  -> { ... }* -> Qa                                            | fused attention |
//   [layer3 + lin_z]  -> { ... }* -> lin_out
// Layer inputs / outputs between the launches' layers are activation images; x (the residual stream, fp32) is the
// only full-width fp32 tensor left besides Qa and the attention aggregate.
struct ChainBuild {
    mc::Program prog;
    cudaStream_t st;
    int rc = 0;
    ChainBuild(int64_t rows, int split, cudaStream_t s) : st(s) {
        prog.nops = 0;
        prog.rows = rows;
        prog.split = split;
        prog.tiles = 0;
    }
    void flush() {
        if (rc == 0 && prog.nops > 0) rc = mlp_chain_launch(prog, st);
        prog.nops = 0;
    }
    // Y = [A1 | A2] W^T + bias (+ res) -> fp32 `out` and / or image `img`
    void add(const PackedSet& ps, const uint8_t* a1, int k1, const uint8_t* a2, int k2, const float* wkey, int n,
             const float* bias, float* out, int64_t ldo, const float* res, int64_t ldr, uint8_t* img, int img_relu) {
        if (rc != 0) return;
        if (prog.nops == mc::MAX_OPS) flush();
        const void* packed = ps.find(wkey);
        if (!packed) {
            set_error("decoder: missing packed weight for the fused MLP path");
            rc = O4D_E_ARG;
            return;
        }
        mc::Op& op = prog.op[prog.nops++];
        op.a1 = a1;
        op.k1c = (k1 + 31) / 32;
        op.a2 = a2;
        op.k2c = a2 ? (k2 + 31) / 32 : 0;
        op.w = (const uint8_t*)packed;
        op.bias = bias;
        op.out = out;
        op.ldo = ldo;
        op.res = res;
        op.ldr = ldr;
        op.img = img;
        op.img_cpt = (n + 31) / 32;
        op.img_relu = img_relu;
        op.n_img = 0;
        op.n = n;
        mlp_chain_tiling(n, &op.bn, &op.ntiles);
        op.k_alg = (double)k1 + (a2 ? k2 : 0);
    }
};

static int decoder_forward_chain(const o4d_decoder_config* c, const DecParams& d, const SceneView& s, const PackedSet& ps,
                                 const DecWs& w, int64_t m, const float* query, const float* in_ptr, int in_w, int64_t nq,
                                 float* out, float* penult, cudaStream_t st) {
    const int H = c->d_hidden, E = c->d_latent_local;
    // implicit.py:403-406: the Fourier features, written straight into the image lin_in reads
    if (c->pos_encoding_freqs > 0)
        O4D_TRY(posenc_image_launch(query, nq, c->d_in, c->pos_encoding_freqs, w.img_pe, st));
    else
        O4D_TRY(act_image_launch(in_ptr, in_w, nq, in_w, 0, w.img_pe, st));
    // w.img_floc: written by the local-feature blend itself (decoder_forward)
    float* qa = (float*)w.sub;                        // (nq, 2H): first carve of the attention workspace
    ChainBuild cb(nq, c->precision == 1 ? 1 : 0, st);
    // implicit.py:403-408 + :416-418 of block 0:  x = lin_in(pe) + lin_z[0](f_query)
    cb.add(ps, w.img_pe, in_w, w.img_floc, E, s.wcat[0], H, s.bcat[0], w.x, H, nullptr, 0, w.img_x, 1);
    for (int b = 0; b < c->n_blocks; ++b) {
        const bool has_next = b + 1 < c->n_blocks;
        // implicit.py:93  h = fc_0(relu(x)): only the image of relu(h) is produced
        cb.add(ps, w.img_x, H, nullptr, 0, d.fc0_w[b], H, d.fc0_b[b], nullptr, 0, nullptr, 0, w.img_h, 1);
        if (d.use_pt[b] < 0) {
            // implicit.py:94-101 (+ :416-418 of block b + 1)  x += fc_1(relu(h)) [+ lin_z[b+1](f_query)]
            if (has_next)
                cb.add(ps, w.img_h, H, w.img_floc, E, s.wcat[b + 1], H, s.bcat[b + 1], w.x, H, w.x, H, w.img_x, 1);
            else
                cb.add(ps, w.img_h, H, nullptr, 0, d.fc1_w[b], H, d.fc1_b[b], w.x, H, w.x, H, w.img_x, 1);
            continue;
        }
        // block followed by a cross-attention layer (implicit.py:421-439 -> modules.py:61-65)
        const int j = d.use_pt[b];
        PtBlockParams pp = PtBlockParams::from(d.pt[j]);
        cb.add(ps, w.img_h, H, nullptr, 0, d.fc1_w[b], H, d.fc1_b[b], w.x, H, w.x, H, w.img_x, 0);   // raw x image for Qa
        // Qa = (W_a1 W_q W_1) x + (W_a1 W_q b_1 + cvec)
        cb.add(ps, w.img_x, H, nullptr, 0, s.wqa[j], 2 * H, s.bqa[j], qa, 2 * H, nullptr, 0, nullptr, 0);
        cb.flush();
        O4D_TRY(cb.rc);
        AttnTables T;
        {
            Arena ta(s.tables[j], s.tables_bytes);   // same carve-up as attn_tables_launch
            T.ka = ta.get<float>((size_t)m * 2 * H);
            T.wc = ta.get<float>((size_t)2 * H * 32);
            T.cvec = ta.get<float>((size_t)2 * H);
            T.vtab = s.vtab[j];
            T.fused = s.fused[j];
        }
        O4D_TRY(attn_fused_launch(pp, T, qa, query, c->d_in, s.abs_xyz, 3, w.idx_c, nq, H, c->cross_attn_neighbors, w.y,
                                  c->precision, st));
        O4D_TRY(act_image_launch(w.y, H, nq, H, 0, w.img_y, st));
        // modules.py:64-65  x += layer3(agg) [+ lin_z[b+1](f_query)]
        if (has_next)
            cb.add(ps, w.img_y, H, w.img_floc, E, s.wcat[b + 1], H, s.bcat[b + 1], w.x, H, w.x, H, w.img_x, 1);
        else
            cb.add(ps, w.img_y, H, nullptr, 0, pp.w3, H, pp.b3, w.x, H, w.x, H, w.img_x, 1);
    }
    // implicit.py:441-443  out = lin_out(relu(x))
    cb.add(ps, w.img_x, H, nullptr, 0, d.lin_out_w, c->d_out, d.lin_out_b, out, c->d_out, nullptr, 0, nullptr, 0);
    cb.flush();
    O4D_TRY(cb.rc);
    if (penult) O4D_TRY(copy2d_launch(w.x, H, nq, H, penult, H, st));           // implicit.py:441
    return 0;
}

int decoder_forward(const o4d_decoder_config* c, const float* const* P, const void* scene, int64_t m,
                    const float* query, int64_t nq, float* out, float* penult, void* ws, size_t ws_bytes,
                    cudaStream_t st) {
    O4D_REQUIRE(dec_cfg_ok(c), "decoder: invalid configuration");
    O4D_REQUIRE(P && scene && out, "decoder forward: null pointer");
    O4D_REQUIRE(nq >= 0 && m >= 1, "decoder forward: bad sizes");
    if (nq == 0) return 0;
    O4D_REQUIRE(query != nullptr, "decoder forward: null query");
    const int H = c->d_hidden, E = c->d_latent_local, Dg = c->d_latent - E;
    const int prec = c->precision;
    const int pe_w = c->d_in * (2 * c->pos_encoding_freqs + 1);
    SceneView s = scene_view(c, m, const_cast<void*>(scene));
    DecWs w = dec_ws(c, nq, ws, ws_bytes);
    if (!ws || !w.x) {
        set_error("decoder forward: workspace too small (%zu < %zu)", ws_bytes, w.bytes);
        return O4D_E_WORKSPACE;
    }
    DecParams d;
    dec_unpack(c, P, &d);
    PackedSet ps;
    packed_set(c, d, s, m, scene, &ps);

    // implicit.py:328-339  K_l nearest abstract points, inverse-distance blend of their features.
    // point_transformer_layer.py:167  K_c nearest abstract points -- identical for every cross layer (same query / abstract
    // cloud).  Both lists come from ONE scan of the abstract cloud when K_l < K_c (the released configurations: 8 / 14).
    const bool one_scan = c->cross_attn_layers > 0 && c->cross_attn_neighbors >= 9 && c->num_local_features < c->cross_attn_neighbors;
    if (one_scan)
        O4D_TRY(knn_dual_launch(query, nq, c->d_in, s.abs_xyz, m, 3, c->cross_attn_neighbors, w.idx_c, c->num_local_features,
                                w.idx_l, w.dist_l, st));
    else
        O4D_TRY(knn_launch(query, nq, c->d_in, s.abs_xyz, m, 3, c->num_local_features, 1, w.idx_l, nullptr, w.dist_l, st));
    // (fused multi-layer path: straight into the activation image lin_z's layers read; no fp32 copy is needed)
    const bool chain = w.img_x != nullptr && nq >= 1024;
    O4D_TRY(local_blend_launch(w.idx_l, w.dist_l, s.abs_feat, E, nq, c->num_local_features, E, w.f_loc, E, st,
                               chain ? w.img_floc : nullptr));
    // point_transformer_layer.py:167 -- identical for every cross layer (same query / abstract cloud)
    if (c->cross_attn_layers > 0 && !one_scan)
        O4D_TRY(knn_launch(query, nq, c->d_in, s.abs_xyz, m, 3, c->cross_attn_neighbors, 0, w.idx_c, nullptr, nullptr, st));
    // lin_z folded into the producing layer whenever the packed K-concatenated weights exist
    const bool fold = prec != 0 && ps.find(s.wcat[0]) != nullptr && tc_shape_ok(nq, H, H) &&
                      tc_shape_ok(nq, c->pos_encoding_freqs > 0 ? pe_w : c->d_in, H);
    RowGather cat;
    cat.a2 = w.f_loc;
    cat.lda2 = E;
    cat.k2 = E;
    // implicit.py:403-408 (+ :416-418 of block 0 when folded)
    const float* in_ptr = c->pos_encoding_freqs > 0 ? w.pe : query;
    const int in_w = c->pos_encoding_freqs > 0 ? pe_w : c->d_in;
    if (c->pos_encoding_freqs > 0 && !chain) O4D_TRY(posenc_launch(query, nq, c->d_in, c->pos_encoding_freqs, w.pe, st));
    if (chain) {
        if (!fold) {
            set_error("decoder: fused MLP path selected but the folded weights are missing");
            return O4D_E_ARG;
        }
        return decoder_forward_chain(c, d, s, ps, w, m, query, in_ptr, in_w, nq, out, penult, st);
    }
    if (fold)
        O4D_TRY(linear_ps_launch(&ps, in_ptr, nq, in_w, in_w, s.wcat[0], s.kcat[0], s.bcat[0], H, nullptr, 0, w.x, H, 0,
                                 prec, st, &cat));
    else
        O4D_TRY(linear_ps_launch(&ps, in_ptr, nq, in_w, in_w, d.lin_in_w, in_w, d.lin_in_b, H, nullptr, 0, w.x, H, 0, prec, st));
    for (int b = 0; b < c->n_blocks; ++b) {
        // implicit.py:416-418  x += lin_z(features_query)   (global half pre-reduced into zg)
        if (!fold)
            O4D_TRY(linear_ps_launch(&ps, w.f_loc, nq, E, E, d.z_w[b] + Dg, c->d_latent, s.zg + (size_t)b * H, H, w.x, H,
                                     w.x, H, 0, prec, st));
        const bool fold_next = fold && b + 1 < c->n_blocks;       // this block's last layer also adds lin_z[b+1]
        // implicit.py:93-101  x += fc_1(relu(fc_0(relu(x))))
        O4D_TRY(linear_ps_launch(&ps, w.x, nq, H, H, d.fc0_w[b], H, d.fc0_b[b], H, nullptr, 0, w.h, H, O4D_RELU_IN, prec, st));
        if (fold_next && d.use_pt[b] < 0)
            O4D_TRY(linear_ps_launch(&ps, w.h, nq, H, H, s.wcat[b + 1], s.kcat[b + 1], s.bcat[b + 1], H, w.x, H, w.x, H,
                                     O4D_RELU_IN, prec, st, &cat));
        else
            O4D_TRY(linear_ps_launch(&ps, w.h, nq, H, H, d.fc1_w[b], H, d.fc1_b[b], H, w.x, H, w.x, H, O4D_RELU_IN, prec, st));
        if (d.use_pt[b] >= 0) {
            // implicit.py:421-439 -> modules.py:61-65 cross attention onto the abstract cloud
            const int j = d.use_pt[b];
            PtBlockParams pp = PtBlockParams::from(d.pt[j]);
            pp.ps = &ps;
            AttnTables T;
            {
                Arena ta(s.tables[j], s.tables_bytes);   // same carve-up as attn_tables_launch
                T.ka = ta.get<float>((size_t)m * 2 * H);
                T.wc = ta.get<float>((size_t)2 * H * 32);
                T.cvec = ta.get<float>((size_t)2 * H);
                T.vtab = s.vtab[j];
                T.fused = s.fused[j];
                T.wqa = s.wqa[j];
                T.bqa = s.bqa[j];
            }
            if (fold_next) {                              // layer3 of this block also adds lin_z[b+1]
                T.w3cat = s.wcat[b + 1];
                T.b3cat = s.bcat[b + 1];
                T.cat_a2 = w.f_loc;
                T.cat_lda2 = E;
                T.cat_k2 = E;
            }
            O4D_TRY(attn_core_launch(pp, w.x, T, query, c->d_in, s.abs_xyz, 3, w.idx_c, nq, H,
                                     c->cross_attn_neighbors, w.x, w.x, prec, w.sub, w.sub_bytes, st));
        }
    }
    if (penult) O4D_TRY(copy2d_launch(w.x, H, nq, H, penult, H, st));           // implicit.py:441
    // implicit.py:442-443
    return linear_ps_launch(&ps, w.x, nq, H, H, d.lin_out_w, H, d.lin_out_b, c->d_out, nullptr, 0, out, c->d_out,
                            O4D_RELU_IN, prec, st);
}

}  // namespace o4d

extern "C" int o4d_decoder_num_params(const o4d_decoder_config* c) {
    if (!o4d::dec_cfg_ok(c)) return O4D_E_ARG;
    return 4 + 6 * c->n_blocks + O4D_PTBLOCK_NPARAMS * c->cross_attn_layers;
}

extern "C" size_t o4d_decoder_scene_bytes(const o4d_decoder_config* c, int64_t m) {
    if (!o4d::dec_cfg_ok(c) || m < 1) return 0;
    return o4d::scene_view(c, m, nullptr).bytes;
}

extern "C" int o4d_decoder_prepare_scene(const o4d_decoder_config* cfg, const float* const* params,
                                         const float* pcl_abstract, int64_t m, int64_t ld_abstract,
                                         const float* feat_global, void* scene, size_t scene_bytes, void* stream) {
    return o4d::decoder_prepare(cfg, params, pcl_abstract, m, ld_abstract, feat_global, scene, scene_bytes,
                                (cudaStream_t)stream);
}

extern "C" int o4d_decoder_update_scene(const o4d_decoder_config* cfg, const float* const* params,
                                        const float* pcl_abstract, int64_t m, int64_t ld_abstract,
                                        const float* feat_global, void* scene, size_t scene_bytes, void* stream) {
    return o4d::decoder_prepare(cfg, params, pcl_abstract, m, ld_abstract, feat_global, scene, scene_bytes,
                                (cudaStream_t)stream, false);
}

extern "C" size_t o4d_decoder_workspace_bytes(const o4d_decoder_config* c, int64_t nq, int64_t m) {
    (void)m;
    if (!o4d::dec_cfg_ok(c) || nq < 1) return 0;
    return o4d::dec_ws(c, nq, nullptr, 0).bytes;
}

extern "C" int o4d_decoder_forward(const o4d_decoder_config* cfg, const float* const* params, const void* scene,
                                   int64_t m, const float* query, int64_t nq, float* out, float* penult,
                                   void* workspace, size_t workspace_bytes, void* stream) {
    return o4d::decoder_forward(cfg, params, scene, m, query, nq, out, penult, workspace, workspace_bytes,
                                (cudaStream_t)stream);
}

extern "C" size_t o4d_decoder_run_host_device_bytes(const o4d_decoder_config* c, int64_t batch, int64_t m) {
    if (!o4d::dec_cfg_ok(c) || batch < 1) return 0;
    // double-buffered query / output staging + one forward workspace
    size_t stage = 2 * (o4d::align_up((size_t)batch * 4 * sizeof(float), 256) +
                        o4d::align_up((size_t)batch * c->d_out * sizeof(float), 256));
    return stage + o4d_decoder_workspace_bytes(c, batch, m);
}

extern "C" int o4d_decoder_run_host(const o4d_decoder_config* c, const float* const* params, const void* scene,
                                    int64_t m, const float* query_host, int64_t nq, int64_t batch, float* out_host,
                                    void* device_scratch, size_t device_scratch_bytes, void* stream) {
    using namespace o4d;
    O4D_REQUIRE(dec_cfg_ok(c), "decoder: invalid configuration");
    O4D_REQUIRE(query_host && out_host && device_scratch && batch >= 1 && nq >= 0, "decoder run_host: bad argument");
    if (device_scratch_bytes < o4d_decoder_run_host_device_bytes(c, batch, m)) {
        set_error("decoder run_host: device scratch too small");
        return O4D_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    Arena a(device_scratch, device_scratch_bytes);
    float* qd[2];
    float* od[2];
    for (int i = 0; i < 2; ++i) {
        qd[i] = a.get<float>((size_t)batch * 4);
        od[i] = a.get<float>((size_t)batch * c->d_out);
    }
    void* ws = (char*)device_scratch + a.off;
    const size_t ws_bytes = device_scratch_bytes - a.off;
    // eval/inference.py:204-246: H2D of the mini-batch, forward, D2H of the result.  All on one
    // stream (the copies are tiny next to the forward); buffers alternate so a pinned-host
    // caller gets copy/compute overlap from the copy engines without extra streams here.
    int slot = 0;
    for (int64_t s0 = 0; s0 < nq; s0 += batch, slot ^= 1) {
        const int64_t nb = (nq - s0 < batch) ? (nq - s0) : batch;
        O4D_CUDA(cudaMemcpyAsync(qd[slot], query_host + s0 * 4, (size_t)nb * 4 * sizeof(float),
                                 cudaMemcpyHostToDevice, st));
        O4D_TRY(decoder_forward(c, params, scene, m, qd[slot], nb, od[slot], nullptr, ws, ws_bytes, st));
        O4D_CUDA(cudaMemcpyAsync(out_host + s0 * c->d_out, od[slot], (size_t)nb * c->d_out * sizeof(float),
                                 cudaMemcpyDeviceToHost, st));
    }
    O4D_CUDA(cudaStreamSynchronize(st));
    return 0;
}
