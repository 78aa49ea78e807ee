/*
 * o4d.h -- C ABI of libo4d.so, the B200 (sm_100a) implementation of the occlusions-4d
 * encoder / implicit-decoder hot path.
 *
 * The reference (basilevh/occlusions-4d) is pure Python: there is no FFI in it to bind.
 * Each entry point below therefore names the reference *Python* interface it replaces
 * (file:line under /root/reference); INTEGRATION.md shows the ctypes stub a maintainer
 * of the reference would add.  The host-side nn.Module mirror lives in
 * occlusions-4d_b200/o4d/.
 *
 * Conventions
 *   - plain pointers and sizes only; all data pointers are DEVICE pointers unless the
 *     name ends in _host; row-major contiguous fp32 activations with an explicit
 *     leading dimension (ld*, in elements) where a view can be strided;
 *   - neighbour / sample indices are int64 at this boundary (the reference hands out
 *     torch.int64) and int32 inside;
 *   - no host synchronisation inside compute calls; scratch memory is caller-owned (query the size with the matching
 *     *_workspace_bytes call).  One exception to "no allocation": a dense layer given an UN-packed fp32 weight on the
 *     tensor-core path (o4d_linear_f32, and the dense layers inside o4d_encoder_forward / the o4d_*_train entries)
 *     packs it into a stream-ordered temporary (cudaMallocAsync / cudaFreeAsync on `stream`: pool reuse after the
 *     first call, no device synchronisation).  The decoder keeps its images in the scene buffer, and
 *     o4d_linear_pack_f32 / o4d_linear_packed_f32 give every other caller the allocation-free form;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - every function returns 0 on success, <0 for an invalid argument (O4D_E_*), >0 for
 *     a cudaError_t raised by a launch; o4d_last_error() gives a message for the
 *     calling thread.  Nothing aborts.
 *   - re-entrant: no global mutable state besides the per-thread error string.
 */
#ifndef O4D_H_
#define O4D_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define O4D_OK 0
#define O4D_E_ARG (-1)      /* bad size / null pointer / unsupported combination */
#define O4D_E_WORKSPACE (-2) /* workspace too small */
#define O4D_E_UNSUPPORTED (-3)

#define O4D_MAX_K 16        /* largest neighbour count (pt_num_neighbors<=16) */
#define O4D_MAX_BLOCKS 16   /* largest n_blocks / down_blocks accepted */
#define O4D_MAX_OUT 64      /* largest d_out handled by o4d_output_activation_f32 */

const char* o4d_last_error(void);
/* ABI version of this header; bumped on any signature change. */
int o4d_abi_version(void);
/* 1 when the library was built with the tcgen05 (tensor-core) GEMM path. */
int o4d_has_tcgen05(void);

/* Diagnostics (not on the data path).  o4d_launch_count: kernels this library has launched
 * in the process so far.  o4d_profile_enable(1) makes every kernel family record CUDA
 * events on its launching stream; o4d_profile_read sums elapsed ms / algorithmic flops /
 * launch counts per family (0 dense layer, 1 kNN, 2 FPS, 3 attention gather+softmax,
 * 4 misc, 5 fused attention MLP) and synchronises on the recorded events;
 * o4d_profile_enable(0|1) also clears the records. */
uint64_t o4d_launch_count(void);
void o4d_profile_enable(int on);
int o4d_profile_read(int n_families, double* ms_out, double* flops_out, int64_t* count_out);

/* ------------------------------------------------------------------ kNN
 * Brute-force k nearest neighbours in 3-D, ascending (distance, index), distance =
 * fp32 ((dx*dx + dy*dy) + dz*dz) without FMA contraction.
 * Replaces kNN_torch, model/point_transformer_layer.py:76-99 (sqrt_dist = 0, squared
 * distances, the reference argsorts them) and my_knn_torch, utils/geometry.py:458-503
 * (sqrt_dist = 1: ordering and returned distances are Euclidean), and the neighbour
 * search of torch_cluster.knn at model/modules.py:142-146.
 *   query  (nq, ldq>=3)  xyz in the first three columns
 *   ref    (m,  ldr>=3)
 *   idx_out  (nq, k) int64           dist_out (nq, k) fp32 or NULL
 * Requires 1 <= k <= min(m, O4D_MAX_K). */
int o4d_knn_f32(const float* query, int64_t nq, int64_t ldq,
                const float* ref, int64_t m, int64_t ldr,
                int k, int sqrt_dist,
                int64_t* idx_out, float* dist_out, void* stream);

/* Both neighbour lists of the decoder from ONE scan of the reference cloud (same query, same cloud):
 *   idx_out  (nq, k)  the k nearest by SQUARED distance, ascending (d2, index)   -- kNN_torch, point_transformer_layer.py:76-99
 *   idx2_out (nq, k2), dist2_out (nq, k2)  the k2 < k nearest by EUCLIDEAN distance, ascending (sqrt(d2), index), with
 *            their distances                                                      -- my_knn_torch, utils/geometry.py:458-503
 * Results are identical to two o4d_knn_f32 calls (sqrt_dist 0 / 1), including the index tie-break where two different
 * squared distances round to the same root.  Requires 9 <= k <= O4D_MAX_K, k2 < k <= m.
 * workspace: o4d_knn_two_lists_workspace_bytes(nq, k, k2). */
size_t o4d_knn_two_lists_workspace_bytes(int64_t nq, int k, int k2);
int o4d_knn_two_lists_f32(const float* query, int64_t nq, int64_t ldq,
                          const float* ref, int64_t m, int64_t ldr,
                          int k, int k2,
                          int64_t* idx_out, int64_t* idx2_out, float* dist2_out,
                          void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ FPS
 * Farthest point sampling of one cloud, replaces torch_cluster.fps + torch.sort at
 * model/modules.py:133-135.  Start point = start_idx (0 = deterministic, the
 * reference's random_start=False); next = argmax of the running min squared distance
 * (first maximum on ties).  Output is SORTED ascending (what the caller consumes).
 *   xyz (n, ld>=3);  idx_sorted_out (n_out) int64;  order_out (n_out) int64 or NULL
 *   (selection order, for tests).  workspace: o4d_fps_workspace_bytes(n, n_out). */
size_t o4d_fps_workspace_bytes(int64_t n, int64_t n_out);
int o4d_fps_f32(const float* xyz, int64_t n, int64_t ld, int64_t n_out, int64_t start_idx,
                int64_t* idx_sorted_out, int64_t* order_out,
                void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ dense layer
 * C = post( pre(A) @ W^T + bias ) [+ R],  replaces every torch.nn.Linear on the path
 * (cuBLAS SGEMM in the reference).  A (rows, k) lda; W (n, k) row-major exactly as
 * nn.Linear stores it; bias (n) or NULL; R (rows, n) ldr or NULL (may alias C; A must NOT alias C);
 * flags: O4D_RELU_IN applies ReLU to A on load, O4D_RELU_OUT to the result before R.
 * precision: 0 = fp32 CUDA-core FMA, 1 = tcgen05 bf16x3 split (fp32-grade), 2 = tcgen05
 * single bf16 (fast, ~3e-3).  */
#define O4D_RELU_IN 1
#define O4D_RELU_OUT 2
int o4d_linear_f32(const float* A, int64_t rows, int64_t k, int64_t lda,
                   const float* W, const float* bias, int64_t n,
                   const float* R, int64_t ldr,
                   float* C, int64_t ldc, int flags, int precision, void* stream);

/* The same layer with a CALLER-OWNED packed weight: o4d_linear_f32 converts W into its tensor-core image (bf16 hi/lo,
 * shared-memory layout) in a stream-ordered temporary on every call (the only compute entry that allocates); callers that
 * reuse a weight pack it once and keep the image next to it, re-packing when the weight changes:
 *     bytes = o4d_linear_pack_bytes(rows, k, n, precision)   0: this shape / precision runs on CUDA cores, use o4d_linear_f32
 *     o4d_linear_pack_f32(W, n, k, ldw, packed, stream)       W (n, k) row-major with leading dimension ldw
 *     o4d_linear_packed_f32(A, ..., packed, ...)              same arguments and semantics as o4d_linear_f32
 * (o4d/ops.py keeps such a cache per (weight storage, version); the decoder keeps its images in the scene buffer.) */
size_t o4d_linear_pack_bytes(int64_t rows, int64_t k, int64_t n, int precision);
int o4d_linear_pack_f32(const float* W, int64_t n, int64_t k, int64_t ldw, void* packed, void* stream);
int o4d_linear_packed_f32(const float* A, int64_t rows, int64_t k, int64_t lda,
                          const void* packed, const float* bias, int64_t n,
                          const float* R, int64_t ldr,
                          float* C, int64_t ldc, int flags, int precision, void* stream);

/* ------------------------------------------------------------------ Fourier features
 * positional_encode, model/implicit.py:20-43 with base_frequency 0.1 (:184,:405):
 * points (n, d_in) -> out (n, d_in*(2*n_freq+1)) = [x, sin(w_0 x), cos(w_0 x), ...],
 * w_p = 2*pi*0.1*2^p, accurate sinf/cosf. */
int o4d_posenc_f32(const float* points, int64_t n, int d_in, int n_freq, float* out, void* stream);

/* ------------------------------------------------------------------ vector attention
 * One PointTransformerBlock, model/modules.py:18-67 wrapping PointTransformerLayer,
 * model/point_transformer_layer.py:116-183:
 *   z = x + W3 . attn( W1 x + b1 ; pos ; x2 ; pos2 ) + b3
 * self mode: x2 = NULL (keys/values are layer1(x) of the same cloud); cross mode: x2
 * (m, d2) raw abstract features, pos2 (m, 3).
 * Parameter table `p` (device pointers), in state_dict order:
 *   0 layer1.weight (d,d_in) 1 layer1.bias  2 to_q.weight (d,d)  3 to_k.weight (d,d2)
 *   4 to_v.weight (d,d2)  5 pos_mlp.0.weight (32,3)  6 pos_mlp.0.bias  7 pos_mlp.2.weight (d,32)
 *   8 pos_mlp.2.bias  9 attn_mlp.0.weight (2d,d)  10 attn_mlp.0.bias  11 attn_mlp.2.weight (d,2d)
 *   12 attn_mlp.2.bias  13 layer3.weight (d_out,d)  14 layer3.bias
 * d_in == d == d_out in every configuration the reference builds (model.py:89-105,
 * implicit.py:245-248); other shapes return O4D_E_UNSUPPORTED.
 * knn_idx_out: optional (n, k) int64 copy of the neighbour indices. */
#define O4D_PTBLOCK_NPARAMS 15
size_t o4d_pt_block_workspace_bytes(int64_t n, int64_t m, int d, int d2, int k);
int o4d_pt_block_forward(const float* const* p,
                         const float* x, int64_t n, int d, const float* pos, int64_t ldpos,
                         const float* x2, int64_t m, int d2, int64_t ldx2,
                         const float* pos2, int64_t ldpos2,
                         int k, int precision,
                         float* z, int64_t* knn_idx_out,
                         void* workspace, size_t workspace_bytes, void* stream);

/* Bare PointTransformerLayer.forward, model/point_transformer_layer.py:148-183 (what
 * o4d_pt_block_forward wraps with layer1/layer3).  `p` = its 11 tensors in state_dict
 * order: to_q.w, to_k.w, to_v.w, pos_mlp.0.{w,b}, pos_mlp.2.{w,b}, attn_mlp.0.{w,b},
 * attn_mlp.2.{w,b}.  out (n, d).  Workspace: o4d_pt_block_workspace_bytes. */
#define O4D_PTLAYER_NPARAMS 11
int o4d_pt_layer_forward(const float* const* p,
                         const float* x, int64_t n, int d, const float* pos, int64_t ldpos,
                         const float* x2, int64_t m, int d2, int64_t ldx2,
                         const float* pos2, int64_t ldpos2,
                         int k, int precision,
                         float* out, int64_t* knn_idx_out,
                         void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ fused residual block
 * ResnetBlockFC.forward, model/implicit.py:93-101 (ReLU, d_in == d_out so no shortcut layer):
 *     out = x + fc_1( relu( fc_0( relu(x) ) ) )
 * as ONE launch of the fused multi-layer kernel (csrc/mlp_chain.cu): the hidden activation never exists as fp32
 * in memory, both contractions run on tcgen05 (precision 1 = bf16x3, fp32-grade; 2 = single bf16 pass).
 * x (rows, d) ldx;  w0 (d_hidden, d), b0 (d_hidden) | NULL;  w1 (d, d_hidden), b1 (d) | NULL;  out (rows, d) ldo.
 * d and d_hidden must be multiples of 32 (else O4D_E_UNSUPPORTED: compose two o4d_linear_f32 calls); rows and
 * biases 16-byte aligned.  out must not alias x.  workspace: o4d_resblock_workspace_bytes (0 = unsupported shape). */
size_t o4d_resblock_workspace_bytes(int64_t rows, int d, int d_hidden);
int o4d_resblock_forward_f32(const float* x, int64_t rows, int d, int64_t ldx,
                             const float* w0, const float* b0, int d_hidden,
                             const float* w1, const float* b1,
                             float* out, int64_t ldo, int precision,
                             void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ down transition
 * DownTransition.forward, model/modules.py:113-163, one cloud:
 *   idx = sort(fps(pos, ceil(n/factor)));  nbr = knn(pos[idx] -> pos, k)
 *   y = relu( [LayerNorm]( W x + b ) ) on all n rows;  z = max over the k neighbour rows.
 * params: 0 mlp.0.weight (d_out,d_in) 1 mlp.0.bias 2 mlp.1.weight|NULL 3 mlp.1.bias|NULL
 * norm: 0 none, 1 LayerNorm(eps 1e-5).  ('batch' is unused by the released configs:
 * O4D_E_UNSUPPORTED.)   Outputs z (n_out, d_out), pos_out (n_out, 3),
 * fps_idx_out (n_out) int64 or NULL. */
#define O4D_DOWN_NPARAMS 4
size_t o4d_down_workspace_bytes(int64_t n, int d_in, int d_out, int factor, int k);
int o4d_down_forward(const float* const* p, const float* x, int64_t n, int d_in,
                     const float* pos, int64_t ldpos, int d_out, int factor, int k, int norm,
                     int64_t start_idx, int precision,
                     float* z, float* pos_out, int64_t* fps_idx_out,
                     void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ encoder
 * PointCompletionNetV3.forward, model/model.py:148-233, one cloud (B = 1 per call;
 * the host loops over the batch).  Supported = the released configuration:
 * enable_decoder=0, skip_connections=0, output_featurized=1, output_global_emb=1. */
typedef struct o4d_encoder_config {
    int32_t d_in;             /* 8 */
    int32_t d_feat;           /* 36 */
    int32_t down_blocks;      /* 3 */
    int32_t transition_factor;/* 3 */
    int32_t pt_num_neighbors; /* 14 / 16 */
    int32_t down_neighbors;   /* 12 */
    int32_t norm;             /* 0 none, 1 layer */
    int32_t abstract_levels;  /* 1 / 2 */
    int32_t global_dim;       /* 128 */
    int32_t precision;        /* see o4d_linear_f32 */
} o4d_encoder_config;
/* Parameter table order = state_dict order of the reference module (model.py:75-146):
 *   pre_mlp.0.{w,b}, pre_mlp.2.{w,b}, global_mlp.0.{w,b}, global_mlp.2.{w,b},
 *   abstract_skip_mlps.{j}.{w,b} (levels-1 of them),
 *   then per block i: PT block -> its 15 tensors; Down -> mlp.0.{w,b}[, mlp.1.{w,b}].
 * o4d_encoder_num_params() returns the expected count. */
int o4d_encoder_num_params(const o4d_encoder_config* cfg);
int64_t o4d_encoder_num_abstract(const o4d_encoder_config* cfg, int64_t n);
size_t o4d_encoder_workspace_bytes(const o4d_encoder_config* cfg, int64_t n);
/* pcl (n, d_in) -> abstract (M, 3 + d_feat*2^down_blocks), global (global_dim).
 * start_idx_host: HOST array of down_blocks FPS start indices (one per down transition,
 * each < the point count of its level) or NULL for all-zero = the reference's
 * deterministic test-time setting fps_random_start=False (eval/inference.py:59).
 * level_pos_out: optional array of down_blocks+1 device pointers receiving the
 * coordinates at every level ((n_l,3) each; "layer_coords", model.py:161-199). */
int o4d_encoder_forward(const o4d_encoder_config* cfg, const float* const* params,
                        const float* pcl, int64_t n, const int64_t* start_idx_host,
                        float* abstract_out, float* global_out, float* const* level_pos_out,
                        void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ decoder
 * LocalPclResnetFC.forward / do_forward_attention, model/implicit.py:271-445
 * (local_mode='attention', activation='relu', all cross layers type 'c', B = 1). */
typedef struct o4d_decoder_config {
    int32_t d_in;               /* 4 (x,y,z,t) */
    int32_t d_hidden;           /* 416 */
    int32_t d_out;              /* G */
    int32_t d_latent;           /* 416 = global + local */
    int32_t d_latent_local;     /* 288 */
    int32_t n_blocks;           /* 6 */
    int32_t pos_encoding_freqs; /* 8 */
    int32_t num_local_features; /* 8 */
    int32_t cross_attn_neighbors; /* 14 */
    int32_t cross_attn_layers;  /* 2 */
    int32_t precision;          /* see o4d_linear_f32 */
} o4d_decoder_config;
/* Parameter table order = state_dict order (implicit.py:142-150, 236-269):
 *   lin_in.{w,b}, lin_out.{w,b}, blocks.{i}.fc_0.{w,b}, blocks.{i}.fc_1.{w,b} (i<n_blocks),
 *   lin_z.{i}.{w,b} (i<n_blocks), pt_blocks.{j} -> 15 tensors each (j<cross_attn_layers). */
int o4d_decoder_num_params(const o4d_decoder_config* cfg);

/* Scene-constant state (K/V tables of every cross layer, global halves of lin_z, packed
 * abstract coordinates); recomputed once per (weights, scene), reused by every query
 * mini-batch of that scene (the reference recomputes all of it per mini-batch,
 * point_transformer_layer.py:171-172, implicit.py:417).
 *   scene buffer: caller-owned, o4d_decoder_scene_bytes(cfg, m) bytes. */
size_t o4d_decoder_scene_bytes(const o4d_decoder_config* cfg, int64_t m);
int o4d_decoder_prepare_scene(const o4d_decoder_config* cfg, const float* const* params,
                              const float* pcl_abstract, int64_t m, int64_t ld_abstract,
                              const float* feat_global,
                              void* scene, size_t scene_bytes, void* stream);
/* Next scene, same weights: o4d_decoder_prepare_scene also builds everything that depends on the weights only
 * (composite Qa weights, Wc, the K-concatenated lin_z folds, every packed tensor-core image).  When `scene` was prepared
 * before with the SAME cfg, params (unchanged values) and m, this entry rewrites only the scene-dependent parts (abstract
 * coordinates / features, K / V / Ka tables, the global half of lin_z) -- the per-frame call of a video
 * (eval/inference.py:195-212 runs the encoder once per frame and track). */
int o4d_decoder_update_scene(const o4d_decoder_config* cfg, const float* const* params,
                             const float* pcl_abstract, int64_t m, int64_t ld_abstract,
                             const float* feat_global, void* scene, size_t scene_bytes, void* stream);
size_t o4d_decoder_workspace_bytes(const o4d_decoder_config* cfg, int64_t nq, int64_t m);
/* query (nq, 4) -> out (nq, d_out), penult (nq, d_hidden) or NULL. */
int o4d_decoder_forward(const o4d_decoder_config* cfg, const float* const* params,
                        const void* scene, int64_t m,
                        const float* query, int64_t nq,
                        float* out, float* penult,
                        void* workspace, size_t workspace_bytes, void* stream);

/* Host-buffer convenience used for end-to-end timing and by non-torch callers:
 * the eval/inference.py:204-246 mini-batch loop in one call.  query_host / out_host are
 * (pinned or pageable) HOST buffers; device scratch is caller-owned
 * (o4d_decoder_run_host_device_bytes).  Copies in, runs every mini-batch of
 * `batch` queries, copies out, synchronises the stream before returning. */
size_t o4d_decoder_run_host_device_bytes(const o4d_decoder_config* cfg, int64_t batch, int64_t m);
int o4d_decoder_run_host(const o4d_decoder_config* cfg, const float* const* params,
                         const void* scene, int64_t m,
                         const float* query_host, int64_t nq, int64_t batch,
                         float* out_host,
                         void* device_scratch, size_t device_scratch_bytes, void* stream);

/* ================================================================== training path
 * The reference trains through torch.autograd over its eager graph (train.py:282-296 ->
 * pipeline.py:93-212 -> the forward()s above).  Here every gradient is an explicit kernel;
 * occlusions-4d_b200/o4d/autograd.py wraps the pairs below in torch.autograd.Function so
 * that loss.backward() in the unmodified train.py reaches them.  All gradient outputs are
 * OVERWRITTEN (autograd accumulates), indices are int64 as everywhere at this boundary. */

/* out = dy where y > 0 else 0 (ReLU backward from the layer OUTPUT; count elements). */
int o4d_relu_backward_f32(const float* dy, const float* y, int64_t count, float* out, void* stream);

/* Backward of o4d_linear_f32's  Y = pre(A) W^T + b  (flags: O4D_RELU_IN as in the forward; a
 * forward O4D_RELU_OUT is undone by the caller with o4d_relu_backward_f32 on dY first; the
 * residual R passes dY through unchanged).
 *   dA (rows, k) = (dY W) * [A > 0 if RELU_IN]      dW (n, k) = dY^T pre(A)      db (n) = sum_r dY
 * Any of dA / dW / db may be NULL.  W has leading dimension ldw (column slices of lin_z). */
size_t o4d_linear_backward_workspace_bytes(int64_t rows, int64_t k, int64_t n);
int o4d_linear_backward_f32(const float* A, int64_t rows, int64_t k, int64_t lda,
                            const float* W, int64_t ldw, int64_t n,
                            const float* dY, int64_t lddy, int flags,
                            float* dA, int64_t ldda, float* dW, int64_t lddw, float* db,
                            int precision, void* workspace, size_t workspace_bytes, void* stream);

/* Vector-attention core with saved activations (point_transformer_layer.py:174-179 given
 * q = to_q(x) (n,d), ktab = to_k(x2) (m,d), vtab = to_v(x2) (m,d) and nbr (n,k) int64):
 *   p8 = pos_mlp.0.{w,b}, pos_mlp.2.{w,b}, attn_mlp.0.{w,b}, attn_mlp.2.{w,b}
 * forward writes agg (n,d) and fills `saved` (o4d_attn_train_saved_bytes) with r, u, relu-hidden,
 * softmax weights and V+delta; backward consumes it and returns dq, dktab, dvtab and the eight
 * parameter gradients dp8 (same shapes as p8). */
size_t o4d_attn_train_saved_bytes(int64_t n, int d, int k);
size_t o4d_attn_backward_workspace_bytes(int64_t n, int d, int k);
int o4d_attn_forward_train(const float* const* p8, const float* q, const float* ktab, const float* vtab, int64_t m,
                           const float* pos, int64_t ldpos, const float* pos2, int64_t ldpos2,
                           const int64_t* nbr, int64_t n, int d, int k, int precision,
                           float* agg_out, void* saved, size_t saved_bytes, void* stream);
int o4d_attn_backward(const float* const* p8, const float* pos, int64_t ldpos, const float* pos2, int64_t ldpos2,
                      const int64_t* nbr, int64_t n, int64_t m, int d, int k, int precision,
                      const void* saved, size_t saved_bytes, const float* agg, const float* dagg,
                      float* dq, float* dktab, float* dvtab, float* const* dp8,
                      void* workspace, size_t workspace_bytes, void* stream);

/* Inverse-distance feature blend, implicit.py:337-339, and its gradient w.r.t. the abstract
 * features (distances / indices carry no gradient: coordinates are data).
 *   idx (n,k) int64, dist (n,k), feat (m,e) ldfeat -> out (n,e);   dfeat (m,e) lddfeat. */
int o4d_local_blend_f32(const int64_t* idx, const float* dist, const float* feat, int64_t ldfeat,
                        int64_t n, int k, int e, float* out, void* stream);
int o4d_local_blend_backward_f32(const int64_t* idx, const float* dist, const float* dout,
                                 int64_t n, int k, int e, int64_t m, float* dfeat, int64_t lddfeat, void* stream);

/* Neighbourhood max-pool of the down transition, modules.py:156-158, with the winning source
 * row per (output row, channel) recorded (first maximum), and the routing of dz back to it. */
int o4d_gather_max_f32(const float* y, int64_t ldy, const int64_t* nbr, int64_t n_out, int k, int d,
                       float* z, int32_t* arg_out, void* stream);
int o4d_gather_max_backward_f32(const float* dz, const int32_t* arg, int64_t n_out, int d, int64_t n_src,
                                float* dy, int64_t lddy, void* stream);

/* relu(LayerNorm(y)) of the CARLA down transition (modules.py:107-110), out of place, and its
 * gradients (dy, dgamma, dbeta). */
int o4d_layernorm_relu_f32(const float* y, int64_t rows, int d, const float* gamma, const float* beta,
                           float eps, float* out, void* stream);
int o4d_layernorm_relu_backward_f32(const float* y, const float* dout, int64_t rows, int d,
                                    const float* gamma, const float* beta, float eps,
                                    float* dy, float* dgamma, float* dbeta, void* stream);

/* Mean over points (model.py:189) and its gradient (dmean / rows broadcast to every row). */
int o4d_col_mean_f32(const float* x, int64_t rows, int d, float* out, void* stream);
int o4d_col_mean_backward_f32(const float* dmean, int64_t rows, int d, float* dx, void* stream);

/* ================================================================== inference driver pieces
 * (the caller's loop eval/inference.py:175-248 around the decoder.)
 *
 * Test-time query lattice of utils/geometry.py:1246-1262 ('grid' mode) generated on the device,
 * bit-identical to the numpy code: counts = ceil(cbrt(num / volume) * extent) per axis (computed by
 * o4d_grid_query_count, which returns the total and fills counts3), coordinate =
 * (index + 0.5) * (extent / count) + lo in fp32 without FMA, x slowest / z fastest, t = time_idx.
 *   out (total, 4) fp32 device, 16-byte aligned. */
int64_t o4d_grid_query_count(int64_t num_sample, const double* extent3, int32_t* counts3_out);
int o4d_grid_queries_f32(const int32_t* counts3, const double* extent3, const double* lo3, float time_idx,
                         float* out, void* stream);

/* Output squashing of eval/inference.py:218-243, in place on the decoder output (n, g):
 * col_ops_host[c] = 0 keep the logit, 1 sigmoid, 2 clamp to [0, 1]  (HOST array of g bytes). */
int o4d_output_activation_f32(float* out, int64_t n, int g, const uint8_t* col_ops_host, void* stream);

/* ================================================================== training-time query sampler pieces
 * (GuidedImplicitPointSampler, utils/geometry.py:578-1105, called once per frame
 * on the critical path of a training step, pipeline.py:176.)
 *
 * filter_air_solid_gap (utils/geometry.py:1164-1196) + select_safely (utils/geometry.py:1095-1105) in one call:
 * every candidate row cand[i, :d] (first three columns x, y, z) whose Euclidean distance to its nearest
 * target point is > radius is kept, in input order.
 *   num_select == 0: boolean-mask semantics -- out[0..n') = the kept rows, dist_out[0..n') their 1-NN
 *                    distances; out / dist_out must hold n rows.
 *   num_select  > 0: select_safely semantics -- out[j] = kept row (j mod n') for j < num_select (the
 *                    reference's repeated doubling); zeros when nothing was kept.  No host sync is needed.
 *   count_out (device int32) receives n' either way.  dist_out may be NULL.
 * Distances are fp32 sqrt((dx*dx + dy*dy) + dz*dz) without FMA contraction (torch.linalg.norm rounds the
 * same quantity within 2 ulp).  Workspace: o4d_filter_workspace_bytes(n). */
size_t o4d_filter_workspace_bytes(int64_t n);
int o4d_filter_air_solid_gap_f32(const float* cand, int64_t n, int d, int64_t ldc,
                                 const float* target, int64_t m, int64_t ldt, float radius,
                                 int64_t num_select, float* out, int64_t ldo, float* dist_out,
                                 int32_t* count_out, void* ws, size_t ws_bytes, void* stream);

/* filter_pcl_bounds_torch (utils/geometry.py:175-188): rows with lo[c] <= pcl[i, c] <= hi[c] for c = 0..2,
 * in input order; lo3_host / hi3_host are HOST arrays of three floats; out holds up to n rows; count_out
 * (device int32) receives the number kept.  Workspace: o4d_filter_workspace_bytes(n). */
int o4d_filter_bounds_f32(const float* pcl, int64_t n, int d, int64_t ld, const float* lo3_host,
                          const float* hi3_host, float* out, int64_t ldo, int32_t* count_out,
                          void* ws, size_t ws_bytes, void* stream);

/* ================================================================== implicit loss heads
 * (MyLosses.implicit_{density,color,segm,track}_loss, loss.py:50-194, applied to
 * every frame's decoder output in a training step, loss.py:226-246.)
 *
 * output (n, g) logits, target (n, 6) = (density, R, G, B, mark_track, segm) with -1 = not available.
 * color_mode: O4D_COLOR_RGB ('rgb' and 'rgb_nosigmoid': L1 on columns 1..3), O4D_COLOR_HSV (12-way hue CE on
 * columns 1..12 over rows with saturation and value >= 0.2 -- only when at least 16 such rows exist --, L1 on
 * saturation / value columns 13, 14), O4D_COLOR_BINS (9-way CE on columns 1..9); HSV targets are derived from
 * the RGB target as utils/utils.py:169-191 does.  semantic_classes = 0 disables the segmentation head (else it
 * reads the last semantic_classes columns), track_idx < 0 disables the tracking head.
 * Forward: losses4_out (device, 4 floats) = (loss_rgb, loss_dens, loss_segm, loss_track), each the mean over its
 * supervised rows (NaN when there are none, like torch); stats_out (device, O4D_LOSS_STATS doubles) keeps the
 * sums and counts for the backward call.  Deterministic (fixed reduction order).
 * Backward: doutput (n, g) = sum_h dlosses4[h] * d loss_h / d output (every column written; dlosses4 on device). */
#define O4D_COLOR_RGB 0
#define O4D_COLOR_HSV 1
#define O4D_COLOR_BINS 2
#define O4D_LOSS_STATS 12
size_t o4d_implicit_loss_workspace_bytes(int64_t n);
int o4d_implicit_loss_forward_f32(const float* output, int64_t n, int g, int64_t ldo,
                                  const float* target, int64_t ldt, int color_mode,
                                  int semantic_classes, int track_idx, float* losses4_out,
                                  double* stats_out, void* ws, size_t ws_bytes, void* stream);
int o4d_implicit_loss_backward_f32(const float* output, int64_t n, int g, int64_t ldo,
                                   const float* target, int64_t ldt, int color_mode,
                                   int semantic_classes, int track_idx, const double* stats,
                                   const float* dlosses4, float* doutput, int64_t lddo, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* O4D_H_ */
