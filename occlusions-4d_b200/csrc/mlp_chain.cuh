// Program description of the fused multi-layer MLP kernel (mlp_chain.cu) and the activation-image helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace o4d {
namespace mc {

constexpr int BM = 128;                    // rows per tile (UMMA M)
constexpr int BN_MAX = 208;                // widest n-tile (two TMEM accumulators of <= 256 columns, 4 smem stages)
constexpr int IMG_CHUNK_BYTES = 16384;     // one (128 rows x 32 columns) activation-image chunk: [bf16 hi 8 KB][bf16 lo 8 KB]
constexpr int MAX_OPS = 12;

// One dense layer:  Y = [A1 | A2] Wp^T + bias (+ R);  outputs: fp32 rows (out), and / or the activation image of
// Y or relu(Y) for the next layer.  A1 / A2 are activation images (k1c / k2c chunks of 32 columns per tile, all
// consumed); Wp is the pre-packed weight (tc_pack_launch format: per (n-tile, k-chunk) [hi slab][lo slab]).
struct Op {
    const uint8_t* a1;
    const uint8_t* a2;
    const uint8_t* w;
    const float* bias;
    float* out;
    const float* res;
    uint8_t* img;
    int64_t ldo, ldr;
    double k_alg;          // algorithmic K (for the flop count of the profile line)
    int k1c, k2c;
    int n, bn, ntiles;
    int img_cpt, img_relu, n_img;
};

struct Program {
    Op op[MAX_OPS];
    int nops;
    int tiles;
    int split;             // 1: bf16x3, 0: single bf16 pass
    int fence_gpu;         // 1: gpu-scope fence in front of the epilogue's proxy fence (diagnostic)
    int64_t rows;
};

}  // namespace mc

size_t act_image_bytes(int64_t rows, int cols);
int act_image_launch(const float* src, int64_t ld, int64_t rows, int cols, int relu, void* img, cudaStream_t st);
bool mlp_chain_layer_ok(int64_t k_total, int64_t n);
void mlp_chain_tiling(int64_t n, int* bn_out, int* ntiles_out);
int mlp_chain_launch(mc::Program& prog, cudaStream_t st);
bool mlp_chain_pair();
// packs a weight in the format the chain kernel of this process reads (single-CTA or pair)
int mlp_chain_pack_launch(const float* W, int64_t n, int64_t k, int64_t ldw, void* packed, cudaStream_t st);

}  // namespace o4d
