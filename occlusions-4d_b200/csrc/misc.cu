// Small bandwidth-bound kernels of the path: Fourier encoding, inverse-distance feature
// blend, neighbourhood max-pool, LayerNorm+ReLU, row gathers, column mean.
#include "o4d_common.cuh"
#include <cuda_bf16.h>
#include <math.h>

namespace o4d {

// positional_encode, model/implicit.py:20-43 with base_frequency 0.1 (:184,:405).
// out (n, d_in*(2F+1)): raw coords, then per power p: sin(w_p x) [d_in], cos(w_p x) [d_in].
// omega is formed in double like the reference's Python scalar and rounded to fp32 when it
// multiplies the fp32 tensor (torch scalar-multiply semantics); sinf/cosf are the
// accurate versions (arguments reach ~4e3 rad) -- this file is built without fast-math.
__global__ void posenc_kernel(const float* __restrict__ q, int64_t n, int d_in, int n_freq,
                              float* __restrict__ out) {
    const int width = d_in * (2 * n_freq + 1);
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n * width) return;
    const int64_t i = e / width;
    const int c = (int)(e % width);
    float v;
    if (c < d_in) {
        v = q[i * d_in + c];
    } else {
        const int t = c - d_in;
        const int p = t / (2 * d_in);
        const int r = t % (2 * d_in);
        const double omega_d = 0.1 * (double)(1 << p) * 3.141592653589793 * 2.0;
        const float omega = (float)omega_d;
        const float arg = q[i * d_in + (r % d_in)] * omega;
        v = (r < d_in) ? sinf(arg) : cosf(arg);
    }
    out[e] = v;
}

// The same features written straight into the activation image the fused multi-layer kernel reads (mlp_chain.cuh layout:
// per (128-row tile, 32-column chunk) [bf16 hi 8 KB][bf16 lo 8 KB], each [k/8][row/8][8 rows][8 values]; padding rows and
// columns zero).  Thread = (row, 8-column group): a warp's stores are 4 x 128 contiguous bytes per image half.
__global__ void posenc_image_kernel(const float* __restrict__ q, int64_t n, int d_in, int n_freq, uint8_t* __restrict__ img,
                                    int cpt, int64_t tiles) {
    const int width = d_in * (2 * n_freq + 1);
    const int64_t units = tiles * cpt * (128 * 4);
    for (int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u < units; u += (int64_t)gridDim.x * blockDim.x) {
        const int kc = (int)(u & 3);
        const int trow = (int)((u >> 2) % 128);
        const int64_t tc = u / (128 * 4);
        const int chunk = (int)(tc % cpt);
        const int64_t tile = tc / cpt;
        const int64_t i = tile * 128 + trow;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int c = chunk * 32 + kc * 8 + e;
            float x = 0.f;
            if (i < n && c < width) {
                if (c < d_in) {
                    x = q[i * d_in + c];
                } else {
                    const int t = c - d_in;
                    const int p = t / (2 * d_in);
                    const int r = t % (2 * d_in);
                    const float omega = (float)(0.1 * (double)(1 << p) * 3.141592653589793 * 2.0);
                    const float arg = q[i * d_in + (r % d_in)] * omega;
                    x = (r < d_in) ? sinf(arg) : cosf(arg);
                }
            }
            v[e] = x;
        }
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
            hi[e] = *reinterpret_cast<const uint32_t*>(&h);
            const float ha = __uint_as_float(hi[e] << 16), hb = __uint_as_float(hi[e] & 0xffff0000u);
            const __nv_bfloat162 l = __floats2bfloat162_rn(v[2 * e] - ha, v[2 * e + 1] - hb);
            lo[e] = *reinterpret_cast<const uint32_t*>(&l);
        }
        uint8_t* dst = img + ((size_t)tile * cpt + chunk) * 16384 + kc * 2048 + (trow >> 3) * 128 + (trow & 7) * 16;
        *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(dst + 8192) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
}

int posenc_image_launch(const float* q, int64_t n, int d_in, int n_freq, void* img, cudaStream_t st) {
    if (n == 0) return 0;
    O4D_REQUIRE(n_freq >= 0 && n_freq <= 24, "posenc: bad frequency count %d", n_freq);
    const int width = d_in * (2 * n_freq + 1);
    const int cpt = (width + 31) / 32;
    const int64_t tiles = cdiv(cdiv(n, 128), 2) * 2;          // image buffers cover an even number of row tiles
    const int64_t units = tiles * cpt * (128 * 4);
    int64_t blocks = cdiv(units, 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    posenc_image_kernel<<<(unsigned)blocks, 256, 0, st>>>(q, n, d_in, n_freq, (uint8_t*)img, cpt, tiles);
    O4D_LAUNCH_CHECK();
    return 0;
}

int posenc_launch(const float* q, int64_t n, int d_in, int n_freq, float* out, cudaStream_t st) {
    if (n == 0) return 0;
    O4D_REQUIRE(n_freq >= 0 && n_freq <= 24, "posenc: bad frequency count %d", n_freq);
    const int64_t total = n * d_in * (2 * n_freq + 1);
    posenc_kernel<<<(unsigned)cdiv(total, 256), 256, 0, st>>>(q, n, d_in, n_freq, out);
    O4D_LAUNCH_CHECK();
    return 0;
}

// model/implicit.py:337-339: w = 1/(dist+1e-4); w /= sum|w| (F.normalize p=1, eps 1e-12);
// out_i = sum_k w_k feat[idx_k].  One warp per query, lanes stride over channels.
// IMG: write the bf16 hi / lo activation image of the fused multi-layer kernel (mlp_chain.cuh layout, ceil(e / 32)
// chunks per 128-row tile, padding columns zero) instead of fp32 rows -- lin_z's operand, read by six layers.
template <bool IMG>
__global__ void __launch_bounds__(256)
local_blend_kernel(const int32_t* __restrict__ idx, const float* __restrict__ dist,
                   const float* __restrict__ feat, int64_t ldfeat, int64_t n, int k, int e,
                   float* __restrict__ out, int64_t ldout, uint8_t* __restrict__ img) {
    const int lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n) return;
    float w[O4D_MAX_K];
    int id[O4D_MAX_K];
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < O4D_MAX_K; ++j) {
        if (j < k) {
            w[j] = 1.0f / (dist[i * k + j] + 1e-4f);
            id[j] = idx[i * k + j];
            sum += fabsf(w[j]);
        }
    }
    const float denom = fmaxf(sum, 1e-12f);
#pragma unroll
    for (int j = 0; j < O4D_MAX_K; ++j)
        if (j < k) w[j] = w[j] / denom;
    const int cpt = (e + 31) / 32;
    // four channels per lane and step (16-byte loads of the gathered rows) when the rows allow it
    const bool vec = (e % 4 == 0) && (ldfeat % 4 == 0) && ((reinterpret_cast<uintptr_t>(feat) & 15) == 0) &&
                     (IMG || ((ldout % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0)));
    if (vec) {
        const int cend = IMG ? cpt * 32 : e;
        for (int c = lane * 4; c < cend; c += 128) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < e) {
#pragma unroll
                for (int j = 0; j < O4D_MAX_K; ++j) {
                    if (j < k) {
                        const float4 f = __ldg(reinterpret_cast<const float4*>(feat + (int64_t)id[j] * ldfeat + c));
                        acc.x = fmaf(w[j], f.x, acc.x); acc.y = fmaf(w[j], f.y, acc.y);
                        acc.z = fmaf(w[j], f.z, acc.z); acc.w = fmaf(w[j], f.w, acc.w);
                    }
                }
            }
            if (!IMG) {
                *reinterpret_cast<float4*>(out + i * ldout + c) = acc;
            } else {
                const __nv_bfloat162 h0 = __floats2bfloat162_rn(acc.x, acc.y), h1 = __floats2bfloat162_rn(acc.z, acc.w);
                const uint32_t u0 = *reinterpret_cast<const uint32_t*>(&h0), u1 = *reinterpret_cast<const uint32_t*>(&h1);
                const __nv_bfloat162 l0 = __floats2bfloat162_rn(acc.x - __uint_as_float(u0 << 16), acc.y - __uint_as_float(u0 & 0xffff0000u));
                const __nv_bfloat162 l1 = __floats2bfloat162_rn(acc.z - __uint_as_float(u1 << 16), acc.w - __uint_as_float(u1 & 0xffff0000u));
                uint8_t* dst = img + ((size_t)(i >> 7) * cpt + (c >> 5)) * 16384 + ((c & 31) >> 3) * 2048 + (((int)i & 127) >> 3) * 128 +
                               ((int)i & 7) * 16 + (c & 7) * 2;
                *reinterpret_cast<uint2*>(dst) = make_uint2(u0, u1);
                *reinterpret_cast<uint2*>(dst + 8192) = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
            }
        }
        return;
    }
    for (int c = lane; c < (IMG ? cpt * 32 : e); c += 32) {
        float acc = 0.f;
        if (c < e) {
#pragma unroll
            for (int j = 0; j < O4D_MAX_K; ++j) {
                if (j < k) acc = fmaf(w[j], feat[(int64_t)id[j] * ldfeat + c], acc);
            }
        }
        if (!IMG) {
            out[i * ldout + c] = acc;
        } else {
            const __nv_bfloat16 hi = __float2bfloat16_rn(acc);
            const __nv_bfloat16 lo = __float2bfloat16_rn(acc - __bfloat162float(hi));
            const uint32_t h = __bfloat16_as_ushort(hi), l = __bfloat16_as_ushort(lo);
            const uint32_t h1 = __shfl_down_sync(0xffffffffu, h, 1), l1 = __shfl_down_sync(0xffffffffu, l, 1);
            if ((lane & 1) == 0) {
                uint8_t* dst = img + ((size_t)(i >> 7) * cpt + (c >> 5)) * 16384 + (lane >> 3) * 2048 + (((int)i & 127) >> 3) * 128 +
                               ((int)i & 7) * 16 + (lane & 7) * 2;
                *reinterpret_cast<uint32_t*>(dst) = h | (h1 << 16);
                *reinterpret_cast<uint32_t*>(dst + 8192) = l | (l1 << 16);
            }
        }
    }
}

int local_blend_launch(const int32_t* idx, const float* dist, const float* feat, int64_t ldfeat,
                       int64_t n, int k, int e, float* out, int64_t ldout, cudaStream_t st, void* img) {
    if (n == 0) return 0;
    O4D_REQUIRE(k >= 1 && k <= O4D_MAX_K, "local blend: k=%d outside [1,%d]", k, O4D_MAX_K);
    if (img)
        local_blend_kernel<true><<<(unsigned)cdiv(n, 8), 256, 0, st>>>(idx, dist, feat, ldfeat, n, k, e, out, ldout, (uint8_t*)img);
    else
        local_blend_kernel<false><<<(unsigned)cdiv(n, 8), 256, 0, st>>>(idx, dist, feat, ldfeat, n, k, e, out, ldout, nullptr);
    O4D_LAUNCH_CHECK();
    return 0;
}

// model/modules.py:156-158: z_i = max_j y[nbr[i,j]].
__global__ void gather_max_kernel(const float* __restrict__ y, int64_t ldy, const int32_t* __restrict__ nbr,
                                  int64_t n_out, int k, int d, float* __restrict__ z) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_out * d) return;
    const int64_t i = e / d;
    const int c = (int)(e % d);
    float m = y[(int64_t)nbr[i * k] * ldy + c];
    for (int j = 1; j < k; ++j) m = fmaxf(m, y[(int64_t)nbr[i * k + j] * ldy + c]);
    z[e] = m;
}

int gather_max_launch(const float* y, int64_t ldy, const int32_t* nbr, int64_t n_out, int k, int d,
                      float* z, cudaStream_t st) {
    if (n_out == 0) return 0;
    gather_max_kernel<<<(unsigned)cdiv(n_out * d, 256), 256, 0, st>>>(y, ldy, nbr, n_out, k, d, z);
    O4D_LAUNCH_CHECK();
    return 0;
}

// torch.nn.LayerNorm(d) (biased variance, eps inside the sqrt) followed by ReLU, in place.
// One warp per row (model/modules.py:107-110 via :152).
__global__ void __launch_bounds__(256)
layernorm_relu_kernel(float* __restrict__ y, int64_t rows, int d, const float* __restrict__ gamma,
                      const float* __restrict__ beta, float eps) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= rows) return;
    float* row = y + r * d;
    float s = 0.f;
    for (int c = lane; c < d; c += 32) s += row[c];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    const float mean = s / (float)d;
    float v = 0.f;
    for (int c = lane; c < d; c += 32) {
        float t = row[c] - mean;
        v = fmaf(t, t, v);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    const float rstd = rsqrtf(v / (float)d + eps);
    for (int c = lane; c < d; c += 32) {
        float t = (row[c] - mean) * rstd * gamma[c] + beta[c];
        row[c] = fmaxf(t, 0.f);
    }
}

int layernorm_relu_launch(float* y, int64_t rows, int d, const float* gamma, const float* beta,
                          float eps, cudaStream_t st) {
    if (rows == 0) return 0;
    layernorm_relu_kernel<<<(unsigned)cdiv(rows, 8), 256, 0, st>>>(y, rows, d, gamma, beta, eps);
    O4D_LAUNCH_CHECK();
    return 0;
}

__global__ void gather_rows_kernel(const float* __restrict__ src, int64_t ldsrc, const int32_t* __restrict__ idx,
                                   int64_t n, int d, float* __restrict__ dst, int64_t lddst) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n * d) return;
    const int64_t i = e / d;
    const int c = (int)(e % d);
    dst[i * lddst + c] = src[(int64_t)idx[i] * ldsrc + c];
}

int gather_rows_launch(const float* src, int64_t ldsrc, const int32_t* idx, int64_t n, int d, float* dst,
                       int64_t lddst, cudaStream_t st) {
    if (n == 0) return 0;
    gather_rows_kernel<<<(unsigned)cdiv(n * d, 256), 256, 0, st>>>(src, ldsrc, idx, n, d, dst, lddst);
    O4D_LAUNCH_CHECK();
    return 0;
}

// torch.mean(x, dim=points), model/model.py:189.  One block per 32 channels; rows are
// split over the 8 warps and combined through shared memory (fixed order: deterministic).
__global__ void __launch_bounds__(256)
col_mean_kernel(const float* __restrict__ x, int64_t rows, int d, float* __restrict__ out) {
    __shared__ float part[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane;
    float s = 0.f;
    if (c < d)
        for (int64_t r = warp; r < rows; r += 8) s += x[r * d + c];
    part[warp][lane] = s;
    __syncthreads();
    if (warp == 0 && c < d) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += part[w][lane];
        out[c] = t / (float)rows;
    }
}

int col_mean_launch(const float* x, int64_t rows, int d, float* out, cudaStream_t st) {
    O4D_REQUIRE(rows >= 1, "mean over an empty cloud");
    col_mean_kernel<<<(unsigned)cdiv(d, 32), 256, 0, st>>>(x, rows, d, out);
    O4D_LAUNCH_CHECK();
    return 0;
}

__global__ void copy2d_kernel(const float* __restrict__ src, int64_t ldsrc, int64_t rows, int cols,
                              float* __restrict__ dst, int64_t lddst) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= rows * cols) return;
    const int64_t r = e / cols;
    const int c = (int)(e % cols);
    dst[r * lddst + c] = src[r * ldsrc + c];
}

int copy2d_launch(const float* src, int64_t ldsrc, int64_t rows, int cols, float* dst, int64_t lddst,
                  cudaStream_t st) {
    if (rows == 0 || cols == 0) return 0;
    copy2d_kernel<<<(unsigned)cdiv(rows * cols, 256), 256, 0, st>>>(src, ldsrc, rows, cols, dst, lddst);
    O4D_LAUNCH_CHECK();
    return 0;
}

__global__ void fill_col_kernel(float* __restrict__ dst, int64_t ld, int64_t rows, int col, float value) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < rows) dst[r * ld + col] = value;
}

int fill_col_launch(float* dst, int64_t ld, int64_t rows, int col, float value, cudaStream_t st) {
    if (rows == 0) return 0;
    fill_col_kernel<<<(unsigned)cdiv(rows, 256), 256, 0, st>>>(dst, ld, rows, col, value);
    O4D_LAUNCH_CHECK();
    return 0;
}

__global__ void widen_idx_kernel(const int32_t* __restrict__ in, int64_t count, int64_t* __restrict__ out) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < count) out[e] = in[e];
}

int widen_idx_launch(const int32_t* in, int64_t count, int64_t* out, cudaStream_t st) {
    if (count == 0) return 0;
    widen_idx_kernel<<<(unsigned)cdiv(count, 256), 256, 0, st>>>(in, count, out);
    O4D_LAUNCH_CHECK();
    return 0;
}

// ---- test-time query lattice, utils/geometry.py:1246-1262 ('grid' mode), generated on the device ----
// numpy: axis = (arange(c, dtype=float32) + 0.5) * (ext / c) + lo with Python-float scalars, i.e. every
// operation rounded to fp32 and no FMA; meshgrid(indexing='ij') + ravel => x slowest, z fastest.
__global__ void grid_queries_kernel(int cx, int cy, int cz, float sx, float sy, float sz, float lx, float ly, float lz,
                                    float t, float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)cx * cy * cz;
    if (i >= total) return;
    const int iz = (int)(i % cz);
    const int iy = (int)((i / cz) % cy);
    const int ix = (int)(i / ((int64_t)cz * cy));
    float4 q;
    q.x = __fadd_rn(__fmul_rn(__fadd_rn((float)ix, 0.5f), sx), lx);
    q.y = __fadd_rn(__fmul_rn(__fadd_rn((float)iy, 0.5f), sy), ly);
    q.z = __fadd_rn(__fmul_rn(__fadd_rn((float)iz, 0.5f), sz), lz);
    q.w = t;
    reinterpret_cast<float4*>(out)[i] = q;
}

// ---- output squashing of the inference loop, eval/inference.py:218-243, in place on (n, g) ----
struct ColOps {
    uint8_t op[O4D_MAX_OUT];   // 0 keep (logit), 1 sigmoid, 2 clamp to [0, 1]
};
__global__ void output_activation_kernel(float* __restrict__ out, int64_t n, int g, ColOps ops) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n * g) return;
    const int c = (int)(e % g);
    const float x = out[e];
    if (ops.op[c] == 1) out[e] = 1.0f / (1.0f + expf(-x));
    else if (ops.op[c] == 2) out[e] = fminf(fmaxf(x, 0.0f), 1.0f);
}

}  // namespace o4d

extern "C" int64_t o4d_grid_query_count(int64_t num_sample, const double* extent3, int32_t* counts3_out) {
    if (!extent3 || !counts3_out || num_sample < 1 || !(extent3[0] > 0 && extent3[1] > 0 && extent3[2] > 0)) return O4D_E_ARG;
    const double per_unit = cbrt((double)num_sample / (extent3[0] * extent3[1] * extent3[2]));   // geometry.py:1248
    int64_t total = 1;
    for (int a = 0; a < 3; ++a) {
        counts3_out[a] = (int32_t)ceil(per_unit * extent3[a]);
        total *= counts3_out[a];
    }
    return total;
}

extern "C" int o4d_grid_queries_f32(const int32_t* counts3, const double* extent3, const double* lo3, float time_idx,
                                    float* out, void* stream) {
    O4D_REQUIRE(counts3 && extent3 && lo3 && out, "o4d_grid_queries_f32: null pointer");
    O4D_REQUIRE(counts3[0] >= 1 && counts3[1] >= 1 && counts3[2] >= 1, "o4d_grid_queries_f32: empty lattice");
    O4D_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "o4d_grid_queries_f32: output must be 16-byte aligned");
    const int64_t total = (int64_t)counts3[0] * counts3[1] * counts3[2];
    // (ext / c) and lo are Python floats in the reference: rounded to fp32 when they meet the fp32 array
    o4d::grid_queries_kernel<<<(unsigned)o4d::cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(
        counts3[0], counts3[1], counts3[2], (float)(extent3[0] / counts3[0]), (float)(extent3[1] / counts3[1]),
        (float)(extent3[2] / counts3[2]), (float)lo3[0], (float)lo3[1], (float)lo3[2], time_idx, out);
    O4D_LAUNCH_CHECK();
    return 0;
}

extern "C" int o4d_output_activation_f32(float* out, int64_t n, int g, const uint8_t* col_ops_host, void* stream) {
    O4D_REQUIRE(out && col_ops_host && n >= 0 && g >= 1 && g <= O4D_MAX_OUT, "o4d_output_activation_f32: bad argument (g <= %d)",
                O4D_MAX_OUT);
    if (n == 0) return 0;
    o4d::ColOps ops;
    for (int c = 0; c < O4D_MAX_OUT; ++c) ops.op[c] = c < g ? col_ops_host[c] : 0;
    o4d::output_activation_kernel<<<(unsigned)o4d::cdiv(n * g, 256), 256, 0, (cudaStream_t)stream>>>(out, n, g, ops);
    O4D_LAUNCH_CHECK();
    return 0;
}

extern "C" int o4d_posenc_f32(const float* points, int64_t n, int d_in, int n_freq, float* out, void* stream) {
    O4D_REQUIRE(points && out && n >= 0 && d_in >= 1, "o4d_posenc_f32: bad argument");
    return o4d::posenc_launch(points, n, d_in, n_freq, out, (cudaStream_t)stream);
}
