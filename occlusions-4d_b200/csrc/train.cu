// Training path: backward kernels of the dense layer and of the vector-attention core, the
// training-mode (materialising) attention forward, and the backward of the small
// gather / pool / normalisation stages.  The reference gets all of this from torch.autograd
// over its eager graph (train.py:282-296 -> pipeline.py:93-212); here every gradient is an
// explicit kernel behind the C ABI and torch.autograd only routes tensors between them
// (occlusions-4d_b200/o4d/autograd.py).
//
// Gradient formulas (row i, neighbour j = nbr[i, jj], channel c; d = width, s = 1/sqrt(d)):
//   forward   r = relu(Wp1 (p_i - p2_j) + bp1)           delta = Wp2 r + bp2
//             u = q_i - K_j + delta                      h = relu(Wa1 u + ba1)
//             a = Wa2 h + ba2                            w = softmax_j(a s)   (per channel)
//             vd = V_j + delta                           agg_i = sum_j w vd
//   backward  dvd = w * dagg_i                           da = w * dagg_i * (vd - agg_i) * s
//             dh = (da Wa2) * [h > 0]                    du = dh Wa1
//             dq_i = sum_j du     dK_j -= du     dV_j += dvd     ddelta = du + dvd
//             dr = (ddelta Wp2) * [r > 0]                dWp1 = dr^T (p_i - p2_j)
//   and for every dense layer Y = pre(A) W^T + b:  dA = (dY W) * [A > 0 if pre = relu],
//   dW = dY^T pre(A), db = column sums of dY.
#include "o4d_common.cuh"
#include <initializer_list>

namespace o4d {

constexpr int POS_HID_T = 32;

// ------------------------------------------------------------------------------ elementwise
__global__ void relu_mask_kernel(const float* __restrict__ dy, const float* __restrict__ y, int64_t n,
                                 float* __restrict__ out) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) out[e] = y[e] > 0.f ? dy[e] : 0.f;
}

// W (n, k) ldw -> Wt (k, n) contiguous
__global__ void transpose_kernel(const float* __restrict__ W, int64_t ldw, int n, int k, float* __restrict__ Wt) {
    __shared__ float tile[32][33];
    const int k0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int gn = n0 + r, gk = k0 + threadIdx.x;
        tile[r][threadIdx.x] = (gn < n && gk < k) ? W[(int64_t)gn * ldw + gk] : 0.f;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int gk = k0 + r, gn = n0 + threadIdx.x;
        if (gk < k && gn < n) Wt[(int64_t)gk * n + gn] = tile[threadIdx.x][r];
    }
}

// ------------------------------------------------------------------------------ weight gradient
// part[s][p][q] = sum over the rows of split s of dY[r, p] * pre(A[r, q]).   64 x 64 tile per
// CTA, 4 x 4 outputs per thread, rows walked 16 at a time through shared memory (both operand
// tiles are read along their contiguous dimension, so the loads coalesce without a transpose).
constexpr int WG_T = 64, WG_R = 16;

template <bool RELU>
__global__ void __launch_bounds__(256)
wgrad_simt_kernel(const float* __restrict__ dY, int64_t lddy, const float* __restrict__ A, int64_t lda, int64_t rows,
                  int n, int k, int64_t rows_per_split, float* __restrict__ part) {
    __shared__ float sY[WG_R][WG_T + 4];
    __shared__ float sA[WG_R][WG_T + 4];
    const int q0 = blockIdx.x * WG_T, p0 = blockIdx.y * WG_T;
    const int64_t r_lo = (int64_t)blockIdx.z * rows_per_split;
    const int64_t r_hi = min(rows, r_lo + rows_per_split);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int lr = threadIdx.x >> 4, lc = (threadIdx.x & 15) * 4;   // loader: row lr, columns lc..lc+3
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int64_t r0 = r_lo; r0 < r_hi; r0 += WG_R) {
        const int64_t gr = r0 + lr;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int gp = p0 + lc + e, gq = q0 + lc + e;
            float vy = 0.f, va = 0.f;
            if (gr < r_hi) {
                if (gp < n) vy = dY[gr * lddy + gp];
                if (gq < k) {
                    va = A[gr * lda + gq];
                    if (RELU) va = fmaxf(va, 0.f);
                }
            }
            sY[lr][lc + e] = vy;
            sA[lr][lc + e] = va;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < WG_R; ++r) {
            const float4 y4 = *reinterpret_cast<const float4*>(&sY[r][ty * 4]);
            const float4 a4 = *reinterpret_cast<const float4*>(&sA[r][tx * 4]);
            const float yv[4] = {y4.x, y4.y, y4.z, y4.w};
            const float av[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(yv[i], av[j], acc[i][j]);
        }
        __syncthreads();
    }
    float* dst = part + (int64_t)blockIdx.z * n * k;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gp = p0 + ty * 4 + i;
        if (gp >= n) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gq = q0 + tx * 4 + j;
            if (gq < k) dst[(int64_t)gp * k + gq] = acc[i][j];
        }
    }
}

// out[p * ldout + q] = sum_s part[s][p * k + q]   (fixed order: deterministic)
__global__ void reduce_partials_kernel(const float* __restrict__ part, int splits, int64_t count, int k,
                                       float* __restrict__ out, int64_t ldout) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= count) return;
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += part[(int64_t)z * count + e];
    out[(e / k) * ldout + (e % k)] = s;
}

// part[s][c] = sum over the rows of split s of dY[r, c]
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const float* __restrict__ dY, int64_t lddy, int64_t rows, int n, int64_t rows_per_split,
                      float* __restrict__ part) {
    __shared__ float red[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane;
    const int64_t r_lo = (int64_t)blockIdx.y * rows_per_split;
    const int64_t r_hi = min(rows, r_lo + rows_per_split);
    float s = 0.f;
    if (c < n)
        for (int64_t r = r_lo + warp; r < r_hi; r += 8) s += dY[r * lddy + c];
    red[warp][lane] = s;
    __syncthreads();
    if (warp == 0 && c < n) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[w][lane];
        part[(int64_t)blockIdx.y * n + c] = t;
    }
}

static int wgrad_splits(int64_t rows, int64_t n, int64_t k) {
    const int64_t tiles = cdiv(n, WG_T) * cdiv(k, WG_T);
    int64_t s = cdiv(148 * 4, tiles);
    const int64_t max_by_rows = cdiv(rows, 256);
    if (s > max_by_rows) s = max_by_rows;
    if (s > 128) s = 128;
    if (s < 1) s = 1;
    return (int)s;
}

static int colsum_splits(int64_t rows) {
    int64_t s = cdiv(rows, 2048);
    if (s > 64) s = 64;
    if (s < 1) s = 1;
    return (int)s;
}

// tcgen05 weight gradient (wgrad_tc.cu)
bool wgrad_tc_ok(int64_t rows, int64_t n, int64_t k);
int wgrad_tc_splits(int64_t rows, int64_t n, int64_t k);
bool wgrad_tc_aligned(const float* dY, int64_t lddy, const float* A, int64_t lda);
int wgrad_tc_launch(const float* dY, int64_t lddy, const float* A, int64_t lda, int64_t rows, int64_t n, int64_t k,
                    bool relu_a, int precision, float* part, float* bpart, int splits, cudaStream_t st);

static int part_splits(int64_t rows, int64_t n, int64_t k) {
    int s = wgrad_splits(rows, n, k);
    if (wgrad_tc_ok(rows, n, k)) {
        const int t = wgrad_tc_splits(rows, n, k);
        s = s > t ? s : t;
    }
    return s;
}

// the tcgen05 weight gradient also produces the bias partials (one per row split)
static int bias_splits(int64_t rows, int64_t n, int64_t k) {
    const int c = colsum_splits(rows), p = part_splits(rows, n, k);
    return c > p ? c : p;
}

size_t linear_bwd_ws_bytes(int64_t rows, int64_t k, int64_t n) {
    Arena a(nullptr, 0);
    a.get<float>((size_t)k * n);                                   // W^T
    a.get<float>((size_t)part_splits(rows, n, k) * n * k);         // weight-gradient partials
    a.get<float>((size_t)bias_splits(rows, n, k) * n);             // bias-gradient partials
    return a.off;
}

// Y = pre(A) W^T + b  ->  dA (rows, k), dW (n, k), db (n); every output optional, overwritten.
int linear_bwd_launch(const float* A, int64_t rows, int64_t k, int64_t lda, const float* W, int64_t ldw, int64_t n,
                      const float* dY, int64_t lddy, int flags, float* dA, int64_t ldda, float* dW, int64_t lddw,
                      float* db, int precision, void* ws, size_t ws_bytes, cudaStream_t st) {
    O4D_REQUIRE(dY && rows >= 0 && k >= 1 && n >= 1, "linear backward: bad argument");
    O4D_REQUIRE(lddy >= n && (!dA || (W && ldw >= k && ldda >= k)) && (!dW || (A && lda >= k && lddw >= k)),
                "linear backward: bad leading dimension / missing operand");
    O4D_REQUIRE(!(flags & O4D_RELU_IN) || A, "linear backward: relu_in needs the forward input");
    Arena a(ws, ws_bytes);
    float* wt = a.get<float>((size_t)k * n);
    const bool use_tc = precision != 0 && dW && wgrad_tc_ok(rows, n, k) && wgrad_tc_aligned(dY, lddy, A, lda);
    const int splits = use_tc ? wgrad_tc_splits(rows, n, k) : wgrad_splits(rows, n, k);
    float* part = a.get<float>((size_t)part_splits(rows, n, k) * n * k);
    const int csplits = colsum_splits(rows);
    float* bpart = a.get<float>((size_t)bias_splits(rows, n, k) * n);
    const bool bias_fused = use_tc && db;
    if (!a.ok || !ws) {
        set_error("linear backward: workspace too small (%zu < %zu)", ws_bytes, a.off);
        return O4D_E_WORKSPACE;
    }
    if (rows == 0) {
        if (dW) O4D_CUDA(cudaMemset2DAsync(dW, lddw * sizeof(float), 0, k * sizeof(float), n, st));
        if (db) O4D_CUDA(cudaMemsetAsync(db, 0, n * sizeof(float), st));
        return 0;
    }
    if (dA) {
        dim3 tb(32, 8), tg((unsigned)cdiv(k, 32), (unsigned)cdiv(n, 32));
        transpose_kernel<<<tg, tb, 0, st>>>(W, ldw, (int)n, (int)k, wt);
        O4D_LAUNCH_CHECK();
        // dA = dY (rows, n) . W (n, k) = dense layer with weight W^T (k, n)
        // a ReLU in front of the layer: its backward (dA = 0 where A <= 0) rides in the GEMM epilogue as a mask operand
        const bool mask = (flags & O4D_RELU_IN) != 0;
        O4D_TRY(linear_ldw_launch(dY, rows, n, lddy, wt, n, nullptr, k, mask ? A : nullptr, mask ? lda : 0, dA, ldda,
                                  mask ? O4D_MASK_RES : 0, precision, st));
    }
    if (dW) {
        ProfScope prof(PROF_LINEAR, 2.0 * (double)rows * (double)k * (double)n, st);
        if (use_tc) {
            O4D_TRY(wgrad_tc_launch(dY, lddy, A, lda, rows, n, k, (flags & O4D_RELU_IN) != 0, precision, part,
                                    bias_fused ? bpart : nullptr, splits, st));
        } else {
            int64_t rps = cdiv(cdiv(rows, splits), WG_R) * WG_R;
            dim3 grid((unsigned)cdiv(k, WG_T), (unsigned)cdiv(n, WG_T), (unsigned)splits);
            if (flags & O4D_RELU_IN)
                wgrad_simt_kernel<true><<<grid, 256, 0, st>>>(dY, lddy, A, lda, rows, (int)n, (int)k, rps, part);
            else
                wgrad_simt_kernel<false><<<grid, 256, 0, st>>>(dY, lddy, A, lda, rows, (int)n, (int)k, rps, part);
            O4D_LAUNCH_CHECK();
        }
        reduce_partials_kernel<<<(unsigned)cdiv(n * k, 256), 256, 0, st>>>(part, splits, n * k, (int)k, dW, lddw);
        O4D_LAUNCH_CHECK();
    }
    if (bias_fused) {
        reduce_partials_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(bpart, splits, n, (int)n, db, n);
        O4D_LAUNCH_CHECK();
    } else if (db) {
        const int64_t rps = cdiv(rows, csplits);
        dim3 grid((unsigned)cdiv(n, 32), (unsigned)csplits);
        colsum_partial_kernel<<<grid, 256, 0, st>>>(dY, lddy, rows, (int)n, rps, bpart);
        O4D_LAUNCH_CHECK();
        reduce_partials_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(bpart, csplits, n, (int)n, db, n);
        O4D_LAUNCH_CHECK();
    }
    return 0;
}

// ------------------------------------------------------------------------------ attention (training)
// r[row, t] = relu(Wp1[t] . (p_i - p2_j) + bp1[t]);  rel[row, :] = p_i - p2_j  (optional)
__global__ void __launch_bounds__(256)
posrelu64_kernel(const float* __restrict__ pos, int64_t ldpos, const float* __restrict__ pos2, int64_t ldpos2,
                 const int64_t* __restrict__ nbr, int64_t n_rows, int k, const float* __restrict__ wp1,
                 const float* __restrict__ bp1, float* __restrict__ R, float* __restrict__ rel) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_rows * POS_HID_T) return;
    const int64_t row = e / POS_HID_T;
    const int t = (int)(e % POS_HID_T);
    const int64_t i = row / k;
    const int64_t j = nbr[row];
    const float rx = pos[i * ldpos + 0] - pos2[j * ldpos2 + 0];
    const float ry = pos[i * ldpos + 1] - pos2[j * ldpos2 + 1];
    const float rz = pos[i * ldpos + 2] - pos2[j * ldpos2 + 2];
    if (R) {
        const float h = fmaf(wp1[t * 3 + 2], rz, fmaf(wp1[t * 3 + 1], ry, fmaf(wp1[t * 3 + 0], rx, bp1[t])));
        R[e] = fmaxf(h, 0.f);
    }
    if (rel && t < 3) rel[row * 3 + t] = (t == 0) ? rx : (t == 1 ? ry : rz);
}

// u = q_i - K_j + delta;  vd = V_j + delta (written over delta)
__global__ void __launch_bounds__(256)
attn_u_kernel(const float* __restrict__ q, const float* __restrict__ ktab, const float* __restrict__ vtab,
              const int64_t* __restrict__ nbr, int64_t n_rows, int k, int d, float* __restrict__ u,
              float* __restrict__ delta_vd) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_rows * d) return;
    const int64_t row = e / d;
    const int c = (int)(e % d);
    const int64_t i = row / k;
    const int64_t j = nbr[row];
    const float dl = delta_vd[e];
    u[e] = q[i * d + c] - ktab[j * d + c] + dl;
    delta_vd[e] = vtab[j * d + c] + dl;
}

// the same, four channels per thread (d % 4 == 0, 16-byte aligned rows): 16-byte loads / stores
__global__ void __launch_bounds__(256)
attn_u_vec_kernel(const float4* __restrict__ q, const float4* __restrict__ ktab, const float4* __restrict__ vtab,
                  const int64_t* __restrict__ nbr, int64_t n_rows, int k, int d4, float4* __restrict__ u,
                  float4* __restrict__ delta_vd) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_rows * d4) return;
    const int64_t row = e / d4;
    const int c = (int)(e - row * d4);
    const int64_t i = row / k;
    const int64_t j = nbr[row];
    const float4 dl = delta_vd[e];
    const float4 qv = __ldg(q + i * d4 + c), kv = __ldg(ktab + j * d4 + c), vv = __ldg(vtab + j * d4 + c);
    u[e] = make_float4(qv.x - kv.x + dl.x, qv.y - kv.y + dl.y, qv.z - kv.z + dl.z, qv.w - kv.w + dl.w);
    delta_vd[e] = make_float4(vv.x + dl.x, vv.y + dl.y, vv.z + dl.z, vv.w + dl.w);
}

// per (query, channel): w = softmax_j(a * s) written over the logits; agg = sum_j w * vd
template <int KMAX>
__global__ void __launch_bounds__(256)
softmax_agg_train_kernel(float* __restrict__ logits_w, const float* __restrict__ vd, int64_t n, int d, int k,
                         float scale, float* __restrict__ agg) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n * d) return;
    const int64_t i = e / d;
    const int c = (int)(e % d);
    float a[KMAX];
    float mx = -3.4e38f;
#pragma unroll
    for (int j = 0; j < KMAX; ++j)
        if (j < k) {
            a[j] = logits_w[(i * k + j) * d + c] * scale;
            mx = fmaxf(mx, a[j]);
        }
    float den = 0.f;
#pragma unroll
    for (int j = 0; j < KMAX; ++j)
        if (j < k) {
            a[j] = expf(a[j] - mx);
            den += a[j];
        }
    float num = 0.f;
#pragma unroll
    for (int j = 0; j < KMAX; ++j)
        if (j < k) {
            const float w = a[j] / den;
            logits_w[(i * k + j) * d + c] = w;
            num = fmaf(w, vd[(i * k + j) * d + c], num);
        }
    agg[e] = num;
}

// four channels per thread: the same arithmetic per channel, in the same order (bit-identical to the scalar kernel)
template <int KMAX>
__global__ void __launch_bounds__(128)
softmax_agg_train_vec_kernel(float4* __restrict__ logits_w, const float4* __restrict__ vd, int64_t n, int d4, int k,
                             float scale, float4* __restrict__ agg) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n * d4) return;
    const int64_t i = e / d4;
    const int c = (int)(e - i * d4);
    float4* lw = logits_w + i * k * d4 + c;
    const float4* vp = vd + i * k * d4 + c;
    float4 a[KMAX];
    float4 mx = make_float4(-3.4e38f, -3.4e38f, -3.4e38f, -3.4e38f);
#pragma unroll
    for (int j = 0; j < KMAX; ++j)
        if (j < k) {
            const float4 x = lw[(int64_t)j * d4];
            a[j] = make_float4(x.x * scale, x.y * scale, x.z * scale, x.w * scale);
            mx = make_float4(fmaxf(mx.x, a[j].x), fmaxf(mx.y, a[j].y), fmaxf(mx.z, a[j].z), fmaxf(mx.w, a[j].w));
        }
    float4 den = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < KMAX; ++j)
        if (j < k) {
            a[j] = make_float4(expf(a[j].x - mx.x), expf(a[j].y - mx.y), expf(a[j].z - mx.z), expf(a[j].w - mx.w));
            den.x += a[j].x; den.y += a[j].y; den.z += a[j].z; den.w += a[j].w;
        }
    float4 num = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < KMAX; ++j)
        if (j < k) {
            const float4 w = make_float4(a[j].x / den.x, a[j].y / den.y, a[j].z / den.z, a[j].w / den.w);
            const float4 v = vp[(int64_t)j * d4];
            lw[(int64_t)j * d4] = w;
            num.x = fmaf(w.x, v.x, num.x); num.y = fmaf(w.y, v.y, num.y);
            num.z = fmaf(w.z, v.z, num.z); num.w = fmaf(w.w, v.w, num.w);
        }
    agg[e] = num;
}

// da = w * dagg * (vd - agg) * s;  dvd = w * dagg
__global__ void __launch_bounds__(256)
softmax_agg_bwd_kernel(const float* __restrict__ w, const float* __restrict__ vd, const float* __restrict__ dagg,
                       const float* __restrict__ agg, int64_t n_rows, int k, int d, float scale,
                       float* __restrict__ da, float* __restrict__ dvd) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_rows * d) return;
    const int64_t row = e / d;
    const int c = (int)(e % d);
    const int64_t i = row / k;
    const float g = w[e] * dagg[i * d + c];
    dvd[e] = g;
    da[e] = g * (vd[e] - agg[i * d + c]) * scale;
}

// dq_i = sum_j du;  dK[nbr] -= du;  dV[nbr] += dvd;  ddelta = du + dvd (written over dvd)
template <int KMAX>
__global__ void __launch_bounds__(256)
attn_du_scatter_kernel(const float* __restrict__ du, float* __restrict__ dvd_ddelta, const int64_t* __restrict__ nbr,
                       int64_t n, int k, int d, float* __restrict__ dq, float* __restrict__ dktab,
                       float* __restrict__ dvtab) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n * d) return;
    const int64_t i = e / d;
    const int c = (int)(e % d);
    float s = 0.f;
    for (int j = 0; j < k; ++j) {
        const int64_t row = i * k + j;
        const int64_t jj = nbr[row];
        const float g = du[row * d + c];
        const float gv = dvd_ddelta[row * d + c];
        s += g;
        atomicAdd(dktab + jj * d + c, -g);
        atomicAdd(dvtab + jj * d + c, gv);
        dvd_ddelta[row * d + c] = g + gv;
    }
    dq[e] = s;
}

__global__ void __launch_bounds__(256)
softmax_agg_bwd_vec_kernel(const float4* __restrict__ w, const float4* __restrict__ vd, const float4* __restrict__ dagg,
                           const float4* __restrict__ agg, int64_t n_rows, int k, int d4, float scale,
                           float4* __restrict__ da, float4* __restrict__ dvd) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_rows * d4) return;
    const int64_t row = e / d4;
    const int c = (int)(e - row * d4);
    const int64_t i = row / k;
    const float4 wv = w[e], dg = __ldg(dagg + i * d4 + c), ag = __ldg(agg + i * d4 + c), v = vd[e];
    const float4 g = make_float4(wv.x * dg.x, wv.y * dg.y, wv.z * dg.z, wv.w * dg.w);
    dvd[e] = g;
    da[e] = make_float4(g.x * (v.x - ag.x) * scale, g.y * (v.y - ag.y) * scale, g.z * (v.z - ag.z) * scale,
                        g.w * (v.w - ag.w) * scale);
}

// four channels per thread; the table gradients go out as one 16-byte vector reduction per neighbour (red.global.add.v4.f32)
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(128)
attn_du_scatter_vec_kernel(const float4* __restrict__ du, float4* __restrict__ dvd_ddelta, const int64_t* __restrict__ nbr,
                           int64_t n, int k, int d4, float4* __restrict__ dq, float* __restrict__ dktab,
                           float* __restrict__ dvtab) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n * d4) return;
    const int64_t i = e / d4;
    const int c = (int)(e - i * d4);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0; j < k; ++j) {
        const int64_t row = i * k + j;
        const int64_t jj = nbr[row];
        const float4 g = du[row * d4 + c];
        const float4 gv = dvd_ddelta[row * d4 + c];
        s.x += g.x; s.y += g.y; s.z += g.z; s.w += g.w;
        red_add_v4(dktab + (jj * d4 + c) * 4, -g.x, -g.y, -g.z, -g.w);
        red_add_v4(dvtab + (jj * d4 + c) * 4, gv.x, gv.y, gv.z, gv.w);
        dvd_ddelta[row * d4 + c] = make_float4(g.x + gv.x, g.y + gv.y, g.z + gv.z, g.w + gv.w);
    }
    dq[e] = s;
}

static inline bool vec4_ok(int d, std::initializer_list<const void*> ptrs) {
    if (d % 4) return false;
    for (const void* p : ptrs)
        if ((uintptr_t)p % 16) return false;
    return true;
}

struct AttnSaved {
    float *r, *u, *h, *w, *vd;
    size_t bytes;
};

static AttnSaved attn_saved(int64_t n, int d, int k, void* base, size_t cap, bool* ok) {
    Arena a(base, cap);
    AttnSaved s;
    const size_t rows = (size_t)n * k;
    s.r = a.get<float>(rows * POS_HID_T);
    s.u = a.get<float>(rows * d);
    s.h = a.get<float>(rows * 2 * d);
    s.w = a.get<float>(rows * d);
    s.vd = a.get<float>(rows * d);
    s.bytes = a.off;
    if (ok) *ok = a.ok;
    return s;
}

struct AttnBwdWs {
    float *da, *dvd, *dh, *du, *dr, *rel;
    char* lin;
    size_t lin_bytes, bytes;
};

static AttnBwdWs attn_bwd_ws(int64_t n, int d, int k, void* base, size_t cap, bool* ok) {
    Arena a(base, cap);
    AttnBwdWs w;
    const size_t rows = (size_t)n * k;
    w.da = a.get<float>(rows * d);
    w.dvd = a.get<float>(rows * d);
    w.dh = a.get<float>(rows * 2 * d);
    w.du = a.get<float>(rows * d);
    w.dr = a.get<float>(rows * POS_HID_T);
    w.rel = a.get<float>(rows * 3);
    size_t lb = linear_bwd_ws_bytes((int64_t)rows, 2 * d, d);
    size_t t = linear_bwd_ws_bytes((int64_t)rows, d, 2 * d);
    lb = lb > t ? lb : t;
    t = linear_bwd_ws_bytes((int64_t)rows, POS_HID_T, d);
    lb = lb > t ? lb : t;
    t = linear_bwd_ws_bytes((int64_t)rows, 3, POS_HID_T);
    lb = lb > t ? lb : t;
    w.lin_bytes = lb;
    w.lin = a.get<char>(lb);
    w.bytes = a.off;
    if (ok) *ok = a.ok;
    return w;
}

// p8: pos_mlp.0.{w,b}, pos_mlp.2.{w,b}, attn_mlp.0.{w,b}, attn_mlp.2.{w,b}
int attn_train_forward(const float* const* p8, const float* q, const float* ktab, const float* vtab, int64_t m,
                       const float* pos, int64_t ldpos, const float* pos2, int64_t ldpos2, const int64_t* nbr,
                       int64_t n, int d, int k, int precision, float* agg, void* saved, size_t saved_bytes,
                       cudaStream_t st) {
    O4D_REQUIRE(p8 && q && ktab && vtab && pos && pos2 && nbr && agg, "attention (train): null pointer");
    O4D_REQUIRE(k >= 1 && k <= O4D_MAX_K && d >= 1 && m >= 1 && n >= 0, "attention (train): bad shape");
    if (n == 0) return 0;
    bool ok;
    AttnSaved s = attn_saved(n, d, k, saved, saved_bytes, &ok);
    if (!ok || !saved) {
        set_error("attention (train): saved-activation buffer too small (%zu < %zu)", saved_bytes, s.bytes);
        return O4D_E_WORKSPACE;
    }
    const int64_t rows = n * k;
    posrelu64_kernel<<<(unsigned)cdiv(rows * POS_HID_T, 256), 256, 0, st>>>(pos, ldpos, pos2, ldpos2, nbr, rows, k, p8[0],
                                                                           p8[1], s.r, nullptr);
    O4D_LAUNCH_CHECK();
    // delta = Wp2 r + bp2  (into the vd buffer)
    O4D_TRY(linear_launch(s.r, rows, POS_HID_T, POS_HID_T, p8[2], p8[3], d, nullptr, 0, s.vd, d, 0, precision, st));
    const bool vec = vec4_ok(d, {q, ktab, vtab, s.u, s.vd, s.w, agg});
    if (vec)
        attn_u_vec_kernel<<<(unsigned)cdiv(rows * (d / 4), 256), 256, 0, st>>>(
            reinterpret_cast<const float4*>(q), reinterpret_cast<const float4*>(ktab), reinterpret_cast<const float4*>(vtab), nbr,
            rows, k, d / 4, reinterpret_cast<float4*>(s.u), reinterpret_cast<float4*>(s.vd));
    else
        attn_u_kernel<<<(unsigned)cdiv(rows * d, 256), 256, 0, st>>>(q, ktab, vtab, nbr, rows, k, d, s.u, s.vd);
    O4D_LAUNCH_CHECK();
    O4D_TRY(linear_launch(s.u, rows, d, d, p8[4], p8[5], 2 * d, nullptr, 0, s.h, 2 * d, O4D_RELU_OUT, precision, st));
    O4D_TRY(linear_launch(s.h, rows, 2 * d, 2 * d, p8[6], p8[7], d, nullptr, 0, s.w, d, 0, precision, st));
    if (vec)
        softmax_agg_train_vec_kernel<O4D_MAX_K><<<(unsigned)cdiv(n * (d / 4), 128), 128, 0, st>>>(
            reinterpret_cast<float4*>(s.w), reinterpret_cast<const float4*>(s.vd), n, d / 4, k, (float)(1.0 / sqrt((double)d)),
            reinterpret_cast<float4*>(agg));
    else
        softmax_agg_train_kernel<O4D_MAX_K><<<(unsigned)cdiv(n * d, 256), 256, 0, st>>>(
            s.w, s.vd, n, d, k, (float)(1.0 / sqrt((double)d)), agg);
    O4D_LAUNCH_CHECK();
    return 0;
}

int attn_train_backward(const float* const* p8, const float* pos, int64_t ldpos, const float* pos2, int64_t ldpos2,
                        const int64_t* nbr, int64_t n, int64_t m, int d, int k, int precision, const void* saved,
                        size_t saved_bytes, const float* agg, const float* dagg, float* dq, float* dktab, float* dvtab,
                        float* const* dp8, void* ws, size_t ws_bytes, cudaStream_t st) {
    O4D_REQUIRE(p8 && pos && pos2 && nbr && saved && agg && dagg && dq && dktab && dvtab && dp8,
                "attention backward: null pointer");
    O4D_REQUIRE(k >= 1 && k <= O4D_MAX_K && d >= 1 && m >= 1 && n >= 0, "attention backward: bad shape");
    bool ok, ok2;
    AttnSaved s = attn_saved(n, d, k, const_cast<void*>(saved), saved_bytes, &ok);
    AttnBwdWs w = attn_bwd_ws(n, d, k, ws, ws_bytes, &ok2);
    if (!ok || !ok2 || !ws) {
        set_error("attention backward: buffer too small (saved %zu/%zu, workspace %zu/%zu)", saved_bytes, s.bytes,
                  ws_bytes, w.bytes);
        return O4D_E_WORKSPACE;
    }
    O4D_CUDA(cudaMemsetAsync(dktab, 0, (size_t)m * d * sizeof(float), st));
    O4D_CUDA(cudaMemsetAsync(dvtab, 0, (size_t)m * d * sizeof(float), st));
    const int64_t rows = n * k;
    if (n == 0) {
        const int64_t sizes[8] = {POS_HID_T * 3, POS_HID_T, (int64_t)d * POS_HID_T, d, 2LL * d * d, 2 * d, 2LL * d * d, d};
        for (int i = 0; i < 8; ++i)
            if (dp8[i]) O4D_CUDA(cudaMemsetAsync(dp8[i], 0, sizes[i] * sizeof(float), st));
        return 0;
    }
    const bool vec = vec4_ok(d, {s.w, s.vd, dagg, agg, w.da, w.dvd, w.du, dq, dktab, dvtab});
    if (vec)
        softmax_agg_bwd_vec_kernel<<<(unsigned)cdiv(rows * (d / 4), 256), 256, 0, st>>>(
            reinterpret_cast<const float4*>(s.w), reinterpret_cast<const float4*>(s.vd), reinterpret_cast<const float4*>(dagg),
            reinterpret_cast<const float4*>(agg), rows, k, d / 4, (float)(1.0 / sqrt((double)d)),
            reinterpret_cast<float4*>(w.da), reinterpret_cast<float4*>(w.dvd));
    else
        softmax_agg_bwd_kernel<<<(unsigned)cdiv(rows * d, 256), 256, 0, st>>>(s.w, s.vd, dagg, agg, rows, k, d,
                                                                             (float)(1.0 / sqrt((double)d)), w.da, w.dvd);
    O4D_LAUNCH_CHECK();
    // a = Wa2 h + ba2   (h is post-ReLU: relu(h) = h and [h > 0] = [pre-activation > 0])
    O4D_TRY(linear_bwd_launch(s.h, rows, 2 * d, 2 * d, p8[6], 2 * d, d, w.da, d, O4D_RELU_IN, w.dh, 2 * d, dp8[6], 2 * d,
                              dp8[7], precision, w.lin, w.lin_bytes, st));
    // h_pre = Wa1 u + ba1
    O4D_TRY(linear_bwd_launch(s.u, rows, d, d, p8[4], d, 2 * d, w.dh, 2 * d, 0, w.du, d, dp8[4], d, dp8[5], precision,
                              w.lin, w.lin_bytes, st));
    if (vec)
        attn_du_scatter_vec_kernel<<<(unsigned)cdiv(n * (d / 4), 128), 128, 0, st>>>(
            reinterpret_cast<const float4*>(w.du), reinterpret_cast<float4*>(w.dvd), nbr, n, k, d / 4,
            reinterpret_cast<float4*>(dq), dktab, dvtab);
    else
        attn_du_scatter_kernel<O4D_MAX_K><<<(unsigned)cdiv(n * d, 256), 256, 0, st>>>(w.du, w.dvd, nbr, n, k, d, dq, dktab,
                                                                                     dvtab);
    O4D_LAUNCH_CHECK();
    // delta = Wp2 r + bp2   (r post-ReLU)
    O4D_TRY(linear_bwd_launch(s.r, rows, POS_HID_T, POS_HID_T, p8[2], POS_HID_T, d, w.dvd, d, O4D_RELU_IN, w.dr,
                              POS_HID_T, dp8[2], POS_HID_T, dp8[3], precision, w.lin, w.lin_bytes, st));
    // r_pre = Wp1 (p_i - p2_j) + bp1
    posrelu64_kernel<<<(unsigned)cdiv(rows * POS_HID_T, 256), 256, 0, st>>>(pos, ldpos, pos2, ldpos2, nbr, rows, k, p8[0],
                                                                           p8[1], nullptr, w.rel);
    O4D_LAUNCH_CHECK();
    O4D_TRY(linear_bwd_launch(w.rel, rows, 3, 3, p8[0], 3, POS_HID_T, w.dr, POS_HID_T, 0, nullptr, 0, dp8[0], 3, dp8[1],
                              precision, w.lin, w.lin_bytes, st));
    return 0;
}

// ------------------------------------------------------------------------------ local feature blend
// forward (int64 indices): out_i = sum_j wn_j feat[idx_j],  wn = normalise_1(1 / (dist + 1e-4))
__global__ void __launch_bounds__(256)
local_blend64_kernel(const int64_t* __restrict__ idx, const float* __restrict__ dist, const float* __restrict__ feat,
                     int64_t ldfeat, int64_t n, int k, int e, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n) return;
    float w[O4D_MAX_K];
    int64_t id[O4D_MAX_K];
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < O4D_MAX_K; ++j)
        if (j < k) {
            w[j] = 1.0f / (dist[i * k + j] + 1e-4f);
            id[j] = idx[i * k + j];
            sum += fabsf(w[j]);
        }
    const float denom = fmaxf(sum, 1e-12f);
    for (int c = lane; c < e; c += 32) {
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < O4D_MAX_K; ++j)
            if (j < k) acc = fmaf(w[j] / denom, feat[id[j] * ldfeat + c], acc);
        out[i * e + c] = acc;
    }
}

// dfeat[idx_j] += wn_j dout_i   (dfeat zero-initialised by the launcher)
__global__ void __launch_bounds__(256)
local_blend_bwd_kernel(const int64_t* __restrict__ idx, const float* __restrict__ dist, const float* __restrict__ dout,
                       int64_t n, int k, int e, float* __restrict__ dfeat, int64_t lddfeat) {
    const int lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n) return;
    float w[O4D_MAX_K];
    int64_t id[O4D_MAX_K];
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < O4D_MAX_K; ++j)
        if (j < k) {
            w[j] = 1.0f / (dist[i * k + j] + 1e-4f);
            id[j] = idx[i * k + j];
            sum += fabsf(w[j]);
        }
    const float denom = fmaxf(sum, 1e-12f);
    for (int c = lane; c < e; c += 32) {
        const float g = dout[i * e + c];
#pragma unroll
        for (int j = 0; j < O4D_MAX_K; ++j)
            if (j < k) atomicAdd(dfeat + id[j] * lddfeat + c, (w[j] / denom) * g);
    }
}

// ------------------------------------------------------------------------------ neighbourhood max-pool
// z_i = max_j y[nbr[i, j]] with the winning source row (first maximum, like torch.max) recorded
__global__ void gather_max_arg_kernel(const float* __restrict__ y, int64_t ldy, const int64_t* __restrict__ nbr,
                                      int64_t n_out, int k, int d, float* __restrict__ z, int32_t* __restrict__ arg) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_out * d) return;
    const int64_t i = e / d;
    const int c = (int)(e % d);
    int64_t best = nbr[i * k];
    float m = y[best * ldy + c];
    for (int j = 1; j < k; ++j) {
        const int64_t s = nbr[i * k + j];
        const float v = y[s * ldy + c];
        if (v > m) {
            m = v;
            best = s;
        }
    }
    z[e] = m;
    arg[e] = (int32_t)best;
}

// dy[arg[i, c], c] += dz[i, c]   (dy zero-initialised by the launcher)
__global__ void scatter_arg_kernel(const float* __restrict__ dz, const int32_t* __restrict__ arg, int64_t n_out, int d,
                                   float* __restrict__ dy, int64_t lddy) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_out * d) return;
    const int c = (int)(e % d);
    atomicAdd(dy + (int64_t)arg[e] * lddy + c, dz[e]);
}

// ------------------------------------------------------------------------------ LayerNorm + ReLU
// out = relu(LN(y) * gamma + beta), y kept (training needs the pre-normalisation input)
__global__ void __launch_bounds__(256)
layernorm_relu_fwd_kernel(const float* __restrict__ y, int64_t rows, int d, const float* __restrict__ gamma,
                          const float* __restrict__ beta, float eps, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= rows) return;
    const float* row = y + r * d;
    float s = 0.f;
    for (int c = lane; c < d; c += 32) s += row[c];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    const float mean = s / (float)d;
    float v = 0.f;
    for (int c = lane; c < d; c += 32) {
        const float t = row[c] - mean;
        v = fmaf(t, t, v);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    const float rstd = rsqrtf(v / (float)d + eps);
    for (int c = lane; c < d; c += 32) out[r * d + c] = fmaxf((row[c] - mean) * rstd * gamma[c] + beta[c], 0.f);
}

// dy, dgamma, dbeta of out = relu(xhat * gamma + beta); dgamma / dbeta zero-initialised by the
// launcher, accumulated per block in shared memory then with one atomic per channel per block.
constexpr int LN_MAXD = 1024;
__global__ void __launch_bounds__(256)
layernorm_relu_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dout, int64_t rows, int d,
                          const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                          float* __restrict__ dy, float* __restrict__ dgamma, float* __restrict__ dbeta) {
    __shared__ float sg[LN_MAXD], sb[LN_MAXD];
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
        sg[c] = 0.f;
        sb[c] = 0.f;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (int64_t r = (int64_t)blockIdx.x * nwarp + warp; r < rows; r += (int64_t)gridDim.x * nwarp) {
        const float* row = y + r * d;
        float s = 0.f;
        for (int c = lane; c < d; c += 32) s += row[c];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        const float mean = s / (float)d;
        float v = 0.f;
        for (int c = lane; c < d; c += 32) {
            const float t = row[c] - mean;
            v = fmaf(t, t, v);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        const float rstd = rsqrtf(v / (float)d + eps);
        // pass 1: sums of dxhat and dxhat * xhat
        float s1 = 0.f, s2 = 0.f;
        for (int c = lane; c < d; c += 32) {
            const float xh = (row[c] - mean) * rstd;
            const float o = xh * gamma[c] + beta[c];
            const float g = o > 0.f ? dout[r * d + c] : 0.f;
            const float dxh = g * gamma[c];
            s1 += dxh;
            s2 = fmaf(dxh, xh, s2);
            atomicAdd(&sb[c], g);
            atomicAdd(&sg[c], g * xh);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, off);
            s2 += __shfl_xor_sync(0xffffffffu, s2, off);
        }
        const float m1 = s1 / (float)d, m2 = s2 / (float)d;
        for (int c = lane; c < d; c += 32) {
            const float xh = (row[c] - mean) * rstd;
            const float o = xh * gamma[c] + beta[c];
            const float g = o > 0.f ? dout[r * d + c] : 0.f;
            dy[r * d + c] = rstd * (g * gamma[c] - m1 - xh * m2);
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
        atomicAdd(dgamma + c, sg[c]);
        atomicAdd(dbeta + c, sb[c]);
    }
}

// dx[r, c] = dmean[c] / rows
__global__ void col_mean_bwd_kernel(const float* __restrict__ dmean, int64_t rows, int d, float* __restrict__ dx) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= rows * d) return;
    dx[e] = dmean[e % d] / (float)rows;
}

}  // namespace o4d

// ================================================================================== C ABI
using namespace o4d;

extern "C" int o4d_relu_backward_f32(const float* dy, const float* y, int64_t count, float* out, void* stream) {
    O4D_REQUIRE(dy && y && out && count >= 0, "o4d_relu_backward_f32: bad argument");
    if (count == 0) return 0;
    relu_mask_kernel<<<(unsigned)cdiv(count, 256), 256, 0, (cudaStream_t)stream>>>(dy, y, count, out);
    O4D_LAUNCH_CHECK();
    return 0;
}

extern "C" size_t o4d_linear_backward_workspace_bytes(int64_t rows, int64_t k, int64_t n) {
    if (rows < 0 || k < 1 || n < 1) return 0;
    return linear_bwd_ws_bytes(rows, k, n);
}

extern "C" int o4d_linear_backward_f32(const float* A, int64_t rows, int64_t k, int64_t lda, const float* W, int64_t ldw,
                                       int64_t n, const float* dY, int64_t lddy, int flags, float* dA, int64_t ldda,
                                       float* dW, int64_t lddw, float* db, int precision, void* workspace,
                                       size_t workspace_bytes, void* stream) {
    return linear_bwd_launch(A, rows, k, lda, W, ldw, n, dY, lddy, flags, dA, ldda, dW, lddw, db, precision, workspace,
                             workspace_bytes, (cudaStream_t)stream);
}

extern "C" size_t o4d_attn_train_saved_bytes(int64_t n, int d, int k) {
    if (n < 0 || d < 1 || k < 1) return 0;
    return attn_saved(n, d, k, nullptr, 0, nullptr).bytes;
}

extern "C" size_t o4d_attn_backward_workspace_bytes(int64_t n, int d, int k) {
    if (n < 0 || d < 1 || k < 1) return 0;
    return attn_bwd_ws(n, d, k, nullptr, 0, nullptr).bytes;
}

extern "C" int o4d_attn_forward_train(const float* const* p8, const float* q, const float* ktab, const float* vtab,
                                      int64_t m, const float* pos, int64_t ldpos, const float* pos2, int64_t ldpos2,
                                      const int64_t* nbr, int64_t n, int d, int k, int precision, float* agg_out,
                                      void* saved, size_t saved_bytes, void* stream) {
    return attn_train_forward(p8, q, ktab, vtab, m, pos, ldpos, pos2, ldpos2, nbr, n, d, k, precision, agg_out, saved,
                              saved_bytes, (cudaStream_t)stream);
}

extern "C" int o4d_attn_backward(const float* const* p8, const float* pos, int64_t ldpos, const float* pos2,
                                 int64_t ldpos2, const int64_t* nbr, int64_t n, int64_t m, int d, int k, int precision,
                                 const void* saved, size_t saved_bytes, const float* agg, const float* dagg, float* dq,
                                 float* dktab, float* dvtab, float* const* dp8, void* workspace, size_t workspace_bytes,
                                 void* stream) {
    return attn_train_backward(p8, pos, ldpos, pos2, ldpos2, nbr, n, m, d, k, precision, saved, saved_bytes, agg, dagg,
                               dq, dktab, dvtab, dp8, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int o4d_local_blend_f32(const int64_t* idx, const float* dist, const float* feat, int64_t ldfeat, int64_t n,
                                   int k, int e, float* out, void* stream) {
    O4D_REQUIRE(idx && dist && feat && out && n >= 0 && e >= 1 && ldfeat >= e, "o4d_local_blend_f32: bad argument");
    O4D_REQUIRE(k >= 1 && k <= O4D_MAX_K, "local blend: k=%d outside [1,%d]", k, O4D_MAX_K);
    if (n == 0) return 0;
    local_blend64_kernel<<<(unsigned)cdiv(n, 8), 256, 0, (cudaStream_t)stream>>>(idx, dist, feat, ldfeat, n, k, e, out);
    O4D_LAUNCH_CHECK();
    return 0;
}

extern "C" int o4d_local_blend_backward_f32(const int64_t* idx, const float* dist, const float* dout, int64_t n, int k,
                                            int e, int64_t m, float* dfeat, int64_t lddfeat, void* stream) {
    O4D_REQUIRE(idx && dist && dout && dfeat && n >= 0 && m >= 1 && e >= 1 && lddfeat >= e,
                "o4d_local_blend_backward_f32: bad argument");
    O4D_REQUIRE(k >= 1 && k <= O4D_MAX_K, "local blend: k=%d outside [1,%d]", k, O4D_MAX_K);
    cudaStream_t st = (cudaStream_t)stream;
    O4D_CUDA(cudaMemset2DAsync(dfeat, lddfeat * sizeof(float), 0, e * sizeof(float), m, st));
    if (n == 0) return 0;
    local_blend_bwd_kernel<<<(unsigned)cdiv(n, 8), 256, 0, st>>>(idx, dist, dout, n, k, e, dfeat, lddfeat);
    O4D_LAUNCH_CHECK();
    return 0;
}

extern "C" int o4d_gather_max_f32(const float* y, int64_t ldy, const int64_t* nbr, int64_t n_out, int k, int d, float* z,
                                  int32_t* arg_out, void* stream) {
    O4D_REQUIRE(y && nbr && z && arg_out && n_out >= 0 && k >= 1 && d >= 1 && ldy >= d, "o4d_gather_max_f32: bad argument");
    if (n_out == 0) return 0;
    gather_max_arg_kernel<<<(unsigned)cdiv(n_out * d, 256), 256, 0, (cudaStream_t)stream>>>(y, ldy, nbr, n_out, k, d, z,
                                                                                           arg_out);
    O4D_LAUNCH_CHECK();
    return 0;
}

extern "C" int o4d_gather_max_backward_f32(const float* dz, const int32_t* arg, int64_t n_out, int d, int64_t n_src,
                                           float* dy, int64_t lddy, void* stream) {
    O4D_REQUIRE(dz && arg && dy && n_out >= 0 && n_src >= 1 && d >= 1 && lddy >= d,
                "o4d_gather_max_backward_f32: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    O4D_CUDA(cudaMemset2DAsync(dy, lddy * sizeof(float), 0, d * sizeof(float), n_src, st));
    if (n_out == 0) return 0;
    scatter_arg_kernel<<<(unsigned)cdiv(n_out * d, 256), 256, 0, st>>>(dz, arg, n_out, d, dy, lddy);
    O4D_LAUNCH_CHECK();
    return 0;
}

extern "C" int o4d_layernorm_relu_f32(const float* y, int64_t rows, int d, const float* gamma, const float* beta,
                                      float eps, float* out, void* stream) {
    O4D_REQUIRE(y && gamma && beta && out && rows >= 0 && d >= 1, "o4d_layernorm_relu_f32: bad argument");
    if (rows == 0) return 0;
    layernorm_relu_fwd_kernel<<<(unsigned)cdiv(rows, 8), 256, 0, (cudaStream_t)stream>>>(y, rows, d, gamma, beta, eps, out);
    O4D_LAUNCH_CHECK();
    return 0;
}

extern "C" int o4d_layernorm_relu_backward_f32(const float* y, const float* dout, int64_t rows, int d,
                                               const float* gamma, const float* beta, float eps, float* dy,
                                               float* dgamma, float* dbeta, void* stream) {
    O4D_REQUIRE(y && dout && gamma && beta && dy && dgamma && dbeta && rows >= 0 && d >= 1 && d <= LN_MAXD,
                "o4d_layernorm_relu_backward_f32: bad argument (d <= %d)", LN_MAXD);
    cudaStream_t st = (cudaStream_t)stream;
    O4D_CUDA(cudaMemsetAsync(dgamma, 0, d * sizeof(float), st));
    O4D_CUDA(cudaMemsetAsync(dbeta, 0, d * sizeof(float), st));
    if (rows == 0) return 0;
    int64_t blocks = cdiv(rows, 8 * 16);   // ~16 rows per warp: few atomics per channel
    if (blocks > 148 * 4) blocks = 148 * 4;
    if (blocks < 1) blocks = 1;
    layernorm_relu_bwd_kernel<<<(unsigned)blocks, 256, 0, st>>>(y, dout, rows, d, gamma, beta, eps, dy, dgamma, dbeta);
    O4D_LAUNCH_CHECK();
    return 0;
}

extern "C" int o4d_col_mean_f32(const float* x, int64_t rows, int d, float* out, void* stream) {
    O4D_REQUIRE(x && out && d >= 1, "o4d_col_mean_f32: bad argument");
    return col_mean_launch(x, rows, d, out, (cudaStream_t)stream);
}

extern "C" int o4d_col_mean_backward_f32(const float* dmean, int64_t rows, int d, float* dx, void* stream) {
    O4D_REQUIRE(dmean && dx && rows >= 1 && d >= 1, "o4d_col_mean_backward_f32: bad argument");
    col_mean_bwd_kernel<<<(unsigned)cdiv(rows * d, 256), 256, 0, (cudaStream_t)stream>>>(dmean, rows, d, dx);
    O4D_LAUNCH_CHECK();
    return 0;
}
