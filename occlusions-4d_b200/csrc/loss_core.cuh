// Per-row arithmetic of the implicit loss heads (loss.cu), written so that the SAME source also compiles for
// the host: tests/test_host_core.py builds it with g++ (-ffp-contract=off) and checks it against the reference's
// golden losses and gradients on CPU, so the kernel arithmetic is verified without a GPU.  Reference lines:
// loss.py:50-194, utils/utils.py:169-191.
#pragma once
#include <math.h>
#include <stdint.h>
#include "../../include/o4d.h"

#if defined(__CUDACC__)
#define O4D_HD __host__ __device__ __forceinline__
#else
#define O4D_HD static inline
#endif

namespace o4d {
namespace lossk {

// one rounding per operation (no FMA contraction): intrinsics on the device, plain operators on the host
O4D_HD float rn_add(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    return a + b;
#endif
}
O4D_HD float rn_sub(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fsub_rn(a, b);
#else
    return a - b;
#endif
}
O4D_HD float rn_mul(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    return a * b;
#endif
}
O4D_HD float rn_div(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fdiv_rn(a, b);
#else
    return a / b;
#endif
}

constexpr int NSTAT = O4D_LOSS_STATS;
enum {
    S_DENS = 0, S_DENS_N = 1, S_L1A = 2, S_COLOR_N = 3, S_L1B = 4, S_CE = 5, S_CE_N = 6,
    S_SEGM = 7, S_SEGM_N = 8, S_TRACK = 9, S_TRACK_N = 10
};

struct Head {
    int g, color_mode, semantic_classes, track_idx;
};

O4D_HD float bce_logits(float x, float t) {
    // max(x, 0) - x t + log(1 + exp(-|x|))
    return fmaxf(x, 0.f) - x * t + log1pf(expf(-fabsf(x)));
}
O4D_HD float sigmoidf(float x) { return 1.f / (1.f + expf(-x)); }

// utils/utils.py:169-191 with every fp32 operation rounded separately; the arg-min tie goes to the first channel.
O4D_HD void rgb_to_hsv(float r, float g, float b, float& h, float& s, float& v) {
    const float eps = 1e-10f;
    float mx = fmaxf(r, fmaxf(g, b));
    float mn = r;
    int am = 0;
    if (g < mn) { mn = g; am = 1; }
    if (b < mn) { mn = b; am = 2; }
    const float span = rn_add(rn_sub(mx, mn), eps);
    float num, off;
    if (am == 0) { num = rn_sub(b, g); off = 180.f; }
    else if (am == 1) { num = rn_sub(r, b); off = 300.f; }
    else { num = rn_sub(g, r); off = 60.f; }
    h = rn_add(rn_div(rn_mul(60.f, num), span), off);
    s = rn_div(span, rn_add(mx, eps));
    v = mx;
}

// class of the colour target: 12 hue bins (hsv) or 6 hue bins + black / gray / white (bins); `vivid` says whether
// the hue head supervises this row (always true for bins).
O4D_HD int color_class(int color_mode, float h, float s, float v, bool& vivid) {
    if (color_mode == O4D_COLOR_HSV) {
        int c = (int)rintf(rn_mul(rn_div(h, 360.f), 12.f));
        if (c == 12) c = 0;
        vivid = (s >= 0.2f) && (v >= 0.2f);
        return c;
    }
    int c = (int)rintf(rn_mul(rn_div(h, 360.f), 6.f));
    if (c == 6) c = 0;
    const bool bland = (s < 0.3f) || (v < 0.3f);
    if (bland) c = v < 0.2f ? 6 : (v < 0.6f ? 7 : 8);
    vivid = true;
    return c;
}

O4D_HD float logsumexp(const float* x, int n) {
    float m = x[0];
    for (int c = 1; c < n; ++c) m = fmaxf(m, x[c]);
    float s = 0.f;
    for (int c = 0; c < n; ++c) s += expf(x[c] - m);
    return m + logf(s);
}

struct RowInfo {
    bool color, vivid, segm, track;
    int cls, tag;
    float sat, val;
};

O4D_HD RowInfo classify(const Head& hd, const float* t) {
    RowInfo ri;
    const bool solid = t[0] >= 0.1f;
    ri.color = solid && (t[1] >= 0.f);
    ri.track = hd.track_idx >= 0 && solid && (t[4] >= 0.f);
    ri.tag = (int)t[5];                          // .type(torch.int64) truncates
    // (a tag >= semantic_classes makes torch's cross_entropy raise; here the point is left out instead of read past z)
    ri.segm = hd.semantic_classes > 0 && ri.tag >= 0 && ri.tag < hd.semantic_classes;
    ri.vivid = false;
    ri.cls = 0;
    ri.sat = ri.val = 0.f;
    if (ri.color && hd.color_mode != O4D_COLOR_RGB) {
        float h;
        rgb_to_hsv(t[1], t[2], t[3], h, ri.sat, ri.val);
        ri.cls = color_class(hd.color_mode, h, ri.sat, ri.val, ri.vivid);
    }
    return ri;
}

// sums and counts of every head for one row (acc has NSTAT entries)
O4D_HD void row_accumulate(const Head& hd, const float* o, const float* t, double* acc) {
    const RowInfo ri = classify(hd, t);
    acc[S_DENS] += (double)bce_logits(o[0], t[0]);
    acc[S_DENS_N] += 1.0;
    if (ri.color) {
        acc[S_COLOR_N] += 1.0;
        if (hd.color_mode == O4D_COLOR_RGB) {
            acc[S_L1A] += (double)fabsf(o[1] - t[1]) + (double)fabsf(o[2] - t[2]) + (double)fabsf(o[3] - t[3]);
        } else if (hd.color_mode == O4D_COLOR_HSV) {
            acc[S_L1A] += (double)fabsf(o[13] - ri.sat);
            acc[S_L1B] += (double)fabsf(o[14] - ri.val);
            if (ri.vivid) {
                acc[S_CE] += (double)(logsumexp(o + 1, 12) - o[1 + ri.cls]);
                acc[S_CE_N] += 1.0;
            }
        } else {
            acc[S_CE] += (double)(logsumexp(o + 1, 9) - o[1 + ri.cls]);
            acc[S_CE_N] += 1.0;
        }
    }
    if (ri.segm) {
        const float* z = o + hd.g - hd.semantic_classes;
        acc[S_SEGM] += (double)(logsumexp(z, hd.semantic_classes) - z[ri.tag]);
        acc[S_SEGM_N] += 1.0;
    }
    if (ri.track) {
        acc[S_TRACK] += (double)bce_logits(o[hd.track_idx], t[4]);
        acc[S_TRACK_N] += 1.0;
    }
}

// torch: the mean of an empty selection is NaN
O4D_HD double mean_or_nan(double s, double c) { return c > 0.0 ? s / c : (double)NAN; }

// (rgb, dens, segm, track) as loss.py returns them, from the reduced sums and counts
O4D_HD void finalize_losses(const Head& hd, const double* st, float* losses4) {
        double rgb;
    if (hd.color_mode == O4D_COLOR_RGB) {
        rgb = mean_or_nan(st[S_L1A], 3.0 * st[S_COLOR_N]);
    } else if (hd.color_mode == O4D_COLOR_HSV) {
        const double hue = st[S_CE_N] >= 16.0 ? st[S_CE] / st[S_CE_N] / 2.0 : 0.0;   // loss.py:105-111
        rgb = (hue + mean_or_nan(st[S_L1A], st[S_COLOR_N]) + mean_or_nan(st[S_L1B], st[S_COLOR_N])) / 3.0;
    } else {
        rgb = mean_or_nan(st[S_CE], st[S_CE_N]) / 3.0;
    }
    losses4[0] = (float)rgb;
    losses4[1] = (float)mean_or_nan(st[S_DENS], st[S_DENS_N]);
    losses4[2] = hd.semantic_classes > 0 ? (float)mean_or_nan(st[S_SEGM], st[S_SEGM_N]) : 0.f;
    losses4[3] = hd.track_idx >= 0 ? (float)mean_or_nan(st[S_TRACK], st[S_TRACK_N]) : 0.f;
}

O4D_HD float sign0(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); }

// d(sum_h w_h * loss_h) / d output of one row; w = upstream gradients of (rgb, dens, segm, track)
O4D_HD void row_backward(const Head& hd, const float* o, const float* t, const double* stats, const float* w, float* d) {
    const RowInfo ri = classify(hd, t);
    for (int c = 0; c < hd.g; ++c) d[c] = 0.f;
    const float w_rgb = w[0], w_dens = w[1], w_segm = w[2], w_track = w[3];
    d[0] += w_dens / (float)stats[S_DENS_N] * (sigmoidf(o[0]) - t[0]);
    if (ri.color) {
        const float cn = (float)stats[S_COLOR_N];
        if (hd.color_mode == O4D_COLOR_RGB) {
            for (int c = 1; c <= 3; ++c) d[c] += w_rgb / (3.f * cn) * sign0(o[c] - t[c]);
        } else if (hd.color_mode == O4D_COLOR_HSV) {
            d[13] += w_rgb / 3.f / cn * sign0(o[13] - ri.sat);
            d[14] += w_rgb / 3.f / cn * sign0(o[14] - ri.val);
            if (ri.vivid && stats[S_CE_N] >= 16.0) {
                const float scale = w_rgb / 3.f / 2.f / (float)stats[S_CE_N];
                const float lse = logsumexp(o + 1, 12);
                for (int c = 0; c < 12; ++c) d[1 + c] += scale * (expf(o[1 + c] - lse) - (c == ri.cls ? 1.f : 0.f));
            }
        } else {
            const float scale = w_rgb / 3.f / (float)stats[S_CE_N];
            const float lse = logsumexp(o + 1, 9);
            for (int c = 0; c < 9; ++c) d[1 + c] += scale * (expf(o[1 + c] - lse) - (c == ri.cls ? 1.f : 0.f));
        }
    }
    if (ri.segm) {
        const int base = hd.g - hd.semantic_classes;
        const float scale = w_segm / (float)stats[S_SEGM_N];
        const float lse = logsumexp(o + base, hd.semantic_classes);
        for (int c = 0; c < hd.semantic_classes; ++c)
            d[base + c] += scale * (expf(o[base + c] - lse) - (c == ri.tag ? 1.f : 0.f));
    }
    if (ri.track) d[hd.track_idx] += w_track / (float)stats[S_TRACK_N] * (sigmoidf(o[hd.track_idx]) - t[4]);
}

}  // namespace lossk
}  // namespace o4d
