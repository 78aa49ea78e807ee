"""GPU parity tests of the individual kernels, called through the C ABI (o4d.ops -> ctypes
-> libo4d.so) and checked against the CPU oracle / the reference's golden vectors.
Integer / index work is bit-exact; floating point carries its tolerance in the test."""
import math
import os

import numpy as np
import pytest
import torch

from o4d import ops
from oracle import cluster_ops, o4d_oracle as orc
from tests.test_oracle import load, relerr

pytestmark = pytest.mark.gpu
DEV = 'cuda'
TOL_FP32 = 2e-5     # precision 0: same fp32 math, different summation order
TOL_SPLIT = 2e-4    # precision 1: bf16x3 split operands, fp32 accumulate


def precisions():
    from o4d import _lib
    return [0, 1] if _lib.lib().o4d_has_tcgen05() else [0]


# ------------------------------------------------------------------------------ kNN

def test_knn_golden_vectors_bit_exact():
    g = load('knn_golden.npz')
    for tag in 'abc':
        q, r, k = g[tag + '_q'].to(DEV), g[tag + '_r'].to(DEV), int(g[tag + '_k'])
        assert torch.equal(ops.knn(q, r, k).cpu(), g[tag + '_idx_sq'])
        kl = g[tag + '_idx_eu'].shape[1]
        idx, dist = ops.knn(q, r, kl, sqrt_dist=True, return_dist=True)
        assert torch.equal(idx.cpu(), g[tag + '_idx_eu'])
        _, od = orc.knn_indices(g[tag + '_q'], g[tag + '_r'], kl, sqrt=True)
        assert torch.equal(dist.cpu(), od)  # IEEE sqrt of the same fp32 sum: bit-exact vs the oracle


def test_knn_duplicates_follow_the_tie_rule():
    g = load('knn_golden.npz')
    q, r, k = g['dup_q'], g['dup_r'], int(g['dup_k'])
    oi, od = orc.knn_indices(q, r, k)
    idx, dist = ops.knn(q.to(DEV), r.to(DEV), k, return_dist=True)
    assert torch.equal(idx.cpu(), oi) and torch.equal(dist.cpu(), od)
    # vs the reference: same distances, index differences only on exact ties
    assert torch.equal(torch.gather(g['dup_d2'], 1, idx.cpu()), torch.gather(g['dup_d2'], 1, g['dup_idx_sq']))


@pytest.mark.parametrize('nq,m,k,sqrt', [
    (1, 16, 16, False),        # k == m, single query
    (3, 5, 1, True),
    (257, 129, 8, True),       # 4 sub-lanes per query
    (1000, 2124, 14, False),   # 32 sub-lanes, CARLA abstract size
    (32768, 531, 14, False),   # decoder cross-attention shape
    (32768, 531, 8, True),     # decoder local-feature shape
    (4779, 14336, 12, False),  # down-transition shape
    (700001, 40, 3, False),    # one thread per query
])
def test_knn_matches_oracle(nq, m, k, sqrt):
    g = torch.Generator().manual_seed(nq + m)
    q = torch.rand(nq, 4, generator=g) * 10 - 5          # ld 4 like (x,y,z,t) queries
    r = torch.rand(m, 7, generator=g) * 10 - 5           # strided reference rows
    oi, od = orc.knn_indices(q[:, :3], r[:, :3], k, sqrt=sqrt)
    idx, dist = ops.knn(q.to(DEV), r.to(DEV), k, sqrt_dist=sqrt, return_dist=True)
    assert torch.equal(idx.cpu(), oi)
    assert torch.equal(dist.cpu(), od)


def test_non_finite_coordinates_stay_inside_the_tables():
    """NaN / Inf coordinates never pass a distance comparison: the emitted neighbour indices and FPS picks must still be
    valid row numbers (the gathers downstream index tables with them), and finite rows keep their exact result."""
    g = torch.Generator().manual_seed(3)
    q = torch.rand(500, 3, generator=g)
    ref = torch.rand(40, 3, generator=g)
    ref[5:] = float('nan')                                   # only 5 finite reference points, k = 14
    idx, dist = ops.knn(q.to(DEV), ref.to(DEV), 14, sqrt_dist=True, return_dist=True)
    assert int(idx.min()) >= 0 and int(idx.max()) < 40
    want = torch.cdist(q.double(), ref[:5].double()).sort(dim=1)
    assert torch.equal(idx[:, :5].cpu(), want.indices)
    assert bool(torch.isinf(dist[:, 5:]).all())
    i1, i2, _ = ops.knn_two_lists(q.to(DEV), ref.to(DEV), 14, 8)
    assert int(i1.min()) >= 0 and int(i1.max()) < 40 and int(i2.min()) >= 0 and int(i2.max()) < 40
    for n in (700, 3000):                                    # single-SM kernel and the cluster kernel
        cloud = torch.full((n, 3), float('nan'))
        picks = ops.fps(cloud.to(DEV), n // 3, 0)
        assert int(picks.min()) >= 0 and int(picks.max()) < n


def test_knn_encoder_self_shape_bit_exact():
    g = torch.Generator().manual_seed(14336)
    p = torch.rand(14336, 3, generator=g) * 10 - 5
    oi, _ = orc.knn_indices(p, p, 14)
    assert torch.equal(ops.knn(p.to(DEV), p.to(DEV), 14).cpu(), oi)
    assert bool((oi[:, 0] == torch.arange(14336)).all())  # every point is its own nearest neighbour


@pytest.mark.parametrize('nq,m,k,k2', [(5000, 531, 14, 8), (3000, 2124, 14, 8), (700, 300, 16, 15), (64, 20, 9, 3)])
def test_knn_two_lists_from_one_scan_equal_two_scans(nq, m, k, k2):
    """The decoder's two neighbour lists (squared-distance K_c, Euclidean K_l + distances) from one scan must be
    IDENTICAL to the two separate kernels, on clouds with exact duplicates (abstract_levels = 2 duplicates every
    level-2 position) and with many near-equal distances (points on a lattice: equal and nearly equal roots)."""
    g = torch.Generator().manual_seed(nq + m)
    ref = torch.rand(m, 3, generator=g) * 10 - 5
    ref[m // 2:m // 2 + m // 4] = ref[:m // 4]                          # exact duplicates
    ref[-(m // 5):] = torch.round(ref[-(m // 5):] * 2) / 2               # lattice points: many equal distances
    q = torch.rand(nq, 3, generator=g) * 10 - 5
    q[:nq // 4] = torch.round(q[:nq // 4] * 2) / 2 + 0.25
    idx, idx2, d2 = ops.knn_two_lists(q.to(DEV), ref.to(DEV), k, k2)
    want = ops.knn(q.to(DEV), ref.to(DEV), k)
    want2, wantd = ops.knn(q.to(DEV), ref.to(DEV), k2, sqrt_dist=True, return_dist=True)
    assert torch.equal(idx, want)
    assert torch.equal(idx2, want2)
    assert torch.equal(d2, wantd)
    # and against the oracle's canonical rule
    oi, od = orc.knn_indices(q, ref, k2, sqrt=True)
    assert torch.equal(idx2.cpu(), oi) and torch.equal(d2.cpu(), od)


def test_knn_rejects_bad_arguments():
    q = torch.rand(4, 3, device=DEV)
    with pytest.raises(RuntimeError, match='k='):
        ops.knn(q, q, 17)
    with pytest.raises(RuntimeError, match='k <= m'):
        ops.knn(q, q, 5)
    assert ops.knn(q[:0], q, 2).shape == (0, 2)


# ------------------------------------------------------------------------------ FPS

@pytest.mark.parametrize('n,n_out,start', [(76, 26, 0), (2048, 683, 0), (2048, 683, 77), (14336, 4779, 0),
                                           (1593, 531, 5), (20000, 300, 3),
                                           # 8-CTA cluster kernel (2048 < n <= 17066): 1, 2, 4 and 5 points per thread
                                           (2049, 683, 0), (4779, 1593, 11), (8192, 700, 8191), (14336, 4779, 123),
                                           (16384, 200, 1), (17000, 150, 9)])
def test_fps_matches_oracle(n, n_out, start):
    g = torch.Generator().manual_seed(n)
    p = torch.rand(n, 3, generator=g) * 10 - 5
    want = cluster_ops.fps_segment(p, n_out, start)
    if n <= 16384:
        got_sorted, got_order = ops.fps(p.to(DEV), n_out, start, return_order=True)
        assert torch.equal(got_order.cpu(), want)
    else:
        got_sorted = ops.fps(p.to(DEV), n_out, start)
    assert torch.equal(got_sorted.cpu(), torch.sort(want)[0])


def test_fps_zero_padded_cloud_repeats_like_argmax():
    p = torch.zeros(64, 3)
    p[:5] = torch.rand(5, 3) + 1.0
    want = cluster_ops.fps_segment(p, 22, 0)
    got_sorted, got_order = ops.fps(p.to(DEV), 22, 0, return_order=True)
    assert torch.equal(got_order.cpu(), want)
    assert torch.equal(got_sorted.cpu(), torch.sort(want)[0])


def test_fps_cluster_kernel_ties_and_single_sm_kernel_agree(monkeypatch):
    """Zero-padded duplicates (geometry.py:320-322) at a size the cluster kernel takes: exact ties must
    resolve to the lowest index, identically to the oracle."""
    g = torch.Generator().manual_seed(2)
    p = torch.rand(6000, 3, generator=g) * 8 - 4
    p[5000:] = 0.0
    p[100] = p[4000]
    want = cluster_ops.fps_segment(p, 2000, 0)
    got_sorted, got_order = ops.fps(p.to(DEV), 2000, 0, return_order=True)
    assert torch.equal(got_order.cpu(), want)
    assert torch.equal(got_sorted.cpu(), torch.sort(want)[0])


# ------------------------------------------------------------------------------ dense layer

@pytest.mark.parametrize('rows,k,n', [(1, 128, 128), (7, 3, 32), (300, 68, 416), (1000, 416, 832),
                                      (4097, 832, 416), (513, 36, 72), (2000, 416, 9), (129, 288, 416)])
def test_linear_matches_fp64(rows, k, n):
    g = torch.Generator().manual_seed(rows * 7 + n)
    a = torch.randn(rows, k, generator=g)
    w = torch.randn(n, k, generator=g) / math.sqrt(k)
    b = torch.randn(n, generator=g)
    r = torch.randn(rows, n, generator=g)
    want = torch.relu(torch.relu(a.double()) @ w.double().t() + b.double()) + r.double()
    plain = a.double() @ w.double().t()
    for prec in precisions():
        tol = TOL_FP32 if prec == 0 else TOL_SPLIT
        got = ops.linear(a.to(DEV), w.to(DEV), b.to(DEV), residual=r.to(DEV), relu_in=True, relu_out=True,
                         precision=prec)
        assert relerr(got.cpu().double(), want) < tol, (prec, rows, k, n)
        got2 = ops.linear(a.to(DEV), w.to(DEV), precision=prec)
        assert relerr(got2.cpu().double(), plain) < tol, (prec, rows, k, n)


@pytest.mark.parametrize('rows,d,dh', [(1024, 416, 416), (3001, 416, 416), (40000, 416, 416), (2500, 96, 160),
                                       (1500, 288, 832)])
def test_fused_residual_block_matches_fp64(rows, d, dh):
    """ResnetBlockFC.forward (implicit.py:93-101) as ONE launch of the fused multi-layer kernel: ragged last tile,
    several row tiles per CTA (40000 rows = 313 tiles on 148 CTAs), 1 / 2 / 4 n-tiles, image hand-off fc_0 -> fc_1."""
    g = torch.Generator().manual_seed(rows + d)
    x = torch.randn(rows, d, generator=g) * 3.0
    w0 = torch.randn(dh, d, generator=g) / math.sqrt(d)
    b0 = torch.randn(dh, generator=g)
    w1 = torch.randn(d, dh, generator=g) / math.sqrt(dh)
    b1 = torch.randn(d, generator=g)
    h = torch.relu(x.double()) @ w0.double().t() + b0.double()
    want = x.double() + torch.relu(h) @ w1.double().t() + b1.double()
    got = ops.resblock(x.to(DEV), w0.to(DEV), b0.to(DEV), w1.to(DEV), b1.to(DEV), precision=1)
    assert got is not None, 'shape should be inside the fused kernel'
    assert relerr(got.cpu().double(), want) < TOL_SPLIT, relerr(got.cpu().double(), want)
    # same launch twice: persistent-kernel state (barrier phases, counters) starts clean every time
    again = ops.resblock(x.to(DEV), w0.to(DEV), b0.to(DEV), w1.to(DEV), b1.to(DEV), precision=1)
    assert torch.equal(again, got)
    # the per-layer kernels must agree with it to rounding (same bf16x3 arithmetic, different tiling of the sum)
    two = ops.linear(ops.linear(x.to(DEV), w0.to(DEV), b0.to(DEV), relu_in=True, precision=1), w1.to(DEV), b1.to(DEV),
                     residual=x.to(DEV), relu_in=True, precision=1)
    assert relerr(got.cpu().double(), two.cpu().double()) < TOL_SPLIT
    assert ops.resblock(x.to(DEV)[:, :40].contiguous(), w0.to(DEV)[:, :40].contiguous(), b0.to(DEV), w1.to(DEV)[:40].contiguous(),
                        b1.to(DEV)[:40].contiguous(), precision=1) is None      # width 40: not a whole image chunk


def test_linear_packed_weight_cache_follows_in_place_updates():
    """o4d.ops keeps the tensor-core image of a weight (o4d_linear_pack_f32 / o4d_linear_packed_f32: no allocation or
    re-packing inside the compute call) keyed on (storage, version): an optimizer-style in-place update must re-pack."""
    g = torch.Generator().manual_seed(4)
    w = (torch.randn(416, 416, generator=g) / 20.0).to(DEV)
    a = torch.randn(2000, 416, generator=g).to(DEV)
    y1 = ops.linear(a, w, precision=1)
    entries = len(ops._PACKED)
    assert entries >= 1
    assert torch.equal(ops.linear(a, w, precision=1), y1) and len(ops._PACKED) == entries      # cache hit
    want = a.cpu().double() @ w.cpu().double().t()
    assert relerr(y1.cpu().double(), want) < TOL_SPLIT
    w.mul_(-2.0)                                                                               # version bump, same storage
    y2 = ops.linear(a, w, precision=1)
    assert len(ops._PACKED) == entries + 1
    assert relerr(y2.cpu().double(), -2.0 * want) < TOL_SPLIT
    # a shape below the tensor-core threshold packs nothing and still works
    assert relerr(ops.linear(a[:100], w, precision=1).cpu().double(), -2.0 * want[:100]) < TOL_SPLIT


def test_fused_residual_block_cta_pair_variant_in_a_subprocess():
    """O4D_CHAIN_PAIR=1 selects the cta_group::2 instantiation of the fused multi-layer kernel (different packed-weight
    format, cluster launch); the switch is read once per process, so the same parity test runs in a child process."""
    import subprocess
    import sys
    env = dict(os.environ, O4D_CHAIN_PAIR='1')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, '-m', 'pytest', os.path.join(root, 'tests', 'test_gpu_kernels.py'), '-q', '-m', 'gpu',
                          '-k', 'fused_residual_block_matches_fp64 and (3001 or 1500)', '-p', 'no:cacheprovider'],
                         env=env, cwd=root, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and '2 passed' in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_linear_strided_views_and_inplace_residual():
    g = torch.Generator().manual_seed(9)
    base = torch.randn(500, 291, generator=g)
    w = torch.randn(416, 288, generator=g) / 17.0
    want = base[:, 3:].double() @ w.double().t()
    for prec in precisions():
        got = ops.linear(base.to(DEV)[:, 3:], w.to(DEV), precision=prec)     # ld 291, offset 3 view
        assert relerr(got.cpu().double(), want) < (TOL_FP32 if prec == 0 else TOL_SPLIT)
        x = torch.randn(500, 416, generator=g).to(DEV)
        h = torch.randn(500, 416, generator=g).to(DEV)
        w2 = (torch.randn(416, 416, generator=g) / 20.0).to(DEV)
        ref = x.cpu().double() + torch.relu(h.cpu().double()) @ w2.cpu().double().t()
        # x += W relu(h): R aliases C (allowed); A must not alias C (other tiles still read it)
        out = ops.linear(h, w2, residual=x, relu_in=True, precision=prec, out=x)
        assert out.data_ptr() == x.data_ptr()
        assert relerr(out.cpu().double(), ref) < (TOL_FP32 if prec == 0 else TOL_SPLIT)


def test_posenc_matches_oracle():
    import o4d
    g = torch.Generator().manual_seed(1)
    q = torch.rand(5000, 4, generator=g) * 10 - 5
    q[:, 3] = 11.0
    want = orc.posenc(q, 8)
    got = o4d.positional_encode(q.to(DEV), 0.1, 8).cpu()
    assert got.shape == (5000, 68)
    assert torch.equal(got[:, :4], q)
    # accurate sinf/cosf on identical fp32 arguments: a few ulp of a value in [-1, 1]
    assert float((got - want).abs().max()) < 5e-7


# ------------------------------------------------------------------------------ attention / down

def _rand_block_state(d, d2, g):
    sd = {}
    def lin(name, n, k, bias=True):
        sd[name + '.weight'] = torch.randn(n, k, generator=g) / math.sqrt(k)
        if bias:
            sd[name + '.bias'] = torch.randn(n, generator=g) * 0.1
    lin('layer1', d, d); lin('layer2.to_q', d, d, False); lin('layer2.to_k', d, d2, False)
    lin('layer2.to_v', d, d2, False); lin('layer2.pos_mlp.0', 32, 3); lin('layer2.pos_mlp.2', d, 32)
    lin('layer2.attn_mlp.0', 2 * d, d); lin('layer2.attn_mlp.2', d, 2 * d); lin('layer3', d, d)
    return sd


BLOCK_ORDER = ['layer1.weight', 'layer1.bias', 'layer2.to_q.weight', 'layer2.to_k.weight', 'layer2.to_v.weight',
               'layer2.pos_mlp.0.weight', 'layer2.pos_mlp.0.bias', 'layer2.pos_mlp.2.weight',
               'layer2.pos_mlp.2.bias', 'layer2.attn_mlp.0.weight', 'layer2.attn_mlp.0.bias',
               'layer2.attn_mlp.2.weight', 'layer2.attn_mlp.2.bias', 'layer3.weight', 'layer3.bias']


@pytest.mark.parametrize('n,m,d,d2,k', [(300, 0, 36, 36, 14), (257, 0, 72, 72, 16), (500, 97, 80, 64, 6),
                                        (1000, 531, 416, 288, 14)])
def test_pt_block_matches_oracle(n, m, d, d2, k):
    g = torch.Generator().manual_seed(n + d)
    sd = _rand_block_state(d, d2, g)
    x = torch.randn(n, d, generator=g)
    pos = torch.rand(n, 3, generator=g) * 4
    x2 = torch.randn(m, d2, generator=g) if m else None
    pos2 = torch.rand(m, 3, generator=g) * 4 if m else None
    want = orc.pt_block(sd, '', x, pos, k, x2=x2, pos2=pos2)
    want_layer = orc.pt_layer(sd, 'layer2.', x, pos, x2, pos2, k)
    params = [sd[nm].to(DEV) for nm in BLOCK_ORDER]
    for prec in precisions():
        tol = 5e-5 if prec == 0 else 5e-4
        z, idx = ops.pt_block_forward(params, x.to(DEV), pos.to(DEV), None if x2 is None else x2.to(DEV),
                                      None if pos2 is None else pos2.to(DEV), k, prec, return_idx=True)
        oi, _ = orc.knn_indices(pos, pos if pos2 is None else pos2, k)
        assert torch.equal(idx.cpu(), oi)
        assert relerr(z.cpu(), want) < tol, prec
        y = ops.pt_layer_forward(params[2:13], x.to(DEV), pos.to(DEV), None if x2 is None else x2.to(DEV),
                                 None if pos2 is None else pos2.to(DEV), k, prec)
        assert relerr(y.cpu(), want_layer) < tol, prec


@pytest.mark.parametrize('norm', ['none', 'layer'])
def test_down_transition_matches_oracle(norm):
    g = torch.Generator().manual_seed(5)
    n, d_in, d_out, k = 1000, 36, 72, 12
    sd = {'mlp.0.weight': torch.randn(d_out, d_in, generator=g) / 6, 'mlp.0.bias': torch.randn(d_out, generator=g)}
    if norm == 'layer':
        sd['mlp.1.weight'] = torch.rand(d_out, generator=g) + 0.5
        sd['mlp.1.bias'] = torch.randn(d_out, generator=g) * 0.1
    x = torch.randn(n, d_in, generator=g)
    pos = torch.rand(n, 3, generator=g) * 10
    z, pos_sub, fidx = orc.down_transition(sd, '', x, pos, 3, k, norm)
    params = [sd['mlp.0.weight'].to(DEV), sd['mlp.0.bias'].to(DEV),
              sd['mlp.1.weight'].to(DEV) if norm == 'layer' else None,
              sd['mlp.1.bias'].to(DEV) if norm == 'layer' else None]
    gz, gp, gi = ops.down_forward(params, x.to(DEV), pos.to(DEV), d_out, 3, k, 1 if norm == 'layer' else 0,
                                  0, 0, return_idx=True)
    assert torch.equal(gi.cpu(), fidx)
    assert torch.equal(gp.cpu(), pos_sub)
    assert relerr(gz.cpu(), z) < TOL_FP32


# ------------------------------------------------------------------------------ inference driver pieces

@pytest.mark.parametrize('num,kind,bounds,mode', [(4096, 'greater', 5.0, 4), (524288, 'greater', 5.0, 4),
                                                  (4096, 'carla', 16.0, 4), (524288, 'carla', 16.0, 4),
                                                  (100000, 'carla', 20.0, 2)])
def test_device_grid_queries_bit_identical_to_numpy(num, kind, bounds, mode):
    from o4d import geometry
    want = geometry.sample_implicit_points_blind_numpy(num, -1.0, bounds, 3, kind, mode, 'grid')
    got = geometry.sample_implicit_points_blind_device(num, -1.0, bounds, 3, kind, mode, DEV)
    assert tuple(got.shape) == want.shape
    assert torch.equal(got.cpu(), torch.from_numpy(want))


def test_output_activation_matches_inference_loop():
    from o4d import geometry
    g = torch.Generator().manual_seed(4)
    for d_out, kw in ((9, dict(color_mode='rgb', track_mode='yes')),
                      (33, dict(color_mode='hsv', predict_segmentation=True, semantic_classes=13, track_mode='yes',
                                output_track_idx=15)),
                      (5, dict(color_mode='rgb_nosigmoid'))):
        x = torch.randn(1000, d_out, generator=g) * 4
        ops_ = geometry.inference_column_ops(d_out, **kw)
        want = x.clone()
        for c, op in enumerate(ops_):
            if op == geometry.SIGMOID:
                want[:, c] = torch.sigmoid(x[:, c])
            elif op == geometry.CLAMP01:
                want[:, c] = torch.clamp(x[:, c], 0.0, 1.0)
        got = geometry.output_activation(x.to(DEV), ops_).cpu()
        assert float((got - want).abs().max()) < 1e-6
        keep = [c for c, op in enumerate(ops_) if op == geometry.KEEP]
        assert torch.equal(got[:, keep], x[:, keep])
