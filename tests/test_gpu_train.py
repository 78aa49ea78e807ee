"""GPU tests of the training path (o4d/autograd.py -> backward kernels of libo4d.so).

Checker = torch.autograd over the CPU oracle in fp64 (oracle/o4d_oracle.py is plain torch and
differentiable), i.e. exactly what the reference's train.py gets from its eager graph.
Tolerance: max|ours - ref| / max|ref| per gradient tensor <= 1e-3 (north-star bar); the fp32 /
bf16x3 kernels deliver ~1e-5, the scatter-adds use float atomics (order-dependent rounding)."""
import math
import os

import numpy as np
import pytest
import torch

import o4d
from o4d import autograd as ag
from o4d import ops
from oracle import o4d_oracle as orc
from tests import configs

pytestmark = pytest.mark.gpu
DEV = 'cuda'
TOL = 1e-3


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def rel2(a, b):
    """Frobenius-norm relative error."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


# ReLU gradients are discontinuous at a zero pre-activation: a forward value that differs from the
# fp64 reference by one part in 1e5 (bf16x3) or 1e7 (fp32) flips the mask of the few elements that
# sit that close to zero, and each flip changes one row's gradient by one full term of its sum
# (measured on the B200: 3 of 852k masks flip in the 2048 x 416 case below, each worth ~4 % of the
# largest gradient entry).  PyTorch's own fp32 training has the same property.  Unit tests therefore
# build the fp64 reference with the masks the kernels actually used (mask-exact comparison, max
# norm); whole-model tests are max-norm at precision 0 and Frobenius-norm at precision 1.


# ------------------------------------------------------------------------------ dense layer
@pytest.mark.parametrize('rows,k,n,relu_in,relu_out,res,prec', [
    (1000, 72, 40, False, False, False, 0),
    (1000, 72, 40, True, False, True, 0),
    (777, 33, 129, False, True, False, 0),
    (1, 288, 128, False, True, False, 0),
    (4096, 416, 416, True, False, True, 1),
    (5000, 288, 416, False, False, True, 1),
    (3000, 832, 416, False, False, False, 1),
    (2048, 32, 416, False, True, False, 1),
    # tcgen05 weight gradient (tiled TMA loads, fused bias gradient): ragged last chunk, narrow / partial tiles, a row count
    # that leaves the last row split short, n not a multiple of 4, four q tiles, ReLU mask folded into the dX epilogue
    (4100, 36, 36, False, False, False, 1),
    (4500, 100, 70, True, False, False, 1),
    (6000, 832, 416, True, False, False, 1),
    (9001, 416, 832, False, True, False, 1),
    (0, 16, 8, False, False, False, 0),
])
def test_linear_backward_matches_autograd(rows, k, n, relu_in, relu_out, res, prec):
    g = torch.Generator().manual_seed(rows + k + n)
    x = torch.randn(rows, k, generator=g)
    w = torch.randn(n, k, generator=g) / math.sqrt(k)
    b = torch.randn(n, generator=g)
    r = torch.randn(rows, n, generator=g) if res else None
    dy = torch.randn(rows, n, generator=g)
    # reference: fp64 autograd
    xr, wr, br = (t.double().requires_grad_(True) for t in (x, w, b))
    rr = r.double().requires_grad_(True) if res else None
    # ours
    xo, wo, bo = (t.to(DEV).requires_grad_(True) for t in (x, w, b))
    ro = r.to(DEV).requires_grad_(True) if res else None
    yo = ag.linear(xo, wo, bo, residual=ro, relu_in=relu_in, relu_out=relu_out, precision=prec)
    yo.backward(dy.to(DEV))
    a = torch.relu(xr) if relu_in else xr
    y = a @ wr.t() + br
    if relu_out:      # the mask the kernel used (see the note above); identical except at |y| ~ 1e-6
        y = torch.where(yo.detach().cpu() > 0, y, torch.zeros_like(y))
    if res:
        y = y + rr
    y.backward(dy.double())
    if rows == 0:
        assert float(wo.grad.abs().max()) == 0.0 and float(bo.grad.abs().max()) == 0.0
        return
    assert rel(yo, y) < 1e-4
    assert rel(xo.grad, xr.grad) < 1e-4, 'dx'
    assert rel(wo.grad, wr.grad) < 1e-4, 'dW'
    assert rel(bo.grad, br.grad) < 1e-4, 'db'
    if res:
        assert torch.equal(ro.grad.cpu(), dy)


def test_linear_backward_c_abi_strided_operands():
    """o4d_linear_backward_f32 straight through the C ABI with operands that are column slices of wider buffers: leading
    dimensions that are multiples of 4 floats keep the tcgen05 weight gradient (tensor-map loads), an odd leading
    dimension or a misaligned base pointer must fall back to the CUDA-core kernel -- same results either way."""
    from o4d import _lib
    from o4d.ops import _ptr, workspace
    L = _lib.lib()
    rows, k, n = 4608, 96, 160
    g = torch.Generator().manual_seed(11)
    w = (torch.randn(n, k, generator=g) / math.sqrt(k)).to(DEV)
    for pad_x, pad_y, shift in ((8, 4, 0), (7, 4, 0), (8, 5, 0), (8, 4, 1)):
        xbuf = torch.randn(rows, k + pad_x + shift, generator=g).to(DEV)
        ybuf = torch.randn(rows, n + pad_y, generator=g).to(DEV)
        x, dy = xbuf[:, shift:shift + k], ybuf[:, :n]
        dw = torch.empty(n, k, device=DEV)
        db = torch.empty(n, device=DEV)
        dx = torch.empty(rows, k, device=DEV)
        ws = workspace(torch.device(DEV, 0), L.o4d_linear_backward_workspace_bytes(rows, k, n), slot=2)
        rc = L.o4d_linear_backward_f32(x.data_ptr(), rows, k, xbuf.shape[1], _ptr(w), k, n, dy.data_ptr(), ybuf.shape[1],
                                       ops.RELU_IN, _ptr(dx), k, _ptr(dw), k, _ptr(db), 1, _ptr(ws), ws.numel(),
                                       torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, 'o4d_linear_backward_f32')
        xr = torch.relu(x.double())
        assert rel(dw, dy.double().t() @ xr) < 1e-4, (pad_x, pad_y, shift)
        assert rel(db, dy.double().sum(0)) < 1e-4, (pad_x, pad_y, shift)
        assert rel(dx, torch.where(x.double() > 0, dy.double() @ w.double(), torch.zeros((), dtype=torch.float64, device=DEV))) < 1e-4


# ------------------------------------------------------------------------------ attention core
def _masked_relu(x, mask):
    return torch.relu(x) if mask is None else torch.where(mask, x, torch.zeros_like(x))


def _attn_reference(q, ktab, vtab, pos, pos2, nbr, p8, r_mask=None, h_mask=None):
    wp1, bp1, wp2, bp2, wa1, ba1, wa2, ba2 = p8
    d = q.shape[1]
    rel_pos = pos[:, None, :] - pos2[nbr]
    delta = _masked_relu(rel_pos @ wp1.t() + bp1, r_mask) @ wp2.t() + bp2
    a = q[:, None, :] - ktab[nbr] + delta
    a = _masked_relu(a @ wa1.t() + ba1, h_mask) @ wa2.t() + ba2
    w = torch.softmax(a / math.sqrt(d), dim=-2)
    return (w * (vtab[nbr] + delta)).sum(dim=-2)


@pytest.mark.parametrize('n,m,d,k,prec', [(300, 50, 40, 5, 0), (2000, 97, 64, 14, 1), (513, 513, 36, 16, 1),
                                          (1200, 200, 416, 14, 1)])
def test_attention_core_gradients(n, m, d, k, prec):
    g = torch.Generator().manual_seed(n + d)
    pos, pos2 = torch.rand(n, 3, generator=g) * 4, torch.rand(m, 3, generator=g) * 4
    q, ktab, vtab = (torch.randn(s, d, generator=g) for s in (n, m, m))
    shapes = [(32, 3), (32,), (d, 32), (d,), (2 * d, d), (2 * d,), (d, 2 * d), (d,)]
    p8 = [torch.randn(*s, generator=g) / math.sqrt(s[-1] if len(s) > 1 else 8) for s in shapes]
    nbr = torch.stack([torch.randperm(m, generator=g)[:k] for _ in range(n)])
    dout = torch.randn(n, d, generator=g)
    our_in = [t.to(DEV).requires_grad_(True) for t in [q, ktab, vtab] + p8]
    out = ag.attn_core(our_in[0], our_in[1], our_in[2], pos.to(DEV), pos2.to(DEV), nbr.to(DEV), k, prec, our_in[3:])
    # the ReLU masks the kernels used, read back from the saved-activation buffer
    # (layout of attn_saved in csrc/train.cu: r (rows,32) | u (rows,d) | h (rows,2d) | ..., 256-byte aligned)
    saved = out.grad_fn.saved_tensors[3]
    rows = n * k
    al = lambda b: (b + 255) // 256 * 256
    off_r, off_u = 0, al(rows * 32 * 4)
    off_h = off_u + al(rows * d * 4)
    r_mask = (saved[off_r:off_r + rows * 32 * 4].view(torch.float32).view(n, k, 32) > 0).cpu()
    h_mask = (saved[off_h:off_h + rows * 2 * d * 4].view(torch.float32).view(n, k, 2 * d) > 0).cpu()
    out.backward(dout.to(DEV))
    ref_in = [t.double().requires_grad_(True) for t in [q, ktab, vtab] + p8]
    out_ref = _attn_reference(ref_in[0], ref_in[1], ref_in[2], pos.double(), pos2.double(), nbr, ref_in[3:],
                              r_mask, h_mask)
    out_ref.backward(dout.double())
    assert rel(out, out_ref) < 1e-4
    names = ['dq', 'dK', 'dV', 'dWp1', 'dbp1', 'dWp2', 'dbp2', 'dWa1', 'dba1', 'dWa2', 'dba2']
    for name, a, b in zip(names, our_in, ref_in):
        if name == 'dba2':      # constant over neighbours: cancels in the softmax, gradient is ~0
            assert float(a.grad.abs().max()) < 1e-3 * float(ref_in[9].grad.abs().max())
            continue
        assert rel(a.grad, b.grad) < TOL, name


# ------------------------------------------------------------------------------ small stages
def test_local_blend_gather_max_layernorm_mean_gradients():
    g = torch.Generator().manual_seed(7)
    m, e, n, k = 90, 64, 700, 8
    feat = torch.randn(m, e, generator=g)
    idx = torch.stack([torch.randperm(m, generator=g)[:k] for _ in range(n)])
    dist = torch.rand(n, k, generator=g)
    dout = torch.randn(n, e, generator=g)
    fr = feat.double().requires_grad_(True)
    w = 1.0 / (dist.double() + 1e-4)
    w = w / w.abs().sum(-1, keepdim=True)
    ref = (w[..., None] * fr[idx]).sum(1)
    ref.backward(dout.double())
    fo = feat.to(DEV).requires_grad_(True)
    out = ag.LocalBlendFn.apply(fo, idx.to(DEV), dist.to(DEV))
    out.backward(dout.to(DEV))
    assert rel(out, ref) < 1e-5 and rel(fo.grad, fr.grad) < 1e-4

    # max-pool over neighbour rows (first maximum wins, like torch.max)
    y = torch.randn(300, 48, generator=g)
    y[5] = y[9]                                     # exact duplicates: tie goes to the earlier neighbour slot
    nbr = torch.stack([torch.randperm(300, generator=g)[:6] for _ in range(120)])
    dz = torch.randn(120, 48, generator=g)
    yr = y.double().requires_grad_(True)
    zr = yr[nbr].max(dim=1)[0]
    zr.backward(dz.double())
    yo = y.to(DEV).requires_grad_(True)
    z = ag.GatherMaxFn.apply(yo, nbr.to(DEV))
    z.backward(dz.to(DEV))
    assert torch.equal(z.cpu(), zr.detach().float())
    assert rel(yo.grad.sum(0), yr.grad.sum(0)) < 1e-4     # column totals are tie-rule independent
    untied = torch.ones(300, dtype=torch.bool)
    untied[[5, 9]] = False
    assert rel(yo.grad[untied.to(DEV)], yr.grad[untied]) < 1e-4

    # relu(LayerNorm)
    yy = torch.randn(1500, 144, generator=g) * 2 + 0.3
    gamma, beta = torch.rand(144, generator=g) + 0.5, torch.randn(144, generator=g) * 0.2
    do = torch.randn(1500, 144, generator=g)
    a, gr, br = (t.double().requires_grad_(True) for t in (yy, gamma, beta))
    o = torch.relu(torch.nn.functional.layer_norm(a, (144,), gr, br, 1e-5))
    o.backward(do.double())
    ao, go, bo = (t.to(DEV).requires_grad_(True) for t in (yy, gamma, beta))
    oo = ag.LayerNormReluFn.apply(ao, go, bo, 1e-5)
    oo.backward(do.to(DEV))
    assert rel(oo, o) < 1e-5
    assert rel(ao.grad, a.grad) < 1e-4 and rel(go.grad, gr.grad) < 1e-4 and rel(bo.grad, br.grad) < 1e-4

    # mean over points
    x = torch.randn(531, 288, generator=g)
    dm = torch.randn(288, generator=g)
    xo = x.to(DEV).requires_grad_(True)
    mo = ag.ColMeanFn.apply(xo)
    mo.backward(dm.to(DEV))
    assert rel(mo, x.double().mean(0)) < 1e-5
    assert rel(xo.grad, (dm.double() / 531)[None].expand(531, -1)) < 1e-6


# ------------------------------------------------------------------------------ whole model
def _oracle_grads(cfg, enc, dec, pcl, query, w_out, w_pen):
    sd_e = {k: v.detach().cpu().double().requires_grad_(True) for k, v in enc.state_dict().items()}
    sd_d = {k: v.detach().cpu().double().requires_grad_(True) for k, v in dec.state_dict().items()}
    a, g = orc.encoder_forward(sd_e, cfg['pcl_args'], pcl.double())
    out, pen = orc.decoder_forward(sd_d, cfg['implicit_args'], query.double(), a, g)
    loss = (out * w_out.double()).sum() + (pen * w_pen.double()).sum()
    loss.backward()
    return out.detach(), float(loss.detach()), sd_e, sd_d


@pytest.mark.parametrize('cfg', [configs.TINY_GREATER, configs.TINY_CARLA], ids=['tiny_greater', 'tiny_carla'])
@pytest.mark.parametrize('prec', [0, 1])
def test_model_gradients_match_oracle_autograd(cfg, prec):
    """loss.backward() through the nn.Module API (encoder -> decoder, the pipeline.py:93-212 data
    flow) fills every parameter's .grad with what torch.autograd gives on the reference graph."""
    enc, dec = configs.build_modules(cfg, DEV)
    enc.train()
    dec.train()
    enc.o4d_precision = dec.o4d_precision = prec
    pcl = configs.synthetic_cloud(cfg)
    query = configs.synthetic_queries(cfg)
    g = torch.Generator().manual_seed(3)
    w_out = torch.randn(query.shape[0], cfg['implicit_args']['d_out'], generator=g)
    w_pen = torch.randn(query.shape[0], cfg['implicit_args']['d_hidden'], generator=g) * 0.05
    out_ref, loss_ref, sd_e, sd_d = _oracle_grads(cfg, enc, dec, pcl, query, w_out, w_pen)

    abstract, glob, _ = enc(pcl.to(DEV)[None], False)
    assert abstract.requires_grad and glob.requires_grad
    out, pen = dec(query.to(DEV), abstract[0], glob[0], None)
    loss = (out * w_out.to(DEV)).sum() + (pen * w_pen.to(DEV)).sum()
    loss.backward()
    assert rel(out, out_ref) < 1e-4
    assert abs(float(loss) - loss_ref) < 1e-4 * max(1.0, abs(loss_ref))
    worst = {}
    gmax = max(float(v.grad.abs().max()) for sd in (sd_e, sd_d) for v in sd.values())
    for mod, sd in ((enc, sd_e), (dec, sd_d)):
        for name, p in mod.named_parameters():
            assert p.grad is not None, name
            assert torch.isfinite(p.grad).all(), name
            ref = sd[name].grad
            if float(ref.abs().max()) < 1e-12:       # e.g. attn_mlp.2.bias: cancels in the softmax
                assert float(p.grad.abs().max()) < 1e-4 * gmax, name
                continue
            worst[name] = rel(p.grad, ref) if prec == 0 else rel2(p.grad, ref)
    # precision 0: max norm at the north-star bar; precision 1: Frobenius norm (mask flips, see the note
    # at the top) per tensor, and over all parameters together
    bad = {k: v for k, v in worst.items() if v > (TOL if prec == 0 else 1e-2)}
    assert not bad, bad
    num = sum(float((p.grad.double().cpu() - sd[n_].grad).pow(2).sum()) for mod, sd in ((enc, sd_e), (dec, sd_d))
              for n_, p in mod.named_parameters())
    den = sum(float(sd[n_].grad.pow(2).sum()) for mod, sd in ((enc, sd_e), (dec, sd_d))
              for n_, p in mod.named_parameters())
    assert math.sqrt(num / den) < 3e-3, math.sqrt(num / den)


@pytest.mark.parametrize('color_mode', ['rgb', 'rgb_nosigmoid', 'hsv'])
def test_callers_in_place_squashing_of_the_grad_enabled_output(color_mode):
    """pipeline.py:193-207 writes into the decoder output IN PLACE (sigmoid / clamp of the colour channels) while
    grad mode is on, through the batched 5-argument call form, then backpropagates.  The output must therefore
    not be a view created inside a custom autograd.Function; gradients must match the oracle's autograd."""
    cfg = configs.TINY_CARLA if color_mode == 'hsv' else configs.TINY_GREATER
    cfg = dict(cfg, implicit_args=dict(cfg['implicit_args'], d_out=16 if color_mode == 'hsv' else 5))
    enc, dec = configs.build_modules(cfg, DEV)
    dec.train()
    dec.o4d_precision = 0
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', cfg['name'] + '.npz'))
    query, abstract, glob = configs.synthetic_queries(cfg), torch.from_numpy(z['abstract']), torch.from_numpy(z['glob'])

    def squash(o):          # verbatim data flow of pipeline.py:199-207
        if color_mode == 'rgb':
            o[..., 1:4] = torch.sigmoid(o[..., 1:4])
        elif color_mode == 'rgb_nosigmoid':
            o[..., 1:4] = torch.clamp(o[..., 1:4].clone(), min=0.0, max=1.0)
        else:
            o[..., 13:15] = torch.clamp(o[..., 13:15].clone(), min=0.0, max=1.0)
        return o

    out, pen, extra = dec(query.to(DEV)[None], abstract.to(DEV)[None], glob.to(DEV)[None], None, False)
    assert extra is None and out.requires_grad
    out = squash(out)
    gen = torch.Generator().manual_seed(11)
    w = torch.randn(out.shape, generator=gen)
    (out * w.to(DEV)).sum().backward()
    # oracle: same weights, same squashing, torch.autograd on the CPU in fp64
    sd = {k: v.detach().double().cpu().requires_grad_(True) for k, v in dec.state_dict().items()}
    o_ref, _ = orc.decoder_forward(sd, cfg['implicit_args'], query.double(), abstract.double(), glob.double(),
                                   knn_fp32=True)
    o_ref = squash(o_ref[None].clone())
    (o_ref * w.double()).sum().backward()
    assert rel(out.detach(), o_ref.detach()) < 1e-4
    for name, p in dec.named_parameters():
        assert p.grad is not None, name
        ref = sd[name].grad
        if float(ref.abs().max()) > 1e-12:
            assert rel(p.grad, ref) < TOL, (name, rel(p.grad, ref))


def test_train_and_inference_paths_agree_and_directional_derivative_c2_widths():
    """Full decoder widths (config 2: d_hidden 416, 6 blocks, 2 cross layers, K=14) on 4096 queries:
    the materialising training forward equals the fused inference forward, and the analytic
    directional derivative <grad, v> matches a central finite difference of the INFERENCE kernels."""
    cfg = configs.C2_GREATER
    _, dec = configs.build_modules(cfg, DEV)
    g = torch.Generator().manual_seed(11)
    m, e = 531, cfg['implicit_args']['d_latent_local']
    abstract = torch.cat([torch.rand(m, 3, generator=g) * 10 - 5, torch.randn(m, e, generator=g) * 0.5], 1).to(DEV)
    glob = (torch.randn(128, generator=g) * 0.5).to(DEV)
    query = configs.synthetic_queries(cfg, num=4096, mode='random').to(DEV)
    w_out = (torch.randn(query.shape[0], cfg['implicit_args']['d_out'], generator=g) / query.shape[0]).to(DEV)
    with torch.no_grad():
        out_inf, pen_inf = dec(query, abstract, glob, None)
    dec.train()
    a_req = abstract.clone().requires_grad_(True)
    out_tr, pen_tr = dec(query, a_req, glob, None)
    assert rel(out_tr, out_inf) < 1e-4 and rel(pen_tr, pen_inf) < 1e-4
    (out_tr * w_out).sum().backward()
    params = [p for p in dec.parameters()]
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in params)
    assert a_req.grad is not None and float(a_req.grad[:, :3].abs().max()) == 0.0
    # direction: a random perturbation scaled per tensor
    vs = [torch.randn(p.shape, generator=g).to(DEV) * p.detach().abs().mean() for p in params]
    analytic = sum(float((p.grad.double() * v.double()).sum()) for p, v in zip(params, vs))
    eps = 2e-3

    def loss_at(sign):
        with torch.no_grad():
            for p, v in zip(params, vs):
                p.add_(v, alpha=sign * eps)
            o, _ = dec(query, abstract, glob, None)
            val = float((o.double() * w_out.double()).sum())
            for p, v in zip(params, vs):
                p.add_(v, alpha=-sign * eps)
        return val

    fd = (loss_at(+1) - loss_at(-1)) / (2 * eps)
    assert abs(fd - analytic) < 0.05 * max(abs(fd), abs(analytic)), (fd, analytic)


def test_train_step_config5_shape_runs_and_reports_memory():
    """One CARLA-shaped frame of a training step (config 5: 17,203 queries per frame, M = 2124,
    K_c = 14, d_out 18): forward + backward through the module API, finite gradients."""
    cfg = configs.C3_CARLA
    _, dec = configs.build_modules(cfg, DEV)
    dec.train()
    g = torch.Generator().manual_seed(5)
    m, e = 2124, cfg['implicit_args']['d_latent_local']
    abstract = torch.cat([torch.rand(m, 3, generator=g) * 30, torch.randn(m, e, generator=g) * 0.5], 1).to(DEV)
    abstract.requires_grad_(True)
    glob = (torch.randn(128, generator=g) * 0.5).to(DEV).requires_grad_(True)
    query = torch.cat([torch.rand(17203, 3, generator=g) * 30, torch.full((17203, 1), 3.0)], 1).to(DEV)
    torch.cuda.reset_peak_memory_stats()
    out, _ = dec(query, abstract, glob, None)
    loss = torch.nn.functional.binary_cross_entropy_with_logits(out[:, 0], (out[:, 1] > 0).float())
    loss.backward()
    torch.cuda.synchronize()
    assert torch.isfinite(abstract.grad).all() and torch.isfinite(glob.grad).all()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in dec.parameters())
    print('peak memory %.1f GiB' % (torch.cuda.max_memory_allocated() / 2 ** 30))
