"""Training path: torch.autograd.Function wrappers over the backward kernels of libo4d.so.

The reference trains through torch.autograd over its eager graph (train.py:282-296 ->
pipeline.py:93-212).  Here every forward AND every gradient is a kernel behind the C ABI
(include/o4d.h, "training path"); torch.autograd only routes tensors between them, so
``loss.backward()`` in the unmodified train.py fills ``.grad`` of the same parameters.

Each Function states the reference lines whose derivative it implements.  Index / distance
inputs carry no gradient (coordinates are data; kNN and FPS are piecewise constant).
"""
import ctypes

import torch

from . import _lib, ops
from .ops import _f32, _ptr, _rows, _stream, workspace


def _c(t):
    """contiguous fp32 CUDA tensor"""
    return _f32(t, 'tensor').contiguous()


def _prec(precision):
    return ops.default_precision() if precision is None else int(precision)


class LinearFn(torch.autograd.Function):
    """post(pre(x) W^T + b) [+ residual]: every nn.Linear of the path (cuBLAS SGEMM + autograd's
    AddmmBackward in the reference).  Gradients: o4d_linear_backward_f32."""

    @staticmethod
    def forward(ctx, x, weight, bias, residual, relu_in, relu_out, precision):
        lead = x.shape[:-1]
        x2 = _c(x.reshape(-1, x.shape[-1]))
        w = _c(weight)
        n, k = w.shape
        assert x2.shape[1] == k, 'linear: input width %d != weight in_features %d' % (x2.shape[1], k)
        b = _c(bias) if bias is not None else None
        r = _c(residual.reshape(-1, n)) if residual is not None else None
        prec = _prec(precision)
        flags = (ops.RELU_IN if relu_in else 0) | (ops.RELU_OUT if relu_out else 0)
        with torch.cuda.device(x2.device):
            # allocated in its final shape: a Function must not hand out a VIEW of a tensor it created (the
            # reference's callers write into the decoder output in place, pipeline.py:196-207, and autograd
            # rejects in-place writes to views made inside a custom Function)
            y = torch.empty((*lead, n), dtype=torch.float32, device=x2.device)
            if x2.shape[0] > 0:        # an empty batch has no device pointers to hand over
                # packed-weight cache keyed on (storage, version): a weight is packed once per optimizer step, not
                # once per call (4 decoder frames per training step share every weight)
                ops.linear_call(x2, k, w, b, r, n, y.reshape(-1, n), flags, prec)
        assert not (relu_out and residual is not None), 'linear: relu_out with a residual is not used on the path'
        ctx.save_for_backward(x2, w, y if relu_out else None)
        ctx.meta = (lead, relu_in, relu_out, prec, bias is not None, residual is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x2, w, y = ctx.saved_tensors
        lead, relu_in, relu_out, prec, has_bias, has_res = ctx.meta
        n, k = w.shape
        L = _lib.lib()
        dy2 = _c(dy.reshape(-1, n))
        rows = x2.shape[0]
        need_x, need_w, need_b, need_r = ctx.needs_input_grad[:4]
        if rows == 0:                  # empty batch: zero parameter gradients, empty input gradient
            return (x2.new_zeros((*lead, k)) if need_x else None, torch.zeros_like(w) if need_w else None,
                    w.new_zeros((n,)) if (need_b and has_bias) else None,
                    dy2.reshape(*lead, n) if (need_r and has_res) else None, None, None, None)
        with torch.cuda.device(x2.device):
            st = _stream(x2)
            g = dy2
            if relu_out:
                g = torch.empty_like(dy2)
                _lib.check(L.o4d_relu_backward_f32(_ptr(dy2), _ptr(y.reshape(-1, n)), dy2.numel(), _ptr(g), st),
                           'o4d_relu_backward_f32')
            dx = torch.empty((rows, k), dtype=torch.float32, device=x2.device) if need_x else None
            dw = torch.empty((n, k), dtype=torch.float32, device=x2.device) if need_w else None
            db = torch.empty((n,), dtype=torch.float32, device=x2.device) if (need_b and has_bias) else None
            if dx is not None or dw is not None or db is not None:
                nbytes = L.o4d_linear_backward_workspace_bytes(rows, k, n)
                ws = workspace(x2.device, nbytes, slot=2)
                rc = L.o4d_linear_backward_f32(_ptr(x2), rows, k, k, _ptr(w), k, n, _ptr(g), n,
                                               ops.RELU_IN if relu_in else 0, _ptr(dx), k, _ptr(dw), k, _ptr(db),
                                               prec, _ptr(ws), ws.numel(), st)
                _lib.check(rc, 'o4d_linear_backward_f32')
        dres = dy2.reshape(*lead, n) if (need_r and has_res) else None
        return (dx.reshape(*lead, k) if dx is not None else None, dw, db, dres, None, None, None)


def linear(x, weight, bias=None, residual=None, relu_in=False, relu_out=False, precision=None):
    return LinearFn.apply(x, weight, bias, residual, bool(relu_in), bool(relu_out), precision)


class AttnCoreFn(torch.autograd.Function):
    """point_transformer_layer.py:174-179 given q / K table / V table and the neighbour lists:
    pos-MLP, q - k + delta, attention MLP, per-channel softmax over neighbours, aggregation."""

    @staticmethod
    def forward(ctx, q, ktab, vtab, pos, pos2, nbr, k, precision, *p8):
        L = _lib.lib()
        q, ktab, vtab = _c(q), _c(ktab), _c(vtab)
        pos_t, ldp = _rows(pos, 'pos')
        pos2_t, ldp2 = _rows(pos2, 'pos2')
        nbr = nbr.contiguous()
        assert nbr.dtype == torch.int64 and tuple(nbr.shape) == (q.shape[0], k)
        n, d = q.shape
        m = ktab.shape[0]
        prec = _prec(precision)
        params = [_c(p) for p in p8]
        tab = (ctypes.c_void_p * 8)(*[p.data_ptr() for p in params])
        with torch.cuda.device(q.device):
            agg = torch.empty((n, d), dtype=torch.float32, device=q.device)
            nbytes = L.o4d_attn_train_saved_bytes(n, d, int(k))
            saved = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=q.device)
            rc = L.o4d_attn_forward_train(tab, _ptr(q), _ptr(ktab), _ptr(vtab), m, _ptr(pos_t), ldp, _ptr(pos2_t),
                                          ldp2, _ptr(nbr), n, d, int(k), prec, _ptr(agg), _ptr(saved),
                                          saved.numel(), _stream(q))
        _lib.check(rc, 'o4d_attn_forward_train')
        ctx.save_for_backward(pos_t, pos2_t, nbr, saved, agg, *params)
        ctx.meta = (n, m, d, int(k), prec, ldp, ldp2)
        return agg

    @staticmethod
    def backward(ctx, dagg):
        pos_t, pos2_t, nbr, saved, agg = ctx.saved_tensors[:5]
        params = ctx.saved_tensors[5:]
        n, m, d, k, prec, ldp, ldp2 = ctx.meta
        L = _lib.lib()
        dagg = _c(dagg)
        dev = agg.device
        tab = (ctypes.c_void_p * 8)(*[p.data_ptr() for p in params])
        with torch.cuda.device(dev):
            dq = torch.empty((n, d), dtype=torch.float32, device=dev)
            dk = torch.empty((m, d), dtype=torch.float32, device=dev)
            dv = torch.empty((m, d), dtype=torch.float32, device=dev)
            dps = [torch.empty_like(p) for p in params]
            dtab = (ctypes.c_void_p * 8)(*[p.data_ptr() for p in dps])
            nbytes = L.o4d_attn_backward_workspace_bytes(n, d, k)
            ws = workspace(dev, nbytes, slot=3)
            rc = L.o4d_attn_backward(tab, _ptr(pos_t), ldp, _ptr(pos2_t), ldp2, _ptr(nbr), n, m, d, k, prec,
                                     _ptr(saved), saved.numel(), _ptr(agg), _ptr(dagg), _ptr(dq), _ptr(dk), _ptr(dv),
                                     dtab, _ptr(ws), ws.numel(), _stream(agg))
        _lib.check(rc, 'o4d_attn_backward')
        return (dq, dk, dv, None, None, None, None, None, *dps)


def attn_core(q, ktab, vtab, pos, pos2, nbr, k, precision, p8):
    return AttnCoreFn.apply(q, ktab, vtab, pos, pos2, nbr, int(k), precision, *p8)


class LocalBlendFn(torch.autograd.Function):
    """implicit.py:337-339: inverse-distance blend of the K_l nearest abstract feature rows."""

    @staticmethod
    def forward(ctx, feat, idx, dist):
        f, ld = _rows(feat, 'features_abstract')
        idx, dist = idx.contiguous(), _c(dist)
        n, k = idx.shape
        e = f.shape[1]
        with torch.cuda.device(f.device):
            out = torch.empty((n, e), dtype=torch.float32, device=f.device)
            rc = _lib.lib().o4d_local_blend_f32(_ptr(idx), _ptr(dist), _ptr(f), ld, n, k, e, _ptr(out), _stream(f))
        _lib.check(rc, 'o4d_local_blend_f32')
        ctx.save_for_backward(idx, dist)
        ctx.meta = (f.shape[0], e)
        return out

    @staticmethod
    def backward(ctx, dout):
        idx, dist = ctx.saved_tensors
        m, e = ctx.meta
        dout = _c(dout)
        n, k = idx.shape
        with torch.cuda.device(dout.device):
            dfeat = torch.empty((m, e), dtype=torch.float32, device=dout.device)
            rc = _lib.lib().o4d_local_blend_backward_f32(_ptr(idx), _ptr(dist), _ptr(dout), n, k, e, m, _ptr(dfeat), e,
                                                         _stream(dout))
        _lib.check(rc, 'o4d_local_blend_backward_f32')
        return dfeat, None, None


class GatherMaxFn(torch.autograd.Function):
    """modules.py:156-158: z_i = max over the k neighbour rows of y."""

    @staticmethod
    def forward(ctx, y, nbr):
        y = _c(y)
        nbr = nbr.contiguous()
        n_out, k = nbr.shape
        d = y.shape[1]
        with torch.cuda.device(y.device):
            z = torch.empty((n_out, d), dtype=torch.float32, device=y.device)
            arg = torch.empty((n_out, d), dtype=torch.int32, device=y.device)
            rc = _lib.lib().o4d_gather_max_f32(_ptr(y), d, _ptr(nbr), n_out, k, d, _ptr(z), _ptr(arg), _stream(y))
        _lib.check(rc, 'o4d_gather_max_f32')
        ctx.save_for_backward(arg)
        ctx.meta = (y.shape[0], d)
        return z

    @staticmethod
    def backward(ctx, dz):
        (arg,) = ctx.saved_tensors
        n_src, d = ctx.meta
        dz = _c(dz)
        with torch.cuda.device(dz.device):
            dy = torch.empty((n_src, d), dtype=torch.float32, device=dz.device)
            rc = _lib.lib().o4d_gather_max_backward_f32(_ptr(dz), _ptr(arg), dz.shape[0], d, n_src, _ptr(dy), d,
                                                        _stream(dz))
        _lib.check(rc, 'o4d_gather_max_backward_f32')
        return dy, None


class LayerNormReluFn(torch.autograd.Function):
    """relu(nn.LayerNorm(d)(y)), modules.py:107-110 (eps 1e-5, biased variance)."""

    @staticmethod
    def forward(ctx, y, gamma, beta, eps):
        y, gamma, beta = _c(y), _c(gamma), _c(beta)
        rows, d = y.shape
        with torch.cuda.device(y.device):
            out = torch.empty_like(y)
            rc = _lib.lib().o4d_layernorm_relu_f32(_ptr(y), rows, d, _ptr(gamma), _ptr(beta), float(eps), _ptr(out),
                                                   _stream(y))
        _lib.check(rc, 'o4d_layernorm_relu_f32')
        ctx.save_for_backward(y, gamma, beta)
        ctx.eps = float(eps)
        return out

    @staticmethod
    def backward(ctx, dout):
        y, gamma, beta = ctx.saved_tensors
        dout = _c(dout)
        rows, d = y.shape
        with torch.cuda.device(y.device):
            dy = torch.empty_like(y)
            dg = torch.empty_like(gamma)
            db = torch.empty_like(beta)
            rc = _lib.lib().o4d_layernorm_relu_backward_f32(_ptr(y), _ptr(dout), rows, d, _ptr(gamma), _ptr(beta),
                                                            ctx.eps, _ptr(dy), _ptr(dg), _ptr(db), _stream(y))
        _lib.check(rc, 'o4d_layernorm_relu_backward_f32')
        return dy, dg, db, None


class ColMeanFn(torch.autograd.Function):
    """torch.mean over the point dimension, model.py:189."""

    @staticmethod
    def forward(ctx, x):
        x = _c(x)
        rows, d = x.shape
        with torch.cuda.device(x.device):
            out = torch.empty((d,), dtype=torch.float32, device=x.device)
            rc = _lib.lib().o4d_col_mean_f32(_ptr(x), rows, d, _ptr(out), _stream(x))
        _lib.check(rc, 'o4d_col_mean_f32')
        ctx.meta = (rows, d)
        return out

    @staticmethod
    def backward(ctx, dmean):
        rows, d = ctx.meta
        dmean = _c(dmean)
        with torch.cuda.device(dmean.device):
            dx = torch.empty((rows, d), dtype=torch.float32, device=dmean.device)
            rc = _lib.lib().o4d_col_mean_backward_f32(_ptr(dmean), rows, d, _ptr(dx), _stream(dmean))
        _lib.check(rc, 'o4d_col_mean_backward_f32')
        return dx


# ------------------------------------------------------------------ composed training forwards
# (one cloud each; the nn.Modules loop over the batch exactly like their inference paths)

def pt_layer_train(layer, x, pos, x2, pos2, prec=None, nbr=None):
    """PointTransformerLayer.forward, point_transformer_layer.py:148-183, differentiable.  `nbr`: the (n, k) neighbour
    list of (pos, pos2) when the caller already has it (the decoder's cross-attention layers share one)."""
    if x2 is None:
        x2, pos2 = x, pos
    prec = layer.o4d_precision if prec is None else prec
    k = layer.num_neighbors
    if nbr is None:
        with torch.no_grad():
            nbr = ops.knn(pos, pos2, k)                                        # :167
    q = linear(x, layer.to_q.weight, precision=prec)                          # :170
    ktab = linear(x2, layer.to_k.weight, precision=prec)                      # :171 (before the gather)
    vtab = linear(x2, layer.to_v.weight, precision=prec)                      # :172
    p8 = [layer.pos_mlp[0].weight, layer.pos_mlp[0].bias, layer.pos_mlp[2].weight, layer.pos_mlp[2].bias,
          layer.attn_mlp[0].weight, layer.attn_mlp[0].bias, layer.attn_mlp[2].weight, layer.attn_mlp[2].bias]
    return attn_core(q, ktab, vtab, pos, pos2, nbr, k, prec, p8)               # :174-179


def pt_block_train(block, x, pos, x2, pos2, prec=None, nbr=None):
    """PointTransformerBlock.forward, modules.py:45-67 (x2 goes RAW into layer2 in cross mode)."""
    prec = block.o4d_precision if prec is None else prec
    y = linear(x, block.layer1.weight, block.layer1.bias, precision=prec)     # :61
    y = pt_layer_train(block.layer2, y, pos, x2, pos2, prec, nbr)             # :63
    return linear(y, block.layer3.weight, block.layer3.bias, residual=x, precision=prec)   # :64-65


def down_train(down, x, pos, start_idx, prec=None):
    """DownTransition.forward, modules.py:113-163."""
    prec = down.o4d_precision if prec is None else prec
    n = x.shape[0]
    n_out = -(-n // down.factor)                                               # :126
    with torch.no_grad():
        fidx = ops.fps(pos, n_out, start_idx)                                  # :133-135 (sorted)
        pos_sub = pos[fidx].contiguous()                                       # :137
        nbr = ops.knn(pos_sub, pos, down.knn_k)                                # :142-146
    if down.norm_type == 'layer':
        y = linear(x, down.mlp[0].weight, down.mlp[0].bias, precision=prec)   # :152
        y = LayerNormReluFn.apply(y, down.mlp[1].weight, down.mlp[1].bias, down.mlp[1].eps)
    else:
        y = linear(x, down.mlp[0].weight, down.mlp[0].bias, relu_out=True, precision=prec)
    return GatherMaxFn.apply(y, nbr), pos_sub                                  # :156-158


def encoder_train(net, pcl, starts):
    """PointCompletionNetV3.forward, model.py:148-233, one cloud (N, d_in), differentiable."""
    from . import modules
    prec = net.o4d_precision
    x = linear(pcl, net.pre_mlp[0].weight, net.pre_mlp[0].bias, relu_out=True, precision=prec)    # :167
    x = linear(x, net.pre_mlp[2].weight, net.pre_mlp[2].bias, precision=prec)
    pos = pcl[:, :3].contiguous()                                              # :168
    coords = [pos]
    skips = []
    final_dim = net.d_feat * (2 ** net.down_blocks)
    dim = net.d_feat
    di = 0
    for blk in net.blocks:
        if isinstance(blk, modules.PointTransformerBlock):
            x = pt_block_train(blk, x, pos, None, None, prec)
        else:
            x, pos = down_train(blk, x, pos, starts[di] if starts is not None else 0, prec)
            di += 1
            dim *= 2
            coords.append(pos)
            if net.abstract_levels > 1:                                        # :202-207
                for j, lin in enumerate(net.abstract_skip_mlps):
                    if lin.in_features == dim:
                        y = linear(x, lin.weight, lin.bias, precision=prec)
                        level = torch.full((y.shape[0], 1), float(j + 1), dtype=y.dtype, device=y.device)
                        skips.append(torch.cat([pos, y[:, :-1], level], dim=-1))   # y[..., -1] = level
    x_avg = ColMeanFn.apply(x)                                                 # :189
    g = linear(x_avg[None], net.global_mlp[0].weight, net.global_mlp[0].bias, relu_out=True, precision=0)
    g = linear(g, net.global_mlp[2].weight, net.global_mlp[2].bias, precision=0)[0]
    if net.abstract_levels > 1:                                                # :220-228
        level = torch.full((x.shape[0], 1), float(net.abstract_levels), dtype=x.dtype, device=x.device)
        out = torch.cat(skips + [torch.cat([pos, x[:, :-1], level], dim=-1)], dim=0)
    else:
        out = torch.cat([pos, x], dim=-1)
    assert dim == final_dim
    return out, g, coords


def decoder_train(net, query, pcl_abstract, feat_global):
    """LocalPclResnetFC.forward + do_forward_attention, implicit.py:271-445, unbatched and
    differentiable w.r.t. every parameter, the abstract features and the global embedding."""
    from .implicit import positional_encode
    prec = net.o4d_precision
    dg = net.d_latent - net.d_latent_local
    abs_xyz = pcl_abstract[:, :3].detach().contiguous()                       # :286-290
    abs_feat = pcl_abstract[:, 3:]
    query = query.detach()
    q_xyz = query[:, :3].contiguous()
    # Every cross-attention layer searches the same (query, abstract) pair with the same K (point_transformer_layer.py:167
    # recomputes it per layer): one scan yields that list and the local-feature list (:328), as in the inference path.
    kc = {net.pt_blocks[i].layer2.num_neighbors for i in net.use_pt_inds.values()}
    kc = kc.pop() if len(kc) == 1 else None
    kl = net.num_local_features
    with torch.no_grad():
        nbr_c = None
        if kc is not None and kl < kc and kc >= 9:
            nbr_c, idx, dist = ops.knn_two_lists(q_xyz, abs_xyz, kc, kl)
        else:
            idx, dist = ops.knn(q_xyz, abs_xyz, kl, sqrt_dist=True, return_dist=True)                   # :328
            if kc is not None:
                nbr_c = ops.knn(q_xyz, abs_xyz, kc)
        pe = positional_encode(query, 0.1, net.pos_encoding_freqs) if net.pos_encoding_freqs > 0 else query
    f_loc = LocalBlendFn.apply(abs_feat, idx, dist)                            # :337-339
    x = linear(pe, net.lin_in.weight, net.lin_in.bias, precision=prec)        # :403-408
    for b in range(net.n_blocks):
        wz = net.lin_z[b].weight
        # lin_z(cat[global, local]) = W[:, :Dg] g + b  +  W[:, Dg:] f_local        (:416-418)
        zg = linear(feat_global[None], wz[:, :dg], net.lin_z[b].bias, precision=0)[0]
        x = linear(f_loc, wz[:, dg:], zg, residual=x, precision=prec)
        blk = net.blocks[b]
        h = linear(x, blk.fc_0.weight, blk.fc_0.bias, relu_in=True, precision=prec)              # :93
        x = linear(h, blk.fc_1.weight, blk.fc_1.bias, residual=x, relu_in=True, precision=prec)  # :94-101
        if b in net.use_pt_inds:                                                # :421-439
            pt = net.pt_blocks[net.use_pt_inds[b]]
            x = pt_block_train(pt, x, q_xyz, abs_feat, abs_xyz, prec, nbr_c)
    out = linear(x, net.lin_out.weight, net.lin_out.bias, relu_in=True, precision=prec)          # :441-443
    return out, x
