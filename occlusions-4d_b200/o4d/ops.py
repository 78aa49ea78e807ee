"""Tensor-level wrappers over the C ABI (include/o4d.h).

PyTorch is used for device memory, streams and module plumbing only; every arithmetic
operation below runs in libo4d.so.  All inputs must be CUDA fp32 tensors -- there is no
CPU path.
"""
import collections
import ctypes

import torch

from . import _lib
from ._lib import DecoderConfig, EncoderConfig  # noqa: F401

RELU_IN = 1
RELU_OUT = 2
MAX_K = 16


def default_precision():
    """1 (tcgen05 bf16x3, fp32-grade) when the library has the tensor-core path, else 0."""
    import os
    env = os.environ.get('O4D_PRECISION')
    if env is not None:
        return int(env)
    return 1 if _lib.lib().o4d_has_tcgen05() else 0


def _stream(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _f32(t, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError('%s must be a tensor' % name)
    if not t.is_cuda:
        raise RuntimeError('%s is on %s: the o4d kernels run on CUDA only (no CPU fallback)'
                           % (name, t.device))
    if t.dtype != torch.float32:
        t = t.float()
    return t


def _rows(t, name):
    """2-D view with unit inner stride; returns (tensor, leading dimension)."""
    t = _f32(t, name)
    assert t.dim() == 2, '%s must be 2-D' % name
    if t.stride(1) != 1 or (t.shape[0] > 1 and t.stride(0) < t.shape[1]):
        t = t.contiguous()
    ld = t.stride(0) if t.shape[0] > 1 else max(t.shape[1], t.stride(0))
    return t, ld


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


_workspaces = {}


def workspace(device, nbytes, slot=0):
    """Per (device, stream, slot) scratch buffer from torch's caching allocator, grown on demand."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream, slot)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


def param_table(tensors):
    """Device-pointer table (host array of const float*) + keep-alive list."""
    keep = []
    arr = (ctypes.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        if t is None:
            arr[i] = None
            continue
        t = t.detach()
        if t.dtype != torch.float32 or not t.is_contiguous():
            t = t.float().contiguous()
        if not t.is_cuda:
            raise RuntimeError('module parameters are on %s: move the module to CUDA (no CPU fallback)'
                               % t.device)
        keep.append(t)
        arr[i] = t.data_ptr()
    return arr, keep


# ---------------------------------------------------------------------------- kNN / FPS

def knn(query, ref, k, sqrt_dist=False, return_dist=False):
    """query (N, >=3), ref (M, >=3) -> idx (N, k) int64 [, dist (N, k)]; ascending (distance, index)."""
    q, ldq = _rows(query, 'query')
    r, ldr = _rows(ref, 'ref')
    assert q.shape[1] >= 3 and r.shape[1] >= 3
    with torch.cuda.device(q.device):
        idx = torch.empty((q.shape[0], k), dtype=torch.int64, device=q.device)
        dist = torch.empty((q.shape[0], k), dtype=torch.float32, device=q.device) if return_dist else None
        rc = _lib.lib().o4d_knn_f32(_ptr(q), q.shape[0], ldq, _ptr(r), r.shape[0], ldr, int(k),
                                    int(bool(sqrt_dist)), _ptr(idx), _ptr(dist), _stream(q))
    _lib.check(rc, 'o4d_knn_f32')
    return (idx, dist) if return_dist else idx


def knn_two_lists(query, ref, k, k2):
    """One scan of `ref`: (idx (n, k) by squared distance, idx2 (n, k2), dist2 (n, k2) by Euclidean distance); identical
    to knn(..., k) and knn(..., k2, sqrt_dist=True, return_dist=True)."""
    q, ldq = _rows(query, 'query')
    r, ldr = _rows(ref, 'ref')
    L = _lib.lib()
    n = q.shape[0]
    with torch.cuda.device(q.device):
        idx = torch.empty((n, k), dtype=torch.int64, device=q.device)
        idx2 = torch.empty((n, k2), dtype=torch.int64, device=q.device)
        dist2 = torch.empty((n, k2), dtype=torch.float32, device=q.device)
        nbytes = L.o4d_knn_two_lists_workspace_bytes(n, int(k), int(k2))
        ws = workspace(q.device, nbytes, slot=5)
        rc = L.o4d_knn_two_lists_f32(_ptr(q), n, ldq, _ptr(r), r.shape[0], ldr, int(k), int(k2), _ptr(idx), _ptr(idx2),
                                     _ptr(dist2), _ptr(ws), ws.numel(), _stream(q))
    _lib.check(rc, 'o4d_knn_two_lists_f32')
    return idx, idx2, dist2


def fps(xyz, n_out, start_idx=0, return_order=False):
    """xyz (N, >=3) -> sorted indices (n_out,) int64 [, selection order]."""
    p, ld = _rows(xyz, 'xyz')
    L = _lib.lib()
    with torch.cuda.device(p.device):
        out = torch.empty((n_out,), dtype=torch.int64, device=p.device)
        order = torch.empty((n_out,), dtype=torch.int64, device=p.device) if return_order else None
        nbytes = L.o4d_fps_workspace_bytes(p.shape[0], n_out)
        ws = workspace(p.device, nbytes)
        rc = L.o4d_fps_f32(_ptr(p), p.shape[0], ld, int(n_out), int(start_idx), _ptr(out), _ptr(order),
                           _ptr(ws), ws.numel(), _stream(p))
    _lib.check(rc, 'o4d_fps_f32')
    return (out, order) if return_order else out


# ---------------------------------------------------------------------------- dense layer

# Caller-owned tensor-core images of weights (o4d_linear_pack_f32): one per (weight storage, version, shape, device),
# re-packed when the weight is updated in place (optimizer step) or replaced; bounded, least recently used first out.
_PACKED = collections.OrderedDict()
_PACKED_MAX = 512


def packed_weight(w, nbytes):
    """bf16 hi/lo image of the contiguous fp32 weight `w` (n, k), cached on (data_ptr, _version)."""
    key = (w.data_ptr(), w._version, tuple(w.shape), w.device.index)
    hit = _PACKED.get(key)
    if hit is not None:
        _PACKED.move_to_end(key)
        return hit[0]
    buf = torch.empty(int(nbytes), dtype=torch.uint8, device=w.device)
    rc = _lib.lib().o4d_linear_pack_f32(_ptr(w), w.shape[0], w.shape[1], w.shape[1], _ptr(buf), _stream(w))
    _lib.check(rc, 'o4d_linear_pack_f32')
    # the weight tensor is kept alive with the entry: its address cannot be recycled for another weight meanwhile
    _PACKED[key] = (buf, w)
    while len(_PACKED) > _PACKED_MAX:
        _PACKED.popitem(last=False)
    return buf


def linear_call(a, lda, w, b, r, ldr, c, flags, prec):
    """One dense layer through the C ABI: the packed form (cached image, no allocation inside the call) whenever the
    shape is on the tensor-core path, o4d_linear_f32 otherwise."""
    L = _lib.lib()
    n, k = w.shape
    nbytes = L.o4d_linear_pack_bytes(a.shape[0], k, n, prec) if a.shape[0] > 0 else 0
    if nbytes:
        rc = L.o4d_linear_packed_f32(_ptr(a), a.shape[0], k, lda, _ptr(packed_weight(w, nbytes)), _ptr(b), n, _ptr(r),
                                     ldr or 0, _ptr(c), n, flags, prec, _stream(a))
        _lib.check(rc, 'o4d_linear_packed_f32')
    else:
        rc = L.o4d_linear_f32(_ptr(a), a.shape[0], k, lda, _ptr(w), _ptr(b), n, _ptr(r), ldr or 0, _ptr(c), n, flags, prec,
                              _stream(a))
        _lib.check(rc, 'o4d_linear_f32')


def linear(x, weight, bias=None, residual=None, relu_in=False, relu_out=False, precision=None, out=None):
    """post(pre(x) @ weight.T + bias) [+ residual] over the last dimension of x."""
    x = _f32(x, 'x')
    lead = x.shape[:-1]
    a, lda = _rows(x.reshape(-1, x.shape[-1]), 'x')
    w = _f32(weight.detach(), 'weight').contiguous()
    n, k = w.shape
    assert a.shape[1] == k, 'linear: input width %d != weight in_features %d' % (a.shape[1], k)
    b = _f32(bias.detach(), 'bias').contiguous() if bias is not None else None
    r = ldr = None
    if residual is not None:
        r, ldr = _rows(_f32(residual, 'residual').reshape(-1, n), 'residual')
    with torch.cuda.device(a.device):
        c = out if out is not None else torch.empty((a.shape[0], n), dtype=torch.float32, device=a.device)
        flags = (RELU_IN if relu_in else 0) | (RELU_OUT if relu_out else 0)
        prec = default_precision() if precision is None else int(precision)
        linear_call(a, lda, w, b, r, ldr, c, flags, prec)
    return c.reshape(*lead, n)


def resblock(x, w0, b0, w1, b1, precision=None):
    """x + fc_1(relu(fc_0(relu(x)))) (ResnetBlockFC.forward, implicit.py:93-101) as one launch of the fused
    multi-layer kernel; None when the shape or precision is outside it (the caller composes two dense layers)."""
    prec = default_precision() if precision is None else int(precision)
    x = _f32(x, 'x')
    lead = x.shape[:-1]
    a = x.reshape(-1, x.shape[-1]).contiguous()
    d, dh = a.shape[1], w0.shape[0]
    L = _lib.lib()
    if prec == 0 or tuple(w0.shape) != (dh, d) or tuple(w1.shape) != (d, dh) or a.shape[0] < 1024:
        return None
    nbytes = L.o4d_resblock_workspace_bytes(a.shape[0], d, dh)
    if nbytes == 0:
        return None
    w0c, w1c = _f32(w0.detach(), 'w0').contiguous(), _f32(w1.detach(), 'w1').contiguous()
    b0c = _f32(b0.detach(), 'b0').contiguous() if b0 is not None else None
    b1c = _f32(b1.detach(), 'b1').contiguous() if b1 is not None else None
    with torch.cuda.device(a.device):
        ws = workspace(a.device, nbytes, slot=4)
        out = torch.empty_like(a)
        rc = L.o4d_resblock_forward_f32(_ptr(a), a.shape[0], d, d, _ptr(w0c), _ptr(b0c), dh, _ptr(w1c), _ptr(b1c),
                                        _ptr(out), d, prec, _ptr(ws), ws.numel(), _stream(a))
    _lib.check(rc, 'o4d_resblock_forward_f32')
    return out.reshape(*lead, d)


# ---------------------------------------------------------------------------- attention

def pt_layer_forward(params, x, pos, x2, pos2, k, precision=None, return_idx=False):
    """One cloud: x (N,D), pos (N,3) [x2 (M,D2), pos2 (M,3)] -> (N,D). params = 11 tensors."""
    return _attn_call('o4d_pt_layer_forward', params, x, pos, x2, pos2, k, precision, return_idx)


def pt_block_forward(params, x, pos, x2, pos2, k, precision=None, return_idx=False):
    """One cloud: z = x + layer3(attn(layer1(x))). params = 15 tensors (include/o4d.h)."""
    return _attn_call('o4d_pt_block_forward', params, x, pos, x2, pos2, k, precision, return_idx)


def _attn_call(fn_name, params, x, pos, x2, pos2, k, precision, return_idx):
    L = _lib.lib()
    xx, _ = _rows(x, 'x')
    xx = xx.contiguous()
    pp, ldp = _rows(pos, 'pos')
    n, d = xx.shape
    cross = x2 is not None
    if cross:
        x2t, ldx2 = _rows(x2, 'x2')
        p2t, ldp2 = _rows(pos2, 'pos2')
        m, d2 = x2t.shape
    else:
        x2t = p2t = None
        ldx2 = ldp2 = 0
        m, d2 = 0, 0
    tab, keep = param_table(params)
    with torch.cuda.device(xx.device):
        z = torch.empty((n, d), dtype=torch.float32, device=xx.device)
        idx = torch.empty((n, k), dtype=torch.int64, device=xx.device) if return_idx else None
        nbytes = L.o4d_pt_block_workspace_bytes(n, m, d, d2, int(k))
        ws = workspace(xx.device, nbytes)
        prec = default_precision() if precision is None else int(precision)
        rc = getattr(L, fn_name)(tab, _ptr(xx), n, d, _ptr(pp), ldp, _ptr(x2t), m, d2, ldx2, _ptr(p2t), ldp2,
                                 int(k), prec, _ptr(z), _ptr(idx), _ptr(ws), ws.numel(), _stream(xx))
    _lib.check(rc, fn_name)
    del keep
    return (z, idx) if return_idx else z


def down_forward(params, x, pos, d_out, factor, k, norm, start_idx=0, precision=None, return_idx=False):
    """One cloud DownTransition: -> (z (n_out,d_out), pos_sub (n_out,3) [, fps idx])."""
    L = _lib.lib()
    xx, _ = _rows(x, 'x')
    xx = xx.contiguous()
    pp, ldp = _rows(pos, 'pos')
    n, d_in = xx.shape
    n_out = -(-n // factor)
    tab, keep = param_table(params)
    with torch.cuda.device(xx.device):
        z = torch.empty((n_out, d_out), dtype=torch.float32, device=xx.device)
        ps = torch.empty((n_out, 3), dtype=torch.float32, device=xx.device)
        fidx = torch.empty((n_out,), dtype=torch.int64, device=xx.device) if return_idx else None
        nbytes = L.o4d_down_workspace_bytes(n, d_in, d_out, factor, k)
        ws = workspace(xx.device, nbytes)
        prec = default_precision() if precision is None else int(precision)
        rc = L.o4d_down_forward(tab, _ptr(xx), n, d_in, _ptr(pp), ldp, d_out, factor, k, norm, int(start_idx),
                                prec, _ptr(z), _ptr(ps), _ptr(fidx), _ptr(ws), ws.numel(), _stream(xx))
    _lib.check(rc, 'o4d_down_forward')
    del keep
    return (z, ps, fidx) if return_idx else (z, ps)


# ---------------------------------------------------------------------------- encoder / decoder

def encoder_forward(cfg, params, pcl, start_idx=None, return_levels=False):
    """One cloud pcl (N, d_in) -> (abstract (M, 3+E), global (G,), [level coordinates])."""
    L = _lib.lib()
    x, _ = _rows(pcl, 'pcl')
    x = x.contiguous()
    n = x.shape[0]
    assert x.shape[1] == cfg.d_in, 'encoder: input width %d != d_in %d' % (x.shape[1], cfg.d_in)
    expect = L.o4d_encoder_num_params(ctypes.byref(cfg))
    if expect < 0 or expect != len(params):
        raise RuntimeError('encoder: expected %d parameter tensors, got %d' % (expect, len(params)))
    tab, keep = param_table(params)
    m = L.o4d_encoder_num_abstract(ctypes.byref(cfg), n)
    e = cfg.d_feat << cfg.down_blocks
    with torch.cuda.device(x.device):
        abstract = torch.empty((m, 3 + e), dtype=torch.float32, device=x.device)
        glob = torch.empty((cfg.global_dim,), dtype=torch.float32, device=x.device)
        levels = None
        lev_arr = None
        if return_levels:
            levels, nl = [], n
            for _ in range(cfg.down_blocks + 1):
                levels.append(torch.empty((nl, 3), dtype=torch.float32, device=x.device))
                nl = -(-nl // cfg.transition_factor)
            lev_arr = (ctypes.c_void_p * len(levels))(*[t.data_ptr() for t in levels])
        starts = None
        if start_idx is not None:
            starts = (ctypes.c_int64 * cfg.down_blocks)(*[int(s) for s in start_idx])
        nbytes = L.o4d_encoder_workspace_bytes(ctypes.byref(cfg), n)
        ws = workspace(x.device, nbytes)
        rc = L.o4d_encoder_forward(ctypes.byref(cfg), tab, _ptr(x), n, starts, _ptr(abstract), _ptr(glob),
                                   lev_arr, _ptr(ws), ws.numel(), _stream(x))
    _lib.check(rc, 'o4d_encoder_forward')
    del keep
    return (abstract, glob, levels) if return_levels else (abstract, glob)


class DecoderScene:
    """Scene-constant decoder state (K/V tables, lin_z global halves, packed weights) living in one device buffer."""

    def __init__(self, cfg, params, pcl_abstract, feat_global):
        L = _lib.lib()
        a, ld = _rows(pcl_abstract, 'pcl_abstract')
        g = _f32(feat_global, 'features_global').contiguous().reshape(-1)
        self.m = a.shape[0]
        self.cfg = cfg
        assert a.shape[1] == 3 + cfg.d_latent_local, \
            'abstract cloud width %d != 3 + d_latent_local %d' % (a.shape[1], cfg.d_latent_local)
        assert g.numel() == cfg.d_latent - cfg.d_latent_local
        tab, keep = param_table(params)
        with torch.cuda.device(a.device):
            nbytes = L.o4d_decoder_scene_bytes(ctypes.byref(cfg), self.m)
            self.buf = torch.empty(nbytes, dtype=torch.uint8, device=a.device)
            rc = L.o4d_decoder_prepare_scene(ctypes.byref(cfg), tab, _ptr(a), self.m, ld, _ptr(g),
                                             _ptr(self.buf), nbytes, _stream(a))
        _lib.check(rc, 'o4d_decoder_prepare_scene')
        del keep

    def update(self, params, pcl_abstract, feat_global):
        """Next scene with the SAME weights and abstract-cloud size: only the scene-dependent parts of the buffer are
        rewritten (o4d_decoder_update_scene); the caller guarantees that no parameter changed since __init__."""
        L = _lib.lib()
        a, ld = _rows(pcl_abstract, 'pcl_abstract')
        g = _f32(feat_global, 'features_global').contiguous().reshape(-1)
        assert a.shape[0] == self.m and a.shape[1] == 3 + self.cfg.d_latent_local and a.device == self.buf.device
        tab, keep = param_table(params)
        with torch.cuda.device(a.device):
            rc = L.o4d_decoder_update_scene(ctypes.byref(self.cfg), tab, _ptr(a), self.m, ld, _ptr(g),
                                            _ptr(self.buf), self.buf.numel(), _stream(a))
        _lib.check(rc, 'o4d_decoder_update_scene')
        del keep


def decoder_forward(cfg, params, scene, query, want_penult=True, out=None):
    """query (N, 4) -> (out (N, G), penult (N, H) or None)."""
    L = _lib.lib()
    q, _ = _rows(query, 'points_query')
    q = q.contiguous()
    nq = q.shape[0]
    assert q.shape[1] == cfg.d_in
    tab, keep = param_table(params)
    with torch.cuda.device(q.device):
        if out is None:
            out = torch.empty((nq, cfg.d_out), dtype=torch.float32, device=q.device)
        else:
            assert out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and \
                tuple(out.shape) == (nq, cfg.d_out)
        pen = torch.empty((nq, cfg.d_hidden), dtype=torch.float32, device=q.device) if want_penult else None
        nbytes = L.o4d_decoder_workspace_bytes(ctypes.byref(cfg), max(nq, 1), scene.m)
        ws = workspace(q.device, nbytes)
        rc = L.o4d_decoder_forward(ctypes.byref(cfg), tab, _ptr(scene.buf), scene.m, _ptr(q), nq, _ptr(out),
                                   _ptr(pen), _ptr(ws), ws.numel(), _stream(q))
    _lib.check(rc, 'o4d_decoder_forward')
    del keep
    return out, pen


def decoder_run_host(cfg, params, scene, query_host, batch, out_host=None):
    """The eval/inference.py:204-246 loop through the C ABI with HOST buffers (query (N,4) CPU fp32)."""
    L = _lib.lib()
    assert not query_host.is_cuda and query_host.dtype == torch.float32 and query_host.is_contiguous()
    nq = query_host.shape[0]
    if out_host is None:
        out_host = torch.empty((nq, cfg.d_out), dtype=torch.float32, pin_memory=True)
    dev = scene.buf.device
    tab, keep = param_table(params)
    with torch.cuda.device(dev):
        nbytes = L.o4d_decoder_run_host_device_bytes(ctypes.byref(cfg), int(batch), scene.m)
        ws = workspace(dev, nbytes, slot=1)
        rc = L.o4d_decoder_run_host(ctypes.byref(cfg), tab, _ptr(scene.buf), scene.m,
                                    ctypes.c_void_p(query_host.data_ptr()), nq, int(batch),
                                    ctypes.c_void_p(out_host.data_ptr()), _ptr(ws), ws.numel(),
                                    ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    _lib.check(rc, 'o4d_decoder_run_host')
    del keep
    return out_host
