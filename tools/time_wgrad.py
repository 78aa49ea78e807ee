#!/usr/bin/env python
"""Weight-gradient kernel alone (o4d_linear_backward_f32 with dA = NULL): time and error against an fp64 product for
the shapes of a config-5 decoder frame.  Usage (on a B200): python tools/time_wgrad.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'occlusions-4d_b200'))
import torch
from o4d import _lib, ops
from o4d.ops import _ptr, workspace
L = _lib.lib()
dev = torch.device('cuda', 0)
g = torch.Generator(device='cuda').manual_seed(3)
for rows, k, n, relu in ((240842, 832, 416, 1), (240842, 416, 832, 0), (240842, 64, 416, 1), (17203, 416, 416, 1), (14336, 36, 36, 0)):
    x = torch.randn(rows, k, device=dev, generator=g)
    dy = torch.randn(rows, n, device=dev, generator=g) * 0.1
    w = torch.randn(n, k, device=dev, generator=g)
    xr = x.clamp_min(0) if relu else x
    ref_w = (dy.double().t() @ xr.double()).float()
    ref_b = dy.double().sum(0).float()
    for prec in (1, 2):
        dw = torch.empty(n, k, device=dev); db = torch.empty(n, device=dev)
        nbytes = L.o4d_linear_backward_workspace_bytes(rows, k, n)
        ws = workspace(dev, nbytes, slot=2)
        st = torch.cuda.current_stream().cuda_stream
        def call():
            rc = L.o4d_linear_backward_f32(_ptr(x), rows, k, k, _ptr(w), k, n, _ptr(dy), n, ops.RELU_IN if relu else 0,
                                           None, k, _ptr(dw), k, _ptr(db), prec, _ptr(ws), ws.numel(), st)
            _lib.check(rc, 'bwd')
        for _ in range(2): call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): call()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        ew = float((dw - ref_w).abs().max() / ref_w.abs().max()); eb = float((db - ref_b).abs().max() / ref_b.abs().max())
        if os.environ.get('O4D_LIB'):
            import ctypes
            buf = (ctypes.c_longlong * 8)(); ctypes.CDLL(_lib.LIB_PATH).o4d_debug_read_wgrad(buf); v = list(buf)
            print('   stamps (%d chunks): loader total %d, wait empty staging %d | converter 0 total %d, wait staging full %d, wait UMMA empty %d | MMA total %d, wait UMMA full %d' % (v[7], v[0], v[1], v[4], v[2], v[3], v[6], v[5]))
        print('rows %d k %d n %d relu %d prec %d: %.3f ms (%.0f alg TFLOP/s)  dW err %.2e  db err %.2e' % (rows, k, n, relu, prec, ms, 2.0 * rows * k * n / ms / 1e9, ew, eb), flush=True)
