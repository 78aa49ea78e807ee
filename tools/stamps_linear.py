#!/usr/bin/env python
"""In-kernel cycle stamps (main loop / accumulator wait / epilogue of one CTA) of the tcgen05 dense layer for
several shapes; O4D_TC_PAIR=1 reads the 2-CTA kernel's stamps.  Usage (on a B200): python tools/stamps_linear.py"""
import sys, ctypes, math, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'occlusions-4d_b200'))
import torch
from o4d import ops, _lib
h = ctypes.CDLL(_lib.LIB_PATH)
def run(rows,k,n,res,relu_in=False,reps=5):
    a = torch.randn(rows,k,device='cuda'); w = torch.randn(n,k,device='cuda')/math.sqrt(k); b = torch.randn(n,device='cuda')
    r = torch.randn(rows,n,device='cuda') if res else None
    out = torch.empty(rows,n,device='cuda')
    for _ in range(2): ops.linear(a,w,b,residual=r,relu_in=relu_in,precision=1,out=out)
    torch.cuda.synchronize()
    e0,e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): ops.linear(a,w,b,residual=r,relu_in=relu_in,precision=1,out=out)
    e1.record(); torch.cuda.synchronize()
    buf=(ctypes.c_longlong*16)(); (h.o4d_debug_read_tc2 if os.environ.get("O4D_TC_PAIR")=="1" else h.o4d_debug_read_tc)(buf); v=list(buf)
    print('rows %6d k %4d n %4d res %d: %.1f us  loop %6d wait %5d epi %6d' % (rows,k,n,res,e0.elapsed_time(e1)/reps*1e3, v[1]-v[0], v[2]-v[1], v[3]-v[2]))
for rows in (2048, 8192, 18944, 32768, 65536):
    run(rows,416,416,1); run(rows,416,416,0)
run(32768,32,416,0); run(32768,832,416,0); run(32768,416,832,0); run(32768,288,416,1)
