#!/usr/bin/env python
"""In-kernel cycle stamps of the fused attention kernel (one tile in the middle of the grid) and of the dense
layer, read through o4d_debug_read / o4d_debug_read_tc after decoder passes over one 32768-query mini-batch,
for precision 1 (bf16x3) and 2 (single bf16 pass).  Used for the bottleneck analysis in DESIGN.md section 4.
Usage (on a B200): python tools/stamps_fused.py"""
import sys, ctypes, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'occlusions-4d_b200'))
import numpy as np, torch
from o4d import ops, _lib
from tests import configs
cfg = configs.C2_GREATER
dev = torch.device('cuda', 0)
_, dec = configs.build_modules(cfg, dev)
z = np.load(os.path.join(ROOT, 'tests', 'golden', 'c2_greater_seeded.npz'))
abstract, glob = torch.from_numpy(z['abstract']).to(dev), torch.from_numpy(z['glob']).to(dev)
q = configs.synthetic_queries(cfg)[:32768].contiguous().to(dev)
h = ctypes.CDLL(_lib.LIB_PATH)
for prec, mode in ((1, 1), (1, 2), (2, 1)):
    dec.o4d_precision = prec
    h.o4d_debug_set_fused_passes(mode)
    with torch.no_grad():
        scene = dec.o4d_scene(abstract, glob)
        dcfg, dparams = dec.o4d_config(), dec.o4d_params()
        out = torch.empty((q.shape[0], 9), device=dev)
        for _ in range(3):
            ops.decoder_forward(dcfg, dparams, scene, q, want_penult=False, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ops.decoder_forward(dcfg, dparams, scene, q, want_penult=False, out=out)
        e1.record(); torch.cuda.synchronize()
    buf=(ctypes.c_longlong*16)()
    h.o4d_debug_read(buf); v=list(buf)
    print('prec %d mode %d: %.3f ms/batch; fused stamps: start->loop %d, loop %d, loop_end->acc2 %d, epilogue %d, total %d' % (prec, mode, e0.elapsed_time(e1)/5, v[0]-v[4], v[1]-v[0], v[2]-v[1], v[3]-v[2], v[3]-v[4])); print('   epilogue parts: acc1 wait %d, stage %d, barriers %d, reduce %d, issue_v %d' % (v[5],v[6],v[7],v[8],v[9])); print('   conversion loop of thread 0 (13 chunks of its group): loop top (G_EMPTY wait + gather issue) %d, wait ACC1_FULL %d, tmem ld + gather wait + LDS %d, relu/split %d, tmem st + arrive %d' % (v[10],v[11],v[12],v[13],v[14]))
    h.o4d_debug_read_chain(buf); v=list(buf)   # last chain launch of the mini-batch (layer3 .. lin_out), one CTA
    print('   chain (last launch): MMA thread total %d cycles for %d items: waits acc-empty %d (epilogue-bound), full-stage %d (load-bound); producer waits: empty-stage %d (MMA-bound), dependency %d; epilogue warp 0: wait %d busy %d' % (v[0], v[7], v[1], v[2], v[3], v[4], v[5], v[6]))
