#!/usr/bin/env python
"""Decoder-only driver for ncu: N forward passes of ONE query mini-batch (32768 queries of the config-2
workload) on the golden scene, no encoder, so kernel launch indices are predictable:
per pass 24 linear_tc_kernel launches (lin_in; 6 x [lin_z, fc_0, fc_1]; per cross layer Qa + layer3; ...),
2 attn_fused_kernel launches.  Usage: python tools/prof_decoder.py [passes] [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'occlusions-4d_b200')):
    sys.path.insert(0, p)
import numpy as np
import torch
from o4d import ops
from tests import configs

passes = int(sys.argv[1]) if len(sys.argv) > 1 else 3
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
cfg = configs.C2_GREATER
dev = torch.device('cuda', 0)
_, dec = configs.build_modules(cfg, dev)
z = np.load(os.path.join(ROOT, 'tests', 'golden', 'c2_greater_seeded.npz'))
abstract, glob = torch.from_numpy(z['abstract']).to(dev), torch.from_numpy(z['glob']).to(dev)
q = configs.synthetic_queries(cfg)[:batch].contiguous().to(dev)
with torch.no_grad():
    scene = dec.o4d_scene(abstract, glob)
    dcfg, dparams = dec.o4d_config(), dec.o4d_params()
    out = torch.empty((q.shape[0], cfg['implicit_args']['d_out']), device=dev)
    for _ in range(passes):
        ops.decoder_forward(dcfg, dparams, scene, q, want_penult=False, out=out)
    torch.cuda.synchronize()
print('done', float(out.abs().sum()))
