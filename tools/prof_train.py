#!/usr/bin/env python
"""Training-step driver for ncu: config-5-shape decoder frame (17,203 queries, M = 2124) forward + backward,
N passes.  Usage: python tools/prof_train.py [passes] [precision 1|2]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'occlusions-4d_b200')):
    sys.path.insert(0, p)
import torch
from tests import configs

passes = int(sys.argv[1]) if len(sys.argv) > 1 else 2
cfg = configs.C3_CARLA
dev = torch.device('cuda', 0)
_, dec = configs.build_modules(cfg, dev)
dec.train()
if len(sys.argv) > 2:
    for mod in dec.modules():
        if hasattr(mod, 'o4d_precision'):
            mod.o4d_precision = int(sys.argv[2])
g = torch.Generator().manual_seed(5)
m, e = 2124, cfg['implicit_args']['d_latent_local']
abstract = torch.cat([torch.rand(m, 3, generator=g) * 30, torch.randn(m, e, generator=g) * 0.5], 1).to(dev).requires_grad_(True)
glob = (torch.randn(128, generator=g) * 0.5).to(dev).requires_grad_(True)
query = torch.cat([torch.rand(17203, 3, generator=g) * 30, torch.full((17203, 1), 3.0)], 1).to(dev)
import time
for i in range(passes):
    t0 = time.perf_counter()
    out, _ = dec(query, abstract, glob, None)
    t1 = time.perf_counter()
    out.square().mean().backward()
    t2 = time.perf_counter()
    torch.cuda.synchronize()
    if os.environ.get('VERBOSE'):
        st = torch.cuda.memory_stats()
        print('pass %d: host forward %.1f ms, host backward %.1f ms, with sync %.1f ms; reserved %.2f GB, cudaMalloc calls %d, retries %d' % (
            i, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (time.perf_counter() - t0) * 1e3, st['reserved_bytes.all.current'] / 1e9,
            st['segment.all.allocated'], st['num_alloc_retries']))
    if i == passes - 2:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
e1.record()
torch.cuda.synchronize()
if passes >= 2:
    print('last pass: %.2f ms' % e0.elapsed_time(e1))
if os.environ.get('PROFILE'):      # kernel times of one more pass as they run back to back (not serialised as under ncu)
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        out, _ = dec(query, abstract, glob, None)
        out.square().mean().backward()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=25, max_name_column_width=60))
