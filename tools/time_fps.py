#!/usr/bin/env python
"""Times farthest point sampling at the three encoder levels (O4D_FPS_CLUSTER=0: single-SM kernel).
Usage (on a B200): python tools/time_fps.py"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'occlusions-4d_b200'))
import torch
from o4d import ops
torch.manual_seed(0)
for n, n_out in ((14336, 4779), (4779, 1593), (1593, 531)):
    p = (torch.rand(n, 3) * 10 - 5).cuda()
    for _ in range(2): ops.fps(p, n_out, 0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): ops.fps(p, n_out, 0)
    e1.record(); torch.cuda.synchronize()
    print('cluster=%s n %d -> %d: %.3f ms (%.2f us/pick)' % (os.environ.get('O4D_FPS_CLUSTER','1'), n, n_out, e0.elapsed_time(e1)/5, e0.elapsed_time(e1)/5*1e3/n_out))
