#!/usr/bin/env python
"""kNN launch-shape tuning: time the shapes of the encoder / decoder for the sub-lane count in O4D_KNN_S.
Usage (on a B200): for s in 4 8 16 32; do O4D_KNN_S=$s python tools/time_knn.py; done"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'occlusions-4d_b200'))
import torch
from o4d import ops
dev = torch.device('cuda', 0)
g = torch.Generator(device='cuda').manual_seed(1)
res = []
for nq, m, k, k2 in ((17203, 2124, 14, 8), (32768, 2124, 14, 8), (14336, 14336, 16, 0), (4779, 14336, 12, 0), (4779, 4779, 16, 0), (1593, 4779, 12, 0), (1593, 1593, 16, 0), (531, 531, 16, 0), (2124, 2124, 16, 0)):
    q = torch.rand(nq, 3, device=dev, generator=g) * 30
    r = torch.rand(m, 3, device=dev, generator=g) * 30
    f = (lambda: ops.knn_two_lists(q, r, k, k2)) if k2 else (lambda: ops.knn(q, r, k))
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    res.append('%dx%d k%d%s: %.0f us' % (nq, m, k, '+%d' % k2 if k2 else '', e0.elapsed_time(e1) * 100))
print('S=%s  ' % os.environ.get('O4D_KNN_S', 'auto') + ' | '.join(res))
