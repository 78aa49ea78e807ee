#!/usr/bin/env python
"""Kernel-time table (torch.profiler) of one BASELINE config-5 training step per GPU: encoder + 4 decoder frames
forward/backward + fused loss heads + AdamW, the step bench.py's train_step_bench times.
Usage (on a B200): python tools/prof_train_step.py [precision 1|2]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'occlusions-4d_b200'))
import torch
from torch.profiler import profile, ProfilerActivity
from o4d import loss as o4d_loss
from tests import configs
prec = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cfg = configs.C3_CARLA
dev = torch.device('cuda', 0)
frames, per_frame = 4, 17203
g = torch.Generator().manual_seed(1830)
pcl = configs.synthetic_cloud(cfg).to(dev)
lo, hi = torch.tensor([0.0, -16.0, -1.0]), torch.tensor([40.0, 16.0, 6.4])
queries = []
for f in range(frames):
    q = torch.rand(per_frame, 4, generator=g); q[:, :3] = q[:, :3] * (hi - lo) + lo; q[:, 3] = float(f)
    queries.append(q.to(dev))
target = torch.rand(frames, per_frame, 6, generator=g)
target[..., 0] = (target[..., 0] > 0.5).float(); target[..., 4] = 0.0
target[..., 5] = torch.randint(0, 13, (frames, per_frame), generator=g).float()
target = target.to(dev)
weights = torch.tensor([1.0, 1.0, 0.6, 1.0], device=dev)
enc, dec = configs.build_modules(cfg, dev)
enc.train(); dec.train()
for m in list(enc.modules()) + list(dec.modules()):
    if hasattr(m, 'o4d_precision'):
        m.o4d_precision = prec
opt = torch.optim.AdamW(list(enc.parameters()) + list(dec.parameters()), lr=1e-4)

def step(parts=None):
    ev = lambda: (torch.cuda.Event(enable_timing=True))
    marks = []
    def mark(name):
        if parts is not None:
            e = ev(); e.record(); marks.append((name, e))
    mark('start')
    opt.zero_grad(set_to_none=True)
    abstract, glob, _ = enc(pcl[None], False)
    mark('encoder forward')
    total = 0.0
    for f in range(frames):
        out, _ = dec(queries[f], abstract[0], glob[0], None)
        heads = o4d_loss.implicit_loss_heads(out, target[f], 'rgb', 13, True)
        total = total + (heads * weights).sum()
    mark('4 decoder frames forward + loss heads')
    (total / frames).backward()
    mark('backward')
    opt.step()
    mark('AdamW')
    if parts is not None:
        torch.cuda.synchronize()
        for (n0, e0), (n1, e1) in zip(marks[:-1], marks[1:]):
            parts.append((n1, e0.elapsed_time(e1)))

for _ in range(3):
    step()
torch.cuda.synchronize()
parts = []
step(parts)
print('precision %d; parts of one step (ms): %s; total %.2f' % (prec, ', '.join('%s %.2f' % p for p in parts), sum(p[1] for p in parts)))
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=40, max_name_column_width=70))
