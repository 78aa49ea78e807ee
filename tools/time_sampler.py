"""Timing of the SURVEY.md 8f pieces at BASELINE.json configs[4] shape (CARLA train step: 28,672-point target
frames, 7,168 solid + 10,035 air queries per frame, 17,203 x 18 logits per frame), CUDA events, one JSON line.

    python tools/time_sampler.py > gpurun_out/sampler_timing.json
"""
import json
import logging
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'occlusions-4d_b200'))


def timed(fn, iters=20, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    from o4d import geometry as geo, loss as o4d_loss, _lib
    dev = torch.device('cuda')
    g = torch.Generator().manual_seed(1830)
    lo, hi = torch.tensor([0.0, -16.0, -1.0]), torch.tensor([40.0, 16.0, 6.4])
    frames = []
    for _ in range(4):
        f = torch.rand(1, 28672, 11, generator=g)
        f[0, :, :3] = f[0, :, :3] * (hi - lo) + lo
        f[0, :, 5] = torch.randint(0, 13, (28672,), generator=g).float()
        frames.append(f.to(dev))
    sizes = [torch.tensor([28672]) for _ in range(4)]
    valo, num_valo = torch.zeros(1, 4), torch.zeros(1, dtype=torch.int64)
    res = {'gpu': torch.cuda.get_device_name(0)}

    cand = (torch.rand(20070, 3, generator=g) * (hi - lo) + lo).to(dev)
    target = frames[0][0, :, :3]
    L = _lib.lib()
    n0 = L.o4d_launch_count()
    geo.filter_select(cand, target, 0.2, 10035)
    res['filter_select_launches'] = int(L.o4d_launch_count() - n0)
    ms = timed(lambda: geo.filter_select(cand, target, 0.2, 10035))
    res['filter_select_ms'] = ms
    res['filter_select_pair_evals_per_s'] = 20070 * 28672 / (ms / 1e3)

    for bias, rng in (('none', True), ('none', False), ('low_moving', True)):
        smp = geo.GuidedImplicitPointSampler(
            logging.getLogger('t'), min_z=-1.0, cube_bounds=16.0, point_occupancy_radius=0.2, num_solid=7168,
            num_air=10035, predict_segmentation=True, semantic_classes=13, data_kind='carla',
            point_sample_bias=bias, cube_mode=4, device_rng=rng)
        res['sampler_ms_per_frame[%s,%s]' % (bias, 'device_rng' if rng else 'reference_rng')] = timed(
            lambda: smp(frames, sizes, valo, num_valo, 1), iters=10)

    out = torch.randn(17203, 18, generator=g).to(dev).requires_grad_(True)
    tgt = torch.full((17203, 6), -1.0)
    tgt[:7168, 0] = 1.0
    tgt[7168:, 0] = 0.0
    tgt[:7168, 1:4] = torch.rand(7168, 3, generator=g)
    tgt[:7168, 4] = 0.0
    tgt[:7168, 5] = torch.randint(0, 13, (7168,), generator=g).float()
    tgt = tgt.to(dev)
    w = torch.ones(4, device=dev)

    def fused():
        out.grad = None
        (o4d_loss.implicit_loss_heads(out, tgt, 'rgb', 13, True) * w).sum().backward()

    def eager():   # the reference's formulation (loss.py:50-194) with torch ops on the same GPU
        out.grad = None
        F = torch.nn.functional
        o, t = out[None], tgt[None]
        dens = F.binary_cross_entropy_with_logits(o[..., 0], t[..., 0])
        m = torch.logical_and(t[..., 0] >= 0.1, t[..., 1] >= 0.0)
        rgb = F.l1_loss(o[m][..., 1:4], t[m][..., 1:4])
        st = t[..., -1].type(torch.int64)
        sm = st >= 0
        segm = F.cross_entropy(o[..., -13:][sm], st[sm])
        tm = torch.logical_and(t[..., 0] >= 0.1, t[..., 4] >= 0.0)
        track = F.binary_cross_entropy_with_logits(o[tm][..., 4], t[tm][..., 4])
        (rgb + dens + segm + track).backward()

    res['loss_heads_fwd_bwd_ms'] = timed(fused)
    res['loss_heads_torch_eager_fwd_bwd_ms'] = timed(eager)
    print(json.dumps(res))


if __name__ == '__main__':
    main()
