"""GPU parity tests of the SURVEY.md 8f pieces, called through the C ABI: the sampler's 1-NN rejection /
compaction kernels and the whole sampler against the reference's golden vectors (draw for draw), the
dataset-side FPS subsampling against the oracle, and the fused loss heads (values and gradients) against the
reference's golden vectors.  Index / selection work is bit-exact; floating point carries its tolerance."""
import logging

import numpy as np
import pytest
import torch

from o4d import geometry as geo, loss as o4d_loss, ops
from oracle import cluster_ops, sampler_oracle as so
from tests import sampler_cases as sc
from tests.test_sampler_cpu import golden

pytestmark = pytest.mark.gpu
DEV = 'cuda'


# ------------------------------------------------------------------ filter_air_solid_gap / select_safely

@pytest.mark.parametrize('name', sorted(sc.FILTER_CASES))
def test_filter_matches_reference_golden(name):
    g = golden('sampler_golden.npz')
    n, m, d, radius, num_select = sc.FILTER_CASES[name]
    cand, target = sc.filter_inputs(name)
    rows, dists, ratio = geo.filter_air_solid_gap(cand.to(DEV), target.to(DEV), 700, radius)
    assert torch.equal(rows.cpu(), g['filter_%s_rows' % name])              # same survivors, same order
    torch.testing.assert_close(dists.cpu(), g['filter_%s_dists' % name], rtol=3e-7, atol=0)
    assert torch.equal(dists.cpu(), so.filter_air_solid_gap(cand, target, 0, radius)[1])   # bit-exact vs the oracle
    assert abs(float(ratio) - float(g['filter_%s_ratio' % name])) < 1e-7
    if num_select:
        sel, seld, count = geo.filter_select(cand.to(DEV), target.to(DEV), radius, num_select)
        assert torch.equal(sel.cpu(), g['filter_%s_sel' % name])
        assert int(count.item()) == rows.shape[0]
        assert torch.equal(seld.cpu(), so.select_safely(dists.cpu(), num_select))


def test_filter_strided_inputs_and_edges():
    g = torch.Generator().manual_seed(3)
    wide = torch.rand(1200, 12, generator=g) * 4 - 2
    cloud = torch.rand(900, 9, generator=g) * 4 - 2
    cand, target = wide[:, 2:7], cloud[:, :3]                                  # row strides 12 and 9
    want = so.filter_select(cand, target, 0.3, 700)
    got = geo.filter_select(wide.to(DEV)[:, 2:7], cloud.to(DEV)[:, :3], 0.3, 700)
    assert torch.equal(got[0].cpu(), want[0]) and torch.equal(got[1].cpu(), want[1])
    assert int(got[2].item()) == int(want[2])
    # nothing survives: zero rows, count 0 (the reference's select_safely would never return)
    rows, dists, count = geo.filter_select(cand.to(DEV), target.to(DEV), 100.0, 16)
    assert int(count.item()) == 0 and not rows.any() and not dists.any()
    rows, dists, ratio = geo.filter_air_solid_gap(cand.to(DEV), target.to(DEV), 0, 100.0)
    assert rows.shape == (0, 5) and dists.shape == (0,) and float(ratio) == 0.0
    # no candidates at all
    rows, dists, count = geo.filter_select(torch.zeros(0, 3, device=DEV), target.to(DEV), 0.3, 8)
    assert int(count.item()) == 0 and rows.shape == (8, 3) and not rows.any()


def test_filter_training_size_properties():
    """Config-5 scale (20k candidates against a 28,672-point target frame): every survivor is farther than the
    radius, every dropped candidate is not, order is preserved, the prefix is what select_safely returns."""
    g = torch.Generator().manual_seed(8)
    cand = torch.rand(20000, 3, generator=g) * torch.tensor([40.0, 32.0, 7.4]) + torch.tensor([0.0, -16.0, -1.0])
    target = torch.rand(28672, 3, generator=g) * torch.tensor([40.0, 32.0, 7.4]) + torch.tensor([0.0, -16.0, -1.0])
    radius = 0.3
    rows, dists, ratio = geo.filter_air_solid_gap(cand.to(DEV), target.to(DEV), 8192, radius)
    d_all = so.nn1_dist(cand, target, chunk=512)
    keep = d_all > radius
    assert torch.equal(rows.cpu(), cand[keep]) and torch.equal(dists.cpu(), d_all[keep])
    assert bool((dists > radius).all()) and 0.0 < float(ratio) < 1.0
    sel, seld, count = geo.filter_select(cand.to(DEV), target.to(DEV), radius, 10035)
    assert int(count.item()) == int(keep.sum())
    assert torch.equal(sel.cpu(), so.select_safely(cand[keep], 10035))


def test_filter_bounds_matches_reference_golden():
    g = golden('sampler_golden.npz')
    pcl = sc.bounds_input().to(DEV)
    assert torch.equal(geo.filter_pcl_bounds_carla_output_torch(pcl, min_z=-1.0, other_bounds=16.0, cube_mode=4).cpu(),
                       g['bounds_carla'])
    assert torch.equal(geo.filter_pcl_bounds_torch(pcl, -3.0, 7.5, -2.0, 2.0, 0.0, 1.5).cpu(), g['bounds_box'])
    assert geo.filter_pcl_bounds_torch(pcl, 100.0, 101.0).shape == (0, 11)


# ------------------------------------------------------------------ whole sampler

def _to_dev(inputs):
    frames, sizes, valo, num_valo = inputs
    return [f.to(DEV) for f in frames], [s.to(DEV) for s in sizes], valo.to(DEV), num_valo.to(DEV)


@pytest.mark.parametrize('name', sorted(sc.SAMPLER_CASES))
def test_sampler_reproduces_reference_golden(name, monkeypatch):
    """Seeded run on the GPU = the reference's CPU run sample for sample.  The blind cuboid samples come from the
    device generator in both implementations (torch.rand(device=...)); the golden run was on CPU, so that one
    draw is redirected to the CPU generator here."""
    blind = geo.sample_implicit_points_blind_torch
    monkeypatch.setattr(geo, 'sample_implicit_points_blind_torch',
                        lambda kind, num, mode, bounds, min_z, device: blind(kind, num, mode, bounds, min_z, 'cpu').to(device))
    g = golden('sampler_golden.npz')
    case = sc.SAMPLER_CASES[name]
    smp = geo.GuidedImplicitPointSampler(logging.getLogger('test'), **case['kwargs'])
    sc.seed_all(case['seed'])
    res = smp(*_to_dev(sc.sampler_inputs(name)), case['time_idx'])
    for key, val in zip(sc.SAMPLER_OUTPUTS, res):
        ref = g['sampler_%s_%s' % (name, key)]
        assert val.shape == ref.shape, (name, key)
        assert torch.equal(val.cpu(), ref), (name, key, (val.cpu() - ref).abs().max())
    assert res[0].is_cuda and res[1].is_cuda and res[2].is_cuda and res[3].is_cuda


def test_sampler_device_rng_invariants():
    """device_rng=True draws on the GPU generator: no golden, but the sampler's contract holds -- air queries are
    farther than the radius from every target point of the frame, solid queries sit within half a radius of one."""
    name = 'greater_low_moving'
    case = sc.SAMPLER_CASES[name]
    frames, sizes, valo, num_valo = sc.sampler_inputs(name)
    smp = geo.GuidedImplicitPointSampler(logging.getLogger('test'), device_rng=True, **case['kwargs'])
    sc.seed_all(1)
    torch.cuda.manual_seed(1)
    solid_in, air_in, solid_tgt, air_tgt, _, _ = smp(*_to_dev((frames, sizes, valo, num_valo)), case['time_idx'])
    r = case['kwargs']['point_occupancy_radius']
    for b in range(frames[0].shape[0]):
        cloud = frames[case['time_idx']][b, :int(sizes[case['time_idx']][b])]
        assert bool((so.nn1_dist(air_in[b].cpu(), cloud) > r).all())
        assert bool((so.nn1_dist(solid_in[b].cpu(), cloud) <= r / 2.0 + 1e-6).all())
    assert bool((solid_in[..., 3] == case['time_idx']).all()) and bool((air_in[..., 3] == case['time_idx']).all())
    assert bool((solid_tgt[..., 0] == 1).all()) and bool((air_tgt[..., 0] == 0).all()) and bool((air_tgt[..., 1:] == -1).all())


# ------------------------------------------------------------------ dataset-side FPS subsampling

def test_subsample_farthest_point_matches_oracle():
    g = torch.Generator().manual_seed(4)
    pcl = torch.rand(9000, 8, generator=g)
    pcl[:, :3] = pcl[:, :3] * torch.tensor([10.0, 10.0, 6.0]) + torch.tensor([-5.0, -5.0, -1.0])
    n_desired = 4096
    sub = geo.subsample_pad_pcl_torch(pcl.to(DEV), n_desired, sample_mode='farthest_point')
    order = cluster_ops.fps_segment(pcl[:, :3], n_desired, start=0)
    assert sub.shape == (n_desired, 8)
    assert torch.equal(sub.cpu(), pcl[torch.sort(order)[0]])
    padded = geo.subsample_pad_pcl_torch(pcl.to(DEV)[None], 9100)
    assert padded.shape == (1, 9100, 8) and not padded[0, 9000:].any()


# ------------------------------------------------------------------ loss heads

@pytest.mark.parametrize('name', sorted(sc.LOSS_CASES))
def test_loss_heads_match_reference_golden(name):
    g = golden('loss_golden.npz')
    case = sc.LOSS_CASES[name]
    output, target = sc.loss_inputs(name)
    losses = o4d_loss.MyLosses('train', None, False, 1.0, 1.0, 1.0, 1.0, case['color_mode'],
                               case['semantic_classes'], 1, 0)
    heads = {'dens': losses.implicit_density_loss, 'rgb': losses.implicit_color_loss,
             'segm': losses.implicit_segm_loss, 'track': losses.implicit_track_loss}
    checked = 0
    for key, fn in heads.items():
        gk = 'loss_%s_%s' % (name, key)
        if gk not in g:
            continue
        o = output.to(DEV).requires_grad_(True)
        val = fn(o[None], target.to(DEV)[None])
        val.backward()
        torch.testing.assert_close(val.cpu(), g[gk], rtol=3e-6, atol=0)
        torch.testing.assert_close(o.grad.cpu(), g[gk + '_grad'], rtol=2e-5, atol=1e-9)
        checked += 1
    assert checked >= 2
    # all heads in one pass, weighted: the gradient is the weighted sum of the per-head gradients
    w = torch.tensor([0.7, 1.3, 0.4, 0.9])
    o = output.to(DEV).requires_grad_(True)
    all4 = o4d_loss.implicit_loss_heads(o, target.to(DEV), case['color_mode'], case['semantic_classes'], case['track'])
    (all4 * w.to(DEV)).sum().backward()
    want = sum(w[i] * g['loss_%s_%s_grad' % (name, k)] for i, k in enumerate(('rgb', 'dens', 'segm', 'track'))
               if 'loss_%s_%s' % (name, k) in g)
    torch.testing.assert_close(o.grad.cpu(), want, rtol=2e-5, atol=2e-9)
    # run-to-run deterministic
    again = o4d_loss.implicit_loss_heads(output.to(DEV), target.to(DEV), case['color_mode'], case['semantic_classes'],
                                         case['track'])
    assert torch.equal(again, all4.detach())


def test_loss_per_example_matches_oracle_and_handles_empty_heads():
    names = ['hsv_seg']
    case = sc.LOSS_CASES['hsv_seg']
    output, target = sc.loss_inputs('hsv_seg')
    frames_o = [output[:2000].to(DEV)[None], output[2000:].to(DEV)[None]]
    frames_t = [target[:2000].to(DEV)[None], target[2000:].to(DEV)[None]]
    losses = o4d_loss.MyLosses('train', None, False, 1.0, 1.0, 1.0, 0.0, case['color_mode'],
                               case['semantic_classes'], 2, 0)
    rgb, dens, segm, track = losses.per_example([torch.zeros(1, 8, 11)], [8], frames_o, frames_t)
    assert track is None
    want = [so.implicit_losses(output[s], target[s], case['color_mode'], case['semantic_classes'])
            for s in (slice(0, 2000), slice(2000, None))]
    for got, key in ((rgb, 'rgb'), (dens, 'dens'), (segm, 'segm')):
        torch.testing.assert_close(got.cpu(), torch.stack([w[key] for w in want]).mean(), rtol=3e-6, atol=0)
    total = losses.entire_batch(0, rgb, dens, segm, track, None, None, None)
    torch.testing.assert_close(total[0].cpu(), (rgb + dens + segm).cpu())
    # a frame of air points only: colour / segmentation / tracking heads have nothing to supervise -> NaN, like torch
    air = target.clone()
    air[:, 0] = 0.0
    air[:, 1:] = -1.0
    vals = o4d_loss.implicit_loss_heads(output.to(DEV), air.to(DEV), 'hsv', 13, True).cpu()
    assert torch.isnan(vals[0]) and torch.isfinite(vals[1]) and torch.isnan(vals[2]) and torch.isnan(vals[3])
    del names
