"""CPU tests of the SURVEY.md 8f pieces (training-time sampler, dataset-side subsampling, loss heads):
the oracle against the reference's golden vectors (and against the reference itself in the build
container), the sampler's host logic draw-for-draw against the reference, and the loss-head kernel
arithmetic compiled for the host (csrc/loss_core.cuh, the source the CUDA kernels include)."""
import ctypes
import logging
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import ref_loader, sampler_oracle as so
from tests import sampler_cases as sc

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(REPO, 'tests', 'golden')


def golden(name):
    g = np.load(os.path.join(GOLDEN, name))
    return {k: torch.from_numpy(np.asarray(g[k])) for k in g.files}


# ------------------------------------------------------------------ oracle vs golden

def test_oracle_filter_matches_reference_golden():
    g = golden('sampler_golden.npz')
    for name, (n, m, d, radius, num_select) in sc.FILTER_CASES.items():
        cand, target = sc.filter_inputs(name)
        rows, dists, ratio = so.filter_air_solid_gap(cand, target, 700, radius)
        assert torch.equal(rows, g['filter_%s_rows' % name]), name          # same survivors, same order
        # torch.linalg.norm rounds the same quantity within 2 ulp of sqrt((dx2 + dy2) + dz2)
        torch.testing.assert_close(dists, g['filter_%s_dists' % name], rtol=3e-7, atol=0)
        assert abs(float(ratio) - float(g['filter_%s_ratio' % name])) < 1e-7
        if num_select:
            sel, seld, count = so.filter_select(cand, target, radius, num_select)
            assert torch.equal(sel, g['filter_%s_sel' % name]), name
            assert int(count) == rows.shape[0]
            assert torch.equal(seld, so.select_safely(dists, num_select))


def test_oracle_bounds_match_reference_golden():
    g = golden('sampler_golden.npz')
    pcl = sc.bounds_input()
    carla = so.filter_pcl_bounds(pcl, 0.0, 40.0, -16.0, 16.0, -1.0, 16.0 * 0.4)
    assert torch.equal(carla, g['bounds_carla'])
    assert (carla[:3, :3] == pcl[:3, :3]).all() and carla.shape[0] < pcl.shape[0]   # faces kept, 40.000004 dropped
    assert torch.equal(so.filter_pcl_bounds(pcl, -3.0, 7.5, -2.0, 2.0, 0.0, 1.5), g['bounds_box'])


def _oracle_losses_and_grads(name):
    case = sc.LOSS_CASES[name]
    output, target = sc.loss_inputs(name)
    res = {}
    o = output.clone().requires_grad_(True)
    for key, val in so.implicit_losses(o, target, case['color_mode'], case['semantic_classes']).items():
        if key == 'track' and not case['track']:
            continue
        grad, = torch.autograd.grad(val, o, retain_graph=True)
        res[key] = (val.detach(), grad)
    return res


def test_oracle_loss_heads_match_reference_golden():
    g = golden('loss_golden.npz')
    for name in sc.LOSS_CASES:
        res = _oracle_losses_and_grads(name)
        keys = [h for h in ('rgb', 'dens', 'segm', 'track') if 'loss_%s_%s' % (name, h) in g]
        assert sorted(keys) == sorted(res), (name, keys, sorted(res))
        for key, (val, grad) in res.items():
            torch.testing.assert_close(val, g['loss_%s_%s' % (name, key)], rtol=2e-6, atol=0)
            torch.testing.assert_close(grad, g['loss_%s_%s_grad' % (name, key)], rtol=1e-5, atol=1e-9)


# ------------------------------------------------------------------ oracle vs the reference itself

@pytest.mark.reference
def test_oracle_sampler_ops_match_reference_run():
    geo = ref_loader.load()['geometry']
    g = torch.Generator().manual_seed(5)
    cand = torch.rand(900, 5, generator=g) * 4 - 2
    target = torch.rand(1300, 3, generator=g) * 4 - 2
    a = geo.filter_air_solid_gap(cand, target, 400, 0.2)
    b = so.filter_air_solid_gap(cand, target, 400, 0.2)
    assert torch.equal(a[0], b[0]) and float(a[2]) == float(b[2])
    torch.testing.assert_close(a[1], b[1], rtol=3e-7, atol=0)
    assert torch.equal(geo.filter_pcl_bounds_torch(cand, -1.0, 1.5, -0.5, 0.5, -2.0, 0.0),
                       so.filter_pcl_bounds(cand, -1.0, 1.5, -0.5, 0.5, -2.0, 0.0))
    utils = ref_loader.load()['utils']
    rgb = torch.rand(500, 3, generator=g)
    rgb[::9] = rgb[::9][:, [1, 1, 1]]
    assert torch.equal(utils.rgb_to_hsv(rgb), so.rgb_to_hsv(rgb))


# ------------------------------------------------------------------ sampler host logic, draw for draw

@pytest.fixture
def geometry_on_cpu(monkeypatch):
    """o4d.geometry with its three device ops replaced by the CPU oracle: what is left is the host logic
    (bias shares, pool construction, draw order), which must reproduce the reference sample for sample."""
    from o4d import geometry as geo
    monkeypatch.setattr(geo, 'filter_select', so.filter_select)
    monkeypatch.setattr(geo, 'filter_air_solid_gap', so.filter_air_solid_gap)
    monkeypatch.setattr(geo, 'filter_pcl_bounds_torch', so.filter_pcl_bounds)
    return geo


@pytest.mark.parametrize('name', sorted(sc.SAMPLER_CASES))
def test_sampler_host_logic_reproduces_reference(geometry_on_cpu, name):
    g = golden('sampler_golden.npz')
    case = sc.SAMPLER_CASES[name]
    smp = geometry_on_cpu.GuidedImplicitPointSampler(logging.getLogger('test'), **case['kwargs'])
    sc.seed_all(case['seed'])
    res = smp(*sc.sampler_inputs(name), case['time_idx'])
    assert len(res) == len(sc.SAMPLER_OUTPUTS)
    for key, val in zip(sc.SAMPLER_OUTPUTS, res):
        ref = g['sampler_%s_%s' % (name, key)]
        assert val.shape == ref.shape, (name, key)
        assert torch.equal(val, ref), (name, key, (val - ref).abs().max())


def test_sampler_rejects_small_clouds(geometry_on_cpu):
    case = sc.SAMPLER_CASES['greater_none']
    frames, sizes, valo, num_valo = sc.sampler_inputs('greater_none')
    sizes[1][0] = 100
    smp = geometry_on_cpu.GuidedImplicitPointSampler(logging.getLogger('test'), **case['kwargs'])
    with pytest.raises(RuntimeError, match='cur_tgt_pcl_count'):
        smp(frames, sizes, valo, num_valo, 1)


def test_subsample_pad_host_branches():
    from o4d import geometry as geo
    pcl = torch.arange(40, dtype=torch.float32).reshape(1, 10, 4)
    padded = geo.subsample_pad_pcl_torch(pcl, 14)
    assert padded.shape == (1, 14, 4) and torch.equal(padded[:, :10], pcl) and (padded[:, 10:] == 0).all()
    assert geo.subsample_pad_pcl_torch(pcl[0], 10).shape == (10, 4)
    with pytest.raises(RuntimeError, match='Too few input points'):
        geo.subsample_pad_pcl_torch(pcl, 14, subsample_only=True)
    np.random.seed(3)
    sub = geo.subsample_pad_pcl_torch(pcl, 6, sample_mode='random')
    np.random.seed(3)
    inds = np.sort(np.random.choice(np.arange(10), 6, replace=False))
    assert torch.equal(sub, pcl[:, inds])


@pytest.mark.reference
def test_subsample_pad_random_matches_reference_run():
    from o4d import geometry as geo
    ref_geo = ref_loader.load()['geometry']
    g = torch.Generator().manual_seed(9)
    # 2-D clouds: the reference's subsampling branch asserts on shape[0] and only works un-batched
    pcl = torch.rand(300, 7, generator=g)
    pcl[:, 5] = torch.randint(0, 12, (300,), generator=g).float()
    for kwargs in (dict(), dict(retain_vehped=True, segm_idx=5)):
        np.random.seed(4)
        a = ref_geo.subsample_pad_pcl_torch(pcl, 120, sample_mode='random', **kwargs)
        np.random.seed(4)
        b = geo.subsample_pad_pcl_torch(pcl, 120, sample_mode='random', **kwargs)
        assert torch.equal(a.reshape(-1, 7), b.reshape(-1, 7))


# ------------------------------------------------------------------ loss-head kernel arithmetic on the host

@pytest.fixture(scope='module')
def host_core(tmp_path_factory):
    so_path = str(tmp_path_factory.mktemp('host_core') / 'loss_core_host.so')
    src = os.path.join(REPO, 'tests', 'host_core', 'loss_core_host.cpp')
    subprocess.run(['g++', '-O2', '-ffp-contract=off', '-shared', '-fPIC', '-o', so_path, src], check=True)
    return ctypes.CDLL(so_path)


@pytest.mark.parametrize('name', sorted(sc.LOSS_CASES))
def test_loss_kernel_arithmetic_on_host_matches_reference_golden(host_core, name):
    g = golden('loss_golden.npz')
    case = sc.LOSS_CASES[name]
    output, target = sc.loss_inputs(name)
    n, width = output.shape
    o, t = output.numpy(), target.numpy()
    mode = {'rgb': 0, 'rgb_nosigmoid': 0, 'hsv': 1, 'bins': 2}[case['color_mode']]
    ti = so.track_idx(case['color_mode']) if case['track'] else -1
    P = ctypes.c_void_p
    losses, stats = np.zeros(4, np.float32), np.zeros(12, np.float64)
    host_core.host_loss_forward(P(o.ctypes.data), ctypes.c_long(n), width, P(t.ctypes.data), mode,
                                case['semantic_classes'], ti, P(losses.ctypes.data), P(stats.ctypes.data))
    checked = 0
    for slot, key in enumerate(('rgb', 'dens', 'segm', 'track')):
        gk = 'loss_%s_%s' % (name, key)
        if gk not in g:
            assert losses[slot] == 0.0
            continue
        assert abs(losses[slot] - float(g[gk])) <= 2e-6 * abs(float(g[gk])), (name, key)
        w = np.zeros(4, np.float32)
        w[slot] = 1.0
        d = np.full((n, width), np.nan, np.float32)
        host_core.host_loss_backward(P(o.ctypes.data), ctypes.c_long(n), width, P(t.ctypes.data), mode,
                                     case['semantic_classes'], ti, P(stats.ctypes.data), P(w.ctypes.data),
                                     P(d.ctypes.data))
        np.testing.assert_allclose(d, g[gk + '_grad'].numpy(), rtol=1e-5, atol=1e-9)
        checked += 1
    assert checked >= 2


# ------------------------------------------------------------------ C ABI argument checks (no GPU work)

def test_new_entry_points_validate_arguments_without_gpu():
    from o4d import _lib
    L = _lib.lib()
    assert L.o4d_filter_workspace_bytes(1000) >= 1000 * 8
    assert L.o4d_filter_air_solid_gap_f32(None, 10, 2, 2, None, 5, 3, 0.1, 0, None, 2, None, None, None, 0, None) == -1
    assert b'd >= 3' in L.o4d_last_error()
    assert L.o4d_filter_air_solid_gap_f32(None, 10, 3, 3, None, 5, 3, 0.1, 0, None, 3, None, None, None, 0, None) == -1
    assert b'count_out' in L.o4d_last_error()
    lo = (ctypes.c_float * 3)(0, 0, 0)
    assert L.o4d_filter_bounds_f32(None, 4, 3, 2, lo, lo, None, 3, None, None, 0, None) == -1
    assert L.o4d_implicit_loss_workspace_bytes(100000) >= 12 * 8
    assert L.o4d_implicit_loss_forward_f32(None, 8, 5, 5, None, 6, 7, 0, 4, None, None, None, 0, None) == -1
    assert b'color_mode' in L.o4d_last_error()
    assert L.o4d_implicit_loss_forward_f32(None, 8, 10, 10, None, 6, 1, 0, -1, None, None, None, 0, None) == -1
    assert b'too narrow' in L.o4d_last_error()                       # hsv needs 15 columns
    assert L.o4d_implicit_loss_backward_f32(None, 8, 5, 5, None, 5, 0, 0, 4, None, None, None, 5, None) == -1


def test_loss_module_surface():
    from o4d import loss as o4d_loss
    assert [o4d_loss.get_track_idx(m) for m in ('rgb', 'rgb_nosigmoid', 'hsv', 'bins')] == [4, 4, 15, 10]
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        o4d_loss.implicit_loss_heads(torch.zeros(4, 5), torch.zeros(4, 6))


def test_inference_driver_has_no_cpu_path():
    from o4d import inference
    with pytest.raises(AssertionError, match='CUDA'):
        inference.query_frame(None, torch.zeros(8, 291), torch.zeros(128), torch.zeros(16, 4))
    # 'random' queries keep the reference's numpy draws
    np.random.seed(2)
    a = inference.query_points(50, -1.0, 5.0, 3, 'greater', 4, 'random', 'cpu')
    np.random.seed(2)
    from o4d import geometry as geo
    b = geo.sample_implicit_points_blind_numpy(50, -1.0, 5.0, 3, 'greater', 4, 'random')
    assert np.array_equal(a.numpy(), b) and a.shape == (50, 4)


@pytest.mark.reference
def test_compat_geometry_shim_keeps_reference_names_and_overrides_the_sampler_path():
    """compat/geometry.py first on sys.path: the reference's own helpers stay importable, the sampler path is o4d's."""
    code = r'''
import importlib, os, sys, warnings
warnings.simplefilter('ignore')
repo, ref = sys.argv[1], sys.argv[2]
sys.path.insert(0, repo)
sys.path.insert(0, os.path.join(repo, 'oracle', 'ref_stubs'))
os.chdir(ref)
sys.path.insert(0, ref)
sys.path.insert(0, os.path.join(repo, 'occlusions-4d_b200', 'compat'))
importlib.import_module('__init__')
import geometry, torch
assert geometry.__file__.endswith(os.path.join('compat', 'geometry.py'))
assert geometry.GuidedImplicitPointSampler.__module__ == 'o4d.geometry'
assert geometry.filter_air_solid_gap.__module__ == 'o4d.geometry'
assert geometry.point_cloud_from_rgbd.__module__ == '_reference_geometry'
assert geometry.subsample_pad_pcl_torch(torch.rand(10, 4), 12).shape == (12, 4)      # CPU cloud: reference path
assert geometry.filter_pcl_bounds_torch(torch.rand(10, 4)).shape == (10, 4)
import loss
print('ok')
'''
    out = subprocess.run([os.sys.executable, '-c', code, REPO, ref_loader.REF_ROOT], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip().endswith('ok'), out.stderr[-2000:]
