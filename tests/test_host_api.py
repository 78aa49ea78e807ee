"""CPU tests of the host side: C-ABI library loads and exports every symbol include/o4d.h
declares, module constructors / state_dict layout match the reference's checkpoint
contract, call-form tolerance, loud failure without CUDA, sharding arithmetic."""
import ctypes
import os
import re

import pytest
import torch

import o4d
from o4d import _lib, parallel
from tests import configs

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(REPO, 'include', 'o4d.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(o4d_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_header_symbol():
    names = header_symbols()
    assert len(names) >= 20
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), 'libo4d.so does not export %s' % n
    # and the binding declares a signature for each of them
    assert sorted(_lib.SIGNATURES) == names


def test_library_probes():
    L = _lib.lib()
    assert L.o4d_abi_version() == 2
    assert L.o4d_has_tcgen05() in (0, 1)
    assert L.o4d_last_error() is not None


def test_config_structs_match_header():
    text = open(os.path.join(REPO, 'include', 'o4d.h')).read()
    for cname, cls in (('o4d_encoder_config', _lib.EncoderConfig), ('o4d_decoder_config', _lib.DecoderConfig)):
        body = re.search(r'typedef struct %s \{(.*?)\} %s;' % (cname, cname), text, flags=re.S).group(1)
        fields = re.findall(r'int32_t\s+(\w+);', body)
        assert fields == [f[0] for f in cls._fields_]


def test_argument_validation_without_gpu():
    """Entry points reject bad arguments before touching the device (no compute here)."""
    L = _lib.lib()
    assert L.o4d_knn_f32(None, 4, 3, None, 4, 3, 2, 0, None, None, None) == -1
    assert b'knn' in L.o4d_last_error()
    cfg = _lib.DecoderConfig(d_in=4, d_hidden=416, d_out=9, d_latent=416, d_latent_local=288, n_blocks=6,
                             pos_encoding_freqs=8, num_local_features=8, cross_attn_neighbors=14,
                             cross_attn_layers=2, precision=0)
    assert L.o4d_decoder_num_params(ctypes.byref(cfg)) == 70
    assert L.o4d_decoder_scene_bytes(ctypes.byref(cfg), 531) > 531 * 416 * 4 * 4
    assert L.o4d_decoder_workspace_bytes(ctypes.byref(cfg), 32768, 531) > 0
    bad = _lib.DecoderConfig(d_in=3)
    assert L.o4d_decoder_num_params(ctypes.byref(bad)) == -1
    ecfg = _lib.EncoderConfig(d_in=8, d_feat=36, down_blocks=3, transition_factor=3, pt_num_neighbors=14,
                              down_neighbors=12, norm=0, abstract_levels=1, global_dim=128, precision=0)
    assert L.o4d_encoder_num_params(ctypes.byref(ecfg)) == 74
    assert L.o4d_encoder_num_abstract(ctypes.byref(ecfg), 14336) == 531
    assert L.o4d_encoder_num_abstract(ctypes.byref(ecfg), 2048) == 76
    ecfg.abstract_levels, ecfg.norm = 2, 1
    assert L.o4d_encoder_num_params(ctypes.byref(ecfg)) == 82
    assert L.o4d_encoder_num_abstract(ctypes.byref(ecfg), 14336) == 1593 + 531


GREATER_ENC_KEYS = 74
CARLA_ENC_KEYS = 82


def test_state_dict_layout_matches_checkpoint_contract():
    enc, dec = configs.build_modules(configs.C2_GREATER)
    ek = list(enc.state_dict().keys())
    assert len(ek) == GREATER_ENC_KEYS
    assert ek[:8] == ['pre_mlp.0.weight', 'pre_mlp.0.bias', 'pre_mlp.2.weight', 'pre_mlp.2.bias',
                      'global_mlp.0.weight', 'global_mlp.0.bias', 'global_mlp.2.weight', 'global_mlp.2.bias']
    assert ek[8:23] == ['blocks.0.layer1.weight', 'blocks.0.layer1.bias', 'blocks.0.layer2.to_q.weight',
                        'blocks.0.layer2.to_k.weight', 'blocks.0.layer2.to_v.weight',
                        'blocks.0.layer2.pos_mlp.0.weight', 'blocks.0.layer2.pos_mlp.0.bias',
                        'blocks.0.layer2.pos_mlp.2.weight', 'blocks.0.layer2.pos_mlp.2.bias',
                        'blocks.0.layer2.attn_mlp.0.weight', 'blocks.0.layer2.attn_mlp.0.bias',
                        'blocks.0.layer2.attn_mlp.2.weight', 'blocks.0.layer2.attn_mlp.2.bias',
                        'blocks.0.layer3.weight', 'blocks.0.layer3.bias']
    assert ek[23:25] == ['blocks.1.mlp.0.weight', 'blocks.1.mlp.0.bias']
    sd = enc.state_dict()
    assert tuple(sd['blocks.6.layer2.attn_mlp.0.weight'].shape) == (576, 288)
    assert tuple(sd['global_mlp.0.weight'].shape) == (128, 288)
    dk = list(dec.state_dict().keys())
    assert len(dk) == 70
    assert dk[:4] == ['lin_in.weight', 'lin_in.bias', 'lin_out.weight', 'lin_out.bias']
    assert dk[4] == 'blocks.0.fc_0.weight' and dk[28] == 'lin_z.0.weight' and dk[40] == 'pt_blocks.0.layer1.weight'
    dsd = dec.state_dict()
    assert tuple(dsd['lin_in.weight'].shape) == (416, 68)
    assert tuple(dsd['pt_blocks.1.layer2.to_k.weight'].shape) == (416, 288)
    assert tuple(dsd['pt_blocks.0.layer2.attn_mlp.0.weight'].shape) == (832, 416)
    # the table handed to the C ABI is the state_dict order
    assert [id(p) for p in dec.o4d_params()] == [id(p) for p in dec.parameters()]
    enc3, _ = configs.build_modules(configs.C3_CARLA)
    k3 = list(enc3.state_dict().keys())
    assert len(k3) == CARLA_ENC_KEYS
    assert k3[8:10] == ['abstract_skip_mlps.0.weight', 'abstract_skip_mlps.0.bias']
    assert 'blocks.1.mlp.1.weight' in k3 and tuple(enc3.state_dict()['abstract_skip_mlps.0.weight'].shape) == (288, 144)
    names = {id(p): n for n, p in enc3.named_parameters()}
    assert len(enc3.o4d_params()) == CARLA_ENC_KEYS and all(id(p) in names for p in enc3.o4d_params())
    assert len(set(id(p) for p in enc3.o4d_params())) == CARLA_ENC_KEYS


def test_use_pt_inds_and_unsupported_modes():
    _, dec = configs.build_modules(configs.C2_GREATER)
    assert dec.use_pt_inds == {2: 0, 4: 1}
    _, dec1 = configs.build_modules(configs.C1_GREATER)
    assert dec1.use_pt_inds == {0: 0}
    with pytest.raises(NotImplementedError):
        o4d.LocalPclResnetFC(num_local_features=8, cross_attn_layers=1, cr_attn_type='s', d_latent=32, d_hidden=32)
    with pytest.raises(ValueError):
        o4d.LocalPclResnetFC(num_local_features=8, cross_attn_layers=1, cr_attn_type='x', d_latent=32, d_hidden=32)
    with pytest.raises(NotImplementedError):
        o4d.PointCompletionNetV3(enable_decoder=True)
    with pytest.raises(ValueError):
        o4d.DownTransition(8, 16, norm_type='weird')


def test_cpu_inputs_fail_loudly():
    """No CPU fallback: CPU tensors must raise, not silently compute elsewhere."""
    enc, dec = configs.build_modules(configs.TINY_GREATER)
    with torch.no_grad():
        with pytest.raises(RuntimeError, match='CUDA only|move the module to CUDA'):
            enc(torch.zeros(1, 64, 8), False)
        with pytest.raises(RuntimeError, match='CUDA only|move the module to CUDA'):
            dec(torch.zeros(5, 4), torch.zeros(30, 3 + 64), torch.zeros(16), None)
        with pytest.raises(RuntimeError):
            o4d.kNN_torch(torch.zeros(1, 8, 3), torch.zeros(1, 8, 3), 2)


def test_training_path_has_no_cpu_fallback_either():
    """With grad enabled the modules take the autograd path (o4d/autograd.py); CPU tensors must
    still raise instead of silently differentiating through torch ops."""
    enc, dec = configs.build_modules(configs.TINY_GREATER)
    enc.train()
    dec.train()
    with pytest.raises(RuntimeError, match='CUDA only|move the module to CUDA'):
        enc(torch.zeros(1, 64, 8), False)
    with pytest.raises(RuntimeError, match='CUDA only|move the module to CUDA'):
        dec(torch.zeros(5, 4), torch.zeros(30, 3 + 64), torch.zeros(16), None)


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', '/nonexistent/libo4d.so')
    with pytest.raises(RuntimeError, match='no fallback'):
        _lib.lib()


def test_shard_ranges_cover_exactly():
    for n in (0, 1, 7, 534528, 2097152):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_grid_query_count_matches_numpy_sampler():
    """o4d_grid_query_count is host-only arithmetic (utils/geometry.py:1248-1250): same lattice as the numpy mirror."""
    import ctypes
    import numpy as np
    from o4d import geometry
    L = _lib.lib()
    for num, kind, bounds, mode in ((4096, 'greater', 5.0, 4), (524288, 'greater', 5.0, 4), (2097152, 'greater', 5.0, 4),
                                    (4096, 'carla', 16.0, 4), (524288, 'carla', 16.0, 4), (2097152, 'carla', 16.0, 4),
                                    (100000, 'carla', 20.0, 1), (77, 'carla', 20.0, 2), (1000, 'carla', 20.0, 3)):
        (x0, x1), (y0, y1), (z0, z1) = geometry.cuboid_bounds(-1.0, bounds, kind, mode)
        ext = (ctypes.c_double * 3)(x1 - x0, y1 - y0, z1 - z0)
        counts = (ctypes.c_int32 * 3)()
        total = L.o4d_grid_query_count(num, ext, counts)
        want = geometry.sample_implicit_points_blind_numpy(num, -1.0, bounds, 0, kind, mode, 'grid')
        assert total == want.shape[0] >= num
        per_axis = [len(np.unique(want[:, a])) for a in range(3)]
        assert list(counts) == per_axis
    assert L.o4d_grid_query_count(0, ext, counts) == -1


def test_inference_column_ops_follow_the_reference_loop():
    """eval/inference.py:218-243: which output columns get a sigmoid / clamp."""
    from o4d import geometry as G
    assert G.inference_column_ops(5) == [G.SIGMOID] * 4 + [G.KEEP]
    assert G.inference_column_ops(5, track_mode='yes') == [G.SIGMOID] * 5
    ops = G.inference_column_ops(33, color_mode='hsv', predict_segmentation=True, semantic_classes=13,
                                 track_mode='yes', output_track_idx=15)
    assert ops[0] == G.SIGMOID and ops[1:13] == [G.SIGMOID] * 12 and ops[13:15] == [G.CLAMP01] * 2
    assert ops[15] == G.SIGMOID and ops[20:] == [G.SIGMOID] * 13 and ops[16:20] == [G.KEEP] * 4
    assert G.inference_column_ops(10, color_mode='bins') == [G.SIGMOID] * 10


def test_grad_mode_routes_to_the_training_path():
    """With grad enabled the modules must take o4d/autograd.py (and still refuse CPU tensors); under
    no_grad they take the fused inference path."""
    from o4d.point_transformer_layer import _wants_grad
    enc, dec = configs.build_modules(configs.TINY_GREATER)
    x = torch.zeros(4, 8)
    assert _wants_grad(enc, x)
    with torch.no_grad():
        assert not _wants_grad(enc, x)
    for p in dec.parameters():
        p.requires_grad_(False)
    assert not _wants_grad(dec, x)
    assert _wants_grad(dec, x.clone().requires_grad_(True))


def test_header_is_plain_c99_and_links_from_c(tmp_path):
    """include/o4d.h is the drop-in boundary: a C99 translation unit includes it (-pedantic, no C++ / torch
    types), takes the address of every declared entry point, links against libo4d.so and runs (no GPU work)."""
    import subprocess
    names = header_symbols()
    src = tmp_path / 'abi.c'
    src.write_text('#include "o4d.h"\n#include <stdio.h>\ntypedef void (*fn)(void);\n'
                   'static fn table[] = {%s};\n' % ', '.join('(fn)%s' % n for n in names) +
                   'int main(void) {\n  unsigned i, ok = 0;\n'
                   '  for (i = 0; i < sizeof(table) / sizeof(table[0]); ++i) ok += table[i] != 0;\n'
                   '  printf("%u %d %d\\n", ok, o4d_abi_version(), (int)sizeof(o4d_decoder_config));\n'
                   '  return o4d_abi_version() == 2 ? 0 : 1;\n}\n')
    exe = tmp_path / 'abi'
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(['gcc', '-std=c99', '-Wall', '-Wextra', '-pedantic', '-Werror', '-I', os.path.join(REPO, 'include'),
                    str(src), '-o', str(exe), '-L', libdir, '-lo4d', '-Wl,-rpath,' + libdir], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    assert int(out[0]) == len(names) and int(out[1]) == 2
    assert int(out[2]) == ctypes.sizeof(_lib.DecoderConfig)
