"""CPU tests of the oracle: (a) against the UNMODIFIED reference imported from
/root/reference (build container only, marker `reference`), (b) against the committed
golden vectors the reference produced (runs anywhere)."""
import os

import numpy as np
import pytest
import torch

from oracle import cluster_ops, o4d_oracle as orc, ref_loader
from tests import configs

GOLD = os.path.join(os.path.dirname(__file__), 'golden')
TOL = 2e-5  # fp32 reassociation only: the oracle restates the same fp32 math


def load(name):
    z = np.load(os.path.join(GOLD, name))
    return {k: torch.from_numpy(z[k]) if z[k].ndim else z[k] for k in z.files}


def split_state(g):
    enc = {k[4:]: v for k, v in g.items() if k.startswith('enc.')}
    dec = {k[4:]: v for k, v in g.items() if k.startswith('dec.')}
    return enc, dec


def relerr(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def boundary_tie_free(query_xyz, ref_xyz, ks):
    """Rows whose k-th and (k+1)-th neighbour distances differ for every k in ks: there the
    neighbour SET is unambiguous even when the cloud holds duplicate positions."""
    kmax = max(ks) + 1
    _, d = orc.knn_indices(query_xyz, ref_xyz, min(kmax, ref_xyz.shape[0]))
    ok = torch.ones(query_xyz.shape[0], dtype=torch.bool)
    for k in ks:
        if k < d.shape[1]:
            ok &= d[:, k - 1] != d[:, k]
    return ok


# ------------------------------------------------------------------ golden vectors (anywhere)

def test_knn_golden_exact():
    g = load('knn_golden.npz')
    for tag in 'abc':
        q, r, k = g[tag + '_q'], g[tag + '_r'], int(g[tag + '_k'])
        idx, _ = orc.knn_indices(q, r, k)
        assert torch.equal(idx, g[tag + '_idx_sq']), 'kNN_torch golden mismatch (%s)' % tag
        kl = g[tag + '_idx_eu'].shape[1]
        idx2, d2 = orc.knn_indices(q, r, kl, sqrt=True)
        assert torch.equal(idx2, g[tag + '_idx_eu']), 'my_knn_torch golden mismatch (%s)' % tag
        # torch.linalg.norm rounds differently from sqrt((dx2+dy2)+dz2): <= 2 ulp
        assert float((d2 - g[tag + '_dist_eu']).abs().max()) <= 1e-6


def test_knn_golden_duplicates_are_true_ties():
    g = load('knn_golden.npz')
    q, r, k = g['dup_q'], g['dup_r'], int(g['dup_k'])
    idx, d = orc.knn_indices(q, r, k)
    ref_idx = g['dup_idx_sq']
    d2 = g['dup_d2']
    # same distances row by row; any index difference sits on an exact tie
    assert torch.equal(torch.gather(d2, 1, idx), torch.gather(d2, 1, ref_idx))
    assert torch.equal(d, torch.gather(d2, 1, idx))
    # canonical rule: ascending index among equal distances
    same = d[:, 1:] == d[:, :-1]
    assert bool((idx[:, 1:][same] > idx[:, :-1][same]).all())


@pytest.mark.parametrize('name', ['tiny_greater', 'tiny_carla'])
def test_oracle_matches_golden_tiny(name):
    g = load(name + '.npz')
    cfg = configs.TINY_GREATER if name == 'tiny_greater' else configs.TINY_CARLA
    sd_e, sd_d = split_state(g)
    abstract, glob, aux = orc.encoder_forward(sd_e, cfg['pcl_args'], g['pcl'], return_aux=True)
    for l in range(1, cfg['pcl_args']['down_blocks'] + 1):
        assert torch.equal(aux['pos'][l], g['level_pos_%d' % l]), 'FPS level %d differs' % l
    assert relerr(abstract, g['abstract']) < TOL
    assert relerr(glob, g['glob']) < TOL
    out, pen = orc.decoder_forward(sd_d, cfg['implicit_args'], g['query'], g['abstract'], g['glob'])
    ok = boundary_tie_free(g['query'][:, :3], g['abstract'][:, :3],
                           [cfg['implicit_args']['num_local_features'],
                            cfg['implicit_args']['cross_attn_neighbors']])
    assert ok.float().mean() > 0.5
    assert relerr(out[ok], g['out'][ok]) < TOL
    assert relerr(pen[ok][:, :16], g['penult'][ok]) < TOL
    if cfg['pcl_args']['abstract_levels'] == 1:
        assert bool(ok.all())


def _seeded_state(cfg):
    """Weights re-created from the seed through o4d's module constructors (CPU tensors)."""
    enc, dec = configs.build_modules(cfg)
    return enc, dec


def test_seeded_init_reproduces_reference_weights():
    for name, cfg in (('c1_greater_seeded', configs.C1_GREATER), ('c2_greater_seeded', configs.C2_GREATER),
                      ('c3_carla_seeded', configs.C3_CARLA)):
        g = load(name + '.npz')
        enc, dec = _seeded_state(cfg)
        assert np.allclose(configs.weight_checksum(enc), g['enc_checksum'], rtol=0, atol=1e-9), name
        assert np.allclose(configs.weight_checksum(dec), g['dec_checksum'], rtol=0, atol=1e-9), name


def test_oracle_matches_golden_c1():
    g = load('c1_greater_seeded.npz')
    cfg = configs.C1_GREATER
    enc, dec = _seeded_state(cfg)
    sd_e, sd_d = orc.cast_state(enc.state_dict(), torch.float32), orc.cast_state(dec.state_dict(), torch.float32)
    assert torch.equal(configs.synthetic_cloud(cfg), g['pcl'])
    assert torch.equal(configs.synthetic_queries(cfg), g['query'])
    abstract, glob = orc.encoder_forward(sd_e, cfg['pcl_args'], g['pcl'])
    assert torch.equal(abstract[:, :3], g['abstract'][:, :3])
    assert relerr(abstract, g['abstract']) < TOL
    assert relerr(glob, g['glob']) < TOL
    out, pen = orc.decoder_forward(sd_d, cfg['implicit_args'], g['query'], g['abstract'], g['glob'])
    assert relerr(out, g['out']) < TOL
    assert relerr(pen[:, :16], g['penult']) < TOL


def test_oracle_decoder_matches_golden_c2_subset():
    g = load('c2_greater_seeded.npz')
    cfg = configs.C2_GREATER
    _, dec = _seeded_state(cfg)
    sd_d = orc.cast_state(dec.state_dict(), torch.float32)
    q = g['query'][:512]
    out, pen = orc.decoder_forward(sd_d, cfg['implicit_args'], q, g['abstract'], g['glob'])
    assert relerr(out, g['out'][:512]) < TOL
    assert relerr(pen[:, :16], g['penult'][:512]) < TOL


def test_fps_properties():
    torch.manual_seed(0)
    p = torch.rand(500, 3)
    sel = cluster_ops.fps_segment(p, 167, 0)
    assert sel[0] == 0 and len(set(sel.tolist())) == 167
    # every pick is the farthest point from the picks before it
    for i in (1, 2, 50, 166):
        d = ((p[:, None, :] - p[sel[:i]][None]) ** 2).sum(-1).min(dim=1)[0]
        assert abs(float(d[sel[i]]) - float(d.max())) < 1e-6
    # zero padding: more samples than distinct points -> index 0 repeats (first maximum of zeros)
    p2 = torch.zeros(10, 3)
    p2[:3] = torch.rand(3, 3)
    sel2 = cluster_ops.fps_segment(p2, 6, 0)
    assert sorted(set(sel2.tolist())) == sorted(set(sel2[:4].tolist()))


def test_use_pt_inds_rule():
    assert orc.use_pt_inds(6, 2) == {2: 0, 4: 1}
    assert orc.use_pt_inds(1, 1) == {0: 0}
    assert orc.use_pt_inds(1, 2) == {0: 1}  # collision: later layer wins (implicit.py:269)


# ------------------------------------------------------------------ live reference (container)

@pytest.mark.reference
@pytest.mark.parametrize('which', ['tiny_greater', 'tiny_carla', 'c1'])
def test_oracle_matches_live_reference(which):
    cfg = {'tiny_greater': configs.TINY_GREATER, 'tiny_carla': configs.TINY_CARLA, 'c1': configs.C1_GREATER}[which]
    ref = ref_loader.load()
    with ref_loader.quiet():
        torch.manual_seed(cfg['seed'] + 7)
        enc = ref['model'].PointCompletionNetV3(**cfg['pcl_args']).eval()
        dec = ref['implicit'].LocalPclResnetFC(**cfg['implicit_args']).eval()
        pcl = configs.synthetic_cloud(cfg)
        query = configs.synthetic_queries(cfg, num=min(cfg['num_query'], 1500), mode='random')
        with torch.no_grad():
            a, gl, _ = enc(pcl[None], False)
            o, p = dec(query, a[0], gl[0], None)
    sd_e = orc.cast_state(enc.state_dict(), torch.float32)
    sd_d = orc.cast_state(dec.state_dict(), torch.float32)
    a2, g2 = orc.encoder_forward(sd_e, cfg['pcl_args'], pcl)
    assert torch.equal(a2[:, :3], a[0][:, :3])
    assert relerr(a2, a[0]) < TOL and relerr(g2, gl[0]) < TOL
    o2, p2 = orc.decoder_forward(sd_d, cfg['implicit_args'], query, a[0], gl[0])
    ok = boundary_tie_free(query[:, :3], a[0][:, :3], [cfg['implicit_args']['num_local_features'],
                                                      cfg['implicit_args']['cross_attn_neighbors']])
    assert relerr(o2[ok], o[ok]) < TOL and relerr(p2[ok], p[ok]) < TOL


@pytest.mark.reference
def test_knn_matches_live_reference_exactly():
    ref = ref_loader.load()
    g = torch.Generator().manual_seed(5)
    pos = torch.rand(1, 3000, 3, generator=g) * 10 - 5
    ri = ref['point_transformer_layer'].kNN_torch(pos, pos, 16)[0]
    oi, _ = orc.knn_indices(pos[0], pos[0], 16)
    assert torch.equal(ri, oi)


@pytest.mark.reference
def test_query_grid_matches_live_reference():
    ref = ref_loader.load()
    for cfg, n, mode in ((configs.C2_GREATER, 4096, 'grid'), (configs.C3_CARLA, 50000, 'grid'),
                         (configs.C2_GREATER, 777, 'random')):
        ours = configs.synthetic_queries(cfg, n, mode).numpy()
        np.random.seed(cfg['seed'])
        theirs = ref['geometry'].sample_implicit_points_blind_numpy(
            n, cfg['min_z'], cfg['cr_cube_bounds'], 3, cfg['kind'], cfg['cube_mode'], mode)
        assert np.array_equal(ours, theirs)


def test_resnetfc_oracle_matches_reference_golden():
    """the global-only mode (implicit.py:152-208) against outputs of the unmodified reference."""
    g = load('resnetfc_golden.npz')
    kw = {k.split('.kw.')[1]: int(v) for k, v in g.items() if k.startswith('a.kw.')}
    sd = {k.split('.sd.')[1]: v for k, v in g.items() if k.startswith('a.sd.')}
    gen = torch.Generator().manual_seed(5)
    pts = torch.rand(2, 1500, 4, generator=gen) * 8 - 4
    f_glob = torch.randn(2, kw['d_latent'], generator=gen)
    f_pt = torch.randn(2, 1500, kw['d_latent'], generator=gen)
    o1, p1 = orc.resnetfc_forward(sd, kw, pts, f_glob)
    o2, p2 = orc.resnetfc_forward(sd, kw, pts, f_pt)
    assert relerr(o1, g['a.out_glob']) < 1e-5 and relerr(p1[..., :16], g['a.pen_glob']) < 1e-5
    assert relerr(o2, g['a.out_pt']) < 1e-5 and relerr(p2[..., :16], g['a.pen_pt']) < 1e-5
    assert relerr(o1, g['a.out_local0']) < 1e-5
