"""GPU parity tests of the encoder / decoder modules (nn.Module API -> C ABI -> CUDA)
against the reference's golden vectors and the CPU oracle, plus size-independent
properties at BASELINE.json's full sizes.

Tolerance (north star): outputs within 1e-3 of the fp32 reference, measured as
max|ours - ref| / max|ref| per tensor (logits reach O(100), so element-wise relative error
near zero crossings is meaningless); kNN / FPS indices bit-exact."""
import os

import numpy as np
import pytest
import torch

import o4d
from o4d import ops, parallel
from oracle import o4d_oracle as orc
from tests import configs
from tests.test_oracle import GOLD, boundary_tie_free, load, relerr, split_state

pytestmark = pytest.mark.gpu
DEV = 'cuda'
TOL = 1e-3          # the north-star bar
TOL_TIGHT = 1e-4    # what the fp32 / bf16x3 paths actually deliver


def modules_from_golden(cfg, g):
    enc, dec = configs.build_modules(cfg)
    if any(k.startswith('enc.') for k in g):
        sd_e, sd_d = split_state(g)
        enc.load_state_dict(sd_e, strict=True)
        dec.load_state_dict(sd_d, strict=True)
    else:
        assert np.allclose(configs.weight_checksum(enc), g['enc_checksum'], rtol=0, atol=1e-9)
        assert np.allclose(configs.weight_checksum(dec), g['dec_checksum'], rtol=0, atol=1e-9)
    return enc.to(DEV).eval(), dec.to(DEV).eval()


CASES = [('tiny_greater.npz', configs.TINY_GREATER), ('tiny_carla.npz', configs.TINY_CARLA),
         ('c1_greater_seeded.npz', configs.C1_GREATER), ('c2_greater_seeded.npz', configs.C2_GREATER),
         ('c3_carla_seeded.npz', configs.C3_CARLA)]


@pytest.mark.parametrize('name,cfg', CASES, ids=[c[0][:-4] for c in CASES])
def test_encoder_and_decoder_match_reference_golden(name, cfg):
    g = load(name)
    enc, dec = modules_from_golden(cfg, g)
    with torch.no_grad():
        abstract, glob, coords = enc(g['pcl'].to(DEV)[None], True)
    assert abstract.shape == (1,) + tuple(g['abstract'].shape) and glob.shape == (1,) + tuple(g['glob'].shape)
    # FPS picks are bit-exact at every level (coords after each down transition)
    for l in range(1, cfg['pcl_args']['down_blocks'] + 1):
        assert torch.equal(coords[1 + 2 * l].cpu()[0], g['level_pos_%d' % l]), 'level %d coordinates' % l
    assert torch.equal(abstract.cpu()[0][:, :3], g['abstract'][:, :3])
    e_abs, e_glob = relerr(abstract.cpu()[0], g['abstract']), relerr(glob.cpu()[0], g['glob'])
    assert e_abs < TOL_TIGHT and e_glob < TOL_TIGHT, (e_abs, e_glob)
    # decoder on the REFERENCE's abstract cloud (isolates it from encoder rounding)
    with torch.no_grad():
        out, pen = dec(g['query'].to(DEV), g['abstract'].to(DEV), g['glob'].to(DEV), None)
    ok = boundary_tie_free(g['query'][:, :3], g['abstract'][:, :3],
                           [cfg['implicit_args']['num_local_features'], cfg['implicit_args']['cross_attn_neighbors']])
    if cfg['pcl_args']['abstract_levels'] == 1:
        assert bool(ok.all())
    assert ok.float().mean() > 0.5
    e_out, e_pen = relerr(out.cpu()[ok], g['out'][ok]), relerr(pen.cpu()[ok][:, :16], g['penult'][ok])
    assert e_out < TOL_TIGHT and e_pen < TOL_TIGHT, (e_out, e_pen)
    # end to end (our encoder feeding our decoder) stays inside the north-star bar
    with torch.no_grad():
        out2, _ = dec(g['query'].to(DEV), abstract[0], glob[0], None)
    assert relerr(out2.cpu()[ok], g['out'][ok]) < TOL


def test_tie_rows_match_the_oracle_canonical_rule():
    """Duplicate abstract positions (abstract_levels=2): the reference's topk leaves the choice
    open; we must agree with the oracle's (distance, index) rule on EVERY row."""
    g = load('tiny_carla.npz')
    cfg = configs.TINY_CARLA
    _, dec = modules_from_golden(cfg, g)
    _, sd_d = split_state(g)
    want, _ = orc.decoder_forward(sd_d, cfg['implicit_args'], g['query'], g['abstract'], g['glob'])
    with torch.no_grad():
        out, _ = dec(g['query'].to(DEV), g['abstract'].to(DEV), g['glob'].to(DEV), None)
    assert relerr(out.cpu(), want) < TOL_TIGHT


def test_call_forms_of_both_reference_callers():
    g = load('tiny_greater.npz')
    enc, dec = modules_from_golden(configs.TINY_GREATER, g)
    pcl = g['pcl'].to(DEV)[None]
    with torch.no_grad():
        r3 = enc(pcl, False)                 # eval/inference.py:195
        r4 = enc(pcl, False, False)          # pipeline.py:93-94
        assert len(r3) == 3 and len(r4) == 4 and r3[2] is None and r4[3] is None
        assert torch.equal(r3[0], r4[0])
        q = g['query'].to(DEV)
        o2 = dec(q, r3[0][0], r3[1][0], None)                    # inference.py:211 (unbatched)
        o3 = dec(q[None], r3[0], r3[1], None, False)             # pipeline.py:193 (batched, extra flag)
        assert len(o2) == 2 and len(o3) == 3 and o3[2] is None
        assert o2[0].shape == (q.shape[0], 5) and o3[0].shape == (1, q.shape[0], 5)
        assert torch.equal(o2[0], o3[0][0])
        # separated coordinates / features (implicit.py:286-290 else-branch)
        o4 = dec(q, r3[0][0][:, :3], r3[1][0], r3[0][0][:, 3:])
        assert torch.equal(o4[0], o2[0])
        # the caller post-processes in place (inference.py:218): result must be writable
        o2[0][..., 0] = torch.sigmoid(o2[0][..., 0])
        with pytest.raises(AssertionError):
            dec(q[None].expand(2, -1, -1), r3[0].expand(2, -1, -1), r3[1].expand(2, -1), None)  # B must be 1


def test_next_scene_updates_the_scene_buffer_in_place():
    """A second scene with the same weights and abstract-cloud size takes o4d_decoder_update_scene (the weight-only parts
    of the scene buffer are reused); results must equal a freshly prepared scene bit for bit, and a parameter change
    must force a full prepare."""
    g = load('c1_greater_seeded.npz')
    cfg = configs.C1_GREATER
    _, dec = modules_from_golden(cfg, g)
    _, fresh = modules_from_golden(cfg, g)
    q = g['query'].to(DEV)
    a1, g1 = g['abstract'].to(DEV), g['glob'].to(DEV)
    gen = torch.Generator().manual_seed(3)
    a2 = torch.cat([a1[:, :3].cpu() + 0.25 * torch.randn(a1.shape[0], 3, generator=gen),
                    a1[:, 3:].cpu() * 1.5], dim=1).to(DEV)
    g2 = (g1.cpu() + torch.randn(g1.shape, generator=gen)).to(DEV)
    with torch.no_grad():
        o1, _ = dec(q, a1, g1, None)
        scene1 = dec._o4d_scene
        o2, _ = dec(q, a2, g2, None)                       # same weights, same m -> in-place update
        assert dec._o4d_scene is scene1
        want2, _ = fresh(q, a2, g2, None)                  # fresh module: full prepare on scene 2
        assert torch.equal(o2, want2)
        assert relerr(o1.cpu(), g['out']) < TOL_TIGHT
        o1b, _ = dec(q, a1, g1, None)                      # and back
        assert torch.equal(o1b, o1)
        dec.lin_out.bias.add_(1.0)                         # parameter version bump -> full prepare, new buffer contents
        o3, _ = dec(q, a1, g1, None)
        assert dec._o4d_scene is not scene1
        assert relerr((o3 - 1.0).cpu(), o1.cpu()) < 1e-6


def test_batched_encoder_equals_per_cloud():
    g = load('tiny_greater.npz')
    enc, _ = modules_from_golden(configs.TINY_GREATER, g)
    other = configs.synthetic_cloud(dict(configs.TINY_GREATER, seed=99))
    both = torch.stack([g['pcl'], other]).to(DEV)
    with torch.no_grad():
        a, gl, _ = enc(both, False)
        a0, g0, _ = enc(both[:1], False)
        a1, g1, _ = enc(both[1:], False)
    assert torch.equal(a[0], a0[0]) and torch.equal(a[1], a1[0]) and torch.equal(gl[1], g1[0])


def test_zero_padded_cloud_runs_and_matches_oracle():
    cfg = configs.TINY_GREATER
    g = load('tiny_greater.npz')
    enc, _ = modules_from_golden(cfg, g)
    pcl = configs.synthetic_cloud(cfg, duplicates=100)      # geometry.py:320-322 style padding
    sd_e, _ = split_state(g)
    want, want_g = orc.encoder_forward(sd_e, cfg['pcl_args'], pcl)
    with torch.no_grad():
        a, gl, _ = enc(pcl.to(DEV)[None], False)
    assert torch.equal(a.cpu()[0][:, :3], want[:, :3])
    assert relerr(a.cpu()[0], want) < TOL_TIGHT and relerr(gl.cpu()[0], want_g) < TOL_TIGHT


def test_random_fps_start_is_used_in_training_configuration():
    cfg = dict(configs.TINY_GREATER)
    cfg['pcl_args'] = dict(cfg['pcl_args'], fps_random_start=True)
    enc, _ = configs.build_modules(cfg)
    enc = enc.to(DEV).eval()
    pcl = configs.synthetic_cloud(cfg).to(DEV)[None]
    with torch.no_grad():
        torch.manual_seed(1)
        a = enc(pcl, False)[0]
        torch.manual_seed(2)
        b = enc(pcl, False)[0]
        torch.manual_seed(1)
        c = enc(pcl, False)[0]
    assert torch.equal(a, c) and not torch.equal(a[..., :3], b[..., :3])


# ------------------------------------------------------------ full-size properties (config 2)

@pytest.fixture(scope='module')
def c2():
    g = load('c2_greater_seeded.npz')
    enc, dec = modules_from_golden(configs.C2_GREATER, g)
    return g, enc, dec


def test_full_size_minibatch_invariance_and_host_loop(c2):
    """524,288-query grid (534,528 points): per-query results must not depend on how the
    frame is cut into mini-batches, and the host-buffer C-ABI loop must equal the module."""
    g, enc, dec = c2
    cfg = configs.C2_GREATER
    q = configs.synthetic_queries(cfg)
    assert q.shape[0] == 534528
    abstract, glob = g['abstract'].to(DEV), g['glob'].to(DEV)
    qd = q.to(DEV)
    with torch.no_grad():
        full = torch.cat([dec(qd[s:s + 32768], abstract, glob, None)[0] for s in range(0, q.shape[0], 32768)])
        sel = torch.linspace(0, q.shape[0] - 1, 4096).long()
        # golden subset of the reference
        assert relerr(full[sel.to(DEV)].cpu(), g['out']) < 1e-4
        # different mini-batch size -> same answers
        odd = torch.cat([dec(qd[s:s + 8192], abstract, glob, None)[0] for s in range(0, 65536, 8192)])
        assert relerr(odd.cpu(), full[:65536].cpu()) < 1e-6
        # host-buffer loop through the C ABI (what bench.py's e2e leg times)
        scene = dec.o4d_scene(abstract, glob)
        host = ops.decoder_run_host(dec.o4d_config(), dec.o4d_params(), scene, q[:100000].contiguous(), 32768)
    assert torch.equal(host, full[:100000].cpu())
    assert bool(torch.isfinite(full).all())


def test_full_size_sharded_decode_single_process(c2):
    g, enc, dec = c2
    q = configs.synthetic_queries(configs.C2_GREATER)[:70000].to(DEV)
    abstract, glob = g['abstract'].to(DEV), g['glob'].to(DEV)
    with torch.no_grad():
        fn = lambda b: dec(b, abstract, glob, None)[0]
        whole = parallel.decode_sharded(fn, q, 32768)
        parts = []
        for r in range(3):   # emulate three ranks' contiguous shards
            a, b = parallel.shard_range(q.shape[0], r, 3)
            parts.append(torch.cat([fn(q[s:min(s + 32768, b)]) for s in range(a, b, 32768)]))
    assert relerr(torch.cat(parts).cpu(), whole.cpu()) < 1e-6


def test_full_size_encoder_is_deterministic(c2):
    g, enc, dec = c2
    pcl = g['pcl'].to(DEV)[None]
    with torch.no_grad():
        a1, g1, _ = enc(pcl, False)
        a2, g2, _ = enc(pcl, False)
    assert torch.equal(a1, a2) and torch.equal(g1, g2)
    assert a1.shape == (1, 531, 291)


# ------------------------------------------------------------ released checkpoints

# Distance to the fp64 arbitration run.  Measured on the B200 (tools/ckpt_error.py, profiles/r2_a_ckpt_error.json): with the
# released weights (logits to 362) every fp32-accumulating implementation sits at 2-5e-4 from fp64 -- the unmodified
# reference's own forward 2.3e-4 (GREATER) / 1.6e-4 (CARLA), our fp32 CUDA-core path 3.9e-4 / 3.4e-4, bf16x3 3.9e-4 /
# 9.7e-4 -- so the north-star bar (1e-3) is the bound, and the path must also stay within 8x of the reference's own
# fp32 rounding distance.
TOL_FP64 = 1e-3


@pytest.mark.parametrize('which', ['greater', 'carla'])
def test_released_checkpoint_parity(which):
    """The released weights (pretrained/*.pth, loaded as eval/inference.py:39-73 does) through encoder and
    decoder, against (i) the unmodified reference's fp32 outputs and (ii) the oracle run in fp64 on the same
    neighbour sets.  Fixture: tests/golden/make_golden.py --ckpt.  A missing fixture is a FAILURE: real-weight
    parity is part of the bar, not an optional extra."""
    path = os.path.join(GOLD, '_ckpt', which + '_nets.pt')
    assert os.path.isfile(path), 'checkpoint fixture missing: run tests/golden/make_golden.py --ckpt'
    ck = torch.load(path, map_location='cpu', weights_only=True)
    enc = o4d.PointCompletionNetV3(**ck['pcl_args'])
    dec = o4d.LocalPclResnetFC(**ck['implicit_args'])
    enc.load_state_dict(ck['pcl_net'], strict=True)
    dec.load_state_dict(ck['implicit_net'], strict=True)
    enc, dec = enc.to(DEV).eval(), dec.to(DEV).eval()
    with torch.no_grad():
        a, gl, _ = enc(ck['pcl'].to(DEV)[None], False)
        out, pen = dec(ck['query'].to(DEV), ck['abstract'].to(DEV), ck['glob'].to(DEV), None)
    out, pen = out.cpu(), pen.cpu()
    assert torch.equal(a.cpu()[0][:, :3], ck['abstract'][:, :3])
    assert relerr(a.cpu()[0], ck['abstract']) < TOL and relerr(gl.cpu()[0], ck['glob']) < TOL
    # (i) the reference itself, on rows whose neighbour sets are unambiguous (abstract_levels=2 duplicates positions)
    ok = boundary_tie_free(ck['query'][:, :3], ck['abstract'][:, :3],
                           [ck['implicit_args']['num_local_features'], ck['implicit_args']['cross_attn_neighbors']])
    assert ok.float().mean() > 0.5
    e_ref, e_ref_pen = relerr(out[ok], ck['out'][ok]), relerr(pen[ok][:, :16], ck['penult'][ok])
    assert e_ref < TOL and e_ref_pen < TOL, (e_ref, e_ref_pen)
    # (ii) fp64 arbitration on EVERY row (same canonical tie rule, fp32 neighbour sets)
    e64 = relerr(out.double(), ck['out64'])
    e64_pen = relerr(pen[:, :16].double(), ck['penult64'])
    ref64 = relerr(ck['out'][ok].double(), ck['out64'][ok])
    print('%s checkpoint: largest |logit| %.1f; ours vs reference %.2e, ours vs fp64 %.2e, reference vs fp64 %.2e'
          % (which, float(out.abs().max()), e_ref, e64, ref64))
    assert e64 < TOL_FP64 and e64_pen < TOL_FP64, (e64, e64_pen)
    assert e64 < 8 * max(ref64, 1e-4), (e64, ref64)
    # the large-logit regime is on record (SURVEY hard part 1: GREATER logits reach several hundred)
    assert abs(float(out.abs().max()) - float(ck['max_abs_logit'])) < 1e-2 * float(ck['max_abs_logit'])
    if which == 'greater':
        assert float(ck['max_abs_logit']) > 100.0
    assert bool(torch.isfinite(out).all())


# ------------------------------------------------------------ ResnetFC.do_forward

@pytest.mark.parametrize('tag', ['a', 'b'])
def test_resnetfc_global_mode_matches_reference_golden(tag):
    """implicit.py:152-208: global-only conditioning, (B, D) and (B, N, D) features, 2-D inputs, and the
    num_local_features == 0 route of LocalPclResnetFC (implicit.py:365-367)."""
    g = load('resnetfc_golden.npz')
    kw = {k.split('.kw.')[1]: int(v) for k, v in g.items() if k.startswith(tag + '.kw.')}
    torch.manual_seed(77)
    net = o4d.ResnetFC(**kw).eval()
    if tag == 'a':
        net.load_state_dict({k.split('.sd.')[1]: v for k, v in g.items() if k.startswith(tag + '.sd.')}, strict=True)
    assert np.allclose(configs.weight_checksum(net), g[tag + '.checksum'].numpy(), rtol=0, atol=1e-9)
    gen = torch.Generator().manual_seed(5)
    pts = torch.rand(2, 1500, 4, generator=gen) * 8 - 4
    f_glob = torch.randn(2, kw['d_latent'], generator=gen)
    f_pt = torch.randn(2, 1500, kw['d_latent'], generator=gen)
    assert np.allclose([float(pts.double().sum()), float(f_pt.double().sum())], g[tag + '.in_checksum'].numpy())
    net = net.to(DEV)
    with torch.no_grad():
        o1, p1 = net(pts.to(DEV), f_glob.to(DEV))
        o2, p2 = net(pts.to(DEV), f_pt.to(DEV))
        o3, p3 = net(pts[0].to(DEV), f_pt[0].to(DEV))
        torch.manual_seed(77)
        loc = o4d.LocalPclResnetFC(num_local_features=0, local_mode='attention', cross_attn_layers=0, **kw).eval()
        if tag == 'a':
            loc.load_state_dict(net.state_dict(), strict=True)
        o4, _ = loc.to(DEV)(pts.to(DEV), None, f_glob.to(DEV), None)
    assert o1.shape == (2, 1500, kw['d_out']) and p1.shape == (2, 1500, kw['d_hidden'])
    assert o3.shape == (1500, kw['d_out']) and torch.equal(o3, o2[0])
    for got, want in ((o1, g[tag + '.out_glob']), (p1[..., :16], g[tag + '.pen_glob']), (o2, g[tag + '.out_pt']),
                      (p2[..., :16], g[tag + '.pen_pt']), (o4, g[tag + '.out_local0'])):
        assert relerr(got.cpu(), want) < TOL_TIGHT, relerr(got.cpu(), want)
    with pytest.raises(AssertionError):
        net(pts.to(DEV), f_glob[:1].to(DEV))                 # batch mismatch (implicit.py:171)
    with pytest.raises(AssertionError):
        net(pts.to(DEV), f_pt[:, :10].to(DEV))               # per-point features of the wrong length (:176)


# ------------------------------------------------------------ fused attention kernel coverage

@pytest.mark.parametrize('d_hidden,d_local,k_cross,n_query,m_abs', [
    (416, 288, 14, 1, 300),       # single query: one 9-query tile, 8 of them padding
    (416, 288, 14, 1003, 531),    # ragged tail tile
    (416, 288, 16, 777, 100),     # 8 queries x 16 neighbours fill the 128-row tile exactly
    (416, 288, 12, 500, 2124),    # 10 queries per tile, CARLA-sized abstract cloud
    (288, 224, 14, 640, 200),     # narrower decoder: two 144-column MMA tiles, 18 hidden chunks
    (320, 256, 9, 333, 64),       # 14 queries per tile
])
def test_fused_attention_decoder_shapes_match_oracle(d_hidden, d_local, k_cross, n_query, m_abs):
    args = dict(configs.C2_GREATER['implicit_args'], d_hidden=d_hidden, d_latent=d_hidden, d_latent_local=d_local,
                cross_attn_neighbors=k_cross, n_blocks=2, cross_attn_layers=1, cr_attn_type='c', d_out=7)
    torch.manual_seed(d_hidden + k_cross)
    dec = o4d.LocalPclResnetFC(**args).eval()
    sd = orc.cast_state(dec.state_dict(), torch.float32)
    g = torch.Generator().manual_seed(n_query)
    abstract = torch.cat([torch.rand(m_abs, 3, generator=g) * 10 - 5, torch.randn(m_abs, d_local, generator=g)], dim=1)
    glob = torch.randn(d_hidden - d_local, generator=g)
    query = torch.cat([torch.rand(n_query, 3, generator=g) * 10 - 5, torch.full((n_query, 1), 3.0)], dim=1)
    want, want_pen = orc.decoder_forward(sd, args, query, abstract, glob)
    dec = dec.to(DEV)
    outs = {}
    for prec in (0, 1):
        dec.o4d_precision = prec
        with torch.no_grad():
            out, pen = dec(query.to(DEV), abstract.to(DEV), glob.to(DEV), None)
        outs[prec] = out.cpu()
        assert relerr(out.cpu(), want) < TOL_TIGHT, (prec, relerr(out.cpu(), want))
        assert relerr(pen.cpu(), want_pen) < TOL_TIGHT, prec
    # single-pass bf16 (precision 2) is the fast, lossy mode: documented ~1e-2, must stay sane
    dec.o4d_precision = 2
    with torch.no_grad():
        out2, _ = dec(query.to(DEV), abstract.to(DEV), glob.to(DEV), None)
    assert relerr(out2.cpu(), want) < 5e-2


# ------------------------------------------------------------ test-time driver

def test_inference_frame_loop_equals_the_reference_formulation():
    """o4d.inference.run_frame (device lattice, mini-batches into one buffer, fused squashing, one D2H) against
    the eval/inference.py:175-248 formulation: numpy lattice, per-mini-batch upload / decode / torch sigmoid."""
    from o4d import geometry, inference
    cfg = configs.C1_GREATER
    g = load('c1_greater_seeded.npz')
    enc, dec = modules_from_golden(cfg, g)
    pcl = g['pcl'].to(DEV)[None]
    res = inference.run_frame(enc, dec, pcl, 4096, -1.0, 5.0, 2, 'greater', 4, point_sample_mode='grid',
                              batch_size=1000, color_mode='rgb', density_threshold=0.5, to_host=True)
    pts = geometry.sample_implicit_points_blind_numpy(4096, -1.0, 5.0, 2, 'greater', 4, 'grid')
    assert res['points_query'].shape == (4332, 4) and np.array_equal(res['points_query'].cpu().numpy(), pts)
    with torch.no_grad():
        abstract, glob, _ = enc(pcl, False)
        chunks = []
        for s in range(0, pts.shape[0], 1000):
            o = dec(torch.from_numpy(pts[s:s + 1000]).to(DEV), abstract[0], glob[0], None)[0]
            o[..., 0] = torch.sigmoid(o[..., 0])
            o[..., 1:4] = torch.sigmoid(o[..., 1:4])
            chunks.append(o.cpu().numpy())
    want = np.concatenate(chunks, axis=0)
    assert res['implicit_output'].shape == want.shape
    np.testing.assert_allclose(res['implicit_output'], want, rtol=0, atol=2e-6)
    assert np.array_equal(res['points_io'][:, :4], pts)
    assert np.array_equal(res['solid_mask'], res['implicit_output'][:, 0] >= 0.5)
    # device-resident form, random queries (the reference's numpy draws, uploaded once)
    np.random.seed(5)
    q = inference.query_points(3000, -1.0, 5.0, 1, 'greater', 4, 'random', DEV)
    np.random.seed(5)
    assert np.array_equal(q.cpu().numpy(),
                          geometry.sample_implicit_points_blind_numpy(3000, -1.0, 5.0, 1, 'greater', 4, 'random'))
    dev = inference.query_frame(dec, abstract[0], glob[0], q, batch_size=4096, color_mode='rgb')
    assert dev['implicit_output'].is_cuda and dev['points_io'].shape == (3000, 4 + cfg['implicit_args']['d_out'])
    assert bool(((dev['implicit_output'][:, :4] >= 0) & (dev['implicit_output'][:, :4] <= 1)).all())
