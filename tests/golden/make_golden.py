"""Generates the golden vectors under tests/golden/ by running the UNMODIFIED reference
(imported from /root/reference through oracle/ref_loader.py) on CPU in the build container.

    python tests/golden/make_golden.py            # all committed fixtures
    python tests/golden/make_golden.py --ckpt     # + real-checkpoint fixtures (git-ignored)

The reference has no tests or golden vectors of its own, so these files are
the pin: `tiny_*` carry full weights (small widths), `c*_seeded` carry only the seed
(weights are re-created by seeding torch and constructing the module -- o4d's modules
create parameters in the reference's order, checked by tests/test_module_api.py) plus the
reference's outputs.  torch_cluster is replaced by oracle/cluster_ops.py when the reference
runs (the extension is not vendored / installable: parity unpinned at that boundary).
"""
import argparse
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'occlusions-4d_b200'))

from oracle import ref_loader  # noqa: E402
from tests import configs  # noqa: E402


def run_reference(ref, pcl_args, imp_args, seed, pcl, query, state=None):
    with ref_loader.quiet():
        torch.manual_seed(seed)
        enc = ref['model'].PointCompletionNetV3(**pcl_args).eval()
        dec = ref['implicit'].LocalPclResnetFC(**imp_args).eval()
        if state is not None:
            enc.load_state_dict(state[0])
            dec.load_state_dict(state[1])
        with torch.no_grad():
            abstract, glob, coords = enc(pcl[None], True)
            out, pen = dec(query, abstract[0], glob[0], None)
    return enc, dec, abstract[0], glob[0], coords, out, pen


def save(name, **arrays):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **{k: (v.numpy() if isinstance(v, torch.Tensor) else v)
                                 for k, v in arrays.items()})
    print('wrote %s (%.1f KB)' % (name, os.path.getsize(path) / 1024))


def weight_checksum(module):
    return np.array([float(sum(p.double().sum() for p in module.parameters())),
                     float(sum(p.double().abs().sum() for p in module.parameters()))])


def make_model_fixture(ref, name, cfg, store_weights, query_subset=None):
    pcl = configs.synthetic_cloud(cfg)
    query = configs.synthetic_queries(cfg)
    if query_subset is not None and query.shape[0] > query_subset:
        sel = torch.linspace(0, query.shape[0] - 1, query_subset).long()
        query = query[sel]
    enc, dec, abstract, glob, coords, out, pen = run_reference(
        ref, cfg['pcl_args'], cfg['implicit_args'], cfg['seed'], pcl, query)
    arrays = dict(pcl=pcl, query=query, abstract=abstract, glob=glob, out=out,
                  penult=pen[:, :16].contiguous(),
                  enc_checksum=weight_checksum(enc), dec_checksum=weight_checksum(dec))
    for i, c in enumerate(coords):
        if i >= 2 and i % 2 == 1:  # coords after every down transition (indices 3,5,7)
            arrays['level_pos_%d' % ((i - 1) // 2)] = c[0]
    if store_weights:
        for k, v in enc.state_dict().items():
            arrays['enc.' + k] = v
        for k, v in dec.state_dict().items():
            arrays['dec.' + k] = v
    save(name, **arrays)


def make_knn_fixture(ref):
    g = torch.Generator().manual_seed(1830)
    ptl, geo = ref['point_transformer_layer'], ref['geometry']
    arrays = {}
    for tag, (nq, m, k) in {'a': (1000, 531, 14), 'b': (777, 2000, 16), 'c': (64, 20, 8)}.items():
        q = torch.rand(nq, 3, generator=g) * 10 - 5
        r = torch.rand(m, 3, generator=g) * 10 - 5
        idx = ptl.kNN_torch(q[None], r[None], k)[0]
        kl = min(8, m)
        idx2, dist2 = geo.my_knn_torch(q, r, kl, return_inds=True, return_knn=False, return_dists=True)
        arrays.update({tag + '_q': q, tag + '_r': r, tag + '_k': np.array(k), tag + '_idx_sq': idx,
                       tag + '_idx_eu': idx2, tag + '_dist_eu': dist2})
    # zero-padded duplicates (geometry.py:320-322 pads clouds with zeros): exact ties.
    r = torch.rand(300, 3, generator=g) * 10 - 5
    r[250:] = 0.0
    q = torch.rand(200, 3, generator=g) * 2 - 1
    arrays.update({'dup_q': q, 'dup_r': r, 'dup_k': np.array(14),
                   'dup_idx_sq': ptl.kNN_torch(q[None], r[None], 14)[0],
                   'dup_d2': ptl.square_distance(q[None], r[None])[0]})
    save('knn_golden.npz', **arrays)


def make_ckpt_fixture(ref, which):
    ck = ref_loader.load_checkpoint(which)
    cfg = configs.checkpoint_config(which, ck['pcl_args'], ck['implicit_args'])
    pcl = configs.synthetic_cloud(cfg)
    query = configs.synthetic_queries(cfg)
    sel = torch.linspace(0, query.shape[0] - 1, 8192).long()
    query = query[sel]
    enc, dec, abstract, glob, coords, out, pen = run_reference(
        ref, ck['pcl_args'], ck['implicit_args'], 0, pcl, query, state=(ck['pcl_net'], ck['implicit_net']))
    # fp64 arbitration: the oracle in double on the reference's own abstract cloud ("fp64 via
    # .double() copies to arbitrate").  fp32 implementations are judged by their distance to this.
    from oracle import o4d_oracle as orc
    sd64 = orc.cast_state(ck['implicit_net'], torch.float64)
    out64, pen64 = orc.decoder_forward(sd64, ck['implicit_args'], query.double(), abstract.double(), glob.double(),
                                       knn_fp32=True)
    os.makedirs(os.path.join(HERE, '_ckpt'), exist_ok=True)
    torch.save({'pcl_args': ck['pcl_args'], 'implicit_args': ck['implicit_args'],
                'pcl_net': ck['pcl_net'], 'implicit_net': ck['implicit_net'],
                'pcl': pcl, 'query': query, 'abstract': abstract, 'glob': glob, 'out': out,
                'penult': pen[:, :16].contiguous(), 'out64': out64, 'penult64': pen64[:, :16].contiguous(),
                'max_abs_logit': out.abs().max()},
               os.path.join(HERE, '_ckpt', which + '_nets.pt'))
    from tests.test_oracle import boundary_tie_free
    ok = boundary_tie_free(query[:, :3], abstract[:, :3],
                           [ck['implicit_args']['num_local_features'], ck['implicit_args']['cross_attn_neighbors']])
    print('wrote _ckpt/%s_nets.pt  (largest |logit| %.1f, reference fp32 vs fp64 oracle on %d tie-free rows %.2e)'
          % (which, float(out.abs().max()), int(ok.sum()),
             float((out.double() - out64)[ok].abs().max() / out64[ok].abs().max())))


def make_resnetfc_fixture(ref):
    """ResnetFC.do_forward (implicit.py:152-208): the global-only mode LocalPclResnetFC falls back to when
    num_local_features == 0 (implicit.py:365-367), with (B, D) and (B, N, D) features."""
    arrays = {}
    with ref_loader.quiet():
        for tag, kw in {'a': dict(d_in=4, d_hidden=96, d_out=5, d_latent=48, n_blocks=3, pos_encoding_freqs=4),
                        'b': dict(d_in=4, d_hidden=416, d_out=9, d_latent=128, n_blocks=2, pos_encoding_freqs=8)}.items():
            torch.manual_seed(77)
            net = ref['implicit'].ResnetFC(**kw).eval()
            g = torch.Generator().manual_seed(5)
            pts = torch.rand(2, 1500, 4, generator=g) * 8 - 4
            f_glob = torch.randn(2, kw['d_latent'], generator=g)
            f_pt = torch.randn(2, 1500, kw['d_latent'], generator=g)
            with torch.no_grad():
                o1, p1 = net(pts, f_glob)
                o2, p2 = net(pts, f_pt)
                o3, p3 = net(pts[0], f_pt[0])          # 2-D inputs are auto-batched (implicit.py:164-169)
                # LocalPclResnetFC with num_local_features=0 routes here (implicit.py:365-367)
                torch.manual_seed(77)
                loc = ref['implicit'].LocalPclResnetFC(num_local_features=0, local_mode='attention',
                                                       cross_attn_layers=0, **kw).eval()
                o4, p4 = loc(pts, None, f_glob, None)
            assert torch.equal(o3, o2[0])
            if tag == 'a':      # small: full weights travel; 'b' is re-created from the seed (checksum below)
                for k, v in net.state_dict().items():
                    arrays['%s.sd.%s' % (tag, k)] = v
            arrays[tag + '.checksum'] = weight_checksum(net)
            # inputs are re-drawn in the test from the same CPU generator (seed 5): only outputs are stored
            arrays.update({tag + '.in_checksum': np.array([float(pts.double().sum()), float(f_pt.double().sum())]),
                           tag + '.out_glob': o1,
                           tag + '.pen_glob': p1[..., :16].contiguous(), tag + '.out_pt': o2,
                           tag + '.pen_pt': p2[..., :16].contiguous(), tag + '.out_local0': o4})
            for k, v in kw.items():
                arrays['%s.kw.%s' % (tag, k)] = np.array(v)
    save('resnetfc_golden.npz', **arrays)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--ckpt', action='store_true')
    ap.add_argument('--only', default='')
    args = ap.parse_args()
    ref = ref_loader.load()
    todo = {
        'knn': lambda: make_knn_fixture(ref),
        'tiny_greater': lambda: make_model_fixture(ref, 'tiny_greater.npz', configs.TINY_GREATER, True),
        'tiny_carla': lambda: make_model_fixture(ref, 'tiny_carla.npz', configs.TINY_CARLA, True),
        'c1': lambda: make_model_fixture(ref, 'c1_greater_seeded.npz', configs.C1_GREATER, False),
        'c2': lambda: make_model_fixture(ref, 'c2_greater_seeded.npz', configs.C2_GREATER, False, 4096),
        'c3': lambda: make_model_fixture(ref, 'c3_carla_seeded.npz', configs.C3_CARLA, False, 4096),
        'resnetfc': lambda: make_resnetfc_fixture(ref),
    }
    for k, fn in todo.items():
        if (not args.only and not args.ckpt) or k in args.only.split(','):
            fn()
    if args.ckpt and args.only in ('', 'ckpt'):
        for which in ('greater', 'carla'):
            make_ckpt_fixture(ref, which)


if __name__ == '__main__':
    main()
