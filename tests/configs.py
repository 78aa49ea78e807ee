"""Shared test/bench configurations and seeded synthetic inputs.

Synthetic inputs: cloud (N, 8) = (x, y, z, R, G, B, t, mark_track) with xyz uniform in the
dataset cuboid (GREATER: data_greater.py:414-417; CARLA cube_mode 4: geometry.py:216-219),
RGB uniform, t = randint(video_len), mark_track = 0 (layout pipeline.py:71).  Queries come
from o4d.geometry.sample_implicit_points_blind_numpy (mirror of geometry.py:1199-1283).
"""
import copy

import numpy as np
import torch


def _pcl_args(**kw):
    base = dict(mixed_precision=False, n_input=14336, n_output=14336, d_in=8, d_out=1, d_feat=36,
                down_blocks=3, up_blocks=3, transition_factor=3, pt_num_neighbors=14, pt_norm_type='none',
                down_neighbors=12, abstract_levels=1, skip_connections=False, enable_decoder=False,
                output_featurized=True, output_global_emb=True, global_dim=128, fps_random_start=False)
    base.update(kw)
    return base


def _imp_args(**kw):
    base = dict(mixed_precision=False, d_in=4, d_hidden=416, d_out=9, d_latent=416, n_blocks=6,
                pos_encoding_freqs=8, activation='relu', num_local_features=8, local_mode='attention',
                d_latent_local=288, cross_attn_neighbors=14, cross_attn_layers=2, cr_attn_type='cc')
    base.update(kw)
    return base


GREATER_CUBE = dict(kind='greater', pt_bounds=((-5, 5), (-5, 5), (-1, 5)), cr_cube_bounds=5.0, min_z=-1.0,
                    cube_mode=4)
CARLA_CUBE = dict(kind='carla', pt_bounds=((-14, 50), (-20, 20), (-1, 10)), cr_cube_bounds=16.0, min_z=-1.0,
                  cube_mode=4)

# tiny widths: fixtures carry the full weights.
TINY_GREATER = dict(
    name='tiny_greater', seed=1830, n_points=512, video_len=4, num_query=300, query_mode='random',
    pcl_args=_pcl_args(n_input=512, n_output=512, d_feat=8, pt_num_neighbors=6, down_neighbors=5, global_dim=16),
    implicit_args=_imp_args(d_hidden=80, d_latent=80, d_latent_local=64, d_out=5, n_blocks=3,
                            pos_encoding_freqs=4, num_local_features=4, cross_attn_neighbors=6),
    **GREATER_CUBE)
TINY_CARLA = dict(
    name='tiny_carla', seed=1831, n_points=600, video_len=4, num_query=300, query_mode='random',
    pcl_args=_pcl_args(n_input=600, n_output=600, d_feat=8, pt_num_neighbors=8, down_neighbors=6, global_dim=16,
                       pt_norm_type='layer', abstract_levels=2),
    implicit_args=_imp_args(d_hidden=80, d_latent=80, d_latent_local=64, d_out=7, n_blocks=2,
                            pos_encoding_freqs=4, num_local_features=4, cross_attn_neighbors=6,
                            cross_attn_layers=1, cr_attn_type='c'),
    **CARLA_CUBE)
# BASELINE.json configs[0]: pipeline.py smoke shapes (1 MLP block, 1 cross layer after block 0).
C1_GREATER = dict(
    name='c1_greater', seed=1830, n_points=2048, video_len=4, num_query=4096, query_mode='grid',
    pcl_args=_pcl_args(n_input=2048, n_output=2048),
    implicit_args=_imp_args(d_out=5, n_blocks=1, cross_attn_layers=1, cr_attn_type='c'),
    **GREATER_CUBE)
# BASELINE.json configs[1]: the configuration the metric is quoted on.
C2_GREATER = dict(
    name='c2_greater', seed=1830, n_points=14336, video_len=12, num_query=524288, query_mode='grid',
    pcl_args=_pcl_args(), implicit_args=_imp_args(), **GREATER_CUBE)
# BASELINE.json configs[2]: CARLA-4D, two abstract levels, LayerNorm, K=16, segmentation head.
C3_CARLA = dict(
    name='c3_carla', seed=1830, n_points=14336, video_len=12, num_query=524288, query_mode='grid',
    pcl_args=_pcl_args(pt_num_neighbors=16, pt_norm_type='layer', abstract_levels=2),
    implicit_args=_imp_args(d_out=18), **CARLA_CUBE)


def checkpoint_config(which, pcl_args, implicit_args):
    cube = GREATER_CUBE if which == 'greater' else CARLA_CUBE
    return dict(name=which + '_ckpt', seed=1830, n_points=pcl_args['n_input'], video_len=12,
                num_query=524288, query_mode='grid', pcl_args=copy.deepcopy(pcl_args),
                implicit_args=copy.deepcopy(implicit_args), **cube)


def synthetic_cloud(cfg, duplicates=0):
    """(N, 8) fp32 seeded cloud; `duplicates` rows at the end are zero-padded (geometry.py:320-322)."""
    g = torch.Generator().manual_seed(cfg['seed'])
    n = cfg['n_points']
    pcl = torch.rand(n, 8, generator=g)
    for c, (lo, hi) in enumerate(cfg['pt_bounds']):
        pcl[:, c] = pcl[:, c] * (hi - lo) + lo
    pcl[:, 6] = torch.randint(0, cfg['video_len'], (n,), generator=g).float()
    pcl[:, 7] = 0.0
    if duplicates:
        pcl[n - duplicates:] = 0.0
    return pcl


def synthetic_queries(cfg, num=None, mode=None, time_idx=3):
    from o4d import geometry
    state = np.random.get_state()
    np.random.seed(cfg['seed'])
    q = geometry.sample_implicit_points_blind_numpy(
        num or cfg['num_query'], cfg['min_z'], cfg['cr_cube_bounds'], time_idx, cfg['kind'], cfg['cube_mode'],
        mode or cfg['query_mode'])
    np.random.set_state(state)
    return torch.from_numpy(q)


def build_modules(cfg, device='cpu'):
    """o4d modules with the seeded default init of the config (same RNG stream as the reference)."""
    import o4d
    torch.manual_seed(cfg['seed'])
    enc = o4d.PointCompletionNetV3(**cfg['pcl_args']).eval()
    dec = o4d.LocalPclResnetFC(**cfg['implicit_args']).eval()
    return enc.to(device), dec.to(device)


def weight_checksum(module):
    return np.array([float(sum(p.detach().double().sum() for p in module.parameters())),
                     float(sum(p.detach().double().abs().sum() for p in module.parameters()))])
