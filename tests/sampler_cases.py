"""Seeded synthetic inputs shared by tests/golden/make_golden_sampler.py (reference run, build container)
and the CPU / GPU tests of the sampler, subsampling and loss-head pieces."""
import numpy as np
import torch

SAMPLER_OUTPUTS = ('solid_input', 'air_input', 'solid_target', 'air_target', 'solid_sbs', 'air_sbs')


def seed_all(seed):
    torch.manual_seed(seed)
    np.random.seed(seed)


def _gen(seed):
    g = torch.Generator()
    g.manual_seed(seed)
    return g


def _uniform(g, n, lo, hi):
    lo = torch.tensor(lo, dtype=torch.float32)
    hi = torch.tensor(hi, dtype=torch.float32)
    return torch.rand(n, len(lo), generator=g) * (hi - lo) + lo


# ------------------------------------------------------------------ filter_air_solid_gap / select_safely
# name -> (candidates, targets, columns, radius, num_select (0 = mask form only))
FILTER_CASES = {
    'air': (3000, 2500, 3, 0.25, 2000),        # more survivors than wanted: a prefix
    'short': (600, 2000, 3, 0.45, 1500),       # fewer survivors than wanted: wrap-around
    'wide': (1500, 1800, 9, 0.5, 0),           # whole rows of a target cloud (the 'moving' bias)
    'tiles': (5000, 4500, 4, 0.3, 4096),       # several scan tiles and several kNN tiles
    'one': (1, 1, 3, 0.1, 3),
}


def filter_inputs(name):
    n, m, d, radius, num_select = FILTER_CASES[name]
    g = _gen(100 + len(name) * 7 + n)
    cand = torch.cat([_uniform(g, n, [-5, -5, -1], [5, 5, 5]), torch.rand(n, d - 3, generator=g)], dim=1)
    target = _uniform(g, m, [-5, -5, -1], [5, 5, 5])
    if name == 'one':
        target = cand[:, :3] + 1.0
    return cand.contiguous(), target.contiguous()


def bounds_input():
    g = _gen(77)
    pcl = torch.cat([_uniform(g, 4000, [-10, -25, -3], [50, 25, 10]), torch.rand(4000, 8, generator=g)], dim=1)
    # rows exactly on the faces of the CARLA output cuboid (closed intervals keep them)
    pcl[0, :3] = torch.tensor([0.0, 0.0, 0.0])
    pcl[1, :3] = torch.tensor([40.0, 16.0, 6.4])
    pcl[2, :3] = torch.tensor([40.0, -16.0, -1.0])
    pcl[3, :3] = torch.tensor([40.000004, 0.0, 0.0])
    return pcl.contiguous()


# ------------------------------------------------------------------ whole sampler
SAMPLER_CASES = {
    'greater_none': dict(seed=11, time_idx=1, frames=3, batch=2, points=1600, moving=(0, 0),
                         kwargs=dict(min_z=-1.0, cube_bounds=5.0, point_occupancy_radius=0.2, num_solid=256,
                                     num_air=512, data_kind='greater', point_sample_bias='none')),
    'greater_low_moving': dict(seed=12, time_idx=2, frames=3, batch=2, points=1600, moving=(400, 90),
                               kwargs=dict(min_z=-1.0, cube_bounds=5.0, point_occupancy_radius=0.2,
                                           num_solid=320, num_air=640, data_kind='greater',
                                           point_sample_bias='low_moving')),
    'carla_all': dict(seed=13, time_idx=0, frames=2, batch=1, points=6000, moving=(500,),
                      kwargs=dict(min_z=-1.0, cube_bounds=16.0, point_occupancy_radius=0.3, num_solid=512,
                                  num_air=768, predict_segmentation=True, semantic_classes=13,
                                  data_kind='carla', point_sample_bias='low_moving_vehped_ivalo_sembal',
                                  cube_mode=4)),
}


def sampler_inputs(name):
    """-> (pcl_target list-T of (B, M, E), pcl_target_size list-T of (B,), valo_ids (B, R), num_valo_ids (B,))."""
    case = SAMPLER_CASES[name]
    carla = case['kwargs']['data_kind'] == 'carla'
    g = _gen(1000 + case['seed'])
    T, B, M = case['frames'], case['batch'], case['points']
    lo, hi = ([-4, -20, -2], [46, 20, 8]) if carla else ([-5, -5, -1], [5, 5, 5])
    E = 11 if carla else 9
    static = []
    for b in range(B):
        n_static = M - case['moving'][b] - 37 * b          # ragged valid sizes
        rows = torch.zeros(n_static, E)
        rows[:, :3] = _uniform(g, n_static, lo, hi)
        static.append(rows)
    frames, sizes = [], []
    for t in range(T):
        frame = torch.zeros(B, M, E)
        size = torch.zeros(B, dtype=torch.int64)
        for b in range(B):
            nm = case['moving'][b]
            blob = torch.zeros(nm, E)
            if nm:
                centre = torch.tensor([12.0 + 6.0 * t, -3.0 + 2.0 * t, 0.8]) if carla \
                    else torch.tensor([-3.0 + 2.5 * t, 1.0 - 1.5 * t, 1.0])
                blob[:, :3] = centre + (torch.rand(nm, 3, generator=g) - 0.5) * torch.tensor([2.0, 2.0, 1.2])
            rows = torch.cat([static[b], blob], dim=0)
            n = rows.shape[0]
            feat = torch.rand(n, 4, generator=g)
            feat[:, 3] = (feat[:, 3] > 0.7).float()           # mark_track
            rows[:, -4:] = feat
            if carla:
                rows[:, 3] = torch.rand(n, generator=g) * 2 - 1                       # cosine_angle
                tags = torch.randint(0, 23, (n,), generator=g)
                tags[torch.rand(n, generator=g) < 0.12] = 4                           # pedestrians
                tags[torch.rand(n, generator=g) < 0.15] = 10                          # vehicles
                inst = torch.randint(1, 7, (n,), generator=g)
                view = torch.randint(0, 3, (n,), generator=g)
                view[inst == 5] = 1                                                    # instance 5 is never seen
                rows[:, 4], rows[:, 5], rows[:, 6] = inst.float(), tags.float(), view.float()
            else:
                rows[:, 3] = torch.randint(0, 8, (n,), generator=g).float()           # instance id
                rows[:, 4] = torch.randint(0, 3, (n,), generator=g).float()           # view index
            rows = rows[torch.randperm(n, generator=g)]
            frame[b, :n] = rows
            size[b] = n
        frames.append(frame)
        sizes.append(size)
    valo = torch.zeros(B, 6)
    num_valo = torch.zeros(B, dtype=torch.int64)
    if carla:
        valo[0, :3] = torch.tensor([5.0, 2.0, 3.0])
        num_valo[0] = 3
    return frames, sizes, valo, num_valo


# ------------------------------------------------------------------ loss heads
LOSS_CASES = {
    'rgb': dict(color_mode='rgb', semantic_classes=0, track=True, n=3000, g=5, seed=21),
    'rgb_seg': dict(color_mode='rgb_nosigmoid', semantic_classes=13, track=True, n=2500, g=18, seed=22),
    'hsv_seg': dict(color_mode='hsv', semantic_classes=13, track=True, n=4000, g=33, seed=23),
    'hsv_bland': dict(color_mode='hsv', semantic_classes=0, track=False, n=300, g=15, seed=24),
    'bins': dict(color_mode='bins', semantic_classes=0, track=True, n=2000, g=11, seed=25),
}


def loss_inputs(name):
    """-> (output (N, G) logits, target (N, 6) = (density, R, G, B, mark_track, segm))."""
    case = LOSS_CASES[name]
    g = _gen(case['seed'])
    n = case['n']
    output = torch.randn(n, case['g'], generator=g) * 2.0
    target = torch.full((n, 6), -1.0)
    solid = torch.rand(n, generator=g) < 0.45
    target[:, 0] = solid.float()
    rgb = torch.rand(n, 3, generator=g)
    if name == 'hsv_bland':
        rgb = 0.5 + (rgb - 0.5) * 0.05                          # nearly gray: fewer than 16 vivid hues
        rgb[:5] = torch.tensor([0.9, 0.1, 0.2])
    rgb[::7] = rgb[::7][:, [0, 0, 0]]                           # exact grays (max == min)
    rgb[5::11, 1] = rgb[5::11, 0]                               # two equal channels
    target[:, 1:4] = rgb
    target[:, 4] = (torch.rand(n, generator=g) < 0.3).float()
    target[:, 5] = torch.randint(0, max(case['semantic_classes'], 1), (n,), generator=g).float()
    no_color = torch.rand(n, generator=g) < 0.1                 # solid rows without colour / track labels
    target[no_color, 1:5] = -1.0
    target[~solid, 1:] = -1.0                                   # air rows: density 0, everything else -1
    return output.contiguous(), target.contiguous()
