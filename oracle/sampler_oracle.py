"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the training-time sampler's device ops, the dataset-side
subsampling and the implicit loss heads.

Plain torch on CPU; every function cites the reference file:line it restates.  The product never
imports this file.

Pinning: ``tests/test_oracle.py`` runs these functions against the unmodified reference imported from
``/root/reference`` (build container only) and against ``tests/golden/sampler_*.npz`` /
``tests/golden/loss_golden.npz`` (reference outputs written by ``tests/golden/make_golden_sampler.py``),
which travel to the GPU box.
"""
import numpy as np
import torch


# --------------------------------------------------------------------------- sampler ops

def nn1_dist(points_xyz, target_xyz, chunk=4096):
    """Euclidean distance of every point to its nearest target point, fp32 sqrt((dx2 + dy2) + dz2).
    utils/geometry.py:1183-1190 (my_knn_torch with K = 1, minimum over target slices: the global minimum)."""
    out = torch.empty(points_xyz.shape[0], dtype=torch.float32)
    t = target_xyz[:, :3].float()
    for s in range(0, points_xyz.shape[0], chunk):
        d = points_xyz[s:s + chunk, None, :3].float() - t[None]
        d = d * d
        out[s:s + chunk] = ((d[..., 0] + d[..., 1]) + d[..., 2]).min(dim=1).values
    # correctly rounded IEEE sqrt (numpy); torch's vectorised CPU sqrt is off by one ulp on ~0.6 % of inputs
    return torch.from_numpy(np.sqrt(out.numpy()))


def filter_air_solid_gap(to_filter, target_coords, target_slice_size, point_occupancy_radius):
    """utils/geometry.py:1164-1196 -> (kept rows, their distances, kept fraction)."""
    del target_slice_size  # slicing only bounds the reference's temporaries; the minimum is global
    dist = nn1_dist(to_filter, target_coords)
    good = dist > point_occupancy_radius
    return to_filter[good], dist[good], good.sum() / dist.shape[0]


def select_safely(rows, num_select):
    """utils/geometry.py:1095-1105: first num_select rows, the content repeated by doubling when short."""
    assert rows.shape[0] > 0
    while rows.shape[0] < num_select:
        rows = torch.cat([rows, rows], dim=0)
    return rows[:num_select].clone()


def filter_select(to_filter, target_coords, point_occupancy_radius, num_select):
    """filter_air_solid_gap + select_safely on rows and distances, as the sampler uses them
    (utils/geometry.py:998-1002, 1022-1025, 1044-1047, 1065-1068) -> (rows, dists, count (1,) int32)."""
    rows, dist, _ = filter_air_solid_gap(to_filter, target_coords, 0, point_occupancy_radius)
    count = torch.tensor([rows.shape[0]], dtype=torch.int32)
    if rows.shape[0] == 0:
        return (torch.zeros((num_select, to_filter.shape[1])), torch.zeros(num_select), count)
    return select_safely(rows, num_select), select_safely(dist, num_select), count


def filter_pcl_bounds(pcl, x_min=-10.0, x_max=10.0, y_min=-10.0, y_max=10.0, z_min=-10.0, z_max=10.0):
    """utils/geometry.py:175-188: rows inside the closed cuboid, in order."""
    lo = torch.tensor([x_min, y_min, z_min], dtype=torch.float32)
    hi = torch.tensor([x_max, y_max, z_max], dtype=torch.float32)
    xyz = pcl[..., :3]
    return pcl[torch.logical_and(lo <= xyz, xyz <= hi).all(dim=-1)]


# --------------------------------------------------------------------------- loss heads

COLOR_MODES = ('rgb', 'rgb_nosigmoid', 'hsv', 'bins')


def track_idx(color_mode):
    """utils/utils.py:204-224."""
    return {'rgb': 4, 'rgb_nosigmoid': 4, 'hsv': 15, 'bins': 10}[color_mode]


def rgb_to_hsv(rgb, epsilon=1e-10):
    """utils/utils.py:169-191 (hue in degrees selected by the arg-min channel)."""
    r, g, b = rgb[:, 0], rgb[:, 1], rgb[:, 2]
    mx = rgb.max(1).values
    mn, argmin = rgb.min(1)
    span = mx - mn + epsilon
    h_b_min = 60.0 * (g - r) / span + 60.0     # blue is the minimum
    h_r_min = 60.0 * (b - g) / span + 180.0    # red is the minimum
    h_g_min = 60.0 * (r - b) / span + 300.0    # green is the minimum
    h = torch.where(argmin == 0, h_r_min, torch.where(argmin == 1, h_g_min, h_b_min))
    return torch.stack((h, span / (mx + epsilon), mx), dim=1)


def implicit_losses(output, target, color_mode='rgb', semantic_classes=0):
    """The four per-frame loss heads of loss.py:50-194 on one frame's (N, G) logits and (N, 6) targets
    (density, R, G, B, mark_track, segm) -> dict of scalar tensors (differentiable in `output`).
    Means over empty selections are NaN, as in torch."""
    F = torch.nn.functional
    res = {}
    res['dens'] = F.binary_cross_entropy_with_logits(output[..., 0], target[..., 0])        # loss.py:57-59
    solid = target[..., 0] >= 0.1
    sel = torch.logical_and(solid, target[..., 1] >= 0.0)                                   # loss.py:73-77
    o, t = output[sel], target[sel]
    if color_mode in ('rgb', 'rgb_nosigmoid'):
        res['rgb'] = F.l1_loss(o[..., 1:4], t[..., 1:4])                                    # loss.py:79-83
    else:
        hsv = rgb_to_hsv(t[..., 1:4])
        sat, val = hsv[..., 1], hsv[..., 2]
        if color_mode == 'hsv':                                                             # loss.py:85-116
            nc = 12
            hue = torch.round(hsv[..., 0] / 360.0 * nc).type(torch.int64)
            hue[hue == nc] = 0
            vivid = torch.logical_and(sat >= 0.2, val >= 0.2)
            loss_hue = F.cross_entropy(o[..., 1:1 + nc][vivid], hue[vivid]) / 2.0 if vivid.sum() >= 16 else 0.0
            res['rgb'] = (loss_hue + F.l1_loss(o[..., 1 + nc], sat) + F.l1_loss(o[..., 2 + nc], val)) / 3.0
        else:                                                                               # loss.py:117-149
            nc = 6
            cls = torch.round(hsv[..., 0] / 360.0 * nc).type(torch.int64)
            cls[cls == nc] = 0
            bland = torch.logical_or(sat < 0.3, val < 0.3)
            cls[torch.logical_and(val < 0.2, bland)] = nc
            cls[torch.logical_and(torch.logical_and(0.2 <= val, val < 0.6), bland)] = nc + 1
            cls[torch.logical_and(0.6 <= val, bland)] = nc + 2
            res['rgb'] = F.cross_entropy(o[..., 1:1 + nc + 3], cls) / 3.0
    if semantic_classes > 0:                                                                # loss.py:163-168
        tag = target[..., -1].type(torch.int64)
        ok = tag >= 0
        res['segm'] = F.cross_entropy(output[..., -semantic_classes:][ok], tag[ok])
    ti = track_idx(color_mode)                                                              # loss.py:182-192
    if output.shape[-1] > ti:
        sel = torch.logical_and(solid, target[..., 4] >= 0.0)
        res['track'] = F.binary_cross_entropy_with_logits(output[sel][..., ti], target[sel][..., 4])
    return res
