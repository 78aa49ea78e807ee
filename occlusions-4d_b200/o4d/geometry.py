"""Host-side helpers mirrored from the reference's utils/geometry.py that sit directly on
either side of the hot path: the test-time query generator and the decoder's local kNN.
"""
import numpy as np

from . import ops


def cuboid_bounds(min_z, cube_bounds, data_kind, cube_mode):
    """Query cuboid of the continuous representation, utils/geometry.py:1217-1245."""
    if data_kind == 'greater':
        return (-cube_bounds, cube_bounds), (-cube_bounds, cube_bounds), (min_z, cube_bounds)
    if data_kind == 'carla':
        x_mul, y_mul, z_mul = {1: (2.0, 1.0, 0.5), 2: (2.4, 0.8, 0.4), 3: (2.2, 1.0, 0.4),
                               4: (2.5, 1.0, 0.4)}[cube_mode]
        return (0.0, cube_bounds * x_mul), (-cube_bounds * y_mul, cube_bounds * y_mul), \
               (min_z, cube_bounds * z_mul)
    raise ValueError(data_kind)


def sample_implicit_points_blind_numpy(num_sample, min_z, cube_bounds, time_idx, data_kind,
                                       cube_mode, point_sample_mode):
    """(N, 4) fp32 query points (x, y, z, t) inside the cuboid, utils/geometry.py:1199-1283.
    'random': exactly num_sample uniform points (np.random.rand, x then y then z);
    'grid': cell-centred lattice with >= num_sample points, x slowest / z fastest."""
    (x_min, x_max), (y_min, y_max), (z_min, z_max) = cuboid_bounds(min_z, cube_bounds, data_kind, cube_mode)
    # plain Python floats: fp32 arrays must stay fp32 when scaled (weak-scalar promotion), as in
    # the reference.
    ext = [float(x_max - x_min), float(y_max - y_min), float(z_max - z_min)]
    lo = [float(x_min), float(y_min), float(z_min)]
    if point_sample_mode == 'random':
        cols = [np.random.rand(num_sample).astype(np.float32) * e + l for e, l in zip(ext, lo)]
        xyz = np.stack(cols, axis=-1)
    elif point_sample_mode == 'grid':
        per_unit = np.cbrt(num_sample / (ext[0] * ext[1] * ext[2]))
        counts = [int(np.ceil(per_unit * e)) for e in ext]
        axes = [(np.arange(c, dtype=np.float32) + 0.5) * (e / c) + l for c, e, l in zip(counts, ext, lo)]
        gx, gy, gz = np.meshgrid(*axes, indexing='ij')
        xyz = np.stack([gx.ravel(), gy.ravel(), gz.ravel()], axis=-1)
    else:
        raise ValueError(point_sample_mode)
    t = np.full((xyz.shape[0], 1), time_idx, dtype=np.float32)
    return np.concatenate([xyz, t], axis=-1).astype(np.float32)


def my_knn_torch(pcl_query, pcl_key, num_neighbors, bidirectional=False,
                 return_inds=False, return_knn=True, return_dists=False):
    """k nearest key rows per query by Euclidean distance, utils/geometry.py:458-503
    (o4d_knn_f32 with sqrt_dist=1; ties resolve to the lower index)."""
    assert return_inds or return_knn or return_dists
    if bidirectional:
        raise NotImplementedError()
    inds, dists = ops.knn(pcl_query, pcl_key, num_neighbors, sqrt_dist=True, return_dist=True)
    result = tuple()
    if return_inds:
        result += (inds,)
    if return_knn:
        result += (pcl_key[inds],)
    if return_dists:
        result += (dists,)
    return result
