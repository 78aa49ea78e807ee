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


def sample_implicit_points_blind_device(num_sample, min_z, cube_bounds, time_idx, data_kind, cube_mode, device):
    """'grid' mode of sample_implicit_points_blind_numpy generated on the GPU (o4d_grid_queries_f32):
    bit-identical values, no host array and no H2D copy of the 8.5 MB lattice per frame."""
    import ctypes
    import torch
    from . import _lib
    (x_min, x_max), (y_min, y_max), (z_min, z_max) = cuboid_bounds(min_z, cube_bounds, data_kind, cube_mode)
    ext = (ctypes.c_double * 3)(float(x_max - x_min), float(y_max - y_min), float(z_max - z_min))
    lo = (ctypes.c_double * 3)(float(x_min), float(y_min), float(z_min))
    counts = (ctypes.c_int32 * 3)()
    L = _lib.lib()
    total = L.o4d_grid_query_count(int(num_sample), ext, counts)
    if total < 0:
        raise RuntimeError('o4d_grid_query_count: bad argument')
    device = torch.device(device)
    with torch.cuda.device(device):
        out = torch.empty((total, 4), dtype=torch.float32, device=device)
        rc = L.o4d_grid_queries_f32(counts, ext, lo, float(time_idx), ops._ptr(out), ops._stream(out))
    _lib.check(rc, 'o4d_grid_queries_f32')
    return out


KEEP, SIGMOID, CLAMP01 = 0, 1, 2


def inference_column_ops(d_out, color_mode='rgb', predict_segmentation=False, semantic_classes=0,
                         track_mode='none', output_track_idx=4):
    """Per-column squashing of eval/inference.py:218-243 as op codes for output_activation."""
    col = [KEEP] * d_out
    col[0] = SIGMOID                                            # density logit -> probability (:218)
    if color_mode == 'rgb':
        col[1:4] = [SIGMOID] * 3
    elif color_mode == 'rgb_nosigmoid':
        col[1:4] = [CLAMP01] * 3
    elif color_mode == 'hsv':
        col[1:13] = [SIGMOID] * 12
        col[13:15] = [CLAMP01] * 2
    elif color_mode == 'bins':
        col[1:10] = [SIGMOID] * 9
    if predict_segmentation:
        col[d_out - semantic_classes:] = [SIGMOID] * semantic_classes
    if track_mode != 'none':
        col[output_track_idx] = SIGMOID
    return col[:d_out]


def output_activation(out, col_ops):
    """In-place squashing of the decoder output (n, g) on the device (o4d_output_activation_f32)."""
    import ctypes
    from . import _lib
    assert out.is_cuda and out.dtype.is_floating_point and out.is_contiguous() and out.dim() == 2
    g = out.shape[1]
    assert len(col_ops) == g
    arr = (ctypes.c_uint8 * g)(*[int(c) for c in col_ops])
    rc = _lib.lib().o4d_output_activation_f32(ops._ptr(out), out.shape[0], g, arr, ops._stream(out))
    _lib.check(rc, 'o4d_output_activation_f32')
    return out


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
