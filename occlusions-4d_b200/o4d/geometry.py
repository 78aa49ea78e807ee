"""Host-side helpers mirrored from the reference's utils/geometry.py that sit directly on
either side of the hot path: the test-time query generator and the decoder's local kNN.
"""
import numpy as np
import torch
import torch.nn.functional as F

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


# ------------------------------------------------------------------------------------------------
# Training-time query sampler and the dataset-side subsampling (row 4).
# Device work: o4d_filter_air_solid_gap_f32 / o4d_filter_bounds_f32 / o4d_fps_f32 (csrc/sampler.cu,
# csrc/fps*.cu).  Random draws follow the reference's order and generators (CPU torch / numpy
# draws moved to the device, torch.rand on the device for the blind cuboid samples), so a seeded run
# reproduces the reference's sample for sample; `device_rng=True` draws everything on the GPU instead.

_CARLA_CUBOID = {1: (2.0, 1.0, 0.5), 2: (2.4, 0.8, 0.4), 3: (2.2, 1.0, 0.4), 4: (2.5, 1.0, 0.4)}


def _filter_call(rows, fn_name, launch):
    """Shared allocation / launch plumbing of the two compaction entry points."""
    import ctypes
    from . import _lib
    L = _lib.lib()
    with torch.cuda.device(rows.device):
        count = torch.empty((1,), dtype=torch.int32, device=rows.device)
        ws = ops.workspace(rows.device, L.o4d_filter_workspace_bytes(rows.shape[0]), slot=4)
        rc = launch(L, count, ws, ctypes)
    _lib.check(rc, fn_name)
    return count


def filter_select(to_filter, target_coords, point_occupancy_radius, num_select):
    """filter_air_solid_gap followed by select_safely on rows and distances (utils/geometry.py:1164-1196,
    1095-1105) as ONE device call with fixed-size outputs and no host synchronisation.
    to_filter (N, D), target_coords (M, >=3) -> rows (num_select, D), dists (num_select,), count (1,) int32
    (device; the number of candidates that survived the filter)."""
    cand, ldc = ops._rows(to_filter, 'to_filter')
    tgt, ldt = ops._rows(target_coords, 'target_coords')
    n, d = cand.shape
    out = torch.empty((num_select, d), dtype=torch.float32, device=cand.device)
    dist = torch.empty((num_select,), dtype=torch.float32, device=cand.device)

    def launch(L, count, ws, ctypes):
        return L.o4d_filter_air_solid_gap_f32(
            ops._ptr(cand), n, d, ldc, ops._ptr(tgt), tgt.shape[0], ldt, float(point_occupancy_radius),
            int(num_select), ops._ptr(out), d, ops._ptr(dist), ops._ptr(count), ops._ptr(ws), ws.numel(),
            ops._stream(cand))
    assert num_select > 0
    count = _filter_call(cand, 'o4d_filter_air_solid_gap_f32', launch)
    return out, dist, count


def filter_air_solid_gap(to_filter, target_coords, target_slice_size, point_occupancy_radius):
    """Reference signature (utils/geometry.py:1164-1196): rows of to_filter (N, D) farther than the radius from
    every target point -> (rows (N', D), dists (N',), good_ratio).  target_slice_size is accepted and ignored:
    no distance matrix is materialised, so the target cloud is never sliced.  One D2H read (N')."""
    del target_slice_size
    cand, ldc = ops._rows(to_filter, 'to_filter')
    tgt, ldt = ops._rows(target_coords, 'target_coords')
    n, d = cand.shape
    out = torch.empty((n, d), dtype=torch.float32, device=cand.device)
    dist = torch.empty((n,), dtype=torch.float32, device=cand.device)

    def launch(L, count, ws, ctypes):
        return L.o4d_filter_air_solid_gap_f32(
            ops._ptr(cand), n, d, ldc, ops._ptr(tgt), tgt.shape[0], ldt, float(point_occupancy_radius),
            0, ops._ptr(out), d, ops._ptr(dist), ops._ptr(count), ops._ptr(ws), ws.numel(), ops._stream(cand))
    count = _filter_call(cand, 'o4d_filter_air_solid_gap_f32', launch)
    kept = int(count.item())
    good_ratio = count[0] / n                          # tensor, like good_mask.sum() / N
    return out[:kept], dist[:kept], good_ratio


def filter_pcl_bounds_torch(pcl, x_min=-10.0, x_max=10.0, y_min=-10.0, y_max=10.0, z_min=-10.0, z_max=10.0):
    """Rows of pcl (N, D) inside the closed cuboid, in order (utils/geometry.py:175-188).  Bounds are compared
    in fp32, as torch does for a Python scalar against an fp32 tensor.  One D2H read (the row count)."""
    rows, ld = ops._rows(pcl, 'pcl')
    n, d = rows.shape
    out = torch.empty((n, d), dtype=torch.float32, device=rows.device)

    def launch(L, count, ws, ctypes):
        lo = (ctypes.c_float * 3)(x_min, y_min, z_min)
        hi = (ctypes.c_float * 3)(x_max, y_max, z_max)
        return L.o4d_filter_bounds_f32(ops._ptr(rows), n, d, ld, lo, hi, ops._ptr(out), d, ops._ptr(count),
                                       ops._ptr(ws), ws.numel(), ops._stream(rows))
    count = _filter_call(rows, 'o4d_filter_bounds_f32', launch)
    return out[:int(count.item())]


def filter_pcl_bounds_carla_output_torch(pcl, min_z=-0.5, other_bounds=16.0, padding=0.0, cube_mode=4):
    """Output cuboid of the CARLA scenes (x >= 0), utils/geometry.py:224-260."""
    x_mul, y_mul, z_mul = _CARLA_CUBOID[cube_mode]
    return filter_pcl_bounds_torch(
        pcl, x_min=0.0 - padding, x_max=other_bounds * x_mul + padding,
        y_min=-other_bounds * y_mul - padding, y_max=other_bounds * y_mul + padding,
        z_min=min_z, z_max=other_bounds * z_mul)


def get_vehped_points(pcl, segm_idx):
    """Pedestrian (tag 4) rows followed by vehicle (tag 10) rows, utils/geometry.py:1323-1332."""
    tag = pcl[..., segm_idx]
    return torch.cat([pcl[tag == 4], pcl[tag == 10]], dim=0)


def sample_random_uniform_3ball(num_points, max_radius, min_radius=0.0, device=None):
    """Uniform points in a 3-ball shell (utils/geometry.py:562-575): normalised Gaussian direction times
    cbrt(U) radius.  device=None draws like the reference (torch CPU generator for the direction, numpy for
    the radius); a CUDA device draws both from that device's torch generator."""
    if device is None:
        direction = F.normalize(torch.randn(num_points, 3, dtype=torch.float32), p=2, dim=-1)
        radius = torch.tensor(np.cbrt(np.random.rand(num_points).astype(np.float32)))
    else:
        direction = F.normalize(torch.randn(num_points, 3, dtype=torch.float32, device=device), p=2, dim=-1)
        radius = torch.rand(num_points, dtype=torch.float32, device=device).pow(1.0 / 3.0)
    radius = radius * (max_radius - min_radius) + min_radius
    return direction * radius[:, None]


def sample_implicit_points_blind_torch(data_kind, num_sample, cube_mode, cube_bounds, min_z, device):
    """(num_sample, 3) uniform points in the output cuboid, drawn on `device` in the reference's order
    (utils/geometry.py:1108-1161): GREATER draws (n, 2) for x, y then (n, 1) for z; CARLA three (n, 1) draws."""
    if data_kind == 'greater':
        xy = torch.rand((num_sample, 2), device=device) * cube_bounds * 2.0 - cube_bounds
        z = torch.rand((num_sample, 1), device=device) * (cube_bounds - min_z) + min_z
        return torch.cat([xy, z], dim=-1)
    if data_kind == 'carla':
        if cube_mode not in _CARLA_CUBOID:
            raise ValueError()
        x_mul, y_mul, z_mul = _CARLA_CUBOID[cube_mode]
        x = torch.rand((num_sample, 1), device=device) * cube_bounds * x_mul
        y = torch.rand((num_sample, 1), device=device) * cube_bounds * (2.0 * y_mul) - cube_bounds * y_mul
        z = torch.rand((num_sample, 1), device=device) * (cube_bounds * z_mul - min_z) + min_z
        return torch.cat([x, y, z], dim=-1)
    raise ValueError()


def subsample_pad_pcl_torch(pcl, n_desired, sample_mode='random', subsample_only=False,
                            retain_vehped=False, segm_idx=None, fps_start_idx=0):
    """Dataset-side size normalisation of a cloud (utils/geometry.py:295-380): zero padding
    when too small, random or farthest-point subsampling (B = 1) when too large.  For CUDA inputs the
    farthest-point branch runs the cluster FPS kernel (o4d_fps_f32) instead of torch_cluster.fps on a CPU
    worker: count = ceil((n_desired / N - 1e-7) * N), sorted indices, deterministic start (fps_start_idx; the
    reference's start is random)."""
    assert sample_mode in ['random', 'farthest_point']
    no_batch = (pcl.dim() == 2)
    if no_batch:
        pcl = pcl.unsqueeze(0)
    (B, N, D) = pcl.shape
    if N < n_desired:
        if subsample_only:
            raise RuntimeError('Too few input points: ' + str(N) + ' vs ' + str(n_desired) + '.')
        result = torch.cat((pcl, torch.zeros((B, n_desired - N, D), dtype=pcl.dtype, device=pcl.device)), dim=1)
    elif N > n_desired:
        assert B == 1
        n_remain = n_desired
        retained = None
        if retain_vehped:
            tag = pcl[0, :, segm_idx]
            retained = pcl[0][(tag == 4) | (tag == 10)]
            pool = torch.nonzero(tag != 10).flatten().cpu().numpy()
            n_remain -= retained.shape[0]
        else:
            pool = np.arange(N)
        if sample_mode == 'random':
            inds = np.random.choice(pool, n_remain, replace=False)
            inds.sort()
            result = pcl[:, torch.as_tensor(inds, device=pcl.device)]
        else:
            assert not retain_vehped
            # torch_cluster: count = ceil(N * ratio) evaluated in src.dtype (fp32)
            n_out = int(np.ceil(np.float32(N) * np.float32(n_remain / N - 1e-7)))
            inds = ops.fps(pcl[0, :, :3], n_out, start_idx=fps_start_idx)
            result = pcl[:, inds]
        if retained is not None:
            result = torch.cat([retained, result[0]], dim=0)[None]
        assert result.shape[1] == n_desired
    else:
        result = pcl
    return result.squeeze(0) if no_batch else result


class GuidedImplicitPointSampler(torch.nn.Module):
    """Training-time sampler of solid / air query points and their targets for one frame
    (utils/geometry.py:578-1105); same constructor, forward signature and 6-tuple result
    (solid_input, air_input, solid_target, air_target, solid_shares, air_shares).

    Differences from the reference, none of which change the samples of a seeded run:
    * every air pool (moving / near-solid-query / near-target / regular) is ONE device call
      (`filter_select`): 1-NN rejection against the whole target cloud, ordered compaction and
      select_safely's wrap-around, fixed-size outputs, no `.item()` / boolean-mask sync per pool
      (the reference syncs twice per pool and slices the target cloud to bound its distance matrix);
    * insufficient-pool warnings (select_safely) need the kept count on the host, so they are only
      emitted with `sync_warnings=True`;
    * `device_rng=True` draws indices and offsets on the GPU generator (no H2D copies); the default
      keeps the reference's CPU draws so results are reproducible against it.
    """

    def __init__(self, logger, min_z=-1.0, cube_bounds=10.0, point_occupancy_radius=0.25,
                 num_solid=1024, num_air=1024, predict_segmentation=False, semantic_classes=13,
                 predict_tracking=False, data_kind='', point_sample_bias='none', cube_mode=4,
                 device_rng=False, sync_warnings=False):
        super().__init__()
        self.logger = logger
        self.min_z = min_z
        self.cube_bounds = cube_bounds
        self.point_occupancy_radius = point_occupancy_radius
        self.num_solid = num_solid
        self.num_air = num_air
        self.predict_segmentation = predict_segmentation
        self.semantic_classes = semantic_classes
        self.predict_tracking = predict_tracking
        self.data_kind = data_kind
        self.point_sample_bias = point_sample_bias
        self.cube_mode = cube_mode
        self.low_prefer_min_z = 0.0
        self.low_prefer_max_z = 2.0
        self.device_rng = device_rng
        self.sync_warnings = sync_warnings

    # ---- random draws (reference order: index draw first, then the offset ball) ----
    def _pick(self, pool, count):
        """count rows of pool drawn with replacement (torch.randint on the CPU generator)."""
        if self.device_rng:
            return pool[torch.randint(0, pool.shape[0], (count,), device=pool.device)]
        return pool[torch.randint(0, pool.shape[0], (count,)).to(pool.device)]

    def _ball(self, count, max_radius, min_radius, device):
        if self.device_rng:
            return sample_random_uniform_3ball(count, max_radius, min_radius, device=device)
        return sample_random_uniform_3ball(count, max_radius, min_radius).to(device)

    def _columns(self):
        carla = self.data_kind == 'carla'
        return (4 if carla else 3), (5 if carla else 3), (6 if carla else 4)   # instance, semantic, view

    def _frame_cloud(self, frame, sizes, i, what):
        """Valid rows of batch item i, cropped to the output cuboid for CARLA (:670-689)."""
        cloud = frame[i, :int(sizes[i].item())]
        if self.data_kind == 'carla':
            cloud = filter_pcl_bounds_carla_output_torch(
                cloud, min_z=self.min_z, other_bounds=self.cube_bounds, cube_mode=self.cube_mode)
        if cloud.shape[0] < 256:
            raise RuntimeError('Invalid due to %s: %d' % (what, cloud.shape[0]))
        return cloud

    def forward(self, pcl_target, pcl_target_size, valo_ids, num_valo_ids, time_idx):
        frame = pcl_target[time_idx]
        sizes = pcl_target_size[time_idx]
        (B, M, E) = frame.shape
        assert torch.all(sizes <= M)
        assert E == {'greater': 9, 'carla': 11}.get(self.data_kind, E)

        other_frame = other_sizes = None
        if len(pcl_target) > 1:                                  # :650-657
            other_time = np.random.randint(len(pcl_target) - 1)
            if other_time == time_idx:
                other_time += 1
            other_frame, other_sizes = pcl_target[other_time], pcl_target_size[other_time]

        per_item = []
        for i in range(B):
            cloud = self._frame_cloud(frame, sizes, i, 'cur_tgt_pcl_count')
            ids = sorted(list(valo_ids[i, :int(num_valo_ids[i].item())].detach().cpu().numpy()))
            tgt_unique = other_unique = None
            if 'moving' in self.point_sample_bias:               # :697-735
                other = self._frame_cloud(other_frame, other_sizes, i, 'cur_other_pcl_count')
                max_slice = int((2 ** 27) // self.num_air)
                head = cloud.shape[0] // int(np.ceil(cloud.shape[0] / max_slice)) + 1
                a, b = cloud[:head], other[:head]
                gap = self.point_occupancy_radius * 2.0
                tgt_unique = filter_air_solid_gap(a, b[..., :3], head, gap)[0]
                other_unique = filter_air_solid_gap(b, a[..., :3], head, gap)[0]
            solid = self.construct_solid_input_target(cloud, tgt_unique, ids, time_idx)
            air = self.construct_air_input_target(cloud, other_unique, solid[0], ids, time_idx)
            per_item.append((solid[0], air[0], solid[1], air[1], solid[2], air[2]))
        return tuple(torch.stack(col) for col in zip(*per_item))

    def construct_solid_input_target(self, cur_tgt_pcl, cur_tgt_unique, cur_valo_ids, time_idx):
        """(S, 4) solid queries near target points and their (S, 6) targets (:764-938)."""
        inst_idx, segm_idx, view_idx = self._columns()
        bias = self.point_sample_bias
        # (regular, low, moving, vehped, ivalo, sembal) -- fp32 arithmetic as in the reference
        shares = torch.tensor([1.0, 0.0, 0.0, 0.0, 0.0, 0.0])
        pools = {}

        def ramp(rows, top):
            # full share at >= 256 rows, proportional from 16 rows, nothing below
            if rows >= 256:
                return top
            return rows * top / 256.0 if rows >= 16 else 0.0

        if 'low' in bias:
            z = cur_tgt_pcl[..., 2]
            pools[1] = cur_tgt_pcl[torch.logical_and(self.low_prefer_min_z <= z, z <= self.low_prefer_max_z)]
            if pools[1].shape[0] >= 256:
                shares[1] += 1.0
        if 'moving' in bias:
            pools[2] = cur_tgt_unique
            inc = ramp(cur_tgt_unique.shape[0], 0.4)
            if inc:
                shares[2] += inc
        if 'vehped' in bias:
            assert self.data_kind == 'carla'
            pools[3] = get_vehped_points(cur_tgt_pcl, segm_idx)
            inc = ramp(pools[3].shape[0], 0.2)
            if inc:
                shares[3] += inc
        if 'ivalo' in bias:
            assert self.data_kind == 'carla'
            if len(cur_valo_ids) > 0:
                seen = cur_tgt_pcl[..., view_idx] == 0
                visible = get_vehped_points(cur_tgt_pcl[seen], segm_idx)[..., inst_idx]
                visible_ids = set(visible.type(torch.int32).unique().detach().cpu().numpy().tolist())
                hidden = get_vehped_points(cur_tgt_pcl[~seen], segm_idx)
                parts = []
                for valo_id in cur_valo_ids:
                    inst = hidden[hidden[..., inst_idx] == valo_id]
                    # an instance that is never visible in this frame counts twice
                    parts += [inst] if int(valo_id) in visible_ids else [inst, inst]
                pools[4] = torch.cat(parts, dim=0)
                inc = ramp(pools[4].shape[0], 0.2)
                if inc:
                    shares[4] += min(inc, 0.2)
        if 'sembal' in bias:
            assert self.data_kind == 'carla'
            shares[5] += 0.4
        shares /= shares.sum()

        picked = []
        counts = [0] * 6
        for slot in (1, 2, 3, 4):
            counts[slot] = int(shares[slot] * self.num_solid)
            if counts[slot] > 0:
                picked.append(self._pick(pools[slot], counts[slot]))
        want_sembal = int(shares[5] * self.num_solid)
        if want_sembal > 0:                                      # :884-903
            tags = cur_tgt_pcl[..., segm_idx]
            present = list(tags.type(torch.int32).unique().detach().cpu().numpy())
            for tag in present:
                rows = cur_tgt_pcl[tags == tag]
                if rows.shape[0] >= 16:
                    picked.append(self._pick(rows, want_sembal // len(present)))
                    counts[5] += want_sembal // len(present)
        counts[0] = self.num_solid - sum(counts[1:])
        if counts[0] > 0:
            picked.append(self._pick(cur_tgt_pcl, counts[0]))

        chosen = torch.cat(picked, dim=0)
        assert chosen.shape[0] == self.num_solid
        xyz = chosen[..., :3] + self._ball(self.num_solid, self.point_occupancy_radius / 2.0, 0.0,
                                           chosen.device)
        t_col = torch.full_like(xyz[..., 0:1], float(time_idx))
        ones = torch.ones_like(t_col)
        if self.predict_segmentation:
            last = chosen[..., segm_idx:segm_idx + 1].clone()
            last[last >= self.semantic_classes] = 3              # = Other
        else:
            last = -ones
        # (x, y, z, t) and (density = 1, R, G, B, mark_track, segm)
        return (torch.cat([xyz, t_col], dim=-1), torch.cat([ones, chosen[..., -4:], last], dim=-1), shares)

    def _air_pool(self, seeds, target_xyz, keep, warn):
        rows, dists, count = filter_select(seeds, target_xyz, self.point_occupancy_radius, keep)
        if warn and self.sync_warnings:
            kept = int(count.item())
            while 0 < kept < keep:
                self.logger.warning('Size %d is insufficient for %d!' % (kept, keep))
                kept *= 2
        return rows, dists

    def construct_air_input_target(self, cur_tgt_pcl, cur_other_unique, cur_solid_input, cur_valo_ids,
                                   time_idx):
        """(A, 4) free-space queries at least one radius away from every target point and their
        (A, 6) targets (:940-1093)."""
        target_xyz = cur_tgt_pcl[..., :3]
        r = self.point_occupancy_radius
        dev = cur_tgt_pcl.device
        # (regular, moving, near a solid query, near a target point)
        shares = torch.tensor([0.5, 0.0, 0.3, 0.2])
        if 'moving' in self.point_sample_bias:
            rows = cur_other_unique.shape[0]
            if rows >= 256:
                shares[1] += 0.4
            elif rows >= 16:
                shares[1] += rows * 0.4 / 256.0
        shares /= shares.sum()

        rows_all, dists_all = [], []
        num_moving = int(shares[1] * self.num_air)
        if num_moving > 0:
            draw = int(num_moving * 1.6)
            seeds = self._pick(cur_other_unique, draw)[..., :3] + self._ball(draw, r * 2.0, 0.0, dev)
            out = self._air_pool(seeds, target_xyz, num_moving, warn=False)
            rows_all.append(out[0]); dists_all.append(out[1])
        num_hsq = int(shares[2] * self.num_air)
        if num_hsq > 0:
            draw = int(num_hsq * 2.0)
            seeds = self._pick(cur_solid_input, draw)[..., :3] + self._ball(draw, r * 3.0, r, dev)
            out = self._air_pool(seeds, target_xyz, num_hsq, warn=True)
            rows_all.append(out[0]); dists_all.append(out[1])
        num_ht = int(shares[3] * self.num_air)
        if num_ht > 0:
            draw = int(num_ht * 2.0)
            seeds = self._pick(cur_tgt_pcl, draw)[..., :3] + self._ball(draw, r * 3.0, r, dev)
            out = self._air_pool(seeds, target_xyz, num_ht, warn=True)
            rows_all.append(out[0]); dists_all.append(out[1])
        num_regular = self.num_air - num_moving - num_hsq - num_ht
        if num_regular > 0:
            draw = int(num_regular * {'greater': 1.3, 'carla': 1.1}[self.data_kind])
            seeds = sample_implicit_points_blind_torch(
                self.data_kind, draw, self.cube_mode, self.cube_bounds, self.min_z, dev)
            out = self._air_pool(seeds, target_xyz, num_regular, warn=True)
            rows_all.append(out[0]); dists_all.append(out[1])

        xyz = torch.cat(rows_all, dim=0)
        assert xyz.shape[0] == self.num_air
        self.last_air_solid_dists = torch.cat(dists_all, dim=0)   # the reference builds and drops these
        air_input = torch.cat([xyz, torch.full_like(xyz[..., 0:1], float(time_idx))], dim=-1)
        # density 0; colour, mark_track and segmentation unavailable (-1)
        air_target = torch.full((self.num_air, 6), -1.0, device=dev, dtype=cur_tgt_pcl.dtype)
        air_target[..., 0] = 0.0
        return (air_input, air_target, shares)
