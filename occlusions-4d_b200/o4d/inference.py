"""Test-time driver of the continuous representation: the per-frame loop of
eval/inference.py:175-248 and the solid / air split of :283-284 with everything resident on the GPU.

The reference builds the query lattice with numpy, uploads every mini-batch, squashes the outputs with five
small torch kernels, and copies every mini-batch back to the host (two host syncs per mini-batch).  Here the
lattice is generated on the device (o4d_grid_queries_f32, bit-identical to the numpy code), every mini-batch
of the decoder writes straight into its rows of ONE (N, G) device buffer, the squashing is one in-place kernel
over the whole frame (o4d_output_activation_f32), and there is at most one D2H copy per frame.
"""
import numpy as np
import torch

from . import geometry, ops


def query_points(num_sample, min_z, cube_bounds, time_idx, data_kind, cube_mode, point_sample_mode, device):
    """(N, 4) fp32 query points on `device` (eval/inference.py:175-177).  'grid' is generated on the GPU; 'random'
    keeps the reference's numpy draws (np.random.rand) and uploads them once."""
    if point_sample_mode == 'grid':
        return geometry.sample_implicit_points_blind_device(num_sample, min_z, cube_bounds, time_idx, data_kind,
                                                            cube_mode, device)
    pts = geometry.sample_implicit_points_blind_numpy(num_sample, min_z, cube_bounds, time_idx, data_kind,
                                                      cube_mode, point_sample_mode)
    return torch.from_numpy(pts).to(device)


def query_frame(implicit_net, pcl_abstract, features_global, points_query, batch_size=32768, color_mode='rgb',
                predict_segmentation=False, semantic_classes=0, track_mode='none', density_threshold=0.5,
                to_host=False):
    """One track of one frame: decoder over all mini-batches + squashing + solid / air split.
    pcl_abstract (M, 3 + E), features_global (D,), points_query (N, 4) on the GPU ->
    dict(implicit_output (N, G) squashed as eval/inference.py:218-243, solid_mask (N,) bool = density >=
    density_threshold, points_io (N, 4 + G) = queries next to outputs (:280-281)); numpy arrays with
    to_host=True (one D2H copy), device tensors otherwise."""
    assert points_query.is_cuda and points_query.dim() == 2, 'points_query must be an (N, 4) CUDA tensor'
    assert pcl_abstract.dim() == 2 and features_global.dim() == 1
    cfg = implicit_net.o4d_config()
    params = implicit_net.o4d_params()
    n = points_query.shape[0]
    with torch.no_grad():
        scene = implicit_net.o4d_scene(pcl_abstract, features_global)
        q = ops._f32(points_query, 'points_query').contiguous()
        out = torch.empty((n, cfg.d_out), dtype=torch.float32, device=q.device)
        for s in range(0, n, int(batch_size)):
            e = min(s + int(batch_size), n)
            ops.decoder_forward(cfg, params, scene, q[s:e], want_penult=False, out=out[s:e])
        col_ops = geometry.inference_column_ops(
            cfg.d_out, color_mode=color_mode, predict_segmentation=predict_segmentation,
            semantic_classes=semantic_classes, track_mode=track_mode,
            output_track_idx={'rgb': 4, 'rgb_nosigmoid': 4, 'hsv': 15, 'bins': 10}[color_mode])
        if n > 0:
            geometry.output_activation(out, col_ops)
        points_io = torch.cat([q, out], dim=-1)
        solid = out[:, 0] >= density_threshold
    res = {'implicit_output': out, 'solid_mask': solid, 'points_io': points_io}
    if to_host:
        host = points_io.cpu().numpy()                       # the frame's single D2H copy
        res = {'implicit_output': host[:, q.shape[1]:], 'points_io': host,
               'solid_mask': host[:, q.shape[1]] >= np.float32(density_threshold)}
    return res


def run_frame(pcl_net, implicit_net, pcl_input, num_sample, min_z, cube_bounds, time_idx, data_kind, cube_mode,
              point_sample_mode='grid', **kwargs):
    """Encoder + query_frame for one clip and frame (eval/inference.py:195-248 with track_mode none / one):
    pcl_input (1, N, d_in) on the GPU -> query_frame's dict plus pcl_abstract / features_global."""
    with torch.no_grad():
        pcl_abstract, features_global, _ = pcl_net(pcl_input, False)
    pcl_abstract, features_global = pcl_abstract.squeeze(0), features_global.squeeze(0)
    points_query = query_points(num_sample, min_z, cube_bounds, time_idx, data_kind, cube_mode, point_sample_mode,
                                pcl_input.device)
    res = query_frame(implicit_net, pcl_abstract, features_global, points_query, **kwargs)
    res['pcl_abstract'], res['features_global'] = pcl_abstract, features_global
    res['points_query'] = points_query
    return res
