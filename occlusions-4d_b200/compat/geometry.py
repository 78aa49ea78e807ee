"""Flat-name shim for the reference's ``utils/geometry.py``: everything the reference module defines stays
available (data-loader helpers, camera maths, ...), and the functions on the training-time sampler path are
replaced by the o4d implementations.  Put this directory FIRST on sys.path
(before the reference's utils/); the reference's own file is located further down sys.path and loaded under
a private name."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_here = _os.path.dirname(_os.path.abspath(__file__))
_pkg_root = _os.path.dirname(_here)
if _pkg_root not in _sys.path:
    _sys.path.insert(0, _pkg_root)


def _load_reference_geometry():
    for entry in list(_sys.path) + [_os.path.join(_os.getcwd(), 'utils')]:
        cand = _os.path.join(entry or '.', 'geometry.py')
        if _os.path.isfile(cand) and _os.path.abspath(_os.path.dirname(cand)) != _here:
            spec = _ilu.spec_from_file_location('_reference_geometry', cand)
            mod = _ilu.module_from_spec(spec)
            _sys.modules['_reference_geometry'] = mod
            spec.loader.exec_module(mod)
            return mod
    return None


_ref = _load_reference_geometry()
if _ref is not None:
    globals().update({k: v for k, v in vars(_ref).items() if not k.startswith('__')})

from o4d import geometry as _o4d_geometry  # noqa: E402

O4D_OVERRIDES = ('GuidedImplicitPointSampler', 'filter_air_solid_gap', 'my_knn_torch', 'sample_implicit_points_blind_numpy',
                 'sample_implicit_points_blind_torch', 'sample_random_uniform_3ball', 'get_vehped_points')
for _name in O4D_OVERRIDES:
    globals()[_name] = getattr(_o4d_geometry, _name)

if _ref is not None:
    _ref_subsample = _ref.subsample_pad_pcl_torch
    _ref_bounds = _ref.filter_pcl_bounds_torch

    def subsample_pad_pcl_torch(pcl, *args, **kwargs):
        """CUDA clouds take the o4d path (cluster FPS kernel).  CPU clouds are the dataset pipeline's calls from
        dataloader workers (data_greater.py:477, data_carla.py:538): that pipeline is out of scope and keeps running
        the reference's own function unchanged -- this is not a CPU fallback of an o4d op
        (o4d.geometry.subsample_pad_pcl_torch itself raises on a CPU cloud in farthest_point mode)."""
        fn = _o4d_geometry.subsample_pad_pcl_torch if pcl.is_cuda else _ref_subsample
        return fn(pcl, *args, **kwargs)

    def filter_pcl_bounds_torch(pcl, *args, **kwargs):
        fn = _o4d_geometry.filter_pcl_bounds_torch if (pcl.is_cuda and pcl.dim() == 2) else _ref_bounds
        return fn(pcl, *args, **kwargs)
else:
    subsample_pad_pcl_torch = _o4d_geometry.subsample_pad_pcl_torch
    filter_pcl_bounds_torch = _o4d_geometry.filter_pcl_bounds_torch
filter_pcl_bounds_carla_output_torch = _o4d_geometry.filter_pcl_bounds_carla_output_torch
