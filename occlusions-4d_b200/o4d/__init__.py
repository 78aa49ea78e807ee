"""o4d -- B200 (sm_100a) implementation of the occlusions-4d encoder / implicit-decoder hot
path behind the reference's own nn.Module API.

    import sys; sys.path.insert(0, '<repo>/occlusions-4d_b200')
    from o4d import model, implicit            # PointCompletionNetV3, LocalPclResnetFC

``compat/`` next to this package holds flat shim modules (``model``, ``implicit``,
``modules``, ``point_transformer_layer``) for the reference's sys.path-style imports.
"""
from . import _lib, ops  # noqa: F401
from . import point_transformer_layer, modules, model, implicit  # noqa: F401
from .model import PointCompletionNetV3  # noqa: F401
from .implicit import LocalPclResnetFC, ResnetFC, ResnetBlockFC, positional_encode  # noqa: F401
from .modules import PointTransformerBlock, DownTransition  # noqa: F401
from .point_transformer_layer import PointTransformerLayer, kNN_torch, index_points  # noqa: F401

__all__ = ['PointCompletionNetV3', 'LocalPclResnetFC', 'ResnetFC', 'ResnetBlockFC', 'positional_encode',
           'PointTransformerBlock', 'DownTransition', 'PointTransformerLayer', 'kNN_torch', 'index_points',
           'ops', 'model', 'implicit', 'modules', 'point_transformer_layer']
