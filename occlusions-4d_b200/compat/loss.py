"""Flat-name shim for the reference's ``loss.py`` (``import loss`` in pipeline.py:13).  The reference's loss.py
sits next to its entry scripts, i.e. in sys.path[0], so this shim is picked up only when the reference runs as
``python -m`` / from another directory; otherwise bind it with one line: ``import o4d.loss as loss``
(INTEGRATION.md)."""
import os as _os
import sys as _sys

_pkg_root = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
if _pkg_root not in _sys.path:
    _sys.path.insert(0, _pkg_root)
from o4d.loss import *  # noqa: F401,F403,E402
from o4d.loss import MyLosses, get_track_idx, implicit_loss_heads  # noqa: F401,E402
