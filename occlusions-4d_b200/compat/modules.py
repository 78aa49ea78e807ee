"""Flat-name shim: lets the reference's unmodified callers (`import modules` after its
sys.path hacks, __init__.py:51-55) pick up the o4d implementation.  Put this directory
FIRST on sys.path (before the reference's model/)."""
import os as _os
import sys as _sys

_pkg_root = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
if _pkg_root not in _sys.path:
    _sys.path.insert(0, _pkg_root)
from o4d.modules import *  # noqa: F401,F403,E402
