"""TEST INFRASTRUCTURE ONLY -- imports the UNMODIFIED reference model code.

Imports from ``/root/reference`` where it exists (the build container), else from the byte-for-byte copy
``oracle/_ref/`` that ``oracle/make_ref.py`` makes for the GPU box (baseline timing in bench.py).  Used by
``tests/golden/make_golden.py`` to generate the committed golden vectors and by the
``not gpu`` tests that pin ``oracle/o4d_oracle.py`` against the real reference.  Nothing
that runs on the GPU box may call this (``/root/reference`` is absent there).

Recipe: SURVEY.md Appendix A (stub modules first on sys.path, chdir into the reference
because ``__init__.py:51-55`` appends cwd-relative paths, flat module names).
"""
import argparse
import contextlib
import os
import sys
import warnings

_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(_HERE)
REF_ROOT = os.environ.get('O4D_REFERENCE_ROOT', '/root/reference')
if not os.path.isfile(os.path.join(REF_ROOT, 'model', 'implicit.py')):
    # GPU box: the byte-for-byte copy made by oracle/make_ref.py (git-ignored, shipped with the snapshot).
    # Model code only -- no pretrained checkpoints, no data / eval / train scripts.
    REF_ROOT = os.path.join(_HERE, '_ref')
IS_COPY = REF_ROOT == os.path.join(_HERE, '_ref')


def available():
    return os.path.isfile(os.path.join(REF_ROOT, 'model', 'implicit.py'))


_cache = {}


def load():
    """Returns a dict of the reference's flat modules: model, implicit, modules,
    point_transformer_layer, geometry, utils."""
    if _cache:
        return _cache
    if not available():
        raise RuntimeError('reference tree not present at ' + REF_ROOT)
    stubs = os.path.join(_HERE, 'ref_stubs')
    for p in (_REPO, stubs):
        if p not in sys.path:
            sys.path.insert(0, p)
    # Flat names of the reference would collide with anything of ours already imported.
    for name in ('model', 'implicit', 'modules', 'point_transformer_layer', 'geometry', 'utils'):
        if name in sys.modules:
            raise RuntimeError('module name collision: %s already imported' % name)
    cwd = os.getcwd()
    os.chdir(REF_ROOT)
    sys.path.insert(0, REF_ROOT)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            import importlib
            importlib.import_module('__init__')  # runs the reference's sys.path hacks
            for name in ('model', 'implicit', 'modules', 'point_transformer_layer',
                         'geometry', 'utils'):
                _cache[name] = importlib.import_module(name)
    finally:
        os.chdir(cwd)
    return _cache


def load_extra(name):
    """One more flat module of the reference (e.g. 'loss'), imported from inside the reference tree."""
    load()
    if name in _cache:
        return _cache[name]
    import importlib
    cwd = os.getcwd()
    os.chdir(REF_ROOT)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            _cache[name] = importlib.import_module(name)
    finally:
        os.chdir(cwd)
    return _cache[name]


def load_checkpoint(which):
    """which in {'greater','carla'} -> dict with pcl_args, implicit_args, pcl_net, implicit_net."""
    import torch
    ref = load()
    path = os.path.join(REF_ROOT, 'pretrained', which + '_checkpoint.pth')
    with torch.serialization.safe_globals([argparse.Namespace]):
        ck = torch.load(path, map_location='cpu', weights_only=True)
    ck['pcl_args']['fps_random_start'] = False  # eval/inference.py:59
    ck['implicit_net'] = ref['utils'].rename_state_dict_keys(
        ck['implicit_net'], 'pt_block.', 'pt_blocks.0.')  # eval/inference.py:62-63
    return ck


@contextlib.contextmanager
def quiet():
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        yield
