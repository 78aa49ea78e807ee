"""ctypes binding of libo4d.so (C ABI declared in include/o4d.h).

There is NO fallback: if the shared library is missing or a call fails, the caller gets a
RuntimeError.  Build it with ``python __graft_entry__.py`` (or ``make -C csrc``).
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(os.path.dirname(_HERE), 'csrc')
# O4D_LIB selects another build of the same library (e.g. the cycle-stamp build `make STAMPS=1 OUT=... BUILD=...`)
LIB_PATH = os.environ.get('O4D_LIB') or os.path.join(_HERE, 'libo4d.so')

c_i64 = ctypes.c_int64
c_int = ctypes.c_int
c_size = ctypes.c_size_t
c_ptr = ctypes.c_void_p


class EncoderConfig(ctypes.Structure):
    # mirrors o4d_encoder_config
    _fields_ = [(n, ctypes.c_int32) for n in (
        'd_in', 'd_feat', 'down_blocks', 'transition_factor', 'pt_num_neighbors', 'down_neighbors',
        'norm', 'abstract_levels', 'global_dim', 'precision')]


class DecoderConfig(ctypes.Structure):
    # mirrors o4d_decoder_config
    _fields_ = [(n, ctypes.c_int32) for n in (
        'd_in', 'd_hidden', 'd_out', 'd_latent', 'd_latent_local', 'n_blocks', 'pos_encoding_freqs',
        'num_local_features', 'cross_attn_neighbors', 'cross_attn_layers', 'precision')]


# name -> (restype, argtypes); every symbol include/o4d.h declares.
SIGNATURES = {
    'o4d_last_error': (ctypes.c_char_p, []),
    'o4d_abi_version': (c_int, []),
    'o4d_has_tcgen05': (c_int, []),
    'o4d_launch_count': (ctypes.c_uint64, []),
    'o4d_profile_enable': (None, [c_int]),
    'o4d_profile_read': (c_int, [c_int, c_ptr, c_ptr, c_ptr]),
    'o4d_knn_f32': (c_int, [c_ptr, c_i64, c_i64, c_ptr, c_i64, c_i64, c_int, c_int, c_ptr, c_ptr, c_ptr]),
    'o4d_knn_two_lists_workspace_bytes': (c_size, [c_i64, c_int, c_int]),
    'o4d_knn_two_lists_f32': (c_int, [c_ptr, c_i64, c_i64, c_ptr, c_i64, c_i64, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_size,
                                      c_ptr]),
    'o4d_fps_workspace_bytes': (c_size, [c_i64, c_i64]),
    'o4d_fps_f32': (c_int, [c_ptr, c_i64, c_i64, c_i64, c_i64, c_ptr, c_ptr, c_ptr, c_size, c_ptr]),
    'o4d_linear_f32': (c_int, [c_ptr, c_i64, c_i64, c_i64, c_ptr, c_ptr, c_i64, c_ptr, c_i64, c_ptr, c_i64,
                               c_int, c_int, c_ptr]),
    'o4d_linear_pack_bytes': (c_size, [c_i64, c_i64, c_i64, c_int]),
    'o4d_linear_pack_f32': (c_int, [c_ptr, c_i64, c_i64, c_i64, c_ptr, c_ptr]),
    'o4d_linear_packed_f32': (c_int, [c_ptr, c_i64, c_i64, c_i64, c_ptr, c_ptr, c_i64, c_ptr, c_i64, c_ptr, c_i64,
                                      c_int, c_int, c_ptr]),
    'o4d_resblock_workspace_bytes': (c_size, [c_i64, c_int, c_int]),
    'o4d_resblock_forward_f32': (c_int, [c_ptr, c_i64, c_int, c_i64, c_ptr, c_ptr, c_int, c_ptr, c_ptr, c_ptr, c_i64,
                                         c_int, c_ptr, c_size, c_ptr]),
    'o4d_posenc_f32': (c_int, [c_ptr, c_i64, c_int, c_int, c_ptr, c_ptr]),
    'o4d_pt_block_workspace_bytes': (c_size, [c_i64, c_i64, c_int, c_int, c_int]),
    'o4d_pt_block_forward': (c_int, [c_ptr, c_ptr, c_i64, c_int, c_ptr, c_i64, c_ptr, c_i64, c_int, c_i64,
                                     c_ptr, c_i64, c_int, c_int, c_ptr, c_ptr, c_ptr, c_size, c_ptr]),
    'o4d_pt_layer_forward': (c_int, [c_ptr, c_ptr, c_i64, c_int, c_ptr, c_i64, c_ptr, c_i64, c_int, c_i64,
                                     c_ptr, c_i64, c_int, c_int, c_ptr, c_ptr, c_ptr, c_size, c_ptr]),
    'o4d_down_workspace_bytes': (c_size, [c_i64, c_int, c_int, c_int, c_int]),
    'o4d_down_forward': (c_int, [c_ptr, c_ptr, c_i64, c_int, c_ptr, c_i64, c_int, c_int, c_int, c_int, c_i64,
                                 c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_size, c_ptr]),
    'o4d_encoder_num_params': (c_int, [ctypes.POINTER(EncoderConfig)]),
    'o4d_encoder_num_abstract': (c_i64, [ctypes.POINTER(EncoderConfig), c_i64]),
    'o4d_encoder_workspace_bytes': (c_size, [ctypes.POINTER(EncoderConfig), c_i64]),
    'o4d_encoder_forward': (c_int, [ctypes.POINTER(EncoderConfig), c_ptr, c_ptr, c_i64, c_ptr, c_ptr, c_ptr,
                                    c_ptr, c_ptr, c_size, c_ptr]),
    'o4d_decoder_num_params': (c_int, [ctypes.POINTER(DecoderConfig)]),
    'o4d_decoder_scene_bytes': (c_size, [ctypes.POINTER(DecoderConfig), c_i64]),
    'o4d_decoder_prepare_scene': (c_int, [ctypes.POINTER(DecoderConfig), c_ptr, c_ptr, c_i64, c_i64, c_ptr,
                                          c_ptr, c_size, c_ptr]),
    'o4d_decoder_update_scene': (c_int, [ctypes.POINTER(DecoderConfig), c_ptr, c_ptr, c_i64, c_i64, c_ptr,
                                          c_ptr, c_size, c_ptr]),
    'o4d_decoder_workspace_bytes': (c_size, [ctypes.POINTER(DecoderConfig), c_i64, c_i64]),
    'o4d_decoder_forward': (c_int, [ctypes.POINTER(DecoderConfig), c_ptr, c_ptr, c_i64, c_ptr, c_i64, c_ptr,
                                    c_ptr, c_ptr, c_size, c_ptr]),
    'o4d_decoder_run_host_device_bytes': (c_size, [ctypes.POINTER(DecoderConfig), c_i64, c_i64]),
    'o4d_decoder_run_host': (c_int, [ctypes.POINTER(DecoderConfig), c_ptr, c_ptr, c_i64, c_ptr, c_i64, c_i64,
                                     c_ptr, c_ptr, c_size, c_ptr]),
    'o4d_relu_backward_f32': (c_int, [c_ptr, c_ptr, c_i64, c_ptr, c_ptr]),
    'o4d_linear_backward_workspace_bytes': (c_size, [c_i64, c_i64, c_i64]),
    'o4d_linear_backward_f32': (c_int, [c_ptr, c_i64, c_i64, c_i64, c_ptr, c_i64, c_i64, c_ptr, c_i64, c_int,
                                        c_ptr, c_i64, c_ptr, c_i64, c_ptr, c_int, c_ptr, c_size, c_ptr]),
    'o4d_attn_train_saved_bytes': (c_size, [c_i64, c_int, c_int]),
    'o4d_attn_backward_workspace_bytes': (c_size, [c_i64, c_int, c_int]),
    'o4d_attn_forward_train': (c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_ptr, c_i64, c_ptr, c_i64, c_ptr,
                                       c_i64, c_int, c_int, c_int, c_ptr, c_ptr, c_size, c_ptr]),
    'o4d_attn_backward': (c_int, [c_ptr, c_ptr, c_i64, c_ptr, c_i64, c_ptr, c_i64, c_i64, c_int, c_int, c_int,
                                  c_ptr, c_size, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_size, c_ptr]),
    'o4d_local_blend_f32': (c_int, [c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_int, c_int, c_ptr, c_ptr]),
    'o4d_local_blend_backward_f32': (c_int, [c_ptr, c_ptr, c_ptr, c_i64, c_int, c_int, c_i64, c_ptr, c_i64,
                                             c_ptr]),
    'o4d_gather_max_f32': (c_int, [c_ptr, c_i64, c_ptr, c_i64, c_int, c_int, c_ptr, c_ptr, c_ptr]),
    'o4d_gather_max_backward_f32': (c_int, [c_ptr, c_ptr, c_i64, c_int, c_i64, c_ptr, c_i64, c_ptr]),
    'o4d_layernorm_relu_f32': (c_int, [c_ptr, c_i64, c_int, c_ptr, c_ptr, ctypes.c_float, c_ptr, c_ptr]),
    'o4d_layernorm_relu_backward_f32': (c_int, [c_ptr, c_ptr, c_i64, c_int, c_ptr, c_ptr, ctypes.c_float,
                                                c_ptr, c_ptr, c_ptr, c_ptr]),
    'o4d_col_mean_f32': (c_int, [c_ptr, c_i64, c_int, c_ptr, c_ptr]),
    'o4d_col_mean_backward_f32': (c_int, [c_ptr, c_i64, c_int, c_ptr, c_ptr]),
    'o4d_grid_query_count': (c_i64, [c_i64, c_ptr, c_ptr]),
    'o4d_grid_queries_f32': (c_int, [c_ptr, c_ptr, c_ptr, ctypes.c_float, c_ptr, c_ptr]),
    'o4d_output_activation_f32': (c_int, [c_ptr, c_i64, c_int, c_ptr, c_ptr]),
    'o4d_implicit_loss_workspace_bytes': (c_size, [c_i64]),
    'o4d_implicit_loss_forward_f32': (c_int, [c_ptr, c_i64, c_int, c_i64, c_ptr, c_i64, c_int, c_int, c_int,
                                              c_ptr, c_ptr, c_ptr, c_size, c_ptr]),
    'o4d_implicit_loss_backward_f32': (c_int, [c_ptr, c_i64, c_int, c_i64, c_ptr, c_i64, c_int, c_int, c_int,
                                               c_ptr, c_ptr, c_ptr, c_i64, c_ptr]),
    'o4d_filter_workspace_bytes': (c_size, [c_i64]),
    'o4d_filter_air_solid_gap_f32': (c_int, [c_ptr, c_i64, c_int, c_i64, c_ptr, c_i64, c_i64, ctypes.c_float,
                                             c_i64, c_ptr, c_i64, c_ptr, c_ptr, c_ptr, c_size, c_ptr]),
    'o4d_filter_bounds_f32': (c_int, [c_ptr, c_i64, c_int, c_i64, c_ptr, c_ptr, c_ptr, c_i64, c_ptr, c_ptr,
                                      c_size, c_ptr]),
}

_lib = None


def build(verbose=False):
    """Compile csrc/*.cu for sm_100a into o4d/libo4d.so (nvcc cross-compiles without a GPU)."""
    out = subprocess.run(['make', '-C', _CSRC, '-j', str(os.cpu_count() or 4)],
                         capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout)
        print(out.stderr)
    if out.returncode != 0:
        raise RuntimeError('building libo4d.so failed')
    return LIB_PATH


def lib():
    """The loaded library; raises loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                'libo4d.so not found at %s -- the CUDA library is the product; there is no fallback. '
                'Build it with `python __graft_entry__.py`.' % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().o4d_last_error().decode('utf-8', 'replace')
        if rc in (-1, -2, -3) and 'assert' in what:
            raise AssertionError('%s: %s' % (what, msg))
        raise RuntimeError('%s failed (status %d): %s' % (what, rc, msg))
