"""ctypes binding of libsegland_b200.so (the C ABI in include/segland_b200.h).

There is deliberately no fallback: if the library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_longlong, c_size_t, c_void_p

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, 'lib', 'libsegland_b200.so')

_P = c_void_p
_SIGNATURES = {
    # name: argtypes (restype is c_int unless noted)
    'sl_abi_version': [],
    'sl_check_device': [],
    'sl_env_reload': [],
    'sl_pop_prepare': [_P, c_int, c_int, c_int] + [_P] * 19,
    'sl_pop_fg_lowres': [_P, c_int, c_int, c_int, _P, _P, _P, c_int, _P, c_int, POINTER(c_int), _P],
    'sl_pop_bg_simt': [_P, c_int, c_int, c_int, _P, _P, _P, _P, c_int, c_int, _P],
    'sl_pop_bg_tc': [_P, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P, c_int, _P, _P, c_int, c_int, _P],
    'sl_pop_head_tc': [_P, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P, c_int, _P, _P, _P, c_int, POINTER(c_int), _P, _P,
                       c_int, c_int, _P],
    'sl_pop_head_bwd': [_P, c_int, c_int, c_int, _P, _P, _P, c_int, POINTER(c_int), _P, _P, _P, _P, c_int, c_int] + [_P] * 7
                       + [c_int, _P, _P],
    'sl_pop_prepare_bwd': [_P, c_int, c_int, c_int] + [_P] * 21,
    'sl_views_reduce': [_P, c_int, c_int, c_int, c_int, c_int, POINTER(c_int), c_float, _P, _P],
    'sl_window_accumulate': [_P, c_longlong, c_longlong, c_int, c_int, c_int, c_int, POINTER(c_int), c_int, POINTER(c_int), c_int,
                             POINTER(c_int), c_int, c_int, c_int, _P, _P, _P],
    'sl_upsample_argmax': [_P, c_int, c_int, c_int, c_int, c_int, c_int, _P, c_int, _P, _P, _P, _P, _P, _P],
    'sl_pseudo_label': [_P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P],
    'sl_confusion': [_P, _P, c_longlong, c_int, c_int, _P, _P, _P],
    'sl_inter_union': [_P, _P, c_longlong, c_int, c_int, _P, _P, _P, _P, _P],
    'sl_map_proto': [_P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P],
    'sl_orth_loss': [_P, c_int, _P, c_int, c_int, _P, _P, _P, _P],
    'sl_fuse_argmax_tiles': [POINTER(_P), c_int, c_int, c_int, c_longlong, c_int, _P, _P, _P, c_int, _P, _P],
    'sl_orth_from_sim': [_P, c_int, c_int, _P, _P, _P],
    'sl_fuse_argmax': [POINTER(_P), c_int, c_int, c_longlong, c_int, _P, _P, _P, c_int, _P, _P],
    'sl_upsample_ce_fwd': [_P, c_int, c_int, c_int, c_int, c_int, c_int, _P, c_int, _P, _P, _P, _P],
    'sl_tail_layernorm': [_P, c_int, c_int, c_int, _P, _P, c_float, _P, _P],
    'sl_tail_conv_prepare': [_P, c_int, c_int, _P, _P, _P],
    'sl_tail_bn_relu_conv': [_P, c_int, c_int, c_int, _P, _P, _P, _P, c_float, c_int, _P, _P, _P, c_int, _P, _P, _P],
    'sl_tail_sum': [POINTER(_P), c_int, c_longlong, _P, _P],
    'sl_tail_bn_relu': [_P, c_int, c_int, c_int, _P, _P, _P, _P, c_float, c_int, _P, _P],
    'sl_tail_concat': [POINTER(_P), POINTER(c_int), c_int, c_int, c_int, _P, _P],
    'sl_upsample_ce_bwd': [_P, c_int, c_int, c_int, c_int, c_int, c_int, _P, c_int, _P, _P, _P, _P, _P],
}

_lib = None


class SeglandError(RuntimeError):
    def __init__(self, fn, code, msg):
        super().__init__(f'{fn} failed with code {code}: {msg}')
        self.fn, self.code = fn, code


def lib():
    """Load (once) and return the shared library.  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f'{LIB_PATH} is missing: build it with `python -m segland_b200.build` '
                '(nvcc, sm_100a).  segland_b200 has no CPU or PyTorch fallback.')
        handle = ctypes.CDLL(LIB_PATH)
        for name, argtypes in _SIGNATURES.items():
            fn = getattr(handle, name)           # AttributeError if the .so lacks a declared symbol
            fn.argtypes = argtypes
            fn.restype = c_int
        handle.sl_error_string.argtypes = [c_int]
        handle.sl_error_string.restype = c_char_p
        handle.sl_pop_bg_tc_ws_bytes.argtypes = [c_int, c_int, c_int]
        handle.sl_pop_bg_tc_ws_bytes.restype = c_size_t
        handle.sl_pop_head_bwd_ws_bytes.argtypes = [c_int, c_int, c_int, c_int]
        handle.sl_pop_head_bwd_ws_bytes.restype = c_size_t
        handle.sl_pop_prepare_bwd_ws_bytes.argtypes = [c_int, c_int]
        handle.sl_pop_prepare_bwd_ws_bytes.restype = c_size_t
        handle.sl_pop_prepare_ws_bytes.argtypes = [c_int, c_int]
        handle.sl_pop_prepare_ws_bytes.restype = c_size_t
        handle.sl_upsample_ce_ws_bytes.argtypes = [c_int, c_int, c_int, c_int, c_int]
        handle.sl_upsample_ce_ws_bytes.restype = c_size_t
        handle.sl_tail_bn_relu_conv_ws_bytes.argtypes = [c_int, c_int, c_int]
        handle.sl_tail_bn_relu_conv_ws_bytes.restype = c_size_t
        if handle.sl_abi_version() != 1:
            raise ImportError(f'{LIB_PATH}: ABI version {handle.sl_abi_version()} != 1; rebuild')
        _lib = handle
    return _lib


def set_env(**switches):
    """Set (value) or clear (None) SL_* environment switches and make the library re-read them: the library caches
    its switches at first use, so changing os.environ alone has no effect on a loaded library."""
    for k, v in switches.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = str(v)
    lib().sl_env_reload()


def exported_names():
    return list(_SIGNATURES) + ['sl_error_string', 'sl_pop_bg_tc_ws_bytes', 'sl_pop_prepare_ws_bytes', 'sl_upsample_ce_ws_bytes',
                                     'sl_pop_head_bwd_ws_bytes', 'sl_pop_prepare_bwd_ws_bytes', 'sl_tail_bn_relu_conv_ws_bytes']


def call(name, *args):
    """Invoke an entry point; non-zero return codes raise SeglandError."""
    handle = lib()
    rc = getattr(handle, name)(*args)
    if rc != 0:
        raise SeglandError(name, rc, handle.sl_error_string(rc).decode())
    return rc


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else c_void_p(t.data_ptr())


def int_array(values):
    return (c_int * len(values))(*[int(v) for v in values])


def ptr_array(tensors):
    return (c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
