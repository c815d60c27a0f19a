"""CPU: the C-ABI library loads and exports every symbol include/segland_b200.h declares
(no compute calls -- there is no GPU here)."""
import os
import re

import pytest

from conftest import ROOT


def header_functions():
    src = open(os.path.join(ROOT, 'include', 'segland_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return re.findall(r'SL_API\s+(?:const\s+)?\w+\s*\*?\s*(sl_\w+)\s*\(', src)


def test_header_declares_expected_surface():
    names = header_functions()
    assert len(names) == len(set(names)) >= 15
    for must in ('sl_pop_prepare', 'sl_pop_fg_lowres', 'sl_pop_bg_simt', 'sl_pop_bg_tc', 'sl_upsample_argmax',
                 'sl_confusion', 'sl_inter_union', 'sl_map_proto', 'sl_orth_loss', 'sl_fuse_argmax',
                 'sl_pseudo_label', 'sl_views_reduce'):
        assert must in names


def test_library_exports_every_declared_symbol():
    import ctypes
    from segland_b200 import _cabi
    if not os.path.exists(_cabi.LIB_PATH):
        from segland_b200 import build
        build.build()
    handle = ctypes.CDLL(_cabi.LIB_PATH)
    for name in header_functions():
        assert hasattr(handle, name), f'{name} declared in the header but not exported'
    # and the binding table covers exactly the header
    assert sorted(_cabi.exported_names()) == sorted(header_functions())
    lib = _cabi.lib()
    assert lib.sl_abi_version() == 1
    assert b'alignment' in lib.sl_error_string(-3)


def test_no_cpu_fallback_without_cuda():
    import torch
    from segland_b200 import ops
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        ops.check_device()
    with pytest.raises(RuntimeError):
        ops.PopHead(torch.zeros(7, 64), (torch.zeros(64, 64), torch.zeros(64, 64), torch.zeros(64)))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'segland_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', text, flags=re.M), f
