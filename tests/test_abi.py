"""CPU: the C-ABI library loads and exports every symbol include/spi_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'spi_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(spi_[a-z0-9_]+)\s*\(', text)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for must in ('spi_bias_act', 'spi_upfirdn2d', 'spi_filtered_lrelu', 'spi_filtered_lrelu_act', 'spi_render_forward',
                 'spi_render_backward', 'spi_points_forward', 'spi_rotate', 'spi_adam_step'):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from spi_b200 import _lib, build
    build.build(verbose=False)
    lib = ctypes.CDLL(_lib.lib_path())
    for name in declared_symbols():
        assert hasattr(lib, name), f'{name} declared in include/spi_b200.h but not exported'
    assert set(_lib.EXPORTS) <= set(declared_symbols())


def test_product_fails_loudly_without_cuda():
    import torch
    from spi_b200.torch_utils.ops import bias_act, upfirdn2d, filtered_lrelu
    x = torch.zeros(1, 2, 4, 4)
    for fn in (lambda: bias_act.bias_act(x), lambda: upfirdn2d.upfirdn2d(x, None), lambda: filtered_lrelu.filtered_lrelu(x)):
        with pytest.raises(RuntimeError):
            fn()


def test_product_never_imports_oracle():
    import subprocess, sys
    code = ("import sys; import spi_b200.training.triplane, spi_b200.torch_utils.ops.bias_act; "
            "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'product imported the oracle'")
    subprocess.check_call([sys.executable, '-c', code], cwd=ROOT)
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'spi_b200')):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), f


def test_library_reports_the_abi_version_the_bindings_expect():
    """A stale libspi_b200.so (older signatures) must fail at load, not misread arguments: the loader compares spi_abi_version()."""
    from spi_b200 import _lib
    lib = _lib.load()
    assert lib.spi_abi_version() == _lib.ABI_VERSION == 3
    header = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'include', 'spi_b200.h')).read()
    assert 'spi_abi_version(void);' in header and '3 for this header' in header
