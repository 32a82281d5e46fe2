"""GPU parity of the small reduction helpers: spi_column_sums (decoder bias gradients over the per-sample rows)."""
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('rows,cols', [(1, 64), (1000, 36), (65536 + 17, 64), (300001, 36), (4096, 128)])
def test_column_sums(lib, rows, cols):
    from spi_b200 import _lib
    gen = torch.Generator().manual_seed(rows + cols)
    x = torch.randn(rows, cols, generator=gen)
    out = torch.empty(cols, device='cuda')
    xg = x.cuda()
    _lib.check(lib.spi_column_sums(_lib.ptr(xg), rows, cols, _lib.ptr(out), _lib.stream()))
    assert rel_l2(out, x.double().sum(0)) < 1e-5


def test_column_sums_rejects_bad_shapes(lib):
    from spi_b200 import _lib
    x = torch.zeros(8, 6, device='cuda')
    with pytest.raises(RuntimeError):
        _lib.check(lib.spi_column_sums(_lib.ptr(x), 8, 6, _lib.ptr(torch.empty(6, device='cuda')), _lib.stream()))
