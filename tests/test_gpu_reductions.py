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


@pytest.mark.parametrize('shape', [(1, 64, 32, 32), (2, 128, 16, 24), (1, 512, 4, 4)])
def test_maxpool2x2_matches_torch_including_ties(lib, shape):
    from spi_b200.ops.resize import maxpool2x2
    gen = torch.Generator().manual_seed(sum(shape))
    x = torch.relu(torch.randn(*shape, generator=gen))            # post-ReLU activations: windows of all zeros tie
    dy = torch.randn(shape[0], shape[1], shape[2] // 2, shape[3] // 2, generator=gen)
    xo = x.clone().requires_grad_(True)
    yo = torch.nn.functional.max_pool2d(xo, 2, 2)
    yo.backward(dy)
    xg = x.cuda().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    yg = maxpool2x2(xg)
    yg.backward(dy.cuda())
    assert torch.equal(yg.cpu(), yo) and torch.equal(xg.grad.cpu(), xo.grad)
