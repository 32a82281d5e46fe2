"""GPU parity of the fused LPIPS tail (spi_lpips_tap_forward/backward) against the oracle's restatement of
spi/criteria/lpips/lpips.py:50-71 + utils.normalize_activation, on the five VGG16 tap shapes (value and gradient)."""
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu
TAPS = [(64, 32), (128, 16), (256, 8), (512, 4), (512, 2)]          # (channels, spatial size): VGG16 tap shapes at reduced resolution


def _oracle_tail(xs, ys_norm, lins, n):
    tot = 0.0
    for fx, fy, w in zip(xs, ys_norm, lins):
        xn = fx / (torch.sqrt(torch.sum(fx ** 2, dim=1, keepdim=True)) + 1e-10)
        d = (xn - fy) ** 2
        tot = tot + (d * w.view(1, -1, 1, 1)).sum(1, keepdim=True).mean((2, 3), True).sum()
    return tot / n


@pytest.mark.parametrize('n,ny', [(1, 1), (4, 4), (4, 1)])
def test_lpips_tail_value_and_gradient(lib, n, ny):
    from spi_b200.criteria.lpips.lpips import _LpipsTail
    gen = torch.Generator().manual_seed(10 * n + ny)
    xs = [torch.relu(torch.randn(n, c, s, s, generator=gen)) for c, s in TAPS]
    ys = [torch.relu(torch.randn(ny, c, s, s, generator=gen)) for c, s in TAPS]
    ys = [y / (torch.sqrt(torch.sum(y ** 2, dim=1, keepdim=True)) + 1e-10) for y in ys]
    lins = [torch.rand(c, generator=gen) for c, _ in TAPS]
    xo = [x.double().requires_grad_(True) for x in xs]
    ref = _oracle_tail(xo, [y.double() for y in ys], [w.double() for w in lins], n)
    ref.backward()
    xg = [x.cuda().contiguous(memory_format=torch.channels_last).requires_grad_(True) for x in xs]
    out = _LpipsTail.apply(len(TAPS), *xg, *[y.cuda().contiguous(memory_format=torch.channels_last) for y in ys], *[w.cuda() for w in lins]) / n
    out.backward()
    assert abs(float(out) - float(ref)) <= 2e-6 * abs(float(ref))
    for g, o in zip(xg, xo):
        assert rel_l2(g.grad, o.grad) < 1e-5
