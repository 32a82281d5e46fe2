"""The operator-plugin boundary: `spi_b200.integration.custom_ops_shim.get_plugin` must serve the reference's own Python
op modules.  The reference tree cannot travel to the GPU box, so the call PROTOCOL of its autograd classes is replayed here verbatim --
forward, first-order backward = the same plugin function with swapped arguments (bias_act.py:128-209, upfirdn2d.py:219-275,
filtered_lrelu.py:161-274) -- and every result is checked against the CPU oracle of the op."""
import pytest
import torch

from conftest import rel_l2
from oracle import ops as OO

pytestmark = pytest.mark.gpu


def test_bias_act_plugin_protocol(lib):
    from spi_b200.integration.custom_ops_shim import get_plugin
    P = get_plugin('bias_act_plugin', sources=['bias_act.cpp', 'bias_act.cu'], headers=['bias_act.h'], source_dir='.', extra_cuda_cflags=['--use_fast_math'])
    E = torch.empty([0], device='cuda')
    gen = torch.Generator().manual_seed(0)
    x = torch.randn(3, 8, 17, 5, generator=gen)
    b = torch.randn(8, generator=gen)
    dy = torch.randn(3, 8, 17, 5, generator=gen)
    # lrelu (cuda_idx 3, ref 'y'): forward (bias_act.py:138), backward grad=1 with yref (bias_act.py:173)
    y = P.bias_act(x.cuda(), b.cuda(), E, E, E, 0, 1, 3, 0.2, 2 ** 0.5, 1.5)
    xo = x.clone().requires_grad_(True)
    yo = OO.bias_act(xo, b, dim=1, act='lrelu', gain=2 ** 0.5, clamp=1.5)
    yo.backward(dy)
    assert rel_l2(y, yo) < 1e-6
    dx = P.bias_act(dy.cuda(), b.cuda(), E, y, E, 1, 1, 3, 0.2, 2 ** 0.5, 1.5)
    assert rel_l2(dx, xo.grad) < 1e-6
    # swish (cuda_idx 9, ref 'x'): backward reads xref instead
    y2 = P.bias_act(x.cuda(), b.cuda(), E, E, E, 0, 1, 9, 0.0, 2 ** 0.5, -1.0)
    xo2 = x.clone().requires_grad_(True)
    yo2 = OO.bias_act(xo2, b, dim=1, act='swish')
    yo2.backward(dy)
    dx2 = P.bias_act(dy.cuda(), b.cuda(), x.cuda(), E, E, 1, 1, 9, 0.0, 2 ** 0.5, -1.0)
    assert rel_l2(y2, yo2) < 1e-6 and rel_l2(dx2, xo2.grad) < 1e-5
    with pytest.raises(RuntimeError):                          # TORCH_CHECK -> RuntimeError (bias_act.cpp:45)
        P.bias_act(x.cuda(), torch.randn(7).cuda(), E, E, E, 0, 1, 3, 0.2, 1.0, -1.0)


def test_upfirdn2d_plugin_protocol(lib):
    from spi_b200.integration.custom_ops_shim import get_plugin
    P = get_plugin('upfirdn2d_plugin', sources=['upfirdn2d.cpp', 'upfirdn2d.cu'], headers=['upfirdn2d.h'])
    gen = torch.Generator().manual_seed(1)
    x = torch.randn(2, 6, 19, 23, generator=gen)
    f = OO.setup_filter([1, 3, 3, 1])
    # upsample2d: up 2, pad [2,1,2,1], gain 4 (upfirdn2d.py:344-350); backward = the same op with up <-> down, flipped filter (:258-263)
    up, pad, gain = 2, (2, 1, 2, 1), 4.0
    y = P.upfirdn2d(x.cuda(), f.cuda(), up, up, 1, 1, *pad, False, gain)
    xo = x.clone().requires_grad_(True)
    yo = OO.upfirdn2d(xo, f, up=up, padding=list(pad), gain=gain)
    dy = torch.randn(yo.shape, generator=gen)
    yo.backward(dy)
    assert y.shape == yo.shape and rel_l2(y, yo) < 1e-6
    fw = fh = 4
    _, _, ih, iw = x.shape
    _, _, oh, ow = y.shape
    p = [fw - pad[0] - 1, iw * up - ow + pad[0] - up + 1, fh - pad[2] - 1, ih * up - oh + pad[2] - up + 1]
    dx = P.upfirdn2d(dy.cuda(), f.cuda(), 1, 1, up, up, *p, True, gain)
    assert rel_l2(dx, xo.grad) < 1e-6


def test_filtered_lrelu_plugin_protocol(lib):
    from spi_b200.integration.custom_ops_shim import get_plugin
    P = get_plugin('filtered_lrelu_plugin')
    gen = torch.Generator().manual_seed(2)
    x = torch.randn(2, 4, 16, 16, generator=gen)
    b = torch.randn(4, generator=gen)
    fu = OO.setup_filter([1, 3, 3, 1]) if hasattr(OO, 'setup_filter') else None
    fd = fu
    E = torch.empty([0], device='cuda')
    y, so, rc = P.filtered_lrelu(x.cuda(), fu.cuda(), fd.cuda(), b.cuda(), E, 2, 2, 3, 2, 3, 2, 0, 0, 2 ** 0.5, 0.2, 0.8, False, True)
    assert rc in (0, -1)
    if rc == 0:
        yo = OO.filtered_lrelu(x, fu=fu, fd=fd, b=b, up=2, down=2, padding=[3, 2, 3, 2], gain=2 ** 0.5, slope=0.2, clamp=0.8)
        assert y.shape == yo.shape and rel_l2(y, yo) < 1e-5
        assert so.dtype == torch.uint8 and so.ndim == 4
    so2 = P.filtered_lrelu_act_(x.cuda().clone(), E, 0, 0, 1.0, 0.2, -1.0, True)
    assert so2.dtype == torch.uint8
    with pytest.raises(RuntimeError):
        from spi_b200.integration.custom_ops_shim import get_plugin as gp
        gp('no_such_plugin')
