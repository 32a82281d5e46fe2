"""GPU parity: the L1 operators through the C ABI vs the CPU oracle (and the reference goldens)."""
import math

import numpy as np
import pytest
import torch

from conftest import rel_l2
from oracle import ops as O

pytestmark = pytest.mark.gpu
TOL = 2e-6     # fp32 elementwise / short FIR sums


def T(a):
    return torch.from_numpy(np.asarray(a))


@pytest.fixture(scope='module')
def P():
    from spi_b200.torch_utils.ops import bias_act, filtered_lrelu, upfirdn2d
    import types
    return types.SimpleNamespace(bias_act=bias_act, upfirdn2d=upfirdn2d, filtered_lrelu=filtered_lrelu)


def test_bias_act_forward_golden(P, golden, lib):
    g = golden('ops')
    x, b = T(g['ba_x']).cuda(), T(g['ba_b']).cuda()
    for act in O.ACTS:
        assert rel_l2(P.bias_act.bias_act(x, b, act=act), g[f'ba_{act}_d']) < TOL, act
        assert rel_l2(P.bias_act.bias_act(x, b, act=act, gain=0.7, clamp=0.9, alpha=0.3), g[f'ba_{act}_g']) < TOL, act
        xc = x.contiguous(memory_format=torch.channels_last)
        y = P.bias_act.bias_act(xc, b, act=act)
        assert y.is_contiguous(memory_format=torch.channels_last)
        assert rel_l2(y, g[f'ba_{act}_d']) < TOL, act


def test_bias_act_gradients_golden(P, golden):
    g = golden('ops')
    for act, kw in (('lrelu', dict(gain=math.sqrt(2), clamp=256.)), ('linear', dict(clamp=1.0)), ('lrelu', dict(clamp=0.5))):
        key = f"ba_grad_{act}_{kw.get('clamp')}"
        x = T(g['ba_x']).cuda().requires_grad_(True)
        b = T(g['ba_b']).cuda().requires_grad_(True)
        P.bias_act.bias_act(x, b, act=act, **kw).backward(T(g[key + '_dy']).cuda())
        assert rel_l2(x.grad, g[key + '_dx']) < TOL and rel_l2(b.grad, g[key + '_db']) < 1e-5


@pytest.mark.parametrize('act', list(O.ACTS))
@pytest.mark.parametrize('shape,dim', [((3, 5, 7, 9), 1), ((4, 33), 1), ((2, 8, 16, 16), 1), ((1000003,), 0), ((0, 4), 1)])
def test_bias_act_vs_oracle_with_autograd(P, act, shape, dim):
    gen = torch.Generator().manual_seed(hash((act, shape)) % 1000)
    x = torch.randn(*shape, generator=gen) * 2
    b = torch.randn(shape[dim], generator=gen) if len(shape) > 1 else None
    xo = x.clone().requires_grad_(True)
    yo = O.bias_act(xo, b, dim=dim, act=act, clamp=1.5)
    xg = x.cuda().requires_grad_(True)
    yg = P.bias_act.bias_act(xg, b.cuda() if b is not None else None, dim=dim, act=act, clamp=1.5)
    assert yg.shape == yo.shape
    if x.numel() == 0:
        return
    assert rel_l2(yg, yo) < 1e-5
    dy = torch.randn(*shape, generator=gen)
    yo.backward(dy)
    yg.backward(dy.cuda())
    # the plugin decides the branch from y (= act(x)*gain), torch from x: they differ only where act(x) rounds to 0
    bad = ((xg.grad.cpu() - xo.grad).abs() > 1e-5 * (1 + xo.grad.abs())).float().mean().item()
    assert bad < 1e-5


def test_bias_act_second_order(P):
    x = torch.randn(2, 4, 5, 5, dtype=torch.float32)
    for act in ('tanh', 'sigmoid', 'softplus', 'swish', 'elu', 'selu'):
        xo = x.clone().requires_grad_(True)
        go, = torch.autograd.grad(O.bias_act(xo, act=act).sum(), xo, create_graph=True)
        (go ** 2).sum().backward()
        xg = x.cuda().requires_grad_(True)
        gg, = torch.autograd.grad(P.bias_act.bias_act(xg, act=act).sum(), xg, create_graph=True)
        (gg ** 2).sum().backward()
        assert rel_l2(xg.grad, xo.grad) < 1e-4, act


def test_bias_act_dtypes_and_errors(P):
    x = torch.randn(2, 6, 4, 4)
    b = torch.randn(6)
    ref = O.bias_act(x.double(), b.double(), act='lrelu')
    assert rel_l2(P.bias_act.bias_act(x.double().cuda(), b.double().cuda(), act='lrelu'), ref) < 1e-7      # alpha/gain cross the ABI as float32, as in the plugin
    assert rel_l2(P.bias_act.bias_act(x.half().cuda(), b.half().cuda(), act='lrelu').float(), O.bias_act(x.half().float(), b.half().float(), act='lrelu')) < 2e-3
    with pytest.raises(RuntimeError):
        P.bias_act.bias_act(x.cuda(), torch.randn(5).cuda())


UP_CASES = {
    'blur_after_convT': dict(up=1, down=1, padding=[1, 1, 1, 1], gain=4.0, flip_filter=False),
    'upsample2d': dict(up=2, down=1, padding=[2, 1, 2, 1], gain=4.0, flip_filter=False),
    'upsample2d_bwd': dict(up=1, down=2, padding=[1, 2, 1, 2], gain=4.0, flip_filter=True),
    'generic_a': dict(up=[2, 1], down=[1, 2], padding=[3, 0, 1, 2], gain=1.3, flip_filter=True),
    'crop': dict(up=1, down=1, padding=[-1, 2, 0, -2], gain=1.0, flip_filter=False),
    'down3': dict(up=1, down=3, padding=[2, 2, 2, 2], gain=1.0, flip_filter=False),
}


def test_upfirdn2d_golden(P, golden):
    g = golden('ops')
    x, f = T(g['up_x']).cuda(), T(g['f4']).cuda()
    for name, kw in UP_CASES.items():
        assert rel_l2(P.upfirdn2d.upfirdn2d(x, f, **kw), g['up_' + name]) < TOL, name
        xc = x.contiguous(memory_format=torch.channels_last)
        assert rel_l2(P.upfirdn2d.upfirdn2d(xc, f, **kw), g['up_' + name]) < TOL, name + '/cl'
    xs, f12 = T(g['up_xs']).cuda(), T(g['f12']).cuda()
    assert rel_l2(P.upfirdn2d.upfirdn2d(xs, f12, up=2, padding=[5, 6, 5, 6], gain=4.0), g['up_separable12']) < 1e-5


@pytest.mark.parametrize('case', list(UP_CASES))
@pytest.mark.parametrize('cl', [False, True])
def test_upfirdn2d_vs_oracle_with_autograd(P, case, cl):
    kw = UP_CASES[case]
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(2, 8, 21, 19, generator=gen)
    f = O.setup_filter([1, 3, 3, 1])
    xo = x.clone().requires_grad_(True)
    yo = O.upfirdn2d(xo, f, **kw)
    xg = x.cuda()
    if cl:
        xg = xg.contiguous(memory_format=torch.channels_last)
    xg.requires_grad_(True)
    yg = P.upfirdn2d.upfirdn2d(xg, f.cuda(), **kw)
    assert yg.shape == yo.shape and rel_l2(yg, yo) < TOL
    dy = torch.randn(*yo.shape, generator=gen)
    yo.backward(dy)
    yg.backward(dy.cuda())
    assert rel_l2(xg.grad, xo.grad) < TOL


def test_upsample2d_full_size_property(P):
    """512^2-scale property check: upsample2d of a constant image is constant (DC gain 1), and linearity."""
    f = O.setup_filter([1, 3, 3, 1]).cuda()
    x = torch.full((1, 96, 128, 128), 0.37, device='cuda').contiguous(memory_format=torch.channels_last)
    y = P.upfirdn2d.upsample2d(x, f)
    assert y.shape == (1, 96, 256, 256)
    assert (y[:, :, 4:-4, 4:-4] - 0.37).abs().max() < 1e-6
    a, b = torch.randn(1, 96, 128, 128, device='cuda'), torch.randn(1, 96, 128, 128, device='cuda')
    lhs = P.upfirdn2d.upsample2d(a + 2 * b, f)
    rhs = P.upfirdn2d.upsample2d(a, f) + 2 * P.upfirdn2d.upsample2d(b, f)
    assert rel_l2(lhs, rhs) < 1e-6


FL_CASES = {
    'u2d2': dict(up=2, down=2, padding=[9, 10, 9, 10], gain=math.sqrt(2), slope=0.2, clamp=256., flip_filter=False),
    'u2d1': dict(up=2, down=1, padding=[5, 6, 5, 6], gain=math.sqrt(2), slope=0.2, clamp=None, flip_filter=False),
    'u1d2': dict(up=1, down=2, padding=[5, 5, 5, 5], gain=1.1, slope=0.1, clamp=0.8, flip_filter=True),
    'u1d1': dict(up=1, down=1, padding=[0, 0, 0, 0], gain=math.sqrt(2), slope=0.2, clamp=None, flip_filter=False),
}


@pytest.mark.parametrize('name', list(FL_CASES))
def test_filtered_lrelu_golden_and_grad(P, golden, name):
    g = golden('ops')
    kw = FL_CASES[name]
    xs, f12, fd12, b = T(g['up_xs']), T(g['f12']), T(g['fd12']), T(g['fl_b'])
    fu = f12 if kw['up'] > 1 else (None if name == 'u1d1' else f12)
    fd = fd12 if kw['down'] > 1 else None
    xg = xs.cuda().requires_grad_(True)
    bg = b.cuda().requires_grad_(True)
    yg = P.filtered_lrelu.filtered_lrelu(xg, fu=fu.cuda() if fu is not None else None, fd=fd.cuda() if fd is not None else None, b=bg, **kw)
    assert rel_l2(yg, g['fl_' + name]) < 1e-5
    xo = xs.clone().requires_grad_(True)
    bo = b.clone().requires_grad_(True)
    yo = O.filtered_lrelu(xo, fu=fu, fd=fd, b=bo, **kw)
    dy = torch.randn(*yo.shape, generator=torch.Generator().manual_seed(5))
    yo.backward(dy)
    yg.backward(dy.cuda())
    assert rel_l2(xg.grad, xo.grad) < 1e-5 and rel_l2(bg.grad, bo.grad) < 1e-5


def test_filtered_lrelu_larger_and_channels_last(P):
    gen = torch.Generator().manual_seed(9)
    x = torch.randn(2, 5, 37, 41, generator=gen)
    b = torch.randn(5, generator=gen)
    fu = O.setup_filter([1, 3, 3, 1]) * 1.0
    for cl in (False, True):
        xg = x.cuda().contiguous(memory_format=torch.channels_last) if cl else x.cuda()
        y = P.filtered_lrelu.filtered_lrelu(xg, fu=fu.cuda(), fd=fu.cuda(), b=b.cuda(), up=2, down=2, padding=[3, 3, 3, 3], clamp=1.0)
        assert rel_l2(y, O.filtered_lrelu(x, fu=fu, fd=fu, b=b, up=2, down=2, padding=[3, 3, 3, 3], clamp=1.0)) < 1e-5


def test_modulate_weights_vs_oracle():
    from spi_b200.ops.modulate import modulate_weights
    gen = torch.Generator().manual_seed(4)
    for (n, o, i, k, demod) in ((1, 16, 8, 3, True), (4, 33, 20, 3, True), (2, 7, 64, 1, False), (3, 96, 512, 1, False)):
        W = torch.randn(o, i, k, k, generator=gen)
        s = torch.randn(n, i, generator=gen) + 1
        g = torch.randn(n, o, i, k, k, generator=gen)
        Wo, so = W.clone().requires_grad_(True), s.clone().requires_grad_(True)
        w = Wo[None] * so.reshape(n, 1, i, 1, 1)
        if demod:
            w = w * (w.square().sum(dim=[2, 3, 4]) + 1e-8).rsqrt().reshape(n, o, 1, 1, 1)
        (w * g).sum().backward()
        Wg, sg = W.cuda().requires_grad_(True), s.cuda().requires_grad_(True)
        out = modulate_weights(Wg, sg, demod, layout=["oihw","ohwi","ihwo"][(n + o) % 3])
        assert rel_l2(out, w) < 1e-5
        (out * g.cuda()).sum().backward()
        assert rel_l2(Wg.grad, Wo.grad) < 1e-4 and rel_l2(sg.grad, so.grad) < 1e-4


@pytest.mark.parametrize('n,o,i,k,demod,layout,flip', [
    (1, 512, 512, 3, True, 'ohwi', False),      # backbone 3x3 layers, shared styles
    (2, 512, 512, 3, True, 'ihwo', True),       # up-sampling layers: transposed-conv layout with reversed taps
    (1, 128, 256, 3, True, 'ihwo', True),
    (4, 40, 24, 3, True, 'ohwi', True),
    (3, 40, 24, 3, True, 'oihw', True),
    (2, 33, 70, 3, False, 'ihwo', False),       # ragged tiles of the transposing kernel
    (1, 96, 128, 1, False, 'ohwi', False),      # toRGB: 1x1, no demodulation
    (2, 3, 4200, 3, True, 'ohwi', False),       # row too long for shared memory: the unstaged kernels
])
def test_modulate_weights_layer_shapes(n, o, i, k, demod, layout, flip):
    """Row-staged modulation kernels on the generator's real layer shapes, every layout, with and without reversed taps,
    against the plain fp32 formula of networks_stylegan2.py:58-68 evaluated by torch on the same device."""
    from spi_b200.ops.modulate import modulate_weights
    gen = torch.Generator().manual_seed(n * 1000 + o + i + k)
    W = torch.randn(o, i, k, k, generator=gen).cuda()
    s = (torch.randn(n, i, generator=gen) + 1).cuda()
    g = torch.randn(n, o, i, k, k, generator=gen).cuda()
    Wr, sr = W.clone().requires_grad_(True), s.clone().requires_grad_(True)
    w = Wr[None] * sr.reshape(n, 1, i, 1, 1)
    if demod:
        w = w * (w.double().square().sum(dim=[2, 3, 4]) + 1e-8).rsqrt().float().reshape(n, o, 1, 1, 1)
    if flip:
        w = w.flip([3, 4])
    (w * g).sum().backward()
    Wg, sg = W.clone().requires_grad_(True), s.clone().requires_grad_(True)
    out = modulate_weights(Wg, sg, demod, layout=layout, flip=flip)
    assert out.shape == w.shape and rel_l2(out, w) < 2e-6
    want = {'oihw': (0, 1, 2, 3, 4), 'ohwi': (0, 1, 3, 4, 2), 'ihwo': (0, 2, 3, 4, 1)}[layout]
    assert out.permute(*want).is_contiguous()
    (out * g).sum().backward()
    assert rel_l2(Wg.grad, Wr.grad) < 2e-5 and rel_l2(sg.grad, sr.grad) < 2e-5
    Wn = W.clone()                                           # stage 1: the weights are frozen, only the styles get a gradient
    sn = s.clone().requires_grad_(True)
    (modulate_weights(Wn, sn, demod, layout=layout, flip=flip) * g).sum().backward()
    assert rel_l2(sn.grad, sr.grad) < 2e-5


@pytest.mark.parametrize('shape,act,clamp,with_noise', [
    ((2, 16, 19, 19), 'lrelu', None, True),        # small layer: 4x2-patch kernel
    ((1, 8, 35, 37), 'linear', 0.7, False),        # ragged, no noise (the SR layers run with noise_mode='none')
    ((2, 128, 131, 131), 'lrelu', 256.0, True),    # 8-row strips
    ((4, 128, 259, 259), 'lrelu', None, False),    # 16-row strips
    ((1, 12, 67, 70), 'lrelu', 1.5, True),         # strips with ragged right / bottom edges
])
def test_blur_bias_act_noise_equals_the_two_passes(P, shape, act, clamp, with_noise):
    """The fused tail of an up-sampling layer (4x4 FIR + noise + bias + activation, `spi_blur4_bias_act_noise`) must reproduce
    upfirdn2d followed by bias_act_noise bit for bit in the forward pass, and their gradients."""
    n, c, ih, iw = shape
    gen = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(*shape, generator=gen).cuda().contiguous(memory_format=torch.channels_last)
    f = P.upfirdn2d.setup_filter([1, 3, 3, 1]).cuda()
    pad = [1, 1, 1, 1]
    oh, ow = ih - 1, iw - 1
    b = torch.randn(c, generator=gen).cuda()
    nc = torch.randn(oh, ow, generator=gen).cuda() if with_noise else None
    st = torch.tensor(0.37).cuda() if with_noise else None
    dy = torch.randn(n, c, oh, ow, generator=gen).cuda()
    outs = []
    for fused in (True, False):
        leaves = [t.clone().requires_grad_(True) if t is not None else None for t in (x, b, nc, st)]
        xg, bg, ng, sg = leaves
        if fused:
            y = P.bias_act.blur_bias_act_noise(xg, f, bg, ng, sg, padding=pad, fir_gain=4, act=act, gain=1.3, clamp=clamp)
        else:
            t = P.upfirdn2d.upfirdn2d(xg, f, padding=pad, gain=4)
            y = (P.bias_act.bias_act_noise(t, bg, ng, sg, act=act, gain=1.3, clamp=clamp) if with_noise
                 else P.bias_act.bias_act(t, bg, act=act, gain=1.3, clamp=clamp))
        y.backward(dy)
        outs.append((y.detach(), [l.grad for l in leaves if l is not None]))
    assert outs[0][0].shape == (n, c, oh, ow) and torch.equal(outs[0][0], outs[1][0])
    for a, r in zip(outs[0][1], outs[1][1]):
        assert rel_l2(a, r) < 2e-5          # same kernels on identical inputs; the scalar / per-channel sums use float atomics
    ref = O.bias_act(O.upfirdn2d(x.cpu(), f.cpu(), padding=pad, gain=4) + (nc.cpu() * st.cpu() if with_noise else 0), b.cpu(), act=act,
                     gain=1.3, clamp=clamp)
    assert rel_l2(outs[0][0], ref) < 2e-6


def test_bias_act_noise_vs_oracle(P):
    gen = torch.Generator().manual_seed(6)
    for cl in (False, True):
        x = torch.randn(2, 12, 9, 9, generator=gen)
        b, nc, st = torch.randn(12, generator=gen), torch.randn(9, 9, generator=gen), torch.tensor(0.37)
        xo, bo, no, so = (t.clone().requires_grad_(True) for t in (x, b, nc, st))
        yo = O.bias_act(xo + no * so, bo, act='lrelu', clamp=1.2)
        dy = torch.randn(*yo.shape, generator=gen)
        yo.backward(dy)
        xg = x.cuda().contiguous(memory_format=torch.channels_last) if cl else x.cuda()
        xg.requires_grad_(True)
        bg, ng, sg = (t.cuda().requires_grad_(True) for t in (b, nc, st))
        yg = P.bias_act.bias_act_noise(xg, bg, ng, sg, act='lrelu', clamp=1.2)
        assert rel_l2(yg, yo) < 1e-6
        yg.backward(dy.cuda())
        for a, r in ((xg.grad, xo.grad), (bg.grad, bo.grad), (ng.grad, no.grad), (sg.grad, so.grad)):
            assert rel_l2(a, r) < 1e-5


@pytest.mark.parametrize('shape', [(2, 3, 16, 24), (1, 3, 64, 64), (4, 6, 8, 10), (1, 5, 4, 4)])
def test_bias_gradient_small_channel_counts_channels_last(P, shape):
    """toRGB outputs are channels-last with C = 3: the bias gradient goes through the small-C reduction kernel."""
    gen = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(*shape, generator=gen)
    b = torch.randn(shape[1], generator=gen)
    dy = torch.randn(*shape, generator=gen)
    xo, bo = x.clone().requires_grad_(True), b.clone().requires_grad_(True)
    O.bias_act(xo, bo, act='linear', clamp=1.0).backward(dy)
    xg = x.cuda().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    bg = b.cuda().requires_grad_(True)
    P.bias_act.bias_act(xg, bg, act='linear', clamp=1.0).backward(dy.cuda().contiguous(memory_format=torch.channels_last))
    assert rel_l2(xg.grad, xo.grad) < TOL and rel_l2(bg.grad, bo.grad) < 1e-5


@pytest.mark.parametrize('shape,act,clamp,with_noise', [
    ((1, 128, 64, 64), 'lrelu', 256.0, True),      # generator layer: bias + noise_strength gradients
    ((2, 96, 33, 31), 'linear', 0.9, False),       # 24 vectors per row (not a power of two), clamp masks part of the gradient
    ((1, 512, 8, 8), 'lrelu', None, True),         # small map, many channels
    ((4, 64, 40, 40), 'relu', None, False),        # relu (VGG layers with a trainable bias)
    ((1, 1024, 4, 4), 'lrelu', 1.0, True),         # widest row the kernel takes
])
def test_activation_gradient_and_its_reductions_in_one_pass(P, shape, act, clamp, with_noise):
    """While the generator is tuned noise_const is a buffer: the backward pass of a layer epilogue wants dx, db and d noise_strength only, which
    `spi_bias_act_grad_reduce` produces in ONE pass over (dy, y) -- checked against the oracle's autograd (eg3d bias_act.py:54-88 semantics)."""
    n, c, h, w = shape
    gen = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(*shape, generator=gen)
    b = torch.randn(c, generator=gen)
    nc = torch.randn(h, w, generator=gen) if with_noise else None
    st = torch.tensor(0.41)
    dy = torch.randn(*shape, generator=gen)
    xo, bo, so = x.clone().requires_grad_(True), b.clone().requires_grad_(True), st.clone().requires_grad_(True)
    O.bias_act(xo + (nc * so if with_noise else 0), bo, act=act, gain=1.3, clamp=clamp).backward(dy)
    xg = x.cuda().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    bg, sg = b.cuda().requires_grad_(True), st.cuda().requires_grad_(True)
    from spi_b200 import _lib
    before = _lib.launch_count()
    if with_noise:
        yg = P.bias_act.bias_act_noise(xg, bg, nc.cuda(), sg, act=act, gain=1.3, clamp=clamp)       # noise_const: no gradient wanted
    else:
        yg = P.bias_act.bias_act(xg, bg, act=act, gain=1.3, clamp=clamp)
    fwd = _lib.launch_count() - before
    yg.backward(dy.cuda().contiguous(memory_format=torch.channels_last))
    assert _lib.launch_count() - before - fwd == 1         # the whole epilogue backward is one kernel of this library
    assert rel_l2(xg.grad, xo.grad) < 1e-6 and rel_l2(bg.grad, bo.grad) < 1e-5
    if with_noise:
        assert abs(float(sg.grad) - float(so.grad)) < 1e-4 * float((dy * nc).abs().sum())
