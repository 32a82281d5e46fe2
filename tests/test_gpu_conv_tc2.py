"""GPU parity of the default conv engine (spi_b200/csrc/conv_tc2.cu, conv_wgrad_tc2.cu) against fp64 evaluations of the same
convolutions (oracle.ops has no conv of its own: the reference calls torch's, conv2d_resample.py:30-43).  TF32 operands, fp32
accumulation: tolerance 1e-3 rel-L2 (north_star); measured 3e-4, the same as cuDNN's TF32 kernels."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-3
CL = torch.channels_last


def _mk(n, ci, co, h, wd, k, per_sample, seed):
    gen = torch.Generator().manual_seed(seed)
    x = torch.randn(n, ci, h, wd, generator=gen).cuda().contiguous(memory_format=CL)
    g = n if per_sample else 1
    w = (torch.randn(g, co, ci, k, k, generator=gen) / (ci * k * k) ** 0.5).cuda()
    return gen, x, w


@pytest.mark.parametrize('n,ci,co,h,wd,k,per_sample', [
    (1, 32, 32, 16, 16, 3, False),       # one tile
    (2, 64, 128, 40, 56, 3, True),       # ragged spatial size (TMA clips the border tiles), per-sample weights
    (1, 128, 96, 64, 64, 1, True),       # 1x1 (toRGB shape), Cout = 96
    (3, 96, 320, 24, 16, 3, False),      # three Cout tiles, the last partial
    (1, 512, 512, 4, 4, 3, True),        # 4x4 map: box larger than the image, split over Cin with reduce-add stores
    (1, 512, 512, 32, 32, 3, False),     # split-K path
    (1, 128, 128, 128, 128, 3, False),   # full-size tiles, two M tiles per CTA
])
def test_stride1_forward_input_grad_weight_grad(lib, n, ci, co, h, wd, k, per_sample):
    """conv2d_per_sample through autograd: y, dL/dx and dL/dw all come from this library's kernels (no cuDNN call on the path)."""
    from spi_b200.ops import conv as E
    assert E.ENGINE == 'tc2' and E.WGRAD_ENGINE == 'tc2'
    gen, x, w = _mk(n, ci, co, h, wd, k, per_sample, n * 1000 + ci + co + h)
    xg, wg = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    assert E.tc2_form(xg, wg, 1, k // 2, False) == 's1'
    y = E.conv2d_per_sample(xg, wg, padding=k // 2)
    gy = torch.randn(y.shape, generator=gen).cuda().contiguous(memory_format=CL)
    y.backward(gy)
    assert lib.spi_tc_error() == 0
    xr, wr = x.double().requires_grad_(True), w.double().requires_grad_(True)
    yr = torch.cat([F.conv2d(xr[i:i + 1], wr[i if per_sample else 0], padding=k // 2) for i in range(n)])
    yr.backward(gy.double())
    assert y.shape == yr.shape and y.is_contiguous(memory_format=CL)
    assert rel_l2(y, yr) < TOL and rel_l2(xg.grad, xr.grad) < TOL and rel_l2(wg.grad, wr.grad) < TOL


@pytest.mark.parametrize('n,ci,co,h,wd,per_sample', [
    (1, 32, 32, 16, 16, False),
    (2, 64, 128, 20, 28, True),
    (1, 512, 512, 4, 4, True),
    (1, 256, 128, 64, 64, False),
    (2, 32, 256, 33, 17, True),
])
def test_stride2_transposed_forward_and_grads(lib, n, ci, co, h, wd, per_sample):
    """The up-sampling layers' conv_transpose2d(stride 2) (conv2d_resample.py:117): forward = four output-parity phases, dL/dx = the
    stride-2 correlation over the four input-parity views, dL/dw = the weight-gradient kernel with the operand roles swapped."""
    from spi_b200.ops import conv as E
    gen, x, w = _mk(n, ci, co, h, wd, 3, per_sample, 7 * n + ci + co + h)
    xg, wg = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    assert E.tc2_form(xg, wg, 2, 0, True) == 't2'
    y = E.conv2d_per_sample(xg, wg, stride=2, padding=0, transpose=True)
    gy = torch.randn(y.shape, generator=gen).cuda().contiguous(memory_format=CL)
    y.backward(gy)
    assert lib.spi_tc_error() == 0
    xr, wr = x.double().requires_grad_(True), w.double().requires_grad_(True)
    yr = torch.cat([F.conv_transpose2d(xr[i:i + 1], wr[i if per_sample else 0].transpose(0, 1), stride=2) for i in range(n)])
    yr.backward(gy.double())
    assert y.shape == yr.shape == (n, co, 2 * h + 1, 2 * wd + 1)
    assert rel_l2(y, yr) < TOL and rel_l2(xg.grad, xr.grad) < TOL and rel_l2(wg.grad, wr.grad) < TOL


@pytest.mark.parametrize('act,h', [('lrelu', 128), ('linear', 128), ('relu', 64)])
def test_fused_layer_epilogue_and_its_gradients(lib, act, h):
    """conv + noise*strength + bias -> act*gain -> clamp in the convolution's accumulator read-out (SynthesisLayer tail,
    networks_stylegan2.py:320-329): output and every gradient (x, w, bias, noise_const, noise_strength) against fp64."""
    from spi_b200.ops import conv as E
    n, ci, co = 2, 64, 128
    gen, x, w = _mk(n, ci, co, h, h, 3, False, 99)
    b = torch.randn(co, generator=gen).cuda()
    nz = torch.randn(h, h, generator=gen).cuda()
    st = torch.tensor(0.7).cuda()
    gain, clamp = 2 ** 0.5, (None if act == 'linear' else 2.5)
    leaves = [t.clone().requires_grad_(True) for t in (x, w, b, nz, st)]
    assert E.conv_bias_act_fusable(leaves[0], leaves[1], act) == (h >= 128)
    y = E.conv2d_bias_act(leaves[0], leaves[1], leaves[2], leaves[3], leaves[4], act=act, gain=gain, clamp=clamp)
    gy = torch.randn(y.shape, generator=gen).cuda().contiguous(memory_format=CL)
    y.backward(gy)
    assert lib.spi_tc_error() == 0
    ref = [t.double().requires_grad_(True) for t in (x, w, b, nz, st)]
    pre = F.conv2d(ref[0], ref[1][0], padding=1) + ref[3] * ref[4] + ref[2].view(1, -1, 1, 1)
    a = {'lrelu': lambda t: F.leaky_relu(t, 0.2), 'linear': lambda t: t, 'relu': torch.relu}[act](pre)
    yr = a * gain
    if clamp is not None:
        yr = yr.clamp(-clamp, clamp)
    yr.backward(gy.double())
    assert rel_l2(y, yr) < TOL
    errs = [rel_l2(m.grad, r.grad) for m, r in zip(leaves, ref)]
    print(act, h, 'grad rel-L2 (x, w, b, noise, strength):', errs)
    # linear, unclamped: the gradients are linear in the operands -> TF32 accuracy.  lrelu / relu / clamp: outputs within TF32 noise of a
    # kink take the other branch's derivative than the fp64 evaluation (a fraction ~1e-4 of the elements -> ~1e-2 rel-L2)
    assert max(errs) < (2e-3 if act == 'linear' else 3e-2), errs


def test_vgg_stem_runs_on_the_engine(lib):
    """The 3-channel VGG stem is zero-padded to 32 input channels so that it runs on the same implicit-GEMM kernel (ops/conv.py::vgg_conv)."""
    from spi_b200.ops import conv as E
    conv = torch.nn.Conv2d(3, 64, 3, padding=1).cuda()
    x = torch.randn(2, 3, 64, 64, device='cuda').requires_grad_(True)
    before = lib.spi_launch_count()
    y = E.vgg_conv(x, conv, act='relu')
    assert lib.spi_launch_count() > before
    y.sum().backward()
    xr = x.detach().double().requires_grad_(True)
    yr = torch.relu(F.conv2d(xr, conv.weight.double(), conv.bias.double(), padding=1))
    yr.sum().backward()
    # dx passes through the ReLU mask: outputs within TF32 noise of zero take the other branch than the fp64 evaluation (~5e-3 rel-L2)
    assert rel_l2(y, yr) < TOL and rel_l2(x.grad, xr.grad) < 3e-2


def test_generator_gradients_do_not_depend_on_the_engine(product_G, lib):
    """Same generator, same inputs: the tc2 engine and the cuDNN arm (SPI_CONV_ENGINE=cudnn) agree on the image and on parameter gradients
    to TF32 accuracy, including the cancellation-prone scalar d/d(noise_strength)."""
    import copy
    from oracle import generator as OG
    from oracle import weights
    from spi_b200.ops import conv as E
    keys = ['backbone.synthesis.b128.conv1.noise_strength', 'backbone.synthesis.b64.conv0.weight', 'superresolution.block1.conv1.weight',
            'backbone.synthesis.b256.torgb.weight', 'backbone.synthesis.b32.conv1.bias']
    res = {}
    old = (E.ENGINE, E.WGRAD_ENGINE)
    try:
        for eng in ('cudnn', 'tc2'):
            E.ENGINE = E.WGRAD_ENGINE = eng
            G = copy.deepcopy(product_G).requires_grad_(True)
            jit, u = OG.make_render_noise(1, 128 * 128, {**OG.RENDERING_DEFAULTS, **dict(G.rendering_kwargs)}, seed=3)
            G.renderer.inject_noise(jit.cuda(), u.cuda())
            ws = weights.w_pivot(5).cuda().requires_grad_(True)
            out = G.synthesis(ws, weights.canonical_camera(0.3).cuda(), noise_mode='const')
            gimg = torch.randn(1, 3, 512, 512, generator=torch.Generator().manual_seed(1)).cuda()
            (out['image'] * gimg).sum().backward()
            p = dict(G.named_parameters())
            res[eng] = (out['image'].detach(), ws.grad.clone(), {k: p[k].grad.clone() for k in keys})
    finally:
        E.ENGINE, E.WGRAD_ENGINE = old
    assert lib.spi_tc_error() == 0
    e_img, e_ws = rel_l2(res['tc2'][0], res['cudnn'][0]), rel_l2(res['tc2'][1], res['cudnn'][1])
    errs = {k: rel_l2(res['tc2'][2][k], res['cudnn'][2][k]) for k in keys}
    print('tc2 vs cuDNN: image', e_img, 'dws', e_ws, errs, 'noise_strength grads', float(res['tc2'][2][keys[0]]), float(res['cudnn'][2][keys[0]]))
    assert e_img < 1e-3 and e_ws < 3e-2
    assert max(v for k, v in errs.items() if 'noise_strength' not in k) < 3e-2


@pytest.mark.parametrize('n,ci,co,h,k,split', [
    (1, 128, 128, 256, 3, False), (1, 256, 256, 128, 3, False), (4, 128, 128, 128, 3, False), (1, 128, 96, 256, 1, False),
    (1, 512, 512, 32, 3, True), (1, 512, 512, 8, 3, True),
])
def test_repeatability(lib, n, ci, co, h, k, split):
    """Layers that are not split over Cin have a fixed summation order: repeated launches are bit-identical (a race between the TMA,
    MMA and epilogue roles would show up here); split layers and weight gradients combine partial sums with fp32 reduce-adds in
    arrival order and agree to fp32 rounding."""
    from spi_b200.ops import conv as E
    gen, x, w = _mk(n, ci, co, h, h, k, False, 5)
    wk = E._ohwi(w).view(1, co, k * k, ci)
    ys = [E.tc2_s1(x, wk, k, False) for _ in range(4)]
    gy = torch.randn(n, co, h, h, generator=gen).cuda().contiguous(memory_format=CL)
    gws = [E._weight_grad('s1', gy, x, w, 1, k // 2, False) for _ in range(3)]
    assert lib.spi_tc_error() == 0
    for y in ys[1:]:
        if split:
            assert rel_l2(y, ys[0]) < 1e-6
        else:
            assert torch.equal(y, ys[0])
    for g in gws[1:]:
        assert rel_l2(g, gws[0]) < 1e-6
    e = dict(b=torch.randn(co, generator=gen).cuda(), noise=torch.randn(h, h, generator=gen).cuda(), strength=torch.tensor(0.5).cuda(), act=2, gain=1.4,
             clamp=3.0)
    fs = [E.tc2_s1(x, wk, k, False, e, allow_split=False) for _ in range(3)]
    assert torch.equal(fs[0], fs[1]) and torch.equal(fs[0], fs[2])


@pytest.mark.parametrize('n,ci,h,per_sample', [(1, 128, 64, False), (2, 256, 40, True), (4, 128, 32, False)])
def test_rgb_head_1x1_conv(lib, n, ci, h, per_sample):
    """The Cout = 3 ToRGB convolutions of the super-resolution blocks (streaming kernels, spi_b200/csrc/conv_rgb.cu): y, dx, dw in fp32."""
    from spi_b200.ops import conv as E
    gen, x, w = _mk(n, ci, 3, h, h, 1, per_sample, 3 + n + ci)
    xg, wg = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    assert E.tc2_form(xg, wg, 1, 0, False) == 'rgb'
    before = lib.spi_launch_count()
    y = E.conv2d_per_sample(xg, wg, padding=0)
    gy = torch.randn(y.shape, generator=gen).cuda().contiguous(memory_format=CL)
    y.backward(gy)
    assert lib.spi_launch_count() >= before + 3
    xr, wr = x.double().requires_grad_(True), w.double().requires_grad_(True)
    yr = torch.cat([F.conv2d(xr[i:i + 1], wr[i if per_sample else 0]) for i in range(n)])
    yr.backward(gy.double())
    assert rel_l2(y, yr) < 1e-5 and rel_l2(xg.grad, xr.grad) < 1e-5 and rel_l2(wg.grad, wr.grad) < 1e-4


def test_zero_arena_outputs_equal_self_zeroed_outputs(lib):
    """Accumulate-into outputs (Cin-split convolutions, weight gradients, RGB weight gradient, style gradients of the modulation backward)
    carved from the per-iteration zero arena give the results of the calls that zero-fill their own output; the second iteration of a
    kind gets the arena (sized by the first) and the library's fill is skipped (flag), a third, larger request falls back."""
    from spi_b200.ops import conv as E
    from spi_b200.ops import zero_arena as Z
    from spi_b200.ops.modulate import modulate_weights

    def work(scale=1):
        gen = torch.Generator().manual_seed(11)
        outs = []
        for (n, ci, co, h, k, ps) in [(1, 512, 512, 8 * scale, 3, False), (2, 256, 512, 16, 3, True), (1, 128, 128, 64, 3, False)]:
            x = torch.randn(n, ci, h, h, generator=gen).cuda().contiguous(memory_format=CL).requires_grad_(True)
            w = (torch.randn(n if ps else 1, co, ci, k, k, generator=gen) / (ci * k * k) ** 0.5).cuda().requires_grad_(True)
            y = E.conv2d_per_sample(x, w, padding=k // 2)
            gy = torch.randn(y.shape, generator=gen).cuda().contiguous(memory_format=CL)
            gx, gw = torch.autograd.grad(y, [x, w], gy)
            outs += [y.detach(), gx, gw]
        # stride-2 transposed layer on a small map + modulation backward (style gradients) + RGB head
        x = torch.randn(1, 256, 16, 16, generator=gen).cuda().contiguous(memory_format=CL).requires_grad_(True)
        wt = torch.randn(128, 256, 3, 3, generator=gen).cuda().requires_grad_(True)
        s = (1 + 0.1 * torch.randn(1, 256, generator=gen)).cuda().requires_grad_(True)
        w5 = modulate_weights(wt, s, True, layout='ohwi', flip=True)
        y = E.conv2d_per_sample(x, w5, stride=2, padding=0, transpose=True)
        outs += [y.detach()] + list(torch.autograd.grad(y.square().sum(), [x, wt, s]))
        x = torch.randn(2, 128, 32, 32, generator=gen).cuda().contiguous(memory_format=CL).requires_grad_(True)
        w = torch.randn(2, 3, 128, 1, 1, generator=gen).cuda().requires_grad_(True)
        y = E.conv2d_per_sample(x, w, padding=0)
        outs += [y.detach()] + list(torch.autograd.grad(y.square().sum(), [x, w]))
        return outs

    Z.reset()
    plain = work()
    with Z.iteration('test', torch.device('cuda')):
        first = work()                     # no hint yet: every take() falls back
        assert not Z.active()
    assert Z._hint['test'] > 0
    with Z.iteration('test', torch.device('cuda')):
        assert Z.active()
        second = work()
        used = Z._state[2]
    assert used == Z._hint['test'] and used > 0
    with Z.iteration('test', torch.device('cuda')):
        third = work(scale=2)              # asks for more than the arena holds: the tail of the requests falls back
    assert lib.spi_tc_error() == 0
    # (reduce-adds and atomics arrive in a different order from run to run: fp32 rounding of sums of ~1e3 terms, not bit equality)
    for j, (a, b, c) in enumerate(zip(plain, first, second)):
        assert rel_l2(b, a) < 1e-5 and rel_l2(c, a) < 1e-5, (j, rel_l2(b, a), rel_l2(c, a))
    ref3 = work(scale=2)
    for j, (a, b) in enumerate(zip(ref3, third)):
        assert rel_l2(b, a) < 1e-5, (j, rel_l2(b, a))
    Z.reset()


@pytest.mark.parametrize('rows,cu,cv', [(1 << 16, 64, 32), (1 << 16, 36, 64), (8 * 1237, 64, 32), (4096 * 64 * 4, 36, 64), (1 << 14, 128, 96)])
def test_rows_outer_sum(lib, rows, cu, cv):
    """spi_rows_outer_sum (the decoder-gradient reduction of the renderer: out = u^T v over millions of rows, column sums of u from the same
    pass) against fp64; ragged channel counts (36) and a row count that is not a multiple of the 128-row tile included."""
    from spi_b200 import _lib
    gen = torch.Generator().manual_seed(rows % 1000 + cu)
    u = torch.randn(rows, cu, generator=gen).cuda()
    v = (torch.randn(rows, cv, generator=gen) + 0.3).cuda()
    u[:, 0] += 0.5                                         # a column with a non-zero mean: its sum does not cancel
    out = torch.full((cu, cv), float('nan'), device='cuda')
    usum = torch.full((cu,), float('nan'), device='cuda')
    _lib.check(lib.spi_rows_outer_sum(_lib.ptr(u), _lib.ptr(v), rows, cu, cv, _lib.ptr(out), _lib.ptr(usum), _lib.stream()))
    torch.cuda.synchronize()
    assert lib.spi_tc_error() == 0
    ref = u.double().t() @ v.double()
    assert rel_l2(out, ref) < TOL
    rs = u.double().sum(0)
    assert (usum.double() - rs).abs().max().item() < 1e-3 * u.double().abs().sum(0).max().item()
    out2 = torch.empty_like(out)
    _lib.check(lib.spi_rows_outer_sum(_lib.ptr(u), _lib.ptr(v), rows, cu, cv, _lib.ptr(out2), None, _lib.stream()))
    assert rel_l2(out2, out) < 1e-5


@pytest.mark.parametrize('n,ci,co,h,wd,k,epi', [
    (1, 128, 128, 256, 256, 3, False), (1, 128, 128, 256, 256, 3, True), (2, 96, 128, 130, 204, 3, True), (1, 64, 128, 256, 256, 1, False),
    (4, 32, 128, 96, 96, 3, True),
])
def test_swapped_operand_form_for_128_channel_tiles(lib, n, ci, co, h, wd, k, epi, monkeypatch):
    """Cout = 128 layers on large maps run with the operand roles swapped (weights on M, 256 pixels on N: M128 x N256 instructions, per-warp
    TMA stores of [32 px][32 ch] boxes).  Same results as the unswapped form (flag 1024) and as fp64, ragged right / bottom edges included."""
    from spi_b200.ops import conv as E
    gen, x, w = _mk(n, ci, co, h, wd, k, False, 17 + ci + h)
    wk = E._ohwi(w).view(1, co, k * k, ci)
    e = None
    if epi:
        e = dict(b=torch.randn(co, generator=gen).cuda(), noise=torch.randn(h, wd, generator=gen).cuda(), strength=torch.tensor(0.7).cuda(), act=2, slope=0.2,
                 gain=1.4, clamp=2.5)
    y_swapped = E.tc2_s1(x, wk, k, False, e, allow_split=e is None)
    monkeypatch.setattr(E, 'TC2_FLAGS', E.TC2_FLAGS | 1024)
    y_plain = E.tc2_s1(x, wk, k, False, e, allow_split=e is None)
    assert lib.spi_tc_error() == 0
    ref = F.conv2d(x.double(), w[0].double(), padding=k // 2)
    if epi:
        ref = ref + e['noise'].double() * 0.7 + e['b'].double().view(1, -1, 1, 1)
        ref = (torch.where(ref > 0, ref, ref * 0.2) * 1.4).clamp(-2.5, 2.5)
    assert y_swapped.shape == ref.shape and y_swapped.is_contiguous(memory_format=CL)
    assert rel_l2(y_plain, ref) < TOL and rel_l2(y_swapped, ref) < TOL
    assert rel_l2(y_swapped, y_plain) < 1e-6          # same operands, same accumulation order over (chunk, tap, k): rounding-identical
