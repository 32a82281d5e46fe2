"""Dense convolution engine for the StyleGAN2 / SR / VGG layers (the "implicit GEMM on tensor cores" row of SURVEY.md §8d).
Call-site contract = `_conv2d_wrapper` (eg3d/torch_utils/ops/conv2d_resample.py:30-43): `flip_weight=True` is correlation
(what `F.conv2d` computes), `False` flips the taps first.

Default engine ('tc2'): this library's halo-patch implicit GEMM on tcgen05 + TMEM + TMA (spi_b200/csrc/conv_tc2.cu) for
  * the stride-1 'same' 3x3 / 1x1 correlations (forward, and data gradient = the same kernel on dy with transposed, reversed taps),
    optionally with the SynthesisLayer / bias_act epilogue (noise, bias, lrelu / relu, gain, clamp) applied in the accumulator read-out,
  * the stride-2 transposed 3x3 convolution of the up-sampling layers (forward) and the stride-2 3x3 correlation (its data gradient),
  * the weight gradients of all of these (`spi_conv_wgrad_tc2`).
`SPI_CONV_ENGINE=cudnn` switches every contraction back to cuDNN's TF32 kernels through ATen (A/B timing; round-1 behaviour).  Shapes
the engine does not take (channel counts that are not multiples of 32: the RGB outputs, the 3-channel VGG stem) go to small dedicated
kernels or to ATen, see `conv2d`.

Modulated convolutions carry one weight set per sample (`conv2d_per_sample`): each sample is its own dense problem, executed inside
ONE autograd node so that no slice / cat / zero-fill traffic is generated around the per-sample calls.
"""
import os

import torch
import torch.nn.functional as F

from .. import _lib
from . import zero_arena

ALLOW_TF32 = True
CUDNN_AUTOTUNE = True      # cudnn.benchmark for the cuDNN arm (every shape is first seen in an eager warm-up pass)
ENGINE = os.environ.get('SPI_CONV_ENGINE', 'tc2')          # 'tc2' | 'cudnn'
WGRAD_ENGINE = os.environ.get('SPI_CONV_WGRAD', 'tc2')     # 'tc2' | 'cudnn'
CL = torch.channels_last
FUSE_EPILOGUE = os.environ.get('SPI_CONV_FUSE', '1') != '0'      # 0: convolution and bias_act as separate kernels (A/B timing, debugging)
TC2_FLAGS = int(os.environ.get('SPI_TC2_FLAGS', '0'))            # debug flags of spi_conv2d_tc2 (4: one M tile per CTA, 128: no N = 256 tiles)


def _check(x):
    if not x.is_cuda:
        raise RuntimeError('spi_b200 conv engine: x must reside on a CUDA device (no CPU path in this build)')


def _pair(v):
    return [v, v] if isinstance(v, int) else [int(v[0]), int(v[1])]


def _cudnn_flags():
    torch.backends.cudnn.allow_tf32 = ALLOW_TF32
    torch.backends.cudnn.benchmark = CUDNN_AUTOTUNE


# ---------------------------------------------------------------------------------------------------------------------
# tc2 primitives (raw tensors in, raw tensors out; no autograd)

def tc2_form(x, w5, stride, padding, transpose):
    """Which tc2 kernel takes conv(x, w5[g]) -- 's1', 't2' (stride-2 transposed) or None.  w5 logical [G,O,I,kh,kw]."""
    if ENGINE != 'tc2' or x.dtype != torch.float32 or w5.dtype != torch.float32:
        return None
    o, i, kh, kw = w5.shape[1:]
    st, pd = _pair(stride), _pair(padding)
    if kh == kw == 1 and o <= 4 and i % 128 == 0 and i <= 512 and not transpose and st == [1, 1] and pd == [0, 0]:
        return 'rgb'             # RGB heads: a streaming 1x1 contraction (spi_b200/csrc/conv_rgb.cu)
    if i % 32 or o % 32 or kh != kw:
        return None
    if not transpose and st == [1, 1] and kh in (1, 3) and pd == [kh // 2, kh // 2]:
        return 's1'
    if transpose and st == [2, 2] and kh == 3 and pd == [0, 0]:
        return 't2'
    return None


def _ohwi(w5):
    """[G,O,I,kh,kw] logical -> contiguous memory [G][O][kh*kw][I] (no copy when modulate_weights wrote the 'ohwi' layout)."""
    return w5.permute(0, 1, 3, 4, 2).contiguous()


PREZEROED = 512      # conv flag: the output handed in is already zero (ops/zero_arena.py)


def _tc2_output(form, n, co, ho, wo, h, wd, ci, k, per_sample, epilogue, flags, device):
    """Output tensor of a tc2 convolution: when the call will split Cin (reduce-add partial sums into a zeroed output) and the iteration has
    a zero arena, a slice of the arena plus the flag that tells the library not to fill it again; a plain uninitialised tensor otherwise."""
    if zero_arena.recording():
        splits = _lib.load().spi_conv_tc2_splits(form, n, h, wd, ci, co, k, int(per_sample), int(bool(epilogue)), flags)
        if splits > 1:
            y = zero_arena.take((n, co, ho, wo), channels_last=True)
            if y is not None:
                return y, flags | PREZEROED
    return torch.empty(n, co, ho, wo, device=device, dtype=torch.float32, memory_format=CL), flags


def tc2_s1(x, wk, k, per_sample, epilogue=None, allow_split=True):
    """Stride-1 'same' correlation.  x [N,I,H,W] channels-last, wk memory [G][O][k*k][I].  epilogue = dict(b, noise, strength, act, slope, gain, clamp)."""
    n, ci, h, wd = x.shape
    co = wk.shape[1]
    e = epilogue or {}
    y, flags = _tc2_output(0, n, co, h, wd, h, wd, ci, k, per_sample, epilogue, (0 if allow_split else 32) | TC2_FLAGS, x.device)
    with _lib.timed('conv', 2 * n * h * wd * k * k * ci * co, detail=f's1 k{k} {n}x{ci}->{co} @{h}x{wd}' + (' +epilogue' if epilogue else '')):
        _lib.check(_lib.load().spi_conv2d_tc2(_lib.ptr(x), _lib.ptr(wk), _lib.ptr(y), n, h, wd, ci, co, k, int(per_sample), _lib.ptr(e.get('b')),
                                              _lib.ptr(e.get('noise')), _lib.ptr(e.get('strength')), int(e.get('act', 0)), float(e.get('slope', 0.2)),
                                              float(e.get('gain', 1.0)), float(e.get('clamp', -1.0)), flags, _lib.stream()))
    return y


def tc2_t2(x, wk, per_sample):
    """Stride-2 transposed 3x3 convolution, no padding: [N,I,H,W] -> [N,O,2H+1,2W+1]; wk memory [G][O][9][I]."""
    n, ci, h, wd = x.shape
    co = wk.shape[1]
    y, flags = _tc2_output(1, n, co, 2 * h + 1, 2 * wd + 1, h, wd, ci, 3, per_sample, None, 0, x.device)
    with _lib.timed('conv', 2 * n * h * wd * 9 * ci * co, detail=f't2 {n}x{ci}->{co} @{h}x{wd}'):
        _lib.check(_lib.load().spi_conv_transpose2d_s2_tc2(_lib.ptr(x), _lib.ptr(wk), _lib.ptr(y), n, h, wd, ci, co, int(per_sample), flags, _lib.stream()))
    return y


def tc2_s2(x, wk, per_sample):
    """Stride-2 3x3 correlation, no padding: [N,I,2H+1,2W+1] -> [N,O,H,W]; wk memory [G][O][9][I]."""
    n, ci, hi, wi = x.shape
    h, wd = (hi - 1) // 2, (wi - 1) // 2
    co = wk.shape[1]
    y, flags = _tc2_output(2, n, co, h, wd, h, wd, ci, 3, per_sample, None, 0, x.device)
    with _lib.timed('conv', 2 * n * h * wd * 9 * ci * co, detail=f's2 {n}x{ci}->{co} ->{h}x{wd}'):
        _lib.check(_lib.load().spi_conv2d_s2_tc2(_lib.ptr(x), _lib.ptr(wk), _lib.ptr(y), n, h, wd, ci, co, int(per_sample), flags, _lib.stream()))
    return y


def tc2_wT(wk, reverse):
    """wk [G][O][T][I] -> [G][I][T'][O] (taps reversed when `reverse`): the weights of the data-gradient convolution."""
    g, o, t, i = wk.shape
    wt = torch.empty(g, i, t, o, device=wk.device, dtype=wk.dtype)
    _lib.check(_lib.load().spi_conv_weight_transpose(_lib.ptr(wk), _lib.ptr(wt), g, o, t, i, int(reverse), _lib.stream()))
    return wt


def _cl(t):
    return t if t.is_contiguous(memory_format=CL) else t.contiguous(memory_format=CL)


def rgb_conv(which, a, b, n, pixels, ci, co, per_sample, out):
    with _lib.timed('conv_rgb', (pixels * n * (ci + co)) * 4):
        _lib.check(_lib.load().spi_conv1x1_rgb(which, _lib.ptr(a), _lib.ptr(b), _lib.ptr(out), pixels, n, ci, co, int(per_sample), _lib.stream()))
    return out


def _tc2_forward(form, x, w5, epilogue=None):
    per_sample = w5.shape[0] > 1
    if form == 'rgb':
        n, ci, h, wd = x.shape
        wk = w5.reshape(w5.shape[0], w5.shape[1], ci).contiguous()
        y = torch.empty(n, wk.shape[1], h, wd, device=x.device, dtype=torch.float32, memory_format=CL)
        return rgb_conv(0, x, wk, n, h * wd, ci, wk.shape[1], per_sample, y), wk
    wk = _ohwi(w5)
    k = w5.shape[-1]
    wk = wk.view(wk.shape[0], wk.shape[1], k * k, wk.shape[-1])
    if form == 's1':
        return tc2_s1(x, wk, k, per_sample, epilogue, allow_split=epilogue is None), wk
    return tc2_t2(x, wk, per_sample), wk


_WT_CACHE = {}      # transposed copies of FROZEN weights (the VGG extractors), keyed by storage address + version


def _wT(wk, reverse, frozen):
    if not frozen:
        return tc2_wT(wk, reverse)
    key = (wk.data_ptr(), wk._version, tuple(wk.shape), bool(reverse))
    hit = _WT_CACHE.get(key)
    if hit is None:
        if len(_WT_CACHE) > 256:
            _WT_CACHE.clear()
        hit = _WT_CACHE[key] = (tc2_wT(wk, reverse), wk)        # keeps wk alive: its address cannot be recycled while the entry exists
    return hit[0]


def _tc2_input_grad(form, gy, wk, frozen=False, wT=None):
    """dL/dx.  `wT`: the transposed weights [G][I][taps'][O] when the caller already has them (ops/modulate.py `modulate_bank`)."""
    per_sample = wk.shape[0] > 1
    if form == 'rgb':
        n, co, h, wd = gy.shape
        ci = wk.shape[2]
        gx = torch.empty(n, ci, h, wd, device=gy.device, dtype=torch.float32, memory_format=CL)
        return rgb_conv(1, gy, wk, n, h * wd, ci, co, per_sample, gx)
    if form == 's1':
        k = int(round(wk.shape[2] ** 0.5))
        return tc2_s1(gy, wT if wT is not None else _wT(wk, True, frozen), k, per_sample)
    return tc2_s2(gy, wT if wT is not None else _wT(wk, False, frozen), per_sample)


def _weight_grad(form, gy, x, w5, stride, padding, transpose):
    """dL/dw5 (logical [G,O,I,kh,kw]) of y = conv(x, w5[g]); x, gy channels-last."""
    g = w5.shape[0]
    n = x.shape[0]
    o, i, kh, kw = w5.shape[1:]
    if form == 'rgb':
        gw = zero_arena.take((g, o, i))
        which = 2 if gw is None else 2 | 4                     # + 4: gw is already zero
        if gw is None:
            gw = torch.empty(g, o, i, device=x.device, dtype=torch.float32)
        return rgb_conv(which, x, gy, n, x.shape[2] * x.shape[3], i, o, g > 1, gw).view(g, o, i, 1, 1)
    if WGRAD_ENGINE == 'tc2' and form is not None:
        h, wd = x.shape[2], x.shape[3]
        mode = 0 if form == 's1' else 1
        # mode 0 writes [G][O][taps][I]; mode 1 (transposed convolution: the roles of x and dy swap) writes [G][I][taps][O]
        shape = (g, o, kh, kw, i) if mode == 0 else (g, i, kh, kw, o)
        gw = zero_arena.take(shape)                            # the kernel reduce-adds its partial sums into gw
        zeroed = 0 if gw is None else 4
        if gw is None:
            gw = torch.empty(shape, device=x.device, dtype=torch.float32)
        with _lib.timed('conv', 2 * n * h * wd * kh * kw * i * o, detail=f'wgrad{mode} k{kh} {n}x{i}->{o} @{h}x{wd}'):
            _lib.check(_lib.load().spi_conv_wgrad_tc2(_lib.ptr(x), _lib.ptr(gy), _lib.ptr(gw), n, h, wd, i, o, kh, int(g > 1), mode | zeroed, _lib.stream()))
        return gw.permute(0, 1, 4, 2, 3) if mode == 0 else gw.permute(0, 4, 1, 2, 3)
    _cudnn_flags()
    gw = torch.empty_like(w5) if g > 1 else None          # preserves w5's (conv-native) strides
    for k in ([slice(0, n)] if g == 1 else [slice(j, j + 1) for j in range(n)]):
        wi = 0 if g == 1 else k.start
        wk = (w5[wi].transpose(0, 1) if transpose else w5[wi]).contiguous(memory_format=CL)
        _, gwk, _ = torch.ops.aten.convolution_backward(gy[k], x[k], wk, None, _pair(stride), _pair(padding), [1, 1], transpose, [0, 0], 1,
                                                        [False, True, False])
        gwk = gwk.transpose(0, 1) if transpose else gwk
        if g == 1:
            return gwk.unsqueeze(0)
        gw[wi].copy_(gwk)
    return gw


# ---------------------------------------------------------------------------------------------------------------------
# autograd nodes

class _PerSampleConv(torch.autograd.Function):
    """y[n] = conv(x[n], w[n]) (or conv_transpose) for n in range(N); x [N,C,H,W] channels-last, w [G,O,I,kh,kw] logical, G in {1, N}."""

    @staticmethod
    def forward(ctx, x, w, stride, padding, transpose, wT=None):
        n, _, h, wd = x.shape
        o, kh, kw = w.shape[1], w.shape[3], w.shape[4]
        st, pd = _pair(stride), _pair(padding)
        x = _cl(x)
        form = tc2_form(x, w, stride, padding, transpose)
        ctx.cfg = (stride, padding, transpose, form)
        if form is not None:
            y, wk = _tc2_forward(form, x, w)
            ctx.save_for_backward(x, w, wk, wT)
            return y
        _cudnn_flags()
        if transpose:
            oh, ow = (h - 1) * st[0] - 2 * pd[0] + kh, (wd - 1) * st[1] - 2 * pd[1] + kw
        else:
            oh, ow = (h + 2 * pd[0] - kh) // st[0] + 1, (wd + 2 * pd[1] - kw) // st[1] + 1
        shared = (w.shape[0] == 1 and n > 1)       # one weight set for the whole batch: a single batched call
        ctx.save_for_backward(x, w, None, None)
        if shared or n == 1:
            wk = (w[0].transpose(0, 1) if transpose else w[0]).contiguous(memory_format=CL)
            if transpose:
                return torch.ops.aten.cudnn_convolution_transpose(x, wk, pd, [0, 0], st, [1, 1], 1, False, False, ALLOW_TF32)
            return torch.ops.aten.cudnn_convolution(x, wk, pd, st, [1, 1], 1, False, False, ALLOW_TF32)
        y = torch.empty(n, o, oh, ow, device=x.device, dtype=x.dtype, memory_format=CL)
        for k in [slice(i, i + 1) for i in range(n)]:
            wk = (w[k.start].transpose(0, 1) if transpose else w[k.start]).contiguous(memory_format=CL)
            if transpose:
                torch.ops.aten.cudnn_convolution_transpose.out(x[k], wk, pd, [0, 0], st, [1, 1], 1, False, False, ALLOW_TF32, out=y[k])
            else:
                torch.ops.aten.cudnn_convolution.out(x[k], wk, pd, st, [1, 1], 1, False, False, ALLOW_TF32, out=y[k])
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w, wk, wT = ctx.saved_tensors
        stride, padding, transpose, form = ctx.cfg
        need_x, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        gy = _cl(gy)
        if form is not None:
            gx = _tc2_input_grad(form, gy, wk, frozen=not need_w, wT=wT) if need_x else None
            gw = _weight_grad(form, gy, x, w, stride, padding, transpose) if need_w else None
            return gx, gw, None, None, None, None
        _cudnn_flags()
        n = x.shape[0]
        shared = (w.shape[0] == 1 and n > 1)
        if shared or n == 1:          # one call covers the batch: hand cuDNN's own results to autograd (no slice copies)
            wc = (w[0].transpose(0, 1) if transpose else w[0]).contiguous(memory_format=CL)
            gx, gwk, _ = torch.ops.aten.convolution_backward(gy, x, wc, None, _pair(stride), _pair(padding), [1, 1], transpose, [0, 0], 1,
                                                             [need_x, need_w, False])
            gw = (gwk.transpose(0, 1) if transpose else gwk).unsqueeze(0) if need_w else None
            return gx, gw, None, None, None, None
        gx = torch.empty_like(x) if need_x else None
        gw = torch.empty_like(w) if need_w else None          # preserves w's (conv-native) strides
        for k in [slice(i, i + 1) for i in range(n)]:
            wc = (w[k.start].transpose(0, 1) if transpose else w[k.start]).contiguous(memory_format=CL)
            gxk, gwk, _ = torch.ops.aten.convolution_backward(gy[k], x[k], wc, None, _pair(stride), _pair(padding), [1, 1],
                                                              transpose, [0, 0], 1, [need_x, need_w, False])
            if need_x:
                gx[k].copy_(gxk)
            if need_w:
                gw[k.start].copy_(gwk.transpose(0, 1) if transpose else gwk)
        return gx, gw, None, None, None, None


class _ConvBiasActNoise(torch.autograd.Function):
    """A non-resampling SynthesisLayer in one kernel: y = clamp(act(conv(x, w[g]) + noise*strength + b) * gain), the epilogue of
    networks_stylegan2.py:320-329 applied while the accumulators are read out of tensor memory, so the convolution result never goes
    to HBM un-activated.  Backward: epilogue gradient from the saved OUTPUT (lrelu / relu / linear are sign-preserving), its bias /
    noise reductions, then the data- and weight-gradient convolutions."""

    @staticmethod
    def forward(ctx, x, w, b, noise_const, noise_strength, cfg, wT=None):
        dim, spec, alpha, gain, clamp = cfg
        x = _cl(x)
        nc = noise_const.contiguous() if noise_const is not None else None
        epi = dict(b=b.contiguous() if b is not None else None, noise=nc, strength=noise_strength, act={'linear': 0, 'relu': 1, 'lrelu': 2}[spec.name],
                   slope=alpha, gain=gain, clamp=clamp)
        y, wk = _tc2_forward('s1', x, w, epi)
        ctx.save_for_backward(x, w, wk, b, y, nc, noise_strength, wT)
        ctx.cfg = cfg
        return y

    @staticmethod
    def backward(ctx, dy):
        from ..torch_utils.ops import bias_act as BA
        x, w, wk, b, y, nc, strength, wT = ctx.saved_tensors
        dy = _cl(dy)
        if dy.stride() != y.stride():            # size-1 dimensions: "channels-last contiguous" does not pin their strides
            dy = torch.empty_like(y).copy_(dy)
        need_b = b is not None and ctx.needs_input_grad[2]
        need_n, need_s = nc is not None and ctx.needs_input_grad[3], nc is not None and ctx.needs_input_grad[4]
        dpre, db, pix, ds = BA._grad_and_reductions(dy, b, y, ctx.cfg, noise=nc, want_db=need_b, want_dpix=need_n, want_ds=need_s)
        dn = pix * strength if need_n else None
        gx = _tc2_input_grad('s1', dpre, wk, frozen=not ctx.needs_input_grad[1], wT=wT) if ctx.needs_input_grad[0] else None
        k = w.shape[-1]
        gw = _weight_grad('s1', dpre, x, w, 1, k // 2, False) if ctx.needs_input_grad[1] else None
        return gx, gw, db, dn, ds, None, None


def conv_bias_act_fusable(x, w5, act):
    """True when `conv2d_bias_act` runs as one kernel: tc2 engine, stride-1 'same' 3x3 / 1x1, 32-multiples of channels, act in {linear, relu,
    lrelu}, and a feature map large enough that the layer is not split over Cin (small maps trade the fusion for full SM occupancy)."""
    k = w5.shape[-1]
    return (FUSE_EPILOGUE and act in ('linear', 'relu', 'lrelu') and tc2_form(x, w5, 1, k // 2, False) == 's1'
            and x.shape[0] * x.shape[2] * x.shape[3] * ((w5.shape[1] + 127) // 128) >= 128 * 256)


def conv2d_bias_act(x, w5, b=None, noise_const=None, noise_strength=None, act='lrelu', alpha=None, gain=None, clamp=None, wT=None):
    """`bias_act(conv(x, w5[g]) + noise_const * noise_strength, b, act, gain, clamp)` for a stride-1 'same' convolution; one kernel when
    `conv_bias_act_fusable`, the separate ops otherwise.  `wT`: the tap-reversed transposed weights, when the caller has them."""
    from ..torch_utils.ops import bias_act as BA
    _check(x)
    k = w5.shape[-1]
    if conv_bias_act_fusable(x, w5, act):
        spec = BA.activation_funcs[act]
        spec.name = act
        cfg = (1, spec, float(alpha if alpha is not None else spec.def_alpha), float(gain if gain is not None else spec.def_gain),
               float(clamp if clamp is not None else -1))
        return _ConvBiasActNoise.apply(x, w5, b, noise_const, noise_strength, cfg, wT)
    y = _PerSampleConv.apply(x, w5, 1, k // 2, False, wT)
    if noise_const is not None:
        return BA.bias_act_noise(y, b, noise_const, noise_strength, act=act, alpha=alpha, gain=gain, clamp=clamp)
    return BA.bias_act(y, b, act=act, alpha=alpha, gain=gain, clamp=clamp)


def vgg_conv(x, conv, act='relu'):
    """`act(conv2d(x, weight, padding=1) + bias)` of a frozen 3x3 `nn.Conv2d` (the VGG16 / VGG19 extractors of the losses,
    spi/criteria/lpips/networks.py:53-63, bbox_cx_loss.py:76-91) as one engine kernel.  The 3-channel stem is zero-padded to 32 input
    channels (x: one small copy; weights: cached) so that it runs on the same implicit-GEMM kernel instead of a library call."""
    w = conv.weight
    if w.shape[1] % 32 and ENGINE == 'tc2' and w.shape[0] % 32 == 0 and x.dtype == torch.float32:
        ci = w.shape[1]
        cache = getattr(conv, '_spi_w32', None)
        if cache is None or cache[0] != w._version or cache[1].device != w.device:
            w32 = torch.zeros(w.shape[0], 32, w.shape[2], w.shape[3], device=w.device, dtype=w.dtype).contiguous(memory_format=CL)
            w32[:, :ci] = w.detach()
            cache = (w._version, w32)
            conv._spi_w32 = cache
        xp = torch.zeros(x.shape[0], 32, x.shape[2], x.shape[3], device=x.device, dtype=x.dtype).contiguous(memory_format=CL)
        xp[:, :ci] = x
        x, w = xp, cache[1]
    return conv2d_bias_act(x, w.unsqueeze(0), conv.bias, act=act, gain=1)


def conv2d(x, w, stride=1, padding=0, groups=1, transpose=False, flip_weight=True):
    """Shared-weight convolution (VGG layers, generic conv2d_resample callers).  w [O,I,kh,kw] ([I,O,kh,kw] when `transpose`)."""
    _check(x)
    if not flip_weight and (w.shape[-1] > 1 or w.shape[-2] > 1):
        w = w.flip([2, 3])
    if groups == 1:
        w5 = (w.transpose(0, 1) if transpose else w).unsqueeze(0)
        if tc2_form(x, w5, stride, padding, transpose) is not None:
            return _PerSampleConv.apply(x, w5, stride, padding, transpose)
    _cudnn_flags()
    op = F.conv_transpose2d if transpose else F.conv2d
    return op(x.contiguous(memory_format=CL), w.contiguous(memory_format=CL), stride=stride, padding=padding, groups=groups)


def conv2d_per_sample(x, w, stride=1, padding=0, transpose=False, flip_weight=True, wT=None):
    """x [N,Cin,H,W], w [N,Cout,Cin,kh,kw] (per-sample weights, any strides) -> [N,Cout,H',W'].
    For `transpose=True` the per-sample weight is used as conv_transpose2d's [Cin,Cout,kh,kw] (= w[n].transpose(0,1)).
    `wT` (optional): w as memory [N][Cin][taps][Cout], taps reversed for the stride-1 form -- the data-gradient convolution's weights."""
    _check(x)
    if not flip_weight and (w.shape[-1] > 1 or w.shape[-2] > 1):
        w, wT = w.flip([3, 4]), None
    return _PerSampleConv.apply(x, w, stride, padding, transpose, wT)


# ---------------------------------------------------------------------------------------------------------------------
# First-generation engine (spi_b200/csrc/conv_tc05.cu), kept for its tests and as a timing reference: one TMA box per (tap, chunk).

def conv2d_tc05(x, w, per_sample=False, bias=None, noise=None, noise_strength=None, act='linear', slope=0.2, gain=1.0, clamp=None):
    """Stride-1 'same' correlation on the first tcgen05 engine with the fused SynthesisLayer epilogue.
    x [N,I,H,W] (channels-last), w [G,O,I,kh,kw] logical with memory order [G][O][kh][kw][I] (G = N if per_sample else 1)."""
    _check(x)
    lib = _lib.load()
    n, ci, h, wd = x.shape
    if w.ndim == 4:
        w = w.unsqueeze(0)
    g, co, _, kh, kw = w.shape
    assert g == (n if per_sample else 1)
    if not lib.spi_conv2d_tc_supported(h, wd, ci, co, kh, kw):
        raise RuntimeError(f'conv2d_tc05: unsupported shape {tuple(x.shape)} x {tuple(w.shape)}')
    x = x.contiguous(memory_format=CL)
    wk = w.permute(0, 1, 3, 4, 2).contiguous()                  # [G][O][kh][kw][I]
    y = torch.empty(n, co, h, wd, device=x.device, dtype=x.dtype, memory_format=CL)
    code = {'linear': 0, 'relu': 1, 'lrelu': 2}[act]
    _lib.check(lib.spi_conv2d_tc(_lib.ptr(x), _lib.ptr(wk), _lib.ptr(y), n, h, wd, ci, co, kh, kw, int(per_sample), _lib.ptr(bias),
                                 _lib.ptr(noise), _lib.ptr(noise_strength), code, float(slope), float(gain),
                                 -1.0 if clamp is None else float(clamp), 0, _lib.stream()))
    return y


def conv2d_tc05_input_grad(gy, w, per_sample=False):
    """Gradient of conv2d_tc05 (no epilogue) w.r.t. x: the same kernel on dy with the flipped / transposed weights."""
    lib = _lib.load()
    if w.ndim == 4:
        w = w.unsqueeze(0)
    g, co, ci, kh, kw = w.shape
    wk = w.permute(0, 1, 3, 4, 2).contiguous()
    wt = torch.empty(g, ci, kh, kw, co, device=w.device, dtype=w.dtype)
    _lib.check(lib.spi_conv_weight_flip_transpose(_lib.ptr(wk), _lib.ptr(wt), g, co, kh * kw, ci, _lib.stream()))
    return conv2d_tc05(gy, wt.permute(0, 1, 4, 2, 3), per_sample=per_sample)
