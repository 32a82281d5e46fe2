"""Fused bias + activation (drop-in for `torch_utils.ops.bias_act`, eg3d/torch_utils/ops/bias_act.py).

Same Python surface: `activation_funcs` table, `bias_act(x, b, dim, act, alpha, gain, clamp, impl)`.
The arithmetic runs in `spi_bias_act` (spi_b200/csrc/bias_act.cu) -- forward, first- and second-order gradient modes,
exactly the plugin's `grad` = 0/1/2 protocol (bias_act.cpp:36).  CUDA only: a CPU tensor raises.
"""
import math
from types import SimpleNamespace

import torch

from ... import _lib

# eg3d/torch_utils/ops/bias_act.py:23-33 (cuda_idx = kernel id, ref = which tensor the backward needs)
activation_funcs = {
    'linear':   SimpleNamespace(def_alpha=0,   def_gain=1,            cuda_idx=1, ref='',  has_2nd_grad=False),
    'relu':     SimpleNamespace(def_alpha=0,   def_gain=math.sqrt(2), cuda_idx=2, ref='y', has_2nd_grad=False),
    'lrelu':    SimpleNamespace(def_alpha=0.2, def_gain=math.sqrt(2), cuda_idx=3, ref='y', has_2nd_grad=False),
    'tanh':     SimpleNamespace(def_alpha=0,   def_gain=1,            cuda_idx=4, ref='y', has_2nd_grad=True),
    'sigmoid':  SimpleNamespace(def_alpha=0,   def_gain=1,            cuda_idx=5, ref='y', has_2nd_grad=True),
    'elu':      SimpleNamespace(def_alpha=0,   def_gain=1,            cuda_idx=6, ref='y', has_2nd_grad=True),
    'selu':     SimpleNamespace(def_alpha=0,   def_gain=1,            cuda_idx=7, ref='y', has_2nd_grad=True),
    'softplus': SimpleNamespace(def_alpha=0,   def_gain=1,            cuda_idx=8, ref='y', has_2nd_grad=True),
    'swish':    SimpleNamespace(def_alpha=0,   def_gain=math.sqrt(2), cuda_idx=9, ref='x', has_2nd_grad=True),
}


def _dense_like(x):
    """Memory format the op works in: channels_last if x already is, else contiguous (bias_act.py:137-138)."""
    if x.ndim == 4 and x.stride(1) == 1 and x.shape[1] > 1:
        return torch.channels_last
    return torch.contiguous_format


def _fused_reductions(dx, want_db, noise=None, want_dpix=False, want_ds=False):
    """db / per-pixel sum / dstrength from one pass over a channels-last fp32 dx; None when the fast path does not apply."""
    if not (dx.ndim == 4 and dx.dtype == torch.float32 and dx.is_contiguous(memory_format=torch.channels_last) and dx.data_ptr() % 16 == 0):
        return None
    n, c, h, w = dx.shape
    if c % 4 != 0:          # RGB outputs of the toRGB layers: bias gradient only (small-C kernel)
        if not (c <= 8 and want_db and not want_dpix and not want_ds and dx.numel() % 4 == 0 and dx.numel() > 0):
            return None
        db = torch.empty(c, device=dx.device)
        _lib.check(_lib.load().spi_epilogue_grad_reduce(_lib.ptr(dx), n * h * w, c, h * w, None, _lib.ptr(db), None, None, _lib.stream()))
        return db, None, None
    if not 4 <= c <= 1024:
        return None
    if want_db and want_ds:                      # packed so the library zero-fills both with one memset
        buf = torch.empty(c + 1, device=dx.device)
        db, ds = buf[:c], buf[c]
    else:
        db = torch.empty(c, device=dx.device) if want_db else None
        ds = torch.empty((), device=dx.device) if want_ds else None
    dpix = torch.empty(h, w, device=dx.device) if want_dpix else None
    _lib.check(_lib.load().spi_epilogue_grad_reduce(_lib.ptr(dx), n * h * w, c, h * w, _lib.ptr(noise) if noise is not None else None,
                                                    _lib.ptr(db), _lib.ptr(dpix), ds.data_ptr() if ds is not None else None, _lib.stream()))
    return db, dpix, ds


def _grad_and_reductions(dy, b, y, cfg, noise=None, want_db=False, want_dpix=False, want_ds=False):
    """(dx, db, dpix, ds) of an activation epilogue given its saved output y: dx = bias_act(grad=1), db = sum of dx over pixels, dpix = sum of
    dx over channels ([H,W]), ds = sum dx * noise.  One kernel (`spi_bias_act_grad_reduce`) when only db / ds are wanted -- always the case
    while the generator is tuned: noise_const is a buffer then -- on channels-last fp32 with linear / relu / lrelu; otherwise the gradient
    kernel followed by the reduction kernel (`_fused_reductions`) or ATen sums."""
    dim, spec, alpha, gain, clamp = cfg
    fast = ((want_db or want_ds) and not want_dpix and not torch.is_grad_enabled() and dim == 1 and spec.cuda_idx in (1, 2, 3) and dy.ndim == 4
            and dy.dtype == torch.float32 and y is not None and y.dtype == torch.float32 and dy.is_contiguous(memory_format=torch.channels_last)
            and dy.stride() == y.stride() and dy.shape == y.shape and dy.shape[1] % 4 == 0 and 4 <= dy.shape[1] <= 1024
            and dy.data_ptr() % 16 == 0 and y.data_ptr() % 16 == 0 and dy.numel() >= 4096 and (not want_ds or noise is not None))
    if fast:
        n, c, h, w = dy.shape
        dx = torch.empty_like(dy)
        if want_db and want_ds:                      # packed so the library zero-fills both with one memset
            buf = torch.empty(c + 1, device=dy.device)
            db, ds = buf[:c], buf[c]
        else:
            db = torch.empty(c, device=dy.device) if want_db else None
            ds = torch.empty((), device=dy.device) if want_ds else None
        with _lib.timed('bias_act', dy.numel() * 4 * 3, detail=f'grad1+reduce act{spec.cuda_idx} {tuple(dy.shape)}'):
            _lib.check(_lib.load().spi_bias_act_grad_reduce(_lib.ptr(dy), _lib.ptr(y), _lib.ptr(dx), dy.numel(), c, h * w, spec.cuda_idx, alpha, gain, clamp,
                                                            _lib.ptr(noise) if want_ds else None, _lib.ptr(db), ds.data_ptr() if ds is not None else None,
                                                            _lib.stream()))
        return dx, db, None, ds
    dx = _BiasActGrad.apply(dy, None, b, y, cfg)
    db = dpix = ds = None
    if want_db or want_dpix or want_ds:
        red = _fused_reductions(dx, want_db, noise=noise, want_dpix=want_dpix, want_ds=want_ds) if dim == 1 else None
        if red is not None:
            db, dpix, ds = red
        else:
            if want_db:
                db = dx.sum([i for i in range(dx.ndim) if i != dim])
            if want_dpix or want_ds:
                dpix = dx.sum([0, 1])
                ds = (dpix * noise).sum() if want_ds else None
    return dx, db, dpix, ds


def _plugin_bias_act(x, b, xref, yref, dy, grad, dim, act_idx, alpha, gain, clamp):
    """The plugin entry point `bias_act(x, b, xref, yref, dy, grad, dim, act, alpha, gain, clamp)` (bias_act.cpp:36)."""
    if not x.is_cuda:
        raise RuntimeError('x must reside on CUDA device')
    if b is not None and b.numel():
        if b.ndim != 1:
            raise RuntimeError('b must have rank 1')
        if not (0 <= dim < x.ndim):
            raise RuntimeError('dim is out of bounds')
        if b.numel() != x.shape[dim]:
            raise RuntimeError('b has wrong number of elements')
        if b.dtype != x.dtype:
            raise RuntimeError('b must have the same dtype and device as x')
    y = torch.empty_like(x)        # preserves a dense layout
    if x.numel() == 0:
        return y
    for t in (xref, yref, dy):
        if t is not None and t.numel() and (t.shape != x.shape or t.stride() != x.stride()):
            raise RuntimeError('xref/yref/dy must have the same shape and layout as x')
    has_b = b is not None and b.numel() > 0
    nbytes = x.numel() * x.element_size() * (2 + (xref is not None) + (yref is not None) + (dy is not None))
    with _lib.timed('bias_act', nbytes, detail=f'grad{grad} act{act_idx} {tuple(x.shape)}'):
        _lib.check(_lib.load().spi_bias_act(
        _lib.ptr(x), _lib.ptr(b), _lib.ptr(xref), _lib.ptr(yref), _lib.ptr(dy), _lib.ptr(y), x.numel(),
            b.numel() if has_b else 1, x.stride(dim) if has_b else 1, _lib.dtype_code(x), grad, act_idx,
            alpha, gain, clamp, _lib.stream()))
    return y


class _BiasActGrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, dy, x, b, y, cfg):
        dim, spec, alpha, gain, clamp = cfg
        dx = _plugin_bias_act(dy, b, x, y, None, 1, dim, spec.cuda_idx, alpha, gain, clamp)
        ctx.cfg = cfg
        ctx.save_for_backward(dy if spec.has_2nd_grad else None, x, b, y)
        return dx

    @staticmethod
    def backward(ctx, d_dx):
        dim, spec, alpha, gain, clamp = ctx.cfg
        dy, x, b, y = ctx.saved_tensors
        d_dx = d_dx.contiguous(memory_format=_dense_like(d_dx))
        d_dy = d_x = d_b = None
        if ctx.needs_input_grad[0]:
            d_dy = _BiasActGrad.apply(d_dx, x, b, y, ctx.cfg)
        if spec.has_2nd_grad and (ctx.needs_input_grad[1] or ctx.needs_input_grad[2]):
            d_x = _plugin_bias_act(d_dx, b, x, y, dy, 2, dim, spec.cuda_idx, alpha, gain, clamp)
            if ctx.needs_input_grad[2]:
                d_b = d_x.sum([i for i in range(d_x.ndim) if i != dim])
        return d_dy, d_x, d_b, None, None


class _BiasAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, b, cfg):
        dim, spec, alpha, gain, clamp = cfg
        fmt = _dense_like(x)
        x = x.contiguous(memory_format=fmt)
        b = b.contiguous() if b is not None else None
        y = x
        if spec.cuda_idx != 1 or gain != 1 or clamp >= 0 or b is not None:
            y = _plugin_bias_act(x, b, None, None, None, 0, dim, spec.cuda_idx, alpha, gain, clamp)
        keep_x = 'x' in spec.ref or spec.has_2nd_grad
        # y is also kept when a clamp is active so that the gradient is masked where the output saturated -- the
        # behaviour of the reference's CPU path (`_bias_act_ref`, the oracle).  The reference CUDA plugin drops y for
        # act='linear' (spec.ref == '') and lets the gradient through the clamp; see DESIGN.md 'Known deviations'.
        ctx.save_for_backward(x if keep_x else None, b if keep_x else None, y if ('y' in spec.ref or clamp >= 0) else None)
        ctx.cfg, ctx.fmt = cfg, fmt
        return y

    @staticmethod
    def backward(ctx, dy):
        dim, spec, alpha, gain, clamp = ctx.cfg
        x, b, y = ctx.saved_tensors
        dy = dy.contiguous(memory_format=ctx.fmt)
        dx = db = None
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            dx = dy
            if spec.cuda_idx != 1 or gain != 1 or clamp >= 0:
                if x is None and y is not None and ctx.needs_input_grad[1] and dim == 1:
                    dx, db, _, _ = _grad_and_reductions(dy, b, y, ctx.cfg, want_db=True)     # gradient and bias gradient in one pass
                    return dx, db, None
                dx = _BiasActGrad.apply(dy, x, b, y, ctx.cfg)
        if ctx.needs_input_grad[1]:
            red = _fused_reductions(dx, True) if dim == 1 else None
            db = red[0] if red is not None else dx.sum([i for i in range(dx.ndim) if i != dim])
        return dx, db, None


def bias_act(x, b=None, dim=1, act='linear', alpha=None, gain=None, clamp=None, impl='cuda'):
    """y = clamp(act(x + b) * gain); see eg3d/torch_utils/ops/bias_act.py:54-88 for the argument contract."""
    assert isinstance(x, torch.Tensor)
    assert impl in ['ref', 'cuda']
    assert clamp is None or clamp >= 0
    spec = activation_funcs[act]
    cfg = (dim, spec, float(alpha if alpha is not None else spec.def_alpha),
           float(gain if gain is not None else spec.def_gain), float(clamp if clamp is not None else -1))
    if not x.is_cuda:
        raise RuntimeError('spi_b200.bias_act: x must reside on a CUDA device (no CPU path in this build)')
    return _BiasAct.apply(x, b, cfg)


class _BiasActNoise(torch.autograd.Function):
    """SynthesisLayer epilogue (networks_stylegan2.py:320-329) in one pass: y = clamp(act(x + noise*strength + b) * gain)."""

    @staticmethod
    def forward(ctx, x, b, noise_const, noise_strength, cfg):
        dim, spec, alpha, gain, clamp = cfg
        assert dim == 1 and x.ndim == 4 and 'x' not in spec.ref
        fmt = _dense_like(x)
        x = x.contiguous(memory_format=fmt)
        n, c, h, w = x.shape
        y = torch.empty_like(x)
        nc = noise_const.contiguous()
        _lib.check(_lib.load().spi_bias_act_noise(
            _lib.ptr(x), _lib.ptr(b.contiguous()), _lib.ptr(y), _lib.ptr(nc), _lib.ptr(noise_strength), x.numel(), c, x.stride(1),
            h * w, c if fmt == torch.channels_last else 0, _lib.dtype_code(x), spec.cuda_idx, alpha, gain, clamp, _lib.stream()))
        ctx.save_for_backward(b, y, nc, noise_strength)
        ctx.cfg, ctx.fmt = cfg, fmt
        return y

    @staticmethod
    def backward(ctx, dy):
        dim, spec, alpha, gain, clamp = ctx.cfg
        b, y, nc, strength = ctx.saved_tensors
        dy = dy.contiguous(memory_format=ctx.fmt)
        need_b, need_n, need_s = ctx.needs_input_grad[1], ctx.needs_input_grad[2], ctx.needs_input_grad[3]
        dx, db, pix, ds = _grad_and_reductions(dy, b, y, ctx.cfg, noise=nc, want_db=need_b, want_dpix=need_n, want_ds=need_s)
        dn = pix * strength if need_n else None
        return dx, db, dn, ds, None


def bias_act_noise(x, b, noise_const, noise_strength, act='lrelu', alpha=None, gain=None, clamp=None):
    """`bias_act(x + noise_const * noise_strength, b, ...)` fused (x: [N,C,H,W] fp32, noise_const [H,W], noise_strength [])."""
    assert clamp is None or clamp >= 0
    spec = activation_funcs[act]
    cfg = (1, spec, float(alpha if alpha is not None else spec.def_alpha),
           float(gain if gain is not None else spec.def_gain), float(clamp if clamp is not None else -1))
    if not x.is_cuda:
        raise RuntimeError('spi_b200.bias_act_noise: x must reside on a CUDA device (no CPU path in this build)')
    if x.dtype != torch.float32 or 'x' in spec.ref:
        return bias_act(x + noise_const * noise_strength, b, act=act, alpha=alpha, gain=gain, clamp=clamp)
    return _BiasActNoise.apply(x, b, noise_const, noise_strength, cfg)


class _BlurBiasActNoise(torch.autograd.Function):
    """Tail of an up-sampling SynthesisLayer in one pass (`spi_blur4_bias_act_noise`): the 4x4 FIR that follows the stride-2
    transposed convolution (conv2d_resample.py:117-119) with the noise / bias / activation / clamp epilogue
    (networks_stylegan2.py:320-329) applied in its store, so the blurred tensor never goes to HBM un-activated.  Backward is the
    unfused chain: epilogue gradient from the saved output, its reductions, then the FIR's own backward (upfirdn2d.py:258-263)."""

    @staticmethod
    def forward(ctx, x, f, b, noise_const, noise_strength, fir_cfg, cfg):
        padx0, padx1, pady0, pady1, flip, fir_gain = fir_cfg
        dim, spec, alpha, gain, clamp = cfg
        n, c, ih, iw = x.shape
        oh, ow = ih + pady0 + pady1 - 3, iw + padx0 + padx1 - 3
        y = torch.empty([n, c, oh, ow], dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
        nc = noise_const.contiguous() if noise_const is not None else None
        f = f.to(x.device).contiguous()
        with _lib.timed('upfirdn2d', (x.numel() + y.numel()) * 4, detail=f'blur4+bias_act {tuple(x.shape)}'):
            _lib.check(_lib.load().spi_blur4_bias_act_noise(
                _lib.ptr(x), _lib.ptr(f), _lib.ptr(y), _lib.ptr(b.contiguous()), _lib.ptr(nc), _lib.ptr(noise_strength), n, c, ih, iw,
                _lib.strides4(x), _lib.strides4(y), padx0, padx1, pady0, pady1, int(bool(flip)), float(fir_gain), spec.cuda_idx, alpha,
                gain, clamp, _lib.stream()))
        ctx.save_for_backward(f, b, y, nc, noise_strength)
        ctx.cfg, ctx.fir_cfg, ctx.x_shape = cfg, fir_cfg, x.shape
        return y

    @staticmethod
    def backward(ctx, dy):
        from . import upfirdn2d
        f, b, y, nc, strength = ctx.saved_tensors
        padx0, padx1, pady0, pady1, flip, fir_gain = ctx.fir_cfg
        dy = dy.contiguous(memory_format=torch.channels_last)
        need_b, need_n, need_s = ctx.needs_input_grad[2], nc is not None and ctx.needs_input_grad[3], nc is not None and ctx.needs_input_grad[4]
        dpre, db, pix, ds = _grad_and_reductions(dy, b, y, ctx.cfg, noise=nc, want_db=need_b, want_dpix=need_n, want_ds=need_s)
        dn = pix * strength if need_n else None
        dx = None
        if ctx.needs_input_grad[0]:
            _, _, ih, iw = ctx.x_shape
            _, _, oh, ow = dpre.shape
            p = (4 - padx0 - 1, iw - ow + padx0, 4 - pady0 - 1, ih - oh + pady0)
            dx = upfirdn2d._Upfirdn2d.apply(dpre, f, (1, 1, 1, 1, *p, not flip, fir_gain))
        return dx, None, db, dn, ds, None, None


def blur_bias_act_noise(x, f, b, noise_const=None, noise_strength=None, padding=0, flip_filter=False, fir_gain=1, act='lrelu', alpha=None,
                        gain=None, clamp=None):
    """`bias_act(upfirdn2d(x, f, padding=padding, gain=fir_gain) + noise_const * noise_strength, b, act, gain, clamp)` for a 4x4 filter
    at up = down = 1; one kernel when x is channels-last fp32 with C % 4 == 0 and the activation is linear / lrelu, the two ops otherwise."""
    from . import upfirdn2d
    spec = activation_funcs[act]
    padx0, padx1, pady0, pady1 = upfirdn2d._parse_padding(padding)
    fusable = (x.is_cuda and x.dtype == torch.float32 and x.ndim == 4 and x.shape[1] % 4 == 0 and x.stride(1) == 1 and x.shape[1] > 1
               and f is not None and f.ndim == 2 and tuple(f.shape) == (4, 4) and f.dtype == torch.float32 and act in ('linear', 'lrelu')
               and b is not None and x.data_ptr() % 16 == 0 and all(st % 4 == 0 for st in (x.stride(0), x.stride(2), x.stride(3)))
               and x.shape[2] + pady0 + pady1 >= 4 and x.shape[3] + padx0 + padx1 >= 4)
    if not fusable:
        y = upfirdn2d.upfirdn2d(x, f, padding=padding, flip_filter=flip_filter, gain=fir_gain)
        if noise_const is not None:
            return bias_act_noise(y, b, noise_const, noise_strength, act=act, alpha=alpha, gain=gain, clamp=clamp)
        return bias_act(y, b, act=act, alpha=alpha, gain=gain, clamp=clamp)
    assert clamp is None or clamp >= 0
    cfg = (1, spec, float(alpha if alpha is not None else spec.def_alpha),
           float(gain if gain is not None else spec.def_gain), float(clamp if clamp is not None else -1))
    return _BlurBiasActNoise.apply(x, f, b, noise_const, noise_strength, (padx0, padx1, pady0, pady1, bool(flip_filter), fir_gain), cfg)
