"""Pad / upsample / FIR / downsample (drop-in for `torch_utils.ops.upfirdn2d`, eg3d/torch_utils/ops/upfirdn2d.py).

Same Python surface: `setup_filter`, `upfirdn2d`, `filter2d`, `upsample2d`, `downsample2d`.  The arithmetic runs in
`spi_upfirdn2d` (spi_b200/csrc/upfirdn2d.cu).  Backward = the same op with up<->down swapped, flipped filter and
the padding of upfirdn2d.py:258-263, so gradients of arbitrary order work.  CUDA only.
"""
import numpy as np
import torch

from ... import _lib


def _parse_scaling(scaling):
    if isinstance(scaling, int):
        scaling = [scaling, scaling]
    assert isinstance(scaling, (list, tuple)) and all(isinstance(v, int) for v in scaling)
    sx, sy = scaling
    assert sx >= 1 and sy >= 1
    return sx, sy


def _parse_padding(padding):
    if isinstance(padding, int):
        padding = [padding, padding]
    assert isinstance(padding, (list, tuple)) and all(isinstance(v, (int, np.integer)) for v in padding)
    padding = [int(v) for v in padding]
    if len(padding) == 2:
        padx, pady = padding
        padding = [padx, padx, pady, pady]
    return tuple(padding)


def _get_filter_size(f):
    if f is None:
        return 1, 1
    assert isinstance(f, torch.Tensor) and f.ndim in [1, 2]
    return int(f.shape[-1]), int(f.shape[0])


def setup_filter(f, device=torch.device('cpu'), normalize=True, flip_filter=False, gain=1, separable=None):
    """eg3d/torch_utils/ops/upfirdn2d.py:72-116."""
    if f is None:
        f = 1
    f = torch.as_tensor(f, dtype=torch.float32)
    assert f.ndim in [0, 1, 2] and f.numel() > 0
    if f.ndim == 0:
        f = f[np.newaxis]
    if separable is None:
        separable = (f.ndim == 1 and f.numel() >= 8)
    if f.ndim == 1 and not separable:
        f = f.ger(f)
    assert f.ndim == (1 if separable else 2)
    if normalize:
        f = f / f.sum()
    if flip_filter:
        f = f.flip(list(range(f.ndim)))
    f = f * (gain ** (f.ndim / 2))
    return f.to(device=device)


def _plugin_upfirdn2d(x, f, upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip, gain):
    """The plugin entry point (upfirdn2d.cpp:20); f is rank-2 fp32."""
    if not x.is_cuda:
        raise RuntimeError('x must reside on CUDA device')
    if f.dtype != torch.float32 or f.ndim != 2:
        raise RuntimeError('f must be float32 and rank 2')
    if x.ndim != 4:
        raise RuntimeError('x must be rank 4')
    n, c, ih, iw = x.shape
    fh, fw = f.shape
    ow = (iw * upx + padx0 + padx1 - fw + downx) // downx
    oh = (ih * upy + pady0 + pady1 - fh + downy) // downy
    if ow < 1 or oh < 1:
        raise RuntimeError('output must be at least 1x1')
    fmt = torch.channels_last if (c > 1 and x.stride(1) == 1) else torch.contiguous_format
    y = torch.empty([n, c, oh, ow], dtype=x.dtype, device=x.device, memory_format=fmt)
    if y.numel() == 0:
        return y
    f = f.to(x.device).contiguous()
    with _lib.timed('upfirdn2d', (x.numel() + y.numel()) * x.element_size(), detail=f'up{upx} down{downx} f{fw} {tuple(x.shape)}->{oh}x{ow}'):
        _lib.check(_lib.load().spi_upfirdn2d(
            _lib.ptr(x), _lib.ptr(f), _lib.ptr(y), _lib.dtype_code(x), n, c, ih, iw, _lib.strides4(x), _lib.strides4(y),
            fh, fw, upx, upy, downx, downy, padx0, padx1, pady0, pady1, int(bool(flip)), float(gain), _lib.stream()))
    return y


class _Upfirdn2d(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, f, cfg):
        upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip, gain = cfg
        assert isinstance(x, torch.Tensor) and x.ndim == 4
        if f is None:
            f = torch.ones([1, 1], dtype=torch.float32, device=x.device)
        if f.ndim == 1 and f.shape[0] == 1:
            f = f.square().unsqueeze(0)
        assert f.ndim in [1, 2]
        if f.ndim == 2:
            y = _plugin_upfirdn2d(x, f, upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip, gain)
        else:   # separable: two 1-D passes (upfirdn2d.py:237-239)
            y = _plugin_upfirdn2d(x, f.unsqueeze(0), upx, 1, downx, 1, padx0, padx1, 0, 0, flip, 1.0)
            y = _plugin_upfirdn2d(y, f.unsqueeze(1), 1, upy, 1, downy, 0, 0, pady0, pady1, flip, gain)
        ctx.save_for_backward(f)
        ctx.x_shape, ctx.cfg = x.shape, cfg
        return y

    @staticmethod
    def backward(ctx, dy):
        upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip, gain = ctx.cfg
        f, = ctx.saved_tensors
        _, _, ih, iw = ctx.x_shape
        _, _, oh, ow = dy.shape
        fw, fh = _get_filter_size(f)
        p = (fw - padx0 - 1, iw * upx - ow * downx + padx0 - upx + 1, fh - pady0 - 1, ih * upy - oh * downy + pady0 - upy + 1)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = _Upfirdn2d.apply(dy, f, (downx, downy, upx, upy, *p, not flip, gain))
        assert not ctx.needs_input_grad[1]
        return dx, None, None


def upfirdn2d(x, f, up=1, down=1, padding=0, flip_filter=False, gain=1, impl='cuda'):
    """eg3d/torch_utils/ops/upfirdn2d.py:120-166."""
    assert isinstance(x, torch.Tensor)
    assert impl in ['ref', 'cuda']
    if not x.is_cuda:
        raise RuntimeError('spi_b200.upfirdn2d: x must reside on a CUDA device (no CPU path in this build)')
    upx, upy = _parse_scaling(up)
    downx, downy = _parse_scaling(down)
    cfg = (upx, upy, downx, downy, *_parse_padding(padding), bool(flip_filter), gain)
    return _Upfirdn2d.apply(x, f, cfg)


def filter2d(x, f, padding=0, flip_filter=False, gain=1, impl='cuda'):
    padx0, padx1, pady0, pady1 = _parse_padding(padding)
    fw, fh = _get_filter_size(f)
    p = [padx0 + fw // 2, padx1 + (fw - 1) // 2, pady0 + fh // 2, pady1 + (fh - 1) // 2]
    return upfirdn2d(x, f, padding=p, flip_filter=flip_filter, gain=gain, impl=impl)


def upsample2d(x, f, up=2, padding=0, flip_filter=False, gain=1, impl='cuda'):
    upx, upy = _parse_scaling(up)
    padx0, padx1, pady0, pady1 = _parse_padding(padding)
    fw, fh = _get_filter_size(f)
    p = [padx0 + (fw + upx - 1) // 2, padx1 + (fw - upx) // 2, pady0 + (fh + upy - 1) // 2, pady1 + (fh - upy) // 2]
    return upfirdn2d(x, f, up=up, padding=p, flip_filter=flip_filter, gain=gain * upx * upy, impl=impl)


def downsample2d(x, f, down=2, padding=0, flip_filter=False, gain=1, impl='cuda'):
    downx, downy = _parse_scaling(down)
    padx0, padx1, pady0, pady1 = _parse_padding(padding)
    fw, fh = _get_filter_size(f)
    p = [padx0 + (fw - downx + 1) // 2, padx1 + (fw - downx) // 2, pady0 + (fh - downy + 1) // 2, pady1 + (fh - downy) // 2]
    return upfirdn2d(x, f, down=down, padding=p, flip_filter=flip_filter, gain=gain, impl=impl)
