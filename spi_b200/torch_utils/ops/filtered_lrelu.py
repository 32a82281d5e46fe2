"""Filtered leaky ReLU (drop-in for `torch_utils.ops.filtered_lrelu`, eg3d/torch_utils/ops/filtered_lrelu.py).

bias -> up-FIR -> gain*lrelu*clamp -> down-FIR in one kernel (`spi_filtered_lrelu`, spi_b200/csrc/filtered_lrelu.cu),
with the plugin's 2-bit sign tensor protocol for the backward pass (filtered_lrelu.py:161-274): the gradient is the
same op with up<->down, flipped filters and the sign tensor in read mode.  When the fused kernel reports "no
specialised kernel" (status -2) the generic path upfirdn2d -> `filtered_lrelu_act_` -> upfirdn2d is used, exactly the
reference's fallback (filtered_lrelu.py:225-232).  CUDA only.
"""
import ctypes
import warnings

import numpy as np
import torch

from ... import _lib
from . import upfirdn2d


def _get_filter_size(f):
    if f is None:
        return 1, 1
    assert isinstance(f, torch.Tensor) and 1 <= f.ndim <= 2
    return f.shape[-1], f.shape[0]


def _parse_padding(padding):
    if isinstance(padding, int):
        padding = [padding, padding]
    assert isinstance(padding, (list, tuple)) and all(isinstance(x, (int, np.integer)) for x in padding)
    padding = [int(x) for x in padding]
    if len(padding) == 2:
        px, py = padding
        padding = [px, px, py, py]
    return tuple(padding)


def _as2d(f):
    return f if f.ndim == 2 else torch.outer(f, f)


def _plugin_filtered_lrelu(x, fu, fd, b, si, up, down, px0, px1, py0, py1, sx, sy, gain, slope, clamp, flip_filters, write_signs):
    """Plugin entry point (filtered_lrelu.cpp:20): returns (y, so, return_code)."""
    if not x.is_cuda:
        raise RuntimeError('x must reside on CUDA device')
    if x.dtype not in (torch.float16, torch.float32):
        raise RuntimeError('x and b must be float16 or float32')
    if x.ndim != 4 or x.numel() == 0:
        raise RuntimeError('x must be rank 4 and non-empty')
    if b.ndim != 1 or b.shape[0] != x.shape[1]:
        raise RuntimeError('b must be a vector with the same number of channels as x')
    if up < 1 or down < 1:
        raise RuntimeError('up and down must be at least 1')
    fu2, fd2 = _as2d(fu.float()).contiguous(), _as2d(fd.float()).contiguous()
    n, c, xh, xw = x.shape
    fuh, fuw = fu2.shape
    fdh, fdw = fd2.shape
    cw = xw * up + (px0 + px1) - (fuw - 1)
    ch = xh * up + (py0 + py1) - (fuh - 1)
    if not (cw > fdw - 1 and ch > fdh - 1):
        raise RuntimeError('upsampled buffer must be at least the size of downsampling filter')
    yw = (cw - (fdw - 1) + (down - 1)) // down
    yh = (ch - (fdh - 1) + (down - 1)) // down
    if yw < 1 or yh < 1:
        raise RuntimeError('output must be at least 1x1')
    fmt = torch.channels_last if (c > 1 and x.stride(1) == 1) else torch.contiguous_format
    y = torch.empty([n, c, yh, yw], dtype=x.dtype, device=x.device, memory_format=fmt)
    read_signs = si is not None and si.numel() > 0
    so = None
    s, mode, s_h, s_wb = None, 0, 0, 0
    if write_signs:
        sh, swb = ctypes.c_int(), ctypes.c_int()
        _lib.load().spi_filtered_lrelu_sign_shape(yh, yw, down, fdh, fdw, ctypes.byref(sh), ctypes.byref(swb))
        s = so = torch.zeros([n, c, sh.value, swb.value], dtype=torch.uint8, device=x.device)
        mode, s_h, s_wb = 1, sh.value, swb.value
    elif read_signs:
        if si.dtype != torch.uint8 or si.ndim != 4 or not si.is_contiguous():
            raise RuntimeError('signs must be contiguous uint8 of rank 4')
        s, mode, s_h, s_wb = si, 2, si.shape[2], si.shape[3]
    rc = _lib.load().spi_filtered_lrelu(
        _lib.ptr(x), _lib.ptr(y), _lib.ptr(b.contiguous()), _lib.ptr(s), _lib.ptr(fu2), _lib.ptr(fd2), _lib.dtype_code(x),
        n, c, xh, xw, _lib.strides4(x), _lib.strides4(y), fuh, fuw, fdh, fdw, up, down, px0, px1, py0, py1, s_h, s_wb,
        int(sx), int(sy), float(gain), float(slope), float(clamp), int(bool(flip_filters)), mode, _lib.stream())
    if _lib.check(rc, soft_unsupported=True) == -2:
        return None, None, -1
    return y, so, 0


def _plugin_filtered_lrelu_act_(x, si, sx, sy, gain, slope, clamp, write_signs):
    """Plugin entry point (filtered_lrelu.cpp:217): in-place activation, returns the sign tensor when writing."""
    n, c, h, w = x.shape
    read_signs = si is not None and si.numel() > 0
    so, s, mode, s_h, s_wb = None, None, 0, 0, 0
    if write_signs:
        s_h, s_wb = h, ((w + 15) & ~15) >> 2
        s = so = torch.zeros([n, c, s_h, s_wb], dtype=torch.uint8, device=x.device)
        mode = 1
    elif read_signs:
        s, mode, s_h, s_wb = si, 2, si.shape[2], si.shape[3]
    _lib.check(_lib.load().spi_filtered_lrelu_act(
        _lib.ptr(x), _lib.ptr(s), _lib.dtype_code(x), n, c, h, w, _lib.strides4(x), s_h, s_wb, int(sx), int(sy), float(gain),
        float(slope), float(clamp), mode, _lib.stream()))
    return so


class _FilteredLRelu(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, fu, fd, b, si, sx, sy, cfg):
        up, down, px0, px1, py0, py1, gain, slope, clamp, flip_filter = cfg
        assert isinstance(x, torch.Tensor) and x.ndim == 4
        if fu is None:
            fu = torch.ones([1, 1], dtype=torch.float32, device=x.device)
        if fd is None:
            fd = torch.ones([1, 1], dtype=torch.float32, device=x.device)
        if up == 1 and fu.ndim == 1 and fu.shape[0] == 1:
            fu = fu.square()[None]
        if down == 1 and fd.ndim == 1 and fd.shape[0] == 1:
            fd = fd.square()[None]
        if b is None:
            b = torch.zeros([x.shape[1]], dtype=x.dtype, device=x.device)
        write_signs = (si is None or si.numel() == 0) and (x.requires_grad or b.requires_grad)
        y, so, rc = _plugin_filtered_lrelu(x, fu, fd, b, si, up, down, px0, px1, py0, py1, sx, sy, gain, slope, clamp,
                                           flip_filter, write_signs)
        if rc < 0:
            warnings.warn('filtered_lrelu called with parameters that have no optimized CUDA kernel, using generic fallback',
                          RuntimeWarning)
            y = x.add(b.unsqueeze(-1).unsqueeze(-1))
            y = upfirdn2d.upfirdn2d(x=y, f=fu, up=up, padding=[px0, px1, py0, py1], gain=up ** 2, flip_filter=flip_filter)
            so = _plugin_filtered_lrelu_act_(y, si, sx, sy, gain, slope, clamp, write_signs)
            y = upfirdn2d.upfirdn2d(x=y, f=fd, down=down, flip_filter=flip_filter)
        ctx.save_for_backward(fu, fd, (si if (si is not None and si.numel()) else so))
        ctx.x_shape, ctx.y_shape, ctx.s_ofs, ctx.cfg = x.shape, y.shape, (sx, sy), cfg
        return y

    @staticmethod
    def backward(ctx, dy):
        up, down, px0, px1, py0, py1, gain, slope, clamp, flip_filter = ctx.cfg
        fu, fd, si = ctx.saved_tensors
        _, _, xh, xw = ctx.x_shape
        _, _, yh, yw = ctx.y_shape
        sx, sy = ctx.s_ofs
        dx = db = None
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[3]:
            pp = ((fu.shape[-1] - 1) + (fd.shape[-1] - 1) - px0, xw * up - yw * down + px0 - (up - 1),
                  (fu.shape[0] - 1) + (fd.shape[0] - 1) - py0, xh * up - yh * down + py0 - (up - 1))
            gg = gain * (up ** 2) / (down ** 2)
            sx = sx - (fu.shape[-1] - 1) + px0
            sy = sy - (fu.shape[0] - 1) + py0
            dx = _FilteredLRelu.apply(dy, fd, fu, None, si, sx, sy, (down, up, *pp, gg, slope, float('inf'), not flip_filter))
        if ctx.needs_input_grad[3]:
            db = dx.sum([0, 2, 3])
        return dx, None, None, db, None, None, None, None


def filtered_lrelu(x, fu=None, fd=None, b=None, up=1, down=1, padding=0, gain=np.sqrt(2), slope=0.2, clamp=None,
                   flip_filter=False, impl='cuda'):
    """eg3d/torch_utils/ops/filtered_lrelu.py:58-120."""
    assert isinstance(x, torch.Tensor)
    assert impl in ['ref', 'cuda']
    assert isinstance(up, int) and up >= 1 and isinstance(down, int) and down >= 1
    assert gain == float(gain) and gain > 0 and slope == float(slope) and slope >= 0
    assert clamp is None or (clamp == float(clamp) and clamp >= 0)
    if not x.is_cuda:
        raise RuntimeError('spi_b200.filtered_lrelu: x must reside on a CUDA device (no CPU path in this build)')
    cfg = (up, down, *_parse_padding(padding), float(gain), float(slope), float(clamp if clamp is not None else 'inf'),
           bool(flip_filter))
    return _FilteredLRelu.apply(x, fu, fd, b, None, 0, 0, cfg)
