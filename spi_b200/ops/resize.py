"""Exact half-size resize (bilinear == area == 2x2 mean at scale 1/2, align_corners=False) through `spi_downsample2x`
(spi_b200/csrc/optim.cu).  Replaces F.interpolate at lpips.py:38-39, bbox_cx_loss.py:161-163, w_projector.py:50,83."""
import torch

from .. import _lib


class _Half(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        n, c, h, w = x.shape
        assert h % 2 == 0 and w % 2 == 0
        x = x.contiguous()
        y = torch.empty(n, c, h // 2, w // 2, device=x.device)
        _lib.check(_lib.load().spi_downsample2x(_lib.ptr(x), _lib.ptr(y), n * c, h // 2, w // 2, 0, _lib.stream()))
        return y

    @staticmethod
    def backward(ctx, gy):
        n, c, oh, ow = gy.shape
        gy = gy.contiguous()
        gx = torch.empty(n, c, 2 * oh, 2 * ow, device=gy.device)
        _lib.check(_lib.load().spi_downsample2x(_lib.ptr(gy), _lib.ptr(gx), n * c, oh, ow, 1, _lib.stream()))
        return gx


def downsample2x(x):
    if not x.is_cuda:
        raise RuntimeError('spi_b200.downsample2x: x must reside on a CUDA device (no CPU path in this build)')
    return _Half.apply(x.float())


class _MaxPool2x2(torch.autograd.Function):
    """nn.MaxPool2d(2, 2) on channels-last fp32 activations (`spi_maxpool2x2`): no index tensor, the backward pass re-derives the
    arg-max from the saved input."""

    @staticmethod
    def forward(ctx, x):
        n, c, h, w = x.shape
        x = x.contiguous(memory_format=torch.channels_last)
        y = torch.empty(n, c, h // 2, w // 2, device=x.device, dtype=x.dtype, memory_format=torch.channels_last)
        _lib.check(_lib.load().spi_maxpool2x2(_lib.ptr(x), None, _lib.ptr(y), n, h, w, c, 0, _lib.stream()))
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, = ctx.saved_tensors
        n, c, h, w = x.shape
        dy = dy.contiguous(memory_format=torch.channels_last)
        dx = torch.empty_like(x)
        _lib.check(_lib.load().spi_maxpool2x2(_lib.ptr(x), _lib.ptr(dy), _lib.ptr(dx), n, h, w, c, 1, _lib.stream()))
        return dx


def maxpool2x2(x):
    """MaxPool2d(kernel_size=2, stride=2); falls back to ATen for shapes the kernel does not take (odd sizes, C % 4 != 0)."""
    n, c, h, w = x.shape
    if x.is_cuda and x.dtype == torch.float32 and c % 4 == 0 and h % 2 == 0 and w % 2 == 0 and h >= 2 and w >= 2:
        return _MaxPool2x2.apply(x)
    return torch.nn.functional.max_pool2d(x, 2, 2)
