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
