"""Weight modulation / demodulation (the fused branch of modulated_conv2d, eg3d/training/networks_stylegan2.py:58-68) as
one forward and one backward launch (`spi_modulate_weights*`, spi_b200/csrc/modulate.cu)."""
import torch

from .. import _lib


class _Modulate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, weight, styles, demodulate):
        o, i, kh, kw = weight.shape
        n = styles.shape[0]
        weight, styles = weight.contiguous(), styles.contiguous()
        out = torch.empty(n, o, i, kh, kw, device=weight.device)
        dcoef = torch.empty(n, o, device=weight.device) if demodulate else None
        _lib.check(_lib.load().spi_modulate_weights(_lib.ptr(weight), _lib.ptr(styles), _lib.ptr(out), _lib.ptr(dcoef), n, o, i, kh * kw,
                                                    int(demodulate), _lib.stream()))
        ctx.save_for_backward(weight, styles, dcoef)
        ctx.demodulate = demodulate
        return out

    @staticmethod
    def backward(ctx, g):
        weight, styles, dcoef = ctx.saved_tensors
        o, i, kh, kw = weight.shape
        n = styles.shape[0]
        gw = torch.empty_like(weight) if ctx.needs_input_grad[0] else None
        gs = torch.empty_like(styles) if ctx.needs_input_grad[1] else None
        _lib.check(_lib.load().spi_modulate_weights_backward(_lib.ptr(weight), _lib.ptr(styles), _lib.ptr(dcoef), _lib.ptr(g.contiguous()),
                                                             _lib.ptr(gw), _lib.ptr(gs), n, o, i, kh * kw, int(ctx.demodulate), _lib.stream()))
        return gw, gs, None


def modulate_weights(weight, styles, demodulate=True):
    """weight [O,I,kh,kw], styles [N,I] -> per-sample weights [N,O,I,kh,kw]."""
    if not weight.is_cuda:
        raise RuntimeError('spi_b200.modulate_weights: tensors must reside on a CUDA device (no CPU path in this build)')
    return _Modulate.apply(weight.float(), styles.float(), bool(demodulate))
