"""Weight modulation / demodulation (the fused branch of modulated_conv2d, eg3d/training/networks_stylegan2.py:58-68) as
one forward and one backward launch (`spi_modulate_weights*`, spi_b200/csrc/modulate.cu).

The per-sample weights are produced directly in the memory layout the conv engine consumes (channels-last OHWI, or IHWO
for the stride-2 transposed convolution), so no layout-conversion copy sits between this op and the convolution."""
import torch

from .. import _lib
from . import zero_arena

LAYOUTS = {'oihw': 0, 'ohwi': 1, 'ihwo': 2}
FLIP = 4          # + FLIP: the taps are written reversed (the result equals w.flip([3, 4])), see modulate.cu
PREZEROED = 8     # backward only: grad_styles handed in is already zero (ops/zero_arena.py)


def _alloc(n, o, i, kh, kw, layout, device):
    """Logical [N,O,I,kh,kw] tensor whose memory order is the requested layout."""
    layout &= 3
    if layout == 0:
        return torch.empty(n, o, i, kh, kw, device=device)
    if layout == 1:
        return torch.empty(n, o, kh, kw, i, device=device).permute(0, 1, 4, 2, 3)
    return torch.empty(n, i, kh, kw, o, device=device).permute(0, 4, 1, 2, 3)


class _Modulate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, weight, styles, demodulate, layout):
        o, i, kh, kw = weight.shape
        n = styles.shape[0]
        weight, styles = weight.contiguous(), styles.contiguous()
        out = _alloc(n, o, i, kh, kw, layout, weight.device)
        dcoef = torch.empty(n, o, device=weight.device) if demodulate else None
        _lib.check(_lib.load().spi_modulate_weights(_lib.ptr(weight), _lib.ptr(styles), _lib.ptr(out), _lib.ptr(dcoef), n, o, i, kh * kw,
                                                    int(demodulate), layout, _lib.stream()))
        ctx.save_for_backward(weight, styles, dcoef)
        ctx.demodulate, ctx.layout = demodulate, layout
        return out

    @staticmethod
    def backward(ctx, g):
        weight, styles, dcoef = ctx.saved_tensors
        o, i, kh, kw = weight.shape
        n = styles.shape[0]
        ref = _alloc(n, o, i, kh, kw, ctx.layout, weight.device)
        same = all(gs == rs for gs, rs, sz in zip(g.stride(), ref.stride(), g.shape) if sz > 1)
        if not same:                            # re-lay the incoming gradient only if autograd handed it over differently
            ref.copy_(g)
            g = ref
        gw = torch.empty_like(weight) if ctx.needs_input_grad[0] else None
        gs = zeroed = None
        if ctx.needs_input_grad[1]:           # accumulated with atomics over the output channels: zero on entry
            gs = zeroed = zero_arena.take(styles.shape)
            if gs is None:
                gs = torch.empty_like(styles)
        _lib.check(_lib.load().spi_modulate_weights_backward(_lib.ptr(weight), _lib.ptr(styles), _lib.ptr(dcoef), _lib.ptr(g),
                                                             _lib.ptr(gw), _lib.ptr(gs), n, o, i, kh * kw, int(ctx.demodulate),
                                                             ctx.layout | (PREZEROED if zeroed is not None else 0), _lib.stream()))
        return gw, gs, None, None


def modulate_weights(weight, styles, demodulate=True, layout='oihw', flip=False):
    """weight [O,I,kh,kw], styles [N,I] -> per-sample weights, logical shape [N,O,I,kh,kw], memory order `layout`; `flip=True`
    returns them with the taps reversed (what `_conv2d_wrapper(..., flip_weight=False)` would build with `w.flip([2, 3])`,
    conv2d_resample.py:38-40) at no extra pass."""
    if not weight.is_cuda:
        raise RuntimeError('spi_b200.modulate_weights: tensors must reside on a CUDA device (no CPU path in this build)')
    return _Modulate.apply(weight.float(), styles.float(), bool(demodulate), LAYOUTS[layout] | (FLIP if flip else 0))
