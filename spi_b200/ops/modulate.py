"""Weight modulation / demodulation (the fused branch of modulated_conv2d, eg3d/training/networks_stylegan2.py:58-68) as
one forward and one backward launch (`spi_modulate_weights*`, spi_b200/csrc/modulate.cu) -- per layer (`modulate_weights`) or for
every modulated convolution of a synthesis network at once (`modulate_bank`).

The per-sample weights are produced directly in the memory layout the conv engine consumes (channels-last OHWI, or IHWO
for the stride-2 transposed convolution), so no layout-conversion copy sits between this op and the convolution."""
import ctypes

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


# ---------------------------------------------------------------------------------------------------------------------
# All modulated convolutions of a network in one launch

MAX_LAYERS = 32
ROW_LIMIT = 40 * 1024          # bytes of one [I x kh*kw] weight row staged in shared memory by the grouped kernels


def _table(ctype, values):
    return (ctype * len(values))(*values)


def _ptrs(tensors):
    return _table(ctypes.c_void_p, [None if t is None else t.data_ptr() for t in tensors])


def _layout_strides(o, i, kh, kw, layout):
    """Strides of the logical [N,O,I,kh,kw] tensor `_alloc` returns for `layout`."""
    k = kh * kw
    layout &= 3
    if layout == 0:
        return (o * i * k, i * k, k, kw, 1)
    if layout == 1:
        return (o * k * i, k * i, 1, kw * i, i)
    return (i * k * o, 1, k * o, kw * o, o)


def _strides_match(t, strides):
    return all(a == b for a, b, sz in zip(t.stride(), strides, t.shape) if sz > 1)


def _transpose_many(src, dst, dims):
    """src[l] memory [g][o][taps][i] -> dst[l] memory [g][i][taps'][o]; dims[l] = (g, o, taps, i, reverse)."""
    cols = list(zip(*dims))
    _lib.check(_lib.load().spi_conv_weight_transpose_many(len(src), _ptrs(src), _ptrs(dst), *[_table(ctypes.c_int, [int(v) for v in c]) for c in cols],
                                                          _lib.stream()))


class _ModulateBank(torch.autograd.Function):
    @staticmethod
    def forward(ctx, meta, *tensors):
        """meta[l] = (demodulate, layout flags, transposed copy: None / 'rev' / 'keep'); tensors = W_0, s_0, W_1, s_1, ...
        Returns the per-sample weights of every layer, then the transposed copies (layers that asked for one, in layer order)."""
        ctx.set_materialize_grads(False)            # layers whose weights go unused hand back None, not a tensor of zeros
        weights = [t.contiguous() for t in tensors[0::2]]
        styles = [t.contiguous() for t in tensors[1::2]]
        dev = weights[0].device
        outs, dcoefs, dims = [], [], []
        for (demod, layout, _), w, s in zip(meta, weights, styles):
            o, i, kh, kw = w.shape
            n = s.shape[0]
            outs.append(_alloc(n, o, i, kh, kw, layout, dev))
            dcoefs.append(torch.empty(n, o, device=dev) if demod else None)
            dims.append((n, o, i, kh * kw, int(demod), layout))
        cols = list(zip(*dims))
        _lib.check(_lib.load().spi_modulate_weights_many(len(meta), _ptrs(weights), _ptrs(styles), _ptrs(outs), _ptrs(dcoefs),
                                                         *[_table(ctypes.c_int, list(c)) for c in cols], _lib.stream()))
        wts, src, tdims = [], [], []
        for (demod, layout, transposed), out, (n, o, i, kk, _, _) in zip(meta, outs, dims):
            if transposed is None:
                continue
            assert layout & 3 == 1                      # memory [n][o][kk][i]
            wts.append(torch.empty(n, i, kk, o, device=dev))
            src.append(out)
            tdims.append((n, o, kk, i, transposed == 'rev'))
        if wts:
            _transpose_many(src, wts, tdims)
            ctx.mark_non_differentiable(*wts)
        ctx.save_for_backward(*weights, *styles, *[d for d in dcoefs if d is not None])
        ctx.meta, ctx.dims = meta, dims
        return tuple(outs) + tuple(wts)

    @staticmethod
    def backward(ctx, *grads):
        meta, dims = ctx.meta, ctx.dims
        count = len(meta)
        saved = list(ctx.saved_tensors)
        weights, styles, rest = saved[:count], saved[count:2 * count], saved[2 * count:]
        dcoefs = [rest.pop(0) if m[0] else None for m in meta]
        dev = weights[0].device
        live = [l for l in range(count) if grads[l] is not None]
        gws = [None] * count
        gss = [None] * count
        if live:
            # gradients handed over in the layout of the stride-2 transposed convolution's weight gradient ([n][i][kk][o]) are re-laid in one
            # grouped launch; anything else that differs from the forward layout takes a plain copy
            gin, src, dst, tdims = {}, [], [], []
            for l in live:
                g = grads[l]
                n, o, i, kk, _, layout = dims[l]
                kh, kw = weights[l].shape[2:]
                if _strides_match(g, _layout_strides(o, i, kh, kw, layout)):
                    gin[l] = g
                    continue
                ref = _alloc(n, o, i, kh, kw, layout, dev)
                if layout & 3 == 1 and _strides_match(g, _layout_strides(o, i, kh, kw, 2)):
                    src.append(g)
                    dst.append(ref)
                    tdims.append((n, i, kk, o, False))           # [n][i][kk][o] -> [n][o][kk][i]
                else:
                    ref.copy_(g)
                gin[l] = ref
            if src:
                _transpose_many(src, dst, tdims)
            need_s = [l for l in live if ctx.needs_input_grad[2 + 2 * l]]
            if need_s:              # accumulated with atomics over the output channels: one zero-filled buffer for all layers
                sizes = [styles[l].numel() for l in need_s]
                flat = zero_arena.take((sum(sizes),))
                if flat is None:
                    flat = torch.zeros(sum(sizes), device=dev)
                off = 0
                for l, sz in zip(need_s, sizes):
                    gss[l] = flat[off:off + sz].view(styles[l].shape)
                    off += sz
            for l in live:
                if ctx.needs_input_grad[1 + 2 * l]:
                    gws[l] = torch.empty_like(weights[l])
            cols = list(zip(*[dims[l] for l in live]))
            pick = lambda seq: [seq[l] for l in live]
            _lib.check(_lib.load().spi_modulate_weights_backward_many(len(live), _ptrs(pick(weights)), _ptrs(pick(styles)), _ptrs(pick(dcoefs)),
                                                                      _ptrs([gin[l] for l in live]), _ptrs(pick(gws)), _ptrs(pick(gss)),
                                                                      *[_table(ctypes.c_int, list(c)) for c in cols], _lib.stream()))
        out = [None]
        for l in range(count):
            out += [gws[l], gss[l]]
        return tuple(out)


def bank_usable(entries):
    """entries as for `modulate_bank`: True when the grouped kernels cover every layer (fp32 CUDA tensors, layouts OIHW / OHWI, rows that fit
    in shared memory)."""
    for weight, styles, demodulate, layout, flip, transposed in entries:
        if not (weight.is_cuda and weight.dtype == torch.float32 and styles.dtype == torch.float32):
            return False
        if LAYOUTS[layout] > 1 or weight.shape[1] * weight.shape[2] * weight.shape[3] * 4 > ROW_LIMIT:
            return False
        if transposed is not None and layout != 'ohwi':
            return False
    return len(entries) > 0


def modulate_bank(entries):
    """entries = [(weight [O,I,kh,kw], styles [N,I], demodulate, layout, flip, transposed), ...] -> [(w5, wT), ...]: `w5` as `modulate_weights`
    returns it, `wT` (when `transposed` is 'rev' or 'keep', else None) the same weights as memory [N][I][taps][O] -- taps reversed for 'rev' --
    which is what the data-gradient convolution of the layer reads (ops/conv.py `_tc2_input_grad`); produced here for all layers by one
    launch instead of one per layer in the backward pass."""
    out = []
    for first in range(0, len(entries), MAX_LAYERS):
        chunk = entries[first:first + MAX_LAYERS]
        meta = tuple((bool(d), LAYOUTS[layout] | (FLIP if flip else 0), tr) for _, _, d, layout, flip, tr in chunk)
        tensors = []
        for weight, styles, *_ in chunk:
            tensors += [weight, styles]
        res = list(_ModulateBank.apply(meta, *tensors))
        w5s, wts = res[:len(chunk)], res[len(chunk):]
        for (_, _, tr), w5 in zip(meta, w5s):
            out.append((w5, wts.pop(0) if tr is not None else None))
    return out
