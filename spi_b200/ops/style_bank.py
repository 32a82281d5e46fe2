"""Every affine layer of a synthesis network in one launch (`spi_style_bank_*`, spi_b200/csrc/modulate.cu).

Each SynthesisLayer / ToRGBLayer of the reference maps its latent through `FullyConnectedLayer(w_dim, in_channels, bias_init=1)`
(eg3d/training/networks_stylegan2.py:282,316 and :352,357-358): 20 matrix-vector products in the backbone, 6 in the super-resolution
module, each a few microseconds of work wrapped in five launches per iteration (addmm, the ToRGB gain, and three backward launches).
The bank evaluates them together -- one forward launch, one backward launch (plus one for the latent gradient while the latent is
being optimised) -- and autograd sees a single node whose backward runs once every style gradient has arrived.
"""
import ctypes

import torch

from .. import _lib

MAX_LAYERS = 32


def _table(ctype, values):
    return (ctype * len(values))(*values)


def _ptrs(tensors):
    return _table(ctypes.c_void_p, [None if t is None else t.data_ptr() for t in tensors])


class _StyleBank(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ws, meta, *params):
        """ws [N, L, K] (last dim contiguous); meta = tuple of (in_features, ws index, weight gain, bias gain, output gain); params = W_0, b_0,
        W_1, b_1, ... (b_l may be None)."""
        ctx.set_materialize_grads(False)            # unused styles hand back None, not zeros
        n, num_ws, k = ws.shape
        weights = [p.contiguous() for p in params[0::2]]
        biases = [None if p is None else p.contiguous() for p in params[1::2]]
        outs = [torch.empty(n, m[0], device=ws.device, dtype=torch.float32) for m in meta]
        tabs = (_table(ctypes.c_int, [m[0] for m in meta]), _table(ctypes.c_int, [m[1] for m in meta]), _table(ctypes.c_float, [m[2] for m in meta]),
                _table(ctypes.c_float, [m[3] for m in meta]), _table(ctypes.c_float, [m[4] for m in meta]))
        _lib.check(_lib.load().spi_style_bank_forward(_lib.ptr(ws), ws.stride(0), ws.stride(1), n, k, len(meta), _ptrs(weights), _ptrs(biases), _ptrs(outs),
                                                      *tabs, _lib.stream()))
        ctx.save_for_backward(ws, *weights)
        ctx.meta, ctx.has_bias = meta, [b is not None for b in biases]
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gs):
        ws, *weights = ctx.saved_tensors
        meta = ctx.meta
        n, num_ws, k = ws.shape
        gs = [None if g is None else g.contiguous() for g in gs]
        need = ctx.needs_input_grad
        dW = [torch.empty_like(w) if need[2 + 2 * l] else None for l, w in enumerate(weights)]
        db = [torch.empty(meta[l][0], device=ws.device, dtype=torch.float32) if (ctx.has_bias[l] and need[3 + 2 * l]) else None for l in range(len(meta))]
        dws = torch.empty(n, num_ws, k, device=ws.device, dtype=torch.float32) if need[0] else None
        tabs = (_table(ctypes.c_int, [m[0] for m in meta]), _table(ctypes.c_int, [m[1] for m in meta]), _table(ctypes.c_float, [m[2] for m in meta]),
                _table(ctypes.c_float, [m[3] for m in meta]), _table(ctypes.c_float, [m[4] for m in meta]))
        _lib.check(_lib.load().spi_style_bank_backward(_lib.ptr(ws), ws.stride(0), ws.stride(1), n, k, len(meta), _ptrs(weights), _ptrs(gs), _ptrs(dW), _ptrs(db),
                                                       *tabs, _lib.ptr(dws), num_ws, _lib.stream()))
        grads = [dws, None]
        for l in range(len(meta)):
            grads += [dW[l], db[l]]
        return tuple(grads)


def usable(ws):
    return ws.is_cuda and ws.dtype == torch.float32 and ws.ndim == 3 and ws.stride(2) == 1 and ws.shape[2] % 4 == 0


def style_bank(ws, entries):
    """ws [N, L, K]; entries = [(affine FullyConnectedLayer, index into L, output gain), ...]  ->  [styles_l [N', in_features_l], ...], with
    N' = 1 when `ws` is one latent broadcast over the batch (stride 0: the layers then build one weight set for the whole batch, as
    SynthesisLayer.forward does on its own)."""
    if not usable(ws):
        raise RuntimeError('spi_b200.style_bank: latents must be a float32 CUDA tensor [N, L, K] with a contiguous last dimension (no CPU path)')
    if ws.shape[0] > 1 and ws.stride(0) == 0:
        ws = ws[:1]
    if any((s * 4) % 16 for s in ws.stride()[:2]) or ws.data_ptr() % 16:
        ws = ws.contiguous()
    out = []
    for first in range(0, len(entries), MAX_LAYERS):
        chunk = entries[first:first + MAX_LAYERS]
        meta, params = [], []
        for fc, idx, gain in chunk:
            assert fc.activation == 'linear' and fc.in_features == ws.shape[2]
            meta.append((fc.out_features, int(idx), float(fc.weight_gain), float(fc.bias_gain), float(gain)))
            params += [fc.weight, fc.bias]
        out += list(_StyleBank.apply(ws, tuple(meta), *params))
    return out
