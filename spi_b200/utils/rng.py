"""Random draws of the inversion loop, with an injection queue for parity tests.

Product code draws from torch's CUDA generator with the reference's shapes and order.  Tests push the oracle's tensors
with `inject(...)`; a queued tensor is consumed (and shape-checked) instead of drawing.
"""
import torch

_queue = []


def inject(*tensors):
    _queue.extend(tensors)


def pending():
    return len(_queue)


def _take(shape, device):
    t = _queue.pop(0)
    assert tuple(t.shape) == tuple(shape), f'injected random tensor has shape {tuple(t.shape)}, expected {tuple(shape)}'
    return t.to(device=device, dtype=torch.float32)


def rand(shape, device):
    return _take(shape, device) if _queue else torch.rand(shape, device=device)


def randn_like(t):
    return _take(t.shape, t.device) if _queue else torch.randn_like(t)
