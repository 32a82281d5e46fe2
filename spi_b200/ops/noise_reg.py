"""Noise-buffer regulariser and re-normalisation of the stage-1 projectors (spi/training/projectors/mirror_projector.py:107-115,
128-131; identical loops in w_projector.py and w_plus_projector.py) as one launch each for all buffers
(`spi_noise_reg_forward/backward`, `spi_noise_renorm` in spi_b200/csrc/noise_reg.cu) instead of ~150 ATen launches per step."""
import torch

from .. import _lib

_MAXLEV = 8


class _Table:
    """Device table {x ptr, gradient offset, size} of a fixed set of square buffers (rebuilt if a buffer moved)."""

    def __init__(self, bufs):
        for b in bufs:
            if not b.is_cuda:
                raise RuntimeError('spi_b200.noise_reg: buffers must reside on a CUDA device (no CPU path in this build)')
            assert b.ndim == 2 and b.shape[0] == b.shape[1] and b.dtype == torch.float32 and b.is_contiguous()
            s = b.shape[0]
            assert s <= 256 and (s & (s - 1)) == 0, 'noise buffers are power-of-two squares up to 256'
        self.key = tuple(b.data_ptr() for b in bufs)
        self.sizes = [b.shape[0] for b in bufs]
        self.offsets, off = [], 0
        for s in self.sizes:
            self.offsets.append(off)
            off += s * s
        self.total = off
        rows = [[b.data_ptr(), o, s] for b, o, s in zip(bufs, self.offsets, self.sizes)]
        self.table = torch.tensor(rows, dtype=torch.int64).to(bufs[0].device)
        self.count, self.max_size = len(bufs), max(self.sizes)


_tables = {}


def _table(bufs):
    key = tuple(b.data_ptr() for b in bufs)
    t = _tables.get(key)
    if t is None:
        if len(_tables) > 64:
            _tables.clear()
        t = _tables[key] = _Table(bufs)
    return t


class _NoiseReg(torch.autograd.Function):
    @staticmethod
    def forward(ctx, *bufs):
        t = _table(bufs)
        dev = bufs[0].device
        partial = torch.empty(t.count, device=dev)
        stats = torch.empty(t.count * _MAXLEV * 2, device=dev)
        _lib.check(_lib.load().spi_noise_reg_forward(_lib.ptr(t.table), t.count, t.max_size, _lib.ptr(partial), _lib.ptr(stats), _lib.stream()))
        ctx.t, ctx.stats = t, stats
        return partial.sum()

    @staticmethod
    def backward(ctx, gout):
        t = ctx.t
        g = torch.empty(t.total, device=gout.device)
        gout = gout.contiguous().float()
        _lib.check(_lib.load().spi_noise_reg_backward(_lib.ptr(t.table), t.count, t.max_size, _lib.ptr(ctx.stats), _lib.ptr(gout), _lib.ptr(g),
                                                      _lib.stream()))
        return tuple(g[o:o + s * s].view(s, s) for o, s in zip(t.offsets, t.sizes))


def noise_regulariser(noise_bufs):
    """Sum over buffers and pyramid levels of mean(x*roll(x,1,W))^2 + mean(x*roll(x,1,H))^2 (differentiable)."""
    return _NoiseReg.apply(*list(noise_bufs))


@torch.no_grad()
def renormalise_noise_(noise_bufs):
    """buf -= buf.mean(); buf *= buf.square().mean().rsqrt() for every buffer, in place, one launch."""
    bufs = list(noise_bufs)
    t = _table(bufs)
    _lib.check(_lib.load().spi_noise_renorm(_lib.ptr(t.table), t.count, _lib.stream()))
