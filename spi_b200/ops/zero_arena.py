"""Per-iteration pool of zero-initialised device memory.

Several kernels of an iteration accumulate into their output (the Cin-split convolutions of the small feature maps and every weight
gradient reduce-add partial sums through TMA; the modulation backward adds style gradients with atomics), so their outputs must be
zero on entry -- about a hundred `cudaMemsetAsync` nodes of a few microseconds each per iteration.  An iteration that opens with
`begin(key)` gets ONE zero-filled buffer instead, sized from what the previous iteration of the same kind (`key`) asked for;
`take(shape)` hands out 1 KiB-aligned slices of it and the entry points are told (a flag bit, include/spi_b200.h) to skip their own
fill.  Outside `begin()` / `end()`, or when the buffer is exhausted, `take` returns None and the callee zero-fills as before -- the
first iteration of each kind runs that way and records the size.

The buffer is an ordinary tensor: slices keep it alive, it is released when the last of them dies, and inside a CUDA-graph capture
it lives in the graph's pool like every other temporary."""
import contextlib

import torch

_ALIGN = 1024
_hint = {}              # key -> bytes requested by the last iteration of that kind
_state = None           # [key, buffer (uint8) or None, offset, bytes requested]
ENABLED = True


def begin(key, device):
    global _state
    want = _hint.get(key, 0) if ENABLED else 0
    buf = torch.zeros(want, dtype=torch.uint8, device=device) if want > 0 else None
    _state = [key, buf, 0, 0]


def end():
    global _state
    if _state is not None:
        _hint[_state[0]] = _state[3]
        _state = None


@contextlib.contextmanager
def iteration(key, device):
    """`with zero_arena.iteration(key, device): <forward + backward of one iteration>`"""
    begin(key, device)
    try:
        yield
    finally:
        end()


def active():
    """An iteration is open and has a buffer."""
    return _state is not None and _state[1] is not None


def recording():
    """An iteration is open: `take` either serves the request or (no buffer yet / no room) records its size for the next iteration."""
    return _state is not None and ENABLED


def take(shape, channels_last=False):
    """A zero float32 tensor of `shape` carved from the iteration's buffer (channels-last memory order for a 4-D NCHW `shape` when
    asked), or None when there is no buffer / no room."""
    if _state is None or not ENABLED:
        return None
    numel = 1
    for s in shape:
        numel *= int(s)
    nbytes = (numel * 4 + _ALIGN - 1) // _ALIGN * _ALIGN
    _state[3] += nbytes
    buf, off = _state[1], _state[2]
    if buf is None or off + nbytes > buf.numel():
        return None
    _state[2] = off + nbytes
    flat = buf[off:off + numel * 4].view(torch.float32)
    if channels_last:
        n, c, h, w = shape
        return flat.view(n, h, w, c).permute(0, 3, 1, 2)
    return flat.view(*shape)


def reset():
    global _state
    _state = None
    _hint.clear()
