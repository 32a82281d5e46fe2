"""Oracle (CPU, test infrastructure): the three L1 operators of `eg3d/torch_utils/ops`.

Restates the `impl='ref'` twins the reference itself falls back to on CPU:
  bias_act        eg3d/torch_utils/ops/bias_act.py:93-123
  upfirdn2d       eg3d/torch_utils/ops/upfirdn2d.py:169-211 (+ setup_filter :72-116,
                  upsample2d :315-350, filter padding algebra)
  filtered_lrelu  eg3d/torch_utils/ops/filtered_lrelu.py:123-153
  conv2d_resample eg3d/torch_utils/ops/conv2d_resample.py:48-143 (the branches the generator hits)
"""
import math

import torch
import torch.nn.functional as F

# name -> (fn, default alpha, default gain); ordering follows the plugin's act index 1..9
# (bias_act.py:23-33).
ACTS = {
    'linear':   (lambda x, a: x, 0.0, 1.0),
    'relu':     (lambda x, a: F.relu(x), 0.0, math.sqrt(2)),
    'lrelu':    (lambda x, a: F.leaky_relu(x, a), 0.2, math.sqrt(2)),
    'tanh':     (lambda x, a: torch.tanh(x), 0.0, 1.0),
    'sigmoid':  (lambda x, a: torch.sigmoid(x), 0.0, 1.0),
    'elu':      (lambda x, a: F.elu(x), 0.0, 1.0),
    'selu':     (lambda x, a: F.selu(x), 0.0, 1.0),
    'softplus': (lambda x, a: F.softplus(x), 0.0, 1.0),
    'swish':    (lambda x, a: torch.sigmoid(x) * x, 0.0, math.sqrt(2)),
}
ACT_INDEX = {name: i + 1 for i, name in enumerate(ACTS)}


def bias_act(x, b=None, dim=1, act='linear', alpha=None, gain=None, clamp=None):
    """y = clamp(act(x + b) * gain)  (bias_act.py:93-123)."""
    fn, def_alpha, def_gain = ACTS[act]
    alpha = def_alpha if alpha is None else float(alpha)
    gain = def_gain if gain is None else float(gain)
    if b is not None:
        shape = [1] * x.ndim
        shape[dim] = -1
        x = x + b.reshape(shape)
    x = fn(x, alpha)
    if gain != 1:
        x = x * gain
    if clamp is not None and clamp >= 0:
        x = x.clamp(-clamp, clamp)
    return x


def setup_filter(taps=(1, 3, 3, 1), gain=1.0, flip=False):
    """1-D taps (< 8) -> normalised outer product (upfirdn2d.py:72-116)."""
    f = torch.as_tensor(taps, dtype=torch.float32)
    if f.ndim == 0:
        f = f[None]
    if f.ndim == 1 and f.numel() < 8:
        f = torch.outer(f, f)
    f = f / f.sum()
    if flip:
        f = f.flip(list(range(f.ndim)))
    return f * (gain ** (f.ndim / 2))


def _pad4(p):
    if isinstance(p, int):
        p = [p, p]
    p = [int(v) for v in p]
    if len(p) == 2:
        p = [p[0], p[0], p[1], p[1]]
    return p


def _two(v):
    return (v, v) if isinstance(v, int) else (int(v[0]), int(v[1]))


def upfirdn2d(x, f, up=1, down=1, padding=0, flip_filter=False, gain=1.0):
    """Zero-insert upsample, pad/crop, FIR, decimate (upfirdn2d.py:169-211)."""
    n, c, h, w = x.shape
    ux, uy = _two(up)
    dx, dy = _two(down)
    px0, px1, py0, py1 = _pad4(padding)
    if f is None:
        f = torch.ones(1, 1)
    y = x.new_zeros(n, c, h * uy, w * ux)
    y[:, :, ::uy, ::ux] = x
    y = F.pad(y, [max(px0, 0), max(px1, 0), max(py0, 0), max(py1, 0)])
    y = y[:, :, max(-py0, 0): y.shape[2] - max(-py1, 0), max(-px0, 0): y.shape[3] - max(-px1, 0)]
    k = (f * (gain ** (f.ndim / 2))).to(x.dtype)
    if not flip_filter:
        k = k.flip(list(range(k.ndim)))
    if k.ndim == 2:
        y = F.conv2d(y, k[None, None].repeat(c, 1, 1, 1), groups=c)
    else:
        y = F.conv2d(y, k[None, None, None, :].repeat(c, 1, 1, 1), groups=c)
        y = F.conv2d(y, k[None, None, :, None].repeat(c, 1, 1, 1), groups=c)
    return y[:, :, ::dy, ::dx]


def upsample2d(x, f, up=2, padding=0, flip_filter=False, gain=1.0):
    """upfirdn2d.py:315-350."""
    ux, uy = _two(up)
    px0, px1, py0, py1 = _pad4(padding)
    fw, fh = f.shape[-1], f.shape[0]
    p = [px0 + (fw + ux - 1) // 2, px1 + (fw - ux) // 2, py0 + (fh + uy - 1) // 2, py1 + (fh - uy) // 2]
    return upfirdn2d(x, f, up=up, padding=p, flip_filter=flip_filter, gain=gain * ux * uy)


def downsample2d(x, f, down=2, padding=0, flip_filter=False, gain=1.0):
    """upfirdn2d.py:353-391."""
    dx, dy = _two(down)
    px0, px1, py0, py1 = _pad4(padding)
    fw, fh = f.shape[-1], f.shape[0]
    p = [px0 + (fw - dx + 1) // 2, px1 + (fw - dx) // 2, py0 + (fh - dy + 1) // 2, py1 + (fh - dy) // 2]
    return upfirdn2d(x, f, down=down, padding=p, flip_filter=flip_filter, gain=gain)


def filter2d(x, f, padding=0, flip_filter=False, gain=1.0):
    """upfirdn2d.py:279-312."""
    px0, px1, py0, py1 = _pad4(padding)
    fw, fh = f.shape[-1], f.shape[0]
    p = [px0 + fw // 2, px1 + (fw - 1) // 2, py0 + fh // 2, py1 + (fh - 1) // 2]
    return upfirdn2d(x, f, padding=p, flip_filter=flip_filter, gain=gain)


def filtered_lrelu(x, fu=None, fd=None, b=None, up=1, down=1, padding=0, gain=math.sqrt(2), slope=0.2,
                   clamp=None, flip_filter=False):
    """bias -> up-FIR (gain up^2) -> gain*lrelu, clamp -> down-FIR (filtered_lrelu.py:123-153)."""
    x = bias_act(x, b)
    x = upfirdn2d(x, fu, up=up, padding=padding, gain=up ** 2, flip_filter=flip_filter)
    x = bias_act(x, act='lrelu', alpha=slope, gain=gain, clamp=clamp)
    return upfirdn2d(x, fd, down=down, flip_filter=flip_filter)


def conv2d_resample(x, w, f=None, up=1, padding=0, groups=1, flip_weight=True):
    """The two branches of conv2d_resample.py the FFHQ generator executes:
    up=1 -> plain correlation (:135-137); up=2 -> stride-2 transposed conv then FIR with gain up^2
    (:113-131).  `w` is [O, I/groups, kh, kw]."""
    o, ig, kh, kw = w.shape
    px0, px1, py0, py1 = _pad4(padding)
    if up == 1:
        assert px0 == px1 and py0 == py1
        if not flip_weight:
            w = w.flip([2, 3])
        return F.conv2d(x, w, padding=[py0, px0], groups=groups)
    fw, fh = f.shape[-1], f.shape[0]
    px0 += (fw + up - 1) // 2 - (kw - 1)
    px1 += (fw - up) // 2 - (kw - up)
    py0 += (fh + up - 1) // 2 - (kh - 1)
    py1 += (fh - up) // 2 - (kh - up)
    pxt = max(min(-px0, -px1), 0)
    pyt = max(min(-py0, -py1), 0)
    wt = w.reshape(groups, o // groups, ig, kh, kw).transpose(1, 2).reshape(groups * ig, o // groups, kh, kw)
    if flip_weight:  # wrapper flips when asked for a *convolution* (`not flip_weight` swap at :126)
        wt = wt.flip([2, 3])
    x = F.conv_transpose2d(x, wt, stride=up, padding=[pyt, pxt], groups=groups)
    return upfirdn2d(x, f, padding=[px0 + pxt, px1 + pxt, py0 + pyt, py1 + pyt], gain=up ** 2)
