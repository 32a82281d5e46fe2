"""2-D convolution with optional up-sampling (drop-in for `torch_utils.ops.conv2d_resample`,
eg3d/torch_utils/ops/conv2d_resample.py:48-143).

The padding algebra is the reference's; the dense contraction itself is dispatched to the conv engine in
`spi_b200.ops.conv` (tensor-core implicit GEMM) and every FIR pass to `spi_upfirdn2d`.
"""
import torch

from . import upfirdn2d
from ...ops import conv as conv_engine


def _get_weight_shape(w):
    return [int(s) for s in w.shape]


def conv2d_resample(x, w, f=None, up=1, down=1, padding=0, groups=1, flip_weight=True, flip_filter=False):
    assert isinstance(x, torch.Tensor) and x.ndim == 4
    assert isinstance(w, torch.Tensor) and w.ndim == 4 and w.dtype == x.dtype
    assert f is None or (isinstance(f, torch.Tensor) and f.ndim in [1, 2] and f.dtype == torch.float32)
    assert isinstance(up, int) and up >= 1 and isinstance(down, int) and down >= 1 and isinstance(groups, int) and groups >= 1
    out_channels, in_channels_per_group, kh, kw = _get_weight_shape(w)
    fw, fh = upfirdn2d._get_filter_size(f)
    px0, px1, py0, py1 = upfirdn2d._parse_padding(padding)

    # padding bookkeeping (conv2d_resample.py:77-87)
    if up > 1:
        px0 += (fw + up - 1) // 2; px1 += (fw - up) // 2; py0 += (fh + up - 1) // 2; py1 += (fh - up) // 2
    if down > 1:
        px0 += (fw - down + 1) // 2; px1 += (fw - down) // 2; py0 += (fh - down + 1) // 2; py1 += (fh - down) // 2

    if kw == 1 and kh == 1 and (down > 1 and up == 1):       # 1x1 + downsample: filter first (:89-93)
        x = upfirdn2d.upfirdn2d(x=x, f=f, down=down, padding=[px0, px1, py0, py1], flip_filter=flip_filter)
        return conv_engine.conv2d(x, w, groups=groups, flip_weight=flip_weight)
    if kw == 1 and kh == 1 and (up > 1 and down == 1):       # 1x1 + upsample: convolve first (:95-99)
        x = conv_engine.conv2d(x, w, groups=groups, flip_weight=flip_weight)
        return upfirdn2d.upfirdn2d(x=x, f=f, up=up, padding=[px0, px1, py0, py1], gain=up ** 2, flip_filter=flip_filter)
    if down > 1 and up == 1:                                 # strided conv (:101-105)
        x = upfirdn2d.upfirdn2d(x=x, f=f, padding=[px0, px1, py0, py1], flip_filter=flip_filter)
        return conv_engine.conv2d(x, w, stride=down, groups=groups, flip_weight=flip_weight)
    if up > 1:                                               # transposed strided conv + FIR (:107-126)
        if groups == 1:
            w = w.transpose(0, 1)
        else:
            w = w.reshape(groups, out_channels // groups, in_channels_per_group, kh, kw).transpose(1, 2)
            w = w.reshape(groups * in_channels_per_group, out_channels // groups, kh, kw)
        px0 -= kw - 1; px1 -= kw - up; py0 -= kh - 1; py1 -= kh - up
        pxt = max(min(-px0, -px1), 0)
        pyt = max(min(-py0, -py1), 0)
        x = conv_engine.conv2d(x, w, stride=up, padding=[pyt, pxt], groups=groups, transpose=True, flip_weight=(not flip_weight))
        x = upfirdn2d.upfirdn2d(x=x, f=f, padding=[px0 + pxt, px1 + pxt, py0 + pyt, py1 + pyt], gain=up ** 2, flip_filter=flip_filter)
        if down > 1:
            x = upfirdn2d.upfirdn2d(x=x, f=f, down=down, flip_filter=flip_filter)
        return x
    if up == 1 and down == 1 and px0 == px1 and py0 == py1 and px0 >= 0 and py0 >= 0:   # plain conv (:128-131)
        return conv_engine.conv2d(x, w, padding=[py0, px0], groups=groups, flip_weight=flip_weight)
    # generic fallback (:133-138)
    x = upfirdn2d.upfirdn2d(x=x, f=(f if up > 1 else None), up=up, padding=[px0, px1, py0, py1], gain=up ** 2, flip_filter=flip_filter)
    x = conv_engine.conv2d(x, w, groups=groups, flip_weight=flip_weight)
    if down > 1:
        x = upfirdn2d.upfirdn2d(x=x, f=f, down=down, flip_filter=flip_filter)
    return x
