"""Dense convolution engine for the StyleGAN2 / SR / VGG layers (the "implicit GEMM on tensor cores" row of
SURVEY.md §8d).  Call-site contract = `_conv2d_wrapper` (eg3d/torch_utils/ops/conv2d_resample.py:30-43):
`flip_weight=True` is correlation (what `F.conv2d` computes), `False` flips the taps first.

Round-1 engine: cuDNN through `torch.nn.functional.conv2d/conv_transpose2d` in channels-last layout with TF32
tensor-core math (library call, counted as baseline -- DESIGN.md "Conv engine").  Grouped-by-batch modulated
convolutions are executed sample by sample, each a dense GEMM-shaped problem.
"""
import torch
import torch.nn.functional as F

ALLOW_TF32 = True


def conv2d(x, w, stride=1, padding=0, groups=1, transpose=False, flip_weight=True):
    if not x.is_cuda:
        raise RuntimeError('spi_b200 conv engine: x must reside on a CUDA device (no CPU path in this build)')
    if not flip_weight and (w.shape[-1] > 1 or w.shape[-2] > 1):
        w = w.flip([2, 3])
    torch.backends.cudnn.allow_tf32 = ALLOW_TF32
    op = F.conv_transpose2d if transpose else F.conv2d
    if groups == 1:
        return op(x.contiguous(memory_format=torch.channels_last), w.contiguous(memory_format=torch.channels_last),
                  stride=stride, padding=padding)
    # groups = batch of per-sample weights: run each sample as its own dense problem
    xs = x.reshape(groups, -1, *x.shape[2:])            # [G, Cin, H, W] (x is [1, G*Cin, H, W])
    if transpose:
        ws = w.reshape(groups, w.shape[0] // groups, *w.shape[1:])      # [G, Cin, Cout, kh, kw]
    else:
        ws = w.reshape(groups, w.shape[0] // groups, *w.shape[1:])      # [G, Cout, Cin, kh, kw]
    outs = []
    for g in range(groups):
        xg = xs[g:g + 1].contiguous(memory_format=torch.channels_last)
        outs.append(op(xg, ws[g].contiguous(memory_format=torch.channels_last), stride=stride, padding=padding))
    y = torch.cat(outs, 0)                               # [G, Cout, H', W']
    return y.reshape(1, -1, *y.shape[2:])
