"""Dense convolution engine for the StyleGAN2 / SR / VGG layers (the "implicit GEMM on tensor cores" row of
SURVEY.md §8d).  Call-site contract = `_conv2d_wrapper` (eg3d/torch_utils/ops/conv2d_resample.py:30-43):
`flip_weight=True` is correlation (what `F.conv2d` computes), `False` flips the taps first.

Round-1 engine: cuDNN's sm_100 TF32 tensor-core implicit-GEMM kernels through ATen, channels-last activations and
weights (library call, counted as baseline -- DESIGN.md "Conv engine").  Modulated convolutions carry one weight set per
sample (`conv2d_per_sample`): each sample is its own dense problem, executed inside ONE autograd node so that no
slice / cat / zero-fill traffic is generated around the per-sample calls.
"""
import torch
import torch.nn.functional as F

ALLOW_TF32 = True
CUDNN_AUTOTUNE = True      # cudnn.benchmark: measured +2.7 % end to end over the heuristic choice (every shape is first seen in an eager warm-up pass)


def _check(x):
    if not x.is_cuda:
        raise RuntimeError('spi_b200 conv engine: x must reside on a CUDA device (no CPU path in this build)')


def conv2d(x, w, stride=1, padding=0, groups=1, transpose=False, flip_weight=True):
    """Shared-weight convolution (VGG layers, generic conv2d_resample callers)."""
    _check(x)
    if not flip_weight and (w.shape[-1] > 1 or w.shape[-2] > 1):
        w = w.flip([2, 3])
    torch.backends.cudnn.allow_tf32 = ALLOW_TF32
    torch.backends.cudnn.benchmark = CUDNN_AUTOTUNE
    op = F.conv_transpose2d if transpose else F.conv2d
    return op(x.contiguous(memory_format=torch.channels_last), w.contiguous(memory_format=torch.channels_last),
              stride=stride, padding=padding, groups=groups)


def _pair(v):
    return [v, v] if isinstance(v, int) else [int(v[0]), int(v[1])]


class _PerSampleConv(torch.autograd.Function):
    """y[n] = conv(x[n], w[n]) (or conv_transpose) for n in range(N); x [N,C,H,W] channels-last, w [N,O,I,kh,kw] logical."""

    @staticmethod
    def forward(ctx, x, w, stride, padding, transpose):
        torch.backends.cudnn.allow_tf32 = ALLOW_TF32
        torch.backends.cudnn.benchmark = CUDNN_AUTOTUNE
        n, _, h, wd = x.shape
        o, kh, kw = w.shape[1], w.shape[3], w.shape[4]
        st, pd = _pair(stride), _pair(padding)
        x = x.contiguous(memory_format=torch.channels_last)
        if transpose:
            oh, ow = (h - 1) * st[0] - 2 * pd[0] + kh, (wd - 1) * st[1] - 2 * pd[1] + kw
        else:
            oh, ow = (h + 2 * pd[0] - kh) // st[0] + 1, (wd + 2 * pd[1] - kw) // st[1] + 1
        shared = (w.shape[0] == 1 and n > 1)       # one weight set for the whole batch: a single batched call
        if shared or n == 1:
            # one call covers the batch: take the op's own result (the `.out` overload of the transposed convolution is a
            # functional call + a full-size copy into `out`, 0.17 ms per 4x128x513x513 tensor)
            wk = (w[0].transpose(0, 1) if transpose else w[0]).contiguous(memory_format=torch.channels_last)
            if transpose:
                y = torch.ops.aten.cudnn_convolution_transpose(x, wk, pd, [0, 0], st, [1, 1], 1, False, False, ALLOW_TF32)
            else:
                y = torch.ops.aten.cudnn_convolution(x, wk, pd, st, [1, 1], 1, False, False, ALLOW_TF32)
            ctx.save_for_backward(x, w)
            ctx.cfg = (stride, padding, transpose)
            return y
        # every sample's result is written by cuDNN straight into its slice of the batch tensor (no cat)
        y = torch.empty(n, o, oh, ow, device=x.device, dtype=x.dtype, memory_format=torch.channels_last)
        for k in [slice(i, i + 1) for i in range(n)]:
            wi = k.start
            wk = w[wi].transpose(0, 1) if transpose else w[wi]
            wk = wk.contiguous(memory_format=torch.channels_last)
            if transpose:
                torch.ops.aten.cudnn_convolution_transpose.out(x[k], wk, pd, [0, 0], st, [1, 1], 1, False, False, ALLOW_TF32, out=y[k])
            else:
                torch.ops.aten.cudnn_convolution.out(x[k], wk, pd, st, [1, 1], 1, False, False, ALLOW_TF32, out=y[k])
        ctx.save_for_backward(x, w)
        ctx.cfg = (stride, padding, transpose)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        stride, padding, transpose = ctx.cfg
        torch.backends.cudnn.allow_tf32 = ALLOW_TF32
        torch.backends.cudnn.benchmark = CUDNN_AUTOTUNE
        n = x.shape[0]
        need_x, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        gy = gy.contiguous(memory_format=torch.channels_last)
        shared = (w.shape[0] == 1 and n > 1)
        single = shared or n == 1
        if single:          # one call covers the batch: hand cuDNN's own results to autograd (no slice copies)
            wk = (w[0].transpose(0, 1) if transpose else w[0]).contiguous(memory_format=torch.channels_last)
            gx, gwk, _ = torch.ops.aten.convolution_backward(gy, x, wk, None, _pair(stride), _pair(padding), [1, 1], transpose, [0, 0], 1,
                                                             [need_x, need_w, False])
            gw = (gwk.transpose(0, 1) if transpose else gwk).unsqueeze(0) if need_w else None
            return gx, gw, None, None, None
        gx = torch.empty_like(x) if (need_x and not single) else None
        gw = torch.empty_like(w) if need_w else None          # preserves w's (conv-native) strides
        for k in ([slice(0, n)] if shared else [slice(i, i + 1) for i in range(n)]):
            wi = 0 if shared else k.start
            wk = w[wi].transpose(0, 1) if transpose else w[wi]
            wk = wk.contiguous(memory_format=torch.channels_last)
            gxk, gwk, _ = torch.ops.aten.convolution_backward(gy[k], x[k], wk, None, _pair(stride), _pair(padding), [1, 1],
                                                              transpose, [0, 0], 1, [need_x, need_w, False])
            if need_x:
                if single:
                    gx = gxk
                else:
                    gx[k].copy_(gxk)
            if need_w:
                gw[wi].copy_(gwk.transpose(0, 1) if transpose else gwk)
        return gx, gw, None, None, None


def conv2d_per_sample(x, w, stride=1, padding=0, transpose=False, flip_weight=True):
    """x [N,Cin,H,W], w [N,Cout,Cin,kh,kw] (per-sample weights, any strides) -> [N,Cout,H',W'].
    For `transpose=True` the per-sample weight is used as conv_transpose2d's [Cin,Cout,kh,kw] (= w[n].transpose(0,1))."""
    _check(x)
    if not flip_weight and (w.shape[-1] > 1 or w.shape[-2] > 1):
        w = w.flip([3, 4])
    return _PerSampleConv.apply(x, w, stride, padding, transpose)


# ---------------------------------------------------------------------------------------------------------------------
# Opt-in engine: hand-written tcgen05 + TMA implicit GEMM (spi_b200/csrc/conv_tc05.cu).  Measured on B200 it reaches
# 225-375 TFLOP/s TF32 on the big layers against cuDNN's 540-750, so cuDNN stays the default engine.

def conv2d_tc05(x, w, per_sample=False, bias=None, noise=None, noise_strength=None, act='linear', slope=0.2, gain=1.0, clamp=None):
    """Stride-1 'same' correlation on the tcgen05 engine with the fused SynthesisLayer epilogue.
    x [N,I,H,W] (channels-last), w [G,O,I,kh,kw] logical with memory order [G][O][kh][kw][I] (G = N if per_sample else 1)."""
    from .. import _lib
    _check(x)
    lib = _lib.load()
    n, ci, h, wd = x.shape
    if w.ndim == 4:
        w = w.unsqueeze(0)
    g, co, _, kh, kw = w.shape
    assert g == (n if per_sample else 1)
    if not lib.spi_conv2d_tc_supported(h, wd, ci, co, kh, kw):
        raise RuntimeError(f'conv2d_tc05: unsupported shape {tuple(x.shape)} x {tuple(w.shape)}')
    x = x.contiguous(memory_format=torch.channels_last)
    wk = w.permute(0, 1, 3, 4, 2).contiguous()                  # [G][O][kh][kw][I]
    y = torch.empty(n, co, h, wd, device=x.device, dtype=x.dtype, memory_format=torch.channels_last)
    code = {'linear': 0, 'relu': 1, 'lrelu': 2}[act]
    _lib.check(lib.spi_conv2d_tc(_lib.ptr(x), _lib.ptr(wk), _lib.ptr(y), n, h, wd, ci, co, kh, kw, int(per_sample), _lib.ptr(bias),
                                 _lib.ptr(noise), _lib.ptr(noise_strength), code, float(slope), float(gain),
                                 -1.0 if clamp is None else float(clamp), 0, _lib.stream()))
    return y


def conv2d_tc05_input_grad(gy, w, per_sample=False):
    """Gradient of conv2d_tc05 (no epilogue) w.r.t. x: the same kernel on dy with the flipped / transposed weights."""
    from .. import _lib
    lib = _lib.load()
    if w.ndim == 4:
        w = w.unsqueeze(0)
    g, co, ci, kh, kw = w.shape
    wk = w.permute(0, 1, 3, 4, 2).contiguous()
    wt = torch.empty(g, ci, kh, kw, co, device=w.device, dtype=w.dtype)
    _lib.check(lib.spi_conv_weight_flip_transpose(_lib.ptr(wk), _lib.ptr(wt), g, co, kh * kw, ci, _lib.stream()))
    return conv2d_tc05(gy, wt.permute(0, 1, 4, 2, 3), per_sample=per_sample)
