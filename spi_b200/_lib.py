"""ctypes binding of libspi_b200.so (the C ABI declared in include/spi_b200.h).

PyTorch is only the allocator / stream provider: tensors are passed as raw device pointers on torch's
current CUDA stream.  There is NO fallback: if the shared library is missing, or a tensor is not on a CUDA
device, the call raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, 'libspi_b200.so')
_lib = None

c_void_p, c_int, c_float, c_ll = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_longlong

_SIGS = {
    'spi_bias_act': [c_void_p] * 6 + [c_ll, c_int, c_int, c_int, c_int, c_int, c_float, c_float, c_float, c_void_p],
    'spi_upfirdn2d': [c_void_p] * 3 + [c_int] * 5 + [c_void_p, c_void_p] + [c_int] * 11 + [c_float, c_void_p],
    'spi_blur4_bias_act_noise': [c_void_p] * 6 + [c_int] * 4 + [c_void_p, c_void_p] + [c_int] * 5 + [c_float, c_int] + [c_float] * 3 + [c_void_p],
    'spi_filtered_lrelu_sign_shape': [c_int] * 5 + [c_void_p, c_void_p],
    'spi_filtered_lrelu': [c_void_p] * 6 + [c_int] * 5 + [c_void_p, c_void_p] + [c_int] * 14 + [c_float] * 3 + [c_int, c_int, c_void_p],
    'spi_filtered_lrelu_act': [c_void_p, c_void_p] + [c_int] * 5 + [c_void_p] + [c_int] * 4 + [c_float] * 3 + [c_int, c_void_p],
    'spi_render_forward': [c_void_p] * 9 + [c_float] + [c_void_p] * 6 + [c_int, c_int, c_ll] + [c_int] * 4 + [c_float] * 3 + [c_int, c_void_p],
    'spi_render_backward': [c_void_p] * 9 + [c_float] + [c_void_p] * 7 + [c_int, c_int, c_ll, c_ll] + [c_int] * 4 + [c_float, c_void_p],
    'spi_render_keeps_activations': [c_int, c_int],
    'spi_render_forward_keep': [c_void_p] * 9 + [c_float] + [c_void_p] * 5 + [c_int, c_int, c_ll] + [c_int] * 4 + [c_float] * 3 + [c_int] + [c_void_p] * 4 + [c_void_p],
    'spi_render_backward_kept': [c_void_p] * 9 + [c_float] + [c_void_p] * 5 + [c_int, c_int, c_ll, c_ll] + [c_int] * 4 + [c_float] + [c_void_p] * 3 + [c_void_p],
    'spi_points_forward': [c_void_p] * 6 + [c_float] + [c_void_p] * 2 + [c_int] * 4 + [c_float, c_void_p],
    'spi_points_backward': [c_void_p] * 6 + [c_float] + [c_void_p] * 7 + [c_int] * 4 + [c_float, c_void_p],
    'spi_ray_sampler': [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p],
    'spi_ray_march': [c_void_p] * 3 + [c_int] * 3 + [c_void_p] * 4 + [c_void_p],
    'spi_sample_importance': [c_void_p] * 3 + [c_int] * 3 + [c_void_p] * 3 + [c_void_p],
    'spi_inverse_cdf': [c_void_p] * 3 + [c_int] * 4 + [c_void_p] * 2 + [c_void_p],
    'spi_unify_samples': [c_void_p] * 2 + [c_int] * 3 + [c_void_p] * 2 + [c_void_p],
    'spi_rotate': [c_void_p] * 8 + [c_int] * 3 + [c_ll] * 4 + [c_float, c_void_p],
    'spi_adam_step': [c_void_p] * 4 + [c_ll] + [c_float] * 4 + [c_int, c_void_p, c_int, c_void_p, c_float, c_void_p],
    'spi_adam_step_multi': [c_void_p, c_int, c_int] + [c_float] * 4 + [c_int, c_void_p, c_void_p, c_float, c_void_p],
    'spi_bias_act_noise': [c_void_p] * 5 + [c_ll] + [c_int] * 6 + [c_float] * 3 + [c_void_p],
    'spi_bias_act_grad_reduce': [c_void_p] * 3 + [c_ll, c_int, c_int, c_int] + [c_float] * 3 + [c_void_p] * 3 + [c_void_p],
    'spi_epilogue_grad_reduce': [c_void_p, c_ll, c_int, c_int] + [c_void_p] * 4 + [c_void_p],
    'spi_modulate_weights': [c_void_p] * 4 + [c_int] * 6 + [c_void_p],
    'spi_modulate_weights_backward': [c_void_p] * 6 + [c_int] * 6 + [c_void_p],
    'spi_modulate_weights_many': [c_int] + [c_void_p] * 10 + [c_void_p],
    'spi_modulate_weights_backward_many': [c_int] + [c_void_p] * 12 + [c_void_p],
    'spi_conv_weight_transpose_many': [c_int] + [c_void_p] * 7 + [c_void_p],
    'spi_style_bank_forward': [c_void_p, c_ll, c_ll, c_int, c_int, c_int] + [c_void_p] * 8 + [c_void_p],
    'spi_style_bank_backward': [c_void_p, c_ll, c_ll, c_int, c_int, c_int] + [c_void_p] * 9 + [c_void_p, c_int, c_void_p],
    'spi_downsample2x': [c_void_p, c_void_p, c_ll, c_int, c_int, c_int, c_void_p],
    'spi_noise_reg_forward': [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p],
    'spi_noise_reg_backward': [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p],
    'spi_noise_renorm': [c_void_p, c_int, c_void_p],
    'spi_lpips_tap_forward': [c_void_p] * 3 + [c_int] * 4 + [c_void_p, c_void_p, c_void_p],
    'spi_lpips_tap_backward': [c_void_p] * 3 + [c_int] * 4 + [c_void_p, c_void_p, c_void_p, c_void_p],
    'spi_column_sums': [c_void_p, c_ll, c_int, c_void_p, c_void_p],
    'spi_maxpool2x2': [c_void_p] * 3 + [c_int] * 5 + [c_void_p],
    'spi_conv2d_tc_supported': [c_int] * 6,
    'spi_conv2d_tc': [c_void_p] * 3 + [c_int] * 8 + [c_void_p] * 3 + [c_int] + [c_float] * 3 + [c_int, c_void_p],
    'spi_conv2d_tc_error': [],
    'spi_tc_error': [],
    'spi_conv_weight_flip_transpose': [c_void_p, c_void_p] + [c_int] * 4 + [c_void_p],
    'spi_conv_tc2_supported': [c_int] * 2,
    'spi_conv_tc2_splits': [c_int] * 10,
    'spi_rows_outer_sum': [c_void_p, c_void_p, c_ll, c_int, c_int, c_void_p, c_void_p, c_void_p],
    'spi_conv2d_tc2': [c_void_p] * 3 + [c_int] * 7 + [c_void_p] * 3 + [c_int] + [c_float] * 3 + [c_int, c_void_p],
    'spi_conv_transpose2d_s2_tc2': [c_void_p] * 3 + [c_int] * 7 + [c_void_p],
    'spi_conv2d_s2_tc2': [c_void_p] * 3 + [c_int] * 7 + [c_void_p],
    'spi_conv_weight_transpose': [c_void_p, c_void_p] + [c_int] * 5 + [c_void_p],
    'spi_conv_wgrad_tc2': [c_void_p] * 3 + [c_int] * 8 + [c_void_p],
    'spi_conv1x1_rgb_supported': [c_int] * 2,
    'spi_roi_align': [c_void_p] * 3 + [c_int] * 4 + [c_void_p] + [c_int] * 2 + [c_void_p],
    'spi_roi_align_backward': [c_void_p] * 3 + [c_int] * 4 + [c_void_p] + [c_int] * 2 + [c_void_p],
    'spi_cx_rows_forward': [c_void_p] + [c_int] * 3 + [c_float] + [c_void_p] * 3,
    'spi_cx_rows_backward': [c_void_p] + [c_int] * 3 + [c_float] + [c_void_p] * 5,
    'spi_conv1x1_rgb': [c_int] + [c_void_p] * 3 + [c_ll] + [c_int] * 4 + [c_void_p],
}

ABI_VERSION = 3            # include/spi_b200.h: spi_abi_version()
EXPORTS = sorted(list(_SIGS) + ['spi_last_error', 'spi_launch_count', 'spi_reset_launch_count', 'spi_abi_version'])


def lib_path():
    return _LIB_PATH


def load():
    """Load the shared library (building is `spi_b200.build.build()`'s job).  Raises if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError(f'{_LIB_PATH} is missing: run `python -m spi_b200.build` (there is no CPU fallback)')
        lib = ctypes.CDLL(_LIB_PATH)
        for name, args in _SIGS.items():
            fn = getattr(lib, name)
            fn.argtypes = args
            fn.restype = c_int
        lib.spi_last_error.restype = ctypes.c_char_p
        lib.spi_launch_count.restype = ctypes.c_ulonglong
        lib.spi_abi_version.restype = c_int
        if lib.spi_abi_version() != ABI_VERSION:
            raise RuntimeError(f'{_LIB_PATH} has ABI version {lib.spi_abi_version()}, this package binds version {ABI_VERSION}: rebuild it '
                               '(`python -m spi_b200.build`)')
        _lib = lib
    return _lib


def ptr(t):
    """Device pointer of a CUDA tensor, or NULL for None / empty (the plugin's 'absent operand' convention)."""
    if t is None or t.numel() == 0:
        return None
    if not t.is_cuda:
        raise RuntimeError('spi_b200: expected a CUDA tensor (this build has no CPU path)')
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def strides4(t):
    return (c_ll * 4)(*t.stride())


def dtype_code(t):
    try:
        return {torch.float32: 0, torch.float16: 1, torch.float64: 2}[t.dtype]
    except KeyError:
        raise RuntimeError(f'spi_b200: unsupported dtype {t.dtype}')


def check(rc, soft_unsupported=False):
    if rc == 0:
        return 0
    if rc == -2 and soft_unsupported:
        return rc
    raise RuntimeError(load().spi_last_error().decode() or f'spi_b200 call failed with status {rc}')


def launch_count():
    return int(load().spi_launch_count())


def reset_launch_count():
    load().spi_reset_launch_count()


# ---- optional per-kernel timing hook (bench.py installs an object with .start(tag, units) / .stop(tag)); events are
#      recorded on the launching stream around the C-ABI call, only in eager passes (never inside graph capture)
KERNEL_TIMER = None


class timed:
    def __init__(self, tag, units=1, detail=None):
        self.tag, self.units, self.detail = tag, units, detail

    def __enter__(self):
        if KERNEL_TIMER is not None:
            if self.detail is not None and getattr(KERNEL_TIMER, 'wants_detail', False):
                KERNEL_TIMER.start(self.tag, self.units, detail=self.detail)
            else:
                KERNEL_TIMER.start(self.tag, self.units)

    def __exit__(self, *exc):
        if KERNEL_TIMER is not None:
            KERNEL_TIMER.stop(self.tag)
        return False
