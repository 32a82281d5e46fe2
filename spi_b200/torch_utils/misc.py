"""The two helpers of eg3d/torch_utils/misc.py the inversion path relies on."""
import torch


def assert_shape(tensor, ref_shape):
    """misc.py:79-95 (without the tracing branches)."""
    if tensor.ndim != len(ref_shape):
        raise AssertionError(f'Wrong number of dimensions: got {tensor.ndim}, expected {len(ref_shape)}')
    for idx, (size, ref) in enumerate(zip(tensor.shape, ref_shape)):
        if ref is not None and size != ref:
            raise AssertionError(f'Wrong size for dimension {idx}: got {size}, expected {ref}')


def named_params_and_buffers(module):
    return list(module.named_parameters()) + list(module.named_buffers())


def copy_params_and_buffers(src_module, dst_module, require_all=False):
    """By-name copy (misc.py:157-164): the state-dict contract of SURVEY.md §8b.2."""
    src = dict(named_params_and_buffers(src_module))
    for name, tensor in named_params_and_buffers(dst_module):
        assert (name in src) or (not require_all), name
        if name in src:
            tensor.copy_(src[name].detach()).requires_grad_(tensor.requires_grad)
