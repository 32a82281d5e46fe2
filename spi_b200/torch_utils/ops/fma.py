"""`fma(a, b, c) = a*b + c` (eg3d/torch_utils/ops/fma.py:15-16); plain torch -- it is off the executed path
(only the non-fused modulated-conv branch calls it)."""
import torch


def fma(a, b, c):
    return torch.addcmul(c, a, b)
