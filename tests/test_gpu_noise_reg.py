"""GPU parity of the fused noise-buffer regulariser / re-normalisation (spi_noise_reg_*, spi_noise_renorm) against the oracle's
restatement of mirror_projector.py:107-115,128-131 on the generator's own 13 buffer shapes (4^2 ... 256^2)."""
import pytest
import torch

from conftest import rel_l2
from oracle import criteria as OC

pytestmark = pytest.mark.gpu
SIZES = [4, 8, 8, 16, 16, 32, 32, 64, 64, 128, 128, 256, 256]


def _bufs(seed):
    gen = torch.Generator().manual_seed(seed)
    return [torch.randn(s, s, generator=gen) * (1 + 0.1 * i) + 0.05 * i for i, s in enumerate(SIZES)]


def test_noise_regulariser_value_and_gradients(lib):
    from spi_b200.ops import noise_reg
    cpu = [b.clone().requires_grad_(True) for b in _bufs(0)]
    ref = OC.noise_regulariser(cpu)
    (ref * 1e5).backward()
    dev = [b.cuda().requires_grad_(True) for b in _bufs(0)]
    out = noise_reg.noise_regulariser(dev)
    (out * 1e5).backward()
    assert abs(float(out) - float(ref)) <= 1e-5 * abs(float(ref))
    for d, c in zip(dev, cpu):
        assert rel_l2(d.grad, c.grad) < 1e-5, d.shape


def test_noise_regulariser_subset_of_buffers(lib):
    """w_projector-style call with fewer buffers and a different upstream scale."""
    from spi_b200.ops import noise_reg
    cpu = [b.clone().requires_grad_(True) for b in _bufs(1)[:5]]
    ref = OC.noise_regulariser(cpu)
    (ref * 3.0).backward()
    dev = [b.cuda().requires_grad_(True) for b in _bufs(1)[:5]]
    out = noise_reg.noise_regulariser(dev)
    (out * 3.0).backward()
    assert abs(float(out) - float(ref)) <= 1e-5 * abs(float(ref))
    for d, c in zip(dev, cpu):
        assert rel_l2(d.grad, c.grad) < 1e-5


def test_noise_renormalisation(lib):
    from spi_b200.ops import noise_reg
    cpu = _bufs(2)
    dev = [b.cuda() for b in cpu]
    OC.renormalise_noise_(cpu)
    noise_reg.renormalise_noise_(dev)
    for d, c in zip(dev, cpu):
        assert rel_l2(d, c) < 2e-6
        assert abs(float(d.mean())) < 1e-5 and abs(float(d.square().mean()) - 1) < 1e-5


def test_noise_reg_rejects_cpu_tensors():
    from spi_b200.ops import noise_reg
    with pytest.raises(RuntimeError):
        noise_reg.noise_regulariser([torch.zeros(8, 8)])
