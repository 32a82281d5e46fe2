"""GPU parity of the opt-in tcgen05 + TMA implicit-GEMM convolution (spi_conv2d_tc) against an fp64 evaluation of the same
correlation (oracle.ops has no conv of its own: the reference calls torch's, conv2d_resample.py:30-43).  TF32 operands,
fp32 accumulation: tolerance 1e-3 rel-L2 (north_star), measured 3e-4 = the same as cuDNN's TF32 kernels."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-3


def _ref(x, w, per_sample):
    xs, ws = x.double().cpu(), w.double().cpu()
    k = w.shape[-1]
    return torch.cat([F.conv2d(xs[i:i + 1], ws[i if per_sample else 0], padding=k // 2) for i in range(x.shape[0])])


@pytest.mark.parametrize('n,ci,co,h,wd,k,per_sample', [
    (1, 32, 32, 16, 16, 3, False),       # smallest supported tile
    (2, 64, 128, 40, 56, 3, True),       # ragged spatial size (TMA clips the border tiles), per-sample weights
    (1, 128, 96, 64, 64, 1, True),       # 1x1 (toRGB shape), O not a multiple of the N tile
    (3, 96, 320, 24, 16, 3, False),      # two N tiles, the second partial
    (1, 256, 256, 64, 64, 3, True),      # BN = 256 path
])
def test_conv_tc05_matches_fp64(lib, n, ci, co, h, wd, k, per_sample):
    from spi_b200.ops.conv import conv2d_tc05, conv2d_tc05_input_grad
    gen = torch.Generator().manual_seed(n * 1000 + ci + co + h)
    x = torch.randn(n, ci, h, wd, generator=gen).cuda().contiguous(memory_format=torch.channels_last)
    w = (torch.randn(n if per_sample else 1, co, ci, k, k, generator=gen) / (ci * k * k) ** 0.5).cuda()
    y = conv2d_tc05(x, w, per_sample=per_sample)
    assert lib.spi_conv2d_tc_error() == 0
    assert y.shape == (n, co, h, wd) and y.is_contiguous(memory_format=torch.channels_last)
    assert rel_l2(y, _ref(x, w, per_sample)) < TOL
    # data gradient = the same kernel on dy with the flipped, transposed weights
    gy = torch.randn(n, co, h, wd, generator=gen).cuda().contiguous(memory_format=torch.channels_last)
    xr = x.double().cpu().requires_grad_(True)
    _ref_y = torch.cat([F.conv2d(xr[i:i + 1], w.double().cpu()[i if per_sample else 0], padding=k // 2) for i in range(n)])
    _ref_y.backward(gy.double().cpu())
    gx = conv2d_tc05_input_grad(gy, w, per_sample=per_sample)
    assert lib.spi_conv2d_tc_error() == 0
    assert rel_l2(gx, xr.grad) < TOL


def test_conv_tc05_fused_epilogue(lib):
    """noise + bias + lrelu + gain + clamp in the accumulator read-out = SynthesisLayer tail (networks_stylegan2.py:320-329)."""
    from spi_b200.ops.conv import conv2d_tc05
    gen = torch.Generator().manual_seed(7)
    n, ci, co, h = 2, 64, 96, 40
    x = torch.randn(n, ci, h, h, generator=gen).cuda().contiguous(memory_format=torch.channels_last)
    w = (torch.randn(1, co, ci, 3, 3, generator=gen) / (9 * ci) ** 0.5).cuda()
    b, nz, st = torch.randn(co, generator=gen).cuda(), torch.randn(h, h, generator=gen).cuda(), torch.tensor(0.7).cuda()
    y = conv2d_tc05(x, w, bias=b, noise=nz, noise_strength=st, act='lrelu', gain=2 ** 0.5, clamp=1.5)
    ref = _ref(x, w, False) + (nz.double() * 0.7).cpu() + b.double().cpu().view(1, -1, 1, 1)
    ref = (F.leaky_relu(ref, 0.2) * 2 ** 0.5).clamp(-1.5, 1.5)
    assert lib.spi_conv2d_tc_error() == 0
    assert rel_l2(y, ref) < TOL


def test_conv_tc05_rejects_unsupported_shapes(lib):
    from spi_b200.ops.conv import conv2d_tc05
    x = torch.zeros(1, 3, 32, 32, device='cuda')
    with pytest.raises(RuntimeError):
        conv2d_tc05(x, torch.zeros(1, 64, 3, 3, 3, device='cuda'))
