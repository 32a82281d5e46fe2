"""GPU parity of the mirror-view contextual-loss kernels (spi_b200/csrc/boxcx.cu) against the reference's own expressions
(spi/criteria/bbox_cx_loss.py:41-59 roi_align via torchvision; :93-131,176 the relative-distance / CX chain) evaluated on the CPU."""
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu


def _chain(sim, h):
    """bbox_cx_loss.py:113-131,176 on a similarity matrix (dist = 1 - sim)."""
    d = 1 - sim
    dmin, _ = torch.min(d, dim=2, keepdim=True)
    dt = torch.clamp(d / (dmin + 1e-5), max=10., min=-10)
    w = torch.exp((1 - dt) / h)
    cx = w / torch.sum(w, dim=2, keepdim=True)
    return torch.max(cx, dim=1)[0]


@pytest.mark.parametrize('b,m,n', [(4, 1600, 1600), (2, 37, 53), (1, 5, 2048)])
def test_cx_rows_forward_and_backward(lib, b, m, n):
    from spi_b200.criteria.bbox_cx_loss import _CXRows
    gen = torch.Generator().manual_seed(b * m + n)
    f1 = torch.nn.functional.normalize(torch.randn(b, 16, m, generator=gen), dim=1)
    f2 = torch.nn.functional.normalize(torch.randn(b, 16, n, generator=gen), dim=1)
    sim = torch.bmm(f1.transpose(1, 2), f2)                       # cosine similarities in [-1, 1], as on the path
    gcol = torch.rand(b, n, generator=gen)
    sg = sim.cuda().requires_grad_(True)
    col = _CXRows.apply(sg, 0.5)
    (col * gcol.cuda()).sum().backward()
    sr = sim.double().requires_grad_(True)
    ref = _chain(sr, 0.5)
    (ref * gcol.double()).sum().backward()
    assert rel_l2(col, ref) < 1e-5
    assert rel_l2(sg.grad, sr.grad) < 1e-4


def test_roi_align_matches_torchvision(lib):
    """Boxes as get_landmark_bbox builds them (integer corners, some reaching outside the 256^2 image), output 80 x 80, forward and the
    gradient w.r.t. the image; NCHW and channels-last inputs."""
    from torchvision.ops import roi_align as tv_roi_align
    from spi_b200.criteria.bbox_cx_loss import roi_align
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(4, 3, 256, 256, generator=gen)
    rois = torch.tensor([[0, 90, 150, 170, 205], [1, -8, 30, 70, 75], [2, 200, 210, 262, 250], [3, 60, 60, 61, 61], [1, 10, 20, 190, 250]], dtype=torch.float32)
    g = torch.randn(5, 3, 80, 80, generator=gen)
    xr = x.clone().requires_grad_(True)
    ref = tv_roi_align(xr, boxes=rois, output_size=80)
    (ref * g).sum().backward()
    for fmt in (torch.contiguous_format, torch.channels_last):
        xg = x.cuda().contiguous(memory_format=fmt).requires_grad_(True)
        out = roi_align(xg, rois.cuda(), 80)
        (out * g.cuda()).sum().backward()
        assert rel_l2(out, ref) < 1e-5
        assert rel_l2(xg.grad, xr.grad) < 1e-5
