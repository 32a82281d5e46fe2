"""Robustness of the loss kernels to NaN images (run under compute-sanitizer)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import weights
from spi_b200.criteria.lpips.lpips import LPIPS
from spi_b200.criteria.bbox_cx_loss import BoxCXLoss
from spi_b200.optim import FlatAdam
torch.manual_seed(0)
lp = LPIPS(net_type='vgg').cuda().eval()
cx = BoxCXLoss().cuda().eval()
y = weights.target_image().cuda()
lm = weights.landmarks68().cuda()
for name, bad in (('nan', float('nan')), ('inf', float('inf'))):
    x = (y * 0.9).clone()
    x[0, 1, 100:140, 90:200] = bad
    x.requires_grad_(True)
    l1 = lp(x, y)
    l2 = cx(x.repeat(2, 1, 1, 1), y.repeat(2, 1, 1, 1), lm.repeat(2, 1, 1))
    (l1.sum() + l2).backward()
    torch.cuda.synchronize()
    print(name, 'lpips', float(l1), 'cx', float(l2), flush=True)
p = [torch.randn(1000, device='cuda', requires_grad=True), torch.randn(33, 7, device='cuda', requires_grad=True)]
opt = FlatAdam(p, lr=1e-3, steal_grads=True)
for i in range(12):
    for q in p:
        q.grad = torch.randn_like(q) * (float('nan') if i == 5 else 1.0)
    opt.step()
torch.cuda.synchronize()
print('adam eager ring ok')
