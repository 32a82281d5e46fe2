"""Robustness: NaN / Inf inputs must propagate as NaN, never as an out-of-bounds access (run under compute-sanitizer)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import weights
from spi_b200.utils import load_utils
from spi_b200 import _lib
load_utils.DEPTH_OVERRIDE = (32, 32)
G = load_utils.build_generator(device='cuda', seed=0).requires_grad_(True)
ws = weights.w_pivot(5).cuda()
c = weights.canonical_camera(0.3).cuda()
for name, bad in (('nan', float('nan')), ('inf', float('inf')), ('huge', 1e30)):
    w = ws.clone()
    w[0, 3, :7] = bad
    out = G.synthesis(w, c, noise_mode='const')
    (out['image'].nan_to_num().sum() + out['image_depth'].nan_to_num().sum()).backward()
    torch.cuda.synchronize()
    print(name, 'ok; finite image fraction', float(torch.isfinite(out['image']).float().mean()), 'tc err', _lib.load().spi_tc_error(), flush=True)
print('done')
