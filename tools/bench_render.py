"""Micro-benchmark of the fused renderer alone (CUDA events, L2 flushed between iterations).  Not the headline bench."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from spi_b200.training.triplane import OSGDecoder
from spi_b200.training.volumetric_rendering.ray_sampler import RaySampler
from spi_b200.training.volumetric_rendering.renderer import ImportanceRenderer
from spi_b200.utils.camera_utils import cal_canonical_c


def main(sweep=False):
    """Default: N in {1, 4} x {32+32, 48+48} at 128^2 rays.  `--sweep`: BASELINE configs[4], the ray-march sweep of
    SURVEY.md §8d (5): (Dc, Df) in {32+32, 48+48, 64+64, 96+96} x neural_rendering_resolution in {64, 128}, N = 1."""
    dev = 'cuda'
    torch.manual_seed(0)
    dec = OSGDecoder(32, {'decoder_lr_mul': 1, 'decoder_output_dim': 32}).to(dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    R = ImportanceRenderer()
    cases = [(n, dc, df, 128) for n in (1, 4) for dc, df in ((32, 32), (48, 48))]
    if sweep:
        cases = [(1, dc, df, res) for res in (64, 128) for dc, df in ((32, 32), (48, 48), (64, 64), (96, 96))]
    for n, dc, df, res in cases:
        if True:
            rk = dict(depth_resolution=dc, depth_resolution_importance=df, ray_start=2.25, ray_end=3.3, box_warp=1, clamp_mode='softplus')
            planes = torch.randn(n, 3, 32, 256, 256, device=dev).requires_grad_(True)
            c = torch.cat([cal_canonical_c(0.1 * k, 0, 1, dev) for k in range(n)], 0)
            o, d = RaySampler()(c[:, :16].view(-1, 4, 4), c[:, 16:].view(-1, 3, 3), res)
            for mode, dec_grad in (('planes only', False), ('planes+decoder', True)):
                dec.requires_grad_(dec_grad)
                tf, tb = [], []
                for it in range(6):
                    flush.zero_()
                    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                    e[0].record()
                    feat, depth, _ = R(planes, dec, o, d, rk)
                    e[1].record()
                    (feat.sum() + depth.sum()).backward()
                    e[2].record()
                    torch.cuda.synchronize()
                    if it >= 2:
                        tf.append(e[0].elapsed_time(e[1])); tb.append(e[1].elapsed_time(e[2]))
                    planes.grad = None
                print(f'N={n} rays={res}^2 D={dc}+{df} {mode:15s} fwd {sum(tf) / len(tf) / n:7.3f} ms/img   bwd {sum(tb) / len(tb) / n:7.3f} ms/img', flush=True)


if __name__ == '__main__':
    main(sweep='--sweep' in sys.argv[1:])
