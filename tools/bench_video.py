"""Orbit clip (120 frames, 512^2, 32+32 samples) rendered the reference's way (whole generator per frame,
spi/utils/video_utils.py:147-172) and with one backbone pass + batched views (spi_b200/utils/video_utils.py).
CUDA events; not the headline bench."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from spi_b200.utils import load_utils
from spi_b200.utils.video_utils import orbit_cameras, render_orbit


def timed(fn, reps=3):
    fn()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


def main():
    load_utils.DEPTH_OVERRIDE = (32, 32)
    G = load_utils.build_generator(device='cuda', seed=0)
    ws = torch.randn(1, 14, 512, device='cuda', generator=torch.Generator(device='cuda').manual_seed(5)) * 0.5
    cams, _ = orbit_cameras(120, device='cuda')

    @torch.no_grad()
    def per_frame():
        for i in range(120):
            img = G.synthesis(ws, cams[i:i + 1], noise_mode='const')['image']
            (img * 127.5 + 128).clamp(0, 255).to(torch.uint8)

    t_ref = timed(per_frame)
    print(f'frame by frame (120 generator passes)        {t_ref:8.1f} ms  {t_ref / 120:6.2f} ms/frame', flush=True)
    for b in (4, 8, 16):
        t = timed(lambda: render_orbit(G, ws, w_frames=120, batch=b))
        print(f'one backbone pass, batches of {b:2d} views         {t:8.1f} ms  {t / 120:6.2f} ms/frame  ({t_ref / t:4.2f}x)', flush=True)
    t = timed(lambda: render_orbit(G, ws, w_frames=120, batch=8, image_mode='image_depth'))
    print(f'depth clip (SR skipped), batch 8                {t:8.1f} ms  {t / 120:6.2f} ms/frame', flush=True)


if __name__ == '__main__':
    main()
