"""GPU tests of the orbit-video row (spi/utils/video_utils.py:74-230, base_coach.py:236-237): the
clip rendered with ONE backbone pass and batched views must show the frames the reference's frame-by-frame structure
(whole generator per frame) produces for the same ray-jitter draws."""
import os
import tempfile

import numpy as np
import pytest
import torch

from oracle import weights

pytestmark = pytest.mark.gpu


def _draws(G, frames, seed):
    rk = G.rendering_kwargs
    r = G.neural_rendering_resolution ** 2
    g = torch.Generator(device='cuda').manual_seed(seed)
    jit = torch.rand(frames, r, int(rk['depth_resolution']), 1, device='cuda', generator=g)
    u = torch.rand(frames * r, int(rk['depth_resolution_importance']), device='cuda', generator=g)
    return jit, u, r


def test_orbit_clip_equals_frame_by_frame_rendering(product_G):
    from spi_b200.utils.video_utils import orbit_cameras, render_orbit
    G = product_G
    ws = weights.w_pivot(5).cuda()
    F, B = 6, 4
    jit, u, r = _draws(G, F, 3)
    for s in range(0, F, B):                                   # the clip: batches of B views sharing one tri-plane set
        G.renderer.inject_noise(jit[s:s + B], u[s * r:(s + B) * r])
    frames, poses = render_orbit(G, ws, w_frames=F, batch=B)
    assert frames.shape == (F, 512, 512, 3) and frames.dtype == torch.uint8 and poses.shape == (F, 4, 4)
    assert not G.renderer._noise_queue
    cams, _ = orbit_cameras(F, device='cuda')
    for i in range(F):                                         # the reference's structure: whole generator per frame
        G.renderer.inject_noise(jit[i:i + 1], u[i * r:(i + 1) * r])
        with torch.no_grad():
            img = G.synthesis(ws, cams[i:i + 1], noise_mode='const')['image']
        ref = (img * 127.5 + 128).clamp(0, 255).to(torch.uint8)[0].permute(1, 2, 0)
        d = (frames[i].int() - ref.int()).abs()
        assert int(d.max()) <= 2 and float((d > 0).float().mean()) < 0.05, (i, int(d.max()), float((d > 0).float().mean()))
    assert float((frames[0].float() - frames[F // 2].float()).abs().mean()) > 0.5        # the camera does move


def test_orbit_frames_match_the_oracle(product_G, gen_sd):
    """Frames of the batched clip against the CPU oracle's `synthesis` at the same orbit cameras (pinned to the reference's pose
    expressions in tests/test_postprocess.py) and the same ray-jitter draws: uint8 frames within one grey level, float images 1e-3."""
    from oracle import generator as OG
    from spi_b200.utils.video_utils import orbit_cameras, render_orbit
    G = product_G
    ws = weights.w_pivot(5).cuda()
    F, B = 4, 4
    jit, u, r = _draws(G, F, 11)
    G.renderer.inject_noise(jit, u)
    frames, _ = render_orbit(G, ws, w_frames=F, batch=B)
    cams, _ = orbit_cameras(F, device='cuda')
    rk = {**OG.RENDERING_DEFAULTS, **dict(G.rendering_kwargs)}
    for i in (0, 2):
        ref = OG.synthesis(gen_sd, ws.cpu(), cams[i:i + 1].cpu(), rk, jitter=jit[i:i + 1].cpu(), u=u[i * r:(i + 1) * r].cpu())['image']
        ref8 = (ref * 127.5 + 128).clamp(0, 255).to(torch.uint8)[0].permute(1, 2, 0)
        d = (frames[i].cpu().int() - ref8.int()).abs()
        print(f'frame {i}: max |d| {int(d.max())}, pixels differing {float((d > 0).float().mean()):.4f}')
        assert int(d.max()) <= 1 and float((d > 0).float().mean()) < 0.02


def test_orbit_depth_clip_and_video_file(product_G):
    from spi_b200.utils.video_utils import gen_interp_video, render_orbit
    cv2 = pytest.importorskip('cv2')
    G = product_G
    ws = weights.w_pivot(5).cuda()
    depth, _ = render_orbit(G, ws, w_frames=3, batch=2, image_mode='image_depth')
    assert depth.shape == (3, 128, 128, 3)
    assert int(depth.min()) == 0 and int(depth.max()) == 255      # per-frame min/max stretch (video_utils.py:174-176)
    with tempfile.TemporaryDirectory() as d:
        path = gen_interp_video(G, {'w': ws}, mp4=os.path.join(d, 'clip.mp4'), w_frames=5, batch=4)
        assert os.path.isfile(path)
        cap = cv2.VideoCapture(path)
        n = 0
        while True:
            ok, frame = cap.read()
            if not ok:
                break
            assert frame.shape == (512, 512, 3)
            n += 1
        cap.release()
        assert n == 5
