"""Orbit video of a finished inversion (drop-in for the used part of spi/utils/video_utils.py:30-230; called from
`BaseCoach.log_video`, base_coach.py:236-237, with one latent and the default 120 frames).

Same entry point (`gen_interp_video(G, G_kwargs, mp4, ...)`), same camera trajectory (yaw 3.14/2 + 0.7 sin, pitch
3.14/2 - 0.05 + 0.4 cos, radius 2.7, look-at [0, 0, 0.2]; the reference's `3.14`, not pi), same keyframe interpolation
and frame layout.  Execution differs: the reference renders frame by frame through the whole generator (120 backbone
passes of 94 GF for one constant latent).  Here the camera poses of all frames are built in one batched call and, for a
single keyframe (the only case the coaches use), the latent is handed to `synthesis` as a broadcast view, so the
camera-independent tri-plane backbone runs ONCE and the fused renderer + SR network run over batches of views
(DESIGN.md §6).  Frames go through `imageio` (libx264 .mp4) when it is installed; otherwise through the Motion-JPEG
AVI writer below (same frames, `.avi` next to the requested path), because this image has no mp4 encoder.
Shape extraction (`gen_shapes=True`, needs `shape_utils` / `mrcfile`) is not built.
"""
import io
import os
import struct

import numpy as np
import torch
from PIL import Image

from .camera_utils import LookAtPoseSampler, _intrinsics


def layout_grid(img, grid_w=None, grid_h=1, float_to_uint8=True, chw_to_hwc=True, to_numpy=True):
    """video_utils.py:30-45: [grid_h*grid_w, C, H, W] -> one [grid_h*H, grid_w*W, C] uint8 frame."""
    batch_size, channels, img_h, img_w = img.shape
    if grid_w is None:
        grid_w = batch_size // grid_h
    assert batch_size == grid_w * grid_h
    if float_to_uint8:
        img = (img * 127.5 + 128).clamp(0, 255).to(torch.uint8)
    img = img.reshape(grid_h, grid_w, channels, img_h, img_w).permute(2, 0, 3, 1, 4)
    img = img.reshape(channels, grid_h * img_h, grid_w * img_w)
    if chw_to_hwc:
        img = img.permute(1, 2, 0)
    if to_numpy:
        img = img.cpu().numpy()
    return img


def orbit_cameras(num_frames, cfg='FFHQ', device='cpu', yaw_range=0.7, pitch_range=0.4, radius=2.7):
    """All `num_frames` 25-d cameras of the orbit in one batched call (video_utils.py:153-160 evaluates them one by one:
    the angles are float64 numpy scalars rounded to float32 when they meet the float32 zero-noise tensor, reproduced here)."""
    t = 2 * 3.14 * np.arange(num_frames, dtype=np.float64) / num_frames
    yaw = torch.from_numpy((3.14 / 2 + yaw_range * np.sin(t)).astype(np.float32)).view(-1, 1).to(device)
    pitch = torch.from_numpy((3.14 / 2 - 0.05 + pitch_range * np.cos(t)).astype(np.float32)).view(-1, 1).to(device)
    lookat = torch.tensor([0, 0, 0.2] if cfg == 'FFHQ' else [0, 0, 0], dtype=torch.float32, device=device)
    cam2world = LookAtPoseSampler.sample(yaw, pitch, lookat, radius=radius, batch_size=num_frames, device=device)
    return torch.cat([cam2world.reshape(-1, 16), _intrinsics(num_frames, device)], 1), cam2world


class MJPEGWriter:
    """Minimal Motion-JPEG AVI (RIFF) writer: one `00dc` chunk per frame + `idx1`; plays in ffmpeg / VLC / browsers via
    ffmpeg.  Used only when imageio is not installed."""

    def __init__(self, path, fps=60, quality=90):
        self.path, self.fps, self.quality = path, fps, quality
        self.frames, self.size = [], None

    def append_data(self, frame):
        img = Image.fromarray(np.ascontiguousarray(frame))
        self.size = self.size or img.size
        assert img.size == self.size, 'all frames of a video share one size'
        buf = io.BytesIO()
        img.save(buf, format='JPEG', quality=self.quality)
        self.frames.append(buf.getvalue())

    def close(self):
        w, h = self.size or (0, 0)
        n = len(self.frames)

        def chunk(tag, payload):
            return tag + struct.pack('<I', len(payload)) + payload + (b'\0' if len(payload) & 1 else b'')

        def lst(tag, payload):
            return chunk(b'LIST', tag + payload)

        biggest = max((len(f) for f in self.frames), default=0)
        avih = struct.pack('<14I', 1000000 // self.fps, biggest * self.fps, 0, 0x10, n, 0, 1, biggest, w, h, 0, 0, 0, 0)
        strh = b'vids' + b'MJPG' + struct.pack('<IHHIIIIIIII4H', 0, 0, 0, 0, 1, self.fps, 0, n, biggest, 0xFFFFFFFF, 0, 0, 0, w, h)
        strf = struct.pack('<IiiHH4sIiiII', 40, w, h, 1, 24, b'MJPG', w * h * 3, 0, 0, 0, 0)
        hdrl = lst(b'hdrl', chunk(b'avih', avih) + lst(b'strl', chunk(b'strh', strh) + chunk(b'strf', strf)))
        movi_payload, index, offset = b'', b'', 4
        for f in self.frames:
            c = chunk(b'00dc', f)
            index += b'00dc' + struct.pack('<III', 0x10, offset, len(f))
            movi_payload += c
            offset += len(c)
        body = b'AVI ' + hdrl + lst(b'movi', movi_payload) + chunk(b'idx1', index)
        with open(self.path, 'wb') as fh:
            fh.write(b'RIFF' + struct.pack('<I', len(body)) + body)
        self.frames = []


def open_writer(mp4, fps=60, **video_kwargs):
    """imageio's libx264 writer when available (the reference's call, video_utils.py:133); else the MJPEG fallback."""
    try:
        import imageio
        return imageio.get_writer(mp4, mode='I', fps=fps, codec='libx264', **video_kwargs), mp4
    except ImportError:
        path = os.path.splitext(mp4)[0] + '.avi'
        return MJPEGWriter(path, fps=fps), path


def interpolate_keyframes(ws, num_frames_per_key, kind='cubic', wraps=2):
    """video_utils.py:118-129 for a 1x1 grid: periodic interpolation of the keyframe latents, evaluated at
    frame_idx / w_frames.  ws [K, L, C] -> [K * num_frames_per_key, L, C]."""
    k = ws.shape[0]
    if k == 1:
        return None                                   # constant latent: the caller broadcasts it instead of copying it
    import scipy.interpolate
    x = np.arange(-k * wraps, k * (wraps + 1))
    y = np.tile(ws.detach().cpu().numpy(), [wraps * 2 + 1, 1, 1])
    interp = scipy.interpolate.interp1d(x, y, kind=kind, axis=0)
    return torch.from_numpy(interp(np.arange(k * num_frames_per_key) / num_frames_per_key)).to(ws.dtype)


@torch.no_grad()
def render_orbit(G, ws, w_frames=30 * 4, kind='cubic', wraps=2, cfg='FFHQ', image_mode='image', batch=8, device=None):
    """The frames of the orbit as one uint8 tensor [F, H, W, 3] (kept on the device until the caller encodes them)."""
    device = device if device is not None else ws.device
    num_keyframes = ws.shape[0]
    n = num_keyframes * w_frames
    cams, poses = orbit_cameras(n, cfg=cfg, device=device)
    w_all = interpolate_keyframes(ws, w_frames, kind=kind, wraps=wraps)
    frames = []
    kept = G._last_planes
    for s in range(0, n, batch):
        c = cams[s:s + batch]
        if w_all is None:
            # constant latent: a stride-0 batch (shared styles in the SR network) and the reference's own backbone cache
            # (triplane.py:66-71) -- the tri-planes are synthesised by the first batch and re-used by the others
            w = ws[:1].expand(c.shape[0], -1, -1)
            cache = dict(cache_backbone=(s == 0), use_cached_backbone=(s > 0))
        else:
            w = w_all[s:s + batch].to(device)
            cache = {}
        need_image = (image_mode == 'image')
        img = G.synthesis(ws=w, c=c, noise_mode='const', need_image=need_image, **cache)[image_mode]
        if image_mode == 'image_depth':                   # per-frame min/max stretch, video_utils.py:174-176
            img = -img
            lo = img.amin(dim=(1, 2, 3), keepdim=True)
            hi = img.amax(dim=(1, 2, 3), keepdim=True)
            img = (img - lo) / (hi - lo) * 2 - 1
        img = (img * 127.5 + 128).clamp(0, 255).to(torch.uint8)
        if img.shape[1] == 1:
            img = img.expand(-1, 3, -1, -1)
        frames.append(img.permute(0, 2, 3, 1).contiguous())
    G._last_planes = kept
    return torch.cat(frames), poses


def gen_interp_video(G, G_kwargs, mp4, images=None, shuffle_seed=None, w_frames=30 * 4, kind='cubic', grid_dims=(1, 1),
                     num_keyframes=None, wraps=2, psi=1, truncation_cutoff=14, cfg='FFHQ', image_mode='image', gen_shapes=False,
                     device=None, batch=8, **video_kwargs):
    """video_utils.py:74-230.  `G_kwargs['w']` [K, 14, 512]: K keyframe latents (K = 1 from the coaches).  Returns the path
    actually written."""
    if gen_shapes:
        raise NotImplementedError('gen_shapes needs shape_utils / mrcfile (SURVEY.md §8f: out of the inversion path)')
    if tuple(grid_dims) != (1, 1):
        raise NotImplementedError('the inversion path renders a 1x1 grid (base_coach.py:236-237)')
    if 'planes_s' in G_kwargs or 'wt' in G_kwargs:
        raise NotImplementedError('synthesis_planes / wt variants belong to the editing tools, not to the inversion path')
    ws = G_kwargs['w']
    if num_keyframes is not None and num_keyframes != ws.shape[0]:
        raise ValueError('Number of input latents must equal num_keyframes for a 1x1 grid')
    frames, _ = render_orbit(G, ws, w_frames=w_frames, kind=kind, wraps=wraps, cfg=cfg, image_mode=image_mode, batch=batch,
                             device=device)
    writer, path = open_writer(mp4, fps=60, **video_kwargs)
    host = frames.cpu().numpy()                           # one device->host copy for the whole clip
    for f in host:
        writer.append_data(f)
    writer.close()
    return path
