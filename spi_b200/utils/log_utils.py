"""JPEG dumps (spi/utils/log_utils.py:8-58)."""
import os

import torch
from PIL import Image

from ..configs import paths_config


def _to_uint8_image(img_tensor):
    img = img_tensor[0].permute(1, 2, 0)
    return (img * 127.5 + 128).clamp(0, 255).to(torch.uint8).detach().cpu().numpy()


def log_image_from_w(w, c, G, name):
    if len(w.size()) <= 2:
        w = w.unsqueeze(0)
    with torch.no_grad():
        img_tensor = G.synthesis(w, c, noise_mode='const')['image']
    Image.fromarray(_to_uint8_image(img_tensor)).save(os.path.join(paths_config.experiments_output_dir, name + '.jpg'))
    return img_tensor


def log_image(t, name, vmin=-1, vmax=1, mode='jpg'):
    t = t.detach().float().cpu()
    if t.ndim == 4:
        t = t[0]
    if t.shape[0] == 1:
        v = t[0].numpy()
        v = (v - v.min()) / max(v.max() - v.min(), 1e-12)
    else:
        v = ((t.permute(1, 2, 0).numpy() - vmin) / (vmax - vmin)).clip(0, 1)
    Image.fromarray((v * 255).astype('uint8')).save(os.path.join(paths_config.experiments_output_dir, name + '.' + mode))
