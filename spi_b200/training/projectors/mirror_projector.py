"""W+ projector with the mirrored view, `first_inv_type='mir'` (spi/training/projectors/mirror_projector.py:12-142)."""
import torch

from ._common import LatentProjector


def project(G, target, c, lpips_func, fg_mask, *, initial_w=None, num_steps=1000, w_avg_samples=10000, initial_learning_rate=0.01,
            initial_noise_factor=0.05, lr_rampdown_length=0.25, lr_rampup_length=0.05, noise_ramp_length=0.75,
            regularize_noise_weight=1e5, verbose=False, device: torch.device = None, image_log_step=None, w_name: str = ''):
    # fg_mask only feeds the reference's unused bg_loss (mirror_projector.py:74-79,117-118) and a debug JPEG.
    p = LatentProjector(G, target, c, 'mir', lpips_func=lpips_func, initial_w=initial_w, num_steps=num_steps,
                        w_avg_samples=w_avg_samples, initial_learning_rate=initial_learning_rate,
                        initial_noise_factor=initial_noise_factor, lr_rampdown_length=lr_rampdown_length,
                        lr_rampup_length=lr_rampup_length, noise_ramp_length=noise_ramp_length,
                        regularize_noise_weight=regularize_noise_weight, device=device)
    for step in range(num_steps):
        p.step(step)
    return p.result()
