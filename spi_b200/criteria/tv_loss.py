"""Density total-variation regulariser (spi/criteria/tv_loss.py:9-21); off by default (pt_tv_lambda = 0)."""
import torch

density_reg_p_dist = 0.004


def cal_tv_loss(ws, G):
    initial = torch.rand((ws.shape[0], 1000, 3), device=ws.device) * 2 - 1
    perturbed = initial + torch.randn_like(initial) * density_reg_p_dist
    coords = torch.cat([initial, perturbed], dim=1)
    sigma = G.sample_mixed(coords, torch.randn_like(coords), ws, update_emas=False)['sigma']
    half = sigma.shape[1] // 2
    return torch.nn.functional.l1_loss(sigma[:, :half], sigma[:, half:])
