"""spi/criteria/l2_loss.py:3-8."""
import torch

l2_criterion = torch.nn.MSELoss(reduction='mean')


def l2_loss(real_images, generated_images):
    return l2_criterion(real_images, generated_images)
