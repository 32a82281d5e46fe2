"""PTI baseline coach, `--G_1_type pti` (spi/training/coaches/pti_coach.py:13-98)."""
import os

import torch

from ...configs import global_config, hyperparameters, paths_config
from ...criteria.l2_loss import l2_loss
from .base_coach import BaseCoach


class SingleIDCoach(BaseCoach):
    def __init__(self, data_loader, use_wandb, **kw):
        super().__init__(data_loader, use_wandb, **kw)
        self.coach_name = 'SingleIDCoach'
        self.build_name()

    def calc_loss(self, generated_images, real_images):
        """pti_coach.py:17-28 with use_locality_regularization = False (hyperparameters.py:24)."""
        loss = 0.0
        loss_lpips = None
        if hyperparameters.pt_l2_lambda > 0:
            loss = loss + l2_loss(generated_images, real_images) * hyperparameters.pt_l2_lambda
        if hyperparameters.pt_lpips_lambda > 0:
            loss_lpips = torch.squeeze(self.lpips_loss(generated_images, real_images))
            loss = loss + loss_lpips * hyperparameters.pt_lpips_lambda
        return loss, loss_lpips

    def train_step(self, w_pivot, camera, image):
        """One iteration of pti_coach.py:62-74; returns (loss_lpips, stepped)."""
        generated_images = self.G.synthesis(w_pivot, camera, noise_mode='const')['image']
        loss, loss_lpips = self.calc_loss(generated_images, image)
        self.optimizer.zero_grad()
        if loss_lpips <= hyperparameters.LPIPS_value_threshold:      # one host sync per iteration, as in the reference
            return loss_lpips, False
        loss.backward()
        self.optimizer.step()
        return loss_lpips, True

    def train(self):
        paths_config.experiments_output_dir += f'{self.coach_name}'
        output_dir = paths_config.experiments_output_dir
        for idx, data in enumerate(self.data_loader):
            if self.image_counter >= hyperparameters.max_images_to_invert:
                break
            image_name = data['name'][0]
            image = data['img'].to(global_config.device)
            camera = data['c'].to(global_config.device)
            mask = data['mask'].to(global_config.device)[:, 0]
            fg_mask = 1 - (mask == 0).float()
            paths_config.experiments_output_dir = os.path.join(output_dir, image_name)
            os.makedirs(paths_config.experiments_output_dir, exist_ok=True)
            self.restart_training()
            w_pivot = self.get_inversion(image_name, image, camera, fg_mask=fg_mask)
            for i in range(hyperparameters.G_1_step):
                _, stepped = self.train_step(w_pivot, camera, image)
                if not stepped:
                    break
                global_config.training_step += 1
            self.image_counter += 1
            self.post_process(w_pivot, camera, self.G, image_name)
        paths_config.experiments_output_dir = output_dir
