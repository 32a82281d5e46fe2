"""PTI baseline coach, `--G_1_type pti` (spi/training/coaches/pti_coach.py:13-98)."""
import os

import torch

from ...configs import global_config, hyperparameters, paths_config
from ...graphs import GraphedStep
from ...ops import zero_arena
from ...utils import rng
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

    def _body(self, w_pivot, camera, image):
        # one zero-filled buffer per iteration for every accumulate-into output (ops/zero_arena.py)
        with zero_arena.iteration(('pti', tuple(w_pivot.shape), tuple(image.shape)), w_pivot.device):
            return self._iteration(w_pivot, camera, image)

    def _iteration(self, w_pivot, camera, image):
        """pti_coach.py:62-74 without host synchronisation; the early exit is applied on the device (conditional Adam)."""
        generated_images = self.G.synthesis(w_pivot, camera, noise_mode='const')['image']
        loss, loss_lpips = self.calc_loss(generated_images, image)
        self.optimizer.zero_grad()
        loss.backward()
        with torch.no_grad():
            self._lpips_out.copy_(loss_lpips.detach())
        self.optimizer.skip_if_le = (self._lpips_out, hyperparameters.LPIPS_value_threshold)
        self.optimizer.step(in_graph=True)
        return self._lpips_out

    def train_step(self, w_pivot, camera, image):
        """One iteration; returns (loss_lpips, stepped)."""
        if not hasattr(self, '_lpips_out') or self._lpips_out.device != w_pivot.device:
            self._lpips_out = torch.zeros((), device=w_pivot.device)
            self._graphs = {}
        if hasattr(self.lpips_loss, 'register_target') and getattr(self, '_registered', None) is not image:
            self.lpips_loss.register_target(image)
            self._registered = image
        if self.optimizer.hyper is None:
            self.optimizer.use_device_hyper()
        self.optimizer._reseat()
        self.optimizer.advance()
        eager = (not global_config.use_cuda_graphs) or rng.pending() or bool(self.G.renderer._noise_queue)
        if eager:
            self._body(w_pivot, camera, image)
        else:
            key = (id(w_pivot), id(camera), id(image), id(self.G))
            if key not in self._graphs:
                state = [self.optimizer.arena, self.optimizer.exp_avg, self.optimizer.exp_avg_sq]
                self._graphs[key] = (GraphedStep(lambda: self._body(w_pivot, camera, image), state), w_pivot, camera, image)
            self._graphs[key][0]()
        stepped = bool(self._lpips_out > hyperparameters.LPIPS_value_threshold)
        if not stepped:
            self.optimizer.steps -= 1
        return self._lpips_out, stepped

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
            self.finish_image(w_pivot, image, camera, image_name)
        paths_config.experiments_output_dir = output_dir
        if self.use_wandb:
            self.log_metric()
