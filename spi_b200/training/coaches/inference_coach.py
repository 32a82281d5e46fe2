"""`--G_1_type Inference` (spi/training/coaches/inference_coach.py:10-46): reload a finished checkpoint and re-render."""
import os

from ...configs import global_config, hyperparameters, paths_config
from .base_coach import BaseCoach


class InferenceCoach(BaseCoach):
    def __init__(self, data_loader, use_wandb, **kw):
        super().__init__(data_loader, use_wandb, **kw)
        self.coach_name = 'InferenceCoach'
        self.build_name()

    def train(self):
        for idx, data in enumerate(self.data_loader):
            if self.image_counter >= hyperparameters.max_images_to_invert:
                break
            image_name = data['name'][0]
            ckpt = os.path.join(paths_config.checkpoints_dir, hyperparameters.load_embedding_coach_name, f'{image_name}.pt')
            w_pivot, camera, self.G = self.load(ckpt)
            self.log_image(w_pivot, camera, self.G, path=os.path.join(paths_config.images_output_dir, self.coach_name, image_name + '.jpg'))
            self.image_counter += 1
