"""`--G_1_type Inference` (spi/training/coaches/inference_coach.py:10-46): reload a finished checkpoint `{w, c, G}` of the coach named
by `--load_embedding_coach_name` and render its orbit video; no optimisation.  The per-image masks / mirrored inputs the reference
prepares here (:27-36) are never consumed by its loop body and are not rebuilt."""
import os

from ...configs import hyperparameters, paths_config
from .base_coach import BaseCoach


class InferenceCoach(BaseCoach):
    def __init__(self, data_loader, use_wandb, **kw):
        super().__init__(data_loader, use_wandb, **kw)
        self.coach_name = 'InferenceCoach'
        self.build_name()

    def train(self):
        paths_config.experiments_output_dir += f'{self.coach_name}'
        output_dir = paths_config.experiments_output_dir
        videos = []
        for idx, data in enumerate(self.data_loader):
            if self.image_counter >= hyperparameters.max_images_to_invert:
                break
            image_name = data['name'][0]
            paths_config.experiments_output_dir = os.path.join(output_dir, image_name)
            os.makedirs(paths_config.experiments_output_dir, exist_ok=True)
            ckpt_path = os.path.join(paths_config.checkpoints_dir, hyperparameters.load_embedding_coach_name, f'{image_name}.pt')
            w_pivot, camera, self.G = self.load(ckpt_path)
            os.makedirs(paths_config.video_output_dir, exist_ok=True)
            videos.append(self.log_video(w_pivot, self.G, os.path.join(paths_config.video_output_dir, f'{image_name}.mp4')))
            self.image_counter += 1
        paths_config.experiments_output_dir = output_dir
        return videos
