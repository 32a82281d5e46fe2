"""Outer per-image loop scaffolding (spi/training/coaches/base_coach.py:36-270): generator reload per image, stage-1
dispatch, G-stage Adam over all G.parameters() (lr 3e-4), checkpoint / still-image output, coach naming."""
import abc
import os

import numpy as np
import torch
from PIL import Image

from ...configs import global_config, hyperparameters, paths_config
from ...criteria.lpips.lpips import LPIPS
from ...optim import FlatAdam
from ...utils import load_utils
from ...utils.camera_utils import cal_mirror_c
from ...utils.log_utils import log_image_from_w
from ...utils.metric_utils import Metric, format_metric_log, gather_metric_dic
from ...utils.video_utils import gen_interp_video
from ..projectors import mirror_projector, w_plus_projector, w_projector


def toogle_grad(model, flag=True):
    for p in model.parameters():
        p.requires_grad = flag


def fix_seed():
    """base_coach.py:28-33."""
    torch.manual_seed(0)
    torch.cuda.manual_seed_all(0)
    np.random.seed(0)
    # the reference also forces cudnn.deterministic here; this build does not depend on cuDNN algorithm choice for parity


class BaseCoach:
    def __init__(self, data_loader, use_wandb, lpips_loss=None, vgg16=None, metric=None):
        self.use_wandb = use_wandb
        self.data_loader = data_loader
        self.w_pivots = {}
        self.image_counter = 0
        self.metric_dic = {}
        self.coach_name = 'Base_coach'
        self.lpips_loss = lpips_loss if lpips_loss is not None else LPIPS(net_type='vgg').to(global_config.device).eval()
        self.restart_training()
        self.vgg16 = vgg16 if vgg16 is not None else load_utils.load_sg_vgg().to(global_config.device)
        self._metric = metric

    @property
    def metric(self):
        """base_coach.py:43 builds `Metric()` (IR-SE50 + LPIPS) eagerly; here on first use: only `use_wandb` runs need it."""
        if self._metric is None:
            self._metric = Metric(lpips_loss=self.lpips_loss)
        return self._metric

    def restart_training(self):
        """base_coach.py:53-60: fresh G and original_G per image, new optimiser, then fix_seed().  Everything keyed to the
        previous image is dropped first: its captured iteration graphs (each owns a private memory pool of several GB and can
        never be replayed against the new G), the cached tri-planes of the old original_G, and the LPIPS taps of its targets."""
        self._graphs = {}
        self._stable_key = None
        self._registered = None
        self.G = self.original_G = self.optimizer = None
        if hasattr(getattr(self, 'lpips_loss', None), 'release_targets'):
            self.lpips_loss.release_targets()
        self.G = load_utils.load_eg3d()
        toogle_grad(self.G, True)
        self.original_G = load_utils.load_eg3d()
        self.optimizer = self.configure_optimizers()
        fix_seed()

    def get_inversion(self, image_name, image, camera, fg_mask=None):
        embedding_dir = f'{paths_config.embedding_base_dir}/{self.coach_name}/'
        os.makedirs(embedding_dir, exist_ok=True)
        w_pivot = None
        if hyperparameters.load_embedding_coach_name is not None:
            w_pivot = self.load_inversions(f'{paths_config.embedding_base_dir}/{hyperparameters.load_embedding_coach_name}/', image_name)
        if w_pivot is None:
            w_pivot = self.calc_inversions(image_name, image, camera, fg_mask)
        torch.save(w_pivot, f'{embedding_dir}/{image_name}.pt')
        w_pivot = w_pivot.to(global_config.device)
        if self.use_wandb:                       # base_coach.py:77-83
            w_inv = log_image_from_w(w_pivot, camera, self.G, f'{image_name}_w_inv')
            w_inv_m = log_image_from_w(w_pivot, cal_mirror_c(camera), self.G, f'{image_name}_w_inv_m')
            self.log_video(w_pivot, self.G, os.path.join(paths_config.experiments_output_dir, f'{image_name}_w_inv.mp4'))
            self.cal_metric(w_inv, image, 'w_inv', fake_m=w_inv_m)
        return w_pivot

    def load_inversions(self, embedding_dir, image_name):
        if image_name in self.w_pivots:
            return self.w_pivots[image_name]
        path = f'{embedding_dir}/{image_name}.pt'
        if not os.path.isfile(path):
            print('[ERROR]: No existing w codes.')
            return None
        w = torch.load(path, map_location='cpu').to(global_config.device)
        self.w_pivots[image_name] = w
        return w

    def calc_inversions(self, image_name, image, camera, fg_mask=None):
        assert hyperparameters.first_inv_type in ['sg', 'sgw+', 'mir', 'reg']
        kw = dict(device=torch.device(global_config.device), w_avg_samples=600, num_steps=hyperparameters.first_inv_steps,
                  verbose=self.use_wandb, w_name=image_name, initial_w=None)
        if hyperparameters.first_inv_type == 'sg':
            return w_projector.project(self.G, image, camera, vgg16=self.vgg16, **kw)
        if hyperparameters.first_inv_type == 'sgw+':
            return w_plus_projector.project(self.G, image, camera, lpips_func=self.lpips_loss, **kw)
        if hyperparameters.first_inv_type == 'mir':
            return mirror_projector.project(self.G, image, camera, lpips_func=self.lpips_loss, fg_mask=fg_mask, **kw)
        raise NotImplementedError

    @abc.abstractmethod
    def train(self):
        pass

    def configure_optimizers(self):
        """base_coach.py:132-135: Adam over every G parameter, lr = pti_learning_rate."""
        return FlatAdam(self.G.parameters(), lr=hyperparameters.pti_learning_rate, steal_grads=True)

    def cal_metric(self, fake, gt, name, fake_m=None):
        """base_coach.py:141-154: L2 / LPIPS / ID of a render against the photo, and of the mirrored render against the
        flipped photo."""
        d = self.metric_dic.setdefault(name, {'l2': [], 'lpips': [], 'id': [], 'l2_m': [], 'lpips_m': [], 'id_m': []})
        l2, lpips, id_sim = self.metric.run(gt, fake)
        d['l2'].append(l2)
        d['lpips'].append(lpips)
        d['id'].append(id_sim)
        if fake_m is not None:
            l2, lpips, id_sim = self.metric.run(torch.flip(gt, dims=[3]), fake_m)
            d['l2_m'].append(l2)
            d['lpips_m'].append(lpips)
            d['id_m'].append(id_sim)

    def log_metric(self):
        """base_coach.py:156-198: append the per-image table and the averages to experiments/<coach>/metric_log.txt.  Under
        torchrun every rank holds the rows of its own dataset block: they are gathered in rank (= block = dataset) order and
        rank 0 writes the one file a single-process run would have written."""
        merged, writer = gather_metric_dic(self.metric_dic)
        if writer:
            with open(os.path.join(paths_config.experiments_output_dir, 'metric_log.txt'), 'a') as log_file:
                log_file.write(format_metric_log(self.coach_name, hyperparameters, merged))
        return merged

    def finish_image(self, w_pivot, image, camera, image_name):
        """Tail of the per-image loop (pti_coach.py:87-94, rot_bbox_cx_coach.py:160-167)."""
        if self.use_wandb and hyperparameters.G_1_step > 0:
            G1_inv = log_image_from_w(w_pivot, camera, self.G, f'{image_name}_G1_inv')
            G1_inv_m = log_image_from_w(w_pivot, cal_mirror_c(camera), self.G, f'{image_name}_G1_inv_m')
            self.log_video(w_pivot, self.G, path=os.path.join(paths_config.experiments_output_dir, f'{image_name}_G1_inv.mp4'))
            self.cal_metric(G1_inv, image, 'G1_inv', fake_m=G1_inv_m)
        self.post_process(w_pivot, camera, self.G, image_name)

    def save(self, w, c, G, path):
        torch.save({'w': w.detach().cpu(), 'c': c.detach().cpu(), 'G': {k: v.detach().cpu() for k, v in G.state_dict().items()}}, path)

    def load(self, path):
        ckpt = torch.load(path, map_location='cpu')
        self.G.load_state_dict(ckpt['G'])
        return ckpt['w'].to(global_config.device), ckpt['c'].to(global_config.device), self.G

    def post_process(self, w, c, G, name):
        self.save(w, c, G, path=os.path.join(paths_config.checkpoints_dir, self.coach_name, f'{name}.pt'))
        self.log_image(w, c, G, path=os.path.join(paths_config.images_output_dir, self.coach_name, name + '.jpg'))
        self.log_image(w, cal_mirror_c(c), G, path=os.path.join(paths_config.mirror_images_output_dir, self.coach_name, name + '.jpg'))
        self.log_video(w, G, path=os.path.join(paths_config.video_output_dir, self.coach_name, f'{name}.mp4'))

    def log_image(self, w, c, G, path):
        if len(w.size()) <= 2:
            w = w.unsqueeze(0)
        with torch.no_grad():
            img = G.synthesis(w, c, noise_mode='const')['image'][0].permute(1, 2, 0)
            img = (img * 127.5 + 128).clamp(0, 255).to(torch.uint8).detach().cpu().numpy()
        Image.fromarray(img).save(path)

    def log_video(self, w, G, path):
        """base_coach.py:236-237: the 120-frame orbit; one backbone pass for the clip (utils/video_utils.py)."""
        if len(w.size()) <= 2:
            w = w.unsqueeze(0)
        return gen_interp_video(G, {'w': w.detach().clone()}, mp4=path)

    def build_name(self):
        """base_coach.py:240-270."""
        hp = hyperparameters
        self.coach_name += f'_{hp.first_inv_type}_{hp.first_inv_steps}_{hp.G_1_type}_{hp.G_1_step}'
        if hp.use_encoder:
            self.coach_name += '_wenc'
        if hp.use_G_avg:
            self.coach_name += '_wgavg'
        self.coach_name += f'_rot_{hp.pt_rot_lambda}_mirrorrot_{hp.pt_mirror_rot_lambda}_depth_{hp.pt_depth_lambda}_tv_{hp.pt_tv_lambda}'
        if hp.use_adapt_yaw_range:
            self.coach_name += '_wadyaw'
        if hp.description is not None:
            self.coach_name += '_' + hp.description
        print('[COACH]:', self.coach_name)
        for d in (paths_config.checkpoints_dir, paths_config.embedding_base_dir, paths_config.experiments_output_dir,
                  paths_config.images_output_dir, paths_config.mirror_images_output_dir, paths_config.video_output_dir):
            os.makedirs(os.path.join(d, self.coach_name), exist_ok=True)
