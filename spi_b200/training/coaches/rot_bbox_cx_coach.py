"""SPI coach, `--G_1_type RotBbox` (spi/training/coaches/rot_bbox_cx_coach.py:15-171): PTI reconstruction loss every
iteration plus, every 4th iteration, the rotation (depth-guided warp + LPIPS), mirror (warp of the flipped image +
BoxCX) and depth-regularisation branches.  Gradients of up to four backward() calls accumulate before one Adam step;
the early-exit test precedes the step."""
import os

import torch

from ...configs import global_config, hyperparameters, paths_config
from ...graphs import GraphedStep
from ...ops import zero_arena
from ...utils import rng
from ...criteria.bbox_cx_loss import BoxCXLoss
from ...criteria.l2_loss import l2_loss
from ...criteria.tv_loss import cal_tv_loss
from ...utils.camera_utils import cal_camera_gauss_weight, cal_camera_weight, cal_mirror_c, sample_camera, sample_surrounding_camera
from ...utils.mask_utils import calculate_face_mask
from ...utils.rotate import rotate
from .base_coach import BaseCoach


class SPIState:
    """Per-image constants of the G-stage loop (rot_bbox_cx_coach.py:34-66)."""

    def __init__(self, image, camera, mask, lm):
        self.image, self.camera, self.lm = image, camera, lm
        self.fg_mask = 1 - (mask == 0).float()
        self.face_mask = calculate_face_mask(mask).float()
        self.fg_mask_m = torch.flip(self.fg_mask, dims=[3])
        self.face_mask_m = torch.flip(self.face_mask, dims=[3])
        self.camera_m = cal_mirror_c(camera=camera)
        self.image_m = torch.flip(image, dims=[3])
        self.weight_m = cal_camera_weight(camera)
        self.mirror_on = bool(self.weight_m > 0)           # one sync per image, not per iteration
        self.yaw_range = cal_camera_gauss_weight(camera)[0].item() if hyperparameters.use_adapt_yaw_range else 0.2


class RotBboxCoach(BaseCoach):
    def __init__(self, data_loader, use_wandb, box_cx_loss=None, **kw):
        super().__init__(data_loader, use_wandb, **kw)
        self.coach_name = 'RotBboxCoach'
        self.build_name()
        self.box_cx_loss = box_cx_loss if box_cx_loss is not None else BoxCXLoss().to(global_config.device).eval()

    def _body(self, heavy, st, w_pivot, rot_bs=4):
        # one zero-filled buffer per iteration for every accumulate-into output (ops/zero_arena.py); light and heavy iterations differ in size
        with zero_arena.iteration(('rot', bool(heavy), rot_bs, tuple(w_pivot.shape)), w_pivot.device):
            return self._iteration(heavy, st, w_pivot, rot_bs)

    def _iteration(self, heavy, st, w_pivot, rot_bs=4):
        """One iteration of rot_bbox_cx_coach.py:68-157 without host synchronisation (capturable as a CUDA graph).
        The early exit (`if loss_lpips <= threshold: break` BEFORE `optimizer.step()`, :148-151) is applied on the device:
        the Adam kernel is a no-op when the LPIPS scalar is below the threshold; the host reads the scalar afterwards."""
        hp = hyperparameters
        share = global_config.share_backbone and heavy
        # share=True: the tri-plane backbone (camera-independent) runs once per iteration and is reused by every view; the
        # four losses are summed and back-propagated once.  share=False replays the reference's structure literally.
        kw = dict(use_cached_backbone=True) if share else {}
        rep = (lambda t, k: t.expand(k, -1, -1)) if share else (lambda t, k: t.repeat(k, 1, 1))
        self.optimizer.zero_grad()
        gen = self.G.synthesis(w_pivot, st.camera, noise_mode='const', cache_backbone=share)
        generated_images, generated_depths = gen['image'], gen['image_depth']
        loss = 0.0
        if hp.pt_l2_lambda > 0:
            loss = loss + l2_loss(generated_images, st.image) * hp.pt_l2_lambda
        if hp.pt_lpips_lambda > 0:
            loss_lpips = torch.squeeze(self.lpips_loss(generated_images, st.image))
            loss = loss + loss_lpips * hp.pt_lpips_lambda
        if not share:
            loss.backward()
            loss = 0.0
        if heavy:
            if hp.pt_rot_lambda > 0:
                cams = sample_surrounding_camera(st.camera, batch_size=rot_bs, yaw_range=st.yaw_range, pitch_range=0.1)
                samples = self.G.synthesis(rep(w_pivot, rot_bs), cams, noise_mode='const', **kw)
                with torch.no_grad():       # broadcast sources instead of .repeat(rot_bs, ...) copies
                    warp_img, warp_mask = rotate(target_camera=cams, target_depth=samples['image_depth'], src_image=st.image,
                                                 src_camera=st.camera, src_depth=generated_depths, src_mask=st.face_mask, EPS=5e-2)
                loss_rot = self.lpips_loss(samples['image'] * warp_mask, warp_img) * hp.pt_rot_lambda * rot_bs
                if share:
                    loss = loss + loss_rot
                else:
                    loss_rot.backward()
            if hp.pt_mirror_rot_lambda > 0 and st.mirror_on:
                cams_m = sample_surrounding_camera(st.camera_m, batch_size=rot_bs, yaw_range=st.yaw_range, pitch_range=0.1)
                samples_m = self.G.synthesis(rep(w_pivot, rot_bs), cams_m, noise_mode='const', **kw)
                with torch.no_grad():
                    depths_m = torch.flip(generated_depths, dims=[3])
                    warp_img_m, warp_mask_m = rotate(target_camera=cams_m, target_depth=samples_m['image_depth'], src_image=st.image_m,
                                                     src_camera=st.camera_m, src_depth=depths_m, src_mask=st.face_mask_m, EPS=5e-2)
                    flip_warp_img_m = torch.flip(warp_img_m, dims=[3])
                    flip_warp_mask_m = torch.flip(warp_mask_m, dims=[3])
                lm = st.lm.repeat(rot_bs, 1, 1)
                flip_gen_image = torch.flip(samples_m['image'], dims=[3])
                loss_rot_m = self.box_cx_loss(flip_gen_image * flip_warp_mask_m, flip_warp_img_m, lm) * hp.pt_mirror_rot_lambda * rot_bs
                if share:
                    loss = loss + loss_rot_m
                else:
                    loss_rot_m.backward()
            if hp.pt_depth_lambda > 0:
                new_camera = sample_camera(batch_size=4, yaw_range=0.7, pitch_range=0.4, device=global_config.device)
                new_ws = rep(w_pivot, 4)
                # only image_depth is consumed here (:133-139): the super-resolution network is dead code for this branch
                sample_depth = self.G.synthesis(new_ws, new_camera, noise_mode='const', need_image=not share, **kw)['image_depth']
                with torch.no_grad():
                    if share:
                        # original_G is frozen and w_pivot is constant during this stage: its camera-independent tri-planes are the
                        # same in every iteration, so they are synthesised once per image (identical results)
                        key = (w_pivot.data_ptr(), w_pivot._version, id(self.original_G))
                        hit = getattr(self, '_stable_key', None) == key and self.original_G._last_planes is not None
                        stable_depth = self.original_G.synthesis(new_ws, new_camera, noise_mode='const', need_image=False, cache_backbone=not hit,
                                                                 use_cached_backbone=hit)['image_depth']
                        self._stable_key = key
                    else:
                        stable_depth = self.original_G.synthesis(new_ws, new_camera, noise_mode='const', need_image=True)['image_depth']
                loss_depth = l2_loss(stable_depth, sample_depth) * hp.pt_depth_lambda
                if share:
                    loss = loss + loss_depth
                else:
                    loss_depth.backward()
            if hp.pt_tv_lambda > 0:
                loss_tv = cal_tv_loss(w_pivot, self.G) * hp.pt_tv_lambda
                if share:
                    loss = loss + loss_tv
                else:
                    loss_tv.backward()
        if share:
            loss.backward()
        with torch.no_grad():
            self._lpips_out.copy_(loss_lpips.detach())
        self.optimizer.skip_if_le = (self._lpips_out, hp.LPIPS_value_threshold)
        self.optimizer.step(in_graph=True)
        return self._lpips_out

    def train_step(self, i, st, w_pivot, rot_bs=4):
        """One iteration; returns (loss_lpips, stepped).  Iterations are replayed as CUDA graphs (one for the plain iteration,
        one for the i % 4 == 0 iteration) unless disabled or a test has injected random draws."""
        heavy = (i % rot_bs == 0)
        if not hasattr(self, '_lpips_out') or self._lpips_out.device != w_pivot.device:
            self._lpips_out = torch.zeros((), device=w_pivot.device)
            self._graphs = {}
        if hasattr(self.lpips_loss, 'register_target') and getattr(self, '_registered', None) is not st.image:
            self.lpips_loss.register_target(st.image)          # the inverted image is constant: cache its LPIPS taps
            self._registered = st.image
        if self.optimizer.hyper is None:
            self.optimizer.use_device_hyper()
        self.optimizer._reseat()
        self.optimizer.advance()
        eager = (not global_config.use_cuda_graphs) or rng.pending() or bool(self.G.renderer._noise_queue) or bool(self.original_G.renderer._noise_queue)
        if eager:
            self._body(heavy, st, w_pivot, rot_bs)
        else:
            key = (heavy, id(st), id(w_pivot), id(self.G))
            if key not in self._graphs:
                state = [self.optimizer.arena, self.optimizer.exp_avg, self.optimizer.exp_avg_sq]
                self._graphs[key] = (GraphedStep(lambda: self._body(heavy, st, w_pivot, rot_bs), state), st, w_pivot)
            self._graphs[key][0]()
        loss_lpips = self._lpips_out
        stepped = bool(loss_lpips > hyperparameters.LPIPS_value_threshold)      # one host read per iteration, as the reference's `if`
        if not stepped:
            self.optimizer.steps -= 1
        return loss_lpips, stepped

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
            lm = data['lm'].to(global_config.device)
            st = SPIState(image, camera, mask, lm)
            paths_config.experiments_output_dir = os.path.join(output_dir, image_name)
            os.makedirs(paths_config.experiments_output_dir, exist_ok=True)
            self.restart_training()
            w_pivot = self.get_inversion(image_name, image, camera, fg_mask=st.fg_mask)
            for i in range(hyperparameters.G_1_step):
                _, stepped = self.train_step(i, st, w_pivot)
                if not stepped:
                    break
                global_config.training_step += 1
            self.image_counter += 1
            self.finish_image(w_pivot, image, camera, image_name)
        paths_config.experiments_output_dir = output_dir
        if self.use_wandb:
            self.log_metric()
