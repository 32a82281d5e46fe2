"""Oracle (CPU, test infrastructure): one iteration of each optimisation loop.

Follows:
  stage-1 projectors   spi/training/projectors/w_projector.py:30-113 ('sg'),
                       w_plus_projector.py ('sgw+'), mirror_projector.py:34-137 ('mir')
  stage-2 PTI          spi/training/coaches/pti_coach.py:17-28,62-74
  stage-2 SPI          spi/training/coaches/rot_bbox_cx_coach.py:55-151
  optimiser            torch.optim.Adam defaults (base_coach.py:132-135; *_projector.py:55-58)
All random draws come from a `NoiseSource`, so the CUDA product can be fed the very same tensors.
"""
import copy
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import criteria, generator, geometry


class NoiseSource:
    """Seeded stream of the random tensors an iteration consumes, in call order."""

    def __init__(self, seed=0):
        self.g = torch.Generator().manual_seed(seed)

    def rand(self, *shape):
        return torch.rand(*shape, generator=self.g)

    def randn(self, *shape):
        return torch.randn(*shape, generator=self.g)

    def render(self, n, r, rk):
        """(jitter, u) of one ImportanceRenderer.forward (renderer.py:190,237)."""
        return self.rand(n, r, rk['depth_resolution'], 1), self.rand(n * r, max(rk['depth_resolution_importance'], 1))


def lr_schedule(step, num_steps, w_std, initial_learning_rate=0.01, initial_noise_factor=0.05,
                lr_rampdown_length=0.25, lr_rampup_length=0.05, noise_ramp_length=0.75):
    """mirror_projector.py:84-91 (host float64): returns (lr, w_noise_scale)."""
    t = step / num_steps
    w_noise_scale = w_std * initial_noise_factor * max(0.0, 1.0 - t / noise_ramp_length) ** 2
    ramp = min(1.0, (1.0 - t) / lr_rampdown_length)
    ramp = 0.5 - 0.5 * np.cos(ramp * np.pi)
    ramp = ramp * min(1.0, t / lr_rampup_length)
    return initial_learning_rate * ramp, w_noise_scale


def w_statistics(sd, c, rk, w_avg_samples=600):
    """*_projector.py:36-44: RandomState(123) z samples -> w_avg [1,1,512], w_std scalar."""
    z = np.random.RandomState(123).randn(w_avg_samples, 512)
    w = generator.mapping(sd, torch.from_numpy(z), c.repeat(w_avg_samples, 1), rk)
    w = w[:, :1, :].numpy().astype(np.float32)
    w_avg = np.mean(w, axis=0, keepdims=True)
    w_std = (np.sum((w - w_avg) ** 2) / w_avg_samples) ** 0.5
    return w_avg, float(w_std)


NOISE_KEYS_SUFFIX = 'noise_const'


def noise_buffer_names(sd):
    return [k for k in sd if k.startswith('backbone.synthesis.') and k.endswith(NOISE_KEYS_SUFFIX)]


class Projector:
    """Stage-1 latent optimisation; kind in {'sg','sgw+','mir'}."""

    def __init__(self, sd, target, c, nets, kind='mir', num_steps=500, rk=None, noise=None, w_avg_samples=600,
                 regularize_noise_weight=1e5):
        self.rk = {**generator.RENDERING_DEFAULTS, **(rk or {})}
        self.kind, self.num_steps, self.nets = kind, num_steps, nets
        self.noise = noise or NoiseSource(0)
        self.reg_w = regularize_noise_weight
        self.sd = {k: v.clone() for k, v in sd.items()}          # deepcopy(G) (:34)
        w_avg, self.w_std = w_statistics(self.sd, c, self.rk, w_avg_samples)
        start = w_avg if kind == 'sg' else np.repeat(w_avg, 14, axis=1)
        self.w_opt = torch.tensor(start, dtype=torch.float32, requires_grad=True)
        self.noise_names = noise_buffer_names(self.sd)
        for k in self.noise_names:                                 # :61-63
            self.sd[k] = self.noise.randn(*self.sd[k].shape).requires_grad_(True)
        self.opt = torch.optim.Adam([self.w_opt] + [self.sd[k] for k in self.noise_names], betas=(0.9, 0.999), lr=5e-3)
        self.target, self.c = target, c
        if kind == 'mir':
            self.target_m = torch.flip(target, dims=[3])
            self.c2 = torch.cat([c, geometry.mirror_camera(c)], 0)
            self.weight_m = geometry.camera_weight(geometry.mirror_camera(c))[0]
        if kind == 'sg':
            t = (target + 1) * (255 / 2)
            t = F.interpolate(t, size=(256, 256), mode='area')
            self.target_features = criteria.vgg16_pt_features(t, nets['vgg16'], nets['lin'])

    def step(self, i):
        lr, w_noise_scale = lr_schedule(i, self.num_steps, self.w_std)
        for g in self.opt.param_groups:
            g['lr'] = lr
        ws = self.w_opt + self.noise.randn(*self.w_opt.shape) * w_noise_scale
        if self.kind == 'sg':
            ws = ws.repeat(1, 14, 1)
        if self.kind == 'mir':
            ws = ws.repeat(2, 1, 1)
            jit, u = self.noise.render(2, 128 * 128, self.rk)
            out = generator.synthesis(self.sd, ws, self.c2, self.rk, jitter=jit, u=u)
            img = out['image']
            dist = criteria.lpips(img[:1], self.target, self.nets['vgg16'], self.nets['lin']) + \
                criteria.lpips(img[1:], self.target_m, self.nets['vgg16'], self.nets['lin']) * self.weight_m
        else:
            jit, u = self.noise.render(1, 128 * 128, self.rk)
            img = generator.synthesis(self.sd, ws, self.c, self.rk, jitter=jit, u=u)['image']
            if self.kind == 'sg':
                s = F.interpolate((img + 1) * (255 / 2), size=(256, 256), mode='area')
                dist = (self.target_features - criteria.vgg16_pt_features(s, self.nets['vgg16'], self.nets['lin'])).square().sum()
            else:
                dist = criteria.lpips(img, self.target, self.nets['vgg16'], self.nets['lin'])
        reg = criteria.noise_regulariser([self.sd[k] for k in self.noise_names])
        loss = dist + reg * self.reg_w
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        self.opt.step()
        criteria.renormalise_noise_([self.sd[k] for k in self.noise_names])
        return {'loss': float(loss), 'dist': float(dist), 'reg': float(reg), 'lr': lr, 'image': img.detach()}

    def result(self):
        return self.w_opt.detach().repeat(1, 14, 1) if self.kind == 'sg' else self.w_opt.detach()


PARAM_SKIP = ('resample_filter', 'noise_const', 'w_avg')


def trainable_names(sd):
    """G.parameters() (base_coach.py:134): everything but the registered buffers."""
    return [k for k in sd if not k.endswith(PARAM_SKIP)]


class Coach:
    """Stage-2 generator fine-tuning; kind in {'pti','RotBbox'}."""

    def __init__(self, sd, w_pivot, image, camera, parsing, lm, nets, kind='RotBbox', rk=None, noise=None,
                 lambdas=None, lr=3e-4):
        self.rk = {**generator.RENDERING_DEFAULTS, **(rk or {})}
        self.kind, self.nets = kind, nets
        self.noise = noise or NoiseSource(0)
        self.lam = dict(l2=1.0, lpips=1.0, rot=0.1, mirror=0.05, depth=1.0)
        self.lam.update(lambdas or {})
        self.sd = {k: v.clone() for k, v in sd.items()}
        self.original_sd = {k: v.clone() for k, v in sd.items()}
        self.params = trainable_names(self.sd)
        for k in self.params:
            self.sd[k].requires_grad_(True)
        self.opt = torch.optim.Adam([self.sd[k] for k in self.params], lr=lr)
        self.w = w_pivot.clone().requires_grad_(True)              # leaf with grad, never stepped (§3.5)
        self.image, self.camera, self.lm = image, camera, lm
        self.face_mask = geometry.face_mask(parsing).float()    # parsing: [1,1,H,W] = data['mask'][:, 0]
        self.image_m = torch.flip(image, dims=[3])
        self.face_mask_m = torch.flip(self.face_mask, dims=[3])
        self.camera_m = geometry.mirror_camera(camera)
        self.weight_m = geometry.camera_weight(camera)
        self.yaw_range = 0.2

    def _synth(self, sd, ws, c):
        jit, u = self.noise.render(ws.shape[0], 128 * 128, self.rk)
        return generator.synthesis(sd, ws, c, self.rk, jitter=jit, u=u)

    def _lpips(self, a, b):
        return criteria.lpips(a, b, self.nets['vgg16'], self.nets['lin'])

    def step(self, i):
        info = {}
        self.opt.zero_grad()
        out = self._synth(self.sd, self.w, self.camera)
        img, depth = out['image'], out['image_depth']
        l2v = criteria.l2(img, self.image)
        lp = self._lpips(img, self.image)
        loss = l2v * self.lam['l2'] + lp * self.lam['lpips']
        loss.backward()
        info.update(l2=float(l2v), lpips=float(lp))
        if self.kind == 'RotBbox' and i % 4 == 0:
            bs = 4
            if self.lam['rot'] > 0:
                cams = geometry.sample_surrounding_camera(self.camera, self.noise.rand(bs, 2), self.yaw_range, 0.1)
                gen = self._synth(self.sd, self.w.repeat(bs, 1, 1), cams)
                with torch.no_grad():
                    warp, wmask = geometry.rotate(cams, gen['image_depth'], self.image.repeat(bs, 1, 1, 1),
                                                  self.camera.repeat(bs, 1), depth.repeat(bs, 1, 1, 1),
                                                  self.face_mask.repeat(bs, 1, 1, 1), eps=5e-2)
                lrot = self._lpips(gen['image'] * wmask, warp) * self.lam['rot'] * bs
                lrot.backward()
                info['rot'] = float(lrot)
            if self.lam['mirror'] > 0 and self.weight_m > 0:
                cams = geometry.sample_surrounding_camera(self.camera_m, self.noise.rand(bs, 2), self.yaw_range, 0.1)
                gen = self._synth(self.sd, self.w.repeat(bs, 1, 1), cams)
                with torch.no_grad():
                    depth_m = torch.flip(depth, dims=[3])
                    warp, wmask = geometry.rotate(cams, gen['image_depth'], self.image_m.repeat(bs, 1, 1, 1),
                                                  self.camera_m.repeat(bs, 1), depth_m.repeat(bs, 1, 1, 1),
                                                  self.face_mask_m.repeat(bs, 1, 1, 1), eps=5e-2)
                    warp, wmask = torch.flip(warp, dims=[3]), torch.flip(wmask, dims=[3])
                lmir = criteria.box_cx(torch.flip(gen['image'], dims=[3]) * wmask, warp, self.lm.repeat(bs, 1, 1),
                                       self.nets['vgg19']) * self.lam['mirror'] * bs
                lmir.backward()
                info['mirror'] = float(lmir)
            if self.lam['depth'] > 0:
                cams = geometry.sample_camera(self.noise.rand(4, 2), yaw_range=0.7, pitch_range=0.4)
                ws4 = self.w.repeat(4, 1, 1)
                jit, u = self.noise.render(4, 128 * 128, self.rk)
                d_new = generator.synthesis(self.sd, ws4, cams, self.rk, jitter=jit, u=u)['image_depth']
                with torch.no_grad():
                    jit, u = self.noise.render(4, 128 * 128, self.rk)
                    d_ref = generator.synthesis(self.original_sd, ws4, cams, self.rk, jitter=jit, u=u)['image_depth']
                ld = criteria.l2(d_ref, d_new) * self.lam['depth']
                ld.backward()
                info['depth'] = float(ld)
        info['early_exit'] = bool(lp <= 0.05)                      # rot_bbox_cx_coach.py:148-149
        if not info['early_exit']:
            self.opt.step()
        return info
