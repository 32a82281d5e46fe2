"""Stage-1 latent optimisation shared by the three projectors (spi/training/projectors/{w,w_plus,mirror}_projector.py).

One class, `LatentProjector`, holds the state of the loop and exposes `step(i)`; the three public `project()` functions
keep the reference signatures.  Behaviours kept on purpose: the learning rate written every step is
`initial_learning_rate=0.01` (hyperparameters.first_inv_lr is never effective), `w_avg_samples` z-draws come from
`RandomState(123)`, `bg_loss` of the mirror projector is computed by the reference but never used (dropped here), noise
buffers are re-normalised after every step.
"""
import copy

import numpy as np
import torch
import torch.nn.functional as F

from ...configs import global_config
from ...graphs import GraphedStep
from ...ops import zero_arena
from ...optim import FlatAdam
from ...ops import noise_reg
from ...ops.resize import downsample2x
from ...utils import rng
from ...utils.camera_utils import cal_camera_weight, cal_mirror_c


def noise_regulariser(noise_bufs):
    """mirror_projector.py:107-115 -- one fused launch for all buffers (spi_b200/ops/noise_reg.py)."""
    return noise_reg.noise_regulariser(list(noise_bufs))


def area_256(img):
    """F.interpolate(size=(256,256), mode='area') (w_projector.py:50,83)."""
    if img.shape[2] == 512:
        return downsample2x(img)
    return F.interpolate(img, size=(256, 256), mode='area') if img.shape[2] > 256 else img


class LatentProjector:
    def __init__(self, G, target, c, kind, *, lpips_func=None, vgg16=None, initial_w=None, num_steps=1000, w_avg_samples=10000,
                 initial_learning_rate=0.01, initial_noise_factor=0.05, lr_rampdown_length=0.25, lr_rampup_length=0.05,
                 noise_ramp_length=0.75, regularize_noise_weight=1e5, device=None):
        assert kind in ('sg', 'sgw+', 'mir')
        assert target.shape[1:] == (G.img_channels, G.img_resolution, G.img_resolution)
        device = device or torch.device(global_config.device)
        self.kind, self.num_steps = kind, num_steps
        self.hp = dict(lr0=initial_learning_rate, noise0=initial_noise_factor, down=lr_rampdown_length, up=lr_rampup_length,
                       nramp=noise_ramp_length, regw=regularize_noise_weight)
        self.G = G = copy.deepcopy(G).eval().requires_grad_(False).to(device).float()
        # w statistics (*_projector.py:36-44)
        z = np.random.RandomState(123).randn(w_avg_samples, G.z_dim)
        w_samples = G.mapping(torch.from_numpy(z).to(device), c.repeat(w_avg_samples, 1))
        w_samples = w_samples[:, :1, :].cpu().numpy().astype(np.float32)
        w_avg = np.mean(w_samples, axis=0, keepdims=True)
        self.w_std = (np.sum((w_samples - w_avg) ** 2) / w_avg_samples) ** 0.5
        self.num_ws = G.backbone.mapping.num_ws
        if initial_w is not None:
            start_w = initial_w
        else:
            start_w = w_avg if kind == 'sg' else np.repeat(w_avg, self.num_ws, axis=1)
        self.noise_bufs = {name: buf for (name, buf) in G.backbone.synthesis.named_buffers() if 'noise_const' in name}
        self.w_opt = torch.tensor(start_w, dtype=torch.float32, device=device, requires_grad=True)
        for buf in self.noise_bufs.values():
            buf[:] = rng.randn_like(buf)
            buf.requires_grad = True
        self.optimizer = FlatAdam([self.w_opt] + list(self.noise_bufs.values()), betas=(0.9, 0.999), lr=initial_learning_rate, steal_grads=True)
        self.optimizer.use_device_hyper()
        self.w_noise_scale = torch.zeros((), device=device)          # per-step scalar, read on device (graph replay)
        self._graph = None
        self._out = dict(loss=torch.zeros((), device=device), dist=torch.zeros((), device=device), image=None)
        self.target, self.c = target, c
        self.lpips_func, self.vgg16 = lpips_func, vgg16
        if lpips_func is not None and hasattr(lpips_func, 'register_target'):
            lpips_func.register_target(target)
        if kind == 'mir':
            self.target_m = torch.flip(target, dims=[3])
            if hasattr(lpips_func, 'register_target'):
                lpips_func.register_target(self.target_m)
            camera_m = cal_mirror_c(camera=c)
            self.target_camera = torch.cat([c, camera_m], dim=0)
            self.weight_m = cal_camera_weight(camera_m)[0]
            self._pair_weights = torch.stack([torch.ones((), device=device), self.weight_m.reshape(()).to(device=device, dtype=torch.float32)])
        if kind == 'sg':
            with torch.no_grad():
                self.target_features = vgg16(area_256((target + 1) * (255 / 2)), resize_images=False, return_lpips=True)

    def schedule(self, step):
        """mirror_projector.py:84-91 (host float64)."""
        hp = self.hp
        t = step / self.num_steps
        w_noise_scale = self.w_std * hp['noise0'] * max(0.0, 1.0 - t / hp['nramp']) ** 2
        lr_ramp = min(1.0, (1.0 - t) / hp['down'])
        lr_ramp = 0.5 - 0.5 * np.cos(lr_ramp * np.pi)
        lr_ramp = lr_ramp * min(1.0, t / hp['up'])
        return hp['lr0'] * lr_ramp, w_noise_scale

    def _body(self):
        # one zero-filled buffer per iteration for every accumulate-into output (ops/zero_arena.py)
        with zero_arena.iteration(('projector', self.kind, tuple(self.w_opt.shape)), self.w_opt.device):
            return self._iteration()

    def _iteration(self):
        """One iteration, free of host synchronisation (capturable): mirror_projector.py:93-131."""
        ws = self.w_opt + rng.randn_like(self.w_opt) * self.w_noise_scale
        G = self.G
        if self.kind == 'mir':
            # both views use the same latent: an expanded view lets the generator evaluate the tri-plane backbone once
            ws2 = ws.expand(2, -1, -1) if global_config.share_backbone else ws.repeat(2, 1, 1)
            out = G.synthesis(ws2, self.target_camera, noise_mode='const')
            img = out['image']
            if hasattr(self.lpips_func, 'weighted_pairs'):
                # lpips(img, target) + weight_m * lpips(img_m, target_m) (mirror_projector.py:100-104) with one pass of the VGG trunk
                dist = self.lpips_func.weighted_pairs(img, (self.target, self.target_m), self._pair_weights)
            else:
                dist = self.lpips_func(img[:1], self.target) + self.lpips_func(img[1:], self.target_m) * self.weight_m
        elif self.kind == 'sgw+':
            img = G.synthesis(ws, self.c, noise_mode='const')['image']
            dist = self.lpips_func(img, self.target)
        else:
            img = G.synthesis(ws.repeat([1, self.num_ws, 1]), self.c, noise_mode='const')['image']
            feats = self.vgg16(area_256((img + 1) * (255 / 2)), resize_images=False, return_lpips=True)
            dist = (self.target_features - feats).square().sum()
        reg_loss = noise_regulariser(self.noise_bufs.values())
        loss = dist + reg_loss * self.hp['regw']
        self.optimizer.zero_grad(set_to_none=True)
        loss.backward()
        self.optimizer.step(in_graph=True)
        with torch.no_grad():
            noise_reg.renormalise_noise_(self.noise_bufs.values())         # mirror_projector.py:128-131
            self._out['loss'].copy_(loss.detach())
            self._out['dist'].copy_(dist.detach())
            if self._out['image'] is None:
                self._out['image'] = torch.empty_like(img)
            self._out['image'].copy_(img.detach())
        return self._out

    def step(self, step):
        lr, w_noise_scale = self.schedule(step)
        for g in self.optimizer.param_groups:
            g['lr'] = lr
        self.w_noise_scale.fill_(float(w_noise_scale))        # value travels as a kernel argument (no reused pinned buffer in flight)
        self.optimizer._reseat()
        self.optimizer.advance()
        eager = (not global_config.use_cuda_graphs) or rng.pending() or bool(self.G.renderer._noise_queue)
        if eager:
            self.last = self._body()
        else:
            if self._graph is None:
                st = [self.optimizer.arena, self.optimizer.exp_avg, self.optimizer.exp_avg_sq]
                self._graph = GraphedStep(self._body, st)
            self.last = self._graph()
        return self.last

    def result(self):
        return self.w_opt.repeat([1, self.num_ws, 1]) if self.kind == 'sg' else self.w_opt
