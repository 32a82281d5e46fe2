"""Pins the oracle's `sgw+` stage-1 projector (oracle/loops.py, kind='sgw+') to the reference's own
spi/training/projectors/w_plus_projector.py:30-120: two optimisation steps of num_steps = 500 on identical weights,
inputs and random draws, on the CPU.

TEST INFRASTRUCTURE ONLY; runs in the build container (needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden_sgw

Writes tests/golden/sgw_plus.npz (the reference's latent after the two steps) and the `sgw+/*` rows of
tests/golden/REPORT.json.  (`sg` cannot be pinned the same way: it calls the third-party `vgg16.pt` TorchScript artefact,
SURVEY.md §8c.)"""
import copy
import json
import os

import numpy as np
import torch

from . import generator, loops, ref_shim, weights
from .make_golden import OUT, REPORT, build_ref_losses, injected_rng, make_nets, note, npy


def main():
    assert ref_shim.available(), 'reference tree not found'
    ref_shim.install()
    torch.manual_seed(0)
    sd = weights.generator_state_dict(0)
    G = ref_shim.build_reference_generator()
    G.load_state_dict(sd, strict=True)
    nets = make_nets()
    from spi.configs import global_config
    from spi.training.projectors import w_plus_projector
    global_config.device = 'cpu'
    L, _ = build_ref_losses(nets)
    target, c = weights.target_image(), weights.canonical_camera(0.3)
    rk = {**generator.RENDERING_DEFAULTS, **G.rendering_kwargs}
    nsteps = 2
    mine = loops.Projector(sd, target, c, nets, kind='sgw+', num_steps=500, rk=rk, noise=loops.NoiseSource(300))
    # the same draws for the reference: noise-buffer init (randn_like x13), then per step randn_like(w_opt), rand_like, rand
    rep = loops.NoiseSource(300)
    bufs = [rep.randn(*sd[k].shape) for k in loops.noise_buffer_names(sd)]
    per_step = []
    for _ in range(nsteps):
        wn = rep.randn(1, 14, 512)
        jit, u = rep.render(1, 128 * 128, rk)
        per_step.append((wn, jit, u))
    w_plus_projector.tqdm = lambda it, *a, **k: it
    real_range = range
    w_plus_projector.range = lambda n: real_range(nsteps)          # stop after `nsteps`; the schedules still see num_steps = 500
    for name in ('log_image', 'log_image_from_w'):
        if hasattr(w_plus_projector, name):
            setattr(w_plus_projector, name, lambda *a, **k: None)
    with injected_rng(randn_like=bufs + [s[0] for s in per_step], rand_like=[s[1] for s in per_step], rand=[s[2] for s in per_step]):
        w_ref = w_plus_projector.project(copy.deepcopy(G), target, c, lpips_func=L, device=torch.device('cpu'), w_avg_samples=600,
                                         num_steps=500, w_name='g')
    infos = [mine.step(i) for i in range(nsteps)]
    note('sgw+/w_opt', w_ref.detach(), mine.result())
    REPORT['sgw+/oracle_losses'] = [i['loss'] for i in infos]
    np.savez_compressed(os.path.join(OUT, 'sgw_plus.npz'), w=npy(w_ref.detach()), loss=np.array([i['loss'] for i in infos]),
                        w_std=np.float64(mine.w_std))
    path = os.path.join(OUT, 'REPORT.json')
    old = json.load(open(path))
    old.update({k: v for k, v in REPORT.items() if k.startswith('sgw+/')})
    json.dump(old, open(path, 'w'), indent=1, sort_keys=True)
    print('sgw+ pinned:', {k: v for k, v in REPORT.items() if k.startswith('sgw+/')})


if __name__ == '__main__':
    main()
