"""Golden vectors at the BENCH configuration (32 coarse + 32 importance samples per ray): the reference's own `TriPlaneGenerator.synthesis`
and one PTI optimisation step (pti_coach.py:62-82) run on the CPU with `rendering_kwargs['depth_resolution'(_importance)] = 32`, the way
the reference reads them from the pickle (load_utils.py:27-28).  tests/golden/{synthesis_32,steps_32}.npz; same conventions as make_golden.py.

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden_32
"""
import copy
import json
import os
import time

import numpy as np
import torch

from . import generator, loops, ref_shim, weights
from .make_golden import OUT, REPORT, build_ref_losses, injected_rng, make_nets, note, npy

DEPTH = (32, 32)


def main():
    assert ref_shim.available(), 'reference tree not found'
    ref_shim.install()
    torch.manual_seed(0)
    sd = weights.generator_state_dict(0)
    G = ref_shim.build_reference_generator()
    G.load_state_dict(sd, strict=True)
    G.rendering_kwargs = dict(G.rendering_kwargs, depth_resolution=DEPTH[0], depth_resolution_importance=DEPTH[1])
    rk = {**generator.RENDERING_DEFAULTS, **G.rendering_kwargs}
    nets = make_nets()
    # ---- synthesis
    g = {}
    ws, c = weights.w_pivot(5), weights.canonical_camera(0.3)
    jit, u = generator.make_render_noise(1, 128 * 128, rk, seed=7)
    t = time.time()
    with injected_rng(rand_like=[jit], rand=[u]):
        ref = G.synthesis(ws, c, noise_mode='const', cache_backbone=True)
    print('reference synthesis (32+32) %.1fs' % (time.time() - t))
    mine = generator.synthesis(sd, ws, c, rk, jitter=jit, u=u)
    for k in ('image', 'image_raw', 'image_depth'):
        note(f'synthesis32/{k}', ref[k], mine[k])
    g['ws'], g['c'] = npy(ws), npy(c)
    g['image_raw'], g['image_depth'] = npy(ref['image_raw']), npy(ref['image_depth'])
    g['image_sub'] = npy(ref['image'][:, :, 1::4, 2::4])
    g['image_sqsum'] = np.float64(ref['image'].double().square().sum().item())
    np.savez_compressed(os.path.join(OUT, 'synthesis_32.npz'), **g)
    # ---- one PTI step (pti_coach.py:62-82) with the reference's own modules
    from criteria.l2_loss import l2_loss
    L, _ = build_ref_losses(nets)
    target = weights.target_image()
    parsing, lm = weights.parsing_mask(), weights.landmarks68()
    Gt = copy.deepcopy(G).requires_grad_(True)
    wp = weights.w_pivot(5).requires_grad_(True)
    src, rep = loops.NoiseSource(200), loops.NoiseSource(200)
    coach = loops.Coach(sd, weights.w_pivot(5), target, c, parsing, lm, nets, kind='pti', rk=rk, noise=src)
    info = coach.step(0)
    jit, u = rep.render(1, 128 * 128, rk)
    with injected_rng(rand_like=[jit], rand=[u]):
        out = Gt.synthesis(wp, c, noise_mode='const')
    l2v = l2_loss(out['image'], target)
    lp = torch.squeeze(L(out['image'], target))
    (l2v + lp).backward()
    grads = {k: p.grad.clone() for k, p in Gt.named_parameters() if p.grad is not None}
    s = {'pti_l2': np.float64(float(l2v)), 'pti_lpips': np.float64(float(lp)), 'pti_wgrad': npy(wp.grad)}
    REPORT['pti32/ref_info'] = {'l2': float(l2v), 'lpips': float(lp)}
    REPORT['pti32/oracle_info'] = {k: v for k, v in info.items() if k != 'early_exit'}
    for k in ('decoder.net.0.weight', 'decoder.net.2.bias', 'superresolution.block1.conv1.weight', 'backbone.synthesis.b4.const',
              'backbone.synthesis.b64.conv0.affine.weight', 'backbone.synthesis.b256.torgb.weight'):
        note(f'pti32/grad/{k}', grads[k], coach.sd[k].grad)
        s[f'pti_grad_{k}'] = npy(grads[k].reshape(-1)[::max(1, grads[k].numel() // 4096)])
        s[f'pti_gradnorm_{k}'] = np.float64(grads[k].double().norm().item())
    note('pti32/grad/ws', wp.grad, coach.w.grad)
    np.savez_compressed(os.path.join(OUT, 'steps_32.npz'), **s)
    path = os.path.join(OUT, 'REPORT.json')
    old = json.load(open(path)) if os.path.exists(path) else {}
    old.update(REPORT)
    json.dump(old, open(path, 'w'), indent=1, sort_keys=True)
    print('wrote', path)


if __name__ == '__main__':
    main()
