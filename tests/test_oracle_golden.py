"""CPU: the oracle restatement against the committed outputs of the reference (tests/golden/, minted by
oracle/make_golden.py from /root/reference).  Tolerance 1e-5 rel-L2 covers thread-order noise (5e-7 measured)."""
import json
import math
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_l2
from oracle import criteria, generator, geometry, loops, ops, weights

TOL = 1e-5


def T(a):
    return torch.from_numpy(np.asarray(a))


def test_report_says_oracle_matches_reference():
    rep = json.load(open(os.path.join(GOLDEN, 'REPORT.json')))
    worst = 0.0
    for k, v in rep.items():
        if isinstance(v, dict) and 'rel_l2' in v:
            worst = max(worst, v['rel_l2'])
        if isinstance(v, dict) and 'exact' in v:
            assert v['exact'], k
    assert worst < 1e-5, worst


def test_bias_act(golden):
    g = golden('ops')
    x, b = T(g['ba_x']), T(g['ba_b'])
    for act in ops.ACTS:
        assert rel_l2(ops.bias_act(x, b, act=act), g[f'ba_{act}_d']) < TOL
        assert rel_l2(ops.bias_act(x, b, act=act, gain=0.7, clamp=0.9, alpha=0.3), g[f'ba_{act}_g']) < TOL


def test_upfirdn2d(golden):
    g = golden('ops')
    rep = json.load(open(os.path.join(GOLDEN, 'REPORT.json')))['upfirdn2d_cases']
    x, f = T(g['up_x']), T(g['f4'])
    assert torch.equal(ops.setup_filter([1, 3, 3, 1]), f)
    for name, kw in rep.items():
        assert rel_l2(ops.upfirdn2d(x, f, **kw), g['up_' + name]) < TOL, name
    xs, f12 = T(g['up_xs']), T(g['f12'])
    assert rel_l2(ops.upfirdn2d(xs, f12, up=2, padding=[5, 6, 5, 6], gain=4.0), g['up_separable12']) < TOL


def test_filtered_lrelu(golden):
    g = golden('ops')
    xs, f12, fd12, b = T(g['up_xs']), T(g['f12']), T(g['fd12']), T(g['fl_b'])
    cases = {
        'u2d2': dict(up=2, down=2, padding=[9, 10, 9, 10], gain=math.sqrt(2), slope=0.2, clamp=256., flip_filter=False),
        'u2d1': dict(up=2, down=1, padding=[5, 6, 5, 6], gain=math.sqrt(2), slope=0.2, clamp=None, flip_filter=False),
        'u1d2': dict(up=1, down=2, padding=[5, 5, 5, 5], gain=1.1, slope=0.1, clamp=0.8, flip_filter=True),
        'u1d1': dict(up=1, down=1, padding=[0, 0, 0, 0], gain=math.sqrt(2), slope=0.2, clamp=None, flip_filter=False),
    }
    for name, kw in cases.items():
        fu = f12 if kw['up'] > 1 else (None if name == 'u1d1' else f12)
        fd = fd12 if kw['down'] > 1 else None
        assert rel_l2(ops.filtered_lrelu(xs, fu=fu, fd=fd, b=b, **kw), g['fl_' + name]) < TOL, name


def test_renderer_pieces(golden):
    g = golden('render')
    cam = T(g['cam'])
    o, d = generator.ray_sampler(cam[:, :16].reshape(-1, 4, 4), cam[:, 16:].reshape(-1, 3, 3), 128)
    assert rel_l2(o[:, ::37], g['ray_origins_sub']) < TOL and rel_l2(d[:, ::37], g['ray_dirs_sub']) < TOL
    sd = weights.generator_state_dict(0)
    planes, sel = T(g['planes']), T(g['sel'])
    oo, dd = o[:, sel].contiguous(), d[:, sel].contiguous()
    for dc, df in ((48, 48), (32, 32), (12, 20)):
        tag = f'r{dc}_{df}_'
        rk = dict(generator.RENDERING_DEFAULTS, depth_resolution=dc, depth_resolution_importance=df)
        rgb, depth, wsum, aux = generator.importance_render(sd, planes, oo, dd, rk, T(g[tag + 'jit']), T(g[tag + 'u']), return_aux=True)
        assert rel_l2(rgb, g[tag + 'rgb']) < TOL and rel_l2(depth, g[tag + 'depth']) < TOL and rel_l2(wsum, g[tag + 'wsum']) < TOL
        assert torch.equal(aux['inds'], T(g[tag + 'inds']))
        assert torch.equal(aux['perm'], T(g[tag + 'perm']))
    r1, r2, r3 = generator.ray_march(T(g['rm_col']), T(g['rm_sig']), T(g['rm_dep']), generator.RENDERING_DEFAULTS)
    assert rel_l2(r1, g['rm_rgb']) < TOL and rel_l2(r2, g['rm_depth']) < TOL and rel_l2(r3, g['rm_w']) < TOL


def test_mapping_and_synthesis(golden):
    g = golden('synthesis')
    sd = weights.generator_state_dict(0)
    w = generator.mapping(sd, T(g['z']), T(g['c3']))
    assert rel_l2(w[:, 0], g['w']) < TOL
    rk = generator.RENDERING_DEFAULTS
    jit, u = generator.make_render_noise(1, 128 * 128, rk, seed=7)
    out = generator.synthesis(sd, T(g['ws']), T(g['c']), rk, jitter=jit, u=u)
    assert rel_l2(out['image_raw'], g['image_raw']) < TOL
    assert rel_l2(out['image_depth'], g['image_depth']) < TOL
    assert rel_l2(out['image'][:, :, 1::4, 2::4], g['image_sub']) < TOL
    assert rel_l2(out['planes'][:, ::7, 3::8, 5::8], g['planes_sub']) < TOL
    assert abs(out['image'].double().square().sum().item() / float(g['image_sqsum']) - 1) < 1e-5


def test_geometry_and_losses(golden):
    g = golden('geometry_losses')
    c = weights.canonical_camera(0.3)
    r = T(g['rand42'])
    assert rel_l2(geometry.sample_surrounding_camera(c, r, 0.2, 0.1), g['surround']) < TOL
    assert rel_l2(geometry.sample_camera(r, 0.7, 0.4), g['sampled']) < TOL
    assert rel_l2(geometry.camera_weight(c), g['cam_weight']) < TOL
    parsing = weights.parsing_mask()
    fm = geometry.face_mask(parsing).float()
    assert int(fm.sum()) == int(g['face_mask_sum'])
    sc = T(g['surround'])
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, 128), torch.linspace(-1, 1, 128), indexing='ij')
    base = 2.7 - 0.35 * torch.exp(-(xx ** 2 + yy ** 2) * 2.5)
    sdepth = base[None, None].repeat(4, 1, 1, 1)
    img = weights.target_image().repeat(4, 1, 1, 1)
    wr, wm = geometry.rotate(sc, T(g['rot_tdepth']), img, c.repeat(4, 1), sdepth, fm.repeat(4, 1, 1, 1), eps=5e-2)
    assert rel_l2(wr[:, :, 2::8, 3::8], g['rot_rgb_sub']) < TOL and rel_l2(wm[:, :, 2::8, 3::8], g['rot_mask_sub']) < TOL
    lm = weights.landmarks68().repeat(2, 1, 1)
    for k, bx in enumerate(criteria.landmark_boxes(lm)):
        assert torch.equal(bx, T(g[f'boxes{k}']))


@pytest.mark.slow
def test_pti_step_matches_reference(golden):
    g = golden('steps')
    from oracle.make_golden import make_nets
    sd = weights.generator_state_dict(0)
    nets = make_nets()
    coach = loops.Coach(sd, weights.w_pivot(5), weights.target_image(), weights.canonical_camera(0.3), weights.parsing_mask(),
                        weights.landmarks68(), nets, kind='pti', noise=loops.NoiseSource(200))
    info = coach.step(0)
    assert abs(info['l2'] / float(g['pti_l2']) - 1) < 1e-4 and abs(info['lpips'] / float(g['pti_lpips']) - 1) < 1e-4
    assert rel_l2(coach.w.grad, g['pti_wgrad']) < 1e-4


def test_id_similarity(golden):
    from oracle import idloss
    g = golden('idloss')
    sd = weights.irse50_state_dict(3)
    x = weights.target_image()
    y = torch.flip(weights.target_image(seed=9), dims=[3]) * 0.8
    with torch.no_grad():
        assert rel_l2(idloss.extract_feats(x, sd), g['feats_x']) < TOL
        assert abs(float(idloss.similarity(x, y, sd)) - float(g['sim'])) < 1e-5


def test_sgw_plus_projector_matches_reference(golden):
    """Two steps of the oracle's `sgw+` projector against the latent the reference's w_plus_projector.py produced for the same
    draws (oracle/make_golden_sgw.py)."""
    from oracle.make_golden import make_nets
    g = golden('sgw_plus')
    sd = weights.generator_state_dict(0)
    p = loops.Projector(sd, weights.target_image(), weights.canonical_camera(0.3), make_nets(), kind='sgw+', num_steps=500,
                        noise=loops.NoiseSource(300))
    infos = [p.step(i) for i in range(2)]
    assert rel_l2(p.result(), g['w']) < 1e-6
    assert abs(infos[1]['loss'] / float(g['loss'][1]) - 1) < 1e-4 and abs(p.w_std / float(g['w_std']) - 1) < 1e-6
