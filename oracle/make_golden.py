"""Mint golden vectors from the REFERENCE ITSELF and validate the oracle against it.

TEST INFRASTRUCTURE.  Runs only in the build container (needs `/root/reference`, read-only):

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden [--fast]

For each function on the hot path it (1) runs the reference's own code (imported through
`oracle/ref_shim.py`) on seeded inputs, (2) runs the oracle restatement on the same inputs and records the
discrepancy in `tests/golden/REPORT.json`, (3) stores the reference outputs (or strided subsets + float64
checksums for the big ones) under `tests/golden/*.npz`.  The fixtures travel to the GPU box; the
reference does not.
"""
import argparse
import contextlib
import json
import math
import os
import sys
import time

import numpy as np
import torch

from . import criteria, generator, geometry, loops, ops, ref_shim, weights

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')
REPORT = {}


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def note(name, ref, mine, exact=False):
    if exact:
        ok = bool(torch.equal(ref, mine))
        REPORT[name] = {'exact': ok}
    else:
        REPORT[name] = {'rel_l2': rel_l2(mine, ref)}
    print(f'  {name:40s} {REPORT[name]}', flush=True)


def npy(t):
    return t.detach().cpu().numpy()


@contextlib.contextmanager
def injected_rng(rand_like=(), rand=(), randn_like=()):
    """Feed the reference's global-RNG draws (renderer.py:190,237; projector w-noise) from queues."""
    q1, q2, q3 = list(rand_like), list(rand), list(randn_like)
    o1, o2, o3 = torch.rand_like, torch.rand, torch.randn_like

    def f1(x, *a, **k):
        return q1.pop(0).to(x.dtype) if q1 else o1(x, *a, **k)

    def f2(*a, **k):
        return q2.pop(0) if q2 else o2(*a, **k)

    def f3(x, *a, **k):
        return q3.pop(0).to(x.dtype) if q3 else o3(x, *a, **k)
    torch.rand_like, torch.rand, torch.randn_like = f1, f2, f3
    try:
        yield
    finally:
        torch.rand_like, torch.rand, torch.randn_like = o1, o2, o3


# ----------------------------------------------------------------------------- L1 ops

def golden_ops():
    print('[ops]')
    from torch_utils.ops import bias_act as r_ba, upfirdn2d as r_up, filtered_lrelu as r_fl
    rs = np.random.RandomState(11)
    g = {}
    x = torch.from_numpy(rs.standard_normal((2, 6, 9, 7)).astype(np.float32)) * 2
    b = torch.from_numpy(rs.standard_normal(6).astype(np.float32))
    g['ba_x'], g['ba_b'] = npy(x), npy(b)
    for act in ops.ACTS:
        for tag, kw in (('d', {}), ('g', dict(gain=0.7, clamp=0.9, alpha=0.3))):
            ref = r_ba._bias_act_ref(x, b, act=act, **kw)
            mine = ops.bias_act(x, b, act=act, **kw)
            note(f'bias_act/{act}/{tag}', ref, mine)
            g[f'ba_{act}_{tag}'] = npy(ref)
    # gradients of the hot activations (bias_act.py:156-171 semantics: dx from dy and y)
    for act, kw in (('lrelu', dict(gain=math.sqrt(2), clamp=256.)), ('linear', dict(clamp=1.0)), ('lrelu', dict(clamp=0.5))):
        xr = x.clone().requires_grad_(True)
        br = b.clone().requires_grad_(True)
        y = r_ba._bias_act_ref(xr, br, act=act, **kw)
        dy = torch.from_numpy(np.random.RandomState(12).standard_normal(y.shape).astype(np.float32))
        y.backward(dy)
        key = f"ba_grad_{act}_{kw.get('clamp')}"
        g[key + '_dy'], g[key + '_dx'], g[key + '_db'] = npy(dy), npy(xr.grad), npy(br.grad)
    # upfirdn2d: the three hit specialisations (+ generic cases)
    f = r_up.setup_filter([1, 3, 3, 1])
    note('setup_filter', f, ops.setup_filter([1, 3, 3, 1]))
    g['f4'] = npy(f)
    xu = torch.from_numpy(rs.standard_normal((2, 3, 17, 17)).astype(np.float32))
    g['up_x'] = npy(xu)
    cases = {
        'blur_after_convT': dict(up=1, down=1, padding=[1, 1, 1, 1], gain=4.0, flip_filter=False),
        'upsample2d': dict(up=2, down=1, padding=[2, 1, 2, 1], gain=4.0, flip_filter=False),
        'upsample2d_bwd': dict(up=1, down=2, padding=[1, 2, 1, 2], gain=4.0, flip_filter=True),
        'generic_a': dict(up=[2, 1], down=[1, 2], padding=[3, 0, 1, 2], gain=1.3, flip_filter=True),
        'crop': dict(up=1, down=1, padding=[-1, 2, 0, -2], gain=1.0, flip_filter=False),
        'down3': dict(up=1, down=3, padding=[2, 2, 2, 2], gain=1.0, flip_filter=False),
    }
    for name, kw in cases.items():
        ref = r_up._upfirdn2d_ref(xu, f, **kw)
        mine = ops.upfirdn2d(xu, f, **kw)
        note(f'upfirdn2d/{name}', ref, mine)
        g['up_' + name] = npy(ref)
    REPORT['upfirdn2d_cases'] = {k: {kk: (list(vv) if isinstance(vv, list) else vv) for kk, vv in v.items()} for k, v in cases.items()}
    f12 = r_up.setup_filter(rs.standard_normal(12).astype(np.float32).tolist(), normalize=False)  # separable (>=8 taps)
    assert f12.ndim == 1
    g['f12'] = npy(f12)
    xs = torch.from_numpy(rs.standard_normal((1, 2, 20, 22)).astype(np.float32))
    g['up_xs'] = npy(xs)
    ref = r_up._upfirdn2d_ref(xs, f12, up=2, down=1, padding=[5, 6, 5, 6], gain=4.0)
    note('upfirdn2d/separable12', ref, ops.upfirdn2d(xs, f12, up=2, padding=[5, 6, 5, 6], gain=4.0))
    g['up_separable12'] = npy(ref)
    ref = r_up.upsample2d(xu, f, impl='ref')
    note('upsample2d', ref, ops.upsample2d(xu, f))
    # filtered_lrelu (SG3 layer configs: up 2 / down 2 with 12-tap separable filters, and up 4/down 2)
    fl_cases = {
        'u2d2': dict(up=2, down=2, padding=[9, 10, 9, 10], gain=math.sqrt(2), slope=0.2, clamp=256., flip_filter=False),
        'u2d1': dict(up=2, down=1, padding=[5, 6, 5, 6], gain=math.sqrt(2), slope=0.2, clamp=None, flip_filter=False),
        'u1d2': dict(up=1, down=2, padding=[5, 5, 5, 5], gain=1.1, slope=0.1, clamp=0.8, flip_filter=True),
        'u1d1': dict(up=1, down=1, padding=[0, 0, 0, 0], gain=math.sqrt(2), slope=0.2, clamp=None, flip_filter=False),
    }
    fd12 = r_up.setup_filter(rs.standard_normal(12).astype(np.float32).tolist(), normalize=False)
    g['fd12'] = npy(fd12)
    bs = torch.from_numpy(rs.standard_normal(2).astype(np.float32))
    g['fl_b'] = npy(bs)
    for name, kw in fl_cases.items():
        fu = f12 if kw['up'] > 1 else (None if name == 'u1d1' else f12)
        fd = fd12 if kw['down'] > 1 else None
        ref = r_fl._filtered_lrelu_ref(xs, fu=fu, fd=fd, b=bs, **kw)
        mine = ops.filtered_lrelu(xs, fu=fu, fd=fd, b=bs, **kw)
        note(f'filtered_lrelu/{name}', ref, mine)
        g['fl_' + name] = npy(ref)
    np.savez_compressed(os.path.join(OUT, 'ops.npz'), **g)


# ----------------------------------------------------------------------------- renderer pieces

def golden_render(G, sd):
    print('[renderer pieces]')
    from training.volumetric_rendering.renderer import ImportanceRenderer
    from training.volumetric_rendering.ray_marcher import MipRayMarcher2
    from training.volumetric_rendering.ray_sampler import RaySampler
    rs = np.random.RandomState(21)
    g = {}
    c = torch.cat([weights.canonical_camera(0.3), weights.canonical_camera(-0.45, 0.1)], 0)
    ro, rd = RaySampler()(c[:, :16].view(-1, 4, 4), c[:, 16:].view(-1, 3, 3), 128)
    mo, md = generator.ray_sampler(c[:, :16].reshape(-1, 4, 4), c[:, 16:].reshape(-1, 3, 3), 128)
    note('ray_sampler/origins', ro, mo)
    note('ray_sampler/dirs', rd, md)
    g['cam'] = npy(c)
    g['ray_dirs_sub'] = npy(rd[:, ::37])
    g['ray_origins_sub'] = npy(ro[:, ::37])
    # small random planes, 2 images x 192 rays (rows of the 128^2 grid around the centre)
    n, r = 2, 192
    planes = torch.from_numpy(rs.standard_normal((n, 3, 32, 64, 64)).astype(np.float32))
    sel = torch.arange(r) * 61 + 3000
    o, d = ro[:, sel].contiguous(), rd[:, sel].contiguous()
    g['planes'], g['sel'] = npy(planes), npy(sel)
    for (dc, df) in ((48, 48), (32, 32), (12, 20)):
        rk = dict(generator.RENDERING_DEFAULTS, depth_resolution=dc, depth_resolution_importance=df)
        jit = torch.from_numpy(rs.uniform(size=(n, r, dc, 1)).astype(np.float32))
        u = torch.from_numpy(rs.uniform(size=(n * r, df)).astype(np.float32))
        rend = ImportanceRenderer()
        cap = {}
        orig_pdf = rend.sample_pdf

        def spy(bins, w, N, det=False, eps=1e-5):
            cap['bins'], cap['w'] = bins.clone(), w.clone()
            out = orig_pdf(bins, w, N, det=det, eps=eps)
            cap['fine'] = out.clone()
            return out
        rend.sample_pdf = spy
        with injected_rng(rand_like=[jit], rand=[u]):
            rgb, depth, wsum = rend(planes, G.decoder, o, d, rk)
        m_rgb, m_depth, m_wsum, aux = generator.importance_render(sd, planes, o, d, rk, jit, u, return_aux=True)
        tag = f'{dc}_{df}'
        note(f'render/{tag}/rgb', rgb, m_rgb)
        note(f'render/{tag}/depth', depth, m_depth)
        note(f'render/{tag}/wsum', wsum, m_wsum)
        note(f'render/{tag}/depths_fine', cap['fine'].reshape(n, r, df, 1), aux['depths_fine'])
        note(f'render/{tag}/bins', cap['bins'], aux['bins'], exact=True)
        for k_, v_ in (('jit', jit), ('u', u), ('rgb', rgb), ('depth', depth), ('wsum', wsum), ('fine', cap['fine']),
                       ('bins', cap['bins']), ('pdf_w', cap['w'])):
            g[f'r{tag}_{k_}'] = npy(v_)
        # index-exact pieces, from the reference's own float inputs
        s_fine, inds = generator.inverse_cdf(aux['bins'], aux['cdf'], u)
        g[f'r{tag}_cdf'], g[f'r{tag}_inds'] = npy(aux['cdf']), npy(inds)
        g[f'r{tag}_perm'] = npy(aux['perm'])
        g[f'r{tag}_depths_all'] = npy(aux['depths_all'])
    # stand-alone marcher / unify on random tensors
    col = torch.from_numpy(rs.uniform(size=(2, 50, 24, 32)).astype(np.float32))
    sig = torch.from_numpy(rs.standard_normal((2, 50, 24, 1)).astype(np.float32)) * 3
    dep, _ = torch.sort(torch.from_numpy(rs.uniform(2.25, 3.3, size=(2, 50, 24, 1)).astype(np.float32)), dim=2)
    rk = generator.RENDERING_DEFAULTS
    r1, r2, r3 = MipRayMarcher2()(col, sig, dep, rk)
    m1, m2, m3 = generator.ray_march(col, sig, dep, rk)
    note('ray_march/rgb', r1, m1)
    note('ray_march/depth', r2, m2)
    note('ray_march/weights', r3, m3)
    g.update(rm_col=npy(col), rm_sig=npy(sig), rm_dep=npy(dep), rm_rgb=npy(r1), rm_depth=npy(r2), rm_w=npy(r3))
    np.savez_compressed(os.path.join(OUT, 'render.npz'), **g)


# ----------------------------------------------------------------------------- full generator

def golden_synthesis(G, sd):
    print('[synthesis]')
    g = {}
    z = torch.from_numpy(np.random.RandomState(31).standard_normal((3, 512)).astype(np.float32))
    c3 = torch.cat([weights.canonical_camera(0.3), weights.canonical_camera(0.0), weights.canonical_camera(-0.5, 0.2)], 0)
    w_ref = G.mapping(z, c3)
    w_mine = generator.mapping(sd, z, c3, G.rendering_kwargs)
    note('mapping', w_ref, w_mine)
    g['z'], g['c3'], g['w'] = npy(z), npy(c3), npy(w_ref[:, 0])
    ws = weights.w_pivot(5)
    c = weights.canonical_camera(0.3)
    rk = {**generator.RENDERING_DEFAULTS, **G.rendering_kwargs}
    jit, u = generator.make_render_noise(1, 128 * 128, rk, seed=7)
    t = time.time()
    with injected_rng(rand_like=[jit], rand=[u]):
        ref = G.synthesis(ws, c, noise_mode='const', cache_backbone=True)
    print('  reference synthesis %.1fs' % (time.time() - t))
    planes_ref = G._last_planes
    t = time.time()
    mine = generator.synthesis(sd, ws, c, rk, jitter=jit, u=u)
    print('  oracle synthesis %.1fs' % (time.time() - t))
    note('synthesis/planes', planes_ref, mine['planes'])
    for k in ('image', 'image_raw', 'image_depth'):
        note(f'synthesis/{k}', ref[k], mine[k])
    g['ws'], g['c'] = npy(ws), npy(c)
    g['image_raw'], g['image_depth'] = npy(ref['image_raw']), npy(ref['image_depth'])
    g['image_sub'] = npy(ref['image'][:, :, 1::4, 2::4])
    g['planes_sub'] = npy(planes_ref[:, ::7, 3::8, 5::8])
    g['image_sum'] = np.float64(ref['image'].double().sum().item())
    g['image_sqsum'] = np.float64(ref['image'].double().square().sum().item())
    g['planes_sqsum'] = np.float64(planes_ref.double().square().sum().item())
    # sample_mixed (triplane.py:98-102)
    pts = torch.from_numpy(np.random.RandomState(32).uniform(-0.6, 0.6, size=(1, 500, 3)).astype(np.float32))
    sm = G.sample_mixed(pts, torch.zeros_like(pts), ws, noise_mode='const')
    mm = generator.sample_mixed(sd, pts, ws, rk)
    note('sample_mixed/sigma', sm['sigma'], mm['sigma'])
    note('sample_mixed/rgb', sm['rgb'], mm['rgb'])
    g['sm_pts'], g['sm_sigma'], g['sm_rgb'] = npy(pts), npy(sm['sigma']), npy(sm['rgb'])
    np.savez_compressed(os.path.join(OUT, 'synthesis.npz'), **g)
    return ref, jit, u


# ----------------------------------------------------------------------------- geometry + losses

def build_ref_losses(nets):
    import spi.criteria.lpips.lpips as r_lpips_mod
    import spi.criteria.lpips.utils as r_lpips_utils
    from spi.criteria.bbox_cx_loss import BoxCXLoss
    lin_sd = {f'{i}.1.weight': w for i, w in enumerate(nets['lin'])}
    r_lpips_mod.get_state_dict = lambda *a, **k: lin_sd
    L = r_lpips_mod.LPIPS(net_type='vgg').eval()
    L.net.layers.load_state_dict(nets['vgg16'], strict=False)
    B = BoxCXLoss().eval()
    B.vgg_model.slice1.load_state_dict(nets['vgg19'])
    return L, B


def make_nets():
    return {'vgg16': weights.vgg_state_dict(weights.VGG16_CFG, seed=1), 'lin': weights.lpips_lin_weights(1),
            'vgg19': weights.vgg_state_dict(weights.VGG19_HEAD, seed=2)}


def golden_geometry_losses(ref_out, nets):
    print('[geometry + losses]')
    from spi.utils import camera_utils as r_cam
    from spi.utils.rotate import rotate as r_rotate
    from spi.utils.mask_utils import calculate_face_mask
    g = {}
    c = weights.canonical_camera(0.3)
    note('canonical_camera', r_cam.cal_canonical_c(0.3, 0, 1, 'cpu'), c)
    note('mirror_camera', r_cam.cal_mirror_c(c), geometry.mirror_camera(c))
    import spi.utils.camera_utils as cu
    cu.GAUSS_CONST = torch.sqrt(torch.tensor(2 * torch.pi))
    note('camera_weight', r_cam.cal_camera_weight(c), geometry.camera_weight(c))
    g['cam_weight'] = npy(r_cam.cal_camera_weight(c))
    r1 = torch.from_numpy(np.random.RandomState(41).uniform(size=(4, 2)).astype(np.float32))
    with injected_rng(rand=[r1[:, 0:1].clone(), r1[:, 1:2].clone()]):
        sc = r_cam.sample_surrounding_camera(c, batch_size=4, yaw_range=0.2, pitch_range=0.1)
    note('sample_surrounding_camera', sc, geometry.sample_surrounding_camera(c, r1, 0.2, 0.1))
    with injected_rng(rand=[r1[:, 0:1].clone(), r1[:, 1:2].clone()]):
        sc2 = r_cam.sample_camera(batch_size=4, yaw_range=0.7, pitch_range=0.4, device='cpu')
    note('sample_camera', sc2, geometry.sample_camera(r1, 0.7, 0.4))
    g['rand42'], g['surround'], g['sampled'] = npy(r1), npy(sc), npy(sc2)
    parsing = weights.parsing_mask()
    fm_ref = calculate_face_mask(parsing)
    note('face_mask', fm_ref.float(), geometry.face_mask(parsing).float(), exact=True)
    g['face_mask_sum'] = np.int64(fm_ref.sum().item())
    # rotate: smooth synthetic depths (planar + bump), 4 views
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, 128), torch.linspace(-1, 1, 128), indexing='ij')
    base = 2.7 - 0.35 * torch.exp(-(xx ** 2 + yy ** 2) * 2.5)
    tdepth = torch.stack([base + 0.01 * k * xx for k in range(4)])[:, None]
    sdepth = base[None, None].repeat(4, 1, 1, 1)
    img = weights.target_image().repeat(4, 1, 1, 1)
    fm = geometry.face_mask(parsing).float().repeat(4, 1, 1, 1)
    wr, wm = r_rotate(sc, tdepth, img, c.repeat(4, 1), sdepth, src_mask=fm, EPS=5e-2)
    mr, mm = geometry.rotate(sc, tdepth, img, c.repeat(4, 1), sdepth, fm, eps=5e-2)
    note('rotate/rgb', wr, mr)
    note('rotate/mask', wm, mm)
    g['rot_tdepth'] = npy(tdepth)
    g['rot_rgb_sub'], g['rot_mask_sub'] = npy(wr[:, :, 2::8, 3::8]), npy(wm[:, :, 2::8, 3::8])
    g['rot_rgb_sqsum'] = np.float64(wr.double().square().sum().item())
    g['rot_mask_sum'] = np.float64(wm.double().sum().item())
    # losses
    L, B = build_ref_losses(nets)
    x = ref_out['image']
    y = weights.target_image()
    lp_ref = L(x, y)
    note('lpips', lp_ref.reshape(()), criteria.lpips(x, y, nets['vgg16'], nets['lin']))
    g['lpips'] = npy(lp_ref)
    lm = weights.landmarks68().repeat(2, 1, 1)
    xb = torch.cat([x, torch.flip(x, dims=[3])], 0)
    yb = torch.cat([y, 0.5 * y + 0.1], 0)
    cx_ref = B(xb, yb, lm)
    note('box_cx', cx_ref.reshape(()), criteria.box_cx(xb, yb, lm, nets['vgg19']))
    g['box_cx'] = npy(cx_ref)
    from spi.criteria.bbox_cx_loss import get_landmark_bbox
    for k_, (a, b_) in enumerate(zip(get_landmark_bbox(lm), criteria.landmark_boxes(lm))):
        note(f'landmark_boxes/{k_}', a, b_, exact=True)
        g[f'boxes{k_}'] = npy(a)
    np.savez_compressed(os.path.join(OUT, 'geometry_losses.npz'), **g)


# ----------------------------------------------------------------------------- loop steps

def golden_steps(G, sd, nets, fast):
    print('[loop steps]')
    import copy as _copy
    from spi.configs import global_config, hyperparameters
    from spi.training.projectors import mirror_projector
    import spi.utils.camera_utils as cu
    cu.GAUSS_CONST = torch.sqrt(torch.tensor(2 * torch.pi))
    global_config.device = 'cpu'
    mirror_projector.log_image = lambda *a, **k: None
    L, B = build_ref_losses(nets)
    g = {}
    target, c = weights.target_image(), weights.canonical_camera(0.3)
    parsing, lm = weights.parsing_mask(), weights.landmarks68()
    rk = {**generator.RENDERING_DEFAULTS, **G.rendering_kwargs}
    # ---- stage 1, 'mir', 2 steps of num_steps=500 (mirror_projector.py:81-131)
    src = loops.NoiseSource(100)
    mine = loops.Projector(sd, target, c, nets, kind='mir', num_steps=500, rk=rk, noise=src)
    # replay the same draws into the reference: noise-buffer init randn_like x13, then per step randn_like(w), rand_like, rand
    rep = loops.NoiseSource(100)
    bufs = [rep.randn(*sd[k].shape) for k in loops.noise_buffer_names(sd)]
    nsteps = 2
    per_step = []
    for _ in range(nsteps):
        wn = rep.randn(1, 14, 512)
        jit, u = rep.render(2, 128 * 128, rk)
        per_step.append((wn, jit, u))
    fg = 1 - (parsing == 0).float()
    Gs = _copy.deepcopy(G)
    import tqdm as _tqdm
    mirror_projector.tqdm = lambda it, *a, **k: it
    # run the reference for exactly `nsteps` steps by truncating its range()
    real_range = range
    mirror_projector.range = lambda n: real_range(nsteps)
    # step / num_steps must still use num_steps=500 -> pass num_steps=500
    with injected_rng(randn_like=bufs + [s[0] for s in per_step], rand_like=[s[1] for s in per_step], rand=[s[2] for s in per_step]):
        w_ref = mirror_projector.project(Gs, target, c, lpips_func=L, fg_mask=fg, device=torch.device('cpu'),
                                         w_avg_samples=600, num_steps=500, w_name='g')
    infos = [mine.step(i) for i in range(nsteps)]
    note('mir/w_opt', w_ref.detach(), mine.result())
    g['mir_w'] = npy(w_ref.detach())
    g['mir_loss'] = np.array([i['loss'] for i in infos])
    g['mir_dist'] = np.array([i['dist'] for i in infos])
    g['mir_w_std'] = np.float64(mine.w_std)
    REPORT['mir/oracle_losses'] = [i['loss'] for i in infos]
    # ---- stage 2 (coach steps restated from rot_bbox_cx_coach.py:68-151 using reference modules)
    from spi.utils.rotate import rotate as r_rotate
    from spi.utils import camera_utils as r_cam
    from spi.utils.mask_utils import calculate_face_mask
    from criteria.l2_loss import l2_loss
    kinds = ['pti'] if fast else ['pti', 'RotBbox']
    for kind in kinds:
        Gt = _copy.deepcopy(G).requires_grad_(True)
        Go = _copy.deepcopy(G)
        opt = torch.optim.Adam(Gt.parameters(), lr=3e-4)
        wp = weights.w_pivot(5).requires_grad_(True)
        src = loops.NoiseSource(200)
        rep = loops.NoiseSource(200)
        coach = loops.Coach(sd, weights.w_pivot(5), target, c, parsing, lm, nets, kind=kind, rk=rk, noise=src)
        info = coach.step(0)
        # reference-side replay
        face_mask = calculate_face_mask(parsing).float()
        image_m, face_mask_m, camera_m = torch.flip(target, dims=[3]), torch.flip(face_mask, dims=[3]), r_cam.cal_mirror_c(c)
        opt.zero_grad()
        jit, u = rep.render(1, 128 * 128, rk)
        with injected_rng(rand_like=[jit], rand=[u]):
            out = Gt.synthesis(wp, c, noise_mode='const')
        l2v = l2_loss(out['image'], target)
        lp = torch.squeeze(L(out['image'], target))
        (l2v + lp).backward()
        depth = out['image_depth']
        ref_info = {'l2': float(l2v), 'lpips': float(lp)}
        if kind == 'RotBbox':
            r4 = rep.rand(4, 2)
            with injected_rng(rand=[r4[:, 0:1].clone(), r4[:, 1:2].clone()]):
                cams = r_cam.sample_surrounding_camera(c, batch_size=4, yaw_range=0.2, pitch_range=0.1)
            jit, u = rep.render(4, 128 * 128, rk)
            with injected_rng(rand_like=[jit], rand=[u]):
                gen = Gt.synthesis(wp.repeat(4, 1, 1), cams, noise_mode='const')
            with torch.no_grad():
                warp, wmask = r_rotate(cams, gen['image_depth'], target.repeat(4, 1, 1, 1), c.repeat(4, 1),
                                       depth.repeat(4, 1, 1, 1), src_mask=face_mask.repeat(4, 1, 1, 1), EPS=5e-2)
            lrot = L(gen['image'] * wmask, warp) * 0.1 * 4
            lrot.backward()
            ref_info['rot'] = float(lrot)
            r4 = rep.rand(4, 2)
            with injected_rng(rand=[r4[:, 0:1].clone(), r4[:, 1:2].clone()]):
                cams = r_cam.sample_surrounding_camera(camera_m, batch_size=4, yaw_range=0.2, pitch_range=0.1)
            jit, u = rep.render(4, 128 * 128, rk)
            with injected_rng(rand_like=[jit], rand=[u]):
                gen = Gt.synthesis(wp.repeat(4, 1, 1), cams, noise_mode='const')
            with torch.no_grad():
                warp, wmask = r_rotate(cams, gen['image_depth'], image_m.repeat(4, 1, 1, 1), camera_m.repeat(4, 1),
                                       torch.flip(depth, dims=[3]).repeat(4, 1, 1, 1),
                                       src_mask=face_mask_m.repeat(4, 1, 1, 1), EPS=5e-2)
                warp, wmask = torch.flip(warp, dims=[3]), torch.flip(wmask, dims=[3])
            lmir = B(torch.flip(gen['image'], dims=[3]) * wmask, warp, lm.repeat(4, 1, 1)) * 0.05 * 4
            lmir.backward()
            ref_info['mirror'] = float(lmir)
            r4 = rep.rand(4, 2)
            with injected_rng(rand=[r4[:, 0:1].clone(), r4[:, 1:2].clone()]):
                cams = r_cam.sample_camera(batch_size=4, yaw_range=0.7, pitch_range=0.4, device='cpu')
            jit, u = rep.render(4, 128 * 128, rk)
            with injected_rng(rand_like=[jit], rand=[u]):
                d_new = Gt.synthesis(wp.repeat(4, 1, 1), cams, noise_mode='const')['image_depth']
            jit, u = rep.render(4, 128 * 128, rk)
            with torch.no_grad(), injected_rng(rand_like=[jit], rand=[u]):
                d_ref = Go.synthesis(wp.repeat(4, 1, 1), cams, noise_mode='const')['image_depth']
            ld = l2_loss(d_ref, d_new) * 1.0
            ld.backward()
            ref_info['depth'] = float(ld)
        grads = {k: p.grad.clone() for k, p in Gt.named_parameters() if p.grad is not None}
        opt.step()
        new_sd = Gt.state_dict()
        REPORT[f'{kind}/ref_info'] = ref_info
        REPORT[f'{kind}/oracle_info'] = {k: v for k, v in info.items() if k != 'early_exit'}
        for k in ('decoder.net.0.weight', 'decoder.net.2.bias', 'superresolution.block1.conv1.weight',
                  'backbone.synthesis.b4.const', 'backbone.synthesis.b64.conv0.affine.weight',
                  'backbone.synthesis.b256.torgb.weight', 'backbone.synthesis.b128.conv1.noise_strength'):
            note(f'{kind}/grad/{k}', grads[k], coach.sd[k].grad)
            note(f'{kind}/param/{k}', new_sd[k], coach.sd[k].detach())
            sub = grads[k].reshape(-1)[::max(1, grads[k].numel() // 4096)]
            g[f'{kind}_grad_{k}'] = npy(sub)
            g[f'{kind}_gradnorm_{k}'] = np.float64(grads[k].double().norm().item())
        g[f'{kind}_wgrad'] = npy(wp.grad)
        note(f'{kind}/grad/ws', wp.grad, coach.w.grad)
        for k_, v_ in ref_info.items():
            g[f'{kind}_{k_}'] = np.float64(v_)
    np.savez_compressed(os.path.join(OUT, 'steps.npz'), **g)


def golden_idloss():
    print('[id loss]')
    from spi.criteria.id_loss.model_irse import Backbone
    from . import idloss
    sd = weights.irse50_state_dict(3)
    net = Backbone(input_size=112, num_layers=50, drop_ratio=0.6, mode='ir_se').eval()
    print('  load:', net.load_state_dict({k[len('facenet.'):]: v for k, v in sd.items()}, strict=True))
    x = weights.target_image()
    y = torch.flip(weights.target_image(seed=9), dims=[3]) * 0.8
    pool = torch.nn.AdaptiveAvgPool2d((112, 112))

    def ref_feats(t):      # IDLoss.extract_feats (id_loss.py:17-21)
        return net(pool(t[:, :, 35:223, 32:220]))
    with torch.no_grad():
        fx, fy = ref_feats(x), ref_feats(y)
        sim = fx[0].dot(fy[0])
        note('id/feats', fx, idloss.extract_feats(x, sd))
        note('id/similarity', sim.reshape(()), idloss.similarity(x, y, sd).reshape(()))
    np.savez_compressed(os.path.join(OUT, 'idloss.npz'), feats_x=npy(fx), feats_y=npy(fy), sim=npy(sim))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--fast', action='store_true', help='skip the ~6 min RotBbox reference step')
    ap.add_argument('--only', default=None)
    args = ap.parse_args()
    assert ref_shim.available(), 'reference tree not found'
    ref_shim.install()
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    sd = weights.generator_state_dict(0)
    G = ref_shim.build_reference_generator()
    missing = G.load_state_dict(sd, strict=True)
    print('state dict loaded into the reference module:', missing)
    nets = make_nets()
    only = args.only.split(',') if args.only else None

    def want(k):
        return only is None or k in only
    if want('ops'):
        golden_ops()
    if want('render'):
        golden_render(G, sd)
    if want('synthesis') or want('losses'):
        ref_out, _, _ = golden_synthesis(G, sd)
    if want('losses'):
        golden_geometry_losses(ref_out, nets)
    if want('steps'):
        golden_steps(G, sd, nets, args.fast)
    if want('idloss'):
        golden_idloss()
    path = os.path.join(OUT, 'REPORT.json')
    old = {}
    if only is not None and os.path.exists(path):
        old = json.load(open(path))
    old.update(REPORT)
    old['_meta'] = {'torch': torch.__version__, 'threads': torch.get_num_threads(), 'reference': ref_shim.REF,
                    'note': 'rel_l2 = ||oracle - reference|| / ||reference|| on identical inputs'}
    json.dump(old, open(path, 'w'), indent=1, sort_keys=True)
    print('wrote', path)


if __name__ == '__main__':
    main()
