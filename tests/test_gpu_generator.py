"""GPU parity: full TriPlaneGenerator.synthesis against the reference goldens, gradients against the CPU oracle,
depth-guided warp and Adam."""
import numpy as np
import pytest
import torch

from conftest import rel_l2
from oracle import generator as OG
from oracle import geometry as OGeo
from oracle import weights

pytestmark = pytest.mark.gpu
RENDER_TOL = 1e-3      # north_star: renders within 1e-3 rel-L2 of the reference (TF32 tensor-core contraction)


def T(a):
    return torch.from_numpy(np.asarray(a))


def test_state_dict_contract(product_G, gen_sd):
    names = dict(product_G.named_parameters())
    bufs = dict(product_G.named_buffers())
    assert len(names) == 132 and len(bufs) == 44
    assert set(names) | set(bufs) == set(gen_sd)
    assert product_G.backbone.mapping.num_ws == 14
    assert len([k for k, _ in product_G.backbone.synthesis.named_buffers() if 'noise_const' in k]) == 13


def test_mapping_golden(product_G, golden):
    g = golden('synthesis')
    w = product_G.mapping(T(g['z']).cuda(), T(g['c3']).cuda())
    assert w.shape == (3, 14, 512) and rel_l2(w[:, 0], g['w']) < 1e-4


def test_synthesis_golden(product_G, golden):
    g = golden('synthesis')
    rk = OG.RENDERING_DEFAULTS
    jit, u = OG.make_render_noise(1, 128 * 128, rk, seed=7)
    product_G.renderer.inject_noise(jit.cuda(), u.cuda())
    out = product_G.synthesis(T(g['ws']).cuda(), T(g['c']).cuda(), noise_mode='const', cache_backbone=True)
    assert out['image'].shape == (1, 3, 512, 512) and out['image_raw'].shape == (1, 3, 128, 128) and out['image_depth'].shape == (1, 1, 128, 128)
    assert rel_l2(product_G._last_planes[:, ::7, 3::8, 5::8], g['planes_sub']) < RENDER_TOL
    assert rel_l2(out['image_raw'], g['image_raw']) < RENDER_TOL
    assert rel_l2(out['image_depth'], g['image_depth']) < RENDER_TOL
    assert rel_l2(out['image'][:, :, 1::4, 2::4], g['image_sub']) < RENDER_TOL
    assert abs(out['image'].double().square().sum().item() / float(g['image_sqsum']) - 1) < 2e-3


def test_synthesis_golden_at_the_bench_depth_resolution(product_G, golden):
    """Same check at 32 coarse + 32 importance samples -- the configuration bench.py times (the tcgen05 two-round renderer path) -- against
    the reference's own synthesis at that setting (oracle/make_golden_32.py)."""
    import copy
    g = golden('synthesis_32')
    G = copy.deepcopy(product_G)
    G.rendering_kwargs = dict(G.rendering_kwargs, depth_resolution=32, depth_resolution_importance=32)
    rk = dict(OG.RENDERING_DEFAULTS, depth_resolution=32, depth_resolution_importance=32)
    jit, u = OG.make_render_noise(1, 128 * 128, rk, seed=7)
    G.renderer.inject_noise(jit.cuda(), u.cuda())
    out = G.synthesis(T(g['ws']).cuda(), T(g['c']).cuda(), noise_mode='const')
    errs = (rel_l2(out['image_raw'], g['image_raw']), rel_l2(out['image_depth'], g['image_depth']), rel_l2(out['image'][:, :, 1::4, 2::4], g['image_sub']))
    print('32+32 synthesis rel-L2 (image_raw, image_depth, image):', errs)
    assert max(errs) < RENDER_TOL
    assert abs(out['image'].double().square().sum().item() / float(g['image_sqsum']) - 1) < 2e-3


def test_sample_mixed_golden(product_G, golden):
    g = golden('synthesis')
    pts = T(g['sm_pts']).cuda()
    out = product_G.sample_mixed(pts, torch.zeros_like(pts), T(g['ws']).cuda(), noise_mode='const')
    assert rel_l2(out['sigma'], g['sm_sigma']) < RENDER_TOL and rel_l2(out['rgb'], g['sm_rgb']) < RENDER_TOL


def test_synthesis_gradients_golden(product_G, golden, gen_sd):
    """One PTI loss backward (L2 only here) against the reference's recorded gradients is covered in test_gpu_loop;
    this checks d(sum image)/d(ws) and a few parameter gradients against the CPU oracle at reduced depth samples."""
    import copy
    G = copy.deepcopy(product_G).requires_grad_(True)
    rk = dict(G.rendering_kwargs)
    ws = weights.w_pivot(5)
    c = weights.canonical_camera(0.3)
    jit, u = OG.make_render_noise(1, 128 * 128, {**OG.RENDERING_DEFAULTS, **rk}, seed=3)
    gimg = torch.randn(1, 3, 512, 512, generator=torch.Generator().manual_seed(1))
    gdep = torch.randn(1, 1, 128, 128, generator=torch.Generator().manual_seed(2))
    keys = ['decoder.net.0.weight', 'backbone.synthesis.b4.const', 'backbone.synthesis.b64.conv0.weight',
            'backbone.synthesis.b256.torgb.weight', 'superresolution.block1.conv1.weight', 'superresolution.block0.conv0.affine.weight',
            'backbone.synthesis.b128.conv1.noise_strength', 'backbone.synthesis.b32.conv1.bias']
    sd = {k: v.clone().requires_grad_(k in keys) for k, v in gen_sd.items()}
    wo = ws.clone().requires_grad_(True)
    out = OG.synthesis(sd, wo, c, jitter=jit, u=u)
    ((out['image'] * gimg).sum() + (out['image_depth'] * gdep).sum()).backward()
    wg = ws.cuda().requires_grad_(True)
    G.renderer.inject_noise(jit.cuda(), u.cuda())
    og = G.synthesis(wg, c.cuda(), noise_mode='const')
    ((og['image'] * gimg.cuda()).sum() + (og['image_depth'] * gdep.cuda()).sum()).backward()
    assert rel_l2(og['image'], out['image']) < RENDER_TOL
    assert rel_l2(wg.grad, wo.grad) < 3e-2      # TF32 contraction in fwd and bwd
    params = dict(G.named_parameters())
    errs = {k: rel_l2(params[k].grad, sd[k].grad) for k in keys}
    print('grad rel-L2:', rel_l2(wg.grad, wo.grad), errs)
    # d/d(noise_strength) is ONE scalar = a sum of 4 M signed terms: measured on B200 it moves by +-2 % between two runs of the same code
    # (fp32 reduce-add order of the split layers) and by 1-9 % between engines / tile shapes, i.e. it sits at the TF32 noise floor of
    # this network; every tensor-valued gradient is held to 3e-2
    tol = {k: (0.25 if k.endswith('noise_strength') else 3e-2) for k in keys}
    assert all(errs[k] < tol[k] for k in keys), errs


def test_rotate_golden(golden):
    from spi_b200.utils.rotate import rotate
    g = golden('geometry_losses')
    c = weights.canonical_camera(0.3)
    sc = T(g['surround'])
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, 128), torch.linspace(-1, 1, 128), indexing='ij')
    base = 2.7 - 0.35 * torch.exp(-(xx ** 2 + yy ** 2) * 2.5)
    sdepth = base[None, None].repeat(4, 1, 1, 1)
    img = weights.target_image().repeat(4, 1, 1, 1)
    fm = OGeo.face_mask(weights.parsing_mask()).float().repeat(4, 1, 1, 1)
    wr, wm = rotate(sc.cuda(), T(g['rot_tdepth']).cuda(), img.cuda(), c.repeat(4, 1).cuda(), sdepth.cuda(), src_mask=fm.cuda(), EPS=5e-2)
    assert wr.shape == (4, 3, 512, 512) and wm.shape == (4, 1, 512, 512)
    # the |d - z| < EPS test is discontinuous: allow a handful of boundary pixels to flip
    ref_r, ref_m = T(g['rot_rgb_sub']), T(g['rot_mask_sub'])
    bad = ((wm[:, :, 2::8, 3::8].cpu() - ref_m).abs() > 1e-3).float().mean().item()
    assert bad < 2e-3
    assert rel_l2(wr[:, :, 2::8, 3::8], ref_r) < 2e-2
    assert abs(wm.double().sum().item() / float(g['rot_mask_sum']) - 1) < 2e-3


def test_rotate_decisions_are_exact_where_the_oracle_is_decisive():
    """Depth-guided warp (rotate.py:92-116): the in-bounds test and the |d - z| < EPS test are discontinuous, so a comparison with a
    tolerance hides real errors.  Here the reference's chain is evaluated in float64 to find, per pixel, how far it is from either
    threshold; wherever that margin exceeds 1e-4 (float32 arithmetic cannot flip it) the kernel's 0/1 mask must be torch.equal to the
    oracle's, and the warped colours agree to 1e-3 rel-L2 (north_star).  Pixels inside the margin are counted, not compared."""
    import torch.nn.functional as F
    from spi_b200.utils.rotate import rotate
    n, res, eps = 4, 512, 5e-2
    c = weights.canonical_camera(0.3).repeat(n, 1)
    rand = torch.rand(n, 2, generator=torch.Generator().manual_seed(8))
    tcam = OGeo.sample_surrounding_camera(c[:1], rand, yaw_range=0.2, pitch_range=0.1)
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, 128), torch.linspace(-1, 1, 128), indexing='ij')
    sdepth = (2.7 - 0.35 * torch.exp(-(xx ** 2 + yy ** 2) * 2.5))[None, None].repeat(n, 1, 1, 1)
    tdepth = (2.7 - 0.33 * torch.exp(-((xx - 0.05) ** 2 + yy ** 2) * 2.3))[None, None].repeat(n, 1, 1, 1) + 0.02 * torch.sin(7 * xx)[None, None]
    img = weights.target_image().repeat(n, 1, 1, 1)
    fm = OGeo.face_mask(weights.parsing_mask()).float().repeat(n, 1, 1, 1)
    # float64 evaluation of the oracle's own expressions (oracle/geometry.py::rotate), keeping the intermediates
    D = torch.float64
    tex, tin = tcam[:, :16].reshape(n, 4, 4).to(D), tcam[:, 16:].reshape(n, 3, 3).to(D)
    gex, gin = c[:, :16].reshape(n, 4, 4).to(D), c[:, 16:].reshape(n, 3, 3).to(D)
    up = lambda d: F.interpolate(d.to(D).reshape(n, 1, 128, 128), (res, res), mode='bilinear', align_corners=False).reshape(n, res, res)
    td, gd = up(tdepth), up(sdepth)
    uv, z = OGeo.project(OGeo.unproject(td, tex, tin, res), gex, gin)
    grid = 2 * uv.reshape(n, res, res, 2) - 1
    inb = ~((grid[..., 0] < -1) | (grid[..., 0] > 1) | (grid[..., 1] < -1) | (grid[..., 1] > 1))
    src_d = F.grid_sample(gd.reshape(n, 1, res, res), grid, align_corners=False).reshape(n, res, res)
    diff = (src_d - z.reshape(n, res, res)).abs()
    decision = (diff < eps) & inb
    margin = torch.minimum((diff - eps).abs(), (grid.abs() - 1).abs().amin(dim=-1))
    robust = margin > 1e-4
    ref_rgb = F.grid_sample(img.to(D), grid, align_corners=False) * decision.unsqueeze(1)
    ref_m = F.grid_sample(fm.to(D), grid, align_corners=False)
    # kernel, without and with the source mask
    wr0, wm0 = rotate(tcam.cuda(), tdepth.cuda(), img.cuda(), c.cuda(), sdepth.cuda(), src_mask=None, EPS=eps)
    wr1, wm1 = rotate(tcam.cuda(), tdepth.cuda(), img.cuda(), c.cuda(), sdepth.cuda(), src_mask=fm.cuda(), EPS=eps)
    frac = 1 - robust.double().mean().item()
    print(f'pixels within 1e-4 of a threshold: {frac:.2e}; mask coverage {decision.double().mean().item():.3f}')
    assert frac < 5e-3 and 0.05 < decision.double().mean().item() < 0.999
    m0 = wm0[:, 0].cpu()
    assert set(m0.unique().tolist()) <= {0.0, 1.0}
    assert torch.equal(m0[robust] > 0.5, decision[robust])                     # bit-exact decisions
    rb = robust.unsqueeze(1).expand(-1, 3, -1, -1)
    assert rel_l2(wr0.cpu()[rb], ref_rgb[rb]) < 1e-3
    assert rel_l2(wm1[:, 0].cpu()[robust], (ref_m[:, 0] * decision)[robust]) < 1e-5
    assert rel_l2(wr1.cpu()[rb], (ref_rgb * ref_m)[rb]) < 1e-3


def test_adam_matches_torch():
    from spi_b200.optim import FlatAdam
    gen = torch.Generator().manual_seed(0)
    shapes = [(7, 5), (33,), (), (4, 3, 3, 3)]
    ps = [torch.randn(s, generator=gen) for s in shapes]
    ref = [p.clone().requires_grad_(True) for p in ps]
    mine = [p.clone().cuda().requires_grad_(True) for p in ps]
    o_ref = torch.optim.Adam(ref, lr=3e-4)
    o_mine = FlatAdam(mine, lr=3e-4)
    for step in range(5):
        gs = [torch.randn(s, generator=gen) for s in shapes]
        for p, g in zip(ref, gs):
            p.grad = g.clone()
        o_mine.zero_grad()
        for p, g in zip(mine, gs):
            p.grad.add_(g.cuda())
        o_ref.step()
        o_mine.step()
    for a, b in zip(mine, ref):
        assert rel_l2(a, b) < 1e-6
