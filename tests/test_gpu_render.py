"""GPU parity: fused renderer (forward, backward, stand-alone stages) vs the CPU oracle / reference goldens."""
import numpy as np
import pytest
import torch

from conftest import rel_l2
from oracle import generator as OG
from oracle import weights

pytestmark = pytest.mark.gpu


def T(a):
    return torch.from_numpy(np.asarray(a))


def decoder_module(sd):
    from spi_b200.training.triplane import OSGDecoder
    d = OSGDecoder(32, {'decoder_lr_mul': 1, 'decoder_output_dim': 32})
    d.load_state_dict({k[len('decoder.'):]: v for k, v in sd.items() if k.startswith('decoder.')})
    return d.cuda()


def test_ray_sampler_golden(golden):
    from spi_b200.training.volumetric_rendering.ray_sampler import RaySampler
    g = golden('render')
    cam = T(g['cam']).cuda()
    o, d = RaySampler()(cam[:, :16].reshape(-1, 4, 4), cam[:, 16:].reshape(-1, 3, 3), 128)
    assert rel_l2(o[:, ::37], g['ray_origins_sub']) < 1e-6 and rel_l2(d[:, ::37], g['ray_dirs_sub']) < 1e-6


@pytest.mark.parametrize('dc,df', [(48, 48), (32, 32), (12, 20)])
def test_render_forward_golden(golden, gen_sd, dc, df):
    """Reference outputs on small random planes; tolerance 1e-4 rel-L2 (fp32, different summation order)."""
    from spi_b200.training.volumetric_rendering.renderer import ImportanceRenderer
    g = golden('render')
    cam = T(g['cam'])
    o, d = OG.ray_sampler(cam[:, :16].reshape(-1, 4, 4), cam[:, 16:].reshape(-1, 3, 3), 128)
    sel = T(g['sel'])
    o, d = o[:, sel].contiguous().cuda(), d[:, sel].contiguous().cuda()
    tag = f'r{dc}_{df}_'
    rk = dict(OG.RENDERING_DEFAULTS, depth_resolution=dc, depth_resolution_importance=df)
    R = ImportanceRenderer()
    R.inject_noise(T(g[tag + 'jit']).cuda(), T(g[tag + 'u']).cuda())
    rgb, depth, wsum = R(T(g['planes']).cuda(), decoder_module(gen_sd), o, d, rk)
    assert rel_l2(rgb, g[tag + 'rgb']) < 1e-4
    assert rel_l2(depth, g[tag + 'depth']) < 1e-5
    assert rel_l2(wsum, g[tag + 'wsum']) < 1e-4


@pytest.mark.parametrize('dc,df', [(48, 48), (32, 32), (12, 20)])
def test_index_stages_bit_exact(golden, dc, df):
    """searchsorted bins and the merge permutation are bit-exact given the reference's own float inputs."""
    from spi_b200.training.volumetric_rendering.renderer import ImportanceRenderer
    g = golden('render')
    tag = f'r{dc}_{df}_'
    R = ImportanceRenderer()
    fine, inds = R.sample_pdf_from_cdf(T(g[tag + 'bins']).cuda(), T(g[tag + 'cdf']).cuda(), T(g[tag + 'u']).cuda())
    assert torch.equal(inds.cpu(), T(g[tag + 'inds']))
    assert rel_l2(fine, g[tag + 'fine']) < 1e-6
    n, r = 2, 192
    d_c = OG.stratified_depths(n, r, dict(OG.RENDERING_DEFAULTS, depth_resolution=dc), T(g[tag + 'jit']))
    d_f = T(g[tag + 'fine']).reshape(n, r, df, 1)
    perm, srt = R.sort_permutation(d_c.cuda(), d_f.cuda())
    assert torch.equal(perm.cpu(), T(g[tag + 'perm']))
    assert torch.equal(srt.cpu(), T(g[tag + 'depths_all']))


def test_ray_marcher_golden(golden):
    from spi_b200.training.volumetric_rendering.ray_marcher import MipRayMarcher2
    g = golden('render')
    rgb, depth, w = MipRayMarcher2()(T(g['rm_col']).cuda(), T(g['rm_sig']).cuda(), T(g['rm_dep']).cuda(), OG.RENDERING_DEFAULTS)
    assert rel_l2(rgb, g['rm_rgb']) < 1e-5 and rel_l2(depth, g['rm_depth']) < 1e-5 and rel_l2(w, g['rm_w']) < 1e-5


@pytest.mark.parametrize('n,r,dc,df', [
    (2, 160, 16, 16),      # tcgen05 kernels, full 4-ray groups
    (1, 37, 32, 32),       # tcgen05 kernels, ragged last group (dead ray slots), both 32-sample rounds full
    (1, 30, 24, 20),       # tcgen05 kernels, partially filled rounds in both passes
    (1, 24, 40, 40),       # 80 merged samples = 3 rounds: tcgen05 backward re-runs the forward part of its oldest round
    (1, 12, 72, 72),       # 144 merged samples: tcgen05 forward (6 rounds), mma.sync backward
])
def test_render_backward_vs_oracle(gen_sd, n, r, dc, df):
    """Gradients w.r.t. planes and the four decoder tensors against autograd through the CPU oracle."""
    from spi_b200.training.volumetric_rendering.renderer import ImportanceRenderer
    gen = torch.Generator().manual_seed(11)
    rk = dict(OG.RENDERING_DEFAULTS, depth_resolution=dc, depth_resolution_importance=df)
    planes = torch.randn(n, 3, 32, 48, 48, generator=gen)
    cam = torch.cat([weights.canonical_camera(0.3), weights.canonical_camera(-0.2, 0.1)], 0)[:n]
    o, d = OG.ray_sampler(cam[:, :16].reshape(-1, 4, 4), cam[:, 16:].reshape(-1, 3, 3), 128)
    sel = torch.arange(r) * 97 + 700
    o, d = o[:, sel].contiguous(), d[:, sel].contiguous()
    jit, u = torch.rand(n, r, dc, 1, generator=gen), torch.rand(n * r, df, generator=gen)
    g_rgb, g_depth = torch.randn(n, r, 32, generator=gen), torch.randn(n, r, 1, generator=gen)
    keys = ['decoder.net.0.weight', 'decoder.net.0.bias', 'decoder.net.2.weight', 'decoder.net.2.bias']
    sd = {k: v.clone().requires_grad_(k in keys) for k, v in gen_sd.items() if k.startswith('decoder.')}
    po = planes.clone().requires_grad_(True)
    rgb_o, depth_o, _ = OG.importance_render(sd, po, o, d, rk, jit, u)
    ((rgb_o * g_rgb).sum() + (depth_o * g_depth).sum()).backward()
    dec = decoder_module(gen_sd).requires_grad_(True)
    pg = planes.cuda().requires_grad_(True)
    R = ImportanceRenderer()
    R.inject_noise(jit.cuda(), u.cuda())
    rgb_g, depth_g, _ = R(pg, dec, o.cuda(), d.cuda(), rk)
    assert rel_l2(rgb_g, rgb_o) < 1e-4 and rel_l2(depth_g, depth_o) < 1e-5
    ((rgb_g * g_rgb.cuda()).sum() + (depth_g * g_depth.cuda()).sum()).backward()
    assert rel_l2(pg.grad, po.grad) < 1e-3
    got = {'decoder.net.0.weight': dec.net[0].weight.grad, 'decoder.net.0.bias': dec.net[0].bias.grad,
           'decoder.net.2.weight': dec.net[2].weight.grad, 'decoder.net.2.bias': dec.net[2].bias.grad}
    for k in keys:
        assert rel_l2(got[k], sd[k].grad) < 1e-3, k


def test_render_backward_recompute_mode_matches_kept_mode(gen_sd):
    """Decoder gradients with the forward pass keeping its activations (default) and with the backward pass recomputing them
    (KEEP_ACTIVATIONS = False: per-sample rows written in merged order by the backward kernel) agree."""
    from spi_b200.training.volumetric_rendering import renderer as RM
    gen = torch.Generator().manual_seed(21)
    n, r, dc, df = 1, 50, 32, 32
    rk = dict(OG.RENDERING_DEFAULTS, depth_resolution=dc, depth_resolution_importance=df)
    planes = torch.randn(n, 3, 32, 48, 48, generator=gen)
    cam = weights.canonical_camera(0.3)
    o, d = OG.ray_sampler(cam[:, :16].reshape(-1, 4, 4), cam[:, 16:].reshape(-1, 3, 3), 128)
    sel = torch.arange(r) * 131 + 900
    o, d = o[:, sel].contiguous().cuda(), d[:, sel].contiguous().cuda()
    jit, u = torch.rand(n, r, dc, 1, generator=gen).cuda(), torch.rand(n * r, df, generator=gen).cuda()
    g_rgb, g_depth = torch.randn(n, r, 32, generator=gen).cuda(), torch.randn(n, r, 1, generator=gen).cuda()
    res = {}
    for keep in (True, False):
        RM.KEEP_ACTIVATIONS = keep
        try:
            dec = decoder_module(gen_sd).requires_grad_(True)
            pg = planes.cuda().requires_grad_(True)
            R = RM.ImportanceRenderer()
            R.inject_noise(jit, u)
            rgb, depth, _ = R(pg, dec, o, d, rk)
            ((rgb * g_rgb).sum() + (depth * g_depth).sum()).backward()
            res[keep] = [pg.grad] + [p.grad for p in dec.parameters()]
        finally:
            RM.KEEP_ACTIVATIONS = True
    for a, b in zip(res[True], res[False]):
        assert rel_l2(a, b) < 2e-4


def test_run_model_vs_oracle(gen_sd):
    from spi_b200.training.volumetric_rendering.renderer import ImportanceRenderer
    gen = torch.Generator().manual_seed(12)
    planes = torch.randn(2, 3, 32, 40, 40, generator=gen)
    pts = torch.rand(2, 777, 3, generator=gen) * 1.3 - 0.65      # some points fall outside the box
    rk = OG.RENDERING_DEFAULTS
    sd = {k: v.clone() for k, v in gen_sd.items() if k.startswith('decoder.')}
    po = planes.clone().requires_grad_(True)
    rgb_o, sig_o = OG.run_model(sd, po, pts, rk)
    (rgb_o.sum() + (sig_o ** 2).sum()).backward()
    pg = planes.cuda().requires_grad_(True)
    out = ImportanceRenderer().run_model(pg, decoder_module(gen_sd), pts.cuda(), None, rk)
    assert rel_l2(out['rgb'], rgb_o) < 1e-5 and rel_l2(out['sigma'], sig_o) < 1e-5
    (out['rgb'].sum() + (out['sigma'] ** 2).sum()).backward()
    assert rel_l2(pg.grad, po.grad) < 1e-4


def test_render_edge_cases(gen_sd):
    """Zero importance samples, rays that miss the box entirely (all samples read zeros), empty batch."""
    from spi_b200.training.volumetric_rendering.renderer import ImportanceRenderer
    gen = torch.Generator().manual_seed(13)
    planes = torch.randn(1, 3, 32, 32, 32, generator=gen)
    o = torch.tensor([[[0., 0., 2.7], [5., 5., 5.]]])
    d = torch.tensor([[[0., 0., -1.], [0., 1., 0.]]])
    for dc, df in ((8, 0), (8, 8)):
        rk = dict(OG.RENDERING_DEFAULTS, depth_resolution=dc, depth_resolution_importance=df)
        jit, u = torch.rand(1, 2, dc, 1, generator=gen), torch.rand(2, max(df, 1), generator=gen)
        sd = {k: v for k, v in gen_sd.items() if k.startswith('decoder.')}
        ro, do, wo = OG.importance_render(sd, planes, o, d, rk, jit, u)
        R = ImportanceRenderer()
        R.inject_noise(jit.cuda(), u.cuda())
        rg, dg, wg = R(planes.cuda(), decoder_module(gen_sd), o.cuda(), d.cuda(), rk)
        assert rel_l2(rg, ro) < 1e-4 and rel_l2(dg, do) < 1e-5 and rel_l2(wg, wo) < 1e-4
