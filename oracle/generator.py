"""Oracle (CPU, test infrastructure): EG3D tri-plane generator forward, functional over a state dict.

The state dict uses the reference's parameter/buffer names, so the same dict
drives the reference module (through `oracle/make_golden.py`), this oracle and the CUDA product.

Follows:
  mapping            eg3d/training/networks_stylegan2.py:233-268 (+ triplane.py:48-51)
  fully connected    networks_stylegan2.py:115-128
  modulated conv     networks_stylegan2.py:34-91 (fused grouped-conv branch: G is always eval(),
                     fused_modconv_default='inference_only' -> SURVEY.md §3.5)
  synthesis layer    networks_stylegan2.py:311-330, ToRGB :352-357, block :417-461, net :503-518
  super-resolution   eg3d/training/superresolution.py:279-290 (SuperresolutionHybrid8XDC)
  ray sampler        eg3d/training/volumetric_rendering/ray_sampler.py:24-61
  renderer           eg3d/training/volumetric_rendering/renderer.py:39-65, 88-253
  ray marcher        eg3d/training/volumetric_rendering/ray_marcher.py:25-57
  decoder            eg3d/training/triplane.py:123-135
  synthesis glue     eg3d/training/triplane.py:53-89
All fp32 (`force_fp32` on CPU, networks_stylegan2.py:421-422).
"""
import math

import torch
import torch.nn.functional as F

from . import ops

RENDERING_DEFAULTS = dict(
    depth_resolution=48, depth_resolution_importance=48, ray_start=2.25, ray_end=3.3, box_warp=1.0,
    c_scale=1.0, c_gen_conditioning_zero=False, clamp_mode='softplus', disparity_space_sampling=False,
    superresolution_noise_mode='none', white_back=False)

BACKBONE_RES = (4, 8, 16, 32, 64, 128, 256)
PLANE_AXES = torch.tensor([[[1, 0, 0], [0, 1, 0], [0, 0, 1]],
                           [[1, 0, 0], [0, 0, 1], [0, 1, 0]],
                           [[0, 0, 1], [1, 0, 0], [0, 1, 0]]], dtype=torch.float32)  # renderer.py:22-37


def channels_at(res, channel_base=32768, channel_max=512):
    return min(channel_base // res, channel_max)


# ----------------------------------------------------------------------------- dense layers

def fully_connected(x, weight, bias, act='linear', lr_mul=1.0):
    """networks_stylegan2.py:115-128."""
    w = weight * (lr_mul / math.sqrt(weight.shape[1]))
    b = bias * lr_mul if (bias is not None and lr_mul != 1) else bias
    if act == 'linear' and b is not None:
        return torch.addmm(b[None], x, w.t())
    return ops.bias_act(x @ w.t(), b, act=act)


def second_moment_normalize(x, eps=1e-8):
    return x * (x.square().mean(1, keepdim=True) + eps).rsqrt()


def mapping(sd, z, c, rk=RENDERING_DEFAULTS, num_ws=14, num_layers=2, truncation_psi=1.0, prefix='backbone.mapping.'):
    """TriPlaneGenerator.mapping (triplane.py:48-51) -> MappingNetwork.forward."""
    if rk.get('c_gen_conditioning_zero', False):
        c = torch.zeros_like(c)
    c = c * rk.get('c_scale', 0)
    x = second_moment_normalize(z.float())
    y = second_moment_normalize(fully_connected(c.float(), sd[prefix + 'embed.weight'], sd[prefix + 'embed.bias']))
    x = torch.cat([x, y], 1)
    for i in range(num_layers):
        x = fully_connected(x, sd[f'{prefix}fc{i}.weight'], sd[f'{prefix}fc{i}.bias'], act='lrelu', lr_mul=0.01)
    x = x[:, None].repeat(1, num_ws, 1)
    if truncation_psi != 1:
        x = sd[prefix + 'w_avg'].lerp(x, truncation_psi)
    return x


# ----------------------------------------------------------------------------- modulated conv

def modulated_conv(x, weight, styles, up=1, padding=0, f=None, demodulate=True, flip_weight=True, noise=None):
    """Fused branch of modulated_conv2d (networks_stylegan2.py:58-91)."""
    n = x.shape[0]
    o, i, kh, kw = weight.shape
    w = weight[None] * styles.reshape(n, 1, i, 1, 1)
    if demodulate:
        d = (w.square().sum(dim=[2, 3, 4]) + 1e-8).rsqrt()
        w = w * d.reshape(n, o, 1, 1, 1)
    y = ops.conv2d_resample(x.reshape(1, n * i, *x.shape[2:]), w.reshape(n * o, i, kh, kw), f=f, up=up,
                            padding=padding, groups=n, flip_weight=flip_weight)
    y = y.reshape(n, o, *y.shape[2:])
    if noise is not None:
        y = y + noise
    return y


def synthesis_layer(sd, p, x, w, up=1, noise_mode='const', clamp=None, gain=1.0):
    """SynthesisLayer.forward (networks_stylegan2.py:311-330)."""
    styles = fully_connected(w, sd[p + 'affine.weight'], sd[p + 'affine.bias'])
    noise = None
    if noise_mode == 'const':
        noise = sd[p + 'noise_const'] * sd[p + 'noise_strength']
    x = modulated_conv(x, sd[p + 'weight'], styles, up=up, padding=1, f=sd[p + 'resample_filter'],
                       flip_weight=(up == 1), noise=noise)
    return ops.bias_act(x, sd[p + 'bias'], act='lrelu', gain=math.sqrt(2) * gain,
                        clamp=None if clamp is None else clamp * gain)


def to_rgb(sd, p, x, w, clamp=None):
    """ToRGBLayer.forward (networks_stylegan2.py:352-357)."""
    cin = sd[p + 'weight'].shape[1]
    styles = fully_connected(w, sd[p + 'affine.weight'], sd[p + 'affine.bias']) * (1 / math.sqrt(cin))
    x = modulated_conv(x, sd[p + 'weight'], styles, demodulate=False)
    return ops.bias_act(x, sd[p + 'bias'], clamp=clamp)


def synthesis_block(sd, p, x, img, ws, first=False, noise_mode='const', clamp=None):
    """SynthesisBlock.forward, 'skip' architecture (networks_stylegan2.py:417-461)."""
    n = ws.shape[0]
    wi = 0
    if first:
        x = sd[p + 'const'][None].repeat(n, 1, 1, 1)
    else:
        x = synthesis_layer(sd, p + 'conv0.', x, ws[:, wi], up=2, noise_mode=noise_mode, clamp=clamp)
        wi += 1
    x = synthesis_layer(sd, p + 'conv1.', x, ws[:, wi], noise_mode=noise_mode, clamp=clamp)
    wi += 1
    if img is not None:
        img = ops.upsample2d(img, sd[p + 'resample_filter'])
    y = to_rgb(sd, p + 'torgb.', x, ws[:, wi], clamp=clamp)
    img = y if img is None else img + y
    return x, img


def backbone_synthesis(sd, ws, noise_mode='const', prefix='backbone.synthesis.', clamp=None):
    """SynthesisNetwork.forward (networks_stylegan2.py:503-518): ws[B,14,512] -> planes [B,96,256,256]."""
    x = img = None
    wi = 0
    for res in BACKBONE_RES:
        nconv = 1 if res == 4 else 2
        x, img = synthesis_block(sd, f'{prefix}b{res}.', x, img, ws[:, wi: wi + nconv + 1], first=(res == 4),
                                 noise_mode=noise_mode, clamp=clamp)
        wi += nconv
    return img


def superresolution(sd, rgb, x, ws, noise_mode='none', prefix='superresolution.', clamp=256):
    """SuperresolutionHybrid8XDC.forward (superresolution.py:279-290); input already 128x128.
    conv_clamp = 256 because sr_num_fp16_res = 4 > 0 (superresolution.py:272-276); it is applied in fp32 too."""
    assert x.shape[-1] == 128, 'oracle covers neural_rendering_resolution == 128 (load_utils.py:31)'
    w3 = ws[:, -1:, :].repeat(1, 3, 1)
    x, rgb = synthesis_block(sd, prefix + 'block0.', x, rgb, w3, noise_mode=noise_mode, clamp=clamp)
    x, rgb = synthesis_block(sd, prefix + 'block1.', x, rgb, w3, noise_mode=noise_mode, clamp=clamp)
    return rgb


# ----------------------------------------------------------------------------- rays

def ray_sampler(cam2world, intrinsics, res):
    """ray_sampler.py:24-61.  Ray m = i*res + j is pixel row i, column j; uv at pixel centres."""
    n = cam2world.shape[0]
    fx, fy = intrinsics[:, 0, 0:1], intrinsics[:, 1, 1:2]
    cx, cy, sk = intrinsics[:, 0, 2:3], intrinsics[:, 1, 2:3], intrinsics[:, 0, 1:2]
    t = torch.arange(res, dtype=torch.float32) * (1. / res) + (0.5 / res)
    y_cam = t[:, None].expand(res, res).reshape(1, -1).expand(n, -1)
    x_cam = t[None, :].expand(res, res).reshape(1, -1).expand(n, -1)
    z_cam = torch.ones(n, res * res)
    x_lift = (x_cam - cx + cy * sk / fy - sk * y_cam / fy) / fx * z_cam
    y_lift = (y_cam - cy) / fy * z_cam
    pts = torch.stack([x_lift, y_lift, z_cam, torch.ones_like(z_cam)], -1)
    world = torch.bmm(cam2world, pts.permute(0, 2, 1)).permute(0, 2, 1)[:, :, :3]
    origin = cam2world[:, :3, 3]
    dirs = F.normalize(world - origin[:, None], dim=2)
    return origin[:, None].repeat(1, res * res, 1), dirs


# ----------------------------------------------------------------------------- renderer

def sample_from_planes(planes, coords, box_warp):
    """renderer.py:39-65.  planes [N,3,C,H,W], coords [N,M,3] -> [N,3,M,C]."""
    n, p, c, h, w = planes.shape
    m = coords.shape[1]
    coords = (2 / box_warp) * coords
    xyz = coords[:, None].expand(-1, p, -1, -1).reshape(n * p, m, 3)
    inv = torch.linalg.inv(PLANE_AXES)[None].expand(n, -1, -1, -1).reshape(n * p, 3, 3)
    uv = torch.bmm(xyz, inv)[..., :2]
    out = F.grid_sample(planes.reshape(n * p, c, h, w), uv[:, None].float(), mode='bilinear',
                        padding_mode='zeros', align_corners=False)
    return out.permute(0, 3, 2, 1).reshape(n, p, m, c)


def osg_decoder(sd, feats, prefix='decoder.'):
    """OSGDecoder.forward (triplane.py:123-135): feats [N,3,M,32] -> rgb [N,M,32], sigma [N,M,1]."""
    x = feats.mean(1)
    n, m, c = x.shape
    x = x.reshape(n * m, c)
    x = fully_connected(x, sd[prefix + 'net.0.weight'], sd[prefix + 'net.0.bias'])
    x = F.softplus(x)
    x = fully_connected(x, sd[prefix + 'net.2.weight'], sd[prefix + 'net.2.bias'])
    x = x.reshape(n, m, -1)
    return torch.sigmoid(x[..., 1:]) * (1 + 2 * 0.001) - 0.001, x[..., 0:1]


def run_model(sd, planes, coords, rk):
    """renderer.py:142-149 (density_noise == 0)."""
    return osg_decoder(sd, sample_from_planes(planes, coords, rk['box_warp']))


def ray_march(colors, sigmas, depths, rk):
    """MipRayMarcher2.run_forward (ray_marcher.py:25-57)."""
    deltas = depths[:, :, 1:] - depths[:, :, :-1]
    c_mid = (colors[:, :, :-1] + colors[:, :, 1:]) / 2
    s_mid = (sigmas[:, :, :-1] + sigmas[:, :, 1:]) / 2
    d_mid = (depths[:, :, :-1] + depths[:, :, 1:]) / 2
    assert rk.get('clamp_mode', 'softplus') == 'softplus'
    s_mid = F.softplus(s_mid - 1)
    alpha = 1 - torch.exp(-(s_mid * deltas))
    shifted = torch.cat([torch.ones_like(alpha[:, :, :1]), 1 - alpha + 1e-10], -2)
    weights = alpha * torch.cumprod(shifted, -2)[:, :, :-1]
    rgb = torch.sum(weights * c_mid, -2)
    total = weights.sum(2)
    depth = torch.sum(weights * d_mid, -2) / total
    depth = torch.nan_to_num(depth, float('inf'))
    depth = torch.clamp(depth, torch.min(depths), torch.max(depths))  # whole-tensor min/max
    if rk.get('white_back', False):
        rgb = rgb + 1 - total
    return rgb * 2 - 1, depth, weights


def stratified_depths(n, m, rk, jitter):
    """sample_stratified, scalar ray_start/ray_end branch (renderer.py:186-190).
    `jitter` [n,m,Dc,1] in [0,1) replaces `torch.rand_like` (:190)."""
    dc = rk['depth_resolution']
    start, end = rk['ray_start'], rk['ray_end']
    if rk.get('disparity_space_sampling', False):
        d = torch.linspace(0, 1, dc).reshape(1, 1, dc, 1).repeat(n, m, 1, 1)
        d = d + jitter * (1 / (dc - 1))
        return 1. / (1. / start * (1. - d) + 1. / end * d)
    d = torch.linspace(start, end, dc).reshape(1, 1, dc, 1).repeat(n, m, 1, 1)
    return d + jitter * ((end - start) / (dc - 1))


def importance_cdf(z_vals, weights):
    """Smoothing + pdf/cdf half of sample_importance/sample_pdf (renderer.py:194-232).
    z_vals [R,Dc], weights [R,Dc-1] -> bins [R,Dc-1], cdf [R,Dc-2]."""
    w = F.max_pool1d(weights[:, None].float(), 2, 1, padding=1)
    w = F.avg_pool1d(w, 2, 1).squeeze(1)
    w = w + 0.01
    bins = 0.5 * (z_vals[:, :-1] + z_vals[:, 1:])
    w = w[:, 1:-1] + 1e-5
    pdf = w / torch.sum(w, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    return bins, torch.cat([torch.zeros_like(cdf[:, :1]), cdf], -1)


def inverse_cdf(bins, cdf, u, eps=1e-5):
    """Inverse-CDF half of sample_pdf (renderer.py:241-253).  Returns (samples, inds int64)."""
    ns = cdf.shape[1] - 1
    inds = torch.searchsorted(cdf, u.contiguous(), right=True)
    below = torch.clamp_min(inds - 1, 0)
    above = torch.clamp_max(inds, ns)
    cb, ca = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
    bb, ba = torch.gather(bins, 1, below), torch.gather(bins, 1, above)
    denom = ca - cb
    denom = torch.where(denom < eps, torch.ones_like(denom), denom)
    return bb + (u - cb) / denom * (ba - bb), inds


def unify_samples(d1, c1, s1, d2, c2, s2):
    """renderer.py:157-167.  Returns sorted (depths, colors, sigmas, permutation int64)."""
    d = torch.cat([d1, d2], -2)
    c = torch.cat([c1, c2], -2)
    s = torch.cat([s1, s2], -2)
    _, idx = torch.sort(d, dim=-2)
    return (torch.gather(d, -2, idx), torch.gather(c, -2, idx.expand(-1, -1, -1, c.shape[-1])),
            torch.gather(s, -2, idx), idx)


def importance_render(sd, planes, origins, dirs, rk, jitter, u, return_aux=False):
    """ImportanceRenderer.forward (renderer.py:88-140) with injected randomness:
    jitter [N,R,Dc,1] for `rand_like` (:190) and u [N*R,Df] for `torch.rand` (:237)."""
    n, r, _ = origins.shape
    dc, df = rk['depth_resolution'], rk['depth_resolution_importance']
    d_coarse = stratified_depths(n, r, rk, jitter)
    xyz = (origins[:, :, None] + d_coarse * dirs[:, :, None]).reshape(n, -1, 3)
    rgb_c, sig_c = run_model(sd, planes, xyz, rk)
    rgb_c = rgb_c.reshape(n, r, dc, -1)
    sig_c = sig_c.reshape(n, r, dc, 1)
    aux = {'depths_coarse': d_coarse, 'sigma_coarse': sig_c}
    if df > 0:
        _, _, w = ray_march(rgb_c, sig_c, d_coarse, rk)
        with torch.no_grad():
            bins, cdf = importance_cdf(d_coarse.reshape(n * r, dc), w.reshape(n * r, -1))
            d_fine, inds = inverse_cdf(bins, cdf, u)
            d_fine = d_fine.detach().reshape(n, r, df, 1)
        xyz = (origins[:, :, None] + d_fine * dirs[:, :, None]).reshape(n, -1, 3)
        rgb_f, sig_f = run_model(sd, planes, xyz, rk)
        rgb_f = rgb_f.reshape(n, r, df, -1)
        sig_f = sig_f.reshape(n, r, df, 1)
        d_all, c_all, s_all, perm = unify_samples(d_coarse, rgb_c, sig_c, d_fine, rgb_f, sig_f)
        rgb, depth, w = ray_march(c_all, s_all, d_all, rk)
        aux.update(weights_coarse=w, cdf=cdf, bins=bins, inds=inds, depths_fine=d_fine, perm=perm, depths_all=d_all,
                   sigma_all=s_all)
    else:
        rgb, depth, w = ray_march(rgb_c, sig_c, d_coarse, rk)
    if return_aux:
        return rgb, depth, w.sum(2), aux
    return rgb, depth, w.sum(2)


def make_render_noise(n, r, rk, seed):
    """Seeded stand-ins for the two RNG draws of one render (renderer.py:190,237)."""
    g = torch.Generator().manual_seed(seed)
    jitter = torch.rand(n, r, rk['depth_resolution'], 1, generator=g)
    u = torch.rand(n * r, max(rk['depth_resolution_importance'], 1), generator=g)
    return jitter, u


# ----------------------------------------------------------------------------- full synthesis

def synthesis(sd, ws, c, rk=RENDERING_DEFAULTS, nrr=128, noise_mode='const', jitter=None, u=None, seed=0,
              return_aux=False):
    """TriPlaneGenerator.synthesis (triplane.py:53-89)."""
    rk = {**RENDERING_DEFAULTS, **rk}
    n = ws.shape[0]
    cam2world = c[:, :16].reshape(-1, 4, 4)
    intrinsics = c[:, 16:25].reshape(-1, 3, 3)
    origins, dirs = ray_sampler(cam2world, intrinsics, nrr)
    if jitter is None:
        jitter, u = make_render_noise(n, nrr * nrr, rk, seed)
    planes = backbone_synthesis(sd, ws, noise_mode=noise_mode)
    planes5 = planes.reshape(n, 3, 32, planes.shape[-2], planes.shape[-1])
    out = importance_render(sd, planes5, origins, dirs, rk, jitter, u, return_aux=return_aux)
    feat, depth = out[0], out[1]
    feat_img = feat.permute(0, 2, 1).reshape(n, feat.shape[-1], nrr, nrr).contiguous()
    depth_img = depth.permute(0, 2, 1).reshape(n, 1, nrr, nrr)
    rgb = feat_img[:, :3]
    image = superresolution(sd, rgb, feat_img, ws, noise_mode=rk['superresolution_noise_mode'])
    res = {'image': image, 'image_raw': rgb, 'image_depth': depth_img, 'planes': planes, 'feature_image': feat_img}
    if return_aux:
        res['aux'] = out[3]
    return res


def sample_mixed(sd, coords, ws, rk=RENDERING_DEFAULTS, noise_mode='const'):
    """TriPlaneGenerator.sample_mixed (triplane.py:98-102)."""
    rk = {**RENDERING_DEFAULTS, **rk}
    planes = backbone_synthesis(sd, ws, noise_mode=noise_mode)
    planes = planes.reshape(len(planes), 3, 32, planes.shape[-2], planes.shape[-1])
    rgb, sigma = run_model(sd, planes, coords, rk)
    return {'rgb': rgb, 'sigma': sigma}
