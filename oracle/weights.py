"""Oracle (CPU, test infrastructure): seeded synthetic weights and inputs.

No checkpoint exists in the container, so every network is filled from
`numpy.random.RandomState` streams keyed by the *tensor name* -- independent of module
construction order, so the reference module, the oracle and the CUDA product all receive
bit-identical tensors.  Shapes/names are the reference's state-dict contract (SURVEY.md §8b.2;
eg3d/training/networks_stylegan2.py:96-357, triplane.py:19-46,112-121, superresolution.py:264-277).
"""
import math
import zlib

import numpy as np
import torch

from .generator import BACKBONE_RES, channels_at


def _rs(name, seed):
    return np.random.RandomState((zlib.crc32(name.encode()) + 7919 * seed) % (2 ** 31))


def _randn(name, seed, shape, scale=1.0, shift=0.0):
    a = _rs(name, seed).standard_normal(size=tuple(shape)).astype(np.float32) * np.float32(scale) + np.float32(shift)
    return torch.from_numpy(np.ascontiguousarray(a))


def resample_filter():
    f = torch.tensor([1., 3., 3., 1.])
    f = torch.outer(f, f)
    return f / f.sum()


def _layer(sd, p, cin, cout, res, k, seed, noise=True):
    sd[p + 'weight'] = _randn(p + 'weight', seed, [cout, cin, k, k])
    if noise:
        sd[p + 'noise_strength'] = _randn(p + 'noise_strength', seed, [], 0.1)
    sd[p + 'bias'] = _randn(p + 'bias', seed, [cout], 0.1)
    if noise:
        sd[p + 'resample_filter'] = resample_filter()
        sd[p + 'noise_const'] = _randn(p + 'noise_const', seed, [res, res])
    sd[p + 'affine.weight'] = _randn(p + 'affine.weight', seed, [cin, 512])
    sd[p + 'affine.bias'] = _randn(p + 'affine.bias', seed, [cin], 0.1, 1.0)


def _block(sd, p, cin, cout, res, img_ch, seed):
    if cin == 0:
        sd[p + 'const'] = _randn(p + 'const', seed, [cout, res, res])
    sd[p + 'resample_filter'] = resample_filter()
    if cin != 0:
        _layer(sd, p + 'conv0.', cin, cout, res, 3, seed)
    _layer(sd, p + 'conv1.', cout, cout, res, 3, seed)
    _layer(sd, p + 'torgb.', cout, img_ch, res, 1, seed, noise=False)


def generator_state_dict(seed=0):
    """132 parameters + 44 buffers, reference order and names."""
    sd = {}
    p = 'backbone.synthesis.'
    for res in BACKBONE_RES:
        _block(sd, f'{p}b{res}.', 0 if res == 4 else channels_at(res // 2), channels_at(res), res, 96, seed)
    m = 'backbone.mapping.'
    sd[m + 'w_avg'] = _randn(m + 'w_avg', seed, [512], 0.1)
    sd[m + 'embed.weight'] = _randn(m + 'embed.weight', seed, [512, 25])
    sd[m + 'embed.bias'] = _randn(m + 'embed.bias', seed, [512], 0.1)
    for i in range(2):
        sd[f'{m}fc{i}.weight'] = _randn(f'{m}fc{i}.weight', seed, [512, 1024 if i == 0 else 512], 100.0)
        sd[f'{m}fc{i}.bias'] = _randn(f'{m}fc{i}.bias', seed, [512], 1.0)
    s = 'superresolution.'
    _block(sd, s + 'block0.', 32, 256, 256, 3, seed)
    _block(sd, s + 'block1.', 256, 128, 512, 3, seed)
    d = 'decoder.net.'
    sd[d + '0.weight'] = _randn(d + '0.weight', seed, [64, 32])
    sd[d + '0.bias'] = _randn(d + '0.bias', seed, [64], 0.1)
    sd[d + '2.weight'] = _randn(d + '2.weight', seed, [33, 64])
    sd[d + '2.bias'] = _randn(d + '2.bias', seed, [33], 0.1)
    return sd


VGG16_CFG = (64, 64, 'M', 128, 128, 'M', 256, 256, 256, 'M', 512, 512, 512, 'M', 512, 512, 512)  # features[:30]
VGG19_HEAD = (64, 64, 'M', 128)                                                                    # features[:6]


def vgg_state_dict(cfg, seed, prefix=''):
    """torchvision `vgg*.features` naming: conv at sequential index i -> '{i}.weight/.bias'.
    He-style scale keeps activations O(1) through 13 layers."""
    sd, idx, cin = {}, 0, 3
    for v in cfg:
        if v == 'M':
            idx += 1
            continue
        sd[f'{prefix}{idx}.weight'] = _randn(f'vgg{idx}.weight', seed, [v, cin, 3, 3], math.sqrt(2.0 / (9 * cin)))
        sd[f'{prefix}{idx}.bias'] = _randn(f'vgg{idx}.bias', seed, [v], 0.05)
        cin = v
        idx += 2
    return sd


def lpips_lin_weights(seed=1):
    """Non-negative 1x1 'lin' weights, U[0,1) (stand-in for richzhang v0.1 vgg.pth)."""
    out = []
    for i, nc in enumerate((64, 128, 256, 512, 512)):
        out.append(torch.from_numpy(_rs(f'lin{i}', seed).uniform(size=(1, nc, 1, 1)).astype(np.float32)))
    return out


# ----------------------------------------------------------------------------- synthetic inputs

def canonical_camera(yaw=0.3, pitch=0.0):
    """cal_canonical_c (spi/utils/camera_utils.py:233-240): radius 2.7, look-at [0,0,0.2], focal 4.2647."""
    from .geometry import look_at_pose
    ext = look_at_pose(torch.tensor([[math.pi / 2 + yaw]]), torch.tensor([[math.pi / 2 - 0.2 + pitch]]),
                       torch.tensor([0., 0., 0.2]), 2.7)
    intr = torch.tensor([[4.2647, 0, 0.5, 0, 4.2647, 0.5, 0, 0, 1]])
    return torch.cat([ext.reshape(1, 16), intr], 1)


def target_image(seed=4, res=512):
    """Smooth seeded 512^2 RGB target in [-1,1] (fallback target of SURVEY.md §8d)."""
    low = _randn('target', seed, [1, 3, 16, 16])
    img = torch.nn.functional.interpolate(low, size=(res, res), mode='bicubic', align_corners=False)
    return (img / img.abs().max()).clamp(-1, 1).contiguous()


def parsing_mask(res=512):
    """int64 [1,1,res,res] face-parsing labels: skin ellipse (1), eyes (4,5), nose (10), lips (12,13),
    hair cap (17), background 0 (formats: spi/data/images_dataset.py, labels: spi/utils/mask_utils.py:4-9)."""
    yy, xx = torch.meshgrid(torch.arange(res, dtype=torch.float32), torch.arange(res, dtype=torch.float32), indexing='ij')
    s = res / 512.0
    m = torch.zeros(res, res, dtype=torch.int64)
    m[((xx - 256 * s) / (150 * s)) ** 2 + ((yy - 200 * s) / (170 * s)) ** 2 < 1] = 17
    m[((xx - 256 * s) / (140 * s)) ** 2 + ((yy - 280 * s) / (180 * s)) ** 2 < 1] = 1
    m[((xx - 190 * s) / (30 * s)) ** 2 + ((yy - 230 * s) / (14 * s)) ** 2 < 1] = 4
    m[((xx - 322 * s) / (30 * s)) ** 2 + ((yy - 230 * s) / (14 * s)) ** 2 < 1] = 5
    m[((xx - 256 * s) / (22 * s)) ** 2 + ((yy - 290 * s) / (40 * s)) ** 2 < 1] = 10
    m[((xx - 256 * s) / (50 * s)) ** 2 + ((yy - 364 * s) / (10 * s)) ** 2 < 1] = 12
    m[((xx - 256 * s) / (50 * s)) ** 2 + ((yy - 384 * s) / (10 * s)) ** 2 < 1] = 13
    return m[None, None]


def landmarks68():
    """Fixed 68x2 (x, y) template at 256 px scale; eye/mouth boxes lie well inside the image."""
    pts = np.zeros((68, 2), dtype=np.float32)
    t = np.linspace(0, np.pi, 17)
    pts[0:17] = np.stack([128 - 70 * np.cos(t), 110 + 90 * np.sin(t)], 1)          # jaw
    pts[17:22] = np.stack([np.linspace(75, 115, 5), np.full(5, 95.)], 1)          # brows
    pts[22:27] = np.stack([np.linspace(141, 181, 5), np.full(5, 95.)], 1)
    pts[27:31] = np.stack([np.full(4, 128.), np.linspace(110, 140, 4)], 1)        # nose
    pts[31:36] = np.stack([np.linspace(116, 140, 5), np.full(5, 150.)], 1)
    e = np.linspace(0, 2 * np.pi, 7)[:6]
    pts[36:42] = np.stack([95 + 12 * np.cos(e), 115 + 5 * np.sin(e)], 1)          # eyes
    pts[42:48] = np.stack([161 + 12 * np.cos(e), 115 + 5 * np.sin(e)], 1)
    m = np.linspace(0, 2 * np.pi, 21)[:20]
    pts[48:68] = np.stack([128 + 25 * np.cos(m), 185 + 9 * np.sin(m)], 1)         # mouth
    return torch.from_numpy(pts)[None]


def w_pivot(seed=5):
    return _randn('w_pivot', seed, [1, 14, 512], 0.5)


def irse50_state_dict(seed=3, prefix='facenet.'):
    """Seeded IR-SE50 (spi/criteria/id_loss/model_irse.py:10-42) state dict with the reference's names; BatchNorm running
    statistics are positive, conv weights He-scaled so that 24 residual units stay O(1)."""
    from .idloss import unit_list
    sd = {}

    def bn(p, c):
        sd[p + 'weight'] = _randn(p + 'weight', seed, [c], 0.1, 1.0)
        sd[p + 'bias'] = _randn(p + 'bias', seed, [c], 0.1)
        sd[p + 'running_mean'] = _randn(p + 'running_mean', seed, [c], 0.1)
        sd[p + 'running_var'] = torch.from_numpy(_rs(p + 'running_var', seed).uniform(0.5, 1.5, size=c).astype(np.float32))
        sd[p + 'num_batches_tracked'] = torch.tensor(0, dtype=torch.long)

    def conv(p, cout, cin, k):
        sd[p] = _randn(p, seed, [cout, cin, k, k], math.sqrt(1.0 / (cin * k * k)))

    p = prefix
    conv(p + 'input_layer.0.weight', 64, 3, 3)
    bn(p + 'input_layer.1.', 64)
    sd[p + 'input_layer.2.weight'] = torch.full([64], 0.25)
    for i, (cin, depth, stride) in enumerate(unit_list()):
        q = f'{p}body.{i}.'
        if cin != depth:
            conv(q + 'shortcut_layer.0.weight', depth, cin, 1)
            bn(q + 'shortcut_layer.1.', depth)
        bn(q + 'res_layer.0.', cin)
        conv(q + 'res_layer.1.weight', depth, cin, 3)
        sd[q + 'res_layer.2.weight'] = torch.full([depth], 0.25)
        conv(q + 'res_layer.3.weight', depth, depth, 3)
        bn(q + 'res_layer.4.', depth)
        conv(q + 'res_layer.5.fc1.weight', depth // 16, depth, 1)
        conv(q + 'res_layer.5.fc2.weight', depth, depth // 16, 1)
    bn(p + 'output_layer.0.', 512)
    sd[p + 'output_layer.3.weight'] = _randn(p + 'output_layer.3.weight', seed, [512, 512 * 7 * 7], math.sqrt(1.0 / (512 * 49)))
    sd[p + 'output_layer.3.bias'] = _randn(p + 'output_layer.3.bias', seed, [512], 0.1)
    bn(p + 'output_layer.4.', 512)
    return sd
