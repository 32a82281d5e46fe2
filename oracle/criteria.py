"""Oracle (CPU, test infrastructure): the losses of the inversion loop.

Follows:
  LPIPS (VGG16)        spi/criteria/lpips/lpips.py:32-71, networks.py:53-63,88-96, utils.py:6-8
  vgg16.pt features    RESTATEMENT of the third-party TorchScript artefact `checkpoints/vgg16.pt`
                       (spi/configs/paths_config.py:5; called at spi/training/projectors/w_projector.py:51,86).
                       The artefact is absent (no version pin) -> "parity unpinned" for this one function
                       (SURVEY.md §8c-i); it is restated from the LPIPS modules above with the [0,255] input
                       convention and sqrt(lin)-scaled, spatially averaged unit features so that the squared
                       distance of two feature vectors equals LPIPS.
  BoxCX                spi/criteria/bbox_cx_loss.py:20-59,76-182
  L2                   spi/criteria/l2_loss.py:3-8
  noise regulariser    spi/training/projectors/mirror_projector.py:107-115,128-131
"""
import torch
import torch.nn.functional as F
from torchvision.ops import roi_align

from .weights import VGG16_CFG, VGG19_HEAD

LPIPS_TAPS = (4, 9, 16, 23, 30)  # sequential indices after which features are tapped (networks.py:93)


def vgg_features(x, sd, cfg, taps=None, prefix='', final_relu=True):
    """torchvision `vgg.features` prefix: conv3x3+ReLU / maxpool2, tapped after module index in `taps`.
    `final_relu=False` stops right after the last conv (VGG19 `features[:6]`, bbox_cx_loss.py:79-82)."""
    outs, idx = [], 0
    for j, v in enumerate(cfg):
        if v == 'M':
            x = F.max_pool2d(x, 2)
            idx += 1
        else:
            x = F.conv2d(x, sd[f'{prefix}{idx}.weight'], sd[f'{prefix}{idx}.bias'], padding=1)
            if final_relu or j != len(cfg) - 1:
                x = F.relu(x)
            idx += 2
        if taps is not None and idx in taps:
            outs.append(x)
    return outs if taps is not None else x


def unit_normalize(x, eps=1e-10):
    return x / (torch.sqrt(torch.sum(x ** 2, dim=1, keepdim=True)) + eps)


LPIPS_MEAN = torch.tensor([-.030, -.088, -.188]).reshape(1, 3, 1, 1)
LPIPS_STD = torch.tensor([.458, .448, .450]).reshape(1, 3, 1, 1)


def lpips_taps(x, vgg_sd):
    x = (x - LPIPS_MEAN) / LPIPS_STD
    return [unit_normalize(f) for f in vgg_features(x, vgg_sd, VGG16_CFG, taps=LPIPS_TAPS)]


def lpips(x, y, vgg_sd, lin):
    """LPIPS.forward (lpips.py:32-71), num_scales=1, no mask/conf."""
    n = x.shape[0]
    if x.shape[-1] > 256:
        x = F.interpolate(x, size=(256, 256), mode='bilinear', align_corners=False)
        y = F.interpolate(y, size=(256, 256), mode='bilinear', align_corners=False)
    fx, fy = lpips_taps(x, vgg_sd), lpips_taps(y, vgg_sd)
    res = [F.conv2d((a - b) ** 2, w).mean((2, 3), True) for a, b, w in zip(fx, fy, lin)]
    return torch.sum(torch.cat(res, 0)) / n


def vgg16_pt_features(img255, vgg_sd, lin):
    """Restated `vgg16.pt(img, resize_images=False, return_lpips=True)`; img in [0,255], 256^2."""
    x = img255 / 127.5 - 1
    feats = []
    for f, w in zip(lpips_taps(x, vgg_sd), lin):
        h, wd = f.shape[2:]
        feats.append((f * torch.sqrt(w) / (h * wd) ** 0.5).reshape(f.shape[0], -1))
    return torch.cat(feats, 1)


def l2(a, b):
    return torch.mean((a - b) ** 2)


# ----------------------------------------------------------------------------- BoxCX

def landmark_boxes(lm):
    """get_landmark_bbox (bbox_cx_loss.py:20-37): int64 boxes [ly, lx, ry, rx]; padding 8 for the mouth, 15 for
    the eyes and -- because `p` is never reset -- 15 for the (unused) nose box."""
    pad = 8
    boxes = []
    for k, (a, b) in enumerate(((48, 68), (36, 42), (42, 48), (27, 36))):
        pts = lm[:, a:b]
        ly, ry = pts[:, :, 0].min(1)[0].long(), pts[:, :, 0].max(1)[0].long()
        lx, rx = pts[:, :, 1].min(1)[0].long(), pts[:, :, 1].max(1)[0].long()
        if k in (1, 2):
            pad = 15
        boxes.append(torch.stack([ly - pad, lx - pad, ry + pad, rx + pad], 1))
    return boxes


def crop_boxes(image, fake, lm):
    """get_bbox (bbox_cx_loss.py:41-59): roi_align(output 80, scale 1, adaptive sampling, aligned=False)."""
    assert image.shape[-1] == 256
    boxes = landmark_boxes(lm)
    idx = torch.arange(image.shape[0])[:, None]
    out = []
    for k in range(3):
        rois = torch.cat([idx, boxes[k]], 1).float()
        out.append((roi_align(image, boxes=rois, output_size=80), roi_align(fake, boxes=rois, output_size=80)))
    return out


def contextual_loss(fx, fy, band_width=0.5):
    """compute_cosine_distance / relative_distance / cx (bbox_cx_loss.py:93-130,174-178)."""
    mu = fy.mean(dim=(0, 2, 3), keepdim=True)
    xn = F.normalize(fx - mu, p=2, dim=1)
    yn = F.normalize(fy - mu, p=2, dim=1)
    n, c = fx.shape[:2]
    d = 1 - torch.bmm(xn.reshape(n, c, -1).transpose(1, 2), yn.reshape(n, c, -1))
    dt = d / (d.min(dim=2, keepdim=True)[0] + 1e-5)
    dt = torch.clamp(dt, max=10., min=-10)
    w = torch.exp((1 - dt) / band_width)
    cx = w / torch.sum(w, dim=2, keepdim=True)
    cx = torch.mean(torch.max(cx, dim=1)[0], dim=1)
    return torch.mean(-torch.log(cx + 1e-5))


VGG_MEAN = torch.tensor([0.485, 0.456, 0.406]).reshape(1, 3, 1, 1)
VGG_STD = torch.tensor([0.229, 0.224, 0.225]).reshape(1, 3, 1, 1)


def box_cx(x, y, lm, vgg19_sd):
    """BoxCXLoss.forward (bbox_cx_loss.py:159-182)."""
    if x.shape[-1] > 256:
        x = F.interpolate(x, (256, 256), mode='bilinear', align_corners=False)
    if y.shape[-1] > 256:
        y = F.interpolate(y, (256, 256), mode='bilinear', align_corners=False)
    x = (x - VGG_MEAN) / VGG_STD
    y = (y - VGG_MEAN) / VGG_STD
    loss = 0
    for a, b in crop_boxes(x, y, lm):
        loss = loss + contextual_loss(vgg_features(a, vgg19_sd, VGG19_HEAD, final_relu=False),
                                       vgg_features(b, vgg19_sd, VGG19_HEAD, final_relu=False))
    return loss * 0.1


# ----------------------------------------------------------------------------- projector noise terms

def noise_regulariser(noise_bufs):
    """mirror_projector.py:107-115 (same in w_/w_plus_projector)."""
    reg = 0.0
    for v in noise_bufs:
        n = v[None, None]
        while True:
            reg = reg + (n * torch.roll(n, shifts=1, dims=3)).mean() ** 2
            reg = reg + (n * torch.roll(n, shifts=1, dims=2)).mean() ** 2
            if n.shape[2] <= 8:
                break
            n = F.avg_pool2d(n, kernel_size=2)
    return reg


def renormalise_noise_(noise_bufs):
    """mirror_projector.py:128-131."""
    with torch.no_grad():
        for b in noise_bufs:
            b -= b.mean()
            b *= b.square().mean().rsqrt()
