"""Mirror-view contextual loss on landmark boxes (drop-in for spi/criteria/bbox_cx_loss.py:20-182).

Landmark boxes (int64, the reference's truncation and padding quirks kept), `roi_align` 80x80 crops, VGG19[:6]
features (conv engine), centred cosine distance, relative distance, CX, -log mean max.
"""
import torch
import torch.nn.functional as F

from .. import _lib
from ..ops import conv as conv_engine
from ..ops.resize import downsample2x, maxpool2x2


class _RoiAlign(torch.autograd.Function):
    """torchvision.ops.roi_align(x, rois, output_size, spatial_scale=1, sampling_ratio=-1, aligned=False) (bbox_cx_loss.py:47-57) on
    `spi_roi_align` / `spi_roi_align_backward`."""

    @staticmethod
    def forward(ctx, x, rois, size):
        n, c, h, w = x.shape
        rois = rois.contiguous().float()
        out = torch.empty(rois.shape[0], c, size, size, device=x.device, dtype=torch.float32)
        _lib.check(_lib.load().spi_roi_align(_lib.ptr(x), _lib.ptr(rois), _lib.ptr(out), n, c, h, w, _lib.strides4(x), rois.shape[0], size, _lib.stream()))
        ctx.save_for_backward(rois)
        ctx.meta = (x.shape, size)
        return out

    @staticmethod
    def backward(ctx, g):
        (rois,) = ctx.saved_tensors
        shape, size = ctx.meta
        gin = torch.zeros(shape, device=g.device, dtype=torch.float32)
        g = g.contiguous()
        _lib.check(_lib.load().spi_roi_align_backward(_lib.ptr(g), _lib.ptr(rois), _lib.ptr(gin), shape[0], shape[1], shape[2], shape[3],
                                                      _lib.strides4(gin), rois.shape[0], size, _lib.stream()))
        return gin, None, None


def roi_align(x, boxes, output_size):
    if not x.is_cuda or x.dtype != torch.float32:
        raise RuntimeError('spi_b200 roi_align: x must be a float32 CUDA tensor (no CPU path in this build)')
    return _RoiAlign.apply(x, boxes, int(output_size))


class _CXRows(torch.autograd.Function):
    """compute_relative_distance + compute_cx + max over dim 1 (bbox_cx_loss.py:116-131,176) of a [B, M, N] cosine-similarity matrix in one
    pass forward and one backward (`spi_cx_rows_forward/backward`): returns colmax [B, N] = max_i cx[b, i, j]."""

    @staticmethod
    def forward(ctx, sim, band_width):
        b, m, n = sim.shape
        sim = sim.contiguous()
        stats = torch.empty(b, m, 2, device=sim.device, dtype=torch.float32)
        colmax = torch.empty(b, n, device=sim.device, dtype=torch.float32)
        _lib.check(_lib.load().spi_cx_rows_forward(_lib.ptr(sim), b, m, n, float(band_width), _lib.ptr(stats), _lib.ptr(colmax), _lib.stream()))
        ctx.save_for_backward(sim, stats, colmax)
        ctx.band_width = float(band_width)
        return colmax

    @staticmethod
    def backward(ctx, g):
        sim, stats, colmax = ctx.saved_tensors
        b, m, n = sim.shape
        dsim = torch.empty_like(sim)
        g = g.contiguous()
        _lib.check(_lib.load().spi_cx_rows_backward(_lib.ptr(sim), b, m, n, ctx.band_width, _lib.ptr(stats), _lib.ptr(colmax), _lib.ptr(g),
                                                    _lib.ptr(dsim), _lib.stream()))
        return dsim, None


def cosine_similarity_matrix(x, y):
    """The similarity half of compute_cosine_distance (bbox_cx_loss.py:93-113): centred on y's channel means, unit-normalised, [N, HW_x, HW_y]."""
    y_mu = y.mean(dim=(0, 2, 3), keepdim=True)
    x_n = F.normalize(x - y_mu, p=2, dim=1)
    y_n = F.normalize(y - y_mu, p=2, dim=1)
    N, C = x.shape[:2]
    return torch.bmm(x_n.reshape(N, C, -1).transpose(1, 2), y_n.reshape(N, C, -1))


def get_landmark_bbox(lm, scale=1):
    """bbox_cx_loss.py:20-37.  `p` is not reset after the eyes, so the (unused) nose box also gets 15."""
    p = 8
    bbox = []
    for _i, (a, b) in enumerate(((48, 68), (36, 42), (42, 48), (27, 36))):
        box_lm = lm[:, a:b]
        ly, ry = torch.min(box_lm[:, :, 0], dim=1)[0], torch.max(box_lm[:, :, 0], dim=1)[0]
        lx, rx = torch.min(box_lm[:, :, 1], dim=1)[0], torch.max(box_lm[:, :, 1], dim=1)[0]
        lx, rx, ly, ry = (lx * scale).long(), (rx * scale).long(), (ly * scale).long(), (ry * scale).long()
        if _i == 1 or _i == 2:
            p = 15
        bbox.append(torch.stack([ly - p, lx - p, ry + p, rx + p], dim=1))
    return bbox


def get_bbox(image, fake_image, lm):
    """bbox_cx_loss.py:41-59."""
    assert image.shape[-1] == 256
    bbox = get_landmark_bbox(lm)
    idx = torch.arange(image.shape[0], device=image.device).unsqueeze(1)
    out = []
    for k in range(3):
        rois = torch.cat([idx, bbox[k].to(image.device)], dim=1).float()
        out.append(roi_align(image, boxes=rois, output_size=80))
    for k in range(3):
        rois = torch.cat([idx, bbox[k].to(image.device)], dim=1).float()
        out.append(roi_align(fake_image, boxes=rois, output_size=80))
    return tuple(out)      # gt_mouth, gt_l_eye, gt_r_eye, fake_mouth, fake_l_eye, fake_r_eye


class VGG19(torch.nn.Module):
    """torchvision `vgg19().features[:6]` = conv(3,64) relu conv(64,64) relu maxpool conv(64,128)  (bbox_cx_loss.py:76-91)."""

    def __init__(self, requires_grad=False):
        super().__init__()
        self.slice1 = torch.nn.Sequential()
        mods = [torch.nn.Conv2d(3, 64, 3, padding=1), torch.nn.ReLU(inplace=True), torch.nn.Conv2d(64, 64, 3, padding=1),
                torch.nn.ReLU(inplace=True), torch.nn.MaxPool2d(2, 2), torch.nn.Conv2d(64, 128, 3, padding=1)]
        for i, m in enumerate(mods):
            self.slice1.add_module(str(i), m)
        if not requires_grad:
            for p in self.parameters():
                p.requires_grad = False

    def forward(self, X):
        X = X.contiguous(memory_format=torch.channels_last)
        mods = list(self.slice1)
        k = 0
        while k < len(mods):
            m = mods[k]
            if isinstance(m, torch.nn.Conv2d):
                relu = k + 1 < len(mods) and isinstance(mods[k + 1], torch.nn.ReLU)
                X = conv_engine.vgg_conv(X, m, act=('relu' if relu else 'linear'))
                k += 2 if relu else 1
            else:
                X = maxpool2x2(X) if isinstance(m, torch.nn.MaxPool2d) else m(X)
                k += 1
        return X


def compute_cosine_distance(x, y):
    y_mu = y.mean(dim=(0, 2, 3), keepdim=True)
    x_n = F.normalize(x - y_mu, p=2, dim=1)
    y_n = F.normalize(y - y_mu, p=2, dim=1)
    N, C = x.shape[:2]
    return 1 - torch.bmm(x_n.reshape(N, C, -1).transpose(1, 2), y_n.reshape(N, C, -1))


def compute_relative_distance(dist_raw):
    dist_min, _ = torch.min(dist_raw, dim=2, keepdim=True)
    return torch.clamp(dist_raw / (dist_min + 1e-5), max=10., min=-10)


def compute_cx(dist_tilde, band_width):
    w = torch.exp((1 - dist_tilde) / band_width)
    return w / torch.sum(w, dim=2, keepdim=True)


def flip_landmark(lm):
    """bbox_cx_loss.py:133-137: returns its input unchanged (the flipped copy is discarded)."""
    return lm


class BoxCXLoss(torch.nn.Module):
    def __init__(self, band_width: float = 0.5):
        super().__init__()
        self.band_width = band_width
        self.vgg_model = VGG19()
        self.register_buffer('vgg_mean', torch.tensor([[[0.485]], [[0.456]], [[0.406]]]))
        self.register_buffer('vgg_std', torch.tensor([[[0.229]], [[0.224]], [[0.225]]]))

    @staticmethod
    def _to256(t):
        if t.shape[-1] == 512 and t.shape[-2] == 512:
            return downsample2x(t)
        if t.shape[-1] > 256:
            return F.interpolate(t, (256, 256), mode='bilinear', align_corners=False)
        return t

    def forward(self, x, y, lm):
        if not x.is_cuda:
            raise RuntimeError('spi_b200 BoxCXLoss: tensors must reside on a CUDA device (no CPU path in this build)')
        x, y = self._to256(x), self._to256(y)
        x = x.sub(self.vgg_mean).div(self.vgg_std)       # ImageNet z-score on [-1,1] data, as the reference does (:165-166)
        y = y.sub(self.vgg_mean).div(self.vgg_std)
        gt_mouth, gt_l_eye, gt_r_eye, fake_mouth, fake_l_eye, fake_r_eye = get_bbox(x, y, lm)
        loss = 0
        for _x, _y in ((gt_mouth, fake_mouth), (gt_l_eye, fake_l_eye), (gt_r_eye, fake_r_eye)):
            fx, fy = self.vgg_model(_x), self.vgg_model(_y)
            if fx.shape[2] * fx.shape[3] <= 2048:
                # 1 - S, relative distance, exp, row normalisation and the column maximum in one pass over the [B,1600,1600] matrix
                cx = _CXRows.apply(cosine_similarity_matrix(fx, fy), self.band_width).mean(dim=1)
            else:
                cx = compute_cx(compute_relative_distance(compute_cosine_distance(fx, fy)), self.band_width)
                cx = torch.mean(torch.max(cx, dim=1)[0], dim=1)
            loss = loss + torch.mean(-torch.log(cx + 1e-5))
        return loss * 0.1
