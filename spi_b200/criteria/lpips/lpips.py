"""LPIPS v0.1 on VGG16 (drop-in for spi/criteria/lpips/lpips.py:10-71).

Differences in execution only: the 512->256 bilinear resize is the exact 2x2 mean (`spi_downsample2x`), and when `y`
has been registered as a constant target (`register_target`) its five feature taps are cached instead of
re-running VGG16 on it every call (the reference recomputes them: +40 GF per call).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ...ops.resize import downsample2x
from .networks import LinLayers, get_network
from .utils import get_state_dict


class LPIPS(nn.Module):
    def __init__(self, net_type: str = 'alex', version: str = '0.1', num_scales=1, lin_state_dict=None):
        assert version in ['0.1'], 'v0.1 is only supported now'
        super().__init__()
        self.net = get_network(net_type)
        self.lin = LinLayers(self.net.n_channels_list)
        if lin_state_dict is None:
            try:
                lin_state_dict = get_state_dict(net_type, version)
            except FileNotFoundError:
                lin_state_dict = None           # weights are loaded later via load_state_dict / load_weights
        if lin_state_dict is not None:
            self.lin.load_state_dict(lin_state_dict)
        self.num_scales = num_scales
        self._cache = {}                 # id(y) -> (y, version, resize, feats); holds y alive so its id/ptr cannot be recycled
        self._retired = []
        self._pairs = {}                 # ids of cached target taps -> their concatenation over the batch (weighted_pairs)

    def load_weights(self, vgg_features_state_dict, lin_weights):
        """vgg_features_state_dict: torchvision `vgg16().features` names ('0.weight', ...); lin_weights: 5 x [1,C,1,1]."""
        self.net.layers.load_state_dict(vgg_features_state_dict, strict=False)
        self.lin.load_state_dict({f'{i}.1.weight': w for i, w in enumerate(lin_weights)})
        self._cache.clear()
        return self

    @staticmethod
    def _resize(t):
        if t.shape[-1] == 512 and t.shape[-2] == 512:
            return downsample2x(t)
        return F.interpolate(t, size=(256, 256), mode='bilinear', align_corners=False)

    def register_target(self, y):
        """Declare `y` a constant target (the image being inverted): its five feature taps are computed once and reused by
        every later forward(x, y) with this very tensor object.  Unregistered `y` (e.g. warped images) are never cached."""
        hit = self._cache.get(id(y))
        if hit is None or hit[0] is not y:        # keep an existing entry: captured CUDA graphs hold pointers to its tensors
            self._cache[id(y)] = (y, None, None, None)
        return y

    def release_targets(self):
        """Drop every registered target with its cached taps (and the retired ones).  Call only when no captured graph that
        reads them will be replayed again -- the coaches do so when they start a new image and drop that image's graphs."""
        self._cache.clear()
        self._retired.clear()
        self._pairs.clear()

    def _target_feats(self, y, resize):
        if y.requires_grad:
            return self.net(self._resize(y) if resize else y)
        hit = self._cache.get(id(y))
        if hit is None or hit[0] is not y:
            with torch.no_grad():
                return self.net(self._resize(y) if resize else y)
        if hit[3] is None or hit[1] != y._version or hit[2] != resize:
            with torch.no_grad():
                feats = self.net(self._resize(y) if resize else y)
            if hit[3] is not None:
                self._retired.append(hit[3])      # never free tensors a captured graph may still read
            self._cache[id(y)] = (y, y._version, resize, feats)
            return feats
        return hit[3]

    def weighted_pairs(self, x, targets, weights):
        """sum_n weights[n] * LPIPS(x[n:n+1], targets[n]) for constant targets (one per sample of x) and a device vector `weights` [N]:
        the expression of the mirror projector (mirror_projector.py:100-104: lpips(img, target) + w_m * lpips(img_m, target_m)) with ONE pass
        of the VGG trunk over the batch instead of one per pair -- the trunk's small layers cost the same for 1 and 2 images."""
        assert len(targets) == x.shape[0] and weights.shape == (x.shape[0],) and self.num_scales == 1
        resize = x.shape[-1] > 256
        per_target = [self._target_feats(t, resize) for t in targets]
        if not all(_fusable(f, t) for f, t in zip(per_target, targets)):
            return sum(self.forward(x[i:i + 1], t) * weights[i] for i, t in enumerate(targets))
        registered = all(self._cache.get(id(t), (None,))[0] is t for t in targets)
        key = tuple(id(f[0]) for f in per_target)
        hit = self._pairs.get(key) if registered else None
        if hit is None:               # concatenated taps of this set of targets (registered targets: kept, captured graphs read them)
            hit = ([torch.cat([f[k] for f in per_target], 0) for k in range(len(per_target[0]))], per_target)
            if registered:
                if len(self._pairs) > 8:
                    self._retired.extend(self._pairs.values())
                    self._pairs.clear()
                self._pairs[key] = hit
        feat_x = self.net(self._resize(x) if resize else x, normalize=False)
        lins = [l[1].weight.reshape(-1) for l in self.lin]
        return _LpipsTail.apply(len(feat_x), *feat_x, *hit[0], *lins, weights.float().contiguous())

    def forward(self, x, y, conf_sigma=None, mask=None):
        assert conf_sigma is None and mask is None, 'conf_sigma / mask variants are not used by SPI and not built'
        assert self.num_scales == 1, 'multi-scale LPIPS is not used by SPI and not built'
        n_samples = x.shape[0]
        resize = x.shape[-1] > 256                     # lpips.py:37-39: both images follow x's size test
        feat_y = self._target_feats(y, resize)
        xin = self._resize(x) if resize else x
        if _fusable(feat_y, y):
            # normalise + difference + lin + spatial mean of all five taps in one pass each (spi_lpips_tap_forward/backward)
            feat_x = self.net(xin, normalize=False)
            lins = [l[1].weight.reshape(-1) for l in self.lin]
            return _LpipsTail.apply(len(feat_x), *feat_x, *feat_y, *lins) / n_samples
        feat_x = self.net(xin)
        diff = [(fx - fy) ** 2 for fx, fy in zip(feat_x, feat_y)]
        res = [l(d).mean((2, 3), True) for d, l in zip(diff, self.lin)]
        return torch.sum(torch.cat(res, 0)) / n_samples


def _fusable(feat_y, y):
    ys = y if isinstance(y, (tuple, list)) else (y,)
    return (not any(t.requires_grad for t in ys)) and all((not f.requires_grad) and f.dtype == torch.float32 and f.shape[1] % 4 == 0 and f.shape[1] <= 512
                                         for f in feat_y)


class _LpipsTail(torch.autograd.Function):
    """sum over taps of sum_n mean_hw sum_c lin_c (x_c/(|x|+1e-10) - yn_c)^2 (lpips.py:50-71); gradients flow to the x taps only."""

    @staticmethod
    def forward(ctx, k, *args):
        from ... import _lib
        xs = [a.contiguous(memory_format=torch.channels_last) for a in args[:k]]
        ys = [a.contiguous(memory_format=torch.channels_last) for a in args[k:2 * k]]
        lins = [a.contiguous() for a in args[2 * k:3 * k]]
        sw = args[3 * k] if len(args) > 3 * k else None            # per-sample weights of the sum over the batch
        out = torch.zeros((), device=xs[0].device)
        lib = _lib.load()
        for fx, fy, w in zip(xs, ys, lins):
            n, c, h, wd = fx.shape
            assert fy.shape[1:] == fx.shape[1:] and fy.shape[0] in (1, n)
            _lib.check(lib.spi_lpips_tap_forward(_lib.ptr(fx), _lib.ptr(fy), _lib.ptr(w), n, h * wd, c, fy.shape[0], _lib.ptr(out), _lib.ptr(sw), _lib.stream()))
        ctx.k, ctx.has_sw = k, sw is not None
        ctx.save_for_backward(*xs, *ys, *lins, *([sw] if sw is not None else []))
        return out

    @staticmethod
    def backward(ctx, gout):
        from ... import _lib
        k = ctx.k
        saved = ctx.saved_tensors
        xs, ys, lins = saved[:k], saved[k:2 * k], saved[2 * k:3 * k]
        sw = saved[3 * k] if ctx.has_sw else None
        gout = gout.contiguous().float()
        lib = _lib.load()
        grads = []
        for i, (fx, fy, w) in enumerate(zip(xs, ys, lins)):
            if not ctx.needs_input_grad[1 + i]:
                grads.append(None)
                continue
            n, c, h, wd = fx.shape
            dx = torch.empty_like(fx)
            _lib.check(lib.spi_lpips_tap_backward(_lib.ptr(fx), _lib.ptr(fy), _lib.ptr(w), n, h * wd, c, fy.shape[0], _lib.ptr(gout), _lib.ptr(dx),
                                                  _lib.ptr(sw), _lib.stream()))
            grads.append(dx)
        return (None, *grads, *([None] * (2 * k + (1 if ctx.has_sw else 0))))
