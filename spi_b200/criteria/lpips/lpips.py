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

    def forward(self, x, y, conf_sigma=None, mask=None):
        assert conf_sigma is None and mask is None, 'conf_sigma / mask variants are not used by SPI and not built'
        assert self.num_scales == 1, 'multi-scale LPIPS is not used by SPI and not built'
        n_samples = x.shape[0]
        resize = x.shape[-1] > 256                     # lpips.py:37-39: both images follow x's size test
        feat_x = self.net(self._resize(x) if resize else x)
        feat_y = self._target_feats(y, resize)
        diff = [(fx - fy) ** 2 for fx, fy in zip(feat_x, feat_y)]
        res = [l(d).mean((2, 3), True) for d, l in zip(diff, self.lin)]
        return torch.sum(torch.cat(res, 0)) / n_samples
