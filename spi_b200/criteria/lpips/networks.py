"""LPIPS backbones (drop-in for spi/criteria/lpips/networks.py).  Only the VGG16 variant is on the inversion path
(`LPIPS(net_type='vgg')`, base_coach.py:48); the module tree (`layers.<i>`, `mean`, `std`) matches torchvision's
`vgg16().features` so pretrained state dicts load by name."""
from typing import Sequence

import torch
import torch.nn as nn

from ...ops import conv as conv_engine
from ...ops.resize import maxpool2x2
from ...torch_utils.ops import bias_act
from .utils import normalize_activation

VGG16_FEATURES = (64, 64, 'M', 128, 128, 'M', 256, 256, 256, 'M', 512, 512, 512, 'M', 512, 512, 512, 'M')


def vgg_features(cfg):
    layers, cin = [], 3
    for v in cfg:
        if v == 'M':
            layers.append(nn.MaxPool2d(kernel_size=2, stride=2))
        else:
            layers += [nn.Conv2d(cin, v, kernel_size=3, padding=1), nn.ReLU(inplace=True)]
            cin = v
    return nn.Sequential(*layers)


def get_network(net_type: str):
    if net_type == 'vgg':
        return VGG16()
    raise NotImplementedError("spi_b200 builds the 'vgg' LPIPS backbone only (the one SPI uses)")


class LinLayers(nn.ModuleList):
    def __init__(self, n_channels_list: Sequence[int]):
        super().__init__([nn.Sequential(nn.Identity(), nn.Conv2d(nc, 1, 1, 1, 0, bias=False)) for nc in n_channels_list])
        for p in self.parameters():
            p.requires_grad = False


class BaseNet(nn.Module):
    def __init__(self):
        super().__init__()
        self.register_buffer('mean', torch.Tensor([-.030, -.088, -.188])[None, :, None, None])
        self.register_buffer('std', torch.Tensor([.458, .448, .450])[None, :, None, None])

    def set_requires_grad(self, state: bool):
        for p in list(self.parameters()) + list(self.buffers()):
            p.requires_grad = state

    def z_score(self, x):
        return (x - self.mean) / self.std

    def _weights_channels_last(self):
        """The conv engine consumes channels-last weights; the (frozen) VGG weights are converted once instead of on every call
        (14.7 M parameters = a 59 MB copy per LPIPS evaluation otherwise).  load_state_dict() copies into the converted tensors."""
        for layer in self.layers:
            if isinstance(layer, nn.Conv2d) and not layer.weight.is_contiguous(memory_format=torch.channels_last):
                layer.weight.data = layer.weight.data.contiguous(memory_format=torch.channels_last)

    def forward(self, x, normalize=True):
        """Feature taps; `normalize=False` returns them raw (the fused LPIPS tail normalises on the fly)."""
        if not x.is_cuda:
            raise RuntimeError('spi_b200 LPIPS: tensors must reside on a CUDA device (no CPU path in this build)')
        self._weights_channels_last()
        x = self.z_score(x).contiguous(memory_format=torch.channels_last)
        output = []
        fused_relu = False
        for i, layer in enumerate(self.layers, 1):
            if isinstance(layer, nn.Conv2d):       # conv -> (+bias, ReLU) in one epilogue pass
                x = conv_engine.vgg_conv(x, layer, act='relu')
                fused_relu = True
            elif isinstance(layer, nn.ReLU) and fused_relu:
                fused_relu = False
            else:
                x = maxpool2x2(x) if isinstance(layer, nn.MaxPool2d) else layer(x)
            if i in self.target_layers:
                output.append(normalize_activation(x) if normalize else x)
            if len(output) == len(self.target_layers):
                break
        return output


class VGG16(BaseNet):
    def __init__(self):
        super().__init__()
        self.layers = vgg_features(VGG16_FEATURES)          # torchvision `vgg16().features` naming
        self.target_layers = [4, 9, 16, 23, 30]
        self.n_channels_list = [64, 128, 256, 512, 512]
        self.set_requires_grad(False)
