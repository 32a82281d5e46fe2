"""ArcFace IR-SE50 backbone (drop-in for spi/criteria/id_loss/model_irse.py:10-49 + helpers.py:23-121): same module tree /
state-dict names (`input_layer`, `body.<i>.{shortcut_layer,res_layer}`, `output_layer`).  Forward-only in SPI (metrics,
`metric_utils.py:11-17`); the convolutions go through the conv engine in channels-last layout."""
from collections import namedtuple

import torch
from torch.nn import (AdaptiveAvgPool2d, BatchNorm1d, BatchNorm2d, Conv2d, Dropout, Linear, MaxPool2d, Module, PReLU, ReLU,
                      Sequential, Sigmoid)


class Flatten(Module):
    def forward(self, x):
        return x.reshape(x.size(0), -1)


def l2_norm(x, axis=1):
    return torch.div(x, torch.norm(x, 2, axis, True))


class Bottleneck(namedtuple('Block', ['in_channel', 'depth', 'stride'])):
    """A ResNet unit description."""


def get_block(in_channel, depth, num_units, stride=2):
    return [Bottleneck(in_channel, depth, stride)] + [Bottleneck(depth, depth, 1) for _ in range(num_units - 1)]


def get_blocks(num_layers):
    table = {50: (3, 4, 14, 3), 100: (3, 13, 30, 3), 152: (3, 8, 36, 3)}
    if num_layers not in table:
        raise ValueError('Invalid number of layers: {}. Must be one of [50, 100, 152]'.format(num_layers))
    n = table[num_layers]
    return [get_block(64, 64, n[0]), get_block(64, 128, n[1]), get_block(128, 256, n[2]), get_block(256, 512, n[3])]


class SEModule(Module):
    def __init__(self, channels, reduction):
        super().__init__()
        self.avg_pool = AdaptiveAvgPool2d(1)
        self.fc1 = Conv2d(channels, channels // reduction, kernel_size=1, padding=0, bias=False)
        self.relu = ReLU(inplace=True)
        self.fc2 = Conv2d(channels // reduction, channels, kernel_size=1, padding=0, bias=False)
        self.sigmoid = Sigmoid()

    def forward(self, x):
        return x * self.sigmoid(self.fc2(self.relu(self.fc1(self.avg_pool(x)))))


class bottleneck_IR(Module):
    se = False

    def __init__(self, in_channel, depth, stride):
        super().__init__()
        if in_channel == depth:
            self.shortcut_layer = MaxPool2d(1, stride)
        else:
            self.shortcut_layer = Sequential(Conv2d(in_channel, depth, (1, 1), stride, bias=False), BatchNorm2d(depth))
        layers = [BatchNorm2d(in_channel), Conv2d(in_channel, depth, (3, 3), (1, 1), 1, bias=False), PReLU(depth),
                  Conv2d(depth, depth, (3, 3), stride, 1, bias=False), BatchNorm2d(depth)]
        if self.se:
            layers.append(SEModule(depth, 16))
        self.res_layer = Sequential(*layers)

    def forward(self, x):
        return self.res_layer(x) + self.shortcut_layer(x)


class bottleneck_IR_SE(bottleneck_IR):
    se = True


class Backbone(Module):
    def __init__(self, input_size, num_layers, mode='ir', drop_ratio=0.4, affine=True):
        super().__init__()
        assert input_size in [112, 224] and num_layers in [50, 100, 152] and mode in ['ir', 'ir_se']
        unit = bottleneck_IR if mode == 'ir' else bottleneck_IR_SE
        self.input_layer = Sequential(Conv2d(3, 64, (3, 3), 1, 1, bias=False), BatchNorm2d(64), PReLU(64))
        side = 7 if input_size == 112 else 14
        self.output_layer = Sequential(BatchNorm2d(512), Dropout(drop_ratio), Flatten(), Linear(512 * side * side, 512),
                                       BatchNorm1d(512, affine=affine))
        self.body = Sequential(*[unit(b.in_channel, b.depth, b.stride) for block in get_blocks(num_layers) for b in block])

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError('spi_b200 IR-SE50: tensors must reside on a CUDA device (no CPU path in this build)')
        x = x.contiguous(memory_format=torch.channels_last)
        x = self.body(self.input_layer(x))
        # Flatten follows the reference's NCHW element order (Linear weights are laid out for it)
        x = self.output_layer[0](x).contiguous(memory_format=torch.contiguous_format)
        for m in list(self.output_layer)[1:]:
            x = m(x)
        return l2_norm(x)
