"""Restatement of NVIDIA's TorchScript `vgg16.pt` feature extractor as called by the `sg` projector
(`vgg16(images, resize_images=False, return_lpips=True)`, spi/training/projectors/w_projector.py:51,86).

The artefact itself is third-party, un-versioned and absent (PARITY UNPINNED for this function).  It is
restated from the LPIPS modules: images in [0,255] -> [-1,1] -> z-score -> VGG16 taps -> unit-normalise -> scale by
sqrt(lin weight) / sqrt(H*W) -> flatten + concat, so that the squared distance of two feature vectors equals LPIPS.
"""
import torch

from .networks import LinLayers, VGG16


class VGG16LPIPSFeatures(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.net = VGG16()
        self.lin = LinLayers(self.net.n_channels_list)

    def load_weights(self, vgg_features_state_dict, lin_weights):
        self.net.layers.load_state_dict(vgg_features_state_dict, strict=False)
        self.lin.load_state_dict({f'{i}.1.weight': w for i, w in enumerate(lin_weights)})
        return self

    def forward(self, images, resize_images=False, return_lpips=True):
        assert not resize_images and return_lpips, 'only the call signature used by w_projector.py is built'
        x = images / 127.5 - 1
        feats = []
        for f, l in zip(self.net(x), self.lin):
            h, w = f.shape[2:]
            feats.append((f * torch.sqrt(l[1].weight) / (h * w) ** 0.5).reshape(f.shape[0], -1))
        return torch.cat(feats, 1)
