"""spi/criteria/lpips/utils.py:6-29."""
from collections import OrderedDict

import torch


def normalize_activation(x, eps=1e-10):
    norm_factor = torch.sqrt(torch.sum(x ** 2, dim=1, keepdim=True))
    return x / (norm_factor + eps)


def get_state_dict(net_type='alex', version='0.1', path=None):
    """The reference downloads richzhang's `v0.1/vgg.pth` (utils.py:13-20).  There is no network here: the file is read
    from `path` (or $SPI_LPIPS_LIN) when given; keys are renamed exactly as the reference does."""
    import os
    path = path or os.environ.get('SPI_LPIPS_LIN')
    if path is None or not os.path.isfile(path):
        raise FileNotFoundError('LPIPS lin weights: pass `path=` or set SPI_LPIPS_LIN (the reference downloads them; no network here)')
    old = torch.load(path, map_location='cpu')
    new = OrderedDict()
    for key, val in old.items():
        new[key.replace('lin', '').replace('model.', '')] = val
    return new
