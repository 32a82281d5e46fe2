import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')
    config.addinivalue_line('markers', 'slow: minutes of CPU oracle time; enabled with SPI_SLOW=1')


def pytest_collection_modifyitems(config, items):
    if os.environ.get('SPI_SLOW') == '1':
        return
    skip = pytest.mark.skip(reason='slow CPU oracle case (set SPI_SLOW=1)')
    for item in items:
        if 'slow' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def golden():
    def load(name):
        return dict(np.load(os.path.join(GOLDEN, name + '.npz')))
    return load


@pytest.fixture(scope='session')
def lib():
    """The built C-ABI library (GPU tests must go through it; they fail loudly if it is missing)."""
    from spi_b200 import _lib
    return _lib.load()


def rel_l2(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.fixture(scope='session')
def gen_sd():
    from oracle import weights
    return weights.generator_state_dict(0)


@pytest.fixture(scope='session')
def product_G(gen_sd):
    """spi_b200 TriPlaneGenerator with the seeded weights, on cuda:0, configured as load_eg3d does (load_utils.py:25-32)."""
    from oracle import ref_shim
    from spi_b200.training.triplane import TriPlaneGenerator
    G = TriPlaneGenerator(**ref_shim.FFHQ512_KWARGS).eval().requires_grad_(False)
    G.load_state_dict(gen_sd, strict=True)
    G.neural_rendering_resolution = 128
    return G.to('cuda')
