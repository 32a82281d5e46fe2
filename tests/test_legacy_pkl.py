"""Reading the reference's source-carrying network pickle (eg3d/legacy.py:23, eg3d/torch_utils/persistence.py:120-205) with
spi_b200.legacy -- without executing the embedded source.  The fixture tests/golden/network_small.pkl.gz was pickled by the
REFERENCE's own classes (oracle/make_golden_pkl.py); network_small.json holds the reference's per-tensor checksums and
network_small.npz its render of a seeded latent."""
import gzip
import io
import json
import os
import pickle

import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_l2


def _open():
    return gzip.open(os.path.join(GOLDEN, 'network_small.pkl.gz'), 'rb')


@pytest.fixture(scope='module')
def loaded():
    from spi_b200 import legacy
    with _open() as f:
        return legacy.load_network_pkl(f)


def test_pickle_loads_into_the_product_generator(loaded):
    from spi_b200.legacy import PersistentStub
    from spi_b200.training.triplane import TriPlaneGenerator
    meta = json.load(open(os.path.join(GOLDEN, 'network_small.json')))
    G = loaded['G_ema']
    assert isinstance(G, TriPlaneGenerator) and isinstance(loaded['G'], TriPlaneGenerator)
    assert isinstance(loaded['D'], PersistentStub) and loaded['D'].class_name == 'FullyConnectedLayer'      # off-path classes stay inert
    assert loaded['augment_pipe'] is None and loaded['training_set_kwargs']['resolution'] == 512
    assert not G.training and not any(p.requires_grad for p in G.parameters())
    kw = G.init_kwargs
    assert kw['channel_base'] == meta['init_kwargs']['channel_base'] and kw['rendering_kwargs']['depth_resolution'] == 24
    assert G.neural_rendering_resolution == 128 and G.backbone.mapping.num_ws == meta['num_ws']
    sd = G.state_dict()
    assert set(sd) == set(meta['tensors']) and len(sd) == 176                   # 132 parameters + 44 buffers, reference names
    for k, m in meta['tensors'].items():
        v = sd[k].double()
        assert list(v.shape) == m['shape'], k
        assert abs(float(v.sum()) - m['sum']) <= 1e-9 * max(1.0, abs(m['sum'])), k
        assert abs(float(v.square().sum()) - m['sumsq']) <= 1e-9 * max(1.0, m['sumsq']), k


def test_load_eg3d_accepts_a_pkl_path(tmp_path):
    from spi_b200.utils import load_utils
    p = tmp_path / 'net.pkl'
    with _open() as f:
        p.write_bytes(f.read())
    G = load_utils.load_eg3d(device='cpu', network_pkl=str(p))
    assert G.neural_rendering_resolution == 128 and G.w_dim == 512 and G.z_dim == 64
    G2 = load_utils.load_eg3d(device='cpu', network_pkl=str(p))                 # restart_training(): a fresh copy from the cached template
    # load_utils.py:27-28: the pickled rendering_kwargs ATTRIBUTE wins over the constructor's copy
    kw, sd, rk = load_utils._template[str(p)]
    rk2 = dict(rk, ray_start=2.0, box_warp=1.5)
    load_utils._template['edited'] = (kw, sd, rk2)
    G3 = load_utils.load_eg3d(device='cpu', network_pkl='edited')
    assert G3.rendering_kwargs['ray_start'] == 2.0 and G3.rendering_kwargs['box_warp'] == 1.5 and G3.init_kwargs['rendering_kwargs']['ray_start'] != 2.0
    del load_utils._template['edited']
    assert G2 is not G and all(torch.equal(a, b) for a, b in zip(G.state_dict().values(), G2.state_dict().values()))


def test_embedded_source_is_never_executed_and_foreign_classes_are_refused():
    from spi_b200 import legacy

    class Boom:
        def __reduce__(self):
            return (os.system, ('echo pwned > /tmp/spi_b200_pwned',))
    with pytest.raises(pickle.UnpicklingError):
        legacy.load_network_pkl(io.BytesIO(pickle.dumps(dict(G_ema=Boom()))))
    assert not os.path.exists('/tmp/spi_b200_pwned')
    # a persistent object whose module_src would raise if it were exec'd
    meta = dict(type='class', version=6, module_src='raise SystemExit("module_src was executed")', class_name='Whatever', state={})
    payload = (b'\x80\x04' + b'ctorch_utils.persistence\n_reconstruct_persistent_obj\n' + pickle.dumps((meta,))[2:-1] + b'R.')
    obj = legacy._SafeUnpickler(io.BytesIO(payload)).load()
    assert isinstance(obj, legacy.PersistentStub) and obj.class_name == 'Whatever'
    # gadgets that live UNDER the torch / numpy roots are refused too (explicit allowlist, not a root-package test)
    for payload in (b"ctorch.utils.collect_env\nrun\n(S'echo pwned > /tmp/spi_b200_pwned'\ntR.",
                    b"ctorch.serialization\nload\n(S'/etc/hostname'\ntR.",
                    b"cnumpy\nload\n(S'/etc/hostname'\ntR.",
                    b"ctorch.hub\nload\n(S'x'\nS'y'\ntR."):
        with pytest.raises(pickle.UnpicklingError):
            legacy._SafeUnpickler(io.BytesIO(payload)).load()
    assert not os.path.exists('/tmp/spi_b200_pwned')


@pytest.mark.gpu
def test_loaded_generator_renders_like_the_reference(loaded):
    from oracle import generator as OG
    g = dict(np.load(os.path.join(GOLDEN, 'network_small.npz')))
    G = loaded['G_ema'].to('cuda')
    rk = {**OG.RENDERING_DEFAULTS, **dict(G.rendering_kwargs)}
    jit, u = OG.make_render_noise(1, 128 * 128, rk, seed=7)
    G.renderer.inject_noise(jit.cuda(), u.cuda())
    out = G.synthesis(torch.from_numpy(g['ws']).cuda(), torch.from_numpy(g['c']).cuda(), noise_mode='const')
    assert rel_l2(out['image_depth'], g['image_depth']) < 1e-4
    assert rel_l2(out['image_raw'], g['image_raw']) < 1e-3
    assert rel_l2(torch.nn.functional.avg_pool2d(out['image'], 8), g['image_64']) < 1e-3
