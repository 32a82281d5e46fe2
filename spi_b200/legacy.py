"""Reading the reference's source-carrying network pickles (drop-in for `legacy.load_network_pkl`, eg3d/legacy.py:23-62, as
called by spi/utils/load_utils.py:15-33) WITHOUT executing the Python source embedded in them.

The reference pickles every `@persistence.persistent_class` object as `_reconstruct_persistent_obj(meta)` with
`meta = dict(type='class', version, module_src, class_name, state)` (eg3d/torch_utils/persistence.py:120-128), where `state` is the
module's `__dict__` (`_parameters`, `_buffers`, `_modules`, `_init_args`, `_init_kwargs`, ...), and rebuilds the class by `exec`-ing
`module_src` (persistence.py:181-205, :218-230).  Here a restricted unpickler turns every persistent object into an inert
`PersistentStub(class_name, state)`; a `TriPlaneGenerator` stub is then rebuilt as THIS repo's generator from its constructor
arguments and its parameters / buffers collected by name (`require_all=True`, as load_utils.py:26).  Classes that are not on the
inversion path (the discriminator, the augmentation pipe) stay stubs.  Globals are resolved from an explicit allowlist of exact (module, name) pairs (tensor rebuild functions, storages, dtypes,
OrderedDict, a few parameter-free nn containers); nothing else is imported on behalf of the file and no code from it runs.
"""
import collections
import copy
import pickle

import numpy as np
import torch

from . import dnnlib

_SAFE_BUILTINS = {'set', 'frozenset', 'slice', 'complex', 'range', 'dict', 'list', 'tuple', 'int', 'float', 'bool', 'str', 'bytes', 'bytearray',
                  'object'}


# Exact (module, name) pairs a network pickle needs for tensors and containers.  Everything else under torch.* / numpy.* is
# refused: those packages are full of callables that run commands or unpickle again (torch.utils.collect_env.run, torch.load,
# torch.hub.*, numpy.load, numpy.testing.*), so "any global below the torch root" is not a safe rule.
_ALLOWED_GLOBALS = {
    ('collections', 'OrderedDict'),
    ('copyreg', '_reconstructor'),
    ('_codecs', 'encode'),
    ('numpy', 'dtype'), ('numpy', 'ndarray'),
    ('numpy.core.multiarray', '_reconstruct'), ('numpy.core.multiarray', 'scalar'),
    ('numpy._core.multiarray', '_reconstruct'), ('numpy._core.multiarray', 'scalar'),
    ('torch._utils', '_rebuild_tensor'), ('torch._utils', '_rebuild_tensor_v2'),
    ('torch._utils', '_rebuild_parameter'), ('torch._utils', '_rebuild_parameter_with_state'),
    ('torch.nn.parameter', 'Parameter'),
    ('torch.nn.modules.container', 'Sequential'), ('torch.nn.modules.container', 'ModuleList'), ('torch.nn.modules.container', 'ModuleDict'),
    ('torch.nn.modules.activation', 'Softplus'), ('torch.nn.modules.activation', 'LeakyReLU'), ('torch.nn.modules.activation', 'ReLU'),
    ('torch.nn.modules.linear', 'Identity'),
}
_TORCH_NAMES = {'Size', 'device', 'FloatStorage', 'HalfStorage', 'DoubleStorage', 'BFloat16Storage', 'LongStorage', 'IntStorage', 'ShortStorage',
                'CharStorage', 'ByteStorage', 'BoolStorage', 'float32', 'float16', 'float64', 'bfloat16', 'int64', 'int32', 'int16', 'int8',
                'uint8', 'bool'}


def _load_storage_from_bytes(b):
    """Stand-in for torch.storage._load_from_bytes (plain-pickled storages): the original calls torch.load(weights_only=False) on
    bytes taken from the file, i.e. a second, unrestricted unpickle; this one only accepts tensor data."""
    import io
    return torch.load(io.BytesIO(b), weights_only=True)


class PersistentStub:
    """A pickled persistent object: class name + the pickled `__dict__`; never instantiates the original class."""

    def __init__(self, class_name, state, version=None):
        self.class_name, self.state, self.version = class_name, state, version

    @property
    def init_args(self):
        return copy.deepcopy(self.state.get('_init_args', ()))

    @property
    def init_kwargs(self):
        return dnnlib.EasyDict(copy.deepcopy(self.state.get('_init_kwargs', {})))

    def __repr__(self):
        return f'<PersistentStub {self.class_name}>'


class ClassStub:
    """Stand-in for a NON-persistent class of the reference's own packages (`training.*`, `torch_utils.*`, `dnnlib.*`) that the file
    names by import path (e.g. training.volumetric_rendering.renderer.ImportanceRenderer, triplane.OSGDecoder): the pickle machinery
    creates it with `__new__` and hands its `__dict__` to `__setstate__`; the real class is never imported."""
    class_name = '?'

    def __init__(self, *args, **kwargs):
        self.state = {}

    def __setstate__(self, state):
        self.state = dict(state) if isinstance(state, dict) else {'_state': state}

    def __repr__(self):
        return f'<ClassStub {self.class_name}>'


_stub_classes = {}


def _class_stub(module, name):
    key = f'{module}.{name}'
    if key not in _stub_classes:
        _stub_classes[key] = type(name, (ClassStub,), {'class_name': name, '__module__': __name__})
    return _stub_classes[key]


def _reconstruct_stub(meta):
    if meta.get('type') != 'class':
        raise pickle.UnpicklingError(f"unsupported persistent object type {meta.get('type')!r}")
    return PersistentStub(meta['class_name'], dict(meta['state'] or {}), meta.get('version'))


class _SafeUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module == 'torch_utils.persistence' and name == '_reconstruct_persistent_obj':
            return _reconstruct_stub
        if module == 'dnnlib.util' and name == 'EasyDict':
            return dnnlib.EasyDict
        if module == 'dnnlib.tflib.network':
            raise pickle.UnpicklingError('legacy TensorFlow pickles are not supported (no EG3D checkpoint uses them)')
        if (module, name) == ('torch.storage', '_load_from_bytes'):
            return _load_storage_from_bytes
        if (module, name) in _ALLOWED_GLOBALS or (module == 'torch' and name in _TORCH_NAMES) or (module == 'builtins' and name in _SAFE_BUILTINS):
            return super().find_class(module, name)
        root = module.split('.')[0]
        if root in ('training', 'torch_utils', 'dnnlib'):
            return _class_stub(module, name)
        raise pickle.UnpicklingError(f'refusing to import {module}.{name} while reading a network pickle')


def _children(obj):
    """(parameters, buffers, non-persistent buffer names, sub-modules) of a stub or of a plain torch container."""
    d = obj.state if isinstance(obj, (PersistentStub, ClassStub)) else obj.__dict__
    return (d.get('_parameters') or {}, d.get('_buffers') or {}, d.get('_non_persistent_buffers_set') or set(), d.get('_modules') or {})


def stub_state_dict(stub, prefix=''):
    """The `state_dict()` the original module would have produced, collected from the pickled `__dict__` tree."""
    out = collections.OrderedDict()
    params, bufs, skip, mods = _children(stub)
    for k, v in params.items():
        if v is not None:
            out[prefix + k] = v.detach()
    for k, v in bufs.items():
        if v is not None and k not in skip:
            out[prefix + k] = v.detach()
    for k, m in mods.items():
        if m is not None:
            out.update(stub_state_dict(m, prefix + k + '.'))
    return out


def _build(stub):
    """Stub -> this repo's module for the classes on the inversion path; anything else stays a stub."""
    if not isinstance(stub, PersistentStub) or stub.class_name != 'TriPlaneGenerator':
        return stub
    from .training.triplane import TriPlaneGenerator
    G = TriPlaneGenerator(*stub.init_args, **stub.init_kwargs).eval().requires_grad_(False)
    sd = stub_state_dict(stub)
    missing, unexpected = G.load_state_dict(sd, strict=False)
    if missing or unexpected:            # misc.copy_params_and_buffers(..., require_all=True) semantics (load_utils.py:26)
        raise RuntimeError(f'network pickle does not match TriPlaneGenerator: missing {list(missing)[:5]}, unexpected {list(unexpected)[:5]}')
    for k in ('neural_rendering_resolution', 'rendering_kwargs'):
        if k in stub.state:               # attributes the reference copies over after re-instantiating (load_utils.py:27-28)
            setattr(G, k, copy.deepcopy(stub.state[k]))
    return G


def load_network_pkl(f, force_fp16=False):
    """eg3d/legacy.py:23.  Returns dict(G, D, G_ema, training_set_kwargs, augment_pipe); generators are spi_b200 modules."""
    if force_fp16:
        raise NotImplementedError('force_fp16: this build computes in fp32 (DESIGN.md §5)')
    data = _SafeUnpickler(f).load()
    if not isinstance(data, dict) or 'G_ema' not in data:
        raise RuntimeError('not an EG3D / StyleGAN network pickle (expected a dict with G, D, G_ema)')
    data = dict(data)
    data.setdefault('training_set_kwargs', None)
    data.setdefault('augment_pipe', None)
    for key in ('G', 'G_ema'):
        if key in data:
            data[key] = _build(data[key])
    if not isinstance(data['G_ema'], torch.nn.Module):
        raise RuntimeError(f"G_ema is a {getattr(data['G_ema'], 'class_name', type(data['G_ema']))}, not a TriPlaneGenerator")
    return data
