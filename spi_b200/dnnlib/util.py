import importlib


class EasyDict(dict):
    """dict with attribute access (eg3d/dnnlib/util.py:42-55)."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value

    def __delattr__(self, name):
        del self[name]


def get_obj_by_name(name):
    """'training.superresolution.Foo' -> class object; bare `training.` / `torch_utils.` prefixes map to spi_b200."""
    module_name, _, obj_name = name.rpartition('.')
    candidates = [module_name]
    if module_name.split('.')[0] in ('training', 'torch_utils', 'dnnlib'):
        candidates.insert(0, 'spi_b200.' + module_name)
    last = None
    for cand in candidates:
        try:
            return getattr(importlib.import_module(cand), obj_name)
        except (ImportError, AttributeError) as e:
            last = e
    raise last


def construct_class_by_name(*args, class_name=None, **kwargs):
    """eg3d/dnnlib/util.py:303-306."""
    assert class_name is not None
    return get_obj_by_name(class_name)(*args, **kwargs)
