"""`persistent_class` surface needed by the loaders (eg3d/torch_utils/persistence.py:37): instances expose
`init_args` / `init_kwargs` so `load_utils.load_eg3d` can rebuild the module (spi/utils/load_utils.py:25)."""
import copy
import functools


def persistent_class(cls):
    orig_init = cls.__init__

    @functools.wraps(orig_init)
    def __init__(self, *args, **kwargs):
        if not hasattr(self, '_init_args'):
            self._init_args = copy.deepcopy(args)
            self._init_kwargs = copy.deepcopy(kwargs)
        orig_init(self, *args, **kwargs)

    cls.__init__ = __init__
    cls.init_args = property(lambda self: copy.deepcopy(self._init_args))
    cls.init_kwargs = property(lambda self: EasyDictLike(copy.deepcopy(self._init_kwargs)))
    return cls


class EasyDictLike(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)
