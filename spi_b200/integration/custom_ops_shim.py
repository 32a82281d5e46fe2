"""Drop-in for `custom_ops.get_plugin` (eg3d/torch_utils/custom_ops.py:61-157): instead of JIT-compiling the reference's three CUDA
plugins it returns objects with the same callables, bound to libspi_b200.so through the C ABI of include/spi_b200.h.

    # eg3d/torch_utils/custom_ops.py
    from spi_b200.integration.custom_ops_shim import get_plugin

The reference's Python op modules then run unmodified: `bias_act.py:40,138` calls `_plugin.bias_act(x, b, xref, yref, dy, grad, dim, act,
alpha, gain, clamp)`, `upfirdn2d.py:25,235` calls `_plugin.upfirdn2d(x, f, upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip, gain)`,
`filtered_lrelu.py:25,214,229` call `_plugin.filtered_lrelu(...)` -> `(y, so, return_code)` and `_plugin.filtered_lrelu_act_(...)` -> `so`.
Conventions kept: empty tensor = absent operand, `clamp < 0` = off, `act` = `cuda_idx` 1..9, argument errors raise
`RuntimeError` (as TORCH_CHECK does), `return_code = -1` = "no specialised kernel" so that filtered_lrelu.py:225-232 falls back to its
generic path, work is enqueued on torch's current stream without host synchronisation.
tests/test_gpu_plugin_shim.py drives these objects with the reference's forward / backward / double-backward call protocol.
"""
import torch

from ..torch_utils.ops import bias_act as _ba
from ..torch_utils.ops import filtered_lrelu as _fl
from ..torch_utils.ops import upfirdn2d as _uf


def _opt(t):
    """The plugins receive `torch.empty([0])` for absent operands (bias_act.py:127, filtered_lrelu.py:187)."""
    return None if (t is None or (isinstance(t, torch.Tensor) and t.numel() == 0)) else t


class _BiasActPlugin:                                   # bias_act.cpp:36,100
    @staticmethod
    def bias_act(x, b, xref, yref, dy, grad, dim, act, alpha, gain, clamp):
        return _ba._plugin_bias_act(x, _opt(b), _opt(xref), _opt(yref), _opt(dy), int(grad), int(dim), int(act), float(alpha), float(gain), float(clamp))


class _Upfirdn2dPlugin:                                 # upfirdn2d.cpp:20,108
    @staticmethod
    def upfirdn2d(x, f, upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip, gain):
        return _uf._plugin_upfirdn2d(x, f, int(upx), int(upy), int(downx), int(downy), int(padx0), int(padx1), int(pady0), int(pady1), bool(flip), float(gain))


class _FilteredLReluPlugin:                             # filtered_lrelu.cpp:20,217,300-301
    @staticmethod
    def filtered_lrelu(x, fu, fd, b, si, up, down, px0, px1, py0, py1, sx, sy, gain, slope, clamp, flip_filters, writeSigns):
        y, so, rc = _fl._plugin_filtered_lrelu(x, fu, fd, b, _opt(si), int(up), int(down), int(px0), int(px1), int(py0), int(py1), int(sx), int(sy),
                                               float(gain), float(slope), float(clamp), bool(flip_filters), bool(writeSigns))
        if rc != 0:
            return torch.empty([0], device=x.device), torch.empty([0], device=x.device), rc
        return y, (so if so is not None else torch.empty([0], dtype=torch.uint8, device=x.device)), rc

    @staticmethod
    def filtered_lrelu_act_(x, si, sx, sy, gain, slope, clamp, writeSigns):
        so = _fl._plugin_filtered_lrelu_act_(x, _opt(si), int(sx), int(sy), float(gain), float(slope), float(clamp), bool(writeSigns))
        return so if so is not None else torch.empty([0], dtype=torch.uint8, device=x.device)


_PLUGINS = {'bias_act_plugin': _BiasActPlugin, 'upfirdn2d_plugin': _Upfirdn2dPlugin, 'filtered_lrelu_plugin': _FilteredLReluPlugin}


def get_plugin(module_name, sources=None, headers=None, source_dir=None, **build_kwargs):
    """Same signature as custom_ops.get_plugin; `sources`, `headers`, `source_dir` and the build flags are accepted and ignored."""
    try:
        return _PLUGINS[module_name]
    except KeyError:
        raise RuntimeError(f'spi_b200 has no plugin named {module_name!r} (known: {sorted(_PLUGINS)})')
