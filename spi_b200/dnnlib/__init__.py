"""Minimal `dnnlib` surface used on the inversion path (eg3d/dnnlib/util.py:42,303): EasyDict and
construct_class_by_name.  Module names of the reference (`training.superresolution.X`) resolve to spi_b200's."""
import importlib

from . import util  # noqa: F401
from .util import EasyDict, construct_class_by_name  # noqa: F401
