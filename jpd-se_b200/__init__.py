"""jpd-se_b200: B200-native (sm_100a) hot path of SenseBrain/JPD-SE.

The directory name carries a hyphen, so import it as ``import jpdse_b200`` (alias module at the repo
root) or ``importlib.import_module("jpd-se_b200")``.

Layout
  csrc/      hand-written CUDA kernels + the C ABI (include/jpdse_b200.h) -> libjpdse_b200.so
  _lib.py    ctypes binding of the C ABI (raises if the library is missing: no fallback)
  ops.py     torch-tensor front end of the C ABI
  engine.py  forward plan of the GlobalGenerator on those kernels
  ctu/       host-side mirror of the reference's ``ctu`` interface for this path
"""
from . import _lib  # noqa: F401
from ._lib import JpdseError  # noqa: F401

__all__ = ["JpdseError", "install_into_reference"]


def install_into_reference():
    """Patch an importable reference ``ctu`` package so its generator runs on jpdse_b200."""
    from .ctu import install
    return install()
