"""opensayal_b200 — B200-native (sm_100a) implementation of OpenSayal's per-step simulation path.

The package holds only what the path needs: csrc/ (CUDA kernels + the C ABI of include/sayal.h) and a
ctypes mirror of the reference's `ConfigParser` / `Source` / `Fluid` interface.
"""
from ._abi import LIB_PATH, SayalError, load  # noqa: F401
from .fluid import Config, ConfigParser, Fluid, Source, Visual  # noqa: F401

__all__ = ["Config", "ConfigParser", "Fluid", "Source", "Visual", "SayalError", "load", "LIB_PATH"]
