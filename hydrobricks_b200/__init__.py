"""hydrobricks_b200 — B200-native batched simulation engine for hydrobricks' ModelHydro::Run() hot path.

Layers (bottom-up):
  csrc/hbk_device.cuh, csrc/hbk_engine.cu   CUDA sm_100a kernels + the C-ABI of include/hbk.h (libhbk.so)
  engine.py                                 ctypes binding of the C-ABI
  structure.py                              structure compiler (SettingsModel structure -> process table)
There is no CPU implementation in this package.
"""
from . import engine, structure  # noqa: F401

__all__ = ["engine", "structure"]
