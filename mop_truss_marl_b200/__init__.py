"""B200-native batched 2D-truss FEM environment (drop-in for the hot path of kupc25648/MOP-truss-MARL).

The product path is the CUDA library ``lib/libtfem.so`` (C ABI in ``include/tfem.h``); there is no CPU
fallback -- importing :mod:`mop_truss_marl_b200.capi` without the built library raises.
"""
from .families import FAMILIES, FamilySpec, family_desc  # noqa: F401

__version__ = "0.1.0"
