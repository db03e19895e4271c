"""Drop-in modules named like the reference's (``FEM_2Dtruss``, ``truss2D_GEN``, ``truss2D_ENV``,
``truss2D_RL``).  Put this directory first on ``sys.path`` (``compat.install()``) and the unchanged
``master_DDPG_truss2D_MO.py`` star-imports these instead of the reference modules; every FEM / env-step /
actor evaluation then runs as a batch of one through ``libtfem.so`` (no CPU fallback)."""
import os
import sys

COMPAT_DIR = os.path.dirname(os.path.abspath(__file__))


def install():
    """make ``import FEM_2Dtruss`` etc. resolve to these modules"""
    if COMPAT_DIR not in sys.path:
        sys.path.insert(0, COMPAT_DIR)
    for name in ("FEM_2Dtruss", "truss2D_GEN", "truss2D_ENV", "truss2D_RL"):
        sys.modules.pop(name, None)
