"""TEST INFRASTRUCTURE ONLY -- ctypes loader of the C restatement ``oracle/truss_oracle.c`` (whole batches, OpenMP).

``build()`` compiles it with gcc into ``oracle/_build/libtruss_oracle.so`` (git-ignored; travels to the GPU box with the
snapshot).  The family tables come from the pinned Python oracle (``oracle/truss_oracle.build_mesh``), never from the
product library.  Only ``tests/``, ``__graft_entry__`` and ``bench.py``'s CPU-baseline legs may import this module."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from .truss_oracle import LONG_STRESS, SECTION_TABLE, YOUNG, TrussOracle

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "truss_oracle.c")
LIB = os.path.join(HERE, "_build", "libtruss_oracle.so")


def build(force: bool = False) -> str:
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.run(["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", "-fno-fast-math",
                        "-o", LIB, SRC, "-lm"], check=True)
    return LIB


class _Family(C.Structure):
    _fields_ = [("N", C.c_int32), ("E", C.c_int32), ("ndof", C.c_int32), ("truss_type", C.c_int32),
                ("conn", C.c_void_p), ("tnsc", C.c_void_p), ("res", C.c_void_p), ("top", C.c_void_p), ("pair", C.c_void_p),
                ("sym_src", C.c_void_p), ("sym_pairs", C.c_void_p), ("npairs", C.c_int32), ("pad", C.c_int32),
                ("x", C.c_void_p), ("target", C.c_void_p), ("P", C.c_void_p), ("sec_area", C.c_void_p),
                ("y_min", C.c_double), ("y_max", C.c_double), ("d_min", C.c_double), ("max_def", C.c_double),
                ("young", C.c_double), ("allow", C.c_double), ("int_obj1", C.c_float), ("int_obj2", C.c_float)]


_OUT_FIELDS = ("y", "weak", "section", "move_range", "d", "axial", "ratio", "iscompress", "U", "point", "status")


class _Out(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in _OUT_FIELDS]


class COracle:
    """``COracle("small_bridge").step(...)`` / ``.solve(...)`` for whole batches (numpy arrays in, dict of arrays out)"""

    def __init__(self, family, threads: int | None = None):
        self.py = TrussOracle(family)
        m = self.py.mesh
        self.lib = C.CDLL(build())
        self.lib.truss_oracle_step.argtypes = [C.POINTER(_Family), C.c_int] + [C.c_void_p] * 6 + [C.POINTER(_Out), C.c_int]
        self.lib.truss_oracle_solve.argtypes = [C.POINTER(_Family), C.c_int, C.c_void_p, C.c_void_p, C.POINTER(_Out), C.c_int]
        self.threads = int(threads or os.cpu_count() or 1)
        self.N, self.E, self.ndof = m.N, m.E, m.ndof
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)      # noqa: E731
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)    # noqa: E731
        self._keep = {
            "conn": i32(m.conn), "tnsc": i32(m.tnsc), "res": i32(m.res), "top": i32(m.top), "pair": i32(m.pair),
            "sym_src": i32([m.sym_src_false, m.sym_src_true]), "sym_pairs": i32(m.sym_elem_pairs).reshape(-1, 2),
            "x": f64(m.x), "target": f64([t if t is not None else 0.0 for t in m.target_top]), "P": f64(m.P),
            "sec_area": f64(SECTION_TABLE[:, 0] * 1e-4),
        }
        f = _Family()
        f.N, f.E, f.ndof = m.N, m.E, m.ndof
        f.truss_type = 0 if m.spec.truss_type == "bridge" else 1
        for k, v in self._keep.items():
            setattr(f, k, v.ctypes.data)
        f.npairs = self._keep["sym_pairs"].shape[0]
        f.y_min, f.y_max, f.d_min, f.max_def = float(m.y_min), float(m.y_max), float(m.d_min), float(m.max_deformation)
        f.young, f.allow = YOUNG, LONG_STRESS
        f.int_obj1, f.int_obj2 = self.py.int_obj1, self.py.int_obj2
        self._f = f

    def _outputs(self, B):
        N, E, n = self.N, self.E, self.ndof
        out = {"y": np.empty((B, N)), "weak": np.empty((B, N), dtype=np.uint8), "section": np.empty((B, E), dtype=np.int32),
               "move_range": np.empty((B, N, 2), dtype=np.float32), "d": np.empty((B, n)), "axial": np.empty((B, E)),
               "ratio": np.empty((B, E)), "iscompress": np.empty((B, E), dtype=np.int32), "U": np.empty(B),
               "point": np.empty((B, 4), dtype=np.float32), "status": np.zeros(B, dtype=np.int32)}
        o = _Out()
        for k in _OUT_FIELDS:
            setattr(o, k, out[k].ctypes.data)
        return out, o

    def step(self, set_node, set_element, a_geo, a_topo, coin, stale_range):
        """batched ``_game_modify``: set_node [B,N,12], set_element [B,E,21], a_geo [B,N,2], a_topo [B,N,3] float32 (the
        action arrays are clipped IN PLACE), coin [B] uint8, stale_range [B,N,2] float32 (read only)"""
        B = set_node.shape[0]
        for a in (set_node, set_element, a_geo, a_topo, stale_range):
            assert a.dtype == np.float32 and a.flags.c_contiguous
        coin = np.ascontiguousarray(coin, dtype=np.uint8)
        out, o = self._outputs(B)
        rc = self.lib.truss_oracle_step(C.byref(self._f), B, set_node.ctypes.data, set_element.ctypes.data, a_geo.ctypes.data,
                                        a_topo.ctypes.data, coin.ctypes.data, stale_range.ctypes.data, C.byref(o), self.threads)
        assert rc == 0
        return out

    def solve(self, y64, section):
        y64 = np.ascontiguousarray(y64, dtype=np.float64)
        section = np.ascontiguousarray(section, dtype=np.int32)
        B = y64.shape[0]
        out, o = self._outputs(B)
        rc = self.lib.truss_oracle_solve(C.byref(self._f), B, y64.ctypes.data, section.ctypes.data, C.byref(o), self.threads)
        assert rc == 0
        for k in ("y", "weak", "section", "move_range"):
            out.pop(k)
        return out
