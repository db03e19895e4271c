"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference modules.

This file imports ``FEM_2Dtruss`` / ``truss2D_GEN`` / ``truss2D_ENV`` straight from
``/root/reference/<run dir>/code`` (read-only, never copied) so that

  * ``oracle/truss_oracle.py`` (our CPU restatement) can be pinned against the real thing, and
  * ``tests/golden/make_golden.py`` can dump golden input/output vectors.

``/root/reference`` exists only in the build container, so nothing that runs on the GPU box may
import this module (``available()`` gates the tests that do).  SURVEY.md section 8c describes the
three stub modules the reference needs when TensorFlow / spektral / matplotlib are absent:

  * ``matplotlib{,.pyplot,.animation}``, ``mpl_toolkits.mplot3d{,.axes3d}`` -- plot only
    (truss2D_GEN.py:10-13)
  * ``spektral.utils.degree_power`` (truss2D_ENV.py:3) -- published definition in spektral 1.2.0:
    ``diag(power(A.sum(1), k))`` with inf -> 0.

"FEM-coerced" mode (``fem_fp64=True``): inside ``Model.gen_all`` every node's y coordinate is turned
into a Python float first and the original object is restored afterwards.  The solver therefore runs
in pure float64 on every NumPy version (the parity definition of SURVEY.md section 8c-2), while the
transition, the move-range rule, the observation builders and the objectives keep the reference's raw
float32 / weak-scalar behaviour.
"""
from __future__ import annotations

import contextlib
import importlib
import io
import os
import random
import sys
import types

import numpy as np

REF_ROOT = os.environ.get("TRUSS_REFERENCE_ROOT", "/root/reference")

RUN_DIRS = {
    "small_bridge": "test/00_small_bridge/code",
    "small_roof": "test/01_small_roof/code",
    "large_bridge": "test/02_large_bridge/code",
    "large_roof": "test/03_large_roof/code",
    "train": "train/code",
}

_REF_MODULE_NAMES = ("FEM_2Dtruss", "truss2D_GEN", "truss2D_ENV", "utils", "set_seed_global")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, RUN_DIRS["small_bridge"]))


def _install_stubs() -> None:
    def mod(name):
        m = sys.modules.get(name)
        if m is None:
            m = types.ModuleType(name)
            sys.modules[name] = m
        return m

    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.animation",
                 "mpl_toolkits", "mpl_toolkits.mplot3d", "mpl_toolkits.mplot3d.axes3d",
                 "spektral", "spektral.utils"):
        try:
            importlib.import_module(name)
        except Exception:
            mod(name)
    m3d = sys.modules["mpl_toolkits.mplot3d"]
    if not hasattr(m3d, "Axes3D"):
        m3d.Axes3D = object
    su = sys.modules["spektral.utils"]
    if not hasattr(su, "degree_power"):
        def degree_power(A, k):
            with np.errstate(divide="ignore"):
                deg = np.power(np.array(A.sum(1)), k).ravel()
            deg[np.isinf(deg)] = 0.0
            return np.diag(deg)
        su.degree_power = degree_power


def load_utils(run: str = "small_bridge"):
    """the reference's ``utils.py`` (Pareto filter, hypervolume) of one run directory, imported in isolation"""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    code_dir = os.path.join(REF_ROOT, RUN_DIRS[run])
    saved = {k: sys.modules.pop(k) for k in ("utils", "set_seed_global") if k in sys.modules}
    sys.path.insert(0, code_dir)
    try:
        return importlib.import_module("utils")
    finally:
        sys.path.remove(code_dir)
        for k in ("utils", "set_seed_global"):
            sys.modules.pop(k, None)
        sys.modules.update(saved)


class RefModules:
    """The reference modules of one run directory, imported in isolation."""

    def __init__(self, run: str):
        if not available():
            raise RuntimeError("reference tree not present at %s" % REF_ROOT)
        self.run = run
        self.code_dir = os.path.join(REF_ROOT, RUN_DIRS[run])
        _install_stubs()
        saved = {k: sys.modules.pop(k) for k in _REF_MODULE_NAMES if k in sys.modules}
        sys.path.insert(0, self.code_dir)
        try:
            with self.cwd():
                self.FEM = importlib.import_module("FEM_2Dtruss")
                self.GEN = importlib.import_module("truss2D_GEN")
                self.ENV = importlib.import_module("truss2D_ENV")
        finally:
            sys.path.remove(self.code_dir)
            for k in _REF_MODULE_NAMES:
                sys.modules.pop(k, None)
            sys.modules.update(saved)
        self._orig_gen_all = self.FEM.Model.gen_all

    @contextlib.contextmanager
    def cwd(self):
        old = os.getcwd()
        os.chdir(self.code_dir)
        try:
            yield
        finally:
            os.chdir(old)

    def set_fem_fp64(self, on: bool) -> None:
        orig = self._orig_gen_all
        if not on:
            self.FEM.Model.gen_all = orig
            return

        def gen_all_fp64(model):
            keep = [n.coord[1] for n in model.nodes]
            for n in model.nodes:
                n.coord[1] = float(n.coord[1])
            try:
                orig(model)
            finally:
                for n, y in zip(model.nodes, keep):
                    n.coord[1] = y

        self.FEM.Model.gen_all = gen_all_fp64


# geometry/load constants of the reference drivers (master_DDPG_truss2D_MO.py)
_LARGE_TAR = [3, 2.75, 2.5, 2.25, 2.25, 2, 2, 2, 2, 2, 2, 2.25, 2.25, 2.5, 2.75, 3]
_SMALL_TAR = [4, 3, 2.5, 2, 2, 2.5, 3, 4]
DRIVER_ARGS = {
    # num_x, num_y, span_x, span_y, tar_y, dmin, loadx, loady, truss_type, support_case, topo_code
    "small_bridge": (8, 2, [5] * 7, [8], _SMALL_TAR, 0.3, 0, -75 * 1000, "bridge", 1, None),
    "small_roof": (8, 2, [5] * 7, [8], _SMALL_TAR, 0.3, 0, -120 * 1000, "roof", 1, None),
    "large_bridge": (16, 2, [5] * 15, [6], _LARGE_TAR, 0.3, 0, -7.5 * 1000, "bridge", 1, None),
    "large_roof": (16, 2, [5] * 15, [6], _LARGE_TAR, 0.3, 0, -8 * 1000, "roof", 1, None),
}

# the 6 x 2 training shapes (train/code/master_DDPG_truss2D_MO.py:787-795, :809-823): the driver passes (..., truss_type,
# gen_topo_code, support_case) positionally, so gen_model receives support_case = None (which generates the supports of
# case 1) and topo_code = 1
_TRAIN_TAR = ([1.0, 1.5, 2.0, 2.0, 1.5, 1.0], [1.0, 3.0, 3.0, 2.0, 1.5, 1.0], [1.0, 1.5, 2.0, 3.0, 3.0, 1.0],
              [1.0, 3.0, 2.0, 2.0, 3.0, 1.0], [3.0, 2.0, 1.0, 1.0, 2.0, 3.0])
for _i, _tar in enumerate(_TRAIN_TAR):
    for _tt in ("roof", "bridge"):
        RUN_DIRS["train%d_%s" % (_i, _tt)] = "train/code"
        DRIVER_ARGS["train%d_%s" % (_i, _tt)] = (6, 2, [4.0, 3.0, 5.0, 3.0, 5.0], [5], list(_tar), 0.2, 0, -100000, _tt, None, 1)


class RefGame:
    """One reference ``gen_model`` + ``Game_research04`` pair, driven with explicit coins."""

    def __init__(self, run: str, fem_fp64: bool = True, args=None, mods: RefModules | None = None):
        self.mods = mods or RefModules(run)
        self.mods.set_fem_fp64(fem_fp64)
        self.args = args or DRIVER_ARGS[run]
        with self.mods.cwd(), contextlib.redirect_stdout(io.StringIO()):
            self.gen = self.mods.GEN.gen_model(*self.args)
            self.game = self.mods.ENV.Game_research04(500, self.gen, 3)

    def reset_state(self):
        return self.game._game_get_1_state()

    def step(self, set_node, set_element, nC_e, a_geo, a_topo, coin: bool):
        """``_game_modify`` with the symmetry coin forced (truss2D_ENV.py:460 draws
        ``random.random() >= 0.5``)."""
        env_random = getattr(self.mods.ENV, "random", None)
        if env_random is None:                               # train/code's ENV draws no coin (no symmetry step)
            return self.game._game_modify(set_node, set_element, nC_e, [a_geo, a_topo])
        orig = env_random.random
        env_random.random = (lambda: 0.75) if coin else (lambda: 0.25)
        try:
            return self.game._game_modify(set_node, set_element, nC_e, [a_geo, a_topo])
        finally:
            env_random.random = orig

    # ---- FP64 views of the solved model (for 1e-9 parity) -----------------------------------------
    def fem_fields(self):
        m = self.gen.model
        with np.errstate(all="ignore"):
            out = {
                "y": np.array([float(n.coord[1]) for n in m.nodes], dtype=np.float64),
                "section": np.array([e.section_no for e in m.elements], dtype=np.int32),
                "d": np.array(m.d, dtype=np.float64).reshape(-1),
                "axial": np.array([float(e.e_q[0][0]) for e in m.elements], dtype=np.float64),
                "ratio": np.array([float(e.prop_yeield) for e in m.elements], dtype=np.float64),
                "iscompress": np.array([int(e.iscompress) for e in m.elements], dtype=np.int32),
                "length": np.array([float(e.length) for e in m.elements], dtype=np.float64),
                "U": float(np.asarray(m.U_full).reshape(-1)[0]),
                "reactions": np.array([r for r in m.r if r is not None], dtype=np.float64),
                "max_up": np.array([np.float32(n.max_up) for n in m.nodes], dtype=np.float32),
                "max_down": np.array([np.float32(n.max_down) for n in m.nodes], dtype=np.float32),
                "tnsc": np.array(m.tnsc, dtype=np.int32),
                "ndof": int(m.ndof),
            }
        return out

    def set_move_range(self, max_up, max_down):
        for n, u, d in zip(self.gen.model.nodes, max_up, max_down):
            n.max_up = np.float32(u)
            n.max_down = np.float32(d)


def seed_all(seed: int) -> None:
    random.seed(seed)
    np.random.seed(seed)


# ----------------------------------------------------------------------------------------------------
# MOEA/D benchmark: read_genes of test/benchmarks/MOEAD/<family>.zip
# ----------------------------------------------------------------------------------------------------
MOEAD_ZIPS = {
    "small_bridge": "test/benchmarks/MOEAD/00_small_bridge.zip",
    "small_roof": "test/benchmarks/MOEAD/01_small_roof.zip",
    "large_bridge": "test/benchmarks/MOEAD/02_large_bridge.zip",
    "large_roof": "test/benchmarks/MOEAD/03_large_roof.zip",
}


class RefMoead:
    """The unmodified ``truss2D_GEN.py`` / ``FEM_2Dtruss.py`` of one MOEA/D benchmark zip, extracted to a temporary
    directory (never into the repo) and imported in isolation; ``read_genes`` is the call of
    ``MOEAD_master.py:104`` with ``int_obj1 / int_obj2`` computed as ``MOEAD_master.py:53-64`` does."""

    def __init__(self, run: str):
        import tempfile
        import zipfile
        if not available():
            raise RuntimeError("reference tree not present at %s" % REF_ROOT)
        self.run = run
        self._tmp = tempfile.TemporaryDirectory(prefix="moead_ref_")
        with zipfile.ZipFile(os.path.join(REF_ROOT, MOEAD_ZIPS[run])) as z:
            top = MOEAD_ZIPS[run].rsplit("/", 1)[1][:-4]
            z.extractall(self._tmp.name, [n for n in z.namelist() if n.startswith(top + "/") and
                                          (n.endswith(".py") or "section_data/" in n)])
        self.code_dir = os.path.join(self._tmp.name, top)
        _install_stubs()
        names = ("FEM_2Dtruss", "truss2D_GEN", "set_seed_global")
        saved = {k: sys.modules.pop(k) for k in names if k in sys.modules}
        sys.path.insert(0, self.code_dir)
        old = os.getcwd()
        os.chdir(self.code_dir)
        try:
            self.GEN = importlib.import_module("truss2D_GEN")
            with contextlib.redirect_stdout(io.StringIO()):
                self.gen = self.GEN.gen_model(*DRIVER_ARGS[run])
        finally:
            os.chdir(old)
            sys.path.remove(self.code_dir)
            for k in names:
                sys.modules.pop(k, None)
            sys.modules.update(saved)
        m = self.gen.model
        all_v = np.zeros(len(m.elements), dtype=np.float32)
        for i, e in enumerate(m.elements):
            all_v[i] = e.area * e.length
        all_dt = np.zeros(len(m.nodes), dtype=np.float32)
        for i, n in enumerate(m.nodes):
            if n.top_node == 1:
                all_dt[i] = abs(n.target - n.coord[1])
        self.int_obj1, self.int_obj2 = np.sum(all_v), np.sum(all_dt)

    def read_genes(self, genes):
        with np.errstate(all="ignore"), contextlib.redirect_stdout(io.StringIO()):
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                point = self.gen.read_genes(genes, self.int_obj1, self.int_obj2)
        m = self.gen.model
        return {
            "point": np.array(point, dtype=np.float32),
            "y": np.array([float(n.coord[1]) for n in m.nodes], dtype=np.float64),
            "section": np.array([e.section_no for e in m.elements], dtype=np.int32),
            "d": np.array(m.d, dtype=np.float64).reshape(-1),
            "axial": np.array([float(e.e_q[0][0]) for e in m.elements], dtype=np.float64),
            "ratio": np.array([float(e.prop_yeield) for e in m.elements], dtype=np.float64),
        }
