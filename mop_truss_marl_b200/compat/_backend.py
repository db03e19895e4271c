"""Batch-of-one bridge between the drop-in object model and ``libtfem.so`` (host logic only)."""
import ctypes as C

import numpy as np

from mop_truss_marl_b200 import capi
from mop_truss_marl_b200.batched_env import step_host
from mop_truss_marl_b200.families import SYM_LARGE, SYM_NONE, SYM_SMALL, FamilySpec, family_desc

_DEVICE = 0


def set_device(index: int):
    global _DEVICE
    _DEVICE = int(index)


def _np_or_py(value: float, weak: bool):
    """the object the reference would hold in node.coord[1]"""
    if not weak:
        return np.float32(value)
    return int(value) if float(value).is_integer() else float(value)


class Backend:
    def __init__(self, spec: FamilySpec):
        import torch
        if not torch.cuda.is_available():
            raise capi.TfemError("the drop-in modules need a CUDA device: libtfem has no CPU path")
        self.torch = torch
        self.spec = spec
        self.handle = capi.Handle(family_desc(spec), _DEVICE)
        d = self.handle.dims
        self.N, self.E, self.ndof, self.nres = d.N, d.E, d.ndof, d.nres
        self.tab = {k: self.handle.table(k) for k in ("conn", "tnsc", "res", "top", "pair", "loaded", "loadvec",
                                                      "x", "y0", "target", "A_n", "mask", "nC_e", "int_obj")}
        self.dev = torch.device("cuda", _DEVICE)
        self.analyses = 0               # FEM solves served (the driver's `my_step` counter counts the same thing)

    # ---- FEM only ------------------------------------------------------------------------------------
    def solve_into(self, model):
        torch = self.torch
        y = np.array([[float(n.coord[1]) for n in model.nodes]], dtype=np.float64)
        sec = np.array([[int(e.section_no) for e in model.elements]], dtype=np.int32)
        ty, ts = torch.from_numpy(y).to(self.dev), torch.from_numpy(sec).to(self.dev)
        f64 = dict(dtype=torch.float64, device=self.dev)
        out = {"d": torch.empty(1, self.ndof, **f64), "axial": torch.empty(1, self.E, **f64),
               "ratio": torch.empty(1, self.E, **f64), "U": torch.empty(1, **f64),
               "reactions": torch.empty(1, self.nres, **f64),
               "status": torch.zeros(1, dtype=torch.int32, device=self.dev)}
        p = lambda t: C.c_void_p(t.data_ptr())   # noqa: E731
        st = C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
        capi.check(capi.lib.tfem_solve_only(self.handle.ptr, 1, p(ty), p(ts), p(out["d"]), p(out["axial"]),
                                            p(out["ratio"]), p(out["U"]), p(out["reactions"]), p(out["status"]), st))
        res = {k: v.cpu().numpy() for k, v in out.items()}
        self.analyses += 1
        self.fill_results(model, res)

    def fill_results(self, model, res):
        """res: dict of batch-of-one arrays d, axial, ratio, U, reactions, status"""
        if int(res["status"][0]) != 0:
            raise np.linalg.LinAlgError("Singular matrix")          # FEM_2Dtruss.py:337 raises the same
        tnsc, ndof = self.tab["tnsc"], self.ndof
        d = res["d"][0]
        model.tnsc = tnsc.tolist()
        model.ndof = ndof
        model.jlv = [[v] for v in self.tab["loadvec"].tolist()]
        model.d = d.reshape(-1, 1).copy()
        model.U_full = np.array([float(res["U"][0])])
        for i, n in enumerate(model.nodes):
            dx = float(d[tnsc[i, 0] - 1]) if tnsc[i, 0] <= ndof else 0
            dy = float(d[tnsc[i, 1] - 1]) if tnsc[i, 1] <= ndof else 0
            n.global_d = [[dx], [dy]]
        for e, el in enumerate(model.elements):
            q = float(res["axial"][0, e])
            el.gen_length()
            el.e_q = np.array([[q], [0.0], [-q], [0.0]])
            el.prop_yeield = float(res["ratio"][0, e])
            el.iscompress = 0 if q <= 0 else 1
        r = [None] * (2 * self.N)
        for k in range(self.nres):
            r[ndof + k] = float(res["reactions"][0, k])
        model.r = r
        for n in model.nodes:
            n.set_target()

    # ---- gene-vector objective ------------------------------------------------------------------------
    def read_genes(self, genes, max_height, int_obj1, int_obj2):
        torch = self.torch
        g = torch.from_numpy(np.ascontiguousarray(genes, dtype=np.float64).reshape(1, -1)).to(self.dev)
        if g.shape[1] != self.N + self.E:
            raise ValueError("expected %d genes" % (self.N + self.E))
        f64 = dict(dtype=torch.float64, device=self.dev)
        out = {"point": torch.empty(1, 4, dtype=torch.float32, device=self.dev), "y": torch.empty(1, self.N, **f64),
               "section": torch.empty(1, self.E, dtype=torch.int32, device=self.dev),
               "d": torch.empty(1, self.ndof, **f64), "axial": torch.empty(1, self.E, **f64),
               "ratio": torch.empty(1, self.E, **f64), "U": torch.empty(1, **f64),
               "reactions": torch.empty(1, self.nres, **f64),
               "status": torch.zeros(1, dtype=torch.int32, device=self.dev)}
        o = capi.GenesOut()
        for k, v in out.items():
            setattr(o, k, C.c_void_p(v.data_ptr()))
        st = C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
        capi.check(capi.lib.tfem_read_genes(self.handle.ptr, 1, C.c_void_p(g.data_ptr()), float(max_height),
                                            float(int_obj1), float(int_obj2), C.byref(o), st))
        self.analyses += 1
        return {k: v.cpu().numpy() for k, v in out.items()}

    # ---- full env step -------------------------------------------------------------------------------
    def game_modify(self, set_node, set_element, move_range, a_geo, a_topo, coin):
        out = step_host(self.handle, np.ascontiguousarray(set_node, dtype=np.float32)[None],
                        np.ascontiguousarray(set_element, dtype=np.float32)[None], move_range, a_geo[None], a_topo[None],
                        np.array([1 if coin else 0], dtype=np.uint8))
        self.analyses += 1
        return out

    def reset_state(self):
        torch = self.torch
        N, E = self.N, self.E
        f32 = dict(dtype=torch.float32, device=self.dev)
        f64 = dict(dtype=torch.float64, device=self.dev)
        bufs = {"x_n": torch.empty(1, N, 13, **f32), "A_s": torch.empty(1, N, N, **f32),
                "A_n_ts": torch.empty(1, N, N, **f32), "A_n_cs": torch.empty(1, N, N, **f32),
                "nN_x_n": torch.empty(1, N, 12, **f32), "nN_x_e": torch.empty(1, E, 21, **f32),
                "point": torch.empty(1, 4, **f32), "d": torch.empty(1, self.ndof, **f64),
                "axial": torch.empty(1, E, **f64), "ratio": torch.empty(1, E, **f64), "U": torch.empty(1, **f64),
                "reactions": torch.empty(1, self.nres, **f64),
                "status": torch.zeros(1, dtype=torch.int32, device=self.dev)}
        mr = torch.empty(1, N, 2, **f32)
        o = capi.StepOut()
        for k, v in bufs.items():
            setattr(o, k, C.c_void_p(v.data_ptr()))
        st = C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
        capi.check(capi.lib.tfem_reset(self.handle.ptr, 1, C.c_void_p(mr.data_ptr()), C.byref(o), st))
        self.analyses += 1
        out = {k: v.cpu().numpy() for k, v in bufs.items()}
        out["move_range"] = mr.cpu().numpy()
        return out


def symmetry_of(run_hint, num_x):
    """which truss2D_ENV.py convention a drop-in Game should follow: the 8-column meshes use the
    test/00,01 file, the 16-column ones test/02,03; anything else has no symmetry pass (train/code)"""
    if run_hint is not None:
        return run_hint
    return {8: SYM_SMALL, 16: SYM_LARGE}.get(num_x, SYM_NONE)
