"""Gene-vector objective of the MOEA/D benchmark (``read_genes``, ``test/benchmarks/MOEAD/<family>.zip:
<family>/truss2D_GEN.py:117-228``; caller ``MOEAD_master.py:103-121``): a whole population in one launch of the
env-step kernel through ``tfem_read_genes`` (``include/tfem.h``).  No CPU path."""
from __future__ import annotations

import ctypes as C

import torch

from . import capi
from .families import FAMILIES, family_desc

# ``max_height = 8 # change this`` in the small zips, ``6`` in the large ones (truss2D_GEN.py:127)
MAX_HEIGHT = {8: 8.0, 16: 6.0}


class GeneEvaluator:
    """``GeneEvaluator("small_bridge").read_genes(genes[B, N+E]) -> point[B, 4]`` (float32, on the device);
    the FP64 fields of the solved models stay in ``self.out`` until the next call."""

    def __init__(self, family, device="cuda:0", max_height=None, handle=None):
        self.spec = FAMILIES[family] if isinstance(family, str) else family
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise capi.TfemError("GeneEvaluator needs a CUDA device: libtfem has no CPU path")
        index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", index)
        self.handle = handle or capi.Handle(family_desc(self.spec), index)
        d = self.handle.dims
        self.N, self.E, self.ndof = d.N, d.E, d.ndof
        self.max_height = float(MAX_HEIGHT.get(d.num_x, self.spec.span_y[0]) if max_height is None else max_height)
        self.out = {}

    def read_genes(self, genes, int_obj1=0.0, int_obj2=0.0, fields=("point", "status")):
        """genes: [B, N+E] float64 tensor on the device (or anything ``torch.as_tensor`` takes; host data is copied).
        ``fields``: which of point, point64, y, section, d, axial, ratio, U, reactions, status to produce."""
        g = torch.as_tensor(genes, dtype=torch.float64).to(self.device).contiguous()
        if g.dim() == 1:
            g = g[None]
        if g.dim() != 2 or g.shape[1] != self.N + self.E:
            raise ValueError("genes must have shape [B, %d]" % (self.N + self.E))
        B = g.shape[0]
        shapes = {"point": ((B, 4), torch.float32), "point64": ((B, 4), torch.float64), "y": ((B, self.N), torch.float64),
                  "section": ((B, self.E), torch.int32), "d": ((B, self.ndof), torch.float64),
                  "axial": ((B, self.E), torch.float64), "ratio": ((B, self.E), torch.float64),
                  "U": ((B,), torch.float64), "reactions": ((B, self.handle.dims.nres), torch.float64), "status": ((B,), torch.int32)}
        o = capi.GenesOut()
        self.out = {}
        if B == 0:                                           # empty population: nothing to launch
            self.out = {k: torch.empty(shapes[k][0], dtype=shapes[k][1], device=self.device) for k in fields}
            return self.out.get("point")
        for k in fields:
            shape, dt = shapes[k]
            self.out[k] = torch.empty(shape, dtype=dt, device=self.device)
            setattr(o, k, C.c_void_p(self.out[k].data_ptr()))
        with torch.cuda.device(self.device):
            st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            capi.check(capi.lib.tfem_read_genes(self.handle.ptr, B, C.c_void_p(g.data_ptr()), self.max_height,
                                                float(int_obj1), float(int_obj2), C.byref(o), st))
        self._keep = g
        return self.out.get("point")
