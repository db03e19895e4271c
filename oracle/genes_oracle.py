"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the gene-vector objective of the MOEA/D benchmark.

Follows ``gen_model.read_genes`` of ``test/benchmarks/MOEAD/<family>.zip:<family>/truss2D_GEN.py:117-228`` (called once
per individual from ``MOEAD_master.py:103-121``) on top of the pinned FEM / objective
restatement of ``oracle/truss_oracle.py``.  Genes are float64, so every height is a Python float and the whole path
is float64 (the float32 casts of the objective arrays ``all_s / all_v / all_d / all_dt`` excepted).

Pinned: ``tests/test_oracle_vs_reference.py::test_genes_oracle_vs_reference_moead`` runs it against the unmodified
``truss2D_GEN.py`` + ``FEM_2Dtruss.py`` extracted from the four zips (``oracle/ref_harness.RefMoead``), and
``tests/golden/genes.npz`` holds outputs recorded from those reference files (``tests/golden/make_genes_golden.py``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this module.
"""
from __future__ import annotations

import numpy as np

from .truss_oracle import SECTION_TABLE, W, TrussOracle, fem_solve, objectives

# ``max_height = 8 # change this`` (small zips, :127) / ``= 6`` (large zips)
MAX_HEIGHT = {"small_bridge": 8, "small_roof": 8, "large_bridge": 6, "large_roof": 6}


def py_max(a, b):
    """``max([a, b])``: the first maximal element"""
    return b if b > a else a


def decode_genes(o: TrussOracle, genes, max_height):
    """genes[N+E] -> (y[N] python floats, section[E] ints): truss2D_GEN.py:126-195 of the zip."""
    m = o.mesh
    N, E, nx = m.N, m.E, m.spec.num_x
    h = [float(g) * max_height for g in genes[:N]]                       # :126-128
    sec = [min([len(SECTION_TABLE) - 1, round(float(g) * 4)]) for g in genes[N:N + E]]   # :130-132 (half-to-even)
    y = [float(v) for v in m.y0]                                         # the persistent model returns to this
    if m.spec.truss_type == "roof":                                      # :135-142
        for i in range(N):
            if m.res[i][1] == 0:
                y[i] = py_max(h[i], m.d_min)
        y[N - 1] = 0                                                     # for-else: runs after the loop, i = N-1
    else:                                                                # :144-149
        for i in range(N):
            if m.top[i] == 1:
                y[i] = py_max(h[i], m.d_min)
    sec = [int(s) for s in sec]                                          # :151-152
    for i in range(N):                                                   # :155-159 fix the vertical pair
        if m.top[i] == 1 and y[i] - m.d_min < y[m.pair[i]]:
            y[m.pair[i]] = y[i] - m.d_min
    for i in range(N):                                                   # :162-167 fix heights below y_min
        if m.top[i] == 0 and y[i] < m.y_min:
            y[i] = m.y_min
            y[m.pair[i]] = m.d_min
    # forced symmetry: heights left -> right (:169-176 small, :53-71 large), sections low index <- partner
    src = m.sym_src_false if nx == 8 else m.sym_src_true
    y = [y[src[i]] for i in range(N)]
    for a, b in m.sym_elem_pairs:
        sec[min(a, b)] = sec[max(a, b)]
    return y, sec


def read_genes(o: TrussOracle, genes, max_height=None, int_obj1=None, int_obj2=None):
    """-> dict(point float32[4], point64, y, section, d, axial, ratio, U)"""
    if max_height is None:
        max_height = MAX_HEIGHT[o.spec.name]
    y, sec = decode_genes(o, genes, max_height)
    fem = fem_solve(o.mesh, [float(v) for v in y], sec)
    point, point64 = objectives(o.mesh, [W(v) for v in y], sec, fem,
                                o.int_obj1 if int_obj1 is None else int_obj1,
                                o.int_obj2 if int_obj2 is None else int_obj2)
    return dict(point=point, point64=point64, y=np.array(y, dtype=np.float64), section=np.array(sec, dtype=np.int32),
                d=fem["d"], axial=fem["axial"], ratio=fem["ratio"], U=fem["U"])
