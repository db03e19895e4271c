"""TEST INFRASTRUCTURE ONLY -- the truss solve of ``FEM_2Dtruss.py:164-337`` (``gen_all`` up to ``gen_d``) in EXTENDED
precision, batched: assembly, Cholesky-free LDL^T and one step of iterative refinement in ``np.longdouble`` (x87 80-bit,
eps = 1.1e-19 on this image's hosts).  It is the yardstick for the ill-conditioned geometries (d_min-deep trusses, cond(K)
up to 2e8): there the reference's own float64 LU answer carries an error of about eps * cond, so "within 1e-9 of the
reference" cannot be asked of any other float64 solver -- what can be asked is that it is as close to the exact
displacements as the reference is.  ``tests/test_exact_fem.py`` (CPU: reference / oracle vs exact),
``tests/test_gpu_parity.py::test_ill_conditioned_solves_vs_extended_precision`` (GPU vs exact)."""
from __future__ import annotations

import numpy as np

from .truss_oracle import SECTION_TABLE, YOUNG, Mesh

LD = np.longdouble


def assemble(m: Mesh, y, sec, dtype=LD):
    """K [M,n,n], P [M,n] for M geometries: y [M,N] float64 heights, sec [M,E] section numbers"""
    y = np.asarray(y, dtype=dtype)
    sec = np.asarray(sec)
    M, n = y.shape[0], m.ndof
    K = np.zeros((M, n, n), dtype=dtype)
    x = np.asarray(m.x, dtype=dtype)
    area = np.asarray(SECTION_TABLE[:, 0], dtype=dtype) * dtype(1e-4)
    for e in range(m.E):
        n0, n1 = m.conn[e]
        dx = x[n1] - x[n0]
        dy = y[:, n1] - y[:, n0]
        L = np.sqrt(dx * dx + dy * dy)
        c, s = dx / L, dy / L
        k = dtype(YOUNG) * area[sec[:, e]] / L
        ids = [m.tnsc[n0][0], m.tnsc[n0][1], m.tnsc[n1][0], m.tnsc[n1][1]]
        g = [c, s, -c, -s]                                   # k_global = k * g g^T
        for p in range(4):
            for q in range(4):
                if ids[p] <= n and ids[q] <= n:
                    K[:, ids[p] - 1, ids[q] - 1] += k * g[p] * g[q]
    P = np.broadcast_to(np.asarray(m.P, dtype=dtype).reshape(1, n), (M, n)).copy()
    return K, P


def ldl_solve(K, P):
    """solves K d = P for a batch of SPD matrices by LDL^T without pivoting, in K's dtype"""
    A = K.copy()
    M, n, _ = A.shape
    b = P.copy()
    for j in range(n):
        piv = A[:, j, j].copy()
        if j + 1 < n:
            l = A[:, j + 1:, j] / piv[:, None]               # [M, n-j-1]
            A[:, j + 1:, j + 1:] -= l[:, :, None] * A[:, j, j + 1:][:, None, :]
            b[:, j + 1:] -= l * b[:, j][:, None]
            A[:, j + 1:, j] = l
    d = np.zeros_like(b)
    for j in range(n - 1, -1, -1):
        acc = b[:, j].copy()
        if j + 1 < n:
            acc -= np.einsum("mk,mk->m", A[:, j, j + 1:], d[:, j + 1:])
        d[:, j] = acc / A[:, j, j]
    return d


def exact_displacements(m: Mesh, y, sec):
    """d [M,n] in longdouble: LDL^T + one refinement step with the residual formed in longdouble"""
    K, P = assemble(m, y, sec)
    d = ldl_solve(K, P)
    r = P - np.einsum("mij,mj->mi", K, d)
    return d + ldl_solve(K, r)


def cond2(m: Mesh, y, sec):
    """2-norm condition numbers of the float64 stiffness matrices"""
    K, _ = assemble(m, y, sec, dtype=np.float64)
    w = np.linalg.eigvalsh(K)
    return w[:, -1] / w[:, 0]


def rel_err(d, d_exact):
    """max |d - d_exact| / max |d_exact| per geometry (the parity tests' normwise measure)"""
    d_exact = np.asarray(d_exact, dtype=LD)
    return np.asarray(np.abs(np.asarray(d, dtype=LD) - d_exact).max(axis=1) / np.abs(d_exact).max(axis=1), dtype=np.float64)
