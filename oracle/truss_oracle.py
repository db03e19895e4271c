"""TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the reference truss env-step.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module, and only as the checker / the timed CPU baseline -- never as a product
path (the product is the CUDA library behind ``include/tfem.h`` and fails loudly without it).

Parity status: **pinned**.  ``tests/test_oracle_vs_reference.py`` runs this restatement against the
unmodified reference modules (imported read-only through ``oracle/ref_harness.py``) in the build
container, and ``tests/golden/*.npz`` holds vectors produced by the reference itself
(``tests/golden/make_golden.py``) that travel to the GPU box.

What is restated (reference file:line, canonical copies):
  mesh / supports / loads / targets ... truss2D_GEN.py:181-190, 241-434
  DOF numbering, load vector ........... FEM_2Dtruss.py:207-280
  move-range rule ...................... truss2D_GEN.py:118-133
  action decode, fixes, symmetry ....... test/0{0,2}_*/code/truss2D_ENV.py:361-553 (small) / :361-673 (large)
  element stiffness, assembly, solve ... FEM_2Dtruss.py:284-337
  member forces, energy, reactions ..... FEM_2Dtruss.py:341-431
  observation tensors .................. truss2D_ENV.py:43-196
  objectives / constraint point ........ truss2D_ENV.py:566-589
  reset-time observation ............... truss2D_ENV.py:339-354
  Pareto chain graph ................... truss2D_ENV.py:22-41

Numeric model.  Under NumPy 2 the reference mixes ``np.float32`` scalars (node
heights read back from the float32 state table) with Python ints/floats (constants assigned by the
constraint passes).  Python scalars are "weak": mixed with a float32 they are first rounded to
float32 and the operation is done in float32; two Python scalars operate in float64.  We carry every
height as ``(value, weak)`` and spell that rule out in ``_bin``; float32 operations are evaluated in
float64 and rounded once to float32, which is exact for + - * / (53 >= 2*24+2).  The FEM itself is
evaluated in float64 from ``float(y)`` ("FEM-coerced" mode of ref_harness).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

SECTION_TABLE = np.array(  # section_data/01_brace_rod2.csv : [A cm^2, I cm^4]
    [[9.085, 59.5], [20.41, 300.0], [38.89, 830.0], [81.23, 4230.0], [164.6, 18700.0]], dtype=np.float64)
YOUNG = 2 * 1e11                       # truss2D_GEN.py:59
LONG_STRESS = 235 * 1e6 / 1.5          # FEM_2Dtruss.py:93-94
MAX_FRONT = 50                         # test/*/code/truss2D_ENV.py:15


def f32(x: float) -> float:
    """Round a float64 to the nearest float32 and hand it back as a Python float."""
    return float(np.float32(x))


# ----------------------------------------------------------------------------------------------------
# family description and static tables
# ----------------------------------------------------------------------------------------------------
@dataclass
class FamilySpec:
    name: str
    num_x: int
    span_x: list
    span_y: list
    tar_y: list
    dmin: float
    loadx: float
    loady: float
    truss_type: str            # 'bridge' | 'roof'
    support_case: int = 1
    symmetry: str = "none"     # 'small' (test/00,01) | 'large' (test/02,03) | 'none' (train)


_SMALL_TAR = [4, 3, 2.5, 2, 2, 2.5, 3, 4]
_LARGE_TAR = [3, 2.75, 2.5, 2.25, 2.25, 2, 2, 2, 2, 2, 2, 2.25, 2.25, 2.5, 2.75, 3]
FAMILIES = {
    "small_bridge": FamilySpec("small_bridge", 8, [5] * 7, [8], _SMALL_TAR, 0.3, 0, -75 * 1000, "bridge", 1, "small"),
    "small_roof": FamilySpec("small_roof", 8, [5] * 7, [8], _SMALL_TAR, 0.3, 0, -120 * 1000, "roof", 1, "small"),
    "large_bridge": FamilySpec("large_bridge", 16, [5] * 15, [6], _LARGE_TAR, 0.3, 0, -7.5 * 1000, "bridge", 1, "large"),
    "large_roof": FamilySpec("large_roof", 16, [5] * 15, [6], _LARGE_TAR, 0.3, 0, -8 * 1000, "roof", 1, "large"),
}


# train/code/master_DDPG_truss2D_MO.py:787-795 (trainChoice), :809-815: 6 x 2, mixed spans, no symmetry step in train/code's ENV
_TRAIN_TAR = ([1.0, 1.5, 2.0, 2.0, 1.5, 1.0], [1.0, 3.0, 3.0, 2.0, 1.5, 1.0], [1.0, 1.5, 2.0, 3.0, 3.0, 1.0],
              [1.0, 3.0, 2.0, 2.0, 3.0, 1.0], [3.0, 2.0, 1.0, 1.0, 2.0, 3.0])
for _i, _tar in enumerate(_TRAIN_TAR):
    for _tt in ("roof", "bridge"):
        FAMILIES["train%d_%s" % (_i, _tt)] = FamilySpec("train%d_%s" % (_i, _tt), 6, [4.0, 3.0, 5.0, 3.0, 5.0], [5], _tar, 0.2, 0,
                                                        -100000, _tt, 1, "none")


@dataclass
class Mesh:
    spec: FamilySpec
    N: int
    E: int
    x: list                    # node x (python numbers, as generated)
    y0: list                   # initial node y
    top: list                  # top_node flag
    pair: list                 # vertical pair index
    res: list                  # [rx, ry] per node
    loaded: list               # node carries the Load object
    has_loady: list            # Node.has_loady (overwritten to 1 for loaded nodes)
    target_top: list           # tar_y for top nodes (None for others)
    conn: list                 # [n0, n1] 0-based
    tnsc: list                 # 1-based DOF ids, free first
    ndof: int
    P: list                    # load vector in free-DOF order
    y_max: float = 0.0
    y_min: float = 0.0
    d_min: float = 0.0
    max_deformation: float = 0.0
    sym_src_true: list = field(default_factory=list)    # y[i] <- y[src] when the coin is >= 0.5
    sym_src_false: list = field(default_factory=list)
    sym_elem_pairs: list = field(default_factory=list)


def build_mesh(spec: FamilySpec) -> Mesh:
    nx = spec.num_x
    # gennode (truss2D_GEN.py:181-190): row-major, bottom chord first
    xs = [sum(spec.span_x[:i]) for i in range(nx)]
    ys = [sum(spec.span_y[:i]) for i in range(2)]
    x = [xs[j] for _ in range(2) for j in range(nx)]
    y0 = [ys[i] for i in range(2) for _ in range(nx)]
    N = 2 * nx
    top = [0] * nx + [1] * nx                  # upper node of each vertical (truss2D_GEN.py:307-310)
    pair = [i + nx for i in range(nx)] + [i for i in range(nx)]
    conn = []
    for row in range(2):                       # chords  (:280-295)
        for i in range(nx - 1):
            conn.append([row * nx + i, row * nx + i + 1])
    for i in range(nx):                        # verticals (:299-321)
        conn.append([i, nx + i])
    for i in range(nx - 1):                    # braces '\' (:323-337)
        conn.append([nx + i, i + 1])
    for i in range(nx - 1):                    # braces '/' (:339-353)
        conn.append([i, nx + i + 1])
    E = len(conn)
    # supports (:400-418)
    xset = set(x)
    xc = sorted(xset)
    if spec.support_case == 2:
        xc.remove(max(xc))
    elif spec.support_case == 3:
        xc.remove(min(xc))
    elif spec.support_case == 4:
        xc.remove(min(xc)); xc.remove(max(xc))
    ymin = min(y0)
    res = [[0, 0] for _ in range(N)]
    for i in range(N):
        if y0[i] == ymin and (x[i] == max(xc) or x[i] == min(xc)):
            res[i] = [1, 1]
    # loads (:421-430)
    loaded = [0] * N
    if spec.truss_type == "bridge":
        for i in range(N):
            if y0[i] == ymin and res[i][1] == 0:
                loaded[i] = 1
    elif spec.truss_type == "roof":
        for i in range(N):
            if top[i] == 1:
                loaded[i] = 1
    has_loady = list(loaded)
    target_top = [None] * N
    k = 0
    for i in range(N):
        if top[i] == 1:
            target_top[i] = spec.tar_y[k]; k += 1
    # DOF numbering (FEM_2Dtruss.py:227-261): free first, node-major, x then y
    flat = []
    for i in range(N):
        flat += [res[i][0], res[i][1]]
    ids = [0] * (2 * N)
    c = 1
    for i, r in enumerate(flat):
        if r == 0:
            ids[i] = c; c += 1
    for i, r in enumerate(flat):
        if r == 1:
            ids[i] = c; c += 1
    tnsc = [[ids[2 * i], ids[2 * i + 1]] for i in range(N)]
    ndof = 2 * N - sum(flat)
    P = []                                     # gen_jlv (:264-280)
    for i in range(N):
        for j in range(2):
            if tnsc[i][j] <= ndof:
                P.append((0, spec.loady)[j] if loaded[i] else 0)   # Load.set_size(0, loady), :374
    m = Mesh(spec, N, E, x, y0, top, pair, res, loaded, has_loady, target_top, conn, tnsc, ndof, P)
    m.y_max = spec.span_y[0]
    m.y_min = 0
    m.d_min = spec.dmin
    m.max_deformation = 0.001 * sum(spec.span_x)
    # symmetry tables (truss2D_ENV.py small :460-553 / large :460-673)
    ident = list(range(N))
    if spec.symmetry == "none":
        m.sym_src_true, m.sym_src_false = ident, list(ident)
    else:
        left_from_right, right_from_left = list(ident), list(ident)
        for row in range(2):
            for cidx in range(nx // 2):
                a, b = row * nx + cidx, row * nx + (nx - 1 - cidx)
                if spec.symmetry == "small" and row == 0 and cidx == 0:
                    continue               # the small file does not list the support pair (0,7)
                left_from_right[a] = b
                right_from_left[b] = a
        if spec.symmetry == "small":       # coin True: right -> left
            m.sym_src_true, m.sym_src_false = left_from_right, right_from_left
        else:                              # large: coin True: left -> right
            m.sym_src_true, m.sym_src_false = right_from_left, left_from_right
        nb = nx - 1
        pairs = []
        for row in range(2):
            for kk in range(nb // 2):
                pairs.append((row * nb + kk, row * nb + nb - 1 - kk))
        base = 2 * nb
        for kk in range(nx // 2):
            pairs.append((base + kk, base + nx - 1 - kk))
        b3, b4 = base + nx, base + nx + nb
        for kk in range(nb):
            pairs.append((b3 + kk, b4 + nb - 1 - kk))
        m.sym_elem_pairs = pairs
    return m


# ----------------------------------------------------------------------------------------------------
# weak / float32 scalar arithmetic
# ----------------------------------------------------------------------------------------------------
def _bin(op, a, b):
    """a, b = (value, weak).  NumPy-2 scalar rule: weak op weak -> float64 (weak); otherwise float32."""
    (av, aw), (bv, bw) = a, b
    if aw and bw:
        return (op(av, bv), True)
    return (f32(op(f32(av), f32(bv))), False)


def _add(a, b): return _bin(lambda p, q: p + q, a, b)
def _sub(a, b): return _bin(lambda p, q: p - q, a, b)
def _mul(a, b): return _bin(lambda p, q: p * q, a, b)
def _div(a, b):
    (av, aw), (bv, bw) = a, b
    if aw and bw:
        return (av / bv, True)
    with np.errstate(all="ignore"):
        return (float(np.float32(av) / np.float32(bv)), False)
def _abs(a): return (abs(a[0]), a[1])


def _cmp(a, b):
    """values to compare under the same promotion rule"""
    (av, aw), (bv, bw) = a, b
    if aw and bw:
        return av, bv
    return f32(av), f32(bv)


def _lt(a, b):
    p, q = _cmp(a, b)
    return p < q


def _gt(a, b):
    p, q = _cmp(a, b)
    return p > q


def W(v): return (v, True)       # python scalar
def S(v): return (f32(v), False)  # np.float32 scalar


def round2_f32(v: float) -> float:
    """np.float32.__round__(2): rint(x * 100) / 100 evaluated in float32."""
    return f32(float(np.rint(np.float32(f32(v * 100.0)))) / 100.0)


def np_argmax(vals) -> int:
    """np.argmax on floats: first maximum, a NaN wins and stops the scan."""
    mp, idx = vals[0], 0
    if mp != mp:
        return 0
    for i in range(1, len(vals)):
        v = vals[i]
        if not (v <= mp):
            mp, idx = v, i
            if mp != mp:
                break
    return idx


def pairwise_sum_f32(a) -> float:
    """np.sum of a contiguous float32 vector with n <= 128 (numpy pairwise_sum, 8 accumulators)."""
    a = [float(v) for v in a]
    n = len(a)
    if n < 8:
        r = 0.0
        for v in a:
            r = f32(r + v)
        return r
    assert n <= 128
    r = a[:8]
    i = 8
    while i < n - (n % 8):
        for k in range(8):
            r[k] = f32(r[k] + a[i + k])
        i += 8
    res = f32(f32(f32(r[0] + r[1]) + f32(r[2] + r[3])) + f32(f32(r[4] + r[5]) + f32(r[6] + r[7])))
    while i < n:
        res = f32(res + a[i]); i += 1
    return res


# ----------------------------------------------------------------------------------------------------
# move range (truss2D_GEN.py:118-133)
# ----------------------------------------------------------------------------------------------------
def move_range(m: Mesh, y):
    """y: list of (value, weak).  Returns (max_up, max_down) as typed scalars."""
    up, down = [None] * m.N, [None] * m.N
    for i in range(m.N):
        yp = y[m.pair[i]]
        if m.top[i] == 1:
            up[i] = _abs(_sub(W(m.y_max), y[i]))
            down[i] = _abs(_sub(_sub(y[i], yp), W(m.d_min)))
        elif m.spec.truss_type == "bridge":
            up[i], down[i] = W(0), W(0)
        elif m.spec.truss_type == "roof":
            up[i] = _abs(_sub(_sub(yp, y[i]), W(m.d_min)))
            down[i] = _abs(_sub(y[i], W(m.y_min)))
    return up, down


# ----------------------------------------------------------------------------------------------------
# transition (truss2D_ENV.py:373-553)
# ----------------------------------------------------------------------------------------------------
def clip_actions(a):
    """in-place clip to [0,1] as written at truss2D_ENV.py:379-391 (NaN passes through)."""
    for i in range(a.shape[0]):
        for j in range(a.shape[1]):
            if a[i, j] > 1:
                a[i, j] = 1
            elif a[i, j] < 0:
                a[i, j] = 0


def transition(m: Mesh, y_tab, sec_tab, max_up32, max_down32, a_geo, a_topo, coin: bool):
    """y_tab: float32 heights from the state table; sec_tab: ints; max_up32/max_down32: the STALE move
    range (float32 values).  Returns (y typed list, section list)."""
    N, E = m.N, m.E
    y = [S(float(v)) for v in y_tab]
    sec = [int(s) for s in sec_tab]
    for i in range(N):
        row = [float(a_geo[i, 0]), float(a_geo[i, 1])]
        adj = np_argmax(row)
        v = row[adj]
        amt = S(v) if v < 1 else W(1)            # min([1, a])
        if adj == 0:
            step = _mul(_mul(amt, S(float(max_up32[i]))), W(0.25))
            y[i] = _add(y[i], step)
        elif adj == 1:
            step = _mul(_mul(amt, S(float(max_down32[i]))), W(0.25))
            y[i] = _sub(y[i], step)
    for i in range(N):
        if m.res[i][1] == 1:
            y[i] = W(0)
        if not y[i][1]:
            y[i] = (round2_f32(y[i][0]), False)
        else:
            y[i] = (round(y[i][0], 2), True)
    for e in range(E):
        n0, n1 = m.conn[e]
        pv = [f32(float(a_topo[n0, k]) + float(a_topo[n1, k])) for k in range(3)]
        am = np_argmax(pv)
        if am == 0:
            sec[e] = max(0, sec[e] - 1)
        elif am == 1:
            sec[e] = min(len(SECTION_TABLE) - 1, sec[e] + 1)
    ymin, ymax, dmin = W(m.y_min), W(m.y_max), W(m.d_min)
    for i in range(N):                           # pass (i)
        if _lt(y[i], ymin):
            if m.top[i] == 1:
                y[i] = dmin
                y[m.pair[i]] = ymin
            else:
                y[i] = ymin
    for i in range(N):                           # pass (ii)
        if _gt(y[i], ymax):
            if m.top[i] == 1:
                y[i] = ymax
            else:
                y[i] = _sub(ymax, dmin)
                y[m.pair[i]] = ymax
    for i in range(N):                           # pass (iii)
        if _lt(_abs(_sub(y[i], y[m.pair[i]])), dmin):
            if m.top[i] == 1:
                y[i] = _add(y[m.pair[i]], dmin)
    src = m.sym_src_true if coin else m.sym_src_false
    y = [y[src[i]] for i in range(N)]
    for a, b in m.sym_elem_pairs:
        s = min(sec[a], sec[b])
        sec[a] = s; sec[b] = s
    return y, sec


# ----------------------------------------------------------------------------------------------------
# FEM (FEM_2Dtruss.py:284-431), float64
# ----------------------------------------------------------------------------------------------------
class SingularStiffness(Exception):
    pass


def fem_solve(m: Mesh, y64, sec):
    N, E, n = m.N, m.E, m.ndof
    K = np.zeros((n, n))
    geo = []
    for e in range(E):
        n0, n1 = m.conn[e]
        dx = m.x[n1] - m.x[n0]
        dy = y64[n1] - y64[n0]
        L = (dx ** 2 + dy ** 2) ** 0.5
        c, s = dx / L, dy / L
        A = SECTION_TABLE[sec[e]][0] * 1e-4
        k = YOUNG * A / L
        geo.append((L, c, s, A, k))
        kl = np.array([[k, 0, -k, 0], [0, 0, 0, 0], [-k, 0, k, 0], [0, 0, 0, 0]])
        T = np.array([[c, s, 0, 0], [-s, c, 0, 0], [0, 0, c, s], [0, 0, -s, c]])
        kg = (T.T.dot(kl)).dot(T)
        ids = [m.tnsc[n0][0], m.tnsc[n0][1], m.tnsc[n1][0], m.tnsc[n1][1]]
        for p in range(4):
            for q in range(4):
                if ids[p] <= n and ids[q] <= n:
                    K[ids[p] - 1][ids[q] - 1] += kg[p][q]
    P = np.array(m.P, dtype=np.float64).reshape(n, 1)
    try:
        with np.errstate(all="ignore"):
            d = np.linalg.solve(K, P)
    except np.linalg.LinAlgError as exc:          # the reference process dies here (FEM_2Dtruss.py:337)
        raise SingularStiffness(str(exc))
    d = d.reshape(-1)
    node_d = np.zeros((N, 2))
    for i in range(N):
        for j in range(2):
            if m.tnsc[i][j] <= n:
                node_d[i, j] = d[m.tnsc[i][j] - 1]
    dcol = d.reshape(n, 1)
    U = float(np.dot(np.dot(dcol.transpose(), K), dcol)[0, 0]) * 0.5      # gen_U_full (:374-379)
    axial = np.zeros(E); ratio = np.zeros(E); length = np.zeros(E)
    iscomp = np.zeros(E, dtype=np.int32)
    react = np.zeros(2 * N)
    for e in range(E):
        n0, n1 = m.conn[e]
        L, c, s, A, k = geo[e]
        u0 = c * node_d[n0, 0] + s * node_d[n0, 1]
        u2 = c * node_d[n1, 0] + s * node_d[n1, 1]
        q0 = k * u0 + (-k) * u2
        q = [q0, 0.0, -k * u0 + k * u2, 0.0]
        f = [c * q[0], s * q[0], c * q[2], s * q[2]]
        ids = [m.tnsc[n0][0], m.tnsc[n0][1], m.tnsc[n1][0], m.tnsc[n1][1]]
        for p in range(4):
            if ids[p] > n:
                react[ids[p] - 1] += f[p]
        axial[e] = q0
        length[e] = L
        ratio[e] = abs(q0 / A) / LONG_STRESS
        iscomp[e] = 0 if q0 <= 0 else 1
    return {"d": d, "node_d": node_d, "axial": axial, "ratio": ratio, "iscompress": iscomp,
            "length": length, "U": U, "reactions": react[n:], "K": K}


# ----------------------------------------------------------------------------------------------------
# observations (truss2D_ENV.py:43-196) and objectives (:566-589)
# ----------------------------------------------------------------------------------------------------
def normalized_adjacency(m: Mesh):
    """A_n = D^-1/2 (A + I) D^-1/2 in float32, constant per topology (truss2D_ENV.py:104-110)."""
    A = np.zeros((m.N, m.N), dtype=np.float32)
    for n0, n1 in m.conn:
        A[n0, n1] = 1; A[n1, n0] = 1
    mask = A.copy()
    A = A + np.eye(m.N, dtype=np.float32)
    with np.errstate(divide="ignore"):
        deg = np.power(np.array(A.sum(1)), -1 / 2).ravel()
    deg[np.isinf(deg)] = 0.0
    D = np.diag(deg)
    return np.matmul(D, np.matmul(A, D)), mask


def incidence(m: Mesh):
    c = np.zeros((m.E, m.N), dtype=np.float32)
    for e, (n0, n1) in enumerate(m.conn):
        c[e, n0] = 1; c[e, n1] = 1
    return c


def _node_common(m: Mesh, y, up, down, fem, i):
    """the first eleven entries shared by x_n and nN_x_n, as float32 values"""
    row = [0.0] * 11
    row[0] = f32(m.x[i])
    row[1] = f32(y[i][0])
    row[2] = float(m.res[i][0]); row[3] = float(m.res[i][1])
    row[4] = float(abs(m.has_loady[i]))
    row[5] = float(m.top[i]); row[6] = float(abs(m.top[i] - 1))
    row[7] = f32(up[i][0]); row[8] = f32(down[i][0])
    if m.top[i] == 1:     # target * top_node / (y + 1e-6); non-top nodes multiply by top_node == 0
        row[9] = f32(_div(W(m.target_top[i] * 1), _add(y[i], W(1e-6)))[0])
    row[10] = f32(abs(float(fem["node_d"][i, 1])))
    return row


def observations(m: Mesh, y, sec, up, down, fem):
    N, E = m.N, m.E
    maxdef32 = f32(m.max_deformation)
    x_n = np.zeros((N, 13), dtype=np.float32)
    raw_n = np.zeros((N, 12), dtype=np.float32)
    for i in range(N):
        row = _node_common(m, y, up, down, fem, i)
        r = f32(row[10] / maxdef32)
        x_n[i, :11] = row
        x_n[i, 11] = f32(r * 0.5) if not (r > 1) else 1.0
        x_n[i, 12] = float(int(r > 1))
        raw_n[i, :11] = row
        raw_n[i, 11] = float(int(r >= 1))
    mn, mx = x_n.min(axis=0), x_n.max(axis=0)
    x_n = (x_n - mn) / (mx - mn + 1e-6)
    A_s = np.zeros((N, N), dtype=np.float32)
    A_ts = np.zeros((N, N), dtype=np.float32)
    A_cs = np.zeros((N, N), dtype=np.float32)
    raw_e = np.zeros((E, 21), dtype=np.float32)
    amax = SECTION_TABLE[-1][0] * 1e-4
    for e in range(E):
        n0, n1 = m.conn[e]
        A = SECTION_TABLE[sec[e]][0] * 1e-4
        A_s[n0, n1] = A_s[n1, n0] = A / amax
        py = float(fem["ratio"][e])
        val = min([py, 1]) * (1 if py > 1 else 0.5)
        if fem["iscompress"][e] == 0:
            A_ts[n0, n1] = A_ts[n1, n0] = val
        else:
            A_cs[n0, n1] = A_cs[n1, n0] = val
        raw_e[e, 0] = sec[e]
        raw_e[e, 1] = A
        raw_e[e, 2] = fem["length"][e]
        raw_e[e, 3] = abs(int(fem["iscompress"][e]) - 1)
        raw_e[e, 4] = fem["iscompress"][e]
        raw_e[e, 5] = fem["axial"][e]
        raw_e[e, 6] = int(py > 1)
        for k, nd in ((7, n0), (14, n1)):
            raw_e[e, k + 0] = raw_n[nd, 0]
            raw_e[e, k + 1] = raw_n[nd, 1]
            raw_e[e, k + 2] = raw_n[nd, 2]
            raw_e[e, k + 3] = raw_n[nd, 3]
            raw_e[e, k + 4] = raw_n[nd, 4]
            raw_e[e, k + 5] = raw_n[nd, 10]
            raw_e[e, k + 6] = raw_n[nd, 11]
    return {"x_n": x_n, "A_s": A_s, "A_n_ts": A_ts, "A_n_cs": A_cs, "nN_x_n": raw_n, "nN_x_e": raw_e}


def initial_objectives(m: Mesh):
    """int_obj1 / int_obj2 of Game_research04.__init__ (truss2D_ENV.py:267-277)."""
    y = [W(v) for v in m.y0]
    fem = fem_solve(m, [float(v) for v in m.y0], [len(SECTION_TABLE) - 1] * m.E)
    all_v = [f32(SECTION_TABLE[-1][0] * 1e-4 * fem["length"][e]) for e in range(m.E)]
    all_dt = [f32(_abs(_sub(W(m.target_top[i]), y[i]))[0]) if m.top[i] == 1 else 0.0 for i in range(m.N)]
    return pairwise_sum_f32(all_v), pairwise_sum_f32(all_dt)


def objectives(m: Mesh, y, sec, fem, int_obj1, int_obj2):
    all_s = [f32(fem["ratio"][e]) for e in range(m.E)]
    all_v = [f32(SECTION_TABLE[sec[e]][0] * 1e-4 * fem["length"][e]) for e in range(m.E)]
    all_d = [0.0] * m.N
    all_dt = [0.0] * m.N
    for i in range(m.N):
        if m.top[i] == 1:
            all_dt[i] = f32(_abs(_sub(W(m.target_top[i]), y[i]))[0])
        else:
            all_d[i] = f32(float(fem["node_d"][i, 1]) / m.max_deformation)
    obj1 = pairwise_sum_f32(all_v)
    obj2 = pairwise_sum_f32(all_dt)
    con1 = max(abs(v) for v in all_s)
    con2 = max(abs(v) for v in all_d)
    point = np.array([f32(obj1 / int_obj1), f32(obj2 / int_obj2), con1, con2], dtype=np.float32)
    # FP64 companions (no float32 rounding anywhere) for the 1e-9 check
    v64 = sum(SECTION_TABLE[sec[e]][0] * 1e-4 * fem["length"][e] for e in range(m.E))
    dt64 = sum(abs(m.target_top[i] - y[i][0]) for i in range(m.N) if m.top[i] == 1)
    c1 = max(abs(fem["ratio"][e]) for e in range(m.E))
    c2 = max(abs(float(fem["node_d"][i, 1])) / m.max_deformation for i in range(m.N) if m.top[i] == 0)
    return point, np.array([v64, dt64, c1, c2], dtype=np.float64)


# ----------------------------------------------------------------------------------------------------
# public oracle entry points
# ----------------------------------------------------------------------------------------------------
class TrussOracle:
    """One family; every method works on ONE environment (loop over the batch in the caller)."""

    def __init__(self, family):
        self.spec = FAMILIES[family] if isinstance(family, str) else family
        self.mesh = build_mesh(self.spec)
        self.A_n, self.mask = normalized_adjacency(self.mesh)
        self.nC_e = incidence(self.mesh)
        self.int_obj1, self.int_obj2 = initial_objectives(self.mesh)

    def _evaluate(self, y, sec):
        m = self.mesh
        up, down = move_range(m, y)
        fem = fem_solve(m, [float(v[0]) for v in y], sec)
        obs = observations(m, y, sec, up, down, fem)
        point, point64 = objectives(m, y, sec, fem, self.int_obj1, self.int_obj2)
        out = dict(obs)
        out.update(point=point, point64=point64, d=fem["d"], axial=fem["axial"], ratio=fem["ratio"],
                   iscompress=fem["iscompress"], length=fem["length"], U=fem["U"],
                   reactions=fem["reactions"],
                   y=np.array([v[0] for v in y], dtype=np.float64),
                   y_weak=np.array([v[1] for v in y], dtype=np.bool_),
                   section=np.array(sec, dtype=np.int32),
                   max_up=np.array([f32(v[0]) for v in up], dtype=np.float32),
                   max_down=np.array([f32(v[0]) for v in down], dtype=np.float32))
        return out

    def reset(self):
        """_game_get_1_state on the generated geometry (truss2D_ENV.py:339-354)."""
        m = self.mesh
        y = [W(v) for v in m.y0]
        sec = [len(SECTION_TABLE) - 1] * m.E
        return self._evaluate(y, sec)

    def step(self, nN_x_n, nN_x_e, max_up32, max_down32, a_geo, a_topo, coin):
        """_game_modify (truss2D_ENV.py:373-589).  a_geo / a_topo are clipped IN PLACE like the
        reference does."""
        clip_actions(a_geo); clip_actions(a_topo)
        y_tab = np.asarray(nN_x_n, dtype=np.float32)[:, 1]
        sec_tab = [int(v) for v in np.asarray(nN_x_e)[:, 0]]
        y, sec = transition(self.mesh, y_tab, sec_tab, max_up32, max_down32, a_geo, a_topo, bool(coin))
        return self._evaluate(y, sec)

    def solve_only(self, y64, sec):
        return fem_solve(self.mesh, [float(v) for v in y64], [int(s) for s in sec])


def pareto_state_data(pf, index=0, max_front=MAX_FRONT):
    """truss2D_ENV.py:22-41 -- chain graph over the current front (MAX_FRONT: 50 in test/*/code, 20 in train/code)."""
    n = len(pf)
    x_pf = np.zeros((n, 4), dtype=np.float32)
    for i in range(n):
        x_pf[i, 0] = pf[i][0]; x_pf[i, 1] = pf[i][1]
        if i == index:
            x_pf[i, 2] = 1
        x_pf[i, 3] = n / max_front
    A = np.eye(n, dtype=np.float32)
    for i in range(n - 1):
        A[i, i + 1] = 1; A[i + 1, i] = 1
    with np.errstate(divide="ignore"):
        deg = np.power(np.array(A.sum(1)), -1 / 2).ravel()
    deg[np.isinf(deg)] = 0.0
    D = np.diag(deg)
    return x_pf, np.matmul(D, np.matmul(A, D))
