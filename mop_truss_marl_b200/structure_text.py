"""The reference's structure text format, written inside its hot loop (``master_DDPG_truss2D_MO.py:215-220`` ->
``gen_model.savetxt``, ``truss2D_GEN.py:193-211``) and re-read by the renderers (``gen_model.read_src``,
``test/*/render/truss2D_READ.py:136-172``).  Host logic only (no device code): batches of structures are formatted /
parsed here and evaluated through ``tfem_solve_only``.

One file = one structure, ``\\r\\n`` line ends, every line starts with one blank:

    `` {name}, [fx, fy]``                                  per load   (``Load.__repr__``, FEM_2Dtruss.py:24-25)
    `` {name}, [x, y], [rx, ry], [[{load}], ...]``         per node   (``Node.__repr__``, FEM_2Dtruss.py:75-76)
    `` {name},{n0},{n1},{E},{A},[[I]]``                    per element

Numbers are written the way the reference's pinned NumPy (1.23.5, ``environments/requirements.txt:52``) prints them:
python ints / floats by ``repr``, ``np.float32`` heights by their shortest round-trip form (``3.2``).  NumPy >= 2 would
print ``np.float32(3.2)``, which ``read_src``'s ``ast.literal_eval`` cannot parse -- this writer never emits that form, and
the reader accepts it anyway.
"""
from __future__ import annotations

import ast
import os
import re
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .families import FAMILIES, YOUNG, FamilySpec

_NP_SCALAR = re.compile(r"np\.float(?:16|32|64)\(([^()]*)\)")


def _num(v) -> str:
    """repr of one scalar as NumPy 1.23 / plain Python print it"""
    if isinstance(v, np.floating):
        return str(v)                       # shortest round-trip form of the scalar's own precision
    if isinstance(v, np.integer):
        return str(int(v))
    return repr(v)


def _mesh(spec: FamilySpec):
    nx = spec.num_x
    xs = [sum(spec.span_x[:i]) for i in range(nx)]
    x = xs + xs
    conn = []
    for row in range(2):
        conn += [(row * nx + i, row * nx + i + 1) for i in range(nx - 1)]
    conn += [(i, nx + i) for i in range(nx)]
    conn += [(nx + i, i + 1) for i in range(nx - 1)]
    conn += [(i, nx + i + 1) for i in range(nx - 1)]
    res = [[0, 0] for _ in range(2 * nx)]
    sc = spec.support_case
    left, right = (0 if sc in (1, 2) else 1), (nx - 1 if sc in (1, 3) else nx - 2)
    for i in (left, right):
        res[i] = [1, 1]
    if spec.truss_type == "bridge":
        loaded = [1 if (i < nx and res[i][1] == 0) else 0 for i in range(2 * nx)]
    else:
        loaded = [0] * nx + [1] * nx
    return x, conn, res, loaded


def format_structure(spec, y, section, y_is_float32=None) -> str:
    """text of one structure.  ``y``: N heights -- python numbers / numpy scalars are printed by their own type; a
    float64 array is printed as float32 where ``y_is_float32[i]`` (the ``y`` / ``y_weak`` pair of ``tfem_step_out``:
    ``y_is_float32 = ~y_weak``) and as python int / float elsewhere.  ``section``: E section numbers."""
    spec = FAMILIES[spec] if isinstance(spec, str) else spec
    x, conn, res, loaded = _mesh(spec)
    N, E = len(x), len(conn)
    if len(y) != N or len(section) != E:
        raise ValueError("expected %d heights and %d sections" % (N, E))
    load = "1, [%s, %s]" % (_num(0), _num(spec.loady))
    out = [" " + load]
    for i in range(N):
        yi = y[i]
        if y_is_float32 is not None:
            yi = np.float32(yi) if y_is_float32[i] else (int(yi) if float(yi).is_integer() else float(yi))
        loads = "[[%s]]" % load if loaded[i] else "[]"
        out.append(" %d, [%s, %s], [%d, %d], %s" % (i + 1, _num(x[i]), _num(yi), res[i][0], res[i][1], loads))
    for e, (a, b) in enumerate(conn):
        s = int(section[e])
        area = spec.section_area_cm2[s] * 1e-4
        inertia = spec.section_inertia_cm4[s] * 1e-8
        out.append(" %d,%d,%d,%s,%s,[[%s]]" % (e + 1, a + 1, b + 1, _num(YOUNG), _num(area), _num(inertia)))
    return "\r\n".join(out) + "\r\n"


def parse_structure(text: str, spec=None):
    """``read_src`` without a model: -> dict(loads {name: [fx, fy]}, nodes {name: (x, y, rx, ry)},
    elements {name: (n0, n1, E, A, I)}) and, when ``spec`` is given, ``y`` [N] float64 and ``section`` [E] int32
    (section k where ``A == truss[k][0] * 1e-4`` exactly, like the reference; -1 if no catalogue entry matches)."""
    loads, nodes, elements = {}, {}, {}
    for raw in text.splitlines():
        line = raw.strip()
        if not line:
            continue
        val = ast.literal_eval(_NP_SCALAR.sub(r"\1", line.replace(" ", "")))
        if len(val) == 2:
            loads[val[0]] = [val[1][0], val[1][1]]
        elif len(val) == 4:
            nodes[val[0]] = (val[1][0], val[1][1], val[2][0], val[2][1])
        elif len(val) == 6:
            elements[val[0]] = (val[1], val[2], val[3], val[4], val[5][0][0])
        else:
            raise ValueError("not a structure line: %r" % raw)          # the reference prints 'ERROR' and stops
    out = {"loads": loads, "nodes": nodes, "elements": elements}
    if spec is not None:
        spec = FAMILIES[spec] if isinstance(spec, str) else spec
        N, E = spec.N, spec.E
        out["y"] = np.array([float(nodes[i + 1][1]) for i in range(N)], dtype=np.float64)
        areas = [a * 1e-4 for a in spec.section_area_cm2]
        out["section"] = np.array([areas.index(elements[e + 1][3]) if elements[e + 1][3] in areas else -1
                                   for e in range(E)], dtype=np.int32)
    return out


def write_batch(paths, spec, y, section, y_is_float32=None, workers: int = 8):
    """one file per structure of the batch (``y`` [B,N], ``section`` [B,E], host arrays), written by a thread pool
    so a driver can keep stepping while the files land; returns the futures' results (paths)"""
    y, section = np.asarray(y), np.asarray(section)

    def one(i):
        txt = format_structure(spec, y[i], section[i], None if y_is_float32 is None else y_is_float32[i])
        os.makedirs(os.path.dirname(os.path.abspath(paths[i])), exist_ok=True)
        with open(paths[i], "w", newline="") as f:
            f.write(txt)
        return paths[i]
    with ThreadPoolExecutor(max_workers=max(1, workers)) as ex:
        return list(ex.map(one, range(len(paths))))


def read_batch(paths, spec):
    """-> (y [B,N] float64, section [B,E] int32): the inputs of ``tfem_solve_only`` / ``BatchedTrussEnv.solve_only``"""
    ys, secs = [], []
    for p in paths:
        with open(p, newline="") as f:
            d = parse_structure(f.read(), spec)
        ys.append(d["y"]); secs.append(d["section"])
    return np.stack(ys), np.stack(secs)
