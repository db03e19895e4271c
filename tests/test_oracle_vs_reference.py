"""Pins oracle/truss_oracle.py against the UNMODIFIED reference modules (build container only).

Parity definition = SURVEY.md section 8c: post-transition geometry and flags bit-exact, FEM fields at
1e-9 normwise against the FP64-coerced reference solve, float32 observation tensors within 2 ulp.
"""
import re
import os

import numpy as np
import pytest

from oracle import ref_harness
from oracle.truss_oracle import FAMILIES, TrussOracle, build_mesh, pareto_state_data

pytestmark = pytest.mark.reference

RUNS = ["small_bridge", "small_roof", "large_bridge", "large_roof", "train0_roof", "train3_bridge"]   # + two of the ten train/code shapes


def ulp_diff(a, b):
    a = np.asarray(a, dtype=np.float32); b = np.asarray(b, dtype=np.float32)
    ia = a.view(np.int32).astype(np.int64); ib = b.view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, -(ia & 0x7FFFFFFF), ia); ib = np.where(ib < 0, -(ib & 0x7FFFFFFF), ib)
    return np.abs(ia - ib)


def nrm(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


@pytest.fixture(scope="module", params=RUNS)
def pair(request):
    run = request.param
    return run, ref_harness.RefGame(run, fem_fp64=True), TrussOracle(run)


def test_static_tables(pair):
    run, ref, orc = pair
    m, rm = orc.mesh, ref.gen.model
    assert [[e.nodes[0].name - 1, e.nodes[1].name - 1] for e in rm.elements] == m.conn
    assert [n.res for n in rm.nodes] == m.res
    assert [n.top_node for n in rm.nodes] == m.top
    assert [n.vertical_pair[0].name - 1 for n in rm.nodes] == m.pair
    assert [n.coord[0] for n in rm.nodes] == m.x
    assert [int(len(n.loads) != 0) for n in rm.nodes] == m.loaded
    assert [n.has_loady for n in rm.nodes] == m.has_loady
    assert [n.target for n in rm.nodes if n.top_node == 1] == [t for t in m.target_top if t is not None]
    assert rm.tnsc == m.tnsc and rm.ndof == m.ndof
    assert [v[0] for v in rm.jlv] == m.P
    assert ref.gen.max_deformation == m.max_deformation
    assert (ref.gen.y_max, ref.gen.y_min, ref.gen.d_min) == (m.y_max, m.y_min, m.d_min)
    assert float(ref.game.int_obj1) == orc.int_obj1 and float(ref.game.int_obj2) == orc.int_obj2


def test_symmetry_tables_match_source(pair):
    """parse the hard-coded assignments in truss2D_ENV.py and compare with the generated tables"""
    run, ref, orc = pair
    src = open(os.path.join(ref.mods.code_dir, "truss2D_ENV.py")).read()
    if run.startswith("train"):
        assert "# ASSIGN SYMMETRY NODE" not in src and orc.mesh.spec.symmetry == "none"   # train/code's ENV has no symmetry step
        return
    body = src[src.index("# ASSIGN SYMMETRY NODE"):src.index("# Structural analysis")]
    node_part, elem_part = body.split("# ASSIGN SYMMETRY ELEMENT")
    t_part, f_part = node_part.split("else:")
    rx = re.compile(r"nodes\[(\d+)\]\.coord\[1\] = self\.gen_model\.model\.nodes\[(\d+)\]\.coord\[1\]")
    for part, table in ((t_part, orc.mesh.sym_src_true), (f_part, orc.mesh.sym_src_false)):
        want = list(range(orc.mesh.N))
        for dst, s in rx.findall(part):
            want[int(dst)] = int(s)
        assert want == table
    ex = re.compile(r"elements\[(\d+)\]\.section_no = min\(self\.gen_model\.model\.elements\[(\d+)\]\.section_no,"
                    r"self\.gen_model\.model\.elements\[(\d+)\]\.section_no\)")
    pairs = set()
    for dst, a, b in ex.findall(elem_part):
        pairs.add((min(int(a), int(b)), max(int(a), int(b))))
    assert pairs == set((min(a, b), max(a, b)) for a, b in orc.mesh.sym_elem_pairs)


def compare_state(ref, orc, st_ref, out, point_ref=None):
    f = ref.fem_fields()
    # bit-exact: geometry, sections, move range, flags
    assert np.array_equal(f["y"], out["y"]), (f["y"], out["y"])
    assert np.array_equal(f["section"], out["section"])
    assert np.array_equal(f["max_up"], out["max_up"]) and np.array_equal(f["max_down"], out["max_down"])
    assert np.array_equal(f["iscompress"], out["iscompress"])
    weak_ref = np.array([not isinstance(n.coord[1], np.floating) for n in ref.gen.model.nodes])
    assert np.array_equal(weak_ref, out["y_weak"])
    # FP64 fields at 1e-9
    for k in ("d", "axial", "ratio", "length", "reactions"):
        assert nrm(out[k], f[k]) <= 1e-9, k
    assert abs(out["U"] - f["U"]) <= 1e-9 * abs(f["U"])
    x_n, A_n, A_s, A_ts, A_cs, mask = st_ref[0:6]
    raw_n, raw_e, c_e = st_ref[8], st_ref[9], st_ref[10]
    assert np.array_equal(A_n, orc.A_n) and np.array_equal(mask, orc.mask) and np.array_equal(c_e, orc.nC_e)
    for name, a, b in (("x_n", out["x_n"], x_n), ("A_s", out["A_s"], A_s), ("A_n_ts", out["A_n_ts"], A_ts),
                       ("A_n_cs", out["A_n_cs"], A_cs), ("nN_x_n", out["nN_x_n"], raw_n),
                       ("nN_x_e", out["nN_x_e"], raw_e)):
        assert a.shape == b.shape and a.dtype == np.float32
        assert ulp_diff(a, b).max() <= 2, (name, ulp_diff(a, b).max())
    # flag columns exact
    assert np.array_equal(out["x_n"][:, 12] > 0.5, x_n[:, 12] > 0.5)
    assert np.array_equal(out["nN_x_n"][:, 11], raw_n[:, 11])
    assert np.array_equal(out["nN_x_e"][:, [0, 3, 4, 6, 13, 20]], raw_e[:, [0, 3, 4, 6, 13, 20]])
    if point_ref is not None:
        assert ulp_diff(out["point"], np.array(point_ref, dtype=np.float32)).max() <= 2


def test_reset_state(pair):
    run, ref, orc = pair
    ref.gen.re_value_args = None
    with ref.mods.cwd():
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            ref.gen.re_value(*ref.args)
    st = ref.reset_state()
    compare_state(ref, orc, st, orc.reset())


@pytest.mark.parametrize("mode", ["uniform", "small_actions", "saturated"])
def test_random_walk(pair, mode):
    """self-feeding walk + the driver's pattern (three children from one parent, stale move range)"""
    run, ref, orc = pair
    rng = np.random.RandomState({"uniform": 1, "small_actions": 2, "saturated": 3}[mode])
    import contextlib, io
    with ref.mods.cwd(), contextlib.redirect_stdout(io.StringIO()):
        ref.gen.re_value(*ref.args)
    st = ref.reset_state()
    N = orc.mesh.N
    steps = 40 if N == 16 else 20
    parent = st
    for k in range(steps):
        children = []
        for child in range(3):
            if mode == "uniform":
                a_geo = rng.rand(N, 2); a_topo = rng.rand(N, 3)
            elif mode == "small_actions":
                a_geo = rng.rand(N, 2) * 0.2; a_topo = rng.rand(N, 3) * np.array([0.3, 0.3, 1.0])
            else:
                a_geo = rng.randn(N, 2) * 2 + 0.5; a_topo = rng.randn(N, 3) * 2 + 0.5
            a_geo = a_geo.astype(np.float32); a_topo = a_topo.astype(np.float32)
            if mode == "saturated" and k % 5 == 0:
                # a deliberately wrong stale move range: lets the y<y_min / y>y_max passes fire
                ref.set_move_range(rng.rand(N) * 40, rng.rand(N) * 40)
            coin = bool(rng.rand() >= 0.5)
            f = ref.fem_fields()
            a_geo2, a_topo2 = a_geo.copy(), a_topo.copy()
            try:
                point, S = ref.step(parent[-3], parent[-2], parent[-1], a_geo, a_topo, coin)
            except np.linalg.LinAlgError:
                continue
            out = orc.step(parent[-3], parent[-2], f["max_up"], f["max_down"], a_geo2, a_topo2, coin)
            assert np.array_equal(a_geo, a_geo2) and np.array_equal(a_topo, a_topo2)   # in-place clip
            compare_state(ref, orc, S, out, point)
            children.append(S)
        if children:
            parent = children[rng.randint(len(children))]


def test_adversarial_walk(pair):
    """corner cases of the transition: ties, exact bounds, inf / NaN geometry actions, absurd stale ranges"""
    from util import adversarial_actions, adversarial_move_range
    run, ref, orc = pair
    rng = np.random.RandomState(11)
    import contextlib, io
    with ref.mods.cwd(), contextlib.redirect_stdout(io.StringIO()):
        ref.gen.re_value(*ref.args)
    parent = ref.reset_state()
    N = orc.mesh.N
    done = 0
    for k in range(30 if N == 16 else 15):
        a_geo, a_topo = adversarial_actions(rng, N)
        up, down = adversarial_move_range(rng, N)
        ref.set_move_range(up, down)
        coin = bool(rng.rand() >= 0.5)
        f = ref.fem_fields()
        g2, t2 = a_geo.copy(), a_topo.copy()
        try:
            point, S = ref.step(parent[-3], parent[-2], parent[-1], a_geo, a_topo, coin)
        except (np.linalg.LinAlgError, ZeroDivisionError):
            continue
        out = orc.step(parent[-3], parent[-2], f["max_up"], f["max_down"], g2, t2, coin)
        assert np.array_equal(a_geo, g2, equal_nan=True) and np.array_equal(a_topo, t2)
        f2 = ref.fem_fields()
        assert np.array_equal(f2["y"], out["y"]) and np.array_equal(f2["section"], out["section"])
        assert np.array_equal(f2["max_up"], out["max_up"]) and np.array_equal(f2["max_down"], out["max_down"])
        cond = np.linalg.cond(orc.solve_only(out["y"], out["section"])["K"])
        if cond < 1e6:
            assert nrm(out["d"], f2["d"]) <= 1e-9
            for name, a, b in (("x_n", out["x_n"], S[0]), ("nN_x_n", out["nN_x_n"], S[8]), ("nN_x_e", out["nN_x_e"], S[9])):
                assert ulp_diff(a, b).max() <= 2, name
        parent = S
        done += 1
    assert done >= 5


def test_pareto_state_data(pair):
    run, ref, orc = pair
    rng = np.random.RandomState(5)
    for n in (1, 2, 7, 50):
        pf = [[rng.rand(), rng.rand()] for _ in range(n)]
        a, b = ref.mods.ENV.pareto_state_data(pf, n // 2)
        c, d = pareto_state_data(pf, n // 2, max_front=ref.mods.ENV.MAX_FRONT)
        assert np.array_equal(a, c) and np.array_equal(b, d)


def test_kat_example_3_8():
    """Textbook KAT kept as a comment in FEM_2Dtruss.py:474-558, run through the live reference
    classes."""
    mods = ref_harness.RefModules("small_bridge")
    F = mods.FEM
    l1 = F.Load(); l1.set_name(1); l1.set_size(150, -300)
    coords = [(0, 0), (12 * 12, 0), (24 * 12, 0), (12 * 12, 16 * 12)]  # inches (Example 3.8)
    nodes = []
    for i, (x, y) in enumerate(coords):
        n = F.Node(); n.set_name(i + 1); n.set_coord(x, y)
        if i < 3:
            n.set_res(1, 1)
        nodes.append(n)
    nodes[3].set_load(l1)
    model = F.Model(); model.add_load(l1)
    for n in nodes:
        model.add_node(n)
    for i, (a, A) in enumerate(((0, 8), (1, 6), (2, 8))):
        e = F.Element(); e.set_name(i + 1); e.set_nodes(nodes[a], nodes[3]); e.set_em(29000); e.set_area(A)
        model.add_element(e)
    model.restore(); model.gen_all()
    assert np.allclose(np.array(model.ssm), [[696, 0], [0, 2143.5833333333335]], rtol=1e-12)
    assert np.allclose(np.array(model.d).ravel(), [0.21551724137931033, -0.13995257162850366], rtol=1e-12)
    q = [float(e.e_q[0][0]) for e in model.elements]
    assert np.allclose(q, [-16.770011273957138, 126.83201803833144, 233.22998872604282], rtol=1e-12)
