"""CPU: the oracle restatement reproduces the golden vectors recorded from the reference itself."""
import numpy as np
import pytest

from oracle.truss_oracle import TrussOracle
from util import FAMILY_NAMES, F32_FIELDS, FP64_TOL, assert_f32_close, load_golden, nrm


@pytest.fixture(scope="module", params=FAMILY_NAMES)
def fam(request):
    return request.param, load_golden(request.param), TrussOracle(request.param)


def test_tables(fam):
    name, g, o = fam
    m = o.mesh
    assert np.array_equal(g["conn"], np.array(m.conn)) and np.array_equal(g["tnsc"], np.array(m.tnsc))
    assert int(g["ndof"]) == m.ndof
    assert np.array_equal(g["res"], np.array(m.res)) and np.array_equal(g["top"], np.array(m.top))
    assert np.array_equal(g["pair"], np.array(m.pair)) and np.array_equal(g["loaded"], np.array(m.loaded))
    assert np.array_equal(g["loadvec"], np.array(m.P, dtype=np.float64))
    assert np.array_equal(g["A_n"], o.A_n) and np.array_equal(g["mask"], o.mask) and np.array_equal(g["nC_e"], o.nC_e)
    assert np.array_equal(g["sym_src"][0], np.array(m.sym_src_false)) and np.array_equal(g["sym_src"][1], np.array(m.sym_src_true))
    assert set(map(tuple, g["sym_elem_pairs"])) == set((min(a, b), max(a, b)) for a, b in m.sym_elem_pairs)
    assert float(g["int_obj"][0]) == o.int_obj1 and float(g["int_obj"][1]) == o.int_obj2


def check(out, g, prefix, i=None):
    pick = (lambda k: g[prefix + k]) if i is None else (lambda k: g[prefix + k][i])
    assert np.array_equal(out["y"], pick("y")) and np.array_equal(out["section"], pick("section"))
    assert np.array_equal(out["y_weak"], pick("y_weak"))
    assert np.array_equal(out["max_up"], pick("max_up")) and np.array_equal(out["max_down"], pick("max_down"))
    assert np.array_equal(out["iscompress"], pick("iscompress"))
    for k in ("d", "axial", "ratio", "length", "reactions"):
        assert nrm(out[k], pick(k)) <= FP64_TOL, k
    assert abs(out["U"] - float(pick("U"))) <= FP64_TOL * abs(float(pick("U")))
    for k in F32_FIELDS:
        assert_f32_close(k, out[k], pick(k))


def test_reset(fam):
    name, g, o = fam
    check(o.reset(), g, "reset_")


def test_transitions(fam):
    name, g, o = fam
    T = g["tr_mode"].shape[0]
    for i in range(T):
        a_geo, a_topo = g["tr_in_a_geo"][i].copy(), g["tr_in_a_topo"][i].copy()
        out = o.step(g["tr_in_set_node"][i], g["tr_in_set_element"][i], g["tr_in_max_up"][i], g["tr_in_max_down"][i],
                     a_geo, a_topo, bool(g["tr_in_coin"][i]))
        assert np.array_equal(a_geo, g["tr_out_a_geo"][i]) and np.array_equal(a_topo, g["tr_out_a_topo"][i])
        check(out, g, "tr_out_", i)
        assert_f32_close("point", out["point"], g["tr_out_point"][i])


def test_golden_scalars():
    """initial-geometry scalars"""
    want = {"small_bridge": 445.21800128, "small_roof": 1169.36743409, "large_bridge": 223.90537182,
            "large_roof": 251.17625713}
    for name, U in want.items():
        g = load_golden(name)
        assert abs(float(g["reset_U"]) - U) < 1e-7
    g = load_golden("small_bridge")
    assert np.allclose(g["reset_d"][:4], [-1.5869e-4, -1.25840e-3, -1.6942e-4, -2.09653e-3], rtol=2e-4)
    assert abs(g["reset_ratio"].max() - 0.10914515210085823) < 1e-15
