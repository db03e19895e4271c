"""GPU: the drop-in modules (mop_truss_marl_b200/compat) driven exactly like the reference driver drives the
reference modules, against the golden transitions recorded from the reference."""
import io
import contextlib
import os
import random

import numpy as np
import pytest

from util import FAMILY_NAMES, assert_f32_close, load_golden, nrm

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

ARGS = {
    "small_bridge": (8, 2, [5] * 7, [8], [4, 3, 2.5, 2, 2, 2.5, 3, 4], 0.3, 0, -75 * 1000, "bridge", 1, None),
    "small_roof": (8, 2, [5] * 7, [8], [4, 3, 2.5, 2, 2, 2.5, 3, 4], 0.3, 0, -120 * 1000, "roof", 1, None),
    "large_bridge": (16, 2, [5] * 15, [6], [3, 2.75, 2.5, 2.25, 2.25, 2, 2, 2, 2, 2, 2, 2.25, 2.25, 2.5, 2.75, 3], 0.3, 0,
                     -7.5 * 1000, "bridge", 1, None),
    "large_roof": (16, 2, [5] * 15, [6], [3, 2.75, 2.5, 2.25, 2.25, 2, 2, 2, 2, 2, 2, 2.25, 2.25, 2.5, 2.75, 3], 0.3, 0,
                   -8 * 1000, "roof", 1, None),
}


@pytest.fixture(scope="module")
def mods():
    from mop_truss_marl_b200 import compat
    compat.install()
    import truss2D_GEN, truss2D_ENV, truss2D_RL, FEM_2Dtruss   # noqa: E401
    return truss2D_GEN, truss2D_ENV, truss2D_RL, FEM_2Dtruss


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


@pytest.mark.parametrize("name", FAMILY_NAMES)
def test_dropin_game_matches_golden(mods, name, tmp_path):
    GEN, ENV, RL, FEM = mods
    g = load_golden(name)
    gm = quiet(GEN.gen_model, *ARGS[name])
    game = quiet(ENV.Game_research04, 500, gm, 3)
    env = ENV.ENV(game)
    env.reset()
    assert float(game.int_obj1) == float(g["int_obj"][0]) and float(game.int_obj2) == float(g["int_obj"][1])
    # star-import surface the driver relies on
    assert GEN.os is os and hasattr(GEN, "gen_model") and hasattr(ENV, "pareto_state_data")
    st = game._game_get_1_state()
    assert len(st) == 11
    for k, idx in (("x_n", 0), ("A_s", 2), ("A_n_ts", 3), ("A_n_cs", 4), ("nN_x_n", 8), ("nN_x_e", 9)):
        assert_f32_close("reset " + k, st[idx], g["reset_" + k])
    assert np.array_equal(st[1], g["A_n"]) and np.array_equal(st[5], g["mask"]) and np.array_equal(st[10], g["nC_e"])
    assert st[6].tolist() == [[1, 1, 1, np.float32(1 / 50)]] and st[7].tolist() == [[1]]
    m = gm.model
    assert nrm(np.array(m.d).ravel(), g["reset_d"]) <= 1e-9 and abs(float(m.U_full[0]) - float(g["reset_U"])) <= 1e-9 * float(g["reset_U"])
    assert [type(n.coord[1]) for n in m.nodes] == [int] * len(m.nodes)
    # replay the golden transitions through _game_modify
    T = g["tr_mode"].shape[0]
    orig = random.random
    try:
        for i in range(T):
            for n, u, d in zip(m.nodes, g["tr_in_max_up"][i], g["tr_in_max_down"][i]):
                n.max_up, n.max_down = np.float32(u), np.float32(d)
            random.random = (lambda: 0.75) if g["tr_in_coin"][i] else (lambda: 0.25)
            a_geo, a_topo = g["tr_in_a_geo"][i].copy(), g["tr_in_a_topo"][i].copy()
            point, S = game._game_modify(g["tr_in_set_node"][i], g["tr_in_set_element"][i], g["nC_e"], [a_geo, a_topo])
            assert np.array_equal(a_geo, g["tr_out_a_geo"][i]) and np.array_equal(a_topo, g["tr_out_a_topo"][i])
            assert S[6] is None and S[7] is None and len(S) == 11
            assert_f32_close("point", np.array(point, dtype=np.float32), g["tr_out_point"][i])
            for k, idx in (("x_n", 0), ("A_s", 2), ("A_n_ts", 3), ("A_n_cs", 4), ("nN_x_n", 8), ("nN_x_e", 9)):
                assert_f32_close("tr %d %s" % (i, k), S[idx], g["tr_out_" + k][i])
            # the object model the driver / savetxt look at
            assert [float(n.coord[1]) for n in m.nodes] == g["tr_out_y"][i].tolist()
            assert [not isinstance(n.coord[1], np.floating) for n in m.nodes] == g["tr_out_y_weak"][i].tolist()
            assert [e.section_no for e in m.elements] == g["tr_out_section"][i].tolist()
            assert [e.iscompress for e in m.elements] == g["tr_out_iscompress"][i].tolist()
            assert np.array_equal(np.array([np.float32(n.max_up) for n in m.nodes]), g["tr_out_max_up"][i])
            assert np.array_equal(np.array([np.float32(n.max_down) for n in m.nodes]), g["tr_out_max_down"][i])
            if i % 20 == 0:
                path = os.path.join(tmp_path, "s.txt")
                gm.savetxt(path)
                lines = open(path, newline="").read().split("\r\n")
                assert len(lines) == 1 + len(m.nodes) + len(m.elements) + 1
                assert lines[0].startswith(" 1, [0, ")
    finally:
        random.random = orig
    game.step()
    assert game.game_step == 2
    game.done_counter = 1
    env.check_over()
    assert env.over == 1


def test_dropin_actor_act(mods):
    GEN, ENV, RL, FEM = mods
    mu = [[0.1, 0.1], [0.1, 0.1, 0.1]]
    maddpg = RL.MADDPG(1e-7, 1, 0.95, 0.99, 200, 200, 1000, 3, [2, 3], mu, mu, mu)
    assert len(maddpg.agents) == 3
    gm = quiet(GEN.gen_model, *ARGS["small_bridge"])
    game = quiet(ENV.Game_research04, 500, gm, 3)
    st = game._game_get_1_state()
    np.random.seed(7)
    geo, topo = maddpg.agents[0].act(st[0], st[1], st[2], st[3], st[4], st[6], st[7])
    assert geo.shape == (16, 2) and topo.shape == (16, 3) and geo.dtype == np.float32
    # same seed -> the same NumPy noise stream (80 randn draws, geo first, row-major)
    from oracle.actor_oracle import actor_forward
    g64, t64 = actor_forward(maddpg.agents[0].actor_model.weights, st[0][None], st[1][None], st[2][None], st[3][None],
                             st[4][None], st[6][None], st[7][None])
    np.random.seed(7)
    noise = np.array([np.random.randn(1)[0] for _ in range(80)])
    want_geo = g64[0] + 0.1 * (0.1 - g64[0]) * 1e-4 + 0.1 * noise[:32].reshape(16, 2)
    want_topo = t64[0] + 0.1 * (0.1 - t64[0]) * 1e-4 + 0.1 * noise[32:].reshape(16, 3)
    assert np.abs(geo - want_geo).max() < 5e-5 and np.abs(topo - want_topo).max() < 5e-5
    assert maddpg.agents[0].update_num == 1
    point, S = game._game_modify(st[8], st[9], st[10], [geo, topo])
    assert len(point) == 4 and geo.max() <= 1 and geo.min() >= 0            # clipped in place
    assert maddpg.agents[1].critic_model.load_weights("nowhere").expect_partial() is not None
    # the learner side of the drop-in: remember / train / update as the driver calls them (master...:612-649)
    maddpg.train()                                           # fewer than 32 transitions: returns silently (:491-494)
    state8 = list(st[:8])                                    # (x_n, A_n, A_s, A_n_ts, A_n_cs, mask, x_pf, A_pf)
    child8 = list(S[:5]) + [st[5], st[6], st[7]]
    for i in range(33):
        acts = []
        for k in range(3):
            acts += [np.random.rand(16, 2).astype(np.float32), np.random.rand(16, 3).astype(np.float32)]
        maddpg.remember(state8, *acts, [0.1 * i, -0.2, 0.3], child8, child8, child8, 1 if i == 5 else 0, 16)
    w_before = maddpg.agents[0].actor_model.weights["gcn_l3_1"][0].copy()
    geo_b, _ = maddpg.agents[0].act(st[0], st[1], st[2], st[3], st[4], st[6], st[7])
    maddpg.train()
    w_after = maddpg.agents[0].actor_model.weights["gcn_l3_1"][0]
    # Adam's first step at lr * 0.1 = 1e-8 (:625): one float32 ulp of a 0.07-sized weight is 7.5e-9, so a weight moves
    # by 0, 1 or 2 ulps
    assert 0 < np.abs(w_after - w_before).max() <= 2e-8
    maddpg.update()
    geo_a, _ = maddpg.agents[0].act(st[0], st[1], st[2], st[3], st[4], st[6], st[7])
    assert geo_a.shape == geo_b.shape


@pytest.mark.parametrize("name", ["small_roof", "large_bridge"])
def test_dropin_read_genes_matches_golden(mods, name):
    """gen_model.read_genes (MOEA/D zips, truss2D_GEN.py:117-228) as MOEAD_master.py:104 calls it"""
    GEN = mods[0]
    g = load_golden("genes")
    gm = quiet(GEN.gen_model, *ARGS[name])
    int_obj1, int_obj2 = g[name + "_int_obj"]
    for t in (0, 2, 3, 4, 7):               # well-conditioned vectors (cond(K) < 1e6)
        point = gm.read_genes(g[name + "_genes"][t], int_obj1, int_obj2)
        assert_f32_close("point", np.array(point, dtype=np.float32), g[name + "_point"][t])
        m = gm.model
        assert np.array_equal(np.array([float(n.coord[1]) for n in m.nodes]), g[name + "_y"][t])
        assert [e.section_no for e in m.elements] == g[name + "_section"][t].tolist()
        assert nrm(np.array(m.d).ravel(), g[name + "_d"][t]) <= 1e-9
        assert nrm(np.array([e.prop_yeield for e in m.elements]), g[name + "_ratio"][t]) <= 1e-9
        assert m.elements[0].area == gm.truss[m.elements[0].section_no][0] * 1e-4


def test_dropin_savetxt_read_src_round_trip(mods, tmp_path):
    """savetxt writes the reference's bytes (tests/golden/structure_text.npz, reset state) and read_src brings a saved
    structure back onto a fresh model, which then analyses to the same displacements"""
    GEN = mods[0]
    g = load_golden("structure_text")
    gm = quiet(GEN.gen_model, *ARGS["small_roof"])
    p0 = os.path.join(tmp_path, "reset.txt")
    gm.savetxt(p0)
    assert open(p0, newline="").read() == str(g["small_roof_text"][0])
    p1 = os.path.join(tmp_path, "stepped.txt")
    with open(p1, "w", newline="") as f:
        f.write(str(g["small_roof_text"][2]))
    gm.read_src(p1)
    m = gm.model
    assert np.array_equal(np.array([float(n.coord[1]) for n in m.nodes], dtype=np.float32),
                          g["small_roof_y"][2].astype(np.float32))
    assert [e.section_no for e in m.elements] == g["small_roof_section"][2].tolist()
    m.restore(); m.gen_all()
    d1 = np.array(m.d).ravel().copy()
    p2 = os.path.join(tmp_path, "again.txt")
    gm.savetxt(p2)
    gm2 = quiet(GEN.gen_model, *ARGS["small_roof"])
    gm2.read_src(p2)
    gm2.model.restore(); gm2.model.gen_all()
    assert np.array_equal(np.array(gm2.model.d).ravel(), d1)
