"""CPU: oracle/pareto_oracle.py (restatement of utils.simple_cull / union_rectangles_fastest) against the golden
vectors recorded from the reference and, when the reference tree is present, against the imported reference itself."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import pareto_oracle, ref_harness          # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "pareto.npz")


def test_oracle_matches_golden():
    g = np.load(GOLD)
    for k in range(len(g["counts"])):
        pts = g["points"][k, :g["counts"][k]]
        front, max_d, dis_d, p_cd, sum_d, std_cd = pareto_oracle.front_stats(pts)
        F = int(g["front_len"][k])
        assert len(front) == F
        assert np.array_equal(front[:, :2], g["fronts"][k, :F])
        assert np.allclose([max_d, dis_d, p_cd, sum_d, std_cd], g["stats"][k], rtol=1e-12, atol=1e-14)
        for r, ref in enumerate(g["refs"]):
            assert abs(pareto_oracle.hypervolume(front, ref) - g["hv"][k, r]) <= 1e-12


def test_edge_cases():
    # single point (1,1): the reference short-circuits to 0 (utils.py:476)
    assert pareto_oracle.hypervolume([[1.0, 1.0, 0, 0]]) == 0.0
    assert pareto_oracle.hypervolume([]) == 0.0
    # one feasible point: no distances (dis 1, max 0, sum 0), std_cd 1
    f, max_d, dis_d, p_cd, sum_d, std_cd = pareto_oracle.front_stats([[0.4, 0.6, 0.5, 0.5], [0.1, 0.1, 2.0, 0.0]])
    assert len(f) == 1 and (max_d, dis_d, p_cd, sum_d, std_cd) == (0.0, 1.0, 0.0, 0.0, 1.0)
    assert abs(pareto_oracle.hypervolume(f) - 0.6 * 0.4) < 1e-15
    with pytest.raises(IndexError):                       # all infeasible: the reference indexes an empty list
        pareto_oracle.front_stats([[0.1, 0.1, 2.0, 0.0]])
    # points beyond 1 are clamped for the area, not for the reference-point strip
    assert abs(pareto_oracle.hypervolume([[1.5, 0.5, 0, 0]], (1, 1))) < 1e-15


@pytest.mark.skipif(not ref_harness.available(), reason="reference tree not present")
def test_oracle_matches_reference_utils():
    utils = ref_harness.load_utils("small_bridge")
    rng = np.random.RandomState(3)
    for trial in range(150):
        n = rng.randint(1, 51)
        pts = rng.rand(n, 4)
        pts[:, 2:] *= 1.3
        pts[rng.randint(n), 2:] = 0.2
        ref = [float(0.7 + 0.3 * rng.rand()), float(0.7 + 0.3 * rng.rand())]
        want = utils.simple_cull([list(map(float, p)) for p in pts])
        got = pareto_oracle.front_stats(pts)
        assert np.array_equal(got[0], np.array(want[0]))
        assert np.allclose(got[1:], want[1:], rtol=1e-12, atol=1e-14)
        hv_want = utils.union_rectangles_fastest([list(f) for f in want[0]], +1, -1, ref_point=ref)
        assert abs(pareto_oracle.hypervolume(got[0], ref) - hv_want) <= 1e-12
        # the hypervolume of an arbitrary (dominated) list too
        hv_all = utils.union_rectangles_fastest([list(map(float, p)) for p in pts], +1, -1, ref_point=ref)
        assert abs(pareto_oracle.hypervolume(pts, ref) - hv_all) <= 1e-12


@pytest.mark.skipif(not ref_harness.available(), reason="reference tree not present")
def test_thinning_of_large_fronts_matches_reference():
    """fronts of more than MAX_FRONT = 50 members (the step-end cull of up to 50 + 150 accumulated candidates,
    master_DDPG_truss2D_MO.py:437): the reference's random.sample is replaced by an explicit draw on both sides"""
    utils = ref_harness.load_utils("small_bridge")
    rng = np.random.RandomState(5)
    done = 0
    for trial in range(40):
        n = rng.randint(120, 201)
        t = np.sort(rng.rand(n))
        pts = np.stack([t, 1.0 - t ** (0.5 + rng.rand()) + 0.002 * rng.rand(n), rng.rand(n) * 0.9, rng.rand(n) * 0.9], axis=1)
        pts = pts[rng.permutation(n)]
        F = len(pareto_oracle.front_indices(pts))
        if F <= 50:
            continue
        pick = [int(k) for k in rng.permutation(F - 2)[:48]]
        orig = utils.random.sample
        utils.random.sample = lambda pop, k: [pop[i] for i in pick[:k]]
        try:
            want = utils.simple_cull([list(map(float, p)) for p in pts])
        finally:
            utils.random.sample = orig
        got = pareto_oracle.front_stats(pts, thin_pick=pick)
        assert len(want[0]) == 50 and np.array_equal(got[0], np.array(want[0]))
        assert np.allclose(got[1:], want[1:], rtol=1e-12, atol=1e-14)
        ref = [0.9, 0.95]
        hv_want = utils.union_rectangles_fastest([list(f) for f in want[0]], +1, -1, ref_point=ref)
        assert abs(pareto_oracle.hypervolume(got[0], ref) - hv_want) <= 1e-12
        done += 1
    assert done >= 10
    with pytest.raises(ValueError):
        pareto_oracle.front_stats(pts)                    # a large front needs the draw
