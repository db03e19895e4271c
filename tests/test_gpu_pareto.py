"""GPU: tpareto_front_hv against the golden vectors recorded from the reference's utils.py and against the CPU oracle
on random batches (front membership and order exact, statistics and hypervolume to 1e-12 in float64)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(points64, counts, ref):
    from mop_truss_marl_b200 import pareto
    pts = torch.from_numpy(np.nan_to_num(points64, nan=0.0).astype(np.float32)).cuda()
    out = pareto.front_hv(pts, torch.from_numpy(counts.astype(np.int32)).cuda(), ref)
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in out.items()}, pts.cpu().numpy().astype(np.float64)


def check_against_oracle(out, pts32, counts, ref):
    from oracle import pareto_oracle
    for b in range(len(counts)):
        p = pts32[b, :counts[b]]
        try:
            idx = pareto_oracle.front_indices(p)
        except IndexError:
            assert out["front_len"][b] == 0 and out["hv"][b] == 0.0
            assert np.array_equal(out["stats"][b], [0.0, 1.0, 0.0, 0.0, 1.0])
            continue
        F = len(idx)
        assert out["front_len"][b] == F
        assert np.array_equal(out["front_idx"][b, :F], idx) and (out["front_idx"][b, F:] == -1).all()
        _, max_d, dis_d, p_cd, sum_d, std_cd = pareto_oracle.front_stats(p)
        assert np.allclose(out["stats"][b], [max_d, dis_d, p_cd, sum_d, std_cd], rtol=1e-12, atol=1e-14), b
        assert abs(out["hv"][b] - pareto_oracle.hypervolume(p[idx], ref)) <= 1e-12, b


def test_golden_from_reference():
    g = np.load(os.path.join(ROOT, "tests", "golden", "pareto.npz"))
    for r, ref in enumerate(g["refs"]):
        out, pts32 = run(g["points"], g["counts"], ref)
        # the kernel sees the float32-rounded points: same front (the golden sets have no near-ties), statistics within
        # float32 rounding of the inputs
        assert np.array_equal(out["front_len"], g["front_len"])
        for b in range(len(g["counts"])):
            F = g["front_len"][b]
            got = pts32[b][out["front_idx"][b, :F]][:, :2]
            assert np.allclose(got, g["fronts"][b, :F], rtol=0, atol=1e-7)
        assert np.allclose(out["stats"], g["stats"], rtol=1e-5, atol=1e-6)
        assert np.allclose(out["hv"], g["hv"][:, r], rtol=0, atol=1e-6)
        check_against_oracle(out, pts32, g["counts"], ref)          # and exactly what the oracle gives on those inputs


@pytest.mark.parametrize("B,P", [(1, 1), (257, 50), (4096, 64), (33, 7), (300, 200), (64, 256)])
def test_random_batches_vs_oracle(B, P):
    rng = np.random.RandomState(B + P)
    pts = rng.rand(B, P, 4)
    pts[:, :, 2:] *= 1.4
    counts = rng.randint(1, P + 1, size=B)
    if B > 8:
        pts[3, :, 2] = 5.0                                    # an environment with no feasible point
        pts[5, 1] = pts[5, 0]                                 # an exact duplicate
        pts[6, :, :2] = np.round(pts[6, :, :2], 1)            # many ties in both objectives
        counts[3:7] = P
    ref = (0.95, 0.85)
    out, pts32 = run(pts, counts, ref)
    step = max(1, B // 200)
    sel = np.unique(np.concatenate([np.arange(0, B, step), np.arange(min(B, 8))]))
    sub = {k: v[sel] for k, v in out.items()}
    check_against_oracle(sub, pts32[sel], counts[sel], ref)


@pytest.mark.parametrize("rows,max_front", [(50, 50), (17, 50), (1, 50), (20, 20)])
def test_state_data_matches_oracle_bit_for_bit(rows, max_front):
    """tpareto_state_data = pareto_state_data (truss2D_ENV.py:22-41) + the driver's zero padding / cut to `rows` rows
    (master_DDPG_truss2D_MO.py:499-517): float32 outputs bit-identical to the oracle (itself pinned against the reference in
    tests/test_oracle_vs_reference.py::test_pareto_state_data), with and without the front_idx indirection of front_hv"""
    from mop_truss_marl_b200 import pareto
    from oracle.truss_oracle import pareto_state_data
    rng = np.random.RandomState(rows)
    B, P = 97, 64
    pts = rng.rand(B, P, 4).astype(np.float32)
    lens = rng.randint(1, P + 1, size=B).astype(np.int32)
    lens[:4] = (1, 2, rows, P)
    index = np.array([rng.randint(0, n) for n in lens], dtype=np.int32)
    perm = np.stack([rng.permutation(P) for _ in range(B)]).astype(np.int32)
    for use_idx in (False, True):
        x_p, A_p = pareto.state_data(torch.from_numpy(pts).cuda(), torch.from_numpy(perm).cuda() if use_idx else None,
                                     torch.from_numpy(lens).cuda(), torch.from_numpy(index).cuda(), rows=rows, max_front=max_front)
        torch.cuda.synchronize()
        x_p, A_p = x_p.cpu().numpy(), A_p.cpu().numpy()
        for b in range(B):
            n = int(lens[b])
            src = perm[b, :n] if use_idx else np.arange(n)
            x, A = pareto_state_data([(pts[b, s, 0], pts[b, s, 1]) for s in src], index=int(index[b]), max_front=max_front)
            wx, wA = np.zeros((rows, 4), np.float32), np.zeros((rows, rows), np.float32)
            m = min(n, rows)
            wx[:m], wA[:m, :m] = x[:m], A[:m, :m]
            assert np.array_equal(x_p[b].view(np.uint32), wx.view(np.uint32)), (b, n)
            assert np.array_equal(A_p[b].view(np.uint32), wA.view(np.uint32)), (b, n)


def test_large_fronts_are_thinned_with_the_callers_draw():
    """the step-end cull of up to 50 + 150 accumulated candidates (master_DDPG_truss2D_MO.py:437): fronts of more than
    MAX_FRONT = 50 members are thinned like utils.simple_cull does (:104-131) with the draw given as an input; order of the
    thinned front exact, statistics and hypervolume of the thinned list to 1e-12 against the oracle (itself pinned against
    the reference with the same draw, tests/test_pareto_oracle.py)"""
    from mop_truss_marl_b200 import pareto
    from oracle import pareto_oracle
    rng = np.random.RandomState(9)
    B, P = 48, 200
    pts = np.zeros((B, P, 4), np.float32)
    counts = rng.randint(120, P + 1, size=B).astype(np.int32)
    for b in range(B):
        n = counts[b]
        t = np.sort(rng.rand(n))
        curve = 1.0 - t ** (0.5 + rng.rand()) + (0.002 if b % 3 else 0.2) * rng.rand(n)     # every third batch: a short front
        p = np.stack([t, curve, rng.rand(n) * 0.9, rng.rand(n) * 0.9], axis=1)
        pts[b, :n] = p[rng.permutation(n)]
    picks = np.zeros((B, 48), np.int32)
    lens = []
    for b in range(B):
        F = len(pareto_oracle.front_indices(pts[b, :counts[b]].astype(np.float64)))
        lens.append(F)
        picks[b] = rng.permutation(max(F - 2, 48))[:48] if F > 50 else 0
    assert sum(F > 50 for F in lens) >= 10 and sum(F <= 50 for F in lens) >= 5
    ref = (0.9, 0.95)
    out = pareto.front_hv(torch.from_numpy(pts).cuda(), torch.from_numpy(counts).cuda(), ref, thin_pick=torch.from_numpy(picks).cuda())
    torch.cuda.synchronize()
    out = {k: v.cpu().numpy() for k, v in out.items()}
    for b in range(B):
        p64 = pts[b, :counts[b]].astype(np.float64)
        f, max_d, dis_d, p_cd, sum_d, std_cd = pareto_oracle.front_stats(p64, thin_pick=picks[b] if lens[b] > 50 else None)
        F = len(f)
        assert F == min(lens[b], 50) and out["front_len"][b] == F
        assert np.array_equal(p64[out["front_idx"][b, :F]], f) and (out["front_idx"][b, F:] == -1).all()
        assert np.allclose(out["stats"][b], [max_d, dis_d, p_cd, sum_d, std_cd], rtol=1e-12, atol=1e-14), b
        assert abs(out["hv"][b] - pareto_oracle.hypervolume(f, ref)) <= 1e-12, b
    # without a draw a large front is returned whole
    whole = pareto.front_hv(torch.from_numpy(pts).cuda(), torch.from_numpy(counts).cuda(), ref)
    assert np.array_equal(whole["front_len"].cpu().numpy(), np.array(lens))
