"""GPU: tfem_read_genes (one launch of the env-step kernel in gene mode) against the golden vectors recorded from the
reference's MOEA/D zips, against the oracle on random populations, and -- at the benchmark's population size --
through size-independent properties."""
import numpy as np
import pytest
import torch

from oracle import genes_oracle
from oracle.truss_oracle import TrussOracle
from util import FAMILY_NAMES, FP64_TOL, assert_f32_close, load_golden, nrm

pytestmark = pytest.mark.gpu


def fp64_tol(o, y, sec):
    """1e-9 (north_star) for cond(K) <= 1e6; beyond that (gene vectors with heights scaled to the d_min floor give
    0.3 m deep trusses, cond up to ~1e8) the reference's own LU answer is only good to eps*cond, so the tolerance grows
    linearly with cond -- the rule of tests/test_gpu_parity.py::fp64_tol"""
    cond = float(np.linalg.cond(o.solve_only(y, sec)["K"]))
    return FP64_TOL * max(1.0, cond / 1e6)


ALL = ("point", "point64", "y", "section", "d", "axial", "ratio", "U", "reactions", "status")


@pytest.fixture(scope="module")
def genesmod():
    from mop_truss_marl_b200 import genes
    return genes


@pytest.mark.parametrize("name", FAMILY_NAMES)
def test_read_genes_vs_golden(genesmod, name):
    g = load_golden("genes")
    ev = genesmod.GeneEvaluator(name)
    o = TrussOracle(name)
    point = ev.read_genes(g[name + "_genes"], fields=ALL)
    torch.cuda.synchronize()
    out = {k: v.cpu().numpy() for k, v in ev.out.items()}
    assert int(np.abs(out["status"]).max()) == 0
    assert np.array_equal(out["y"], g[name + "_y"])                      # float64 bit patterns
    assert np.array_equal(out["section"], g[name + "_section"])
    for t in range(out["d"].shape[0]):
        tol = fp64_tol(o, out["y"][t], out["section"][t])
        for k in ("d", "axial", "ratio"):
            assert nrm(out[k][t], g["%s_%s" % (name, k)][t]) <= tol, (t, k, tol)
    assert_f32_close("point", point.cpu().numpy(), g[name + "_point"])   # <= 2 ulp(float32)
    assert ev.handle.launch_count() == 1


@pytest.mark.parametrize("name", ["small_bridge", "large_roof"])
def test_read_genes_vs_oracle_random(genesmod, name):
    o = TrussOracle(name)
    rng = np.random.RandomState(11)
    B = 96
    genes = rng.rand(B, o.mesh.N + o.mesh.E)
    genes[::3, :o.mesh.N] *= 0.05
    genes[1::4, o.mesh.N:] = np.round(genes[1::4, o.mesh.N:] * 8) / 8
    ev = genesmod.GeneEvaluator(name)
    ev.read_genes(genes, fields=ALL)
    torch.cuda.synchronize()
    out = {k: v.cpu().numpy() for k, v in ev.out.items()}
    for b in range(B):
        want = genes_oracle.read_genes(o, genes[b])
        assert np.array_equal(out["y"][b], want["y"]) and np.array_equal(out["section"][b], want["section"])
        tol = fp64_tol(o, want["y"], want["section"])
        for k in ("d", "axial", "ratio"):
            assert nrm(out[k][b], want[k]) <= tol, (b, k, tol)
        assert abs(out["U"][b] - want["U"]) <= tol * abs(want["U"])
        assert_f32_close("point", out["point"][b], want["point"])
        assert nrm(out["point64"][b], want["point64"]) <= tol


def test_read_genes_custom_normalisers_and_errors(genesmod):
    ev = genesmod.GeneEvaluator("small_bridge")
    g = np.random.RandomState(0).rand(8, ev.N + ev.E)
    p0 = ev.read_genes(g).cpu().numpy()
    p1 = ev.read_genes(g, int_obj1=2.0, int_obj2=4.0).cpu().numpy()
    tab = ev.handle.table("int_obj")
    assert np.allclose(p1[:, 0] * 2.0, p0[:, 0] * tab[0], rtol=1e-6) and np.allclose(p1[:, 1] * 4.0, p0[:, 1] * tab[1], rtol=1e-6)
    assert np.array_equal(p0[:, 2:], p1[:, 2:])
    with pytest.raises(ValueError):
        ev.read_genes(np.zeros((2, 5)))
    launches = ev.handle.launch_count()
    empty = ev.read_genes(np.zeros((0, ev.N + ev.E)))                  # empty population
    assert tuple(empty.shape) == (0, 4) and ev.handle.launch_count() == launches
    one = ev.read_genes(g[3])                                          # a single individual, 1-D
    assert np.array_equal(one.cpu().numpy()[0], p0[3])
    odd = ev.read_genes(g[:7]).cpu().numpy()                           # ragged against the 8 warps of a CTA
    assert np.array_equal(odd, p0[:7])


@pytest.mark.parametrize("name,B", [("small_bridge", 18000), ("large_bridge", 12000)])
def test_population_size_properties(genesmod, name, B):
    """MOEAD_master.py evaluates (n_neighbors + 1) * n_iteration = 18 000 / 12 000 individuals per seed: one launch"""
    ev = genesmod.GeneEvaluator(name)
    N, E, nx = ev.N, ev.E, ev.N // 2
    gen = torch.Generator(device="cuda").manual_seed(3)
    genes = torch.rand(B, N + E, dtype=torch.float64, device="cuda", generator=gen)
    ev.read_genes(genes, fields=ALL)
    torch.cuda.synchronize()
    o = ev.out
    assert int(o["status"].abs().max()) == 0
    # (1) forced symmetry: mirrored heights; every section in 0..4
    y = o["y"]
    assert torch.equal(y[:, nx:], y[:, nx:].flip(1)) and torch.equal(y[:, :nx], y[:, :nx].flip(1))
    assert int(o["section"].min()) >= 0 and int(o["section"].max()) <= 4
    # (2) strain energy = half the work of the loads; reactions balance the load
    P = torch.from_numpy(ev.handle.table("loadvec")).cuda()
    work = 0.5 * (o["d"] @ P)
    rel = (o["U"] - work).abs() / work.abs()
    deep = (y[:, nx:] - y[:, :nx]).min(dim=1).values >= 1.0        # at least 1 m deep everywhere: cond(K) < 1e6
    assert int(deep.sum()) > B // 20
    assert float(rel[deep].max()) <= 1e-9 and float(rel.max()) <= 1e-6   # 0.3 m deep individuals: eps * cond
    tnsc = ev.handle.table("tnsc"); ndof = ev.ndof
    ry = sum(o["reactions"][:, tnsc[i, 1] - 1 - ndof] for i in range(N) if tnsc[i, 1] > ndof)
    bal = (ry + P.sum()).abs() / P.sum().abs()
    assert float(bal[deep].max()) <= 1e-9 and float(bal.max()) <= 1e-6
    # (3) idempotence: the plain solve of the decoded geometry gives the same bits; so does the other half alone
    from mop_truss_marl_b200 import capi
    import ctypes as C
    d2 = torch.empty_like(o["d"]); ax2 = torch.empty_like(o["axial"])
    p = lambda t: C.c_void_p(t.data_ptr())   # noqa: E731
    capi.check(capi.lib.tfem_solve_only(ev.handle.ptr, B, p(y), p(o["section"]), p(d2), p(ax2), None, None, None, None,
                                        C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    assert torch.equal(d2, o["d"]) and torch.equal(ax2, o["axial"])
    keep_point, keep_d = o["point"].clone(), o["d"].clone()
    ev.read_genes(genes[B // 2:], fields=("point", "d"))
    torch.cuda.synchronize()
    assert torch.equal(ev.out["point"], keep_point[B // 2:]) and torch.equal(ev.out["d"], keep_d[B // 2:])
    # (4) feasibility predicate of MyProblem._evaluate is well defined (no NaN)
    assert bool(torch.isfinite(keep_point).all())
