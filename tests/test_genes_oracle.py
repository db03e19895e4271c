"""CPU: the gene-vector oracle (oracle/genes_oracle.py) against the golden vectors recorded from the reference's
MOEA/D benchmark zips, and -- in the build container -- against the unmodified reference files themselves."""
import numpy as np
import pytest

from oracle import genes_oracle, ref_harness
from oracle.truss_oracle import TrussOracle
from util import FAMILY_NAMES, FP64_TOL, load_golden, nrm


@pytest.fixture(scope="module")
def golden():
    return load_golden("genes")


@pytest.mark.parametrize("name", FAMILY_NAMES)
def test_genes_oracle_vs_golden(golden, name):
    o = TrussOracle(name)
    assert float(golden[name + "_int_obj"][0]) == o.int_obj1 and float(golden[name + "_int_obj"][1]) == o.int_obj2
    genes = golden[name + "_genes"]
    assert genes.shape[1] == o.mesh.N + o.mesh.E
    for t, g in enumerate(genes):
        out = genes_oracle.read_genes(o, g)
        assert np.array_equal(out["y"], golden[name + "_y"][t]), t                 # float64 bit patterns
        assert np.array_equal(out["section"], golden[name + "_section"][t]), t
        for k in ("d", "axial", "ratio"):
            assert nrm(out[k], golden["%s_%s" % (name, k)][t]) <= FP64_TOL, (t, k)
        assert np.array_equal(out["point"], golden[name + "_point"][t]), t         # float32, same operations


def test_genes_decode_corner_cases():
    """the quirks of read_genes the restatement must keep: the roof loop's for-else zeroes the last node before the
    pair fix, tops are lifted to d_min, sections round half to even and are mirrored from the right"""
    o = TrussOracle("small_roof")
    N, E = o.mesh.N, o.mesh.E
    g = np.zeros(N + E)
    y, sec = genes_oracle.decode_genes(o, g, 8)
    assert y[N - 1] == y[o.mesh.N // 2] and all(v in (0, 0.3) or abs(v) < 1e-12 or v == 0.3 for v in y)
    g = np.full(N + E, 0.5)
    g[N:] = 0.125                    # round(0.5) == 0 (half to even)
    _, sec = genes_oracle.decode_genes(o, g, 8)
    assert set(sec) == {0}
    g[N:] = 0.375                    # round(1.5) == 2
    _, sec = genes_oracle.decode_genes(o, g, 8)
    assert set(sec) == {2}
    g[N:] = np.linspace(0, 1, E)
    _, sec = genes_oracle.decode_genes(o, g, 8)
    for a, b in o.mesh.sym_elem_pairs:
        assert sec[a] == sec[b] == min(4, round(g[N + max(a, b)] * 4))


@pytest.mark.reference
@pytest.mark.parametrize("name", FAMILY_NAMES)
def test_genes_oracle_vs_reference_moead(name):
    ref = ref_harness.RefMoead(name)
    o = TrussOracle(name)
    assert float(ref.int_obj1) == o.int_obj1 and float(ref.int_obj2) == o.int_obj2
    rng = np.random.RandomState(7)
    N, E = o.mesh.N, o.mesh.E
    for t in range(25):
        g = rng.rand(N + E)
        if t % 4 == 1:
            g[:N] *= 0.04
        if t % 4 == 2:
            g[N:] = np.round(g[N:] * 8) / 8
        want = ref.read_genes(g)
        got = genes_oracle.read_genes(o, g)
        assert np.array_equal(got["y"], want["y"]) and np.array_equal(got["section"], want["section"])
        for k in ("d", "axial", "ratio"):
            assert nrm(got[k], want[k]) <= FP64_TOL
        assert np.array_equal(got["point"], want["point"])
