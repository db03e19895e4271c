"""CPU: how accurate is the REFERENCE's own float64 solve?  The golden transitions recorded from the reference
(tests/golden/*.npz: tr_out_d is what its np.linalg.solve returned) against the extended-precision solve of the same
geometries (oracle/exact_fem.py).  This is the evidence behind the conditioning-aware FP64 tolerance of the GPU parity
tests: up to cond(K) = 1e6 the reference is within 1e-9 of the exact displacements by a wide margin, beyond that its error
grows like eps * cond and reaches 1e-9 .. 1e-8 on the d_min-deep trusses of the saturated walks -- no float64 solver can
be asked to agree with it more closely than it agrees with the truth."""
import numpy as np
import pytest

from oracle import exact_fem
from oracle.truss_oracle import FAMILIES, build_mesh, fem_solve
from util import FAMILY_NAMES, load_golden

EPS = np.finfo(np.float64).eps


@pytest.mark.parametrize("name", FAMILY_NAMES)
def test_reference_solve_error_grows_with_conditioning(name):
    g = load_golden(name)
    m = build_mesh(FAMILIES[name])
    y, sec, d_ref = g["tr_out_y"], g["tr_out_section"], g["tr_out_d"]
    d_exact = exact_fem.exact_displacements(m, y, sec)
    cond = exact_fem.cond2(m, y, sec)
    err_ref = exact_fem.rel_err(d_ref, d_exact)
    # the extended-precision solve is converged: a second refinement step does not move it
    K, P = exact_fem.assemble(m, y, sec)
    r = P - np.einsum("mij,mj->mi", K, d_exact)
    assert exact_fem.rel_err(d_exact + exact_fem.ldl_solve(K, r), d_exact).max() <= 1e-12   # (1e-19 * cond * cancellation)
    # the reference's LU answer obeys the textbook bound ...
    assert (err_ref <= 20 * EPS * cond).all(), float((err_ref / (EPS * cond)).max())
    # ... is comfortably inside 1e-9 while cond <= 1e6 ...
    assert err_ref[cond <= 1e6].max() <= 1e-10
    # ... and our numpy restatement (same LAPACK call) reproduces its error level case by case
    err_orc = exact_fem.rel_err(np.stack([fem_solve(m, list(y[i]), list(sec[i]))["d"] for i in range(len(y))]), d_exact)
    assert (err_orc <= 20 * EPS * cond).all()
    print("%s: cond %.1e..%.1e, reference error %.1e..%.1e (%.2f eps cond at most)" % (
        name, cond.min(), cond.max(), err_ref.min(), err_ref.max(), (err_ref / (EPS * cond)).max()))


def test_saturated_walks_exceed_flat_tolerance_budget():
    """on the worst golden geometries the reference itself is more than 1e-10 away from the exact displacements: a flat
    1e-9 between two float64 solvers there would be a statement about rounding luck, not about correctness"""
    worst = 0.0
    for name in FAMILY_NAMES:
        g = load_golden(name)
        m = build_mesh(FAMILIES[name])
        d_exact = exact_fem.exact_displacements(m, g["tr_out_y"], g["tr_out_section"])
        worst = max(worst, exact_fem.rel_err(g["tr_out_d"], d_exact).max())
    assert worst >= 1e-10, worst
