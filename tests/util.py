"""shared helpers for the parity tests (tolerances = SURVEY.md section 8c)"""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FAMILY_NAMES = ("small_bridge", "small_roof", "large_bridge", "large_roof")
F32_FIELDS = ("x_n", "A_s", "A_n_ts", "A_n_cs", "nN_x_n", "nN_x_e")
FP64_TOL = 1e-9          # normwise relative tolerance on d, axial, ratio, U, reactions (north_star)
F32_ULP_TOL = 2          # float32 observation tensors


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))


def ulp_diff(a, b):
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    ia = a.view(np.int32).astype(np.int64)
    ib = b.view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, -(ia & 0x7FFFFFFF), ia)
    ib = np.where(ib < 0, -(ib & 0x7FFFFFFF), ib)
    return np.abs(ia - ib)


def nrm(a, b):
    """max |a-b| / max |b|"""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-300))


def assert_f32_close(name, got, want, tol=F32_ULP_TOL):
    got = np.asarray(got); want = np.asarray(want)
    assert got.shape == want.shape, (name, got.shape, want.shape)
    u = ulp_diff(got, want)
    assert u.max() <= tol, "%s: %d ulp at %s (got %r want %r)" % (
        name, u.max(), np.unravel_index(u.argmax(), u.shape), got.flat[u.argmax()], want.flat[u.argmax()])


def adversarial_actions(rng, N):
    """actions that hit the reference's corner cases: exact 0 / 1, ties (np.argmax takes the first), out of
    range, +-inf and NaN in a_geo (clip leaves NaN, argmax picks it, min([1, nan]) == 1).  a_topo stays finite:
    the reference forms nC_e @ a_topo, where one NaN poisons every element through 0*NaN -- not reproduced."""
    pool = np.array([0.0, 1.0, 0.5, 0.5, 0.25, 1.5, -0.5, 0.999999, 1e-7, np.inf, -np.inf, np.nan], dtype=np.float32)
    a_geo = pool[rng.randint(0, len(pool), size=(N, 2))].astype(np.float32)
    mix = rng.rand(N, 2) < 0.4
    a_geo = np.where(mix, rng.rand(N, 2).astype(np.float32), a_geo).astype(np.float32)
    tpool = np.array([0.0, 1.0, 0.5, 0.5, 0.25, 1.5, -0.5, 0.75], dtype=np.float32)
    a_topo = tpool[rng.randint(0, len(tpool), size=(N, 3))].astype(np.float32)
    return a_geo, a_topo


def adversarial_move_range(rng, N):
    pool = np.array([0.0, 0.3, 7.7, 8.0, 40.0, 1e-3, 2.0], dtype=np.float32)
    return pool[rng.randint(0, len(pool), size=N)], pool[rng.randint(0, len(pool), size=N)]
