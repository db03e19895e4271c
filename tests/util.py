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
