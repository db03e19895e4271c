"""CPU: the C-ABI library loads, exports every symbol include/tfem.h declares, builds the same static
tables as the reference (golden) and refuses to compute without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from mop_truss_marl_b200 import FAMILIES, capi, family_desc
from util import FAMILY_NAMES, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exports_match_header():
    hdr = open(os.path.join(ROOT, "include", "tfem.h")).read()
    declared = set(re.findall(r"\b(tfem_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no prototypes found"
    assert declared == set(capi.EXPORTS)
    for name in declared:
        assert hasattr(capi.lib, name), name
    assert b"sm_100a" in capi.lib.tfem_version()
    hdr2 = open(os.path.join(ROOT, "include", "tactor.h")).read()
    declared2 = set(re.findall(r"\b(tactor_[a-z_0-9]+)\s*\(", hdr2))
    from mop_truss_marl_b200 import actor
    assert declared2 == set(actor.ACTOR_EXPORTS)
    for name in declared2:
        assert hasattr(capi.lib, name), name
    hdr4 = open(os.path.join(ROOT, "include", "tpareto.h")).read()
    declared4 = set(re.findall(r"\b(tpareto_[a-z_0-9]+)\s*\(", hdr4))
    from mop_truss_marl_b200 import pareto
    assert declared4 == set(pareto.PARETO_EXPORTS)
    for name in declared4:
        assert hasattr(capi.lib, name), name
    hdr3 = open(os.path.join(ROOT, "include", "trollout.h")).read()
    declared3 = set(re.findall(r"\b(trollout_[a-z_0-9]+)\s*\(", hdr3))
    from mop_truss_marl_b200 import host_pipeline
    assert declared3 == set(host_pipeline.ROLLOUT_EXPORTS)
    for name in declared3:
        assert hasattr(capi.lib, name), name


def test_struct_sizes():
    assert C.sizeof(capi.FamilyDesc) == 16 + 8 * (15 + 1 + 16 + 2 + 5 + 5 + 2)
    assert C.sizeof(capi.StepIn) == 8 * 8 and C.sizeof(capi.StepOut) == 18 * 8 and C.sizeof(capi.Dims) == 32


@pytest.mark.parametrize("name", FAMILY_NAMES)
def test_tables_match_reference(name):
    g = load_golden(name)
    h = capi.Handle(family_desc(FAMILIES[name]), device=-1)
    assert (h.dims.N, h.dims.E, h.dims.ndof) == (g["top"].shape[0], g["conn"].shape[0], int(g["ndof"]))
    for key in ("conn", "tnsc", "res", "top", "pair", "loaded", "loadvec", "x", "y0", "target", "A_n", "mask", "nC_e",
                "sym_src", "int_obj"):
        assert np.array_equal(h.table(key), g[key]), key
    partner = np.arange(h.dims.E)
    for a, b in g["sym_elem_pairs"]:
        partner[a], partner[b] = b, a
    assert np.array_equal(h.table("sym_elem"), partner)
    assert np.array_equal(h.table("scalars")[:4], g["scalars"])
    h.close()


def test_no_cpu_fallback():
    h = capi.Handle(family_desc(FAMILIES["small_bridge"]), device=-1)
    out = capi.StepOut()
    rc = capi.lib.tfem_reset(h.ptr, 1, None, C.byref(out), None)
    assert rc == -2 and b"no CPU path" in capi.lib.tfem_last_error()
    y = np.zeros((1, 16)); s = np.zeros((1, 36), np.int32)
    rc = capi.lib.tfem_solve_only(h.ptr, 1, y.ctypes.data, s.ctypes.data, None, None, None, None, None, None, None)
    assert rc == -2


def test_bad_family_rejected():
    import dataclasses
    spec = dataclasses.replace(FAMILIES["small_bridge"], num_x=7, span_x=(5,) * 6, tar_y=(1,) * 7)
    with pytest.raises(capi.TfemError):
        capi.Handle(family_desc(spec), device=-1)
