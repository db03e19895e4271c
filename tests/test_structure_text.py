"""CPU: the structure text format (savetxt / read_src) against files written by the unmodified
reference ``gen_model.savetxt`` (tests/golden/structure_text.npz, tests/golden/make_structure_text_golden.py)."""
import os

import numpy as np
import pytest

from mop_truss_marl_b200 import structure_text as st
from mop_truss_marl_b200.families import FAMILIES
from util import FAMILY_NAMES, load_golden


@pytest.fixture(scope="module")
def golden():
    return load_golden("structure_text")


@pytest.mark.parametrize("name", FAMILY_NAMES)
def test_writer_is_byte_identical_to_reference_savetxt(golden, name):
    for t in range(golden[name + "_y"].shape[0]):
        txt = st.format_structure(name, golden[name + "_y"][t], golden[name + "_section"][t],
                                  y_is_float32=~golden[name + "_y_weak"][t])
        assert txt == str(golden[name + "_text"][t]), (name, t)      # the pinned NumPy 1.23 form, byte for byte


@pytest.mark.parametrize("name", FAMILY_NAMES)
def test_reader_round_trip_and_numpy2_form(golden, name):
    spec = FAMILIES[name]
    for t in range(golden[name + "_y"].shape[0]):
        for key in ("_text", "_text_np2"):                           # np.float32(...) wrappers are accepted too
            d = st.parse_structure(str(golden[name + key][t]), name)
            # like read_src, the reader returns what the text says (6.2); the model held np.float32(6.2): the shortest
            # repr identifies the float32 uniquely
            assert np.array_equal(d["y"].astype(np.float32), golden[name + "_y"][t].astype(np.float32))
            assert np.array_equal(d["section"], golden[name + "_section"][t])
            assert len(d["loads"]) == 1 and d["loads"][1] == [0, spec.loady]
            assert len(d["nodes"]) == spec.N and len(d["elements"]) == spec.E
            assert d["elements"][1][:3] == (1, 2, 2 * 1e11)


def test_batch_files(tmp_path, golden):
    name = "small_roof"
    y, sec, weak = golden[name + "_y"], golden[name + "_section"], golden[name + "_y_weak"]
    paths = [os.path.join(tmp_path, "sub", "s%d.txt" % i) for i in range(y.shape[0])]
    st.write_batch(paths, name, y, sec, y_is_float32=~weak, workers=2)
    assert open(paths[1], newline="").read() == str(golden[name + "_text"][1])
    y2, sec2 = st.read_batch(paths, name)
    assert np.array_equal(y2.astype(np.float32), y.astype(np.float32)) and np.array_equal(sec2, sec)
    with pytest.raises(ValueError):
        st.parse_structure(" 1, 2, 3\r\n")
    with pytest.raises(ValueError):
        st.format_structure(name, y[0][:3], sec[0])
