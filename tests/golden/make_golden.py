"""Generates tests/golden/<family>.npz by running the UNMODIFIED reference modules (imported read-only
from /root/reference through oracle/ref_harness.py) in FEM-coerced mode.

Run in the build container only:  python tests/golden/make_golden.py

Per family the file holds
  * the static tables the reference generates (connectivity, DOF ids, supports, loads, A_n, mask, c_e,
    the hard-coded symmetry lists parsed from truss2D_ENV.py, int_obj1/2),
  * the reset-time state (_game_get_1_state on the generated geometry),
  * T recorded _game_modify transitions (inputs incl. the stale move range and the forced coin,
    every returned tensor, and the FP64 fields of the solved model).
The walks follow the driver's pattern: three children per parent from the same parent state, the move
range left behind by the previous call.  "saturated" walks use out-of-range actions and a deliberately
wrong stale move range so that the y<y_min / y>y_max / depth passes fire.
"""
import contextlib
import io
import os
import re
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from oracle import ref_harness  # noqa: E402

OUT_DIR = os.path.dirname(os.path.abspath(__file__))
STEPS = {"small_bridge": 12, "small_roof": 12, "large_bridge": 6, "large_roof": 6}   # parents per mode


def parse_symmetry(code_dir, N):
    src = open(os.path.join(code_dir, "truss2D_ENV.py")).read()
    body = src[src.index("# ASSIGN SYMMETRY NODE"):src.index("# Structural analysis")]
    node_part, elem_part = body.split("# ASSIGN SYMMETRY ELEMENT")
    t_part, f_part = node_part.split("else:")
    rx = re.compile(r"nodes\[(\d+)\]\.coord\[1\] = self\.gen_model\.model\.nodes\[(\d+)\]\.coord\[1\]")
    tabs = []
    for part in (f_part, t_part):
        want = list(range(N))
        for dst, s in rx.findall(part):
            want[int(dst)] = int(s)
        tabs.append(want)
    ex = re.compile(r"elements\[(\d+)\]\.section_no = min\(self\.gen_model\.model\.elements\[(\d+)\]\.section_no,"
                    r"self\.gen_model\.model\.elements\[(\d+)\]\.section_no\)")
    pairs = sorted(set((min(int(a), int(b)), max(int(a), int(b))) for _, a, b in ex.findall(elem_part)))
    return np.array(tabs, dtype=np.int32), np.array(pairs, dtype=np.int32)


def state_arrays(prefix, st, out):
    names = ("x_n", "A_n", "A_s", "A_n_ts", "A_n_cs", "mask", None, None, "nN_x_n", "nN_x_e", "nC_e")
    for name, arr in zip(names, st):
        if name in ("x_n", "A_s", "A_n_ts", "A_n_cs", "nN_x_n", "nN_x_e"):
            out.setdefault(prefix + name, []).append(np.asarray(arr, dtype=np.float32))


def fem_arrays(prefix, ref, out):
    f = ref.fem_fields()
    for k in ("y", "section", "d", "axial", "ratio", "iscompress", "length", "reactions", "max_up", "max_down"):
        out.setdefault(prefix + k, []).append(f[k])
    out.setdefault(prefix + "U", []).append(np.float64(f["U"]))
    weak = np.array([not isinstance(n.coord[1], np.floating) for n in ref.gen.model.nodes])
    out.setdefault(prefix + "y_weak", []).append(weak)


def make(run):
    ref = ref_harness.RefGame(run, fem_fp64=True)
    m = ref.gen.model
    N, E = len(m.nodes), len(m.elements)
    data = {}
    st0 = ref.reset_state()
    data["conn"] = np.array([[e.nodes[0].name - 1, e.nodes[1].name - 1] for e in m.elements], dtype=np.int32)
    data["tnsc"] = np.array(m.tnsc, dtype=np.int32)
    data["ndof"] = np.int32(m.ndof)
    data["res"] = np.array([n.res for n in m.nodes], dtype=np.int32)
    data["top"] = np.array([n.top_node for n in m.nodes], dtype=np.int32)
    data["pair"] = np.array([n.vertical_pair[0].name - 1 for n in m.nodes], dtype=np.int32)
    data["loaded"] = np.array([int(len(n.loads) != 0) for n in m.nodes], dtype=np.int32)
    data["loadvec"] = np.array([v[0] for v in m.jlv], dtype=np.float64)
    data["x"] = np.array([n.coord[0] for n in m.nodes], dtype=np.float64)
    data["y0"] = np.array([float(n.coord[1]) for n in m.nodes], dtype=np.float64)
    data["target"] = np.array([n.target if n.top_node else 0.0 for n in m.nodes], dtype=np.float64)
    data["A_n"], data["mask"], data["nC_e"] = st0[1], st0[5], st0[10]
    data["int_obj"] = np.array([ref.game.int_obj1, ref.game.int_obj2], dtype=np.float32)
    data["scalars"] = np.array([ref.gen.y_max, ref.gen.y_min, ref.gen.d_min, ref.gen.max_deformation], dtype=np.float64)
    data["sym_src"], data["sym_elem_pairs"] = parse_symmetry(ref.mods.code_dir, N)
    rec = {}
    state_arrays("reset_", st0, rec)
    fem_arrays("reset_", ref, rec)
    for k, v in rec.items():
        data[k] = v[0]

    tr = {}
    fired = {"lt_min": 0, "gt_max": 0, "weak_nonsupport": 0, "singular": 0}
    for mode_id, mode in enumerate(("uniform", "small_actions", "saturated")):
        rng = np.random.RandomState(100 + mode_id)
        with ref.mods.cwd(), contextlib.redirect_stdout(io.StringIO()):
            ref.gen.re_value(*ref.args)
        parent = ref.reset_state()
        for k in range(STEPS[run]):
            children = []
            for child in range(3):
                if mode == "uniform":
                    a_geo, a_topo = rng.rand(N, 2), rng.rand(N, 3)
                elif mode == "small_actions":
                    a_geo, a_topo = rng.rand(N, 2) * 0.2, rng.rand(N, 3) * np.array([0.3, 0.3, 1.0])
                else:
                    a_geo, a_topo = rng.randn(N, 2) * 2 + 0.5, rng.randn(N, 3) * 2 + 0.5
                a_geo, a_topo = a_geo.astype(np.float32), a_topo.astype(np.float32)
                if mode == "saturated" and (3 * k + child) % 4 == 0:
                    ref.set_move_range(rng.rand(N) * 40, rng.rand(N) * 40)
                coin = bool(rng.rand() >= 0.5)
                f = ref.fem_fields()
                raw_geo, raw_topo = a_geo.copy(), a_topo.copy()
                try:
                    point, S = ref.step(parent[-3], parent[-2], parent[-1], a_geo, a_topo, coin)
                except np.linalg.LinAlgError:
                    fired["singular"] += 1
                    continue
                tr.setdefault("in_set_node", []).append(np.asarray(parent[-3], dtype=np.float32))
                tr.setdefault("in_set_element", []).append(np.asarray(parent[-2], dtype=np.float32))
                tr.setdefault("in_max_up", []).append(f["max_up"])
                tr.setdefault("in_max_down", []).append(f["max_down"])
                tr.setdefault("in_a_geo", []).append(raw_geo)
                tr.setdefault("in_a_topo", []).append(raw_topo)
                tr.setdefault("in_coin", []).append(np.uint8(coin))
                tr.setdefault("out_a_geo", []).append(a_geo)       # clipped in place by the reference
                tr.setdefault("out_a_topo", []).append(a_topo)
                tr.setdefault("out_point", []).append(np.array(point, dtype=np.float32))
                tr.setdefault("mode", []).append(np.int32(mode_id))
                state_arrays("out_", S, tr)
                fem_arrays("out_", ref, tr)
                yw = tr["out_y_weak"][-1]
                sup = data["res"][:, 1] == 1
                fired["weak_nonsupport"] += int(np.any(yw & ~sup))
                children.append(S)
            if children:
                parent = children[rng.randint(len(children))]
    for k, v in tr.items():
        data["tr_" + k] = np.stack(v)
    path = os.path.join(OUT_DIR, run + ".npz")
    np.savez_compressed(path, **data)
    print(run, "transitions:", len(tr["mode"]), "fired:", fired, "->", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    for run in ("small_bridge", "small_roof", "large_bridge", "large_roof"):
        make(run)
