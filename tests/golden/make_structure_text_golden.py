"""Generates tests/golden/structure_text.npz: files written by the UNMODIFIED reference ``gen_model.savetxt``
(``truss2D_GEN.py:193-211``), for every family at reset and after two env steps, with the heights / sections they
encode.  Under this container's NumPy 2 the reference prints float32 heights as ``np.float32(3.2)``; the pinned
NumPy 1.23.5 prints ``3.2`` (the same for the ``np.float64`` inertia inside ``[[I]]``) -- both forms are stored (``*_text_np2`` as written here, ``*_text`` with the wrapper
removed = the pinned environment's bytes).  Run in the build container only."""
import os
import re
import sys
import tempfile
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")
from oracle import ref_harness  # noqa: E402


def main():
    data = {}
    for fam in ("small_bridge", "small_roof", "large_bridge", "large_roof"):
        game = ref_harness.RefGame(fam)
        st = game.reset_state()
        rng = np.random.RandomState(4)
        texts, ys, weak, secs = [], [], [], []
        for step in range(3):
            if step:
                N = st[0].shape[0]
                point, st2 = game.step(st[8], st[9], st[10], rng.rand(N, 2).astype(np.float32),
                                       rng.rand(N, 3).astype(np.float32), bool(step & 1))
                st = list(st); st[8], st[9] = st2[8], st2[9]
            with tempfile.TemporaryDirectory() as d:
                p = os.path.join(d, "s.txt")
                with game.mods.cwd():
                    game.gen.savetxt(p)
                texts.append(open(p, newline="").read())
            m = game.gen.model
            ys.append([float(n.coord[1]) for n in m.nodes])
            weak.append([not isinstance(n.coord[1], np.floating) for n in m.nodes])
            secs.append([e.section_no for e in m.elements])
        data[fam + "_text_np2"] = np.array(texts)
        data[fam + "_text"] = np.array([re.sub(r"np\.float(?:32|64)\(([^()]*)\)", r"\1", t) for t in texts])
        data[fam + "_y"] = np.array(ys, dtype=np.float64)
        data[fam + "_y_weak"] = np.array(weak)
        data[fam + "_section"] = np.array(secs, dtype=np.int32)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "structure_text.npz")
    np.savez_compressed(path, **data)
    print("wrote", path, os.path.getsize(path))
    print(data["small_bridge_text"][1][:400])


if __name__ == "__main__":
    main()
