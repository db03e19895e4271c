"""Writes tests/golden/actor_2000pickle_base.npz: the three trained actors of the reference's checkpoint
``model/2000pickle_base/Agent{1,2,3}_Actor_pickle`` (26 float32 tensors each, 291 805 parameters per agent), read with the
TF-free reader ``mop_truss_marl_b200.tf_checkpoint.load_actor_weights``.  The checkpoint itself cannot travel to the GPU
box (nothing there may read /root/reference), so the GPU tests and bench.py take the trained weights from this fixture.

    python tests/golden/make_actor_golden.py          # run in the build container (needs /root/reference)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mop_truss_marl_b200 import tf_checkpoint  # noqa: E402

REF = os.environ.get("TRUSS_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden", "actor_2000pickle_base.npz")


def main():
    arrays = {}
    for agent in (1, 2, 3):
        w = tf_checkpoint.load_actor_weights(os.path.join(REF, "model", "2000pickle_base", "Agent%d_Actor_pickle" % agent))
        assert list(w) == list(tf_checkpoint.ACTOR_LAYERS)
        for name, (kernel, bias) in w.items():
            arrays["agent%d/%s/kernel" % (agent, name)] = np.asarray(kernel, np.float32)
            arrays["agent%d/%s/bias" % (agent, name)] = np.asarray(bias, np.float32)
    np.savez_compressed(OUT, **arrays)
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(arrays), "tensors")


if __name__ == "__main__":
    main()
