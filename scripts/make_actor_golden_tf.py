"""Pins the actor restatement against the reference's own TensorFlow model (cannot run in the build container: it needs
TensorFlow 2.11 + spektral 1.2.0, the versions the reference's README names).

    python scripts/make_actor_golden_tf.py /path/to/MOP-truss-MARL

For every (agent, family, P) case of ``tests/test_gpu_actor.py::test_trained_checkpoint_on_golden_states`` it builds the
same inputs (golden states + padded Pareto graphs, same seeds), loads ``model/2000pickle_base/Agent<k>_Actor_pickle``
into the reference's ``multimodes_actor`` (``train/code/truss2D_RL.py:49-127``) and writes the inputs and the model's
outputs to ``tests/golden/actor_tf.npz``.  When that file exists the GPU test also checks the float64 oracle against it
(1e-5), which turns "parity unpinned" into a pinned oracle for the actor."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ref = sys.argv[1]
    sys.path.insert(0, os.path.join(ref, "train", "code"))
    import tensorflow as tf  # noqa: F401
    from truss2D_RL import multimodes_actor            # the reference's model class, unmodified
    from test_gpu_actor import padded_pareto_graph
    from util import load_golden
    out = {}
    for agent in (1, 2, 3):
        model = multimodes_actor(200, 2, 3)                # n_hidden, n_action1, n_action2 (truss2D_RL.py:300-303)
        model.load_weights(os.path.join(ref, "model", "2000pickle_base", "Agent%d_Actor_pickle" % agent))
        for family, P in (("small_bridge", 1), ("small_bridge", 17), ("small_roof", 50), ("large_bridge", 50), ("large_roof", 17)):
            g = load_golden(family)
            rng = np.random.RandomState(100 * agent + P)
            x_n = np.concatenate([g["reset_x_n"][None], g["tr_out_x_n"]]).astype(np.float32)
            A_s = np.concatenate([g["reset_A_s"][None], g["tr_out_A_s"]]).astype(np.float32)
            A_ts = np.concatenate([g["reset_A_n_ts"][None], g["tr_out_A_n_ts"]]).astype(np.float32)
            A_cs = np.concatenate([g["reset_A_n_cs"][None], g["tr_out_A_n_cs"]]).astype(np.float32)
            B = x_n.shape[0]
            sizes = rng.randint(1, P + 1, size=B)
            sizes[0] = P
            x_p, A_p = padded_pareto_graph(rng, B, P, sizes)
            A_n = np.broadcast_to(g["A_n"], A_s.shape).astype(np.float32)
            geo, topo = model([x_n, A_n, A_s, A_ts, A_cs, x_p, A_p])      # call signature: truss2D_RL.py:75-84
            key = "agent%d/%s/P%d" % (agent, family, P)
            for name, arr in (("x_n", x_n), ("A_s", A_s), ("A_n_ts", A_ts), ("A_n_cs", A_cs), ("x_p", x_p), ("A_p", A_p),
                              ("geo", np.asarray(geo)), ("topo", np.asarray(topo))):
                out[key + "/" + name] = arr
    path = os.path.join(ROOT, "tests", "golden", "actor_tf.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    main()
