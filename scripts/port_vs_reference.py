"""How fast is the CPU port (oracle/truss_oracle.py, bench.py's `cpu_baseline` and reference arm) next to the reference's
own ``_game_modify`` loop?  Runs only in the build container (/root/reference is needed): both are driven with the same
i.i.d. uniform actions from the same reset state, own state fed back, one core, env-step only (no actor: TensorFlow is not
installable here).  Writes profiles/r2_port_vs_reference.json; bench.py quotes the factor in `cpu_baseline.sample`.

    python scripts/port_vs_reference.py [seconds per family]"""
import json
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")
from oracle import ref_harness  # noqa: E402
from oracle.truss_oracle import TrussOracle  # noqa: E402


def time_reference(family, seconds):
    g = ref_harness.RefGame(family, fem_fp64=False)       # the reference exactly as it runs
    rng = np.random.RandomState(0)
    st = g.reset_state()
    set_node, set_element, nC_e = st[-3], st[-2], st[-1]
    N = set_node.shape[0]
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        a_geo, a_topo = rng.rand(N, 2).astype(np.float32), rng.rand(N, 3).astype(np.float32)
        _, ns = g.step(set_node, set_element, nC_e, a_geo, a_topo, bool(rng.rand() >= 0.5))
        set_node, set_element = ns[-3], ns[-2]
        n += 1
    return n / (time.perf_counter() - t0)


def time_port(family, seconds):
    o = TrussOracle(family)
    rng = np.random.RandomState(0)
    N = o.mesh.N
    st = o.reset()
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        a_geo, a_topo = rng.rand(N, 2).astype(np.float32), rng.rand(N, 3).astype(np.float32)
        st = o.step(st["nN_x_n"], st["nN_x_e"], st["max_up"], st["max_down"], a_geo, a_topo, rng.rand() >= 0.5)
        n += 1
    return n / (time.perf_counter() - t0)


def main():
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 8.0
    out = {}
    for family in ("small_bridge", "small_roof", "large_bridge", "large_roof"):
        ref, port = time_reference(family, seconds), time_port(family, seconds)
        out[family] = {"reference_steps_per_s": ref, "port_steps_per_s": port, "port_over_reference": port / ref,
                       "seconds_each": seconds, "cores": 1, "what": "_game_modify env-step only, uniform actions, own state fed back"}
        print(family, out[family])
    json.dump(out, open(os.path.join(ROOT, "profiles", "r2_port_vs_reference.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
