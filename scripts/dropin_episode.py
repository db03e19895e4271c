"""BASELINE.json configs[0] with the drop-in modules: one test episode of the reference driver's inner loop
(master_DDPG_truss2D_MO.py:167-260) as a batch of ONE -- three agents act on the parent state, each child goes through
game._game_modify -- driven exactly like the unchanged driver drives the reference modules.  Reports the per-call latency
of the batch-of-1 path (every call crosses PCIe and launches kernels for a single environment: this path exists for
drop-in compatibility, the batched path is what the library is built for)."""
import contextlib
import io
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from mop_truss_marl_b200 import compat  # noqa: E402

compat.install()
import truss2D_ENV  # noqa: E402
import truss2D_GEN  # noqa: E402
import truss2D_RL  # noqa: E402

STEPS = int(sys.argv[1]) if len(sys.argv) > 1 else 100
with contextlib.redirect_stdout(io.StringIO()):
    gm = truss2D_GEN.gen_model(8, 2, [5] * 7, [8], [4, 3, 2.5, 2, 2, 2.5, 3, 4], 0.3, 0, -75 * 1000, "bridge", 1, None)
    game = truss2D_ENV.Game_research04(500, gm, 3)
env = truss2D_ENV.ENV(game)
env.reset()
mu = [[0.1, 0.1], [0.1, 0.1, 0.1]]
maddpg = truss2D_RL.MADDPG(1e-7, 1, 0.95, 0.99, 200, 200, 1000, 3, [2, 3], mu, mu, mu)   # random-init actors (no checkpoint here)
np.random.seed(20)
state = game._game_get_1_state()
t_act = t_env = 0.0
n = 0
for step in range(STEPS + 5):
    if step == 5:
        t_act = t_env = 0.0
        n = 0
    children = []
    for k in range(3):
        t0 = time.perf_counter()
        a_geo, a_topo = maddpg.agents[k].act(state[0], state[1], state[2], state[3], state[4], state[6], state[7])
        t1 = time.perf_counter()
        point, child = game._game_modify(state[8], state[9], state[10], [a_geo, a_topo])
        t2 = time.perf_counter()
        t_act += t1 - t0; t_env += t2 - t1; n += 1
        children.append((point, child))
    best = min(range(3), key=lambda i: float(children[i][0][0]) + float(children[i][0][1]))
    child = children[best][1]
    state = list(child)
    state[6], state[7] = truss2D_ENV.pareto_state_data([[1.0, 1.0]], 0)
    game.step()
print(json.dumps({"workload": "drop-in batch of 1, small bridge, %d parent steps x 3 agents" % STEPS,
                  "act_ms_per_call": 1e3 * t_act / n, "game_modify_ms_per_call": 1e3 * t_env / n,
                  "env_steps_per_s": n / (t_act + t_env), "fem_analyses": gm._tfem.analyses}))
