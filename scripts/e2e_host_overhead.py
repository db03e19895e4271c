"""Where the end-to-end step's wall time goes: Python argument checks, the C call (checks, cudaGraphLaunch, wait) and the GPU
time line.  TROLLOUT_HOSTTIME=1 python scripts/e2e_host_overhead.py [family batch]  (development helper, 1 GPU)"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mop_truss_marl_b200 import actor, batched_env, host_pipeline, tf_checkpoint  # noqa: E402
from mop_truss_marl_b200.host_pipeline import HostRollout, STATE_IN  # noqa: E402


def main():
    family = sys.argv[1] if len(sys.argv) > 1 else "small_bridge"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    env = batched_env.BatchedTrussEnv(family, B)
    env.reset()
    pol = actor.BatchedActor(tf_checkpoint.random_actor_weights(seed=20), env.N, B)
    roll = HostRollout(env, pol, pieces=2)
    bufs = [roll.alloc_host(), roll.alloc_host()]
    for k in STATE_IN:
        bufs[0][k].copy_(getattr(env, k))
    torch.cuda.synchronize()
    HostRollout.fill_compact(bufs[0])
    x_p = torch.tensor([[1, 1, 1, 1 / 50]], dtype=torch.float32).repeat(B, 1, 1).pin_memory()
    A_p = torch.ones(B, 1, 1).pin_memory()
    coin = (torch.rand(B) >= 0.5).to(torch.uint8).pin_memory()
    c_time = [0.0]
    real = host_pipeline._lib.trollout_step_host

    def timed(*a):
        t0 = time.perf_counter()
        r = real(*a)
        c_time[0] += time.perf_counter() - t0
        return r
    host_pipeline._lib.trollout_step_host = timed
    for it in range(10):
        roll.step(bufs[it & 1], coin, x_p, A_p, bufs[1 - (it & 1)])
    c_time[0] = 0.0
    K = 200
    t0 = time.perf_counter()
    for it in range(K):
        roll.step(bufs[it & 1], coin, x_p, A_p, bufs[1 - (it & 1)])
    wall = time.perf_counter() - t0
    print(json.dumps({"family": family, "B": B, "wall_us_per_step": wall / K * 1e6, "c_call_us": c_time[0] / K * 1e6,
                      "python_us": (wall - c_time[0]) / K * 1e6}))
    roll.close()


if __name__ == "__main__":
    main()
