"""Banded production solve (tfem_solve_only) vs the north_star's dense blocked Cholesky with DMMA trailing updates
(tfem_solve_dense_dmma) on the same batch: CUDA-event time per launch, L2 flushed between launches.
usage: python scripts/solver_compare.py [family] [batch]"""
import json, sys
import numpy as np, torch
sys.path.insert(0, '.')
from mop_truss_marl_b200 import batched_env

family = sys.argv[1] if len(sys.argv) > 1 else "large_bridge"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
env = batched_env.BatchedTrussEnv(family, 1)
N, E = env.N, env.E
rng = np.random.RandomState(0)
y = np.zeros((B, N)); y[:, N // 2:] = 1.0 + rng.rand(B, N // 2) * 3.0; y[:, 1:N // 2 - 1] = rng.rand(B, N // 2 - 2) * 0.6
sec = rng.randint(0, 5, size=(B, E)).astype(np.int32)
yt, st = torch.from_numpy(y).cuda(), torch.from_numpy(sec).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def timeit(fn, reps=20):
    for _ in range(3): fn()
    ts = []
    for i in range(reps):
        flush.fill_(i & 0xff)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))

t_band = timeit(lambda: env.solve_only(yt, st))
t_dense = timeit(lambda: env.solve_dense_dmma(yt, st))
d0 = env.solve_only(yt, st)["d"]; d1, _ = env.solve_dense_dmma(yt, st)
err = float(((d0 - d1).abs().amax(1) / d0.abs().amax(1)).max())
print(json.dumps({"family": family, "batch": B, "ndof": env.ndof, "banded_ms": t_band, "dense_dmma_ms": t_dense,
                  "banded_solves_per_s": B / t_band * 1e3, "dense_dmma_solves_per_s": B / t_dense * 1e3,
                  "dense_over_banded_time": t_dense / t_band, "max_rel_diff_d": err,
                  "note": "banded time includes member forces, stress ratios, U and reactions; dense computes d only"}))
