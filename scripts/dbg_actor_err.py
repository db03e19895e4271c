import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from mop_truss_marl_b200 import actor, tf_checkpoint
from oracle.actor_oracle import actor_forward
from test_gpu_actor import random_inputs

def truss_mask(N):
    nx = N // 2
    m = np.eye(N, dtype=np.float32)
    for i in range(nx):
        for j in range(nx):
            if abs(i - j) <= 1:
                m[i, j] = m[nx + i, nx + j] = m[i, nx + j] = m[nx + j, i] = 1
    return m

if __name__ == "__main__":
  for N, B, P, mode in [(16, 37, 1, 'dense'), (16, 37, 3, 'sparse'), (32, 9, 50, 'dense'), (32, 21, 2, 'sparse')]:
      rng = np.random.RandomState(N + B)
      w = tf_checkpoint.random_actor_weights(seed=3)
      for k in w:
          w[k] = (w[k][0], (rng.randn(*w[k][1].shape) * 0.05).astype(np.float32))
      inp = list(random_inputs(rng, B, N, P))
      if mode == 'sparse':
          m = truss_mask(N)
          inp[1] = (rng.rand(N, N).astype(np.float32) * 0.3 + 0.05) * m
          for i in (2, 3, 4):
              inp[i] = (rng.rand(B, N, N).astype(np.float32) * (rng.rand(B, N, N) < 0.7)) * (m - np.eye(N, dtype=np.float32))
      a = actor.BatchedActor(w, N, max_batch=B)
      dev = [torch.from_numpy(np.ascontiguousarray(t)).cuda() for t in inp]
      geo, topo = a.forward(*dev)
      torch.cuda.synchronize()
      g64, t64 = actor_forward(w, *inp)
      eg = np.abs(geo.cpu().numpy() - g64).reshape(B * N, 2).max(1)
      et = np.abs(topo.cpu().numpy() - t64).reshape(B * N, 3).max(1)
      try:
          a.check(); st = "status ok"
      except Exception as ex:
          st = "status " + str(ex)
      print(N, B, mode, "max err geo %.3e topo %.3e" % (eg.max(), et.max()), st, flush=True)
      if max(eg.max(), et.max()) > 2e-5:
          for t in range((B * N + 127) // 128):
              print("  tile", t, "geo %.2e topo %.2e" % (eg[t*128:(t+1)*128].max(), et[t*128:(t+1)*128].max()))
