import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from mop_truss_marl_b200 import actor, tf_checkpoint
from oracle.actor_oracle import actor_forward
from test_gpu_actor import random_inputs
N, B, P = 16, 37, 1
rng = np.random.RandomState(N + B)
w = tf_checkpoint.random_actor_weights(seed=3)
inp = random_inputs(rng, B, N, P)
a = actor.BatchedActor(w, N, max_batch=B)
dev = [torch.from_numpy(t).cuda() for t in inp]
geo, topo = a.forward(*dev)
torch.cuda.synchronize()
g64, t64 = actor_forward(w, *inp)
eg = np.abs(geo.cpu().numpy() - g64).reshape(B * N, 2).max(1)
et = np.abs(topo.cpu().numpy() - t64).reshape(B * N, 3).max(1)
print("max err geo %.3e topo %.3e" % (eg.max(), et.max()))
for t in range((B * N + 127) // 128):
    print("tile", t, "geo %.2e topo %.2e" % (eg[t*128:(t+1)*128].max(), et[t*128:(t+1)*128].max()))
try:
    a.check(); print("status ok")
except Exception as e:
    print("status", e)
