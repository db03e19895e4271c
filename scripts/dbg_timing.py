import ctypes as C, numpy as np, torch, sys
sys.path.insert(0, '.')
from mop_truss_marl_b200 import actor, tf_checkpoint, capi
from scripts.dbg_actor_err import truss_mask
B, N = 4096, 16
a = actor.BatchedActor(tf_checkpoint.random_actor_weights(1), N, B)
g = torch.Generator(device='cuda').manual_seed(0)
r = lambda *s: torch.rand(*s, device='cuda', generator=g)
m = torch.from_numpy(truss_mask(N)).cuda()
inp = (r(B, N, 13), r(N, N) * m, r(B, N, N) * m, r(B, N, N) * m, r(B, N, N) * m, r(B, 1, 4), r(B, 1, 1))
for _ in range(3): a.forward(*inp)
torch.cuda.synchronize()
buf = np.zeros(512, np.int64)
capi.lib.tactor_debug_dump(a._h, buf.ctypes.data_as(C.c_void_p))
names = {0: "generator warp 0 [setup, wait H, wait a_empty, compute+mix, chunk total(after wait), layer-1 x, split+st+wait::st, t_end]",
         1: "epilogue warp 8  [wait acc_full, drain+H/heads, head tail, -, -, -, -, t_end]",
         2: "issuer           [wait A, wait W, mma+commit issue, wait acc_empty, -, -, -, t_end]"}
for role in range(3):
    print(names[role])
    d = buf[16 + role * 64: 16 + role * 64 + 56].reshape(7, 8)
    for g in range(7): print("  ", g, d[g].tolist())
