import ctypes as C, numpy as np, torch, sys
sys.path.insert(0, '.')
from mop_truss_marl_b200 import actor, tf_checkpoint, capi
B, N = 4096, 16
a = actor.BatchedActor(tf_checkpoint.random_actor_weights(1), N, B)
g = torch.Generator(device='cuda').manual_seed(0)
r = lambda *s: torch.rand(*s, device='cuda', generator=g)
inp = (r(B, N, 13), r(N, N), r(B, N, N), r(B, N, N), r(B, N, N), r(B, 1, 4), r(B, 1, 1))
for _ in range(3): a.forward(*inp)
torch.cuda.synchronize()
buf = np.zeros(512, np.int64)
capi.lib.tactor_debug_dump(a._h, buf.ctypes.data_as(C.c_void_p))
for who, off in (("generator thread 0", 16), ("issuer lane", 144), ("producer lane", 272)):
    d = buf[off:off + 28].reshape(7, 4)
    print(who, ": per GEMM [setup+sync, main loop (to acc done), epilogue, start offset]")
    for g in range(7): print(g, d[g].tolist())
    acc = buf[off + 32: off + 32 + 56].reshape(7, 8)
    per = np.diff(np.vstack([np.zeros((1, 8), np.int64), acc]), axis=0)
    print("  per GEMM [iss wait A, iss wait W, iss mma+commit, prod wait empty, prod issue_w | gen wait empty, gen fill, gen fence+arrive]")
    for g in range(7): print("  ", g, per[g].tolist())
