"""GPU (needs 2 devices: `gpurun --gpus 2`; skipped on a one-GPU box): the NCCL twin of
tests/test_learner_cpu.py::test_gradient_allreduce_keeps_ranks_identical -- two CUDA ranks with different device-resident
replays stay bit-identical through MADDPG updates whose gradients go through learner.allreduce_flat over NCCL, and the
updated actor lands in each rank's CUDA actor handle (tactor_set_weights)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from mop_truss_marl_b200 import actor as actor_mod
    from mop_truss_marl_b200 import learner
    N, P, B = 16, 3, 48
    lrn = learner.MADDPGLearner(lr=1e-3, hidden=200, n_q=32, batch_size=8, device=dev, seed=5)    # same models on both ranks
    lrn.device_replay = learner.DeviceReplay(256, N, P, dev, seed=100 + rank)
    g = torch.Generator(device=dev).manual_seed(1000 + rank)                                       # different replay per rank
    r = lambda *s: torch.rand(*s, device=dev, generator=g)                                         # noqa: E731

    def state():
        return {"x_n": r(B, N, 13), "A_n": torch.full((N, N), 1.0 / N, device=dev), "A_s": r(B, N, N) / N, "A_n_ts": r(B, N, N) / N,
                "A_n_cs": r(B, N, N) / N, "x_p": r(B, P, 4), "A_p": r(B, P, P) / P}
    lrn.device_replay.push(state(), [(r(B, N, 2), r(B, N, 3)) for _ in range(3)], r(B, 3), [state() for _ in range(3)], 0.0)
    for _ in range(3):
        assert lrn.train()
    flat = torch.cat([p.detach().reshape(-1) for a in lrn.agents for m in (a.actor, a.critic) for p in m.parameters()])
    # the updated actor drives the CUDA actor of this rank
    pol = actor_mod.BatchedActor(lrn.actor_weights(0), N, B, device=dev)
    s = state()
    geo, topo = pol.forward(s["x_n"], s["A_n"], s["A_s"], s["A_n_ts"], s["A_n_cs"], s["x_p"], s["A_p"])
    pol.check()
    with torch.no_grad():
        g2, t2 = lrn.agents[0].actor(s["x_n"], s["A_n"], s["A_s"], s["A_n_ts"], s["A_n_cs"], s["x_p"], s["A_p"])
    torch.save({"flat": flat.cpu(), "elements": lrn.allreduced_elements, "backend": dist.get_backend(),
                "actor_vs_twin": float(max((geo - g2).abs().max(), (topo - t2).abs().max()))}, os.path.join(out_dir, "rank%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_nccl_gradient_allreduce_keeps_cuda_ranks_identical(tmp_path):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (torch.load(os.path.join(tmp_path, "rank%d.pt" % k)) for k in range(2))
    assert r0["backend"] == "nccl"
    assert torch.equal(r0["flat"], r1["flat"])                   # averaged gradients -> bit-identical weights on both GPUs
    assert r0["elements"] == r1["elements"] > 0
    assert r0["actor_vs_twin"] <= 2e-5 and r1["actor_vs_twin"] <= 2e-5   # CUDA actor == differentiable twin after the update


def test_captured_update_equals_eager_update():
    """MADDPGLearner.capture_graph(): one train() as a replayed CUDA graph gives the same parameters as the eager update on
    the same sampled indices (the warm-up updates of the capture are undone), here on one GPU without a process group"""
    from mop_truss_marl_b200 import learner
    dev = torch.device("cuda", 0)
    N, P, B = 16, 3, 64

    def make():
        lrn = learner.MADDPGLearner(lr=1e-3, hidden=200, n_q=32, batch_size=8, device=dev, seed=5)
        lrn.device_replay = learner.DeviceReplay(256, N, P, dev, seed=7)
        g = torch.Generator(device=dev).manual_seed(11)
        r = lambda *s: torch.rand(*s, device=dev, generator=g)                                     # noqa: E731

        def state():
            return {"x_n": r(B, N, 13), "A_n": torch.full((N, N), 1.0 / N, device=dev), "A_s": r(B, N, N) / N,
                    "A_n_ts": r(B, N, N) / N, "A_n_cs": r(B, N, N) / N, "x_p": r(B, P, 4), "A_p": r(B, P, P) / P}
        lrn.device_replay.push(state(), [(r(B, N, 2), r(B, N, 3)) for _ in range(3)], r(B, 3), [state() for _ in range(3)], 0.0)
        return lrn
    eager, graphed = make(), make()
    graphed.capture_graph()
    graphed.device_replay.gen.set_state(eager.device_replay.gen.get_state())      # the capture drew one index set
    for _ in range(4):
        assert eager.train() and graphed.train()
    torch.cuda.synchronize()
    for a, b in zip(eager.agents, graphed.agents):
        for m1, m2 in ((a.actor, b.actor), (a.critic, b.critic)):
            for p1, p2 in zip(m1.parameters(), m2.parameters()):
                assert torch.allclose(p1, p2, rtol=0, atol=1e-6), float((p1 - p2).abs().max())
