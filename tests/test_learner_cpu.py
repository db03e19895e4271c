"""CPU: the PyTorch-level DDPG learner (mop_truss_marl_b200/learner.py) -- network restatements against the numpy
actor oracle, the update rules of MADDPG.train (truss2D_RL.py:463-689), and the flat-buffer gradient all-reduce on a
world_size-2 gloo group (the only collective of the system, BASELINE.json configs[4])."""
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist          # noqa: E402
import torch.multiprocessing as mp        # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mop_truss_marl_b200 import learner, tf_checkpoint          # noqa: E402


def random_state(rng, N, P, batch=None):
    lead = () if batch is None else (batch,)
    A = (rng.rand(*lead, N, N) < 0.3).astype(np.float32)
    return (rng.rand(*lead, N, 13).astype(np.float32), (rng.rand(N, N) * 0.3).astype(np.float32) if batch is None
            else np.broadcast_to((rng.rand(N, N) * 0.3).astype(np.float32), (batch, N, N)).copy(),
            (A * rng.rand(*lead, N, N)).astype(np.float32), (A * rng.rand(*lead, N, N) * 0.5).astype(np.float32),
            (A * rng.rand(*lead, N, N) * 0.5).astype(np.float32), rng.rand(*lead, P, 4).astype(np.float32),
            (rng.rand(*lead, P, P) * 0.5).astype(np.float32))


def test_actor_net_matches_numpy_oracle():
    from oracle.actor_oracle import actor_forward
    rng = np.random.RandomState(0)
    w = tf_checkpoint.random_actor_weights(seed=4)
    for k in w:
        w[k] = (w[k][0], (rng.randn(*w[k][1].shape) * 0.05).astype(np.float32))
    net = learner.ActorNet()
    net.import_weights(w)
    for N, P in ((16, 1), (32, 7)):
        st = random_state(rng, N, P, batch=5)
        geo, topo = net(*(torch.from_numpy(s) for s in st))
        g64, t64 = actor_forward(w, *st)
        assert np.abs(geo.detach().numpy() - g64).max() < 2e-5 and np.abs(topo.detach().numpy() - t64).max() < 2e-5
    back = net.export_weights()
    assert all(np.array_equal(back[k][0], w[k][0]) and np.array_equal(back[k][1], w[k][1]) for k in w)


def test_critic_net_structure():
    rng = np.random.RandomState(1)
    c = learner.CriticNet(hidden=200, n_q=200)
    names = [n for n, _ in c.named_parameters()]
    assert len(c.l1) == 10 and len(c.l2) == 11 and c.dense_1.in_features == 2200 and c.dense_out.out_features == 1
    assert [m.kernel.shape[0] for m in c.l1] == [13, 13, 13, 4, 2, 3, 2, 3, 2, 3] and len(names) == 2 * 21 + 6
    N, B = 16, 3
    st = [torch.from_numpy(s) for s in random_state(rng, N, 2, batch=B)]
    acts = [torch.rand(B, N, k) for k in (2, 3, 2, 3, 2, 3)]
    q = c(*st, *acts)
    assert q.shape == (B, 1) and torch.isfinite(q).all()
    # the Pareto embedding is tiled and reshaped, not transposed (truss2D_RL.py:190-195)
    pooled = torch.arange(6.0).view(1, 6)
    assert torch.equal(learner._tile_reshape(pooled, 3)[0],
                       torch.from_numpy(np.stack([pooled[0].numpy()] * 3, axis=-1).reshape(3, 6)))


def fill(lrn, rng, n, N=16, P=1):
    for _ in range(n):
        s = random_state(rng, N, P)
        acts = [(rng.rand(N, 2).astype(np.float32), rng.rand(N, 3).astype(np.float32)) for _ in range(3)]
        lrn.remember(s, acts, rng.randn(3).astype(np.float32), [random_state(rng, N, P) for _ in range(3)], int(rng.rand() < 0.1))


def test_train_step_rules():
    rng = np.random.RandomState(2)
    lrn = learner.MADDPGLearner(lr=1e-3, hidden=32, n_q=16, batch_size=8, seed=3)
    assert lrn.train() is False                                  # fewer than batch_size transitions (:491-494)
    fill(lrn, rng, 20)
    before_a = [p.detach().clone() for p in lrn.agents[0].actor.parameters()]
    before_c = [p.detach().clone() for p in lrn.agents[0].critic.parameters()]
    before_t = [p.detach().clone() for p in lrn.agents[0].target_actor.parameters()]
    assert lrn.train() is True
    da = max(float((p - b).abs().max()) for p, b in zip(lrn.agents[0].actor.parameters(), before_a))
    dc = max(float((p - b).abs().max()) for p, b in zip(lrn.agents[0].critic.parameters(), before_c))
    # a NEW Adam every call: the actor moves by at most lr * 0.1 per weight (:625); the critic by at most ~lr
    assert 0 < da <= 1e-3 * 0.1 * 1.01 and 0 < dc <= 1e-3 * 1.01
    assert all(torch.equal(p, b) for p, b in zip(lrn.agents[0].target_actor.parameters(), before_t))   # targets untouched
    # soft update only when an agent's own counter hits 1000, checked every 300 calls of agent 0 (:392-398, :692-697)
    assert lrn.update() == [False, False, False]
    lrn.agents[0].update_num = 1
    assert lrn.update() is None
    for a in lrn.agents:
        a.update_num = 1000
    lrn.agents[0].update_num = 1000 + 200                       # 1200 % 300 == 0 but != 1000
    assert lrn.update() == [False, True, True] and lrn.agents[1].update_num == 0
    moved = max(float((p - b).abs().max()) for p, b in zip(lrn.agents[1].target_actor.parameters(),
                                                          lrn.agents[1].actor.parameters()))
    assert moved > 0                                            # blended by tau, not copied
    # the critic regression reduces its loss on a fixed replay
    losses = []
    for _ in range(15):
        lrn.rng.seed(0)
        lrn.train()
        losses.append(lrn.last_losses[0][0])
    assert losses[-1] < losses[0]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lrn = learner.MADDPGLearner(lr=1e-3, hidden=16, n_q=8, batch_size=4, seed=5)     # same initial weights on every rank
    fill(lrn, np.random.RandomState(100 + rank), 12)                                  # different replay per rank
    for _ in range(3):
        assert lrn.train()
    flat = torch.cat([p.detach().reshape(-1) for a in lrn.agents for m in (a.actor, a.critic) for p in m.parameters()])
    torch.save({"flat": flat, "elements": lrn.allreduced_elements}, os.path.join(out_dir, "rank%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_allreduce_keeps_ranks_identical(tmp_path):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (torch.load(os.path.join(tmp_path, "rank%d.pt" % r)) for r in range(2))
    assert torch.equal(r0["flat"], r1["flat"])                   # averaged gradients -> identical weights
    n_params = sum(p.numel() for m in (learner.ActorNet(16), learner.CriticNet(16, 8)) for p in m.parameters())
    assert r0["elements"] == 3 * 3 * n_params                    # one flat all-reduce per model update: 3 steps x 3 agents
    # without the collective the two replays would have driven the ranks apart
    solo = learner.MADDPGLearner(lr=1e-3, hidden=16, n_q=8, batch_size=4, seed=5)
    fill(solo, np.random.RandomState(100), 12)
    for _ in range(3):
        solo.train()
    flat = torch.cat([p.detach().reshape(-1) for a in solo.agents for m in (a.actor, a.critic) for p in m.parameters()])
    assert not torch.equal(flat, r0["flat"])
