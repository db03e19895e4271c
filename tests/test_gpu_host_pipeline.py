"""GPU: host_pipeline.HostRollout / trollout_step_host (state tuple resident on the host, batch cut into pieces on three streams) gives
exactly what the resident path gives: environments are independent, so cutting the batch must not change a bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.mark.parametrize("family,B,pieces,compact", [("small_bridge", 77, 3, True), ("large_roof", 70, 2, True),
                                                     ("small_roof", 8, 1, True), ("small_bridge", 77, 3, False),
                                                     ("large_bridge", 40, 2, False),
                                                     ("small_bridge", 300, [32, 192, 76], True),    # trollout_set_pieces
                                                     ("small_bridge", 1500, "auto", True), ("small_roof", 60, "auto", True)])
def test_host_rollout_equals_resident_path(family, B, pieces, compact):
    """compact: the state tuple carries node_y / element_section (the two table columns _set_model reads) and uploads
    those instead of the full raw tables; without them the full tables go up.  Same bits either way."""
    from mop_truss_marl_b200 import actor, batched_env, tf_checkpoint
    from mop_truss_marl_b200.host_pipeline import HostRollout, STATE_IN, STATE_OUT
    w = tf_checkpoint.random_actor_weights(seed=11)
    envs = [batched_env.BatchedTrussEnv(family, B) for _ in range(2)]
    N = envs[0].N
    pols = [actor.BatchedActor(w, N, B, sigma=0.0, theta=0.0) for _ in range(2)]    # no exploration noise: deterministic
    for e in envs:
        e.reset()
    dev = envs[0].device
    g = torch.Generator(device=dev).manual_seed(3)
    x_p = torch.rand(B, 2, 4, device=dev, generator=g)
    A_p = torch.rand(B, 2, 2, device=dev, generator=g)
    x_p_host, A_p_host = x_p.cpu().pin_memory(), A_p.cpu().pin_memory()
    roll = HostRollout(envs[1], pols[1], pieces=pieces)
    if pieces == "auto":                                 # ~5.7 MB of download per piece: 2 pieces for 1500 small bridges, 1 for 60
        assert len(roll.ranges) == (2 if B == 1500 else 1) and all(lo % 32 == 0 for lo, _ in roll.ranges)
    else:
        assert len(roll.ranges) == (pieces if isinstance(pieces, int) else len(pieces))
    assert roll.ranges[0][0] == 0 and roll.ranges[-1][1] == B
    bufs = [roll.alloc_host(compact), roll.alloc_host(compact)]
    for k in STATE_IN:
        bufs[0][k].copy_(getattr(envs[1], k))
    torch.cuda.synchronize()
    if compact:
        HostRollout.fill_compact(bufs[0])
        for k in ("nN_x_n", "nN_x_e"):                  # not read on this path: poison them to prove it
            bufs[0][k].fill_(float("nan"))
    rng = np.random.RandomState(5)
    for it in range(4):
        coin = torch.from_numpy((rng.rand(B) >= 0.5).astype(np.uint8))
        # resident reference path
        a_geo, a_topo = pols[0].act(envs[0].x_n, envs[0].A_n, envs[0].A_s, envs[0].A_n_ts, envs[0].A_n_cs, x_p, A_p)
        a_geo_in = a_geo.clone()
        envs[0].step(a_geo, a_topo, coin.to(dev))
        # host-resident, pipelined path
        src, dst = bufs[it & 1], bufs[1 - (it & 1)]
        roll.step(src, coin.pin_memory(), x_p_host, A_p_host, dst)
        torch.cuda.synchronize()
        for k in STATE_OUT:
            assert torch.equal(dst[k], getattr(envs[0], k).cpu()), (it, k)
        assert torch.equal(dst["a_geo"], a_geo.cpu()) and torch.equal(dst["a_topo"], a_topo.cpu())
        assert torch.all((dst["a_geo"] >= 0) & (dst["a_geo"] <= 1))            # clipped in place by the step
        assert a_geo_in.shape == a_geo.shape
        if compact:
            assert torch.equal(dst["node_y"], dst["nN_x_n"][:, :, 1]) and torch.equal(dst["element_section"], dst["nN_x_e"][:, :, 0])
    h2d, d2h = roll.bytes_per_step(compact=compact)
    h2d_full, d2h_full = roll.bytes_per_step(compact=False)
    assert h2d > 0 and d2h > h2d and h2d_full >= h2d and (not compact or h2d_full - h2d == 4 * B * (11 * N + 20 * envs[0].E))


@pytest.mark.parametrize("family,B,pieces", [("small_bridge", 200, 4), ("large_bridge", 96, 2)])
def test_graph_replay_equals_direct_enqueue(family, B, pieces, monkeypatch):
    """trollout_step_host captures a buffer set into a CUDA graph the second time it sees it and replays it afterwards;
    replayed steps (OU noise included: seed and call index reach the kernels through device memory) must give the bits
    of the directly enqueued steps, and the launch counters must count the replayed kernels"""
    from mop_truss_marl_b200 import actor, batched_env, tf_checkpoint
    from mop_truss_marl_b200.host_pipeline import HostRollout, STATE_IN, STATE_OUT
    w = tf_checkpoint.random_actor_weights(seed=5)
    rolls, bufs, envs, pols = [], [], [], []
    for use_graph in (True, False):
        monkeypatch.setenv("TROLLOUT_NO_GRAPH", "0" if use_graph else "1")
        env = batched_env.BatchedTrussEnv(family, B)
        env.reset()
        pol = actor.BatchedActor(w, env.N, B, seed=123)          # default OU noise on
        roll = HostRollout(env, pol, pieces=pieces)
        b = [roll.alloc_host(), roll.alloc_host()]
        for k in STATE_IN:
            b[0][k].copy_(getattr(env, k))
        torch.cuda.synchronize()
        HostRollout.fill_compact(b[0])
        rolls.append(roll); bufs.append(b); envs.append(env); pols.append(pol)
    torch.cuda.synchronize()
    x_p = torch.tensor([1.0, 1.0, 1.0, 1.0 / 50]).repeat(B, 1, 1).contiguous().pin_memory()
    A_p = torch.ones(B, 1, 1).pin_memory()
    coin = (torch.rand(B, generator=torch.Generator().manual_seed(1)) >= 0.5).to(torch.uint8).pin_memory()
    counts = []
    for it in range(8):
        for r in range(2):
            src, dst = bufs[r][it & 1], bufs[r][1 - (it & 1)]
            rolls[r].step(src, coin, x_p, A_p, dst)
        torch.cuda.synchronize()
        for k in STATE_OUT + ("a_geo", "a_topo"):
            assert torch.equal(bufs[0][1 - (it & 1)][k], bufs[1][1 - (it & 1)][k]), (it, k)
        counts.append((envs[0].launch_count(), pols[0].launch_count(), envs[1].launch_count(), pols[1].launch_count()))
    # both arms launched the same number of kernels every step (1 env-step + 2 actor kernels per piece: the OU noise is applied inside the actor kernel)
    assert counts[-1][0] == counts[-1][2] and counts[-1][1] == counts[-1][3]
    per_step = [(b[0] - a[0], b[1] - a[1]) for a, b in zip(counts, counts[1:])]
    assert all(p == (pieces, 2 * pieces) for p in per_step), per_step
    assert int(bufs[0][0]["status"].abs().max()) == 0
    pols[0].check()


def test_graph_replay_follows_set_weights(monkeypatch):
    """a CUDA graph captured by trollout_step_host must act with the weights of the moment it is REPLAYED: the power-of-two
    scales of the fp16 operand images are read from device memory, so tactor_set_weights with 4x larger weights (every
    layer's scale exponent changes) is seen by the replayed kernels exactly as by directly enqueued ones"""
    from mop_truss_marl_b200 import actor, batched_env, tf_checkpoint
    from mop_truss_marl_b200.host_pipeline import HostRollout, STATE_IN, STATE_OUT
    family, B = "small_bridge", 96
    w0 = tf_checkpoint.random_actor_weights(seed=5)
    w1 = {k: (v[0] * 4.0, v[1] * 4.0) for k, v in tf_checkpoint.random_actor_weights(seed=6).items()}
    rolls, bufs, pols = [], [], []
    for use_graph in (True, False):
        monkeypatch.setenv("TROLLOUT_NO_GRAPH", "0" if use_graph else "1")
        env = batched_env.BatchedTrussEnv(family, B)
        env.reset()
        pol = actor.BatchedActor(w0, env.N, B, sigma=0.0, theta=0.0)
        roll = HostRollout(env, pol, pieces=2)
        b = [roll.alloc_host(), roll.alloc_host()]
        for k in STATE_IN:
            b[0][k].copy_(getattr(env, k))
        torch.cuda.synchronize()
        HostRollout.fill_compact(b[0])
        rolls.append(roll); bufs.append(b); pols.append(pol)
    x_p = torch.tensor([1.0, 1.0, 1.0, 1.0 / 50]).repeat(B, 1, 1).contiguous().pin_memory()
    A_p = torch.ones(B, 1, 1).pin_memory()
    coin = torch.zeros(B, dtype=torch.uint8).pin_memory()
    for it in range(10):
        if it == 6:                                         # both buffer sets have been captured by now
            for pol in pols:
                pol.set_weights(w1)
        for r in range(2):
            rolls[r].step(bufs[r][it & 1], coin, x_p, A_p, bufs[r][1 - (it & 1)])
        torch.cuda.synchronize()
        for k in STATE_OUT + ("a_geo", "a_topo"):
            assert torch.equal(bufs[0][1 - (it & 1)][k], bufs[1][1 - (it & 1)][k]), (it, k)
        if it == 5:
            before = bufs[0][1 - (it & 1)]["a_geo"].clone()
    assert not torch.equal(before, bufs[0][0]["a_geo"])     # the new weights really act
    rolls[0].forget_buffers()
    rolls[0].step(bufs[0][0], coin, x_p, A_p, bufs[0][1])   # works after the cached graphs were dropped
    for pol in pols:
        pol.check()


@pytest.mark.parametrize("family,B", [("small_bridge", 1024), ("large_roof", 256)])
def test_closed_loop_episode_stays_healthy(family, B):
    """BASELINE.json configs[0] runs one episode of 500 game steps: the resident actor -> env loop for a whole episode must
    keep every environment solvable (status 0), every tensor finite, the actor inside the fp16 range of its split product
    (tactor_status), and the geometry inside the reference's own bounds"""
    from mop_truss_marl_b200 import actor, batched_env, tf_checkpoint
    env = batched_env.BatchedTrussEnv(family, B)
    env.reset()
    pol = actor.BatchedActor(tf_checkpoint.random_actor_weights(seed=20), env.N, B, seed=20)
    dev = env.device
    x_p = torch.tensor([1.0, 1.0, 1.0, 1.0 / 50], device=dev).repeat(B, 1, 1).contiguous()
    A_p = torch.ones(B, 1, 1, device=dev)
    g = torch.Generator(device=dev).manual_seed(0)
    bad = torch.zeros((), dtype=torch.int32, device=dev)
    for t in range(500):
        a_geo, a_topo = pol.act(env.x_n, env.A_n, env.A_s, env.A_n_ts, env.A_n_cs, x_p, A_p)
        coin = (torch.rand(B, device=dev, generator=g) >= 0.5).to(torch.uint8)
        env.step(a_geo, a_topo, coin)
        bad |= env.status.abs().max()
    torch.cuda.synchronize()
    pol.check()
    assert int(bad) == 0
    for k in ("x_n", "A_s", "A_n_ts", "A_n_cs", "nN_x_n", "nN_x_e", "point", "move_range"):
        assert bool(torch.isfinite(getattr(env, k)).all()), k
    y = env.nN_x_n[:, :, 1]
    scal = env.handle.table("scalars")                       # y_max, y_min, d_min, ...
    assert float(y.min()) >= float(scal[1]) and float(y.max()) <= float(scal[0]) + 1e-6
    nx = env.N // 2
    assert float((y[:, nx:] - y[:, :nx]).min()) >= float(scal[2]) - 1e-6      # at least d_min deep everywhere
    assert env.game_step == 501                                # the reference counts from 1 (truss2D_ENV.py:242)


def test_state_tuple_without_raw_tables():
    """HostRollout.alloc_host(raw_tables=False): the raw tables travel as their live columns in both directions; every other
    member of the tuple is bit-identical to the full-tuple path over several fed-back steps"""
    from mop_truss_marl_b200 import actor, batched_env, tf_checkpoint
    from mop_truss_marl_b200.host_pipeline import HostRollout, STATE_IN
    B, dev = 160, torch.device("cuda", 0)
    env = batched_env.BatchedTrussEnv("small_bridge", B, device=dev)
    env.reset()
    w = tf_checkpoint.random_actor_weights(seed=3)
    pols = [actor.BatchedActor(w, env.N, B, device=dev, seed=9) for _ in range(2)]
    rolls = [HostRollout(env, pols[k], pieces=2) for k in range(2)]
    full = [rolls[0].alloc_host(), rolls[0].alloc_host()]
    slim = [rolls[1].alloc_host(raw_tables=False), rolls[1].alloc_host(raw_tables=False)]
    assert "nN_x_e" not in slim[0] and "element_section" in slim[0]
    for k in STATE_IN:
        full[0][k].copy_(getattr(env, k))
    torch.cuda.synchronize()
    HostRollout.fill_compact(full[0])
    for k in slim[0]:
        if k in full[0]:
            slim[0][k].copy_(full[0][k])
    x_p = torch.tensor([1.0, 1.0, 1.0, 1.0 / 50]).repeat(B, 1, 1).contiguous().pin_memory()
    A_p = torch.ones(B, 1, 1).pin_memory()
    rng = np.random.RandomState(1)
    for it in range(4):
        coin = torch.from_numpy((rng.rand(B) >= 0.5).astype(np.uint8)).pin_memory()
        rolls[0].step(full[it & 1], coin, x_p, A_p, full[1 - (it & 1)])
        rolls[1].step(slim[it & 1], coin, x_p, A_p, slim[1 - (it & 1)])
        torch.cuda.synchronize()
        for k in slim[1 - (it & 1)]:
            assert torch.equal(slim[1 - (it & 1)][k], full[1 - (it & 1)][k]), (it, k)
    up_f, down_f = rolls[0].bytes_per_step()
    up_s, down_s = rolls[1].bytes_per_step(raw_tables=False)
    assert up_s == up_f and down_f - down_s == 4 * B * (12 * env.N + 21 * env.E)
