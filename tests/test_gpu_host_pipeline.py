"""GPU: host_pipeline.HostRollout / trollout_step_host (state tuple resident on the host, batch cut into pieces on three streams) gives
exactly what the resident path gives: environments are independent, so cutting the batch must not change a bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.mark.parametrize("family,B,pieces", [("small_bridge", 77, 3), ("large_roof", 70, 2), ("small_roof", 8, 1)])
def test_host_rollout_equals_resident_path(family, B, pieces):
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
    assert len(roll.ranges) == pieces and roll.ranges[0][0] == 0 and roll.ranges[-1][1] == B
    bufs = [roll.alloc_host(), roll.alloc_host()]
    for k in STATE_IN:
        bufs[0][k].copy_(getattr(envs[1], k))
    torch.cuda.synchronize()
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
    h2d, d2h = roll.bytes_per_step()
    assert h2d > 0 and d2h > h2d
