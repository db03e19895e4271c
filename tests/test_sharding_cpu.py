"""CPU, world_size 2 (gloo): the multi-GPU path shards environments with no data-path collective; the only
collectives are bench.py's barrier and max-over-ranks of the timings.  This test runs that host logic."""
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist          # noqa: E402
import torch.multiprocessing as mp        # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def shard_bounds(total, world, rank):
    """contiguous block of environments per rank"""
    per = total // world
    extra = total % world
    lo = rank * per + min(rank, extra)
    return lo, lo + per + (1 if rank < extra else 0)


def _worker(rank, world, port, total, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_bounds(total, world, rank)
    # every rank steps its own slice with the oracle as a stand-in for the device (host logic under test:
    # slicing, seeding, barrier, max-over-ranks reduction of the step time)
    from oracle.truss_oracle import TrussOracle
    o = TrussOracle("small_bridge")
    rng = np.random.RandomState(0)
    a_geo = rng.rand(total, 16, 2).astype(np.float32)
    a_topo = rng.rand(total, 16, 3).astype(np.float32)
    st0 = o.reset()
    pts = []
    for b in range(lo, hi):
        out = o.step(st0["nN_x_n"], st0["nN_x_e"], st0["max_up"], st0["max_down"], a_geo[b].copy(), a_topo[b].copy(), b % 2)
        pts.append(out["point"])
    np.save(os.path.join(out_dir, "pts_%d.npy" % rank), np.array(pts))
    dist.barrier()
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)      # pretend step time
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == float(world)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding(tmp_path):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    total, world = 7, 2
    mp.spawn(_worker, args=(world, port, total, str(tmp_path)), nprocs=world, join=True)
    got = np.concatenate([np.load(os.path.join(tmp_path, "pts_%d.npy" % r)) for r in range(world)])
    assert got.shape == (total, 4)
    # single-process result over the whole batch is identical (independent environments)
    sys.path.insert(0, ROOT)
    from oracle.truss_oracle import TrussOracle
    o = TrussOracle("small_bridge")
    rng = np.random.RandomState(0)
    a_geo = rng.rand(total, 16, 2).astype(np.float32); a_topo = rng.rand(total, 16, 3).astype(np.float32)
    st0 = o.reset()
    for b in range(total):
        out = o.step(st0["nN_x_n"], st0["nN_x_e"], st0["max_up"], st0["max_down"], a_geo[b].copy(), a_topo[b].copy(), b % 2)
        assert np.array_equal(out["point"], got[b])


def test_shard_bounds_cover_everything():
    for total in (0, 1, 7, 4096, 16384):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
