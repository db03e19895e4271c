"""TF-free reader of the reference's actor checkpoints (host logic)."""
import os

import numpy as np
import pytest

from mop_truss_marl_b200 import tf_checkpoint
from oracle import ref_harness

CKPT = os.path.join(ref_harness.REF_ROOT, "model", "2000pickle_base")


@pytest.mark.reference
@pytest.mark.parametrize("agent", [1, 2, 3])
def test_reads_reference_actor_checkpoint(agent):
    w = tf_checkpoint.load_actor_weights(os.path.join(CKPT, "Agent%d_Actor_pickle" % agent))
    shapes = {k: (a.shape, b.shape) for k, (a, b) in w.items()}
    assert shapes["gcn_l1_1"] == ((13, 200), (200,)) and shapes["gcn_l1_4"] == ((4, 200), (200,))
    assert shapes["gcn_l2_5"] == ((200, 200), (200,)) and shapes["gcn_l4_1"] == ((200, 2), (2,))
    assert shapes["gcn_l4_2"] == ((200, 3), (3,))
    assert sum(a.size + b.size for a, b in w.values()) == 291805
    assert all(np.isfinite(a).all() and np.isfinite(b).all() for a, b in w.values())
    assert 0.05 < w["gcn_l2_1"][0].std() < 0.09                           # Glorot-scale kernels


@pytest.mark.reference
def test_critic_data_missing_is_reported():
    with pytest.raises(FileNotFoundError):
        tf_checkpoint.load_checkpoint(os.path.join(CKPT, "Agent1_Critic_pickle"))


def test_random_weights_have_reference_shapes():
    w = tf_checkpoint.random_actor_weights(0)
    assert list(w) == list(tf_checkpoint.ACTOR_LAYERS)
    assert sum(a.size + b.size for a, b in w.values()) == 291805


def test_actor_oracle_scramble_quirk():
    """x14b[b,n,h] = pooled[b,(n*200+h)//N] (tf.reshape instead of transpose, truss2D_RL.py:89-95)"""
    from oracle.actor_oracle import actor_forward
    rng = np.random.RandomState(0)
    w = tf_checkpoint.random_actor_weights(0)
    N, B, P = 16, 2, 3
    z = np.zeros
    geo, topo = actor_forward(w, rng.rand(B, N, 13), np.eye(N), z((B, N, N)), z((B, N, N)), z((B, N, N)),
                              rng.rand(B, P, 4), rng.rand(B, P, P))
    assert geo.shape == (B, N, 2) and topo.shape == (B, N, 3) and np.all((geo > 0) & (geo < 1))
