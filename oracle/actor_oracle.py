"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the node-agent GCN actor forward pass.

Follows ``multimodes_actor.call`` (``train/code/truss2D_RL.py:75-127``) with spektral 1.2.0's published
``GCNConv`` (``output = A . (X . W) + b``, ``use_bias=True``, no activation) and ``GlobalSumPool`` (sum over
the node axis); ``act`` adds Ornstein-Uhlenbeck noise to every entry (``:41-48, 328-354``).

Parity status: **unpinned** -- TensorFlow / spektral are not installable here and the reference holds no
recorded actor output, so this restatement can only be checked against itself (shapes, determinism,
float32 CUDA vs float64 numpy agreement).  SURVEY.md section 8c.
"""
from __future__ import annotations

import numpy as np


def _gcn(X, A, W, b):
    return np.matmul(A, np.matmul(X, W)) + b


def _relu(x):
    return np.maximum(x, 0)


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def actor_forward(weights, x_n, A_n, A_s, A_n_ts, A_n_cs, x_p, A_p, n_pf=None, dtype=np.float64):
    """Batched forward.  x_n [B,N,13]; A_* [B,N,N] (A_n may be [N,N]); x_p [B,P,4]; A_p [B,P,P];
    n_pf [B] = number of valid Pareto rows (rows >= n_pf[b] are ignored, i.e. ``act`` on the unpadded
    graph).  Returns (geo [B,N,2], topo [B,N,3])."""
    w = {k: (np.asarray(a, dtype=dtype), np.asarray(b, dtype=dtype)) for k, (a, b) in weights.items()}
    x_n = np.asarray(x_n, dtype=dtype)
    B, N, _ = x_n.shape
    A_n = np.broadcast_to(np.asarray(A_n, dtype=dtype), (B, N, N))
    A_s, A_ts, A_cs = (np.asarray(a, dtype=dtype) for a in (A_s, A_n_ts, A_n_cs))
    x_p, A_p = np.asarray(x_p, dtype=dtype), np.asarray(A_p, dtype=dtype)
    x11 = _relu(_gcn(x_n, A_n, *w["gcn_l1_1"]))
    x12 = _relu(_gcn(x_n, A_n, *w["gcn_l1_2"]))
    x13 = _relu(_gcn(x_n, A_n, *w["gcn_l1_3"]))
    x14 = _relu(_gcn(x_p, A_p, *w["gcn_l1_4"]))                     # [B,P,H]
    if n_pf is not None:
        valid = (np.arange(x_p.shape[1])[None, :] < np.asarray(n_pf)[:, None]).astype(dtype)
        x14 = x14 * valid[:, :, None]
    pooled = x14.sum(axis=1)                                         # GlobalSumPool -> [B,H]
    H = pooled.shape[1]
    # tf.ragged.stack([pooled]*N, axis=-1) -> [B,H,N]; tf.reshape(..., (B,N,H)): NOT a transpose (:89-95)
    x14b = np.stack([pooled] * N, axis=-1).reshape(B, N, H)
    x21 = _relu(_gcn(x11, A_n, *w["gcn_l2_1"]))
    x22 = _relu(_gcn(x12, A_ts, *w["gcn_l2_2"]))
    x23 = _relu(_gcn(x12, A_cs, *w["gcn_l2_3"]))
    x24 = _relu(_gcn(x13, A_s, *w["gcn_l2_4"]))
    x25 = _relu(_gcn(x14b, A_n, *w["gcn_l2_5"]))
    s = x21 + x22 + x23 + x24 + x25
    x31 = _relu(_gcn(s, A_n, *w["gcn_l3_1"]))
    x32 = _relu(_gcn(s, A_s, *w["gcn_l3_2"]))
    geo = _sigmoid(_gcn(x31, A_n, *w["gcn_l4_1"]))
    topo = _sigmoid(_gcn(x32, A_n, *w["gcn_l4_2"]))
    return geo, topo


class OUNoise:
    """``OUNoise.gen_noise`` (truss2D_RL.py:41-48): theta*(mu-x)*dt + sigma*np.random.randn(1)"""

    def __init__(self, mu, theta, sigma, rng=None):
        self.mu, self.theta, self.sigma, self.dt = mu, theta, sigma, 0.0001
        self.rng = rng if rng is not None else np.random

    def gen_noise(self, x):
        return self.theta * (self.mu - x) * self.dt + self.sigma * self.rng.randn(1)


def act(weights, state, mu=0.1, theta=0.1, sigma=0.1, rng=None):
    """``multimodals_OneAgent.act`` for ONE environment: forward + OU noise on every entry, drawn
    row-major geo first then topo (truss2D_RL.py:341-350).  Returns float32 arrays like the reference."""
    x_n, A_n, A_s, A_ts, A_cs, x_p, A_p = state
    geo, topo = actor_forward(weights, x_n[None], A_n[None], A_s[None], A_ts[None], A_cs[None], x_p[None], A_p[None],
                              dtype=np.float32)
    geo, topo = geo[0].astype(np.float32), topo[0].astype(np.float32)
    ng, nt = OUNoise(mu, theta, sigma, rng), OUNoise(mu, theta, sigma, rng)
    for i in range(geo.shape[0]):
        for j in range(geo.shape[1]):
            geo[i, j] += ng.gen_noise(geo[i, j])[0]
    for i in range(topo.shape[0]):
        for j in range(topo.shape[1]):
            topo[i, j] += nt.gen_noise(topo[i, j])[0]
    return geo, topo
