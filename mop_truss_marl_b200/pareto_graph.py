"""Host-side helper: the Pareto graph of a full front as ``pareto_state_data`` builds it
(``test/00_small_bridge/code/truss2D_ENV.py:22-41``), written out here so that the bench's product path does not import
``oracle/``."""
import numpy as np


def chain_graph(P, index=0, max_front=50):
    """x_p [1,P,4] (obj1 ascending, obj2 descending on [0,1], one-hot of `index`, P / MAX_FRONT) and the symmetrically
    normalised chain adjacency with self loops A_p [1,P,P], float32"""
    x = np.zeros((1, P, 4), np.float32)
    x[0, :, 0] = np.linspace(0.1, 0.9, P, dtype=np.float32)
    x[0, :, 1] = np.linspace(0.9, 0.1, P, dtype=np.float32)
    x[0, index, 2] = 1.0
    x[0, :, 3] = P / max_front
    A = np.eye(P, dtype=np.float32)
    for i in range(P - 1):
        A[i, i + 1] = A[i + 1, i] = 1.0
    d = np.power(A.sum(1), -0.5).astype(np.float32)
    return x, (d[:, None] * A * d[None, :])[None].astype(np.float32)
