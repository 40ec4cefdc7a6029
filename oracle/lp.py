"""CPU oracle for the test-time label propagation path (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py for the
pinning status).  Restates Util/Tool.py:435-468 (LaplacianMatSym_XYZRGB_DirectComp) and
Util/ProbLabelPropagation.py:17-42 with numpy; the solve uses a dense inverse in fp64 exactly as written
(tf.linalg.inv [TF]); distances follow the SmoothConstraint flavour of oracle/knn_oracle.c."""
from __future__ import annotations

import numpy as np

from . import knn as oknn


def laplacian_sym(X: np.ndarray, RGB: np.ndarray) -> np.ndarray:
    """X, RGB (B,N,3) -> Lsym (B,N,N) fp32."""
    d1 = oknn.pairwise_distance(X, oknn.SMOOTH)                       # Tool.py:444-448 (clamped)
    d2 = oknn.pairwise_distance(RGB, oknn.SMOOTH)                     # :452-456
    W = np.exp(-d1 * np.float32(1e3)) * np.exp(-d2 * np.float32(1e1))  # :449,:457,:459
    d = W.sum(-1)                                                     # :461
    Dm = np.zeros_like(W)
    idx = np.arange(W.shape[1])
    Dm[:, idx, idx] = d + np.float32(1e-8)                            # :462
    inv = (d ** -0.5)[:, :, None] * (d ** -0.5)[:, None, :]           # :463,:465
    return ((Dm - W) * inv).astype(np.float32)


def point_weights(G: np.ndarray) -> np.ndarray:
    """ComputeWeight4EachPoint (ProbLabelPropagation.py:31-42)"""
    G = G.astype(np.float64)
    K = G.shape[-1]
    return 1.0 - (-(G * np.log(G + 1e-5) / np.log(2.0)).sum(1)) / (np.log(K) / np.log(2.0))


def solve(Lm: np.ndarray, G: np.ndarray, alpha=1.0, beta=1.0):
    """SolveLabelProp (ProbLabelPropagation.py:19-23): returns Y, Y_prob, w (fp64)."""
    N = G.shape[0]
    w = point_weights(G)
    A = alpha * Lm.astype(np.float64) + beta * np.diag(w) + 1e-5 * np.eye(N)
    Y = beta * np.linalg.inv(A) @ np.diag(w) @ G.astype(np.float64)
    return Y, Y / Y.sum(-1, keepdims=True), w
