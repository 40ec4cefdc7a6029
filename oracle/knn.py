"""ctypes front-end of oracle/knn_oracle.c plus a pure-numpy cross-check.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Reference call sites:
tf_util.pairwise_distance / knn (Networks/dgcnn/utils/tf_util.py:638-671),
SmoothConstraint Dmat/top_k (Util/SmoothConstraint.py:141-154).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libwspc_oracle.so")
TFUTIL, SMOOTH = 0, 1
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "knn_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _SO


def _load():
    global _lib
    if _lib is None:
        try:
            _lib = ctypes.CDLL(build())
        except OSError:
            _lib = ctypes.CDLL(build(force=True))
        f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
        i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
        ci = ctypes.c_int
        _lib.oracle_pairwise_distance.argtypes = [f32p, ci, ci, ci, ci, ci, ci, f32p]
        _lib.oracle_knn.argtypes = [f32p, ci, ci, ci, ci, ci, ci, ci, i32p, ctypes.c_void_p]
        _lib.oracle_topk_rows.argtypes = [f32p, ctypes.c_longlong, ci, ci, i32p, ctypes.c_void_p]
    return _lib


def pairwise_distance(x: np.ndarray, flavour: int = TFUTIL, coff: int = 0, D: int | None = None) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    B, N, C = x.shape
    D = C - coff if D is None else D
    adj = np.empty((B, N, N), np.float32)
    assert _load().oracle_pairwise_distance(x, B, N, C, coff, D, flavour, adj) == 0
    return adj


def knn(x: np.ndarray, k: int, flavour: int = TFUTIL, coff: int = 0, D: int | None = None, return_dist=False):
    x = np.ascontiguousarray(x, dtype=np.float32)
    B, N, C = x.shape
    D = C - coff if D is None else D
    idx = np.empty((B, N, k), np.int32)
    dist = np.empty((B, N, k), np.float32) if return_dist else None
    dp = dist.ctypes.data_as(ctypes.c_void_p) if return_dist else None
    assert _load().oracle_knn(x, B, N, C, coff, D, k, flavour, idx, dp) == 0
    return (idx, dist) if return_dist else idx


def topk_rows(adj: np.ndarray, k: int, return_vals=False):
    adj = np.ascontiguousarray(adj, dtype=np.float32)
    ncols = adj.shape[-1]
    rows = adj.size // ncols
    idx = np.empty(adj.shape[:-1] + (k,), np.int32)
    vals = np.empty(adj.shape[:-1] + (k,), np.float32) if return_vals else None
    vp = vals.ctypes.data_as(ctypes.c_void_p) if return_vals else None
    assert _load().oracle_topk_rows(adj, rows, ncols, k, idx, vp) == 0
    return (idx, vals) if return_vals else idx


# ---- independent numpy restatement (small cases only) -------------------------------------
def _fma32(a, b, c):
    """fp32 fused multiply-add emulated in fp64 (a*b is exact in fp64; one extra rounding that
    only matters on exact fp32 half-way cases)."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def pairwise_distance_numpy(x: np.ndarray, flavour: int = TFUTIL) -> np.ndarray:
    x = np.asarray(x, np.float32)
    B, N, D = x.shape
    dot = np.zeros((B, N, N), np.float32)
    sq = np.zeros((B, N), np.float32)
    for c in range(D):
        dot = _fma32(x[:, :, None, c], x[:, None, :, c], dot)
        sq = _fma32(x[:, :, c], x[:, :, c], sq)
    if flavour == TFUTIL:
        return (sq[:, :, None] + np.float32(-2.0) * dot) + sq[:, None, :]
    d = (sq[:, :, None] + sq[:, None, :]) - np.float32(2.0) * dot
    return np.where(d > 0, d, np.float32(0.0)).astype(np.float32)


def knn_numpy(adj: np.ndarray, k: int) -> np.ndarray:
    """tf.nn.top_k(-adj, k).indices == first k of a stable ascending argsort."""
    return np.argsort(adj, axis=-1, kind="stable")[..., :k].astype(np.int32)
