"""Thin tensor-level wrappers over the C ABI (include/wspc.h).

Each function checks shapes, allocates outputs with torch (device memory and
streams are torch's job; the arithmetic is libwspc's) and calls the matching
``wspc_*`` entry point on the current stream.
"""
from __future__ import annotations

import torch

from . import _lib as L
from ._lib import DIST_SMOOTH, DIST_TFUTIL  # noqa: F401


def _as_bnc(x: torch.Tensor):
    """(B,N,C) or (B,N,1,C) -> (B,N,C) view (reference squeeze quirk: tf_util.py:647-650)."""
    if x.dim() == 4:
        if x.shape[2] != 1:
            raise L.WspcError(f"expected (B,N,1,C), got {tuple(x.shape)}")
        x = x[:, :, 0, :]
    if x.dim() != 3:
        raise L.WspcError(f"expected (B,N,C), got {tuple(x.shape)}")
    return x


def knn_fused(x: torch.Tensor, k: int, flavour: int = DIST_TFUTIL, coff: int = 0, D: int | None = None,
              return_dist: bool = False):
    """Fused distance + kNN on channels [coff, coff+D) of x (B,N,C). -> idx int32 (B,N,k)[, dist]."""
    x = _as_bnc(x)
    L.require_cuda(x)
    if x.dtype != torch.float32:
        raise L.WspcError("knn_fused takes fp32")
    B, N, C = x.shape
    D = C - coff if D is None else D
    idx = torch.empty((B, N, k), dtype=torch.int32, device=x.device)
    dist = torch.empty((B, N, k), dtype=torch.float32, device=x.device) if return_dist else None
    nbytes = L.lib().wspc_knn_workspace_bytes(B, N, D)
    ws = L.workspace(nbytes, x.device, "knn")
    L.check(L.lib().wspc_knn_fused(L.ptr(x), B, N, C, coff, D, k, flavour, L.ptr(idx), L.ptr(dist), L.ptr(ws),
                                   ws.numel(), L.stream()))
    return (idx, dist) if return_dist else idx


def pairwise_distance(x: torch.Tensor, flavour: int = DIST_TFUTIL, coff: int = 0, D: int | None = None):
    x = _as_bnc(x)
    L.require_cuda(x)
    B, N, C = x.shape
    D = C - coff if D is None else D
    adj = torch.empty((B, N, N), dtype=torch.float32, device=x.device)
    nbytes = L.lib().wspc_knn_workspace_bytes(B, N, D)
    ws = L.workspace(nbytes, x.device, "knn")
    L.check(L.lib().wspc_pairwise_distance(L.ptr(x), B, N, C, coff, D, flavour, L.ptr(adj), L.ptr(ws), ws.numel(),
                                           L.stream()))
    return adj


def topk_rows(adj: torch.Tensor, k: int, return_vals: bool = False):
    """k smallest per row (== tf.nn.top_k(-adj, k)), ascending, ties -> lower index."""
    L.require_cuda(adj)
    ncols = adj.shape[-1]
    rows = adj.numel() // ncols
    idx = torch.empty(adj.shape[:-1] + (k,), dtype=torch.int32, device=adj.device)
    vals = torch.empty(adj.shape[:-1] + (k,), dtype=torch.float32, device=adj.device) if return_vals else None
    L.check(L.lib().wspc_topk_rows(L.ptr(adj), rows, ncols, k, L.ptr(idx), L.ptr(vals), L.stream()))
    return (idx, vals) if return_vals else idx
