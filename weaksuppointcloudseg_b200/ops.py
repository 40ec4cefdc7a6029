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


def batch_gather(X: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """Tool.batch_gather_v1 (Util/Tool.py:72-104): X (B,N,D), idx (B,N,k) -> (B,N,k,D)."""
    X = _as_bnc(X)
    L.require_cuda(X, idx)
    B, N, D = X.shape
    k = idx.shape[-1]
    idx = idx.to(torch.int32).contiguous()
    out = torch.empty((B, N, k, D), dtype=torch.float32, device=X.device)
    L.check(L.lib().wspc_gather(L.ptr(X), L.ptr(idx), B, N, k, D, D, 0, L.ptr(out), L.stream()))
    return out


def get_edge_feature(point_cloud: torch.Tensor, nn_idx: torch.Tensor, k: int = 20) -> torch.Tensor:
    """tf_util.get_edge_feature (tf_util.py:674-706): (B,N,[1,]C) + (B,N,k) -> (B,N,k,2C) = [x_i | x_j - x_i]."""
    X = _as_bnc(point_cloud).contiguous()
    L.require_cuda(X, nn_idx)
    B, N, C = X.shape
    idx = nn_idx.to(torch.int32).contiguous()
    out = torch.empty((B, N, idx.shape[-1], 2 * C), dtype=torch.float32, device=X.device)
    L.check(L.lib().wspc_gather(L.ptr(X), L.ptr(idx), B, N, idx.shape[-1], C, C, 1, L.ptr(out), L.stream()))
    return out


def smooth_loss(Z: torch.Tensor, X: torch.Tensor, gamma: float = 1e-1, knn: int = 10, want_grad: bool = False):
    """SmoothConstraint.Loss_SpatialColorSmooth_add_SelfContain (Util/SmoothConstraint.py:130-167)."""
    Z, X = _as_bnc(Z).contiguous(), _as_bnc(X).contiguous()
    L.require_cuda(Z, X)
    B, N, C = Z.shape
    idx, dist = knn_fused(X, knn, DIST_SMOOTH, return_dist=True)
    loss = torch.empty(1, dtype=torch.float32, device=Z.device)
    dZ = torch.zeros_like(Z) if want_grad else None
    ws = L.workspace(256, Z.device, "smooth")
    L.check(L.lib().wspc_smooth_loss(L.ptr(Z), L.ptr(idx), L.ptr(dist), B, N, C, knn, gamma, L.ptr(dZ), L.ptr(loss),
                                     L.ptr(ws), ws.numel(), L.stream()))
    return (loss[0], dZ) if want_grad else loss[0]


SMOOTH_SUM_C, SMOOTH_WEIGHTS, SMOOTH_GLOBAL_SS = 1, 2, 4     # include/wspc.h WSPC_SMOOTH_*


def smooth_loss_graph(Z: torch.Tensor, idx: torch.Tensor, dist_or_w: torch.Tensor, gamma: float = 1e-1, flags: int = 0,
                      idx_match: torch.Tensor = None, want_grad: bool = False):
    """The smoothness term on a GIVEN kNN graph (wspc_smooth_loss_ex): the variants of Util/SmoothConstraint.py differ in
    channel reduction (sum / mean), where the weights come from and an optional slot mask."""
    Z = _as_bnc(Z).contiguous()
    idx = idx.to(torch.int32).contiguous()
    dist_or_w = dist_or_w.to(torch.float32).contiguous()
    L.require_cuda(Z, idx, dist_or_w)
    B, N, C = Z.shape
    knn = idx.shape[-1]
    if tuple(idx.shape) != (B, N, knn) or tuple(dist_or_w.shape) != (B, N, knn):
        raise L.WspcError(f"smooth_loss_graph: idx {tuple(idx.shape)} / weights {tuple(dist_or_w.shape)} do not match Z {tuple(Z.shape)}")
    if idx_match is not None:
        idx_match = idx_match.to(torch.int32).contiguous()
        L.require_cuda(idx_match)
        if idx_match.shape != idx.shape:
            raise L.WspcError("smooth_loss_graph: idx_match must have the shape of idx")
    loss = torch.empty(1, dtype=torch.float32, device=Z.device)
    dZ = torch.zeros_like(Z) if want_grad else None
    ws = L.workspace(256, Z.device, "smooth")
    L.check(L.lib().wspc_smooth_loss_ex(L.ptr(Z), L.ptr(idx), L.ptr(dist_or_w), L.ptr(idx_match), B, N, C, knn, float(gamma),
                                        int(flags), L.ptr(dZ), L.ptr(loss), L.ptr(ws), ws.numel(), L.stream()))
    return (loss[0], dZ) if want_grad else loss[0]


def laplacian_sym(X, RGB, scale_xyz: float = 1e3, scale_rgb: float = 1e1) -> torch.Tensor:
    """TF_Computation.LaplacianMatSym_XYZRGB_DirectComp.Eval (Util/Tool.py:435-468): (B,N,3),(B,N,3) -> (B,N,N)."""
    X = torch.as_tensor(X, dtype=torch.float32)
    RGB = torch.as_tensor(RGB, dtype=torch.float32)
    dev = X.device if X.is_cuda else torch.device("cuda", torch.cuda.current_device())
    X, RGB = X.to(dev).contiguous(), RGB.to(dev).contiguous()
    B, N, D1 = X.shape
    D2 = RGB.shape[-1]
    out = torch.empty((B, N, N), dtype=torch.float32, device=dev)
    deg = torch.empty((B, N), dtype=torch.float32, device=dev)
    L.check(L.lib().wspc_laplacian_sym(L.ptr(X), L.ptr(RGB), B, N, D1, D2, scale_xyz, scale_rgb, L.ptr(deg), L.ptr(out),
                                       L.stream()))
    return out


LP_MAX_ITER = 600      # Jacobi-PCG needs 10-20 iterations on confident predictions and ~150 when w ~ 0 (near-uniform G)


def lp_blocks_on_graph(Lmat, G, alpha: float = 1.0, beta: float = 1.0, max_iter: int = LP_MAX_ITER, tol: float = 1e-6):
    """LabelPropagation_TF.SolveLabelProp (Util/ProbLabelPropagation.py:44-57) for a batch of blocks: L (B,N,N), G (B,N,K) ->
    Y, Y_prob (B,N,K), w (B,N), info {iters, resid, converged} (device tensors; nothing here waits for the GPU)."""
    Lmat, G = Lmat.contiguous(), G.to(torch.float32).contiguous()
    L.require_cuda(Lmat, G)
    B, N, K = G.shape
    if tuple(Lmat.shape) != (B, N, N):
        raise L.WspcError(f"lp_blocks: L {tuple(Lmat.shape)} does not match G {tuple(G.shape)}")
    dev = G.device
    Y = torch.empty((B, N, K), dtype=torch.float32, device=dev)
    Yp = torch.empty((B, N, K), dtype=torch.float32, device=dev)
    w = torch.empty((B, N), dtype=torch.float32, device=dev)
    iters = torch.empty((B,), dtype=torch.int32, device=dev)
    done = torch.empty((B,), dtype=torch.int32, device=dev)
    resid = torch.empty((B,), dtype=torch.float32, device=dev)
    ws = L.workspace(L.lib().wspc_lp_blocks_workspace_bytes(B, N, K, max_iter), dev, "lp")
    L.check(L.lib().wspc_lp_blocks(L.ptr(Lmat), L.ptr(G), B, N, K, alpha, beta, max_iter, tol, L.ptr(Y), L.ptr(Yp), L.ptr(w),
                                   L.ptr(iters), L.ptr(resid), L.ptr(done), L.ptr(ws), ws.numel(), L.stream()))
    return Y, Yp, w, {"iters": iters, "resid": resid, "converged": done}


def lp_blocks(xyz, rgb, G, alpha: float = 1.0, beta: float = 1.0, max_iter: int = LP_MAX_ITER, tol: float = 1e-6,
              scale_xyz: float = 1e3, scale_rgb: float = 1e1):
    """The test-time stage of S3DIS_Trainer.Test / ShapeNet_Trainer.Test for every block of a batch at once: symmetric
    Laplacian of (xyz, rgb) (Util/Tool.py:435-468) -> label propagation of the network's probabilities G.  The (B,N,N)
    Laplacians live in a cached workspace (67 MB per block at N = 4096)."""
    xyz, rgb, G = xyz.to(torch.float32).contiguous(), rgb.to(torch.float32).contiguous(), G.to(torch.float32).contiguous()
    L.require_cuda(xyz, rgb, G)
    B, N, D1 = xyz.shape
    D2 = rgb.shape[-1]
    buf = L.workspace(B * N * N * 4 + B * N * 4, xyz.device, "lp_laplacian")
    Lm = buf[:B * N * N * 4].view(torch.float32).view(B, N, N)
    deg = buf[B * N * N * 4:B * N * N * 4 + B * N * 4].view(torch.float32)
    L.check(L.lib().wspc_laplacian_sym(L.ptr(xyz), L.ptr(rgb), B, N, D1, D2, scale_xyz, scale_rgb, L.ptr(deg), L.ptr(Lm),
                                       L.stream()))
    return lp_blocks_on_graph(Lm, G, alpha, beta, max_iter, tol)


def lp_solve(Lmat, G, alpha: float = 1.0, beta: float = 1.0, max_iter: int = LP_MAX_ITER, tol: float = 1e-6):
    """LabelPropagation_TF.SolveLabelProp (Util/ProbLabelPropagation.py:44-57): L (N,N), G (N,K) -> Y, Y_prob, w.
    `lp_solve.last_info` holds the device-side {iters, resid, converged} of the call."""
    Lmat = torch.as_tensor(Lmat, dtype=torch.float32)
    G = torch.as_tensor(G, dtype=torch.float32)
    dev = Lmat.device if Lmat.is_cuda else torch.device("cuda", torch.cuda.current_device())
    Y, Yp, w, info = lp_blocks_on_graph(Lmat.to(dev)[None], G.to(dev)[None], alpha, beta, max_iter, tol)
    lp_solve.last_info = info
    return Y[0], Yp[0], w[0]
