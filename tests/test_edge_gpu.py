"""Factored EdgeConv first layer (csrc/edge.cu): y_ij = x_i (W1 - W2) + x_j W2 + b against the reference formulation
conv2d(get_edge_feature(x, idx)) (tf_util.py:674-706 feeding tf_util.py:115-173), forward and backward, in fp64 torch."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(cuda, B, N, k, Cx, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn((B, N, Cx), generator=g)
    idx = torch.randint(0, N, (B, N, k), generator=g, dtype=torch.int32)
    W = torch.randn((2 * Cx, 64), generator=g) * 0.3
    b = torch.randn((64,), generator=g) * 0.1
    return x, idx, W, b


@pytest.mark.parametrize("B,N,k,Cx", [(2, 96, 20, 64), (3, 50, 7, 9), (1, 130, 20, 3)])
def test_edge_factored_forward_backward(cuda, B, N, k, Cx):
    from weaksuppointcloudseg_b200 import _lib as L
    x, idx, W, b = _setup(cuda, B, N, k, Cx, 5 + Cx)
    P, R = B * N, B * N * k
    # ---- reference formulation in fp64
    xd = x.double().requires_grad_(True)
    Wd = W.double().requires_grad_(True)
    bd = b.double().requires_grad_(True)
    gidx = (idx.long() + (torch.arange(B).view(B, 1, 1) * N)).reshape(-1)
    xi = xd.reshape(P, Cx).repeat_interleave(k, dim=0)
    xj = xd.reshape(P, Cx)[gidx]
    y_ref = torch.cat([xi, xj - xi], -1) @ Wd + bd                                   # (R, 64)
    c1, c2, c3 = (torch.randn(64, dtype=torch.float64) for _ in range(3))
    G = torch.randn((R, 64), dtype=torch.float64)
    dy = c1 * G + c2 + c3 * y_ref.detach()
    y_ref.backward(dy)
    # ---- device path
    dev = cuda
    f32 = lambda t: t.detach().float().contiguous().to(dev)   # noqa: E731
    Wc = torch.empty((Cx, 128), device=dev)
    # every device tensor is bound to a name: a temporary would be recycled by the caching allocator before the launch
    Wg, bg, idxg, Gg, c1g, c2g, c3g = f32(W), f32(b), idx.to(dev), f32(G), f32(c1), f32(c2), f32(c3)
    L.check(L.lib().wspc_edge_split_weights(L.ptr(Wg), Cx, 64, L.ptr(Wc), L.stream()))
    assert torch.allclose(Wc.cpu(), torch.cat([W[:Cx] - W[Cx:], W[Cx:]], 1), atol=1e-7)
    UV = (f32(x).reshape(P, Cx) @ Wc).contiguous()          # the P-row GEMM itself is covered by tests/test_kernels_gpu.py
    y = torch.empty((R, 64), device=dev)
    stats = torch.zeros((2, 64), dtype=torch.float64, device=dev)
    L.check(L.lib().wspc_edge_combine_fwd(L.ptr(UV), 128, L.ptr(idxg), L.ptr(bg), P, k, N, 64, L.ptr(y), L.ptr(stats), L.stream()))
    yr = y_ref.detach()
    assert float((y.cpu().double() - yr).abs().max()) <= 2e-5 * float(yr.abs().max())
    assert torch.allclose(stats[0].cpu(), yr.sum(0), rtol=1e-5, atol=1e-3)
    assert torch.allclose(stats[1].cpu(), (yr * yr).sum(0), rtol=1e-5, atol=1e-3)
    DUV = torch.zeros((P, 128), device=dev)
    yg = f32(yr)
    L.check(L.lib().wspc_edge_combine_bwd(L.ptr(Gg), L.ptr(yg), L.ptr(c1g), L.ptr(c2g), L.ptr(c3g), L.ptr(idxg), P, k, N, 64,
                                          L.ptr(DUV), 128, L.stream()))
    dWc = (f32(x).reshape(P, Cx).t() @ DUV).contiguous()
    dW = torch.empty((2 * Cx, 64), device=dev)
    db = torch.empty(64, device=dev)
    dbc = DUV.sum(0).contiguous()
    L.check(L.lib().wspc_edge_merge_wgrad(L.ptr(dWc), L.ptr(dbc), Cx, 64, L.ptr(dW), L.ptr(db), L.stream()))
    dX = DUV @ Wc.t()
    scale = lambda t: float(t.abs().max())   # noqa: E731
    assert float((dW.cpu().double() - Wd.grad).abs().max()) <= 1e-4 * scale(Wd.grad)
    assert float((db.cpu().double() - bd.grad).abs().max()) <= 1e-4 * scale(bd.grad)
    assert float((dX.cpu().double() - xd.grad.reshape(P, Cx)).abs().max()) <= 1e-4 * scale(xd.grad)


def test_edge_bwd_without_bn(cuda):
    """c1 == NULL: dy = G (a layer without batch norm)."""
    from weaksuppointcloudseg_b200 import _lib as L
    B, N, k = 2, 64, 5
    P, R = B * N, B * N * k
    g = torch.Generator().manual_seed(1)
    G = torch.randn((R, 64), generator=g)
    idx = torch.randint(0, N, (B, N, k), generator=g, dtype=torch.int32)
    DUV = torch.zeros((P, 128), device=cuda)
    Gg, idxg = G.to(cuda), idx.to(cuda)
    L.check(L.lib().wspc_edge_combine_bwd(L.ptr(Gg), None, None, None, None, L.ptr(idxg), P, k, N, 64, L.ptr(DUV), 128,
                                          L.stream()))
    du = G.reshape(P, k, 64).sum(1)
    dv = torch.zeros((P, 64))
    gidx = (idx.long() + (torch.arange(B).view(B, 1, 1) * N)).reshape(-1)
    dv.index_add_(0, gidx, G)
    assert torch.allclose(DUV[:, :64].cpu(), du, atol=1e-5)
    assert torch.allclose(DUV[:, 64:].cpu(), dv, atol=1e-5)


def test_edge_errors_are_loud(cuda):
    from weaksuppointcloudseg_b200 import _lib as L
    t = torch.zeros(256, device=cuda)
    with pytest.raises(L.WspcError):
        L.check(L.lib().wspc_edge_split_weights(L.ptr(t), 4, 32, L.ptr(t), L.stream()))      # Cout != 64
    with pytest.raises(L.WspcError):
        L.check(L.lib().wspc_edge_combine_fwd(None, 128, L.ptr(t), None, 16, 4, 16, 64, L.ptr(t), None, L.stream()))
