"""Drop-in for Util/SmoothConstraint.py: manifold smoothness losses on a kNN graph of the input points.

`Loss_SpatialColorSmooth_add_SelfContain` (reference :130-167) is the variant the trainers call
(S3DIS_DGCNN_trainer.py:137, ShapeNet_DGCNN_trainer.py:133); it is computed by the fused CUDA kernels
(kNN with the SmoothConstraint distance flavour + graph kernel).  The sibling names are kept for API
compatibility and map onto the same kernels with the argument conventions of their reference definitions.
All functions take CUDA fp32 tensors and return a 0-dim CUDA tensor."""
from __future__ import annotations

from . import ops


def Loss_SpatialColorSmooth_add_SelfContain(Z, X, gamma=1e-1, knn=10):
    '''
    function to return spatial and color smoothness constraint loss
    :param Z: Input point cloud feature embedding float B*N*D
    :param X: Input point cloud  float B*N*6    XYZRGB
    :return: spatial smooth loss
    '''
    return ops.smooth_loss(Z, X, gamma, knn)


def Loss_SpatialSmooth_SelfContain(X, gamma=1e-1, knn=5):
    '''reference :36-67 — the smoothed quantity and the graph are both X (B*N*3)'''
    return ops.smooth_loss(X, X, gamma, knn)


def Loss_SpatialColorSmooth_SelfContain(Z, X, gamma=1e-1, knn=10):
    '''reference :70-128 — separate XYZ and RGB graphs, losses added'''
    return ops.smooth_loss(Z, X[:, :, 0:3].contiguous(), gamma, knn) + ops.smooth_loss(Z, X[:, :, 3:6].contiguous(), gamma, knn)


def Loss_SpatialColorSmoothAdd_UnknownBatch_SelfContain(Z, X, gamma=1e-1, knn=10):
    '''reference :169-219 — the `_add_` loss for graphs whose batch size is not static; identical arithmetic'''
    return ops.smooth_loss(Z, X, gamma, knn)


def Loss_SpatialSmooth(X, W, Ind):
    Z = X
    '''reference :9-33 — weights W (B,N,knn) and indices Ind (B,N,knn) supplied by the caller:
    mean_{b,n,j} W * mean_c (Z_i - Z_j)^2.  W = exp(-d/gamma) is inverted to the distance the kernel expects.'''
    import torch
    from . import _lib as L
    Zc = Z.contiguous()
    B, N, C = Zc.shape
    knn = Ind.shape[-1]
    dist = (-torch.log(W.clamp_min(1e-38))).contiguous()       # kernel computes exp(-dist/1)
    loss = torch.empty(1, dtype=torch.float32, device=Zc.device)
    ws = L.workspace(256, Zc.device, "smooth")
    L.check(L.lib().wspc_smooth_loss(L.ptr(Zc), L.ptr(Ind.to(torch.int32).contiguous()), L.ptr(dist), B, N, C, knn, 1.0, None,
                                     L.ptr(loss), L.ptr(ws), ws.numel(), L.stream()))
    return loss[0]


def ComputeW(Target, X, knn):
    '''reference :224-230 — `Target` was (sess, {'W': graph}) there and is ignored; returns the kNN weights
    W = exp(-d/0.1) (B,N,knn) of the input points'''
    import torch
    _, dist = ops.knn_fused(torch.as_tensor(X, dtype=torch.float32).cuda().contiguous(), knn, ops.DIST_SMOOTH,
                            return_dist=True)
    return torch.exp(-dist / 0.1)
