"""Drop-in for Util/SmoothConstraint.py: manifold smoothness losses on kNN graphs of the input points.

`Loss_SpatialColorSmooth_add_SelfContain` (reference :130-167) is the variant the trainers call
(S3DIS_DGCNN_trainer.py:137, ShapeNet_DGCNN_trainer.py:133) and the one fused into the training engines.  The sibling
definitions of the reference module differ from it in the channel reduction (sum instead of mean), in where the edge weights
come from, and in an extra slot mask; each is computed with its own arithmetic by `wspc_smooth_loss_ex` (flags in
include/wspc.h) on graphs from `wspc_knn_fused` (SmoothConstraint distance flavour: |x|^2 + |y|^2 - 2xy clamped at 0, ties to
the lower index like tf.nn.top_k / tf.argsort).  tests/test_util_variants_gpu.py holds every function to what the reference's
own module returns on the same inputs (tests/golden/make_util_golden.py).

All functions take CUDA fp32 tensors and return a 0-dim CUDA tensor."""
from __future__ import annotations

from . import ops


def _graph(X, knn):
    return ops.knn_fused(ops._as_bnc(X).contiguous(), knn, ops.DIST_SMOOTH, return_dist=True)


def Loss_SpatialColorSmooth_add_SelfContain(Z, X, gamma=1e-1, knn=10):
    '''
    smoothness of the embedding Z on the kNN graph of the points X (all channels of X span the graph)
    :param Z: embedding / class probabilities, float B*N*D
    :param X: points, float B*N*6 (XYZRGB)
    :return: mean over edges of exp(-d/gamma) * mean_c (Z_i - Z_j)^2          (reference :130-167)
    '''
    return ops.smooth_loss(Z, X, gamma, knn)


def Loss_SpatialColorSmoothAdd_UnknownBatch_SelfContain(Z, X, gamma=1e-1, knn=10):
    '''
    the `_add_` loss for graphs whose batch size is not static; the reference SUMS the squared differences over channels
    here (:215) where `_add_` averages them (:163), so the value is D times larger
    :return: mean over edges of exp(-d/gamma) * sum_c (Z_i - Z_j)^2            (reference :169-219)
    '''
    idx, dist = _graph(X, knn)
    return ops.smooth_loss_graph(Z, idx, dist, gamma, ops.SMOOTH_SUM_C)


def Loss_SpatialSmooth(X, W, Ind):
    '''
    smoothness of X itself under caller-supplied weights and neighbour lists
    :param X: float B*N*3;  W: float B*N*Knn edge weights;  Ind: int B*N*Knn neighbour indices
    :return: mean over edges of W * sum_c (X_i - X_j)^2                        (reference :9-33)
    '''
    return ops.smooth_loss_graph(X, Ind, W, 1.0, ops.SMOOTH_SUM_C | ops.SMOOTH_WEIGHTS)


def Loss_SpatialSmooth_SelfContain(X, gamma=1e-1, knn=5):
    '''
    the reference (:64) multiplies the weights by a reduce_sum WITHOUT axis, i.e. by the squared differences summed over
    every edge and channel of the batch; reproduced as written:
    :return: (sum_edges exp(-d/gamma)) * (sum_edges sum_c (X_i - X_j)^2) / (B*N*Knn)          (reference :36-67)
    '''
    idx, dist = _graph(X, knn)
    return ops.smooth_loss_graph(X, idx, dist, gamma, ops.SMOOTH_SUM_C | ops.SMOOTH_GLOBAL_SS)


def Loss_SpatialColorSmooth_SelfContain(Z, X, gamma=1e-1, knn=10):
    '''
    separate graphs on X[..., 0:3] and X[..., 3:6]; an edge slot contributes only where both graphs hold the same
    neighbour in that slot (knn_mask, :113), with the two graphs' weights added
    :return: mean over slots of mask * (W_xyz + W_rgb) * sum_c (Z_i - Z_j)^2   (reference :70-128)
    '''
    X = ops._as_bnc(X)
    ix, dx = _graph(X[:, :, 0:3], knn)
    ir, dr = _graph(X[:, :, 3:6], knn)
    f = ops.SMOOTH_SUM_C
    return ops.smooth_loss_graph(Z, ix, dx, gamma, f, idx_match=ir) + ops.smooth_loss_graph(Z, ir, dr, gamma, f, idx_match=ix)


def ComputeW(Target, X, knn):
    '''reference :224-230 — `Target` was (sess, {'W': graph}) there and is ignored; returns the kNN weights
    W = exp(-d/0.1) (B,N,knn) of the input points'''
    import torch
    _, dist = ops.knn_fused(torch.as_tensor(X, dtype=torch.float32).cuda().contiguous(), knn, ops.DIST_SMOOTH,
                            return_dist=True)
    return torch.exp(-dist / 0.1)
