"""Drop-in for the parts of the reference's Util/Tool.py that the hot path and the trainers use.

Host helpers (vectorised; same results as the reference's per-point Python loops):
  OnehotEncode (Util/Tool.py:4-28), IoU (:142-169), IoU_detail (:171-195), printout (:215-229)
Device helpers (CUDA through the C ABI):
  batch_gather_v1 (:72-104), TF_Computation.LaplacianMatSym_XYZRGB_DirectComp (:435-468)
"""
from __future__ import annotations

import numpy as np


def OnehotEncode(Y, K, dtype=np.float64):
    """Onehot key encoding of input label Y (B*N, N, or scalar) -> float64 one-hot like the reference.  The epoch loops ask
    for float32, the dtype of the feed: building 128 x 4096 x 13 in float64 and converting costs 75 ms of host time per
    step, building it in float32 directly 4 ms."""
    Y = np.asarray(Y)
    out = np.zeros(Y.shape + (K,), dtype)
    if Y.ndim == 0:
        out[int(Y)] = 1
        return out
    np.put_along_axis(out, Y.astype(np.int64)[..., None], 1, axis=-1)
    return out


def IoU_detail(pred, gt, K):
    '''
    function to measure IoU for batch input
    :param pred: B*N
    :param gt: B*N
    :return: iou, intersect, union  (B*K each)
    '''
    pred, gt = np.asarray(pred), np.asarray(gt)
    ks = np.arange(K)
    p1 = pred[..., None] == ks
    g1 = gt[..., None] == ks
    intersect = np.sum(p1 & g1, axis=1)
    union = np.sum(p1, axis=1) + np.sum(g1, axis=1) - intersect
    iou = intersect / (union + 1e-6)
    return iou, intersect, union


def IoU(pred, gt, K):
    return IoU_detail(pred, gt, K)[0]


def printout(str, write_flag=False, fid=None, end=''):
    '''
    function to print the string (str) and write into a file if fid is provided
    '''
    print(str, end=end)
    if write_flag:
        fid.write(str + end)


def pdist_np(X):
    "Euclidean distance matrix of the rows of X (N*D) -> N*N, negatives from cancellation clamped (Util/Tool.py:57-70)"
    X = np.asarray(X)
    sq = np.sum(X ** 2, axis=-1, keepdims=True)
    return np.sqrt(np.maximum(sq + sq.T - 2 * (X @ X.T), 0.))


def L2NormVec(x):
    "The reference's `L2 normalize` (Util/Tool.py:197-204) -- note it returns sqrt(x / sum(x^2)), kept as written"
    return np.sqrt(x / np.sum(x ** 2))


def L1NormVec(x):
    "x / sum(|x|) (Util/Tool.py:206-213)"
    return x / np.sum(np.abs(x))


def ResamplePointCloud(X, target_num_pts):
    """Resample N*D points to target_num_pts rows: without replacement when shrinking, with replacement when growing
    (Util/Tool.py:270-289; same np.random.choice calls).  -> (points, sample indices)"""
    N = X.shape[0]
    if N == target_num_pts:
        return X, np.arange(0, N)
    samp_idx = np.random.choice(np.arange(0, N), target_num_pts, N < target_num_pts)
    return X[samp_idx, :], samp_idx


def pdist2(X, Y):
    "Squared distances between the rows of X (B*N*D) and Y (B*M*D) -> B*N*M, CUDA tensors (Util/Tool.py:30-42)"
    import torch
    return (X ** 2).sum(-1, keepdim=True) + (Y ** 2).sum(-1).unsqueeze(1) - 2 * torch.einsum('ijk,ilk->ijl', X, Y)


def pdist(X):
    "Euclidean distance matrix per cloud, X B*N*D -> B*N*N (Util/Tool.py:44-55); squared distances from the wspc kernel"
    from . import ops
    return ops.pairwise_distance(X, ops.DIST_SMOOTH).sqrt()


def batch_gather_v1(X, idx):
    '''
    batch gather function (Util/Tool.py:72-104)
    :param X: Input tensor to be sliced/gathered B*N*D   (CUDA fp32)
    :param idx: Slicing/Gather index B*N*Knn              (CUDA int32)
    :return: Xgather: Sliced/Gathered X B*N*Knn*D
    '''
    from . import ops
    return ops.batch_gather(X, idx)


class TF_Computation:
    """Namespace kept for signature compatibility (Util/Tool.py:291-468)."""

    class LaplacianMatSym_XYZRGB_DirectComp():
        """Lsym = D^-1/2 (D + 1e-8 - W) D^-1/2 with W = exp(-1e3 d_xyz) * exp(-10 d_rgb) (Util/Tool.py:435-468)."""

        def __init__(self):
            pass

        def Eval(self, sess, X, RGB):
            from . import ops
            return ops.laplacian_sym(X, RGB)
