"""Seeded synthetic inputs with the tensor contract of the reference loaders.

S3DIS-like blocks mimic S3DIS_IO.NextBatch_TrainSet_v1 (S3DIS/DataIO_S3DIS.py:127-154, :320-337, :431-433):
9 channels = xyz (xy block-centred metres) | rgb/255 | xyz / room extent, 5 % exact duplicate points (short
block padding), labels in 0..12, exactly `n_labelled` labelled points per cloud (SampIndex_m-*.mat), and
the Full-style feed layout of TrainOneEpoch_Full (S3DIS_DGCNN_trainer.py:246-314): rows interleaved
[sample, augmented sample, ...] with labels / mask duplicated.
ShapeNet-like clouds mimic ShapeNetIO (ShapeNet/DataIO_ShapeNet.py:145-193, :338-344) and the jitter +
mirror augmentation of ShapeNet_DGCNN_trainer.py:261-283.
"""
from __future__ import annotations

import numpy as np


def s3dis_batch(n_samples: int, N: int = 4096, n_labelled: int = 40, num_classes: int = 13, seed: int = 1234,
                dup_frac: float = 0.05):
    """-> X (2*n_samples, N, 9) f32, Y one-hot (2n, N, 13) f32, Mask (2n, N) f32, seg (2n, N) int64."""
    rng = np.random.default_rng(seed)
    X = np.empty((n_samples, N, 9), np.float32)
    X[:, :, 0:2] = rng.uniform(-0.5, 0.5, (n_samples, N, 2))
    X[:, :, 2] = rng.uniform(0.0, 3.0, (n_samples, N))
    X[:, :, 3:6] = rng.uniform(0.0, 1.0, (n_samples, N, 3))
    off = rng.uniform(0.5, 5.0, (n_samples, 1, 3)).astype(np.float32)
    ext = rng.uniform(6.0, 12.0, (n_samples, 1, 3)).astype(np.float32)
    X[:, :, 6:9] = (X[:, :, 0:3] + off) / ext
    ndup = int(N * dup_frac)
    for s in range(n_samples):          # duplicated points (DataIO_S3DIS.py:431-433)
        src = rng.integers(0, N, ndup)
        dst = rng.integers(0, N, ndup)
        X[s, dst] = X[s, src]
    seg = rng.integers(0, num_classes, (n_samples, N))
    mask = np.zeros((n_samples, N), np.float32)
    for s in range(n_samples):
        mask[s, rng.choice(N, n_labelled, replace=False)] = 1.0
    # Siamese partner: one of the axis-swap / mirror augmentations applied to a COPY
    Xa = X.copy()
    for s in range(n_samples):
        mode = rng.integers(0, 4)
        if mode & 1:
            Xa[s, :, [0, 1]] = Xa[s, :, [1, 0]]
            Xa[s, :, [6, 7]] = Xa[s, :, [7, 6]]
        if mode & 2:
            Xa[s, :, 0] = -Xa[s, :, 0]
            Xa[s, :, 6] = Xa[s, :, 6].max() - Xa[s, :, 6]
    Xf = np.empty((2 * n_samples, N, 9), np.float32)
    Xf[0::2], Xf[1::2] = X, Xa
    segf = np.repeat(seg, 2, axis=0)
    maskf = np.repeat(mask, 2, axis=0)
    Y = np.zeros((2 * n_samples, N, num_classes), np.float32)
    np.put_along_axis(Y, segf[..., None], 1.0, axis=-1)
    return Xf, Y, maskf, segf


AIRPLANE_PARTS = (0, 1, 2, 3)
# part-id ranges per category (16 categories, 50 parts), DataIO_ShapeNet.py:36-41
CAT_PART_RANGES = [(0, 4), (4, 6), (6, 8), (8, 12), (12, 16), (16, 19), (19, 22), (22, 24), (24, 28), (28, 30), (30, 36),
                   (36, 38), (38, 41), (41, 44), (44, 47), (47, 50)]


def shapenet_batch(n_samples: int, N: int = 2048, n_labelled: int = 204, seed: int = 1234, category: int | None = None):
    """-> X (2n, N, 3), label one-hot (2n, 16), Y one-hot (2n, N, 50), Mask (2n, N), seg (2n, N)."""
    rng = np.random.default_rng(seed)
    X = rng.uniform(-1, 1, (n_samples, N, 3)).astype(np.float32)
    X -= X.mean(axis=1, keepdims=True)
    X /= np.sqrt((X ** 2).sum(-1)).max(axis=1)[:, None, None]       # pc_normalize, DataIO_ShapeNet.py:338-344
    cat = np.full((n_samples,), category) if category is not None else rng.integers(0, 16, (n_samples,))
    seg = np.stack([rng.integers(*CAT_PART_RANGES[c], (N,)) for c in cat])
    mask = np.zeros((n_samples, N), np.float32)
    for s in range(n_samples):
        mask[s, rng.choice(N, n_labelled, replace=False)] = 1.0
    Xa = X.copy()
    for s in range(n_samples):                                      # ShapeNet_DGCNN_trainer.py:266-275
        extent = X[s].max(0) - X[s].min(0)
        Xa[s] = X[s] + (2e-3 * extent * rng.standard_normal((N, 3))).astype(np.float32)
        if rng.random() > 0.5:
            Xa[s, :, 2] = -Xa[s, :, 2]
    Xf = np.empty((2 * n_samples, N, 3), np.float32)
    Xf[0::2], Xf[1::2] = X, Xa
    catf = np.repeat(cat, 2)
    lab = np.zeros((2 * n_samples, 16), np.float32)
    lab[np.arange(2 * n_samples), catf] = 1.0
    segf = np.repeat(seg, 2, axis=0)
    Y = np.zeros((2 * n_samples, N, 50), np.float32)
    np.put_along_axis(Y, segf[..., None], 1.0, axis=-1)
    return Xf, lab, Y, np.repeat(mask, 2, axis=0), segf
