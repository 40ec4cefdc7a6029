"""Generates the golden fixtures in this directory from the CPU oracle (run from the repo root:
`python tests/golden/make_golden.py`).  The reference ships no golden vectors and TensorFlow 1.14 cannot be
installed, so these pin the ORACLE (SURVEY §8c "parity unpinned"): the CPU suite checks that the oracle still
reproduces them bit-for-bit (kNN) / to 1e-6 (floating point), the GPU suite checks the CUDA path against them."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import dgcnn as od  # noqa: E402
from oracle import knn as oknn  # noqa: E402
from oracle import lp as olp  # noqa: E402
from weaksuppointcloudseg_b200 import synthetic as syn  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def knn_fixture():
    X, _, _, _ = syn.s3dis_batch(1, N=512, n_labelled=8, seed=101)          # includes 5 % duplicated points
    feats = np.maximum(np.random.default_rng(102).standard_normal((2, 512, 64)), 0).astype(np.float32)
    out = dict(X=X, feats=feats)
    out["idx_xyz"] = oknn.knn(X, 20, oknn.TFUTIL, coff=6, D=3)
    out["idx_feat"] = oknn.knn(feats, 20, oknn.TFUTIL)
    i, d = oknn.knn(X, 10, oknn.SMOOTH, coff=0, D=6, return_dist=True)
    out["idx_smooth"], out["dist_smooth"] = i, d
    np.savez_compressed(os.path.join(HERE, "knn_golden.npz"), **out)


def s3dis_fixture():
    n_samples, N = 2, 192
    X, Y, M, _ = syn.s3dis_batch(n_samples, N=N, n_labelled=8, seed=103)
    params = od.init_params(od.S3DIS_LAYERS, seed=104)
    mask = np.floor(0.7 + np.random.default_rng(105).random((2 * n_samples, N, 256))).astype(np.float32)
    p = od.to_torch(params)
    opt = od.AdamTF(p, od.trainable_names(p))
    rec = {}
    out = od.train_step_s3dis(p, opt, torch.from_numpy(X), torch.from_numpy(Y), torch.from_numpy(M), step=0,
                              dropout_mask=torch.from_numpy(mask), rec=rec)
    np.savez_compressed(
        os.path.join(HERE, "s3dis_step_golden.npz"), X=X, Y=Y, Mask=M, dropout_mask=mask.astype(np.uint8),
        knn1=rec["knn1/idx"].numpy().astype(np.int32), knn2=rec["knn2/idx"].numpy().astype(np.int32),
        knn3=rec["knn3/idx"].numpy().astype(np.int32), logits=out["Z"].detach().numpy(),
        losses=np.array([float(out[k].detach()) for k in ("loss_seg", "loss_siamese", "loss_inexact", "loss_smooth", "loss")],
                        np.float64),
        seed_params=np.array([104]))


def lp_fixture():
    rng = np.random.default_rng(106)
    N, K = 128, 13
    xyz = (np.concatenate([rng.uniform(-0.5, 0.5, (1, N, 2)), rng.uniform(0, 3, (1, N, 1))], -1) * 0.15).astype(np.float32)
    rgb = rng.uniform(0, 1, (1, N, 3)).astype(np.float32)
    G = rng.dirichlet(np.ones(K) * 0.3, N).astype(np.float32)
    Lm = olp.laplacian_sym(xyz, rgb)
    Y, Yp, w = olp.solve(Lm[0], G)
    np.savez_compressed(os.path.join(HERE, "lp_golden.npz"), xyz=xyz, rgb=rgb, G=G, L=Lm, Y_prob=Yp, w=w)


if __name__ == "__main__":
    knn_fixture()
    s3dis_fixture()
    lp_fixture()
    print("golden fixtures written to", HERE)
