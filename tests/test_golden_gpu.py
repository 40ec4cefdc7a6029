"""GPU: the CUDA path against the committed golden fixtures (no oracle execution at run time)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_knn_against_golden(cuda):
    from weaksuppointcloudseg_b200 import ops
    g = np.load(os.path.join(G, "knn_golden.npz"))
    X = torch.from_numpy(g["X"]).to(cuda)
    assert np.array_equal(ops.knn_fused(X, 20, ops.DIST_TFUTIL, coff=6, D=3).cpu().numpy(), g["idx_xyz"])
    assert np.array_equal(ops.knn_fused(torch.from_numpy(g["feats"]).to(cuda), 20).cpu().numpy(), g["idx_feat"])
    i, d = ops.knn_fused(X, 10, ops.DIST_SMOOTH, coff=0, D=6, return_dist=True)
    assert np.array_equal(i.cpu().numpy(), g["idx_smooth"]) and np.array_equal(d.cpu().numpy(), g["dist_smooth"])


def test_s3dis_step_against_golden(cuda):
    from oracle import dgcnn as od   # only for the seeded initialiser (numpy), nothing is executed on the CPU path
    from weaksuppointcloudseg_b200.engine_s3dis import S3DISEngine
    g = np.load(os.path.join(G, "s3dis_step_golden.npz"))
    B, N = g["X"].shape[:2]
    eng = S3DISEngine(od.init_params(od.S3DIS_LAYERS, seed=int(g["seed_params"][0])), B, N, device=cuda)
    ov = {f"knn{i}": torch.from_numpy(g[f"knn{i}"]).to(cuda) for i in (2, 3)}
    losses = eng.train_step(torch.from_numpy(g["X"]).to(cuda), torch.from_numpy(g["Y"]).to(cuda),
                            torch.from_numpy(g["Mask"]).to(cuda), lr=1e-3, bn_decay=0.5,
                            dropout_mask=torch.from_numpy(g["dropout_mask"].astype(np.float32)).to(cuda), knn_override=ov)
    assert np.array_equal(eng.idx[0].cpu().numpy(), g["knn1"])
    Z = eng.Z.cpu().numpy()
    assert np.abs(Z - g["logits"]).max() <= 1e-3 * np.abs(g["logits"]).max()
    assert np.allclose(losses.cpu().numpy(), g["losses"], rtol=1e-3)


def test_lp_against_golden(cuda):
    from weaksuppointcloudseg_b200 import ops
    g = np.load(os.path.join(G, "lp_golden.npz"))
    Lm = ops.laplacian_sym(torch.from_numpy(g["xyz"]).to(cuda), torch.from_numpy(g["rgb"]).to(cuda))
    assert np.abs(Lm.cpu().numpy() - g["L"]).max() <= 1e-4 * np.abs(g["L"]).max()
    _, Yp, w = ops.lp_solve(Lm[0], torch.from_numpy(g["G"]).to(cuda))
    assert np.abs(Yp.cpu().numpy() - g["Y_prob"]).max() <= 1e-3 * np.abs(g["Y_prob"]).max()
    assert np.abs(w.cpu().numpy() - g["w"]).max() <= 1e-4
