"""CPU: the oracle reproduces the committed golden fixtures (tests/golden/make_golden.py), and the invariants that
follow from the reference's definitions hold for the oracle."""
import os

import numpy as np
import torch

from oracle import dgcnn as od
from oracle import knn as oknn
from oracle import lp as olp

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_knn_oracle_reproduces_golden():
    g = np.load(os.path.join(G, "knn_golden.npz"))
    assert np.array_equal(oknn.knn(g["X"], 20, oknn.TFUTIL, coff=6, D=3), g["idx_xyz"])
    assert np.array_equal(oknn.knn(g["feats"], 20, oknn.TFUTIL), g["idx_feat"])
    i, d = oknn.knn(g["X"], 10, oknn.SMOOTH, coff=0, D=6, return_dist=True)
    assert np.array_equal(i, g["idx_smooth"]) and np.array_equal(d, g["dist_smooth"])
    # invariant 2: self at rank 0 with d == 0 unless an exact duplicate with a lower index exists
    rank0 = g["idx_smooth"][..., 0]
    n = np.arange(rank0.shape[1])
    assert np.all(g["dist_smooth"][..., 0] == 0) and np.all(rank0 <= n)


def test_s3dis_oracle_reproduces_golden():
    g = np.load(os.path.join(G, "s3dis_step_golden.npz"))
    p = od.to_torch(od.init_params(od.S3DIS_LAYERS, seed=int(g["seed_params"][0])))
    opt = od.AdamTF(p, od.trainable_names(p))
    out = od.train_step_s3dis(p, opt, torch.from_numpy(g["X"]), torch.from_numpy(g["Y"]), torch.from_numpy(g["Mask"]), step=0,
                              dropout_mask=torch.from_numpy(g["dropout_mask"].astype(np.float32)))
    assert np.abs(out["Z"].detach().numpy() - g["logits"]).max() <= 1e-5 * np.abs(g["logits"]).max()
    got = np.array([float(out[k].detach()) for k in ("loss_seg", "loss_siamese", "loss_inexact", "loss_smooth", "loss")])
    assert np.allclose(got, g["losses"], rtol=1e-5)


def test_lp_oracle_reproduces_golden():
    g = np.load(os.path.join(G, "lp_golden.npz"))
    Lm = olp.laplacian_sym(g["xyz"], g["rgb"])
    assert np.allclose(Lm, g["L"], rtol=1e-6, atol=1e-7)
    _, Yp, w = olp.solve(Lm[0], g["G"])
    assert np.allclose(Yp, g["Y_prob"], rtol=1e-6) and np.allclose(w, g["w"], rtol=1e-6)
    # invariant 8: rows of Lsym have diagonal (d_i + 1e-8 - 1)/d_i, L symmetric PSD
    assert np.abs(Lm[0] - Lm[0].T).max() < 1e-6
    assert np.linalg.eigvalsh(Lm[0].astype(np.float64)).min() > -1e-5


def test_oracle_invariants():
    from weaksuppointcloudseg_b200 import synthetic as syn
    X, Y, M, _ = syn.s3dis_batch(2, N=128, n_labelled=6, seed=1)
    p = od.to_torch(od.init_params(od.S3DIS_LAYERS, seed=2), requires_grad=False)
    Xi = torch.from_numpy(X).clone()
    Xi[1::2] = Xi[0::2]
    Z = od.get_model_s3dis(p, Xi, False)
    Zp = torch.softmax(Z, -1)
    assert float(od.siamese_loss(Zp, 10.0)) == 0.0                                   # invariant 3
    assert float(od.smooth_loss(torch.full_like(Zp, 1 / 13), Xi[:, :, :6])) == 0.0   # invariant 4
    assert float(od.smooth_loss(Zp, Xi[:, :, :6])) >= 0.0
    ones = torch.ones(Z.shape[:2])                                                   # invariant 5
    seg = torch.from_numpy(Y).argmax(-1)
    ce = torch.nn.functional.cross_entropy(Z.reshape(-1, 13), seg.reshape(-1))
    assert abs(float(od.seg_loss(Z, torch.from_numpy(Y), ones)) - float(ce)) < 1e-5
    # invariant 6: EdgeConv stack is permutation equivariant
    perm = torch.randperm(128, generator=torch.Generator().manual_seed(0))
    Zperm = od.get_model_s3dis(p, Xi[:, perm], False)
    assert float((Zperm - Z[:, perm]).abs().max()) <= 2e-4 * float(Z.abs().max())
    # invariant 1: T-net is the identity at initialisation
    Xs, lab, _, _, _ = syn.shapenet_batch(1, N=96, n_labelled=8, seed=3)
    ps = od.to_torch(od.init_params(od.SHAPENET_LAYERS, seed=4, shapenet=True), requires_grad=False)
    rec = {}
    od.get_model_shapenet(ps, torch.from_numpy(Xs), torch.from_numpy(lab), True, bn_decay=0.5, rec=rec)
    assert torch.equal(rec["transform"], torch.eye(3).expand(2, 3, 3)) and torch.equal(rec["pct"], torch.from_numpy(Xs))
    assert torch.equal(rec["knn0/idx"], rec["knn1/idx"])


def test_oracle_loss_gradients_by_finite_differences():
    """invariant 9: analytic gradients of the weak losses (fp64 oracle) vs central finite differences."""
    rng = np.random.default_rng(0)
    B, N, C = 2, 24, 5
    Z = torch.tensor(rng.normal(size=(B, N, C)), dtype=torch.float64, requires_grad=True)
    X = torch.tensor(rng.uniform(size=(B, N, 6)), dtype=torch.float64)
    Y = torch.nn.functional.one_hot(torch.tensor(rng.integers(0, C, (B, N))), C).double()
    M = torch.tensor((rng.random((B, N)) < 0.3).astype(np.float64))
    graph = od.smooth_graph(X.float(), knn=4)

    def total(z):
        return od.weak_sup_losses(z, X, Y, M, 10.0, graph)["loss"]
    g, = torch.autograd.grad(total(Z), Z)
    eps = 1e-6
    for _ in range(12):
        b, n, c = rng.integers(0, B), rng.integers(0, N), rng.integers(0, C)
        d = torch.zeros_like(Z)
        d[b, n, c] = eps
        fd = (total(Z.detach() + d) - total(Z.detach() - d)) / (2 * eps)
        assert abs(float(fd) - float(g[b, n, c])) <= 1e-6 + 1e-5 * abs(float(fd))


def test_cfg1_shapenet_airplane_one_cloud_forward_ce():
    """BASELINE cfg-1, the reference's CPU-runnable plumbing case: one Airplane cloud, N=2048, k=20, Plain style
    (forward + masked cross-entropy + backward) through the oracle at full size; properties only (no TF to compare with):
    identity input transform at initialisation, self-inclusive kNN, CE near ln 50 for Xavier weights, finite gradients."""
    import torch
    from oracle import dgcnn as od
    from weaksuppointcloudseg_b200 import synthetic as syn

    X, lab, Y, M, seg = syn.shapenet_batch(1, N=2048, n_labelled=204, seed=7, category=0)
    X, lab, Y, M = (torch.from_numpy(a[0:1]) for a in (X, lab, Y, M))
    assert int(lab.argmax()) == 0 and set(np.unique(seg[0])) <= {0, 1, 2, 3} and int(M.sum()) == 204
    p = od.to_torch(od.init_params(od.SHAPENET_LAYERS, seed=1234, shapenet=True))
    rec = {}
    Z = od.get_model_shapenet(p, X, lab, True, bn_decay=0.5, dropout_masks=(torch.ones(1, 2048, 256), torch.ones(1, 2048, 256)),
                              rec=rec)
    assert Z.shape == (1, 2048, 50)
    assert torch.allclose(rec["transform"][0], torch.eye(3), atol=1e-6)            # T-net: W = 0, b = eye (transform_nets.py:41-50)
    idx0 = rec["knn0/idx"]
    assert idx0.shape == (1, 2048, 20) and bool((idx0[0, :, 0] == torch.arange(2048)).all())
    loss = od.seg_loss(Z, Y, M)
    assert abs(float(loss.detach()) - np.log(50.0)) < 1.5
    loss.backward()
    g = p["adj_conv1/weights"].grad
    assert g is not None and bool(torch.isfinite(g).all()) and float(g.abs().max()) > 0
