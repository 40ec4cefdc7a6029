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


# ------------------------------------------------------------------------------------------------------------
# Fixtures produced by the REFERENCE'S OWN PYTHON on the TF-1.14 shim (tests/golden/make_reference_golden.py).
# kNN stage-wise (the reference's matmul order is unspecified: lists may differ on last-bit ties), everything
# downstream teacher-forced with the reference's neighbour lists: logits / losses within 1e-3, gradients in L2.
# ------------------------------------------------------------------------------------------------------------
import sys
sys.path.insert(0, G)
import refgen_common as rc  # noqa: E402


def _i32(a, dev):
    return torch.from_numpy(a.astype(np.int32)).to(dev)


def _keep(packed, dev):
    return torch.from_numpy(np.unpackbits(packed, axis=-1).astype(np.float32)).to(dev)


def _tie_level(idx_a, idx_b, dist):
    da = torch.gather(dist, -1, idx_a.long())
    db = torch.gather(dist, -1, idx_b.long())
    assert float((da - db).abs().max()) <= 1e-5 * float(dist.abs().max())
    return float((idx_a == idx_b).all(-1).float().mean())


def _smooth_graph(Xs, idx, cuda):
    from weaksuppointcloudseg_b200 import ops
    d = ops.pairwise_distance(Xs.contiguous(), ops.DIST_SMOOTH)
    return idx, torch.gather(d, -1, idx.long())


def _grad_check(got, f, lim):
    gmax = max(np.abs(f[k]).max() for k in f.files if k.startswith("grad/"))
    bad = {}
    for k in f.files:
        if not k.startswith("grad/"):
            continue
        a, b = rc.subsample(got[k[len("grad/"):]])[0].astype(np.float64), f[k].astype(np.float64)
        if np.abs(b).max() < 1e-6 * gmax:
            assert np.abs(a).max() < 1e-4 * gmax, k
            continue
        e = np.linalg.norm(a - b) / np.linalg.norm(b)
        # the T-net's gradients all hang off one (B,3,3) tensor and its FC layers normalise over B=6 clouds only: the routing
        # noise of an UN-forced comparison there is 2-3x the rest of the net and varies from run to run with the summation
        # order of the statistics (tools/flake_probe.py, 40 runs: trunk 2.4e-2 .. 4.0e-2, T-net 7.8e-2 .. 1.3e-1).  The tight
        # statement is the forced-routing test (tests/test_baseline_shapes_gpu.py, tests/test_shapenet_engine_gpu.py).
        if e > (3 * lim if "transform_net1/" in k else lim):
            bad[k] = e
    assert not bad, bad


def test_unit_ops_against_reference_code(cuda):
    from weaksuppointcloudseg_b200 import ops
    f = np.load(os.path.join(G, "ref_unit_ops.npz"))
    X = torch.from_numpy(f["X"]).to(cuda)
    xyz = X[:, :, 6:9].contiguous()
    adj = ops.pairwise_distance(xyz, ops.DIST_TFUTIL)
    assert float((adj.cpu() - torch.from_numpy(f["adj"])).abs().max()) <= 1e-5 * np.abs(f["adj"]).max()
    idx = ops.knn_fused(X, 20, ops.DIST_TFUTIL, coff=6, D=3)
    assert _tie_level(idx, _i32(f["knn"], cuda), adj) >= 0.97
    ef = ops.get_edge_feature(X, _i32(f["knn"], cuda))
    assert np.array_equal(ef.cpu().numpy()[:, ::8], f["edge_feature"])
    P = torch.from_numpy(f["P"]).to(cuda)
    sm = ops.smooth_loss(P, X[:, :, 0:6].contiguous())
    assert abs(float(sm) - float(f["smooth_loss"])) <= 1e-3 * abs(float(f["smooth_loss"]))
    bg = ops.batch_gather(P, _i32(f["gather_idx"], cuda))
    assert np.array_equal(bg.cpu().numpy(), f["gathered"])


def test_s3dis_step_against_reference_code(cuda):
    from weaksuppointcloudseg_b200.engine_s3dis import S3DISEngine
    f = np.load(os.path.join(G, "ref_s3dis_step.npz"))
    params0 = rc.xavier_params(rc.S3DIS_LAYERS, int(f["param_seed"][0]))
    B, N = f["X"].shape[:2]
    bs = int(f["batch_size"][0])
    X, M = (torch.from_numpy(f[k]).to(cuda) for k in ("X", "Mask"))
    Y = torch.from_numpy(f["Y"].astype(np.float32)).to(cuda)
    eng = S3DISEngine(params0, B, N, device=cuda)
    # inference graph on the initial variables (population statistics, no dropout)
    Ze = eng.forward(X, False, knn_override={k: _i32(f[k + "_eval"], cuda) for k in ("knn1", "knn2", "knn3")}).cpu().numpy()
    assert np.abs(Ze - f["Z_eval"]).max() <= 1e-3 * np.abs(f["Z_eval"]).max()
    ov = {k: _i32(f[k], cuda) for k in ("knn2", "knn3")}          # knn1 is computed by the kernel
    sg = _smooth_graph(X[:, :, 0:6], _i32(f["knn_smooth"], cuda), cuda)
    bn_decay = min(0.99, 1 - 0.5 * 0.5 ** ((0 * bs) // 600000))
    losses = eng.train_step(X, Y, M, lr=1e-3, bn_decay=bn_decay,
                            dropout_mask=_keep(f["dropout_keep"], cuda), knn_override=ov, smooth_graph=sg, apply=False)
    torch.cuda.synchronize()
    assert np.array_equal(eng.idx[0].cpu().numpy(), f["knn1"].astype(np.int32))  # no last-bit ties in this fixture
    Z = eng.Z.cpu().numpy()
    assert np.abs(Z - f["Z"]).max() <= 1e-3 * np.abs(f["Z"]).max()
    ref = [float(f[n]) for n in ("loss_seg", "loss_siamese", "loss_inexact", "loss_smooth", "loss")]
    assert np.allclose(losses.cpu().numpy(), ref, rtol=1e-3)
    # 3e-2: ReLU masks / arg-max pools route single-element gradients differently between two correct fp32
    # implementations (measured 1.4e-2 here, one max_pool2d flip; see tests/test_s3dis_engine_gpu.py); the kernels
    # themselves are pinned to 1e-5 on identical inputs in tests/test_kernels_gpu.py
    _grad_check(eng.vs.grads(), f, 3e-2)
    st = eng.vs.export()
    for n in params0:
        if n.endswith("pop_mean") or n.endswith("pop_var"):                      # BN population statistics after the step
            assert np.abs(st[n] - f["after/" + n]).max() <= 1e-4 * max(np.abs(f["after/" + n]).max(), 1.0), n
    # Adam on the engine's own gradients reproduces the reference's updated weights where |g| >> eps
    eng.vs.adam_step(1e-3)
    after = eng.vs.export()
    gmax = max(np.abs(f[k_]).max() for k_ in f.files if k_.startswith("grad/"))
    for n in eng.vs.trainable_names:
        if "after/" + n not in f.files or "grad/" + n not in f.files:
            continue
        g = f["grad/" + n].reshape(params0[n].shape)
        if np.abs(g).max() < 1e-4 * gmax:      # analytically zero gradient (bias in front of a batch norm): pure noise
            continue
        # first Adam step = lr * g / (|g| + eps'): sign-like, so only elements whose sign cannot be flipped by the
        # ~1e-2 gradient noise discussed above are compared
        well = np.abs(g) > 0.2 * np.abs(g).max()
        if well.any():
            assert np.abs(after[n] - f["after/" + n])[well].max() <= 2e-4, n


def test_shapenet_step_against_reference_code(cuda):
    from weaksuppointcloudseg_b200.engine_shapenet import ShapeNetEngine
    f = np.load(os.path.join(G, "ref_shapenet_step.npz"))
    params0 = rc.xavier_params(rc.SHAPENET_LAYERS, int(f["param_seed"][0]), tnet_seed=int(f["param_seed"][1]))
    B, N = f["X"].shape[:2]
    bs = int(f["batch_size"][0])
    X, lab, M = (torch.from_numpy(f[k]).to(cuda) for k in ("X", "label", "Mask"))
    Y = torch.from_numpy(f["Y"].astype(np.float32)).to(cuda)
    eng = ShapeNetEngine(params0, B, N, device=cuda)
    ov = {k: _i32(f[k], cuda) for k in ("knn1", "knn2", "knn3")}  # knn0 (raw xyz) is computed by the kernel
    sg = _smooth_graph(X, _i32(f["knn_smooth"], cuda), cuda)
    masks = [_keep(f[k], cuda) for k in ("dropout_keep1", "dropout_keep2")]
    bn_decay = min(0.99, 1 - 0.5 * 0.5 ** ((0 * bs) // (2 * 16881 * 20)))
    losses = eng.train_step(X, lab, Y, M, lr=1e-3, bn_decay=bn_decay, dropout_masks=masks, knn_override=ov, smooth_graph=sg,
                            apply=False)
    torch.cuda.synchronize()
    assert np.array_equal(eng.idx[0].cpu().numpy(), f["knn0"].astype(np.int32))
    Z = eng.Z.cpu().numpy()
    assert np.abs(Z - f["Z"]).max() <= 1e-3 * np.abs(f["Z"]).max()
    ref = [float(f[n]) for n in ("loss_seg", "loss_siamese", "loss_inexact", "loss_smooth", "loss")]
    assert np.allclose(losses.cpu().numpy(), ref, rtol=1e-3)
    _grad_check(eng.vs.grads(), f, 6e-2)      # same bound and rationale as tests/test_shapenet_engine_gpu.py


def test_label_propagation_against_reference_code(cuda):
    from weaksuppointcloudseg_b200 import ops
    f = np.load(os.path.join(G, "ref_label_prop.npz"))
    Lm = ops.laplacian_sym(torch.from_numpy(f["xyz"]).to(cuda), torch.from_numpy(f["rgb"]).to(cuda))
    assert np.abs(Lm.cpu().numpy() - f["L"]).max() <= 5e-4 * np.abs(f["L"]).max()
    _, Yp, w = ops.lp_solve(torch.from_numpy(f["L"][0]).to(cuda), torch.from_numpy(f["G"]).to(cuda))
    assert np.abs(w.cpu().numpy() - f["w"]).max() <= 1e-4
    assert np.abs(Yp.cpu().numpy() - f["Y_prob"]).max() <= 1e-3 * np.abs(f["Y_prob"]).max()
