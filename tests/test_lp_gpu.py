"""Label propagation path (Laplacian + closed-form solve) and the unfused API ops against the CPU oracle.
Reference: Util/Tool.py:435-468, Util/ProbLabelPropagation.py:8-62, Util/SmoothConstraint.py:130-167."""
import numpy as np
import pytest
import torch

from oracle import dgcnn as od
from oracle import lp as olp

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def _block(rng, N):
    xyz = np.concatenate([rng.uniform(-0.5, 0.5, (1, N, 2)), rng.uniform(0, 3, (1, N, 1))], -1).astype(np.float32)
    xyz = (xyz * 0.15).astype(np.float32)          # dense enough that exp(-1e3 d) links neighbours
    rgb = rng.uniform(0, 1, (1, N, 3)).astype(np.float32)
    return xyz, rgb


def test_laplacian_matches_oracle(cuda):
    from weaksuppointcloudseg_b200 import ops
    rng = np.random.default_rng(0)
    xyz, rgb = _block(rng, 300)
    Lref = olp.laplacian_sym(xyz, rgb)
    Lm = ops.laplacian_sym(torch.from_numpy(xyz).to(cuda), torch.from_numpy(rgb).to(cuda)).cpu().numpy()
    assert rel(Lm, Lref) <= 1e-4
    # SURVEY §4 invariant 8: symmetric, diagonal (d_i + 1e-8 - 1)/d_i
    assert np.abs(Lm[0] - Lm[0].T).max() <= 1e-6


def test_lp_solve_matches_dense_inverse(cuda):
    from weaksuppointcloudseg_b200 import ops
    rng = np.random.default_rng(1)
    N, K = 256, 13
    xyz, rgb = _block(rng, N)
    Lm = olp.laplacian_sym(xyz, rgb)[0]
    logits = rng.normal(0, 2.0, (N, K))
    G = (np.exp(logits) / np.exp(logits).sum(-1, keepdims=True)).astype(np.float32)
    Yref, Ypref, wref = olp.solve(Lm, G)
    Y, Yp, w = ops.lp_solve(torch.from_numpy(Lm).to(cuda), torch.from_numpy(G).to(cuda))
    assert rel(w.cpu().numpy(), wref) <= 1e-4
    assert rel(Yp.cpu().numpy(), Ypref) <= 1e-3          # SURVEY App. A-11 bar
    assert rel(Y.cpu().numpy(), Yref) <= 1e-3


def test_lp_invariants(cuda):
    """SURVEY §4 invariant 7: alpha = 0 returns Y_prob ~ G; w ~ 1 for one-hot G, ~ 0 for uniform G."""
    from weaksuppointcloudseg_b200 import ops
    from weaksuppointcloudseg_b200.ProbLabelPropagation import LabelPropagation_TF
    rng = np.random.default_rng(2)
    N, K = 128, 13
    G = rng.dirichlet(np.ones(K), N).astype(np.float32)
    Lm = olp.laplacian_sym(*_block(rng, N))[0]
    _, Yp, _ = ops.lp_solve(torch.from_numpy(Lm).to(cuda), torch.from_numpy(G).to(cuda), alpha=0.0, beta=1.0)
    assert rel(Yp.cpu().numpy(), G) <= 1e-3
    onehot = np.eye(K, dtype=np.float32)[rng.integers(0, K, N)]
    uni = np.full((N, K), 1.0 / K, np.float32)
    lp = LabelPropagation_TF(1.0, 1.0, 10)
    assert np.abs(lp.EvalWeight4EachPoint(None, onehot)[0] - 1.0).max() < 1e-3
    assert np.abs(lp.EvalWeight4EachPoint(None, uni)[0]).max() < 1e-3


def test_smooth_loss_api_and_gather(cuda):
    from weaksuppointcloudseg_b200 import SmoothConstraint, Tool, tf_util
    rng = np.random.default_rng(3)
    B, N, C = 2, 200, 13
    X = rng.uniform(0, 1, (B, N, 6)).astype(np.float32)
    Z = rng.dirichlet(np.ones(C), (B, N)).astype(np.float32)
    ref = float(od.smooth_loss(torch.from_numpy(Z), torch.from_numpy(X)))
    got = float(SmoothConstraint.Loss_SpatialColorSmooth_add_SelfContain(torch.from_numpy(Z).to(cuda), torch.from_numpy(X).to(cuda)))
    assert abs(got - ref) <= 1e-3 * abs(ref)
    # invariant 4: >= 0, == 0 for constant Z
    const = torch.full((B, N, C), 1.0 / C, device=cuda)
    assert float(SmoothConstraint.Loss_SpatialColorSmooth_add_SelfContain(const, torch.from_numpy(X).to(cuda))) == 0.0
    idx = torch.from_numpy(rng.integers(0, N, (B, N, 7)).astype(np.int32)).to(cuda)
    Xd = torch.from_numpy(X).to(cuda)
    gat = Tool.batch_gather_v1(Xd, idx)
    gidx = (idx.long() + torch.arange(B, device=cuda).view(B, 1, 1) * N).reshape(-1)
    assert torch.equal(gat, Xd.reshape(B * N, 6)[gidx].view(B, N, 7, 6))
    edge = tf_util.get_edge_feature(Xd.unsqueeze(2), idx, k=7)
    assert torch.equal(edge, od.get_edge_feature(torch.from_numpy(X), idx.cpu().long()).to(cuda))


def test_fused_model_matches_unfused_tf_util_graph(cuda):
    """DGCNN_S3DIS.get_model (fused engine) vs the same graph assembled from the unfused tf_util ops, inference mode."""
    from weaksuppointcloudseg_b200 import DGCNN_S3DIS, synthetic as syn, tf_util
    params = od.init_params(od.S3DIS_LAYERS, seed=3)
    rng = np.random.default_rng(4)
    for k in params:   # non-trivial population statistics so inference BN does something
        if k.endswith("pop_mean"):
            params[k] = rng.normal(0, 0.1, params[k].shape).astype(np.float32)
        if k.endswith("pop_var"):
            params[k] = rng.uniform(0.5, 1.5, params[k].shape).astype(np.float32)
    DGCNN_S3DIS.set_variables(params)
    tf_util.VARIABLES.clear()
    tf_util.VARIABLES.update({k: torch.from_numpy(v).to(cuda) for k, v in params.items()})
    X, _, _, _ = syn.s3dis_batch(1, N=256, n_labelled=8, seed=9)
    Xd = torch.from_numpy(X).to(cuda)
    fused = DGCNN_S3DIS.get_model(Xd, False).clone()
    unfused = DGCNN_S3DIS.get_model_unfused(Xd, False)
    err = (fused - unfused).abs() / unfused.abs().max()
    assert float((err <= 1e-3).float().mean()) >= 0.99        # neighbour flips between the two device paths are local
    # and against the CPU oracle
    ref = od.get_model_s3dis(od.to_torch(params, requires_grad=False), torch.from_numpy(X), False)
    err2 = (unfused.cpu() - ref).abs() / ref.abs().max()
    assert float((err2 <= 1e-3).float().mean()) >= 0.99
