"""ShapeNet part-seg DGCNN (T-net + category branch): CUDA engine vs the CPU oracle, same protocol and
tolerances as tests/test_s3dis_engine_gpu.py.  Reference: ShapeNet/DGCNN_ShapeNet.py:15-113,
Networks/dgcnn/models/transform_nets.py:10-56, ShapeNet/ShapeNet_DGCNN_trainer.py:85-133."""
import numpy as np
import pytest
import torch

from oracle import dgcnn as od
from weaksuppointcloudseg_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
TOL = 1e-3


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


@pytest.fixture(scope="module")
def setup(cuda):
    from weaksuppointcloudseg_b200.engine_shapenet import ShapeNetEngine

    # 8 Siamese samples = 16 clouds: the T-net's FC layers batch-normalise over the clouds, and with a handful of
    # clouds one ReLU flip of a near-zero pre-activation moves a whole channel's gradient by several percent
    n_samples, N = 8, 256
    X, lab, Y, M, _ = syn.shapenet_batch(n_samples, N=N, n_labelled=32, seed=21)
    B = 2 * n_samples
    params = od.init_params(od.SHAPENET_LAYERS, seed=8, shapenet=True)
    rng = np.random.default_rng(2)
    # a non-identity transform so the T-net path is exercised (the reference starts it at exactly I)
    params["transform_net1/transform_XYZ/weights"] = rng.normal(0, 0.02, (256, 9)).astype(np.float32)
    params["transform_net1/transform_XYZ/biases"] = rng.normal(0, 0.05, (9,)).astype(np.float32)
    masks = [np.floor(0.6 + rng.random((B, N, 256))).astype(np.float32) for _ in range(2)]
    p = od.to_torch(params)
    opt = od.AdamTF(p, od.trainable_names(p))
    rec = {}
    out = od.train_step_shapenet(p, opt, torch.from_numpy(X), torch.from_numpy(lab), torch.from_numpy(Y),
                                 torch.from_numpy(M), step=0, dropout_masks=[torch.from_numpy(m) for m in masks], rec=rec)
    eng = ShapeNetEngine(params, B, N, device=cuda)
    ov = {f"knn{i}": rec[f"knn{i}/idx"].to(torch.int32).to(cuda) for i in (1, 2, 3)}
    losses = eng.train_step(torch.from_numpy(X).to(cuda), torch.from_numpy(lab).to(cuda), torch.from_numpy(Y).to(cuda),
                            torch.from_numpy(M).to(cuda), lr=1e-3, bn_decay=od.bn_decay(0, n_samples, 16881 * 20),
                            dropout_masks=[torch.from_numpy(m).to(cuda) for m in masks], knn_override=ov)
    torch.cuda.synchronize()
    # fp64 run of the same oracle on the same neighbour graphs: the yardstick for how far two correct fp32
    # implementations may drift apart on the discontinuous (ReLU / arg-max) gradient paths
    p64 = od.to_torch(params, dtype=torch.float64)
    f64 = lambda a: torch.from_numpy(a).to(torch.float64)
    out64 = od.train_step_shapenet(p64, od.AdamTF(p64, od.trainable_names(p64)), f64(X), f64(lab), f64(Y), f64(M), step=0,
                                   dropout_masks=[f64(m) for m in masks], rec={},
                                   knn_override={f"knn{i}": rec[f"knn{i}/idx"] for i in (0, 1, 2, 3)},
                                   smooth_graph_=od.smooth_graph(torch.from_numpy(X)))
    return dict(eng=eng, out=out, out64=out64, rec=rec, p=p, losses=losses.cpu().numpy(), params0=params, X=X)


def test_tnet_and_knn0(setup):
    eng, rec = setup["eng"], setup["rec"]
    assert np.array_equal(eng.idx[0].cpu().numpy(), rec["knn0/idx"].numpy().astype(np.int32))   # bit-exact
    T = eng.Tm.cpu().numpy().reshape(-1, 3, 3) + np.eye(3, dtype=np.float32)
    assert rel(T, rec["transform"].detach().numpy()) <= TOL
    assert rel(eng.Xt.cpu().numpy(), rec["pct"].detach().numpy()) <= TOL


def test_logits_and_losses(setup):
    eng, out = setup["eng"], setup["out"]
    assert rel(eng.Z.cpu().numpy(), out["Z"].detach().numpy()) <= TOL
    for v, n in zip(setup["losses"], ["loss_seg", "loss_siamese", "loss_inexact", "loss_smooth", "loss"]):
        ref = float(out[n].detach())
        assert abs(v - ref) <= TOL * abs(ref), (n, v, ref)


def test_gradients(setup):
    eng, out = setup["eng"], setup["out"]
    got = eng.vs.grads()
    gmax = max(float(g.abs().max()) for g in out["grads"].values() if g is not None)
    bad = {}
    for name, g in out["grads"].items():
        a, b = got[name].astype(np.float64), g.numpy().astype(np.float64)
        bn_bias = name.endswith('/biases') and not (name.startswith('seg/conv4') or 'transform_XYZ' in name)
        if bn_bias or np.abs(b).max() < 1e-6 * gmax:     # analytically zero gradients: rounding noise on both sides
            assert np.abs(a).max() < 1e-4 * gmax, name
            continue
        e = (rel(a, b), np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
        c = setup["out64"]["grads"][name].numpy()
        l2 = lambda u, v: np.linalg.norm(u - v) / max(np.linalg.norm(v), 1e-30)
        # ReLU masks / arg-max pools are discontinuous: activations that differ in the last bits route single-element
        # gradients differently, and that noise compounds backwards through the 16 BN'd layers of this net (it is
        # 1e-4 at seg/conv4 and a few 1e-2 at the T-net, whose gradients are all proportional to the single
        # (B,3,3) tensor dT = X^T dX' and whose FC layers normalise over only B=6 clouds here).  The fp32 oracle
        # itself scatters by ~1 % against its own fp64 run on these tensors (tools/diag_shapenet.py).  The bound
        # below (||a-b||/||b|| <= 6e-2, i.e. cosine >= 0.998) still catches any wiring / scaling / indexing error,
        # and tests/test_kernels_gpu.py pins every kernel to 1e-5 on identical inputs.
        lim = (3e-1, 6e-2)     # (max-norm: one arg-max flip of max_pool2d moves a whole gradient row)
        # A tensor beyond that fixed bound still passes when the engine is as close to the exact (fp64) gradient
        # as the fp32 oracle itself is, within a factor 3: then the gap is fp32 routing noise, not an error.
        as_good_as_fp32 = l2(a, c) <= 3.0 * l2(b, c) + 1e-3
        if (e[0] > lim[0] or e[1] > lim[1]) and not as_good_as_fp32:
            bad[name] = e + (l2(a, c), l2(b, c))
    assert not bad, bad


def test_identity_transform_at_init(cuda):
    """SURVEY §4 invariant 1: with the reference initialisation (W=0, b=0 + eye) the T-net output is exactly I,
    so X' == X bit-exactly and kNN-1 == kNN-0."""
    from weaksuppointcloudseg_b200.engine_shapenet import ShapeNetEngine
    X, lab, Y, M, _ = syn.shapenet_batch(1, N=256, n_labelled=16, seed=5)
    eng = ShapeNetEngine(od.init_params(od.SHAPENET_LAYERS, seed=1, shapenet=True), 2, 256, device=cuda)
    eng.forward(torch.from_numpy(X).to(cuda), torch.from_numpy(lab).to(cuda), True, 0.5)
    assert torch.equal(eng.Xt.cpu(), torch.from_numpy(X))
    assert torch.equal(eng.idx[0], eng.idx[1])
