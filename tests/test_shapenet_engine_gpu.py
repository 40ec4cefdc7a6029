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
    return dict(eng=eng, out=out, rec=rec, p=p, losses=losses.cpu().numpy(), params0=params, X=X, lab=lab, Y=Y, M=M, masks=masks)


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


def test_gradients(setup, cuda):
    """Every trainable tensor against the fp64 oracle that takes the ENGINE's discrete branches (ReLU masks, max-over-k
    splits, arg-max rows of the two max-over-points stages; tests/routing.py): with the routing pinned the comparison is a
    statement about arithmetic, and the bound is the parity bar of SURVEY 8(c), 1e-3 max-rel, for EVERY tensor (measured:
    5e-6 .. 5.8e-4; un-forced the same comparison scatters by 3e-2 .. 1e-1 on the T-net, which is why round 1 carried a 3e-1
    bound and an escape hatch here)."""
    import routing
    from weaksuppointcloudseg_b200 import runtime as rt
    from weaksuppointcloudseg_b200.engine_shapenet import ShapeNetEngine
    params, X, lab, Y, M, masks = (setup[n] for n in ("params0", "X", "lab", "Y", "M", "masks"))
    n_samples, N = X.shape[0] // 2, X.shape[1]
    B = X.shape[0]
    eng = ShapeNetEngine(params, B, N, device=cuda)
    rt.ROUTING = {}
    try:
        eng.train_step(*(torch.from_numpy(a).to(cuda) for a in (X, lab, Y, M)), lr=1e-3, bn_decay=od.bn_decay(0, n_samples, 16881 * 20),
                       dropout_masks=[torch.from_numpy(m).to(cuda) for m in masks], apply=False)
        torch.cuda.synchronize()
        route = routing.export_shapenet(eng, rt.ROUTING)
    finally:
        rt.ROUTING = None
    ov = {f"knn{i}": eng.idx[i].cpu().long() for i in range(4)}
    sg = (eng.idxS.cpu().long(), torch.exp(-eng.dS.cpu().double() / 0.1))
    p64 = od.to_torch(params, dtype=torch.float64)
    f64 = lambda a: torch.from_numpy(a).to(torch.float64)   # noqa: E731
    with od.forced_routing(route):
        ref = od.train_step_shapenet(p64, od.AdamTF(p64, od.trainable_names(p64)), f64(X), f64(lab), f64(Y), f64(M), step=0,
                                     dropout_masks=[f64(m) for m in masks], knn_override=ov, smooth_graph_=sg)
    assert rel(eng.Z.cpu().numpy(), ref["Z"].detach().numpy()) <= TOL
    got = eng.vs.grads()
    gmax = max(float(g.abs().max()) for g in ref["grads"].values() if g is not None)
    worst = {}
    for name, g in ref["grads"].items():
        a, b = got[name].astype(np.float64), g.numpy()
        if np.abs(b).max() < 1e-9 * gmax:     # analytically zero gradients (biases of batch-normalised layers)
            assert np.abs(a).max() < 1e-5 * gmax, name
            continue
        worst[name] = rel(a, b)
    print("ShapeNet (16 x 256) forced-routing gradient errors:", {k_: f"{v:.1e}" for k_, v in worst.items()})
    bad = {k_: v for k_, v in worst.items() if v > TOL}
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
