"""S3DIS DGCNN: CUDA engine (through the C ABI) vs the CPU oracle on the same seeded inputs.

Tolerance (the reference defines none): per tensor max|a-b| / max|b| <= 1e-3, fp32.
kNN-2/3 consume computed features, so logits are compared with teacher forcing (the engine is fed the
oracle's neighbour lists); kNN itself is checked bit-exactly stage-wise (engine features -> oracle kNN).
Biases of BN'd convs have an analytically zero gradient (pure rounding noise) and are excluded
(SURVEY §7.3-8).

Gradients: ReLU masks and the arg-max of the max-over-k / max-over-N pools are discontinuous, so two
correct fp32 implementations whose activations differ in the last bits route a handful of single-element
gradients differently (the fp32 oracle shows the same ~1e-3 scatter against its own fp64 run, see
tools/diag_grads.py), and the effect is largest at the tiny sizes the CPU oracle can afford (it shrinks
like 1/sqrt(rows)).  Weight gradients are therefore held to ||a-b||_2/||b||_2 <= 2e-2 and
max|a-b|/max|b| <= 6e-2 here, while tests/test_kernels_gpu.py pins every backward kernel to 1e-5 against
fp64 formulas evaluated on identical inputs (no discontinuity in play).
"""
import numpy as np
import pytest
import torch

from oracle import dgcnn as od
from oracle import knn as oknn
from weaksuppointcloudseg_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
TOL = 1e-3


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


@pytest.fixture(scope="module")
def setup(cuda):
    from weaksuppointcloudseg_b200.engine_s3dis import S3DISEngine

    n_samples, N = 3, 384
    X, Y, M, _ = syn.s3dis_batch(n_samples, N=N, n_labelled=12, seed=77)
    B = 2 * n_samples
    params = od.init_params(od.S3DIS_LAYERS, seed=5)
    rng = np.random.default_rng(9)
    for k in params:  # non-trivial BN affine so gamma/beta gradients are exercised
        if k.endswith("gamma"):
            params[k] = rng.uniform(0.5, 1.5, params[k].shape).astype(np.float32)
        if k.endswith("beta"):
            params[k] = rng.uniform(-0.2, 0.2, params[k].shape).astype(np.float32)
    mask = np.floor(0.7 + rng.random((B, N, 256))).astype(np.float32)
    # oracle
    p = od.to_torch(params)
    opt = od.AdamTF(p, od.trainable_names(p))
    rec = {}
    out = od.train_step_s3dis(p, opt, torch.from_numpy(X), torch.from_numpy(Y), torch.from_numpy(M), step=0,
                              dropout_mask=torch.from_numpy(mask), rec=rec)
    # engine, teacher-forced kNN 2/3
    eng = S3DISEngine(params, B, N, device=cuda)
    ov = {f"knn{i}": rec[f"knn{i}/idx"].to(torch.int32).to(cuda) for i in (2, 3)}
    Xd, Yd, Md = (torch.from_numpy(a).to(cuda) for a in (X, Y, M))
    losses = eng.train_step(Xd, Yd, Md, lr=1e-3, bn_decay=od.bn_decay(0, n_samples, 300000),
                            dropout_mask=torch.from_numpy(mask).to(cuda), knn_override=ov)
    torch.cuda.synchronize()
    return dict(eng=eng, out=out, rec=rec, p=p, losses=losses.cpu().numpy(), X=X, Y=Y, M=M, params0=params, mask=mask)


def test_knn1_and_smooth_graph_bit_exact(setup):
    eng, rec, X = setup["eng"], setup["rec"], setup["X"]
    assert np.array_equal(eng.idx[0].cpu().numpy(), rec["knn1/idx"].numpy().astype(np.int32))
    idx, d = oknn.knn(X, 10, oknn.SMOOTH, coff=0, D=6, return_dist=True)
    assert np.array_equal(eng.idxS.cpu().numpy(), idx)
    assert np.array_equal(eng.dS.cpu().numpy(), d)


def test_features_and_stagewise_knn(setup):
    eng, rec = setup["eng"], setup["rec"]
    B, N = eng.B, eng.N
    cat = eng.cat.cpu().numpy().reshape(B, N, 192)
    for i, name in enumerate(["net_1", "net_2", "net_3"]):
        assert rel(cat[:, :, 64 * i:64 * (i + 1)], rec[name].detach().numpy()) <= TOL
    # stage-wise kNN: the oracle's kNN on the ENGINE's features == the engine's own fused kNN on them
    from weaksuppointcloudseg_b200 import ops
    for i in range(2):
        feats = np.ascontiguousarray(cat[:, :, 64 * i:64 * (i + 1)])
        got = ops.knn_fused(torch.from_numpy(feats).to(eng.dev), 20).cpu().numpy()
        assert np.array_equal(got, oknn.knn(feats, 20))


def test_logits_probs_losses(setup):
    eng, out = setup["eng"], setup["out"]
    assert rel(eng.Z.cpu().numpy(), out["Z"].detach().numpy()) <= TOL
    assert rel(eng.Zp.cpu().numpy(), out["Z_prob"].detach().numpy()) <= TOL
    names = ["loss_seg", "loss_siamese", "loss_inexact", "loss_smooth", "loss"]
    for v, n in zip(setup["losses"], names):
        ref = float(out[n])
        assert abs(v - ref) <= TOL * abs(ref), (n, v, ref)


def test_gradients(setup, cuda):
    """Every trainable tensor against the fp64 oracle that takes the ENGINE's discrete branches (ReLU masks, max-over-k splits,
    arg-max rows of the max over points; tests/routing.py): max-rel <= 1e-3, the parity bar of SURVEY 8(c).  (Round 1 compared
    un-forced at 6e-2 max-rel / 2e-2 in L2: two correct fp32 implementations route near-tied maxima differently.)"""
    import routing
    from weaksuppointcloudseg_b200 import runtime as rt
    from weaksuppointcloudseg_b200.engine_s3dis import S3DISEngine
    params, X, Y, M, mask = (setup[n] for n in ("params0", "X", "Y", "M", "mask"))
    B, N = X.shape[0], X.shape[1]
    eng = S3DISEngine(params, B, N, device=cuda)
    rt.ROUTING = {}
    try:
        eng.train_step(torch.from_numpy(X).to(cuda), torch.from_numpy(Y).to(cuda), torch.from_numpy(M).to(cuda), lr=1e-3,
                       bn_decay=od.bn_decay(0, B // 2, 300000), dropout_mask=torch.from_numpy(mask).to(cuda), apply=False)
        torch.cuda.synchronize()
        route = routing.export_s3dis(eng, rt.ROUTING)
    finally:
        rt.ROUTING = None
    ov = {f"knn{i + 1}": eng.idx[i].cpu().long() for i in range(3)}
    sg = (eng.idxS.cpu().long(), torch.exp(-eng.dS.cpu().double() / 0.1))
    p64 = od.to_torch(params, dtype=torch.float64)
    f64 = lambda a: torch.from_numpy(a).double()   # noqa: E731
    with od.forced_routing(route):
        ref = od.train_step_s3dis(p64, od.AdamTF(p64, od.trainable_names(p64)), f64(X), f64(Y), f64(M), step=0,
                                  dropout_mask=f64(mask), knn_override=ov, smooth_graph_=sg)
    assert rel(eng.Z.cpu().numpy(), ref["Z"].detach().numpy()) <= 2e-4
    got = eng.vs.grads()
    gmax = max(float(g.abs().max()) for g in ref["grads"].values())
    worst = {}
    for name, g in ref["grads"].items():
        a, b = got[name].astype(np.float64), g.numpy()
        if np.abs(b).max() < 1e-9 * gmax:
            # analytically zero (biases of BN'd convs; adj_conv7's beta, whose per-channel shift of the tiled
            # global feature is removed again by seg/conv1's BN)
            assert np.abs(a).max() < 1e-5 * gmax, name
            continue
        worst[name] = rel(a, b)
    print("S3DIS (6 x 384) forced-routing gradient errors:", {k_: f"{v:.1e}" for k_, v in worst.items()})
    bad = {k_: v for k_, v in worst.items() if v > TOL}
    assert not bad, bad


def test_adam_and_pop_stats(setup):
    """First TF-Adam step moves every weight by ~lr*sign(g): compare the update where the
    gradient is not rounding noise; population BN statistics must match to TOL (tf_util.py:524-525)."""
    eng, p, out = setup["eng"], setup["p"], setup["out"]
    got = eng.vs.export()
    gmax = max(float(g.abs().max()) for g in out["grads"].values())
    for name in got:
        ref = p[name].detach().numpy()
        if name.endswith("pop_mean") or name.endswith("pop_var"):
            assert rel(got[name], ref) <= TOL, name
            continue
        g = out["grads"][name].numpy()
        d_ref = ref - setup["params0"][name]
        d_got = got[name] - setup["params0"][name]
        assert np.abs(d_got).max() <= 1.001e-3, name
        sig = np.abs(g) > 1e-2 * np.abs(g).max()
        # analytically-zero gradients (biases of batch-normalised convs; adj_conv7's beta): the update is the sign of rounding
        # noise.  Named explicitly: the fp32 oracle's own noise on them depends on torch's CPU reduction order (thread count)
        # and sat right at the 1e-6 threshold below
        zero_by_construction = (name.endswith("/biases") and name != "seg/conv3/biases") or name == "adj_conv7/bn/beta"
        if zero_by_construction or np.abs(g).max() < 1e-6 * gmax or not sig.any():
            continue
        assert np.mean(np.abs(d_got[sig] - d_ref[sig]) <= 2e-5) >= 0.999, name


def test_inference_mode_uses_population_stats(setup, cuda):
    """is_training=False: BN uses pop stats (tf_util.py:529-530), dropout is the identity (:632-634)."""
    eng, p = setup["eng"], setup["p"]
    X = torch.from_numpy(setup["X"])
    rec = {}
    Zref = od.get_model_s3dis(p, X, False, rec=rec)
    # engine weights differ slightly after its own Adam step -> load the oracle's
    eng.vs.load({k: v.detach().numpy() for k, v in p.items()})
    ov = {f"knn{i}": rec[f"knn{i}/idx"].to(torch.int32).to(cuda) for i in (2, 3)}
    Z = eng.forward(X.to(cuda), False, knn_override=ov)
    assert rel(Z.cpu().numpy(), Zref.detach().numpy()) <= TOL


def test_siamese_zero_for_identical_pairs(setup, cuda):
    """SURVEY §4 invariant 3: identical pair members, no dropout -> loss_siamese == 0."""
    eng = setup["eng"]
    X = torch.from_numpy(setup["X"]).to(cuda).clone()
    X[1::2] = X[0::2]
    eng.forward(X, False)
    B, N = eng.B, eng.N
    Y = torch.zeros((B, N, 13), device=cuda)
    Y[..., 0] = 1
    M = torch.ones((B, N), device=cuda)
    l = eng.losses_and_grad(Y, M, full=True, want_grad=True).cpu().numpy()
    assert l[1] == 0.0
    assert l[3] >= 0.0


def test_closed_gate_head_pass(setup, cuda):
    """full=2 (Full graph, ramp-up gate closed, S3DIS_DGCNN_trainer.py:100-102): one head pass returns the four loss values
    of the Full graph and exactly the Plain-style gradient."""
    eng = setup["eng"]
    X = torch.from_numpy(setup["X"]).to(cuda)
    Y = torch.from_numpy(setup["Y"]).to(cuda)
    M = torch.from_numpy(setup["M"]).to(cuda)
    eng.forward(X, False)
    l_full = eng.losses_and_grad(Y, M, full=True, want_grad=True).clone()
    dz_full = eng.dZ.clone()
    eng.losses_and_grad(Y, M, full=False, want_grad=True)
    dz_plain = eng.dZ.clone()
    eng.dZ.fill_(7.0)
    l_closed = eng.losses_and_grad(Y, M, full=2, want_grad=True).clone()
    assert torch.equal(eng.dZ, dz_plain)
    assert float((dz_full - dz_plain).abs().max()) > 0
    assert torch.allclose(l_closed[:4], l_full[:4], rtol=1e-6, atol=0)   # fp64 atomics: order-dependent last bit only


def test_cfg4_stress_shape_properties(cuda):
    """BASELINE cfg-4 (N=8192, k=40; one Siamese pair): size-independent properties of a full train step --
    neighbour lists sorted / self-inclusive / duplicate-free, finite losses, Siamese term > 0, every gradient finite,
    and the weights move.  (k=40 exercises the two-slot CUDA-core kNN and the generic k of every edge kernel.)"""
    from weaksuppointcloudseg_b200.engine_s3dis import S3DISEngine
    from weaksuppointcloudseg_b200 import ops
    N, k = 8192, 40
    X, Y, M, _ = syn.s3dis_batch(1, N=N, n_labelled=81, seed=44, dup_frac=0.0)
    eng = S3DISEngine(od.init_params(od.S3DIS_LAYERS, seed=3), 2, N, device=cuda, k=k)
    Xd, Yd, Md = (torch.from_numpy(a).to(cuda) for a in (X, Y, M))
    before = eng.vs.theta.clone()
    losses = eng.train_step(Xd, Yd, Md, lr=1e-3, bn_decay=0.5).cpu().numpy()
    torch.cuda.synchronize()
    assert np.isfinite(losses).all() and losses[1] > 0 and losses[4] > 0
    idx = eng.idx[0]
    _, dist = ops.knn_fused(Xd, k, ops.DIST_TFUTIL, coff=6, D=3, return_dist=True)
    assert bool((dist[..., 1:] >= dist[..., :-1]).all())
    # with the reference's formula (sq_i - 2 x_i.x_j) + sq_j a close neighbour can come out at a slightly negative
    # distance and precede the point itself (d_ii == 0 exactly), so only self-INCLUSION is a property of the lists
    me = torch.arange(N, device=cuda, dtype=torch.int32).view(1, N, 1)
    assert float((idx == me).any(-1).float().mean()) >= 0.999
    s = torch.sort(idx, dim=-1).values
    assert bool((s[..., 1:] != s[..., :-1]).all())
    assert bool(torch.isfinite(eng.vs.grad).all())
    assert float((eng.vs.theta - before).abs().max()) > 0
