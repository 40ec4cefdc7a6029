"""Runs the REFERENCE'S OWN Util/SmoothConstraint.py (every variant, :9-219) and Util/Loss.py (:5-195), imported unmodified
from /root/reference, on the eager tf1_shim and stores what they return:
    python tests/golden/make_util_golden.py   ->   tests/golden/ref_util_variants.npz
The product's SmoothConstraint / Loss modules are held to these values (tests/test_util_variants_*.py).  Only this generator
reads /root/reference."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("WSPC_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "tf1_shim"))
sys.path.insert(1, os.path.join(REF, "Util"))

import tensorflow as tf  # noqa: E402  (the shim)

assert "tf1_shim" in tf.__file__


def _np(t):
    return np.asarray(t.detach().numpy() if hasattr(t, "detach") else t).copy()


def smooth_variants(out):
    import SmoothConstraint as SC

    rng = np.random.default_rng(401)
    B, N, C = 2, 200, 13
    X6 = np.concatenate([rng.uniform(-0.5, 0.5, (B, N, 3)), rng.uniform(0, 1, (B, N, 3))], -1).astype(np.float32)
    # the xyz and rgb graphs of Loss_SpatialColorSmooth_SelfContain only agree slot by slot (knn_mask) when colour follows
    # position: give half of the points a colour that is a function of their coordinates
    X6[:, ::2, 3:6] = X6[:, ::2, 0:3] + 0.5
    Zl = rng.normal(0, 1, (B, N, C)).astype(np.float32)
    P = (np.exp(Zl) / np.exp(Zl).sum(-1, keepdims=True)).astype(np.float32)
    W = rng.uniform(0.05, 1.0, (B, N, 5)).astype(np.float32)
    Ind = rng.integers(0, N, (B, N, 5)).astype(np.int32)
    tf.reset()
    out.update(sm_X6=X6, sm_P=P, sm_W=W, sm_Ind=Ind)
    out["Loss_SpatialSmooth"] = _np(SC.Loss_SpatialSmooth(tf.constant(X6[:, :, 0:3]), tf.constant(W), tf.constant(Ind)))
    out["Loss_SpatialSmooth_SelfContain"] = _np(SC.Loss_SpatialSmooth_SelfContain(tf.constant(X6[:, :, 0:3])))
    out["Loss_SpatialSmooth_SelfContain_g05_k7"] = _np(SC.Loss_SpatialSmooth_SelfContain(tf.constant(X6[:, :, 0:3]), gamma=0.5, knn=7))
    out["Loss_SpatialColorSmooth_SelfContain"] = _np(SC.Loss_SpatialColorSmooth_SelfContain(tf.constant(P), tf.constant(X6)))
    out["Loss_SpatialColorSmooth_add_SelfContain"] = _np(SC.Loss_SpatialColorSmooth_add_SelfContain(tf.constant(P), tf.constant(X6)))
    out["Loss_SpatialColorSmoothAdd_UnknownBatch_SelfContain"] = _np(
        SC.Loss_SpatialColorSmoothAdd_UnknownBatch_SelfContain(tf.constant(P), tf.constant(X6)))
    out["Loss_SpatialColorSmoothAdd_UnknownBatch_SelfContain_g1_k4"] = _np(
        SC.Loss_SpatialColorSmoothAdd_UnknownBatch_SelfContain(tf.constant(P), tf.constant(X6), gamma=1.0, knn=4))
    # how many slots of the two graphs coincide (the masked variant is vacuous if this is ~0)
    d = lambda A: ((A[:, :, None, :] - A[:, None, :, :]) ** 2).sum(-1)   # noqa: E731
    ix, ir = np.argsort(d(X6[..., 0:3]), -1, kind="stable")[..., :10], np.argsort(d(X6[..., 3:6]), -1, kind="stable")[..., :10]
    out["sm_mask_fraction"] = np.array([(ix == ir).mean()])


def loss_module(out):
    import Loss as RL

    rng = np.random.default_rng(411)
    B, N, K = 3, 50, 6
    L = rng.normal(0, 2, (B, N, K)).astype(np.float32)
    L += np.array([9, -9, 0, 0, 0, 0], np.float32)      # a dominant and a dominated class: the Overwhelm penalties are non-zero
    Ypt = (rng.random((B, N, K)) < 0.3).astype(np.float32)
    Ycl = (rng.random((B, K)) < 0.5).astype(np.float32)
    Ycl[0, 0:2] = 1
    Ycl[1, 0:2] = (0, 1)
    alpha = rng.uniform(0.1, 0.9, (B, N, K)).astype(np.float32)
    pw, nw = rng.uniform(0.5, 2, (K,)).astype(np.float32), rng.uniform(0.5, 2, (K,)).astype(np.float32)
    out.update(ls_L=L, ls_Ypt=Ypt, ls_Ycl=Ycl, ls_alpha=alpha, ls_pw=pw, ls_nw=nw)
    c = tf.constant
    out["focal_loss"] = _np(RL.focal_loss(c(L), c(Ypt)))
    out["focal_loss_a4_g3"] = _np(RL.focal_loss(c(L), c(Ypt), alpha=0.4, gamma=3))
    out["focal_loss_v1"] = _np(RL.focal_loss_v1(c(L), c(Ypt)))
    out["focal_loss_v1_alpha"] = _np(RL.focal_loss_v1(c(L), c(Ypt), alpha=c(alpha)))
    out["class_weighted_CE_loss"] = _np(RL.class_weighted_CE_loss(c(L[:, :1]), c(Ypt[:, :1]), c(pw), c(nw)))
    out["SelfEntropy"] = _np(RL.SelfEntropy(c(L)))
    out["OverwhelmLoss_v1"] = _np(RL.OverwhelmLoss_v1(c(L), c(Ycl)))
    l2, p2, n2 = RL.OverwhelmLoss_v2(c(L), c(Ycl))
    out.update(OverwhelmLoss_v2=_np(l2), OverwhelmLoss_v2_pos=_np(p2), OverwhelmLoss_v2_neg=_np(n2))
    l3, f3 = RL.OverwhelmLoss(c(L), c(Ycl))
    out.update(OverwhelmLoss=_np(l3), OverwhelmLoss_full=_np(f3))


def unnorm_xyz_model(out):
    """DGCNN_S3DIS.get_model_unnormXYZ (S3DIS/DGCNN_S3DIS.py:106-186): the first graph is built on the metric xyz channels
    0:3 instead of the room-normalised 6:9; inference mode on seeded variables, neighbour lists recorded."""
    sys.path.insert(1, os.path.join(REF, "S3DIS"))
    sys.path.insert(1, os.path.join(REF, "Networks/dgcnn/utils"))
    sys.path.insert(1, os.path.dirname(os.path.dirname(HERE)))
    sys.path.insert(1, HERE)
    import DGCNN_S3DIS as network
    from refgen_common import S3DIS_LAYERS, xavier_params
    from weaksuppointcloudseg_b200 import synthetic as syn
    X, _, _, _ = syn.s3dis_batch(1, N=192, n_labelled=8, seed=421)
    params0 = xavier_params(S3DIS_LAYERS, seed=422)
    tf.reset()
    tf.preset_variables(params0)
    Z = network.get_model_unnormXYZ(tf.constant(X), tf.constant(False), weight_decay=0., bn_decay=None)
    tk = tf.RECORD["top_k"][-3:]
    out.update(ux_X=X, ux_param_seed=np.array([422]), ux_Z=_np(Z), ux_knn1=tk[0].astype(np.int16), ux_knn2=tk[1].astype(np.int16),
               ux_knn3=tk[2].astype(np.int16))


def classification_model(out):
    """Networks/dgcnn/models/dgcnn.py:20-109: the ModelNet classification DGCNN (get_model in inference mode on seeded variables
    with non-trivial BN statistics and a non-identity input transform) and get_loss (label smoothing 0.2)."""
    sys.path.insert(1, os.path.join(REF, "Networks/dgcnn/models"))
    sys.path.insert(1, os.path.join(REF, "Networks/dgcnn/utils"))
    sys.path.insert(1, HERE)
    import dgcnn as network
    from refgen_common import CLS_LAYERS, xavier_params
    rng = np.random.default_rng(431)
    B, N = 3, 160
    X = rng.uniform(-1, 1, (B, N, 3)).astype(np.float32)
    lab = rng.integers(0, 40, (B,)).astype(np.int32)
    params0 = xavier_params(CLS_LAYERS, seed=432, tnet_seed=433)
    tf.reset()
    tf.preset_variables(params0)
    Z, _ = network.get_model(tf.constant(X), tf.constant(False), bn_decay=None)
    loss = network.get_loss(Z, tf.constant(lab), {})
    tk = tf.RECORD["top_k"][-5:]
    out.update(cls_X=X, cls_label=lab, cls_param_seed=np.array([432, 433]), cls_Z=_np(Z), cls_loss=_np(loss))
    for i, t in enumerate(tk):
        out["cls_knn%d" % i] = t.astype(np.int16)
    missing = [k for k in tf.STATE["variables"] if k not in params0]
    assert not missing, missing


if __name__ == "__main__":
    out = {}
    smooth_variants(out)
    loss_module(out)
    unnorm_xyz_model(out)
    classification_model(out)
    np.savez_compressed(os.path.join(HERE, "ref_util_variants.npz"), **out)
    for k, v in out.items():
        if v.size == 1:
            print(k, float(v.reshape(-1)[0]))
