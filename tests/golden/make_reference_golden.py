"""Runs the REFERENCE'S OWN PYTHON (imported unmodified from /root/reference) on seeded inputs and stores what it
computes as golden fixtures:  `python tests/golden/make_reference_golden.py`  ->  tests/golden/ref_*.npz

TensorFlow 1.14 is not installable in this image, so `import tensorflow` resolves to tests/golden/tf1_shim (an eager
op-by-op stand-in with the documented TF-1.14 semantics, see its docstring).  What these fixtures pin is therefore the
reference's graph construction -- layer order, scopes, shapes, argument order, loss formulas, optimiser wiring, every
quirk of tf_util / DGCNN_S3DIS / DGCNN_ShapeNet / transform_nets / SmoothConstraint / Tool / ProbLabelPropagation /
the trainers' defineNetwork + WeakSupLoss -- executed by the reference's own code, not restated.  The per-op
arithmetic of TensorFlow's binary kernels remains the published semantics.

Nothing under tests/ or the product reads /root/reference at test time: only this generator does.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("WSPC_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "tf1_shim"))
for sub in ("S3DIS", "ShapeNet", "Util", "Networks/dgcnn/utils", "Networks/dgcnn/models"):
    sys.path.insert(1, os.path.join(REF, sub))
sys.path.insert(1, ROOT)
sys.path.insert(1, HERE)

import tensorflow as tf  # noqa: E402  (the shim)
from weaksuppointcloudseg_b200 import synthetic as syn  # noqa: E402  (numpy-only input generators)

assert "tf1_shim" in tf.__file__


from refgen_common import S3DIS_LAYERS, SHAPENET_LAYERS, subsample, xavier_params  # noqa: E402


def _np(t):
    return t.detach().numpy().copy()


def _collect(tr, params0):
    """losses / logits, gradients (large tensors subsampled, L2 norm kept), BN population statistics and the updated
    value of every small trainable tensor after the Adam step."""
    out = {}
    for n in ("Z", "Z_prob", "loss_seg", "loss_siamese", "loss_inexact", "loss_smooth", "loss"):
        out[n] = _np(getattr(tr, n))
    for n, g in tf.STATE["last_grads"].items():
        if g is not None and n in params0:
            full = g.numpy().reshape(params0[n].shape)
            out["grad/" + n], _ = subsample(full)
            out["gradnorm/" + n] = np.array([np.linalg.norm(full.astype(np.float64))])
    for n, v in tf.STATE["variables"].items():
        if n in params0 and (n.endswith("pop_mean") or n.endswith("pop_var") or params0[n].size <= 8192):
            out["after/" + n] = _np(v).reshape(params0[n].shape)
    return out


def s3dis_train_step():
    """One `sess.run([solver, losses, Z])` of S3DIS_Trainer (S3DIS_DGCNN_trainer.py:57-137, Full style, ramp-up gate
    open), preceded by an inference-mode build of DGCNN_S3DIS.get_model on the same initial variables."""
    import S3DIS_DGCNN_trainer as trainer_mod
    import DGCNN_S3DIS as network

    n_samples, N, bs = 2, 160, 2
    X, Y, M, _ = syn.s3dis_batch(n_samples, N=N, n_labelled=8, seed=301)
    B = 2 * n_samples
    params0 = xavier_params(S3DIS_LAYERS, seed=302)
    keep = np.floor(0.7 + np.random.default_rng(303).random((B, N, 1, 256))).astype(np.float32)

    tf.reset()
    tf.preset_variables(params0)
    out = dict(X=X, Y=Y.astype(np.uint8), Mask=M, dropout_keep=np.packbits(keep.reshape(B, N, 256).astype(np.uint8), axis=-1),
               batch_size=np.array([bs]), param_seed=np.array([302]))
    # inference graph (is_training=False: population statistics, no dropout) on the initial variables ...
    Zt = network.get_model(tf.constant(X), tf.constant(False), weight_decay=0., bn_decay=None)
    out["Z_eval"] = _np(Zt)
    out["knn1_eval"], out["knn2_eval"], out["knn3_eval"] = [t.astype(np.int16) for t in tf.RECORD["top_k"][0:3]]
    # ... then the trainer builds (= runs) its training graph on the same variables
    tf.feed(InputPts=X, PartGT=Y, Mask=M, IsTraining=True)
    tf.DROPOUT_MASKS.append(keep)
    tr = trainer_mod.S3DIS_Trainer(test_area=5)
    tr.SetLearningRate(1e-3, bs)
    tr.defineNetwork(B, N, style="Full", rampup=0)
    out.update(_collect(tr, params0))
    tk = tf.RECORD["top_k"][3:]
    assert len(tk) == 4 and [t.shape[-1] for t in tk] == [20, 20, 20, 10], [t.shape for t in tk]
    out.update(knn1=tk[0].astype(np.int16), knn2=tk[1].astype(np.int16), knn3=tk[2].astype(np.int16),
               knn_smooth=tk[3].astype(np.int16))
    np.savez_compressed(os.path.join(HERE, "ref_s3dis_step.npz"), **out)
    print("ref_s3dis_step: loss", float(out["loss"]), "seg/siam/inexact/smooth",
          float(out["loss_seg"]), float(out["loss_siamese"]), float(out["loss_inexact"]), float(out["loss_smooth"]))


def shapenet_train_step():
    """One training `sess.run` of ShapeNet_Trainer (ShapeNet_DGCNN_trainer.py:56-133, Full style)."""
    import ShapeNet_DGCNN_trainer as trainer_mod

    n_samples, N, bs = 3, 192, 3
    X, lab, Y, M, _ = syn.shapenet_batch(n_samples, N=N, n_labelled=20, seed=311)
    B = 2 * n_samples
    rng = np.random.default_rng(312)
    params0 = xavier_params(SHAPENET_LAYERS, seed=313, tnet_seed=314)
    keeps = [np.floor(0.6 + rng.random((B, N, 1, 256))).astype(np.float32) for _ in range(2)]

    tf.reset()
    tf.preset_variables(params0)
    tf.feed(InputPts=X, PartGT=Y.astype(np.int32), Mask=M, IsTraining=True, ShapeGT=lab)
    tf.DROPOUT_MASKS.extend(keeps)
    tr = trainer_mod.ShapeNet_Trainer()
    tr.SetLearningRate(1e-3, bs)
    tr.defineNetwork(B, point_num=N, style="Full", rampup=0)
    out = dict(X=X, label=lab, Y=Y.astype(np.uint8), Mask=M, batch_size=np.array([bs]), param_seed=np.array([313, 314]),
               dropout_keep1=np.packbits(keeps[0].reshape(B, N, 256).astype(np.uint8), axis=-1),
               dropout_keep2=np.packbits(keeps[1].reshape(B, N, 256).astype(np.uint8), axis=-1))
    out.update(_collect(tr, params0))
    tk = tf.RECORD["top_k"]
    assert len(tk) == 5 and [t.shape[-1] for t in tk] == [20, 20, 20, 20, 10], [t.shape for t in tk]
    out.update(knn0=tk[0].astype(np.int16), knn1=tk[1].astype(np.int16), knn2=tk[2].astype(np.int16),
               knn3=tk[3].astype(np.int16), knn_smooth=tk[4].astype(np.int16))
    np.savez_compressed(os.path.join(HERE, "ref_shapenet_step.npz"), **out)
    print("ref_shapenet_step: loss", float(out["loss"]))


def unit_ops():
    """tf_util.pairwise_distance / knn / get_edge_feature, SmoothConstraint, Tool.batch_gather_v1 on their own."""
    import tf_util
    import SmoothConstraint
    import Tool

    rng = np.random.default_rng(321)
    tf.reset()
    X, _, _, _ = syn.s3dis_batch(1, N=200, n_labelled=8, seed=322)          # (2, 200, 9) with duplicated points
    pc = tf.constant(X[:, :, 6:9])
    adj = tf_util.pairwise_distance(pc)
    idx = tf_util.knn(adj, k=20)
    ef = tf_util.get_edge_feature(tf.expand_dims(tf.constant(X), -2), nn_idx=idx, k=20)
    one = tf.constant(X[:1, :, 6:9])                                       # batch of one: the squeeze/expand quirk (:647-650)
    adj1 = tf_util.pairwise_distance(one)
    Z = rng.normal(0, 1, (2, 200, 13)).astype(np.float32)
    P = np.exp(Z) / np.exp(Z).sum(-1, keepdims=True)
    sm = SmoothConstraint.Loss_SpatialColorSmooth_add_SelfContain(tf.constant(P), tf.constant(X[:, :, 0:6]))
    gidx = rng.integers(0, 200, (2, 200, 7)).astype(np.int32)
    bg = Tool.batch_gather_v1(tf.constant(P), tf.constant(gidx))
    np.savez_compressed(os.path.join(HERE, "ref_unit_ops.npz"), X=X, adj=_np(adj), knn=_np(idx).astype(np.int16), edge_feature=_np(ef)[:, ::8].copy(),
                        adj_batch1=_np(adj1), P=P, smooth_loss=_np(sm), smooth_idx=tf.RECORD["top_k"][-1].astype(np.int16),
                        gather_idx=gidx, gathered=_np(bg))
    print("ref_unit_ops: smooth", float(_np(sm)))


def label_propagation():
    """Tool.TF_Computation.LaplacianMatSym_XYZRGB_DirectComp + ProbLabelPropagation.LabelPropagation_TF
    (the test-time stage, S3DIS_DGCNN_trainer.py:139-143 and Test())."""
    import Tool
    import ProbLabelPropagation as PLP

    rng = np.random.default_rng(331)
    N, K = 160, 13
    xyz = (np.concatenate([rng.uniform(-0.5, 0.5, (1, N, 2)), rng.uniform(0, 3, (1, N, 1))], -1) * 0.15).astype(np.float32)
    rgb = rng.uniform(0, 1, (1, N, 3)).astype(np.float32)
    G = rng.dirichlet(np.ones(K) * 0.3, N).astype(np.float32)
    sess = tf.Session()
    tf.reset()
    tf.feed_queue([xyz, rgb])
    comp = Tool.TF_Computation.LaplacianMatSym_XYZRGB_DirectComp()
    L = comp.Eval(sess, xyz, rgb)
    tf.feed_queue([G, L[0], np.float32(1.0), np.float32(1.0)])
    solver = PLP.LabelPropagation_TF(alpha=1e0, beta=1e0, K=10)
    Yv, Yp, w = solver.SolveLabelProp(sess, L[0], G)
    np.savez_compressed(os.path.join(HERE, "ref_label_prop.npz"), xyz=xyz, rgb=rgb, G=G, L=L, Y=Yv, Y_prob=Yp, w=w)
    print("ref_label_prop: |L|", float(np.abs(L).max()), "argmax agreement with G", float((Yp.argmax(-1) == G.argmax(-1)).mean()))


if __name__ == "__main__":
    unit_ops()
    label_propagation()
    s3dis_train_step()
    shapenet_train_step()
