"""CPU oracle for the DGCNN segmentation networks, weak-supervision losses and TF-style optimiser.

TEST INFRASTRUCTURE ONLY — see oracle/__init__.py for the pinning status.  The reference has no tests or
golden vectors for this path and TensorFlow 1.14 cannot be installed here, so this module restates
the reference's graph construction op by op (file:line cited on every function) with the TF op
semantics of SURVEY.md App. A, on torch-CPU tensors (fp32 by default, fp64 for gradient checks); it is
held to the outputs of the reference's own Python run on the op-level TF stand-in
(tests/golden/ref_*.npz, tests/test_reference_golden_cpu.py).
Distances / kNN go through oracle/knn_oracle.c (canonical fp32 arithmetic); everything else uses
torch CPU ops with autograd providing the reference gradients.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch

from . import knn as oknn

BN_EPS = 1e-3  # tf_util.py:527,530


# ------------------------------------------------------------------------------------------------
# parameters (TF variable names: <scope>/weights, <scope>/biases, <scope>/bn/{beta,gamma,pop_mean,pop_var})
# ------------------------------------------------------------------------------------------------
S3DIS_LAYERS = [  # (scope, Cin, Cout, has_bn)          DGCNN_S3DIS.py:36-101
    ("adj_conv1", 18, 64, True), ("adj_conv2", 64, 64, True), ("adj_conv3", 128, 64, True),
    ("adj_conv4", 64, 64, True), ("adj_conv5", 128, 64, True), ("adj_conv7", 192, 1024, True),
    ("seg/conv1", 1216, 512, True), ("seg/conv2", 512, 256, True), ("seg/conv3", 256, 13, False),
]
SHAPENET_LAYERS = [  # DGCNN_ShapeNet.py:27-109, transform_nets.py:18-40
    ("transform_net1/tconv1", 6, 64, True), ("transform_net1/tconv2", 64, 128, True),
    ("transform_net1/tconv3", 128, 1024, True), ("transform_net1/tfc1", 1024, 512, True),
    ("transform_net1/tfc2", 512, 256, True),
    ("adj_conv1", 6, 64, True), ("adj_conv2", 64, 64, True), ("adj_conv3", 128, 64, True),
    ("adj_conv4", 64, 64, True), ("adj_conv5", 128, 64, True), ("adj_conv7", 192, 1024, True),
    ("one_hot_label_expand", 16, 64, True),
    ("seg/conv1", 1280, 256, True), ("seg/conv2", 256, 256, True), ("seg/conv3", 256, 128, True),
    ("seg/conv4", 128, 50, False),
]


def init_params(layers, seed=1234, shapenet=False) -> "OrderedDict[str, np.ndarray]":
    """Xavier-uniform weights U(+-sqrt(6/(Cin+Cout))), zero biases, gamma=1, beta=0, pop_mean=0, pop_var=1
    (tf_util.py:43-47,160-161,513-519; tf.contrib.layers.xavier_initializer [TF]).  numpy so that
    the oracle and the CUDA path load bit-identical values."""
    rng = np.random.default_rng(seed)
    p = OrderedDict()
    for scope, cin, cout, has_bn in layers:
        lim = math.sqrt(6.0 / (cin + cout))
        p[f"{scope}/weights"] = rng.uniform(-lim, lim, (cin, cout)).astype(np.float32)
        p[f"{scope}/biases"] = np.zeros((cout,), np.float32)
        if has_bn:
            p[f"{scope}/bn/beta"] = np.zeros((cout,), np.float32)
            p[f"{scope}/bn/gamma"] = np.ones((cout,), np.float32)
            p[f"{scope}/bn/pop_mean"] = np.zeros((cout,), np.float32)
            p[f"{scope}/bn/pop_var"] = np.ones((cout,), np.float32)
    if shapenet:  # transform_nets.py:45-51: W = 0, b = 0 (+ eye added in the graph)
        p["transform_net1/transform_XYZ/weights"] = np.zeros((256, 9), np.float32)
        p["transform_net1/transform_XYZ/biases"] = np.zeros((9,), np.float32)
    return p


def trainable_names(params) -> list:
    return [k for k in params if not (k.endswith("pop_mean") or k.endswith("pop_var"))]


def to_torch(params, dtype=torch.float32, requires_grad=True):
    out = OrderedDict()
    tn = set(trainable_names(params))
    for k, v in params.items():
        t = torch.tensor(np.asarray(v), dtype=dtype)
        if requires_grad and k in tn:
            t.requires_grad_(True)
        out[k] = t
    return out


# ------------------------------------------------------------------------------------------------
# ops  (Networks/dgcnn/utils/tf_util.py)
# ------------------------------------------------------------------------------------------------
def knn_indices(x: torch.Tensor, k: int, flavour=oknn.TFUTIL, return_dist=False):
    """tf_util.pairwise_distance + knn (tf_util.py:638-671); no gradient (indices only, :670)."""
    xn = x.detach().to(torch.float32).numpy()
    r = oknn.knn(xn, k, flavour, return_dist=return_dist)
    if return_dist:
        return torch.from_numpy(r[0].astype(np.int64)), torch.from_numpy(r[1])
    return torch.from_numpy(r.astype(np.int64))


def get_edge_feature(x: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """tf_util.get_edge_feature (tf_util.py:674-706): concat([x_i tiled, x_j - x_i], -1) -> (B,N,k,2C)."""
    B, N, C = x.shape
    k = idx.shape[-1]
    flat = x.reshape(B * N, C)
    gidx = (idx + (torch.arange(B).view(B, 1, 1) * N)).reshape(-1)       # :696-700
    nbr = flat[gidx].view(B, N, k, C)
    ctr = x.unsqueeze(2).expand(B, N, k, C)                               # :701-703
    return torch.cat([ctr, nbr - ctr], dim=-1)                            # :705


def batch_norm(y, p, scope, is_training, bn_decay, axes):
    """batch_norm_dist_template (tf_util.py:502-535): biased batch variance, eps 1e-3, population
    statistics updated as pop*decay + batch*(1-decay) (decay = bn_decay or 0.9) in training."""
    gamma, beta = p[f"{scope}/bn/gamma"], p[f"{scope}/bn/beta"]
    if is_training:
        mean = y.mean(dim=axes)
        var = ((y - mean) ** 2).mean(dim=axes)                            # tf.nn.moments [TF]
        decay = 0.9 if bn_decay is None else bn_decay                     # :523
        with torch.no_grad():
            p[f"{scope}/bn/pop_mean"].mul_(decay).add_(mean.detach() * (1 - decay))   # :524
            p[f"{scope}/bn/pop_var"].mul_(decay).add_(var.detach() * (1 - decay))     # :525
    else:
        mean, var = p[f"{scope}/bn/pop_mean"], p[f"{scope}/bn/pop_var"]   # :530
    inv = torch.rsqrt(var + BN_EPS) * gamma                               # tf.nn.batch_normalization [TF]
    return y * inv + (beta - mean * inv)


# Forced routing (gradient parity tests): the graph is piecewise linear in its ReLUs / max pools, and two correct fp32
# implementations may take different branches where a value sits within rounding of a threshold.  With ROUTE set to a dict
# of decisions exported from the implementation under test (tests/routing.py) every branch is taken as THAT implementation
# took it, so the remaining difference is arithmetic only and a 1e-3 bound on every gradient tensor means something.
#   "relu/<scope>"  0/1 mask shaped like the layer output         "maxk/<knn tag>"  (B,N,k,C) weights of the max over k
#   "maxn/<scope>"  ((B,C) arg-max rows, (B,C) 0/1 ReLU gate)     "inexact"         (B,N,C) weights of the max over points
ROUTE = None


class forced_routing:
    def __init__(self, route):
        self.route = route

    def __enter__(self):
        global ROUTE
        self.prev, ROUTE = ROUTE, self.route
        return self.route

    def __exit__(self, *exc):
        global ROUTE
        ROUTE = self.prev
        return False


def conv2d(x, p, scope, is_training, bn=True, bn_decay=None, act=True, rec=None):
    """tf_util.conv2d with a [1,1] kernel (tf_util.py:115-173): x W + b -> BN -> ReLU over the last axis."""
    y = x @ p[f"{scope}/weights"] + p[f"{scope}/biases"]                  # :160-165
    if rec is not None:
        rec[f"{scope}/pre"] = y
    if bn:
        y = batch_norm(y, p, scope, is_training, bn_decay, tuple(range(y.dim() - 1)))  # :167-169, axes [0,1,2]
    if act:
        if ROUTE is not None and f"relu/{scope}" in ROUTE:
            y = y * ROUTE[f"relu/{scope}"].to(y.dtype).reshape(y.shape)
        else:
            y = torch.relu(y)                                             # :171-172
    return y


fully_connected = conv2d  # tf_util.fully_connected (tf_util.py:317-354): same arithmetic, BN axes [0]


class _MaxPoolN(torch.autograd.Function):
    """tf_util.max_pool2d([N,1]) (tf_util.py:357-380): max over the point axis; gradient goes to the
    FIRST arg-max row [TF MaxPoolGrad], unlike reduce_max's equal split."""

    @staticmethod
    def forward(ctx, x):  # (B,N,C)
        m = x.max(dim=1, keepdim=True).values
        first = (x == m).to(torch.uint8).argmax(dim=1, keepdim=True)     # argmax of 0/1 -> first max index
        ctx.save_for_backward(first)
        ctx.shape = x.shape
        return m.squeeze(1)

    @staticmethod
    def backward(ctx, g):
        (first,) = ctx.saved_tensors
        gx = torch.zeros(ctx.shape, dtype=g.dtype)
        gx.scatter_(1, first, g.unsqueeze(1))
        return gx


def max_pool_points(x):
    return _MaxPoolN.apply(x)


def reduce_max_k(x):
    """tf.reduce_max(axis=-2) (DGCNN_S3DIS.py:46): gradient split equally among ties [TF _MinOrMaxGrad];
    torch.amax has the same rule."""
    return torch.amax(x, dim=-2)


def dropout(x, is_training, keep_prob, mask=None):
    """tf_util.dropout (tf_util.py:614-635) -> tf.nn.dropout: x * mask / keep, mask = floor(keep + U[0,1)) [TF].
    `mask` (0/1, same shape) is injectable so both sides of a parity test use the same draw."""
    if not is_training:
        return x
    if mask is None:
        mask = torch.floor(keep_prob + torch.rand_like(x))
    return x * mask.to(x.dtype) / keep_prob


# ------------------------------------------------------------------------------------------------
# models
# ------------------------------------------------------------------------------------------------
def _edge_block(x_feat, knn_src, p, scopes, is_training, bn_decay, k, rec, tag, knn_override):
    """kNN -> edge feature -> conv2d x len(scopes) -> max over k  (DGCNN_S3DIS.py:32-46 and repeats)."""
    idx = knn_override.get(tag) if knn_override else None
    if idx is None:
        idx = knn_indices(knn_src, k)
    if rec is not None:
        rec[f"{tag}/idx"] = idx
    net = get_edge_feature(x_feat, idx)
    forced = ROUTE.get(f"maxk/{tag}") if ROUTE is not None else None
    for s in scopes:
        last = s == scopes[-1]
        net = conv2d(net, p, s, is_training, bn=True, bn_decay=bn_decay, rec=rec, act=not (last and forced is not None))
    if forced is not None:      # ReLU + max over k as one routed selection (weights are zero where the ReLU clipped the maximum)
        return (net * forced.to(net.dtype)).sum(dim=-2)
    return reduce_max_k(net)


def get_model_s3dis(p, point_cloud, is_training, bn_decay=None, k=20, dropout_mask=None, rec=None,
                    knn_override=None, unnorm_xyz=False):
    """DGCNN_S3DIS.get_model (S3DIS/DGCNN_S3DIS.py:24-104); unnorm_xyz=True -> get_model_unnormXYZ (:106-186)."""
    B, N, _ = point_cloud.shape
    src = point_cloud[:, :, 0:3] if unnorm_xyz else point_cloud[:, :, 6:9]            # :32 / :114
    net_1 = _edge_block(point_cloud, src, p, ["adj_conv1", "adj_conv2"], is_training, bn_decay, k, rec, "knn1",
                        knn_override)                                                  # :32-46
    net_2 = _edge_block(net_1, net_1, p, ["adj_conv3", "adj_conv4"], is_training, bn_decay, k, rec, "knn2",
                        knn_override)                                                  # :48-62
    net_3 = _edge_block(net_2, net_2, p, ["adj_conv5"], is_training, bn_decay, k, rec, "knn3", knn_override)  # :64-78
    cat = torch.cat([net_1, net_2, net_3], dim=-1)
    mn = ROUTE.get("maxn/adj_conv7") if ROUTE is not None else None
    out7 = conv2d(cat, p, "adj_conv7", is_training, bn=True, bn_decay=bn_decay, rec=rec, act=mn is None)   # :80-83
    if mn is None:
        out_max = max_pool_points(out7)                                                # :85
    else:
        out_max = torch.gather(out7, 1, mn[0].long().unsqueeze(1)).squeeze(1) * mn[1].to(out7.dtype)
    expand = out_max.unsqueeze(1).expand(B, N, out_max.shape[-1])                      # :87
    concat = torch.cat([expand, net_1, net_2, net_3], dim=-1)                          # :89-92
    net = conv2d(concat, p, "seg/conv1", is_training, bn=True, bn_decay=None, rec=rec)  # :95-96 (decay 0.9)
    net = conv2d(net, p, "seg/conv2", is_training, bn=True, bn_decay=None, rec=rec)     # :97-98
    net = dropout(net, is_training, 0.7, dropout_mask)                                 # :99
    net = conv2d(net, p, "seg/conv3", is_training, bn=False, act=False, rec=rec)        # :100-101
    if rec is not None:
        rec.update(net_1=net_1, net_2=net_2, net_3=net_3, out_max=out_max)
    return net


def input_transform_net(p, edge_feature, is_training, bn_decay, rec=None, prefix="transform_net1/"):
    """transform_nets.input_transform_net (Networks/dgcnn/models/transform_nets.py:10-56), K=3."""
    B = edge_feature.shape[0]
    net = conv2d(edge_feature, p, prefix + "tconv1", is_training, bn_decay=bn_decay, rec=rec)   # :18-21
    forced = ROUTE.get("maxk/tnet") if ROUTE is not None else None
    net = conv2d(net, p, prefix + "tconv2", is_training, bn_decay=bn_decay, rec=rec, act=forced is None)   # :22-25
    if forced is not None:      # ReLU + max over k as one routed selection, as in _edge_block
        net = (net * forced.to(net.dtype)).sum(dim=-2)
    else:
        net = reduce_max_k(net)                                                                  # :27
    mn = ROUTE.get("maxn/" + prefix + "tconv3") if ROUTE is not None else None
    net = conv2d(net, p, prefix + "tconv3", is_training, bn_decay=bn_decay, rec=rec, act=mn is None)       # :29-32
    if mn is None:
        net = max_pool_points(net)                                                               # :33-36
    else:
        net = torch.gather(net, 1, mn[0].long().unsqueeze(1)).squeeze(1) * mn[1].to(net.dtype)
    net = fully_connected(net, p, prefix + "tfc1", is_training, bn_decay=bn_decay, rec=rec)     # :37-38
    net = fully_connected(net, p, prefix + "tfc2", is_training, bn_decay=bn_decay, rec=rec)     # :39-40
    W, b = p[prefix + "transform_XYZ/weights"], p[prefix + "transform_XYZ/biases"]
    t = net @ W + (b + torch.eye(3, dtype=net.dtype).flatten())                                  # :51-53
    return t.view(B, 3, 3)


def get_model_shapenet(p, point_cloud, input_label, is_training, bn_decay=None, k=20, dropout_masks=None,
                       rec=None, knn_override=None):
    """DGCNN_ShapeNet.get_model (ShapeNet/DGCNN_ShapeNet.py:15-113). input_label: (B,16) one-hot float."""
    B, N, _ = point_cloud.shape
    ov = knn_override or {}
    idx0 = ov.get("knn0")
    if idx0 is None:
        idx0 = knn_indices(point_cloud, k)                                                       # :23-24
    if rec is not None:
        rec["knn0/idx"] = idx0
    T = input_transform_net(p, get_edge_feature(point_cloud, idx0), is_training, bn_decay, rec)  # :25-28
    pct = point_cloud @ T                                                                        # :29
    if rec is not None:
        rec["transform"] = T
        rec["pct"] = pct
    net_1 = _edge_block(pct, pct, p, ["adj_conv1", "adj_conv2"], is_training, bn_decay, k, rec, "knn1", ov)
    net_2 = _edge_block(net_1, net_1, p, ["adj_conv3", "adj_conv4"], is_training, bn_decay, k, rec, "knn2", ov)
    net_3 = _edge_block(net_2, net_2, p, ["adj_conv5"], is_training, bn_decay, k, rec, "knn3", ov)
    cat = torch.cat([net_1, net_2, net_3], dim=-1)
    mn = ROUTE.get("maxn/adj_conv7") if ROUTE is not None else None
    out7 = conv2d(cat, p, "adj_conv7", is_training, bn_decay=bn_decay, rec=rec, act=mn is None)  # :80-83
    if mn is None:
        out_max = max_pool_points(out7)                                                          # :85
    else:
        out_max = torch.gather(out7, 1, mn[0].long().unsqueeze(1)).squeeze(1) * mn[1].to(out7.dtype)
    lab = conv2d(input_label.to(out_max.dtype), p, "one_hot_label_expand", is_training, bn_decay=bn_decay,
                 rec=rec)                                                                        # :87-91
    g = torch.cat([out_max, lab], dim=-1)                                                        # :92
    expand = g.unsqueeze(1).expand(B, N, g.shape[-1])                                            # :93
    concat = torch.cat([expand, net_1, net_2, net_3], dim=-1)                                    # :95-98
    dm = dropout_masks or (None, None)
    net = conv2d(concat, p, "seg/conv1", is_training, bn_decay=bn_decay, rec=rec)                # :100-101
    net = dropout(net, is_training, 0.6, dm[0])                                                  # :102
    net = conv2d(net, p, "seg/conv2", is_training, bn_decay=bn_decay, rec=rec)                   # :103-104
    net = dropout(net, is_training, 0.6, dm[1])                                                  # :105
    net = conv2d(net, p, "seg/conv3", is_training, bn_decay=bn_decay, rec=rec)                   # :106-107
    net = conv2d(net, p, "seg/conv4", is_training, bn=False, act=False, rec=rec)                 # :108-109
    if rec is not None:
        rec.update(net_1=net_1, net_2=net_2, net_3=net_3, out_max=out_max)
    return net


# ------------------------------------------------------------------------------------------------
# losses  (trainers + Util/SmoothConstraint.py)
# ------------------------------------------------------------------------------------------------
def seg_loss(Z, Y_onehot, Mask):
    """S3DIS_DGCNN_trainer.py:89-90 / ShapeNet_DGCNN_trainer.py:88-89:
    sum(Mask * softmax_cross_entropy_with_logits(Y, Z)) / sum(Mask)."""
    ce = -(Y_onehot.to(Z.dtype) * torch.log_softmax(Z, dim=-1)).sum(-1)
    return (Mask * ce).sum() / Mask.sum()


def siamese_loss(Z_prob, weight):
    """S3DIS_DGCNN_trainer.py:128 (weight 1e1) / ShapeNet_DGCNN_trainer.py:123-124 (weight 1):
    mean over (pair, point) of sum_c (P[0::2] - P[1::2])^2."""
    return weight * ((Z_prob[0::2] - Z_prob[1::2]) ** 2).sum(-1).mean()


def inexact_loss(Z, Y_onehot):
    """S3DIS_DGCNN_trainer.py:131-134: L_gt = max_n Y; L = max_n Z; mean sigmoid_cross_entropy_with_logits
    = max(x,0) - x*z + log(1 + exp(-|x|)) [TF]; reduce_max gradient: equal split among ties [TF]."""
    L_gt = Y_onehot.to(Z.dtype).amax(dim=1)
    if ROUTE is not None and "inexact" in ROUTE:
        L = (Z * ROUTE["inexact"].to(Z.dtype)).sum(dim=1)
    else:
        L = torch.amax(Z, dim=1)
    l = torch.clamp(L, min=0) - L * L_gt + torch.log1p(torch.exp(-L.abs()))
    return l.mean()


def smooth_graph(X, gamma=1e-1, knn=10):
    """kNN graph + weights of Loss_SpatialColorSmooth_add_SelfContain (Util/SmoothConstraint.py:141-158):
    clamped sq. distances on X (B,N,3|6), top_k(-D, knn), W = exp(-D/gamma) gathered at the indices."""
    idx, d = knn_indices(X, knn, flavour=oknn.SMOOTH, return_dist=True)
    g = np.float32(gamma)
    W = torch.from_numpy(np.exp((-d.numpy()) / g).astype(np.float32))                   # :157 exp(-d_i / gamma)
    return idx, W


def smooth_loss(Z_prob, X, gamma=1e-1, knn=10, graph=None):
    """Loss_SpatialColorSmooth_add_SelfContain (Util/SmoothConstraint.py:130-167):
    mean_{b,n,j}( W * mean_c (Z_i - Z_j)^2 ); gradient w.r.t. Z only (X is an input)."""
    idx, W = graph if graph is not None else smooth_graph(X, gamma, knn)
    knn = idx.shape[-1]
    B, N, C = Z_prob.shape
    gidx = (idx + (torch.arange(B).view(B, 1, 1) * N)).reshape(-1)
    Zt = Z_prob.reshape(B * N, C)[gidx].view(B, N, knn, C)                               # batch_gather_v1, Tool.py:72-104
    Ze = Z_prob.unsqueeze(2)                                                             # :160
    loss = W.to(Z_prob.dtype) * ((Ze - Zt) ** 2).mean(-1)                                # :161
    return loss.mean()                                                                   # :163


def weak_sup_losses(Z, X_smooth, Y_onehot, Mask, siamese_weight, smooth_graph_=None):
    """defineNetwork loss block + WeakSupLoss (S3DIS_DGCNN_trainer.py:85-102,120-137) with the ramp-up
    gate = 1.  Returns dict of the four terms and their sum."""
    Zp = torch.softmax(Z, dim=-1)
    out = dict(
        loss_seg=seg_loss(Z, Y_onehot, Mask),
        loss_siamese=siamese_loss(Zp, siamese_weight),
        loss_inexact=inexact_loss(Z, Y_onehot),
        loss_smooth=smooth_loss(Zp, X_smooth, graph=smooth_graph_),
        Z_prob=Zp,
    )
    out["loss"] = out["loss_seg"] + out["loss_siamese"] + out["loss_inexact"] + out["loss_smooth"]
    return out


# ------------------------------------------------------------------------------------------------
# optimiser + schedules
# ------------------------------------------------------------------------------------------------
def learning_rate(step, base_lr, batch_size, decay_step, decay_rate=0.5):
    """get_learning_rate (S3DIS_DGCNN_trainer.py:36-44): staircase exponential decay, clipped at 1e-5."""
    return max(base_lr * decay_rate ** math.floor(step * batch_size / decay_step), 1e-5)


def bn_decay(step, batch_size, decay_step, init=0.5, rate=0.5, clip=0.99):
    """get_bn_decay (S3DIS_DGCNN_trainer.py:46-54): min(clip, 1 - init*rate^floor(step*bs/(2*DECAY_STEP)))."""
    return min(clip, 1 - init * rate ** math.floor(step * batch_size / float(decay_step * 2)))


class AdamTF:
    """tf.train.AdamOptimizer [TF]: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); theta -= lr_t*m/(sqrt(v)+eps)
    (eps NOT bias-corrected) — SURVEY App. A-12."""

    def __init__(self, params, names, b1=0.9, b2=0.999, eps=1e-8):
        self.p, self.names, self.b1, self.b2, self.eps, self.t = params, list(names), b1, b2, eps, 0
        self.m = {n: torch.zeros_like(params[n]) for n in self.names}
        self.v = {n: torch.zeros_like(params[n]) for n in self.names}

    @torch.no_grad()
    def step(self, grads, lr):
        self.t += 1
        lr_t = lr * math.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
        for n in self.names:
            g = grads[n]
            if g is None:
                g = torch.zeros_like(self.p[n])
            self.m[n].mul_(self.b1).add_(g, alpha=1 - self.b1)
            self.v[n].mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
            self.p[n].sub_(lr_t * self.m[n] / (self.v[n].sqrt() + self.eps))


def train_step_s3dis(p, opt, X, Y_onehot, Mask, step=0, base_lr=1e-3, batch_size=None, dropout_mask=None,
                     full=True, rec=None, knn_override=None, smooth_graph_=None, k=20):
    """One `sess.run([solver, loss, ...])` of TrainOneEpoch_Full (S3DIS_DGCNN_trainer.py:317-323)."""
    bs = batch_size if batch_size is not None else X.shape[0] // 2
    decay = bn_decay(step, bs, 300000)
    lr = learning_rate(step, base_lr, bs, 300000)
    Z = get_model_s3dis(p, X, True, bn_decay=decay, k=k, dropout_mask=dropout_mask, rec=rec, knn_override=knn_override)
    if full:
        L = weak_sup_losses(Z, X[:, :, 0:6], Y_onehot, Mask, 10.0, smooth_graph_)
    else:
        L = dict(loss_seg=seg_loss(Z, Y_onehot, Mask), Z_prob=torch.softmax(Z, -1))
        L["loss"] = L["loss_seg"]
    names = opt.names
    grads = torch.autograd.grad(L["loss"], [p[n] for n in names], allow_unused=True)
    gd = dict(zip(names, grads))
    opt.step(gd, lr)
    L["Z"] = Z
    L["grads"] = gd
    return L


def train_step_shapenet(p, opt, X, label, Y_onehot, Mask, step=0, base_lr=1e-3, batch_size=None, dropout_masks=None,
                        full=True, rec=None, knn_override=None, smooth_graph_=None, k=20):
    """One `sess.run([solver, loss, ...])` of ShapeNet TrainOneEpoch_Full (ShapeNet_DGCNN_trainer.py:308-314);
    DECAY_STEP = 16881*20 (:31), Siamese weight 1 (:123-124), smooth term on xyz (:133)."""
    bs = batch_size if batch_size is not None else X.shape[0] // 2
    decay = bn_decay(step, bs, 16881 * 20)
    lr = learning_rate(step, base_lr, bs, 16881 * 20)
    Z = get_model_shapenet(p, X, label, True, bn_decay=decay, k=k, dropout_masks=dropout_masks, rec=rec,
                           knn_override=knn_override)
    if full:
        L = weak_sup_losses(Z, X, Y_onehot, Mask, 1.0, smooth_graph_)
    else:
        L = dict(loss_seg=seg_loss(Z, Y_onehot, Mask), Z_prob=torch.softmax(Z, -1))
        L["loss"] = L["loss_seg"]
    names = opt.names
    grads = torch.autograd.grad(L["loss"], [p[n] for n in names], allow_unused=True)
    gd = dict(zip(names, grads))
    opt.step(gd, lr)
    L["Z"] = Z
    L["grads"] = gd
    return L
