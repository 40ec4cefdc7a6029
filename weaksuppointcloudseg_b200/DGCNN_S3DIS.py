"""Drop-in for S3DIS/DGCNN_S3DIS.py (reference :16-192): S3DIS segmentation DGCNN constructor.

`get_model(point_cloud, is_training, weight_decay=0., bn_decay=None)` keeps the reference signature and returns
the logits (B,N,13) of the fused CUDA executor (engine_s3dis.S3DISEngine), which runs the exact layer sequence of
reference :24-104.  Engines are cached per (B, N, variant); `set_variables` / `get_engine` expose the variables
under their TF names.  `get_model_unfused` builds the same network op by op from tf_util (the reference's own
structure) and is what the API-level parity tests compare against."""
from __future__ import annotations

import torch

from . import _lib as L
from . import tf_util
from .engine_s3dis import LAYERS, S3DISEngine

_ENGINES: dict = {}
_PARAMS = None


def placeholder_inputs(batch_size, num_point):
    """(:16-21) placeholders become pre-allocated CUDA tensors"""
    dev = torch.device("cuda", torch.cuda.current_device())
    return (torch.zeros((batch_size, num_point, 9), dtype=torch.float32, device=dev),
            torch.zeros((batch_size, num_point), dtype=torch.int32, device=dev))


def set_variables(params):
    """variables (TF names -> numpy) used by engines created afterwards"""
    global _PARAMS
    _PARAMS = params
    _ENGINES.clear()


def get_engine(B, N, device, unnorm=False):
    key = (B, N, str(device), unnorm)
    if key not in _ENGINES:
        from .S3DIS_DGCNN_trainer import xavier_params
        _ENGINES[key] = S3DISEngine(_PARAMS if _PARAMS is not None else xavier_params(LAYERS, 0), B, N, device=device,
                                    unnorm_xyz=unnorm)
    return _ENGINES[key]


def get_model(point_cloud, is_training, weight_decay=0., bn_decay=None):
    """ ConvNet baseline, input is BxNx9 (:24-104); kNN on the normalised xyz channels 6:9 """
    B, N, _ = point_cloud.shape
    eng = get_engine(B, N, point_cloud.device)
    return eng.forward(point_cloud.contiguous(), bool(is_training), bn_decay)


def get_model_unnormXYZ(point_cloud, is_training, weight_decay=0., bn_decay=None):
    """(:106-186) identical except that the first kNN runs on channels 0:3 (:114)"""
    B, N, _ = point_cloud.shape
    eng = get_engine(B, N, point_cloud.device, unnorm=True)
    return eng.forward(point_cloud.contiguous(), bool(is_training), bn_decay)


def get_model_unfused(point_cloud, is_training, weight_decay=0., bn_decay=None):
    """the reference graph written with the unfused tf_util ops, line for line in structure (:24-104)"""
    B, N, _ = point_cloud.shape
    k = 20
    input_image = point_cloud.unsqueeze(2)
    adj = tf_util.pairwise_distance(point_cloud[:, :, 6:].contiguous())
    nn_idx = tf_util.knn(adj, k=k)
    edge_feature = tf_util.get_edge_feature(input_image, nn_idx=nn_idx, k=k)
    kw = dict(padding='VALID', stride=[1, 1], bn=True, is_training=is_training, is_dist=True)
    out1 = tf_util.conv2d(edge_feature, 64, [1, 1], scope='adj_conv1', bn_decay=bn_decay, **kw)
    out2 = tf_util.conv2d(out1, 64, [1, 1], scope='adj_conv2', bn_decay=bn_decay, **kw)
    net_1 = out2.amax(dim=-2, keepdim=True)
    adj = tf_util.pairwise_distance(net_1)
    nn_idx = tf_util.knn(adj, k=k)
    edge_feature = tf_util.get_edge_feature(net_1, nn_idx=nn_idx, k=k)
    out3 = tf_util.conv2d(edge_feature, 64, [1, 1], scope='adj_conv3', bn_decay=bn_decay, **kw)
    out4 = tf_util.conv2d(out3, 64, [1, 1], scope='adj_conv4', bn_decay=bn_decay, **kw)
    net_2 = out4.amax(dim=-2, keepdim=True)
    adj = tf_util.pairwise_distance(net_2)
    nn_idx = tf_util.knn(adj, k=k)
    edge_feature = tf_util.get_edge_feature(net_2, nn_idx=nn_idx, k=k)
    out5 = tf_util.conv2d(edge_feature, 64, [1, 1], scope='adj_conv5', bn_decay=bn_decay, **kw)
    net_3 = out5.amax(dim=-2, keepdim=True)
    out7 = tf_util.conv2d(torch.cat([net_1, net_2, net_3], dim=-1), 1024, [1, 1], scope='adj_conv7', bn_decay=bn_decay, **kw)
    out_max = tf_util.max_pool2d(out7, [N, 1], padding='VALID', scope='maxpool')
    expand = out_max.expand(B, N, 1, out_max.shape[-1])
    concat = torch.cat([expand, net_1, net_2, net_3], dim=3)
    net = tf_util.conv2d(concat, 512, [1, 1], scope='seg/conv1', **kw)
    net = tf_util.conv2d(net, 256, [1, 1], scope='seg/conv2', **kw)
    net = tf_util.dropout(net, keep_prob=0.7, is_training=is_training, scope='dp1')
    net = tf_util.conv2d(net, 13, [1, 1], padding='VALID', stride=[1, 1], activation_fn=None, scope='seg/conv3', is_dist=True)
    return net.squeeze(2)


def get_loss(pred, label):
    """ pred: B,N,13; label: B,N -> mean sparse softmax cross entropy (:189-192) """
    B, N, C = pred.shape
    dev = pred.device
    Y = torch.zeros((B, N, C), dtype=torch.float32, device=dev)
    Y.scatter_(2, label.long().unsqueeze(-1), 1.0)
    M = torch.ones((B, N), dtype=torch.float32, device=dev)
    P = torch.empty_like(pred)
    losses = torch.empty(5, dtype=torch.float32, device=dev)
    ws = L.workspace(L.lib().wspc_head_losses_workspace_bytes(B, N, C), dev, "head")
    L.check(L.lib().wspc_head_losses(L.ptr(pred.contiguous()), L.ptr(Y), L.ptr(M), None, None, B, N, C, 0, 0.1, 0.0, 0, 0,
                                     L.ptr(P), None, L.ptr(losses), L.ptr(ws), ws.numel(), L.stream()))
    return losses[0]
