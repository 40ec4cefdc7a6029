"""Drop-in for Networks/dgcnn/models/dgcnn.py (reference :14-109): the ModelNet classification DGCNN.

The weak-supervision trainers never build this graph (unused variant); it is kept for API completeness and is
composed from the unfused `tf_util` ops of this package -- the same hand-written kernels behind the C ABI (tcgen05 kNN,
edge-feature gather, 1x1-conv GEMMs, BN, pools), forward only.  Same function names, argument order and return values:
`get_model(point_cloud, is_training, bn_decay)` -> (logits (B,40), end_points), `get_loss(pred, label, end_points)`.
The reference's batch norm here is the non-distributed template whose population statistics are EMA shadow variables
(tf_util.py:462-499); they are kept under `<scope>/bn/pop_mean`, `<scope>/bn/pop_var` like everywhere else in this package.
Pinned to the reference module run on the TF shim (tests/golden/make_util_golden.py::classification_model)."""
from __future__ import annotations

import torch

from . import _lib as L
from . import tf_util
from .transform_nets import input_transform_net


def placeholder_inputs(batch_size, num_point):
    """(:14-17) the two feed tensors of the classification graph, as CUDA tensors"""
    dev = torch.device("cuda", torch.cuda.current_device())
    return (torch.zeros((batch_size, num_point, 3), dtype=torch.float32, device=dev),
            torch.zeros((batch_size,), dtype=torch.int32, device=dev))


def _edge_conv(points, k, channels, scope, is_training, bn_decay):
    """pairwise_distance -> knn -> get_edge_feature -> conv2d + BN + ReLU -> max over k, keep_dims (:26-31, :38-47)"""
    adj_matrix = tf_util.pairwise_distance(points)
    nn_idx = tf_util.knn(adj_matrix, k=k)
    edge_feature = tf_util.get_edge_feature(points, nn_idx=nn_idx, k=k)
    net = tf_util.conv2d(edge_feature, channels, [1, 1], padding='VALID', stride=[1, 1], bn=True, is_training=is_training,
                         scope=scope, bn_decay=bn_decay)
    return net.amax(dim=-2, keepdim=True)


def get_model(point_cloud, is_training, bn_decay=None):
    """ Classification DGCNN, input is BxNx3, output Bx40 """
    point_cloud = point_cloud.contiguous()
    batch_size, num_point = point_cloud.shape[0], point_cloud.shape[1]
    end_points = {}
    k = 20
    is_training = bool(is_training)

    adj_matrix = tf_util.pairwise_distance(point_cloud)
    nn_idx = tf_util.knn(adj_matrix, k=k)
    edge_feature = tf_util.get_edge_feature(point_cloud, nn_idx=nn_idx, k=k)
    with tf_util.variable_scope('transform_net1'):
        transform = input_transform_net(edge_feature, is_training, bn_decay, K=3)
    # point_cloud_transformed = tf.matmul(point_cloud, transform)   (:33)
    point_cloud_transformed = torch.empty_like(point_cloud)
    L.check(L.lib().wspc_transform_points_fwd(L.ptr(point_cloud), L.ptr(transform.contiguous()), batch_size, num_point, 0,
                                              L.ptr(point_cloud_transformed), L.stream()))

    net1 = _edge_conv(point_cloud_transformed, k, 64, 'dgcnn1', is_training, bn_decay)
    net2 = _edge_conv(net1, k, 64, 'dgcnn2', is_training, bn_decay)
    net3 = _edge_conv(net2, k, 64, 'dgcnn3', is_training, bn_decay)
    net4 = _edge_conv(net3, k, 128, 'dgcnn4', is_training, bn_decay)

    net = tf_util.conv2d(torch.cat([net1, net2, net3, net4], dim=-1), 1024, [1, 1], padding='VALID', stride=[1, 1], bn=True,
                         is_training=is_training, scope='agg', bn_decay=bn_decay)
    net = tf_util.max_pool2d(net, [num_point, 1], padding='VALID', scope='maxpool')          # tf.reduce_max(net, axis=1) (:88)

    # MLP on global point cloud vector
    net = net.reshape(batch_size, -1)
    net = tf_util.fully_connected(net, 512, bn=True, is_training=is_training, scope='fc1', bn_decay=bn_decay)
    net = tf_util.dropout(net, keep_prob=0.5, is_training=is_training, scope='dp1')
    net = tf_util.fully_connected(net, 256, bn=True, is_training=is_training, scope='fc2', bn_decay=bn_decay)
    net = tf_util.dropout(net, keep_prob=0.5, is_training=is_training, scope='dp2')
    net = tf_util.fully_connected(net, 40, activation_fn=None, scope='fc3')
    return net, end_points


def get_loss(pred, label, end_points):
    """ pred: B*NUM_CLASSES, label: B -- softmax cross entropy with label smoothing 0.2, mean over the batch (:100-106) """
    num_classes = pred.shape[-1]
    labels = torch.nn.functional.one_hot(label.long(), num_classes).to(pred.dtype) * (1.0 - 0.2) + 0.2 / num_classes
    loss = -(labels * torch.log_softmax(pred, dim=-1)).sum(-1)
    return loss.mean()
