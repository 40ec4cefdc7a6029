"""Drop-in for Networks/dgcnn/models/transform_nets.py (reference :10-56): the input (XYZ) transform net.

The fused implementation lives in engine_shapenet.ShapeNetEngine (T-net section of forward/backward).  This
module keeps the reference entry point, built from the unfused tf_util ops."""
from __future__ import annotations

import numpy as np
import torch

from . import tf_util


def input_transform_net(edge_feature, is_training, bn_decay=None, K=3, is_dist=False):
    """ Input (XYZ) Transform Net, input is BxNxkx6 edge feature. Return: transformation matrix of size BxKxK """
    B, N = edge_feature.shape[0], edge_feature.shape[1]
    kw = dict(padding='VALID', stride=[1, 1], bn=True, is_training=is_training, bn_decay=bn_decay, is_dist=is_dist)
    net = tf_util.conv2d(edge_feature, 64, [1, 1], scope='tconv1', **kw)
    net = tf_util.conv2d(net, 128, [1, 1], scope='tconv2', **kw)
    net = net.amax(dim=-2, keepdim=True)
    net = tf_util.conv2d(net, 1024, [1, 1], scope='tconv3', **kw)
    net = tf_util.max_pool2d(net, [N, 1], padding='VALID', scope='tmaxpool')
    net = net.reshape(B, -1)
    net = tf_util.fully_connected(net, 512, bn=True, is_training=is_training, scope='tfc1', bn_decay=bn_decay, is_dist=is_dist)
    net = tf_util.fully_connected(net, 256, bn=True, is_training=is_training, scope='tfc2', bn_decay=bn_decay, is_dist=is_dist)
    with tf_util.variable_scope('transform_XYZ'):
        dev = net.device
        W = tf_util._variable_on_cpu("/".join(tf_util._SCOPE + ["weights"]), (256, K * K), lambda s: np.zeros(s, np.float32), dev)
        b = tf_util._variable_on_cpu("/".join(tf_util._SCOPE + ["biases"]), (K * K,), lambda s: np.zeros(s, np.float32), dev)
    with tf_util.variable_scope('transform_XYZ_apply'):
        tf_util.VARIABLES["/".join(tf_util._SCOPE + ["fc", "weights"])] = W
        tf_util.VARIABLES["/".join(tf_util._SCOPE + ["fc", "biases"])] = b + torch.eye(K, device=dev).flatten()
        transform = tf_util.fully_connected(net, K * K, scope='fc', activation_fn=None)
    return transform.view(B, K, K)
