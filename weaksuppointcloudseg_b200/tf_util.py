"""Drop-in for the used subset of Networks/dgcnn/utils/tf_util.py (reference :115-173, :317-380, :502-706).

Same function names, positional order and defaults.  TF's graph/variable-scope protocol collapses to eager
CUDA ops: `scope` strings key a process-wide variable store so variable names stay TF compatible
(`<scope>/weights`, `<scope>/biases`, `<scope>/bn/{beta,gamma,pop_mean,pop_var}`); `is_training` /
`bn_decay` are plain Python values.  Every op runs in the hand-written kernels behind the C ABI.  These
unfused ops materialise their outputs (like the reference); the trainers use the fused engines instead, and
the parity tests check both against the same oracle.  Gradients are provided by the engines, not here.
"""
from __future__ import annotations

import ctypes
import math
from contextlib import contextmanager

import numpy as np
import torch

from . import _lib as L
from . import ops

VARIABLES: dict = {}       # TF-style name -> CUDA tensor
_SCOPE: list = []
_RNG = np.random.default_rng(0)
relu = "relu"              # stands in for tf.nn.relu as the activation_fn default


@contextmanager
def variable_scope(name):
    _SCOPE.append(name)
    try:
        yield "/".join(_SCOPE)
    finally:
        _SCOPE.pop()


def _full(scope, name):
    return "/".join(_SCOPE + [scope, name])


def _device(t=None):
    return t.device if (t is not None and t.is_cuda) else torch.device("cuda", torch.cuda.current_device())


def _variable_on_cpu(name, shape, initializer, dev):
    """tf_util._variable_on_cpu (:12-24): the reference pins variables to the CPU; here they live in HBM."""
    if name not in VARIABLES:
        VARIABLES[name] = torch.as_tensor(initializer(shape), dtype=torch.float32).to(dev)
    return VARIABLES[name]


def _variable_with_weight_decay(name, shape, stddev, wd, use_xavier, dev):
    """tf_util._variable_with_weight_decay (:26-51); wd only ever fed a 'losses' collection nobody reads (App. A-13)."""
    fan_in, fan_out = int(np.prod(shape[:-1])), shape[-1]

    def init(s):
        if use_xavier:
            lim = math.sqrt(6.0 / (fan_in + fan_out))
            return _RNG.uniform(-lim, lim, s).astype(np.float32)
        return np.clip(_RNG.normal(0, stddev, s), -2 * stddev, 2 * stddev).astype(np.float32)
    return _variable_on_cpu(name, shape, init, dev)


def pairwise_distance(point_cloud):
    """(batch_size, num_points, num_dims) -> (batch_size, num_points, num_points)   (:638-657)"""
    return ops.pairwise_distance(point_cloud.contiguous())


def knn(adj_matrix, k=20):
    """(batch_size, num_points, num_points) -> nearest neighbours (batch_size, num_points, k)   (:660-671)"""
    return ops.topk_rows(adj_matrix.contiguous(), k)


def get_edge_feature(point_cloud, nn_idx, k=20):
    """(B,N,1,C) + (B,N,k) -> (B,N,k,2C)   (:674-706)"""
    return ops.get_edge_feature(point_cloud, nn_idx, k)


def _bn(outputs2d, is_training, bn_decay, scope_name, act):
    rows, C = outputs2d.shape
    dev = outputs2d.device
    z, o = (lambda s: np.zeros(s, np.float32)), (lambda s: np.ones(s, np.float32))
    beta = _variable_on_cpu(scope_name + "/beta", (C,), z, dev)
    gamma = _variable_on_cpu(scope_name + "/gamma", (C,), o, dev)
    pm = _variable_on_cpu(scope_name + "/pop_mean", (C,), z, dev)
    pv = _variable_on_cpu(scope_name + "/pop_var", (C,), o, dev)
    stats = torch.zeros((2, C), dtype=torch.float64, device=dev)
    sc, sh = torch.empty(C, device=dev), torch.empty(C, device=dev)
    if is_training:   # batch statistics via a pass over the (already materialised) conv output
        ident = L.Operand(p=outputs2d.data_ptr(), ld=C, C=C)
        eye = torch.eye(C, device=dev)
        tmp = torch.empty_like(outputs2d)
        from . import runtime as rt
        rt.rows_gemm((ident, L.OP_PLAIN), eye, C, 0, rows, C, C, L.Epilogue(out=tmp.data_ptr(), ldo=C, stats=stats.data_ptr()),
                     L.EPI_STORE_STATS)
    L.check(L.lib().wspc_bn_finalize(L.ptr(stats), C, float(rows), L.ptr(gamma), L.ptr(beta), 1e-3,
                                     0.9 if bn_decay is None else float(bn_decay), 1 if is_training else 0, L.ptr(pm),
                                     L.ptr(pv), L.ptr(sc), L.ptr(sh), None, None, L.stream()))
    out = torch.empty_like(outputs2d)
    L.check(L.lib().wspc_bn_apply(L.ptr(outputs2d), L.ptr(sc), L.ptr(sh), rows, C, 1 if act else 0, L.ptr(out), L.stream()))
    return out


def batch_norm_for_conv2d(inputs, is_training, bn_decay, scope, is_dist=False):
    """(:577-592) BN over axes [0,1,2] of a BHWC tensor; both templates share the same arithmetic."""
    C = inputs.shape[-1]
    return _bn(inputs.reshape(-1, C).contiguous(), is_training, bn_decay, "/".join(_SCOPE + [scope]), False).view(inputs.shape)


def batch_norm_for_fc(inputs, is_training, bn_decay, scope, is_dist=False):
    """(:539-553) BN over axis [0]."""
    return batch_norm_for_conv2d(inputs, is_training, bn_decay, scope, is_dist)


def conv2d(inputs, num_output_channels, kernel_size, scope, stride=[1, 1], padding='SAME', use_xavier=True, stddev=1e-3,
           weight_decay=0.0, activation_fn=relu, bn=False, bn_decay=None, is_training=None, is_dist=False):
    """2D convolution with non-linear operation; only the [1,1] kernels the two models use are supported (:115-173)."""
    if list(kernel_size) != [1, 1] or list(stride) != [1, 1]:
        raise L.WspcError("wspc tf_util.conv2d implements the 1x1 convolutions of the DGCNN models only")
    from . import runtime as rt
    x = inputs.contiguous()
    Cin = x.shape[-1]
    rows = x.numel() // Cin
    dev = x.device
    W = _variable_with_weight_decay(_full(scope, "weights"), (Cin, num_output_channels), stddev, weight_decay, use_xavier, dev)
    b = _variable_on_cpu(_full(scope, "biases"), (num_output_channels,), lambda s: np.zeros(s, np.float32), dev)
    y = torch.empty((rows, num_output_channels), dtype=torch.float32, device=dev)
    rt.rows_gemm((L.Operand(p=x.data_ptr(), ld=Cin, C=Cin), L.OP_PLAIN), W, num_output_channels, 0, rows,
                 num_output_channels, Cin, L.Epilogue(out=y.data_ptr(), ldo=num_output_channels, bias=b.data_ptr()), L.EPI_STORE)
    act = activation_fn is not None
    if bn:
        y = _bn(y, bool(is_training), bn_decay, _full(scope, "bn"), act)
    elif act:
        one, zero = torch.ones(num_output_channels, device=dev), torch.zeros(num_output_channels, device=dev)
        out = torch.empty_like(y)
        L.check(L.lib().wspc_bn_apply(L.ptr(y), L.ptr(one), L.ptr(zero), rows, num_output_channels, 1, L.ptr(out), L.stream()))
        y = out
    return y.view(*inputs.shape[:-1], num_output_channels)


def fully_connected(inputs, num_outputs, scope, use_xavier=True, stddev=1e-3, weight_decay=0.0, activation_fn=relu, bn=False,
                    bn_decay=None, is_training=None, is_dist=False):
    """(:317-354) same arithmetic as a 1x1 conv on a (B, C) tensor."""
    return conv2d(inputs, num_outputs, [1, 1], scope, use_xavier=use_xavier, stddev=stddev, weight_decay=weight_decay,
                  activation_fn=activation_fn, bn=bn, bn_decay=bn_decay, is_training=is_training, is_dist=is_dist)


def max_pool2d(inputs, kernel_size, scope, stride=[2, 2], padding='VALID'):
    """(:357-380) only the [num_point, 1] global pooling of the models: (B,N,1,C) -> (B,1,1,C)."""
    B, N, one, C = inputs.shape
    if kernel_size[0] != N or kernel_size[1] != 1:
        raise L.WspcError("wspc tf_util.max_pool2d implements the [num_point,1] pooling of the DGCNN models only")
    x = inputs.contiguous()
    g = torch.empty((B, C), dtype=torch.float32, device=x.device)
    am = torch.empty((B, C), dtype=torch.int32, device=x.device)
    one_, zero_ = torch.ones(C, device=x.device), torch.zeros(C, device=x.device)
    # max_n relu(x) == max_n x for the post-ReLU activations this is applied to; a general max uses a shift
    lo = float(x.min())
    shift = torch.full((C,), -min(lo, 0.0), device=x.device)
    L.check(L.lib().wspc_maxn_bnrelu_fwd(L.ptr(x), L.ptr(one_), L.ptr(shift), B, N, C, L.ptr(g), L.ptr(am), L.stream()))
    return (g - shift).view(B, 1, 1, C)


def dropout(inputs, is_training, scope, keep_prob=0.5, noise_shape=None):
    """(:614-635) x * floor(keep + U) / keep in training, identity otherwise."""
    if not is_training:
        return inputs
    x = inputs.contiguous()
    mask = torch.empty_like(x)
    dropout.counter = getattr(dropout, "counter", 0) + 1
    L.check(L.lib().wspc_dropout_mask(L.ptr(mask), x.numel(), keep_prob, 977, dropout.counter * (x.numel() // 4 + 1), L.stream()))
    out = torch.empty_like(x)
    C = x.shape[-1]
    sc = torch.full((C,), 1.0 / keep_prob, device=x.device)
    zero = torch.zeros(C, device=x.device)
    xm = torch.empty_like(x)
    # x * mask through the bn_apply kernel (scale 1/keep) after an elementwise product on the mask operand
    L.check(L.lib().wspc_bn_apply(L.ptr(x), L.ptr(sc), L.ptr(zero), x.numel() // C, C, 0, L.ptr(xm), L.stream()))
    out = xm * mask
    return out
