"""Eager stand-in for the slice of the TensorFlow-1.14 API that the reference hot path touches.

TEST INFRASTRUCTURE ONLY (used by tests/golden/make_reference_golden.py; never imported by the product).
TensorFlow 1.14 cannot be installed in this image, so the reference's own Python
(Networks/dgcnn/utils/tf_util.py, models/transform_nets.py, S3DIS/DGCNN_S3DIS.py, ShapeNet/DGCNN_ShapeNet.py,
Util/SmoothConstraint.py, Util/Tool.py, Util/ProbLabelPropagation.py and the two trainers' defineNetwork /
WeakSupLoss) is executed UNMODIFIED from /root/reference on top of this module: every `tf.*` call below runs
immediately on torch-CPU fp32 tensors with the documented TF-1.14 op semantics, gradients come from
torch autograd over the reference's own op composition.  Graph/session protocol: placeholders are pre-fed
(`feed(name=value)` / `feed_queue([...])`) before the reference builds its "graph", so building == running once;
`Session.run` returns the values that were computed at build time.
"""
from __future__ import annotations

import contextlib
import math
from collections import OrderedDict

import numpy as _np
import torch as _torch

float32, float16, float64 = _torch.float32, _torch.float16, _torch.float64
int32, int64, bool = _torch.int32, _torch.int64, _torch.bool   # noqa: A001  (tf.bool)
newaxis = None
_pybool = __builtins__["bool"] if isinstance(__builtins__, dict) else __builtins__.bool


# ------------------------------------------------------------------------------------------ tensors
class Dimension:
    def __init__(self, v):
        self.value = v

    def __int__(self):
        return int(self.value)

    __index__ = __int__

    def __eq__(self, o):
        return self.value == (o.value if isinstance(o, Dimension) else o)

    def __repr__(self):
        return "Dimension(%r)" % (self.value,)


class TensorShape(list):
    def __init__(self, dims=()):
        super().__init__(d if isinstance(d, Dimension) else Dimension(d) for d in dims)

    def as_list(self):
        return [d.value for d in self]


class Tensor(_torch.Tensor):
    """torch.Tensor with the handful of tf.Tensor methods the reference calls."""

    def get_shape(self):
        return TensorShape(tuple(_torch.Tensor.size(self)))

    def __iadd__(self, o):      # `biases += tf.constant(...)` / `i += 1` rebinding, never in place
        return self + o


def _t(x, dtype=None):
    if isinstance(x, _torch.Tensor):
        y = x if dtype is None or x.dtype == dtype else x.to(dtype)
        return y.as_subclass(Tensor)
    a = _np.asarray(x)
    if dtype is None:
        dtype = float32 if a.dtype.kind == "f" else (int32 if a.dtype.kind in "iu" else (bool if a.dtype.kind == "b" else None))
    return _torch.as_tensor(a).to(dtype).as_subclass(Tensor)


def _i(v):
    return int(v)


def _axis_kw(axis, keepdims, keep_dims):
    return axis, _pybool(keepdims or keep_dims)


# --------------------------------------------------------------------------- placeholders / session
_FEED_BY_NAME, _FEED_QUEUE = {}, []
RECORD = {"top_k": [], "dropout": []}        # hooks the generator reads back (kNN indices, dropout call count)
DROPOUT_MASKS = []                           # pre-generated keep masks, consumed in call order
STATE = {"variables": OrderedDict(), "trainable": OrderedDict(), "scope": [], "last_grads": None, "adam": {}}


def reset():
    _FEED_BY_NAME.clear()
    del _FEED_QUEUE[:]
    del DROPOUT_MASKS[:]
    RECORD["top_k"], RECORD["dropout"] = [], []
    STATE.update(variables=OrderedDict(), trainable=OrderedDict(), scope=[], last_grads=None, adam={}, preset={})


def feed(**by_name):
    _FEED_BY_NAME.update(by_name)


def feed_queue(values):
    _FEED_QUEUE.extend(values)


def preset_variables(values):
    """name -> ndarray used instead of the initializer (so the oracle and the CUDA path load identical weights)."""
    STATE.setdefault("preset", {}).update(values)


def placeholder(dtype, shape=None, name=None):
    if name is not None and name in _FEED_BY_NAME:
        v = _FEED_BY_NAME[name]
    elif _FEED_QUEUE:
        v = _FEED_QUEUE.pop(0)
    else:
        raise RuntimeError("tf1_shim: placeholder %r must be pre-fed (eager execution)" % (name,))
    t = _t(_np.asarray(v), dtype)
    if shape is not None and shape != ():
        want = [None if s is None else int(s) for s in shape]
        assert len(want) == t.dim() and all(w is None or w == g for w, g in zip(want, t.shape)), (name, want, tuple(t.shape))
    return t


class _GpuOptions:
    allow_growth = False


class ConfigProto:
    def __init__(self, **kw):
        self.gpu_options = _GpuOptions()


def _fetch(f):
    if isinstance(f, (list, tuple)):
        return type(f)(_fetch(x) for x in f) if isinstance(f, tuple) else [_fetch(x) for x in f]
    if isinstance(f, _torch.Tensor):
        return f.detach().numpy().copy()
    return f


class Session:
    def __init__(self, config=None):
        pass

    def run(self, fetches, feed_dict=None):
        for ph, val in (feed_dict or {}).items():      # the value was consumed at build time: it must be the same one
            if isinstance(ph, _torch.Tensor):
                assert _np.array_equal(ph.detach().numpy(), _np.asarray(val).astype(ph.detach().numpy().dtype)), \
                    "tf1_shim: feed_dict differs from the pre-fed value"
        return _fetch(fetches)


def global_variables_initializer():
    return None


def no_op(*a, **k):
    return None


def identity(x, name=None):
    return _t(x) + 0 if isinstance(x, _torch.Tensor) else x


@contextlib.contextmanager
def device(_):
    yield


@contextlib.contextmanager
def control_dependencies(_):
    yield


def add_to_collection(*a, **k):
    return None


# ------------------------------------------------------------------------------------------- variables
class _Scope:
    def __init__(self, name):
        self.name = name


@contextlib.contextmanager
def variable_scope(name, *a, **k):
    STATE["scope"].append(name)
    try:
        yield _Scope("/".join(STATE["scope"]))
    finally:
        STATE["scope"].pop()


def _make_variable(full, value, trainable):
    v = _torch.tensor(_np.asarray(value)).clone().as_subclass(Tensor)
    if trainable and v.dtype.is_floating_point:
        v.requires_grad_(True)
        STATE["trainable"][full] = v
    STATE["variables"][full] = v
    return v


def get_variable(name, shape=None, initializer=None, dtype=float32, trainable=True):
    full = "/".join(STATE["scope"] + [name])
    if full in STATE["variables"]:                      # AUTO_REUSE: a second model build shares the weights
        return STATE["variables"][full]
    preset = STATE.get("preset", {})
    if full in preset:
        val = _np.asarray(preset[full], _np.float32).reshape([int(s) for s in shape])
    else:
        val = initializer([int(s) for s in shape])
    return _make_variable(full, val, trainable)


_VAR_COUNT = [0]


def Variable(initial_value, trainable=True, name=None, dtype=None):
    _VAR_COUNT[0] += 1
    a = _np.asarray(initial_value)
    if a.dtype.kind == "f":
        a = a.astype(_np.float32)
    elif a.dtype.kind in "iu":
        a = a.astype(_np.int32)
    if name is not None:      # named variables live under the enclosing scopes and honour presets (non-dist batch norm: beta, gamma)
        full = "/".join(STATE["scope"] + [name])
        if full in STATE["variables"]:
            return STATE["variables"][full]
        preset = STATE.get("preset", {})
        if full in preset:
            a = _np.asarray(preset[full], a.dtype).reshape(a.shape)
        return _make_variable(full, a, trainable)
    return _make_variable("Variable_%d" % _VAR_COUNT[0], a, trainable)


def assign(ref, value):
    ref.data.copy_(_torch.as_tensor(value).detach())
    return ref


def zeros_initializer():
    return lambda shape: _np.zeros(shape, _np.float32)


def ones_initializer():
    return lambda shape: _np.ones(shape, _np.float32)


def constant_initializer(value=0.0):
    return lambda shape: _np.full(shape, value, _np.float32)


_INIT_RNG = _np.random.default_rng(20260101)


def truncated_normal_initializer(stddev=1.0, mean=0.0):
    def init(shape):
        x = _INIT_RNG.standard_normal(shape)
        bad = _np.abs(x) > 2
        while bad.any():
            x[bad] = _INIT_RNG.standard_normal(int(bad.sum()))
            bad = _np.abs(x) > 2
        return (mean + stddev * x).astype(_np.float32)
    return init


def _xavier(shape):
    rf = int(_np.prod(shape[:-2])) if len(shape) > 2 else 1
    lim = math.sqrt(6.0 / (shape[-2] * rf + shape[-1] * rf))
    return _INIT_RNG.uniform(-lim, lim, shape).astype(_np.float32)


# ---------------------------------------------------------------------------------------- array ops
def constant(value, dtype=None, shape=None, name=None):
    if shape is not None and _np.ndim(value) == 0:        # tf.constant(0.0, shape=[C]) fills the shape
        value = _np.full([_i(d) for d in shape], value, _np.float32 if isinstance(value, float) else None)
    return _t(value, dtype)


def cast(x, dtype, name=None):
    return _t(x).to(dtype).as_subclass(Tensor) if isinstance(x, _torch.Tensor) else _t(x, dtype)


def shape(x, name=None):     # noqa: A001
    return _t(_np.asarray(tuple(_torch.Tensor.size(x)), _np.int32))


def expand_dims(x, axis=None, name=None, dim=None):
    return _t(x).unsqueeze(axis if axis is not None else dim)


def squeeze(x, axis=None, name=None, squeeze_dims=None):
    axis = axis if axis is not None else squeeze_dims
    if axis is None:
        return _t(x).squeeze()
    for a in sorted([axis] if isinstance(axis, int) else list(axis), reverse=True):
        x = _t(x).squeeze(a)
    return x


def transpose(x, perm=None, name=None):
    return _t(x).permute(*perm) if perm is not None else _t(x).t()


def reshape(x, shape, name=None):     # noqa: A002
    return _t(x).reshape([_i(s) for s in shape])


def tile(x, multiples, name=None):
    return _t(x).repeat(*[_i(m) for m in multiples])


def concat(values=None, axis=None, name=None, **kw):
    if isinstance(values, int):          # legacy tf.concat(axis, values)
        values, axis = axis, values
    return _torch.cat([_t(v) for v in values], dim=axis).as_subclass(Tensor)


def stack(values, axis=0, name=None):
    return _torch.stack([_t(v) if isinstance(v, _torch.Tensor) else _t(_np.asarray(v)) for v in values], dim=axis).as_subclass(Tensor)


def unstack(x, num=None, axis=0, name=None):
    return [t.as_subclass(Tensor) for t in _torch.unbind(_t(x), dim=axis)]


def gather(params, indices, axis=0, name=None):
    idx = _t(indices).long()
    axis = axis % params.dim()          # tf.gather accepts negative axes (Util/Loss.py:143)
    return _torch.index_select(_t(params), axis, idx.reshape(-1)).reshape(
        tuple(params.shape[:axis]) + tuple(idx.shape) + tuple(params.shape[axis + 1:])).as_subclass(Tensor)


def batch_gather(params, indices, name=None):
    idx = _t(indices).long()
    assert params.dim() == idx.dim()
    return _torch.gather(_t(params), params.dim() - 1, idx).as_subclass(Tensor)


def range(*a, **k):     # noqa: A001
    return _t(_np.arange(*[_i(v) for v in a], dtype=_np.int32))


def fill(dims, value, name=None):
    return _t(_np.full([_i(d) for d in dims], value, _np.float32 if isinstance(value, float) else _np.int32))


def zeros_like(x, dtype=None, name=None):
    return _torch.zeros_like(_t(x)).as_subclass(Tensor)


def ones_like(x, dtype=None, name=None):
    return _torch.ones_like(_t(x)).as_subclass(Tensor)


def eye(n, dtype=float32, **k):
    return _torch.eye(_i(n), dtype=dtype).as_subclass(Tensor)


def diag(d, name=None):
    return _torch.diag(_t(d)).as_subclass(Tensor)


def matrix_diag(d, name=None):
    return _torch.diag_embed(_t(d)).as_subclass(Tensor)


def while_loop(cond, body, loop_vars, shape_invariants=None, **k):
    vs = tuple(loop_vars)
    while _pybool(cond(*vs)):
        vs = tuple(body(*vs))
    return vs


def cond(pred, true_fn=None, false_fn=None, name=None, fn1=None, fn2=None):
    return (true_fn or fn1)() if _pybool(pred) else (false_fn or fn2)()


# ----------------------------------------------------------------------------------------- math ops
def _red(fn, x, axis, keepdims, keep_dims):
    x = _t(x)
    kd = _pybool(keepdims or keep_dims)
    if axis is None:
        return fn(x)
    return fn(x, dim=axis, keepdim=kd)


def reduce_sum(x, axis=None, keepdims=False, name=None, keep_dims=False, reduction_indices=None):
    return _red(_torch.sum, x, axis if axis is not None else reduction_indices, keepdims, keep_dims)


def reduce_mean(x, axis=None, keepdims=False, name=None, keep_dims=False):
    return _red(_torch.mean, x, axis, keepdims, keep_dims)


def reduce_max(x, axis=None, keepdims=False, name=None, keep_dims=False):
    x = _t(x)
    if not x.dtype.is_floating_point:
        return _red(lambda t, **kw: _torch.amax(t, **kw), x, axis, keepdims, keep_dims)
    # TF MaxGrad: the gradient is split equally among tied maxima -- torch.amax has the same rule
    return _red(lambda t, **kw: _torch.amax(t, **kw), x, axis, keepdims, keep_dims)


def reduce_min(x, axis=None, keepdims=False, name=None, keep_dims=False):
    return _red(lambda t, **kw: _torch.amin(t, **kw), x, axis, keepdims, keep_dims)


def square(x, name=None):
    return _t(x) * _t(x)


def sqrt(x, name=None):
    return _torch.sqrt(_t(x))


def exp(x, name=None):
    return _torch.exp(_t(x))


def log(x, name=None):
    return _torch.log(_t(x) if isinstance(x, _torch.Tensor) else _t(_np.float32(x)))


def maximum(a, b, name=None):
    a, b = (_t(v) if isinstance(v, _torch.Tensor) else _t(_np.float32(v)) for v in (a, b))
    return _torch.maximum(a, b).as_subclass(Tensor)


def minimum(a, b, name=None):
    a, b = (_t(v) if isinstance(v, _torch.Tensor) else _t(_np.float32(v)) for v in (a, b))
    return _torch.minimum(a, b).as_subclass(Tensor)


def multiply(a, b, name=None):
    return _t(a) * b


def matmul(a, b, transpose_a=False, transpose_b=False, name=None):
    a, b = _t(a), _t(b)
    if transpose_a:
        a = a.transpose(-1, -2)
    if transpose_b:
        b = b.transpose(-1, -2)
    return _torch.matmul(a, b).as_subclass(Tensor)


def einsum(eq, *ops):
    return _torch.einsum(eq, *[_t(o) for o in ops]).as_subclass(Tensor)


def less(a, b, name=None):
    return _t(_np.asarray(_pybool(a < b))) if not isinstance(a, _torch.Tensor) and not isinstance(b, _torch.Tensor) else (a < b)


def greater_equal(a, b, name=None):
    if not isinstance(a, _torch.Tensor) and not isinstance(b, _torch.Tensor):
        return _t(_np.asarray(a >= b))
    return a >= b


def argmax(x, axis=None, name=None, **k):
    return _torch.argmax(_t(x), dim=axis).as_subclass(Tensor)


def argsort(x, axis=-1, direction="ASCENDING", stable=False, name=None):
    return _torch.sort(_t(x), dim=axis, descending=direction != "ASCENDING", stable=True).indices.to(int32).as_subclass(Tensor)


def clip_by_value(x, lo, hi, name=None):
    return _torch.clamp(_t(x), lo, hi)


class _Math:
    log = staticmethod(log)
    equal = staticmethod(lambda a, b, name=None: _t(a) == b)


def one_hot(indices, depth, on_value=1.0, off_value=0.0, axis=-1, dtype=float32, name=None):
    return _torch.nn.functional.one_hot(_t(indices).long(), _i(depth)).to(dtype).as_subclass(Tensor)


class _Losses:
    @staticmethod
    def softmax_cross_entropy(onehot_labels, logits, weights=1.0, label_smoothing=0, scope=None, **k):
        """tf.losses.softmax_cross_entropy (TF 1.14): labels*(1-ls) + ls/num_classes, per-example CE, weights = 1 ->
        reduction SUM_BY_NONZERO_WEIGHTS = mean over the batch"""
        lab = _t(onehot_labels).to(logits.dtype)
        if label_smoothing > 0:
            lab = lab * (1 - label_smoothing) + label_smoothing / lab.shape[-1]
        return (-(lab * _torch.log_softmax(_t(logits), dim=-1)).sum(-1)).mean().as_subclass(Tensor)


losses = _Losses()


math = _Math()     # noqa: A001  (shadows the stdlib name inside this module only after its last use above)


class _Linalg:
    inv = staticmethod(lambda x, name=None: _torch.linalg.inv(_t(x)).as_subclass(Tensor))


linalg = _Linalg()


class _Random:
    @staticmethod
    def uniform(shape, minval=0.0, maxval=1.0, dtype=float32, seed=None, name=None):
        return _t(_INIT_RNG.uniform(minval, maxval, [_i(s) for s in shape]).astype(_np.float32))


random = _Random()
random_uniform = _Random.uniform


# -------------------------------------------------------------------------------------------- tf.nn
class _NN:
    @staticmethod
    def relu(x, name=None):
        return _torch.relu(_t(x))

    @staticmethod
    def sigmoid(x, name=None):
        return _torch.sigmoid(_t(x))

    @staticmethod
    def softmax(x, axis=-1, name=None, dim=None):
        return _torch.softmax(_t(x), dim=dim if dim is not None else axis)

    @staticmethod
    def bias_add(x, b, name=None):
        return _t(x) + b

    @staticmethod
    def l2_loss(x, name=None):
        return (_t(x) * x).sum() / 2

    @staticmethod
    def conv2d(x, kernel, strides, padding, name=None, **k):
        """NHWC x HWIO.  The hot path only uses 1x1 kernels with stride 1 (tf_util.py:159-161)."""
        kh, kw, cin, cout = kernel.shape
        assert (kh, kw) == (1, 1) and list(strides) == [1, 1, 1, 1], "tf1_shim.conv2d: only 1x1/stride-1 kernels"
        return _torch.matmul(_t(x), _t(kernel).reshape(cin, cout)).as_subclass(Tensor)

    @staticmethod
    def moments(x, axes, name=None, keep_dims=False, keepdims=False):
        x = _t(x)
        mean = x.mean(dim=list(axes), keepdim=True)
        var = ((x - mean.detach()) ** 2).mean(dim=list(axes), keepdim=True)   # TF: squared_difference(x, stop_gradient(mean))
        if not (keep_dims or keepdims):
            mean, var = mean.reshape(-1) if len(axes) == x.dim() - 1 else mean.squeeze(list(axes)), \
                var.reshape(-1) if len(axes) == x.dim() - 1 else var.squeeze(list(axes))
        return mean, var

    @staticmethod
    def batch_normalization(x, mean, variance, offset, scale, variance_epsilon, name=None):
        inv = _torch.rsqrt(_t(variance) + variance_epsilon)
        if scale is not None:
            inv = inv * scale
        return _t(x) * inv + ((offset - mean * inv) if offset is not None else (-mean * inv))

    @staticmethod
    def max_pool(x, ksize, strides, padding, name=None, **k):
        assert padding == "VALID"
        y = _torch.nn.functional.max_pool2d(_t(x).permute(0, 3, 1, 2), kernel_size=(ksize[1], ksize[2]),
                                            stride=(strides[1], strides[2]))
        return y.permute(0, 2, 3, 1).as_subclass(Tensor)

    @staticmethod
    def avg_pool(x, ksize, strides, padding, name=None, **k):
        y = _torch.nn.functional.avg_pool2d(_t(x).permute(0, 3, 1, 2), kernel_size=(ksize[1], ksize[2]),
                                            stride=(strides[1], strides[2]))
        return y.permute(0, 2, 3, 1).as_subclass(Tensor)

    @staticmethod
    def dropout(x, keep_prob=None, noise_shape=None, seed=None, name=None, rate=None):
        keep = keep_prob if keep_prob is not None else 1.0 - rate
        x = _t(x)
        RECORD["dropout"].append(tuple(x.shape))
        assert DROPOUT_MASKS, "tf1_shim: queue a keep mask in DROPOUT_MASKS before building a training graph"
        mask = _t(_np.asarray(DROPOUT_MASKS.pop(0), _np.float32)).reshape(x.shape)
        return x / keep * mask            # tf.nn.dropout (1.x): div(x, keep_prob) * floor(keep_prob + U[0,1))

    @staticmethod
    def top_k(x, k=1, sorted=True, name=None):      # noqa: A002
        """values descending; among equal values the lower index first (TopK kernel uses a stable comparison)."""
        vals, idx = _torch.sort(_t(x).detach(), dim=-1, descending=True, stable=True)
        idx = idx[..., :_i(k)]
        RECORD["top_k"].append(idx.numpy().astype(_np.int32).copy())
        return _torch.gather(_t(x), -1, idx).as_subclass(Tensor), idx.to(int32).as_subclass(Tensor)

    @staticmethod
    def softmax_cross_entropy_with_logits(labels=None, logits=None, dim=-1, name=None, **k):
        lab = _t(labels).to(logits.dtype)
        return -(lab * _torch.log_softmax(_t(logits), dim=dim)).sum(dim=dim)

    softmax_cross_entropy_with_logits_v2 = softmax_cross_entropy_with_logits

    @staticmethod
    def sparse_softmax_cross_entropy_with_logits(labels=None, logits=None, name=None, **k):
        lp = _torch.log_softmax(_t(logits), dim=-1)
        return -_torch.gather(lp, -1, _t(labels).long().unsqueeze(-1)).squeeze(-1)

    @staticmethod
    def sigmoid_cross_entropy_with_logits(labels=None, logits=None, name=None, **k):
        x, z = _t(logits), _t(labels).to(logits.dtype)
        return _torch.relu(x) - x * z + _torch.log1p(_torch.exp(-_torch.abs(x)))      # max(x,0) - x z + log(1 + e^-|x|)


nn = _NN()


# ----------------------------------------------------------------------------------------- tf.train
class _EMA:
    """tf.train.ExponentialMovingAverage as tf_util.batch_norm_template (:462-499) uses it: the shadow values of (batch_mean,
    batch_var) are the population statistics read in inference mode.  They are kept under `<scope>/pop_mean`, `<scope>/pop_var`
    (the names of the dist template, so one weight dictionary serves both).  Only the inference read is implemented."""

    def __init__(self, decay=None, **k):
        self.decay, self._reads = decay, 0

    def apply(self, tensors):
        raise NotImplementedError("tf1_shim: training-mode ExponentialMovingAverage (zero-debiased shadow update) is not modelled")

    def average(self, tensor):
        name = ("pop_mean", "pop_var")[self._reads % 2]       # batch_norm_template reads mean, then variance
        self._reads += 1
        full = "/".join(STATE["scope"] + [name])
        if full not in STATE["variables"]:
            preset = STATE.get("preset", {})
            val = preset[full] if full in preset else (_np.zeros if name == "pop_mean" else _np.ones)(tensor.shape, _np.float32)
            _make_variable(full, _np.asarray(val, _np.float32).reshape(tuple(tensor.shape)), False)
        return STATE["variables"][full]


class _Saver:
    def __init__(self, *a, **k):
        pass


class _Adam:
    """tf.train.AdamOptimizer: lr_t = lr sqrt(1-b2^t)/(1-b1^t); m,v EMA; var -= lr_t m / (sqrt(v) + eps)."""

    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8, **k):
        self.lr, self.b1, self.b2, self.eps = learning_rate, beta1, beta2, epsilon

    def minimize(self, loss, global_step=None, var_list=None):
        names = list(STATE["trainable"])
        vs = [STATE["trainable"][n] for n in names]
        grads = _torch.autograd.grad(loss, vs, allow_unused=True)
        STATE["last_grads"] = OrderedDict((n, None if g is None else g.detach().clone()) for n, g in zip(names, grads))
        st = STATE["adam"]
        st["t"] = st.get("t", 0) + 1
        t = st["t"]
        lr = float(self.lr)
        lr_t = _np.float32(lr * math_sqrt(1 - self.b2 ** t) / (1 - self.b1 ** t))
        for n, v, g in zip(names, vs, grads):
            if g is None:
                continue
            m = st.setdefault("m/" + n, _torch.zeros_like(g))
            s = st.setdefault("v/" + n, _torch.zeros_like(g))
            m.mul_(self.b1).add_(g, alpha=1 - self.b1)
            s.mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
            v.data.sub_(lr_t * m / (s.sqrt() + self.eps))
        if global_step is not None:
            global_step.data.add_(1)
        return None


def _exponential_decay(learning_rate, global_step, decay_steps, decay_rate, staircase=False, name=None):
    p = _t(global_step).to(float32) / _np.float32(decay_steps)
    if staircase:
        p = _torch.floor(p)
    return (_np.float32(learning_rate) * _torch.pow(_t(_np.float32(decay_rate)), p)).as_subclass(Tensor)


class _Train:
    ExponentialMovingAverage = _EMA
    Saver = _Saver
    AdamOptimizer = _Adam
    exponential_decay = staticmethod(_exponential_decay)


train = _Train()
import math as _pymath   # noqa: E402

math_sqrt = _pymath.sqrt


class _Layers:
    xavier_initializer = staticmethod(lambda *a, **k: _xavier)


class _Contrib:
    layers = _Layers()


contrib = _Contrib()
