"""tf1_shim: names Util/Loss.py uses from tensorflow.python.ops.array_ops (zeros_like / ones_like / where)."""
import torch as _torch

from tensorflow import zeros_like, ones_like, Tensor  # noqa: F401


def where(condition, x=None, y=None, name=None):
    """array_ops.where(cond, x, y): element-wise select (the three-argument form is the only one Util/Loss.py uses)."""
    return _torch.where(condition, x, y).as_subclass(Tensor)
