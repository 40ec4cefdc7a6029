"""Drop-in for ShapeNet/DGCNN_ShapeNet.py (reference :15-122): ShapeNet part-segmentation DGCNN constructor.

`get_model(point_cloud, input_label, is_training, cat_num, part_num, batch_size, num_point, weight_decay,
bn_decay=None)` keeps the reference signature and returns the logits (B,N,part_num) of the fused CUDA executor
(engine_shapenet.ShapeNetEngine)."""
from __future__ import annotations

import torch

from . import _lib as L
from .engine_shapenet import LAYERS, ShapeNetEngine

_ENGINES: dict = {}
_PARAMS = None


def set_variables(params):
    global _PARAMS
    _PARAMS = params
    _ENGINES.clear()


def get_engine(B, N, device):
    key = (B, N, str(device))
    if key not in _ENGINES:
        from .S3DIS_DGCNN_trainer import xavier_params
        _ENGINES[key] = ShapeNetEngine(_PARAMS if _PARAMS is not None else xavier_params(LAYERS, 0, shapenet=True), B, N,
                                       device=device)
    return _ENGINES[key]


def get_model(point_cloud, input_label, is_training, cat_num, part_num, batch_size, num_point, weight_decay, bn_decay=None):
    B, N, _ = point_cloud.shape
    eng = get_engine(B, N, point_cloud.device)
    return eng.forward(point_cloud.contiguous(), input_label.to(torch.float32).reshape(B, cat_num).contiguous(),
                       bool(is_training), bn_decay)


def get_loss(seg_pred, seg):
    """(:116-122) -> seg_loss, per_instance_seg_loss (B), per_instance_seg_pred_res (B,N)"""
    B, N, C = seg_pred.shape
    dev = seg_pred.device
    Y = torch.zeros((B, N, C), dtype=torch.float32, device=dev)
    Y.scatter_(2, seg.long().unsqueeze(-1), 1.0)
    P = torch.empty_like(seg_pred)
    per = torch.empty(B, dtype=torch.float32, device=dev)
    losses = torch.empty(5, dtype=torch.float32, device=dev)
    ws = L.workspace(L.lib().wspc_head_losses_workspace_bytes(1, N, C), dev, "head")
    M = torch.ones((1, N), dtype=torch.float32, device=dev)
    for b in range(B):   # per-instance mean CE: one masked-CE evaluation per cloud
        L.check(L.lib().wspc_head_losses(L.ptr(seg_pred[b].contiguous()), L.ptr(Y[b].contiguous()), L.ptr(M), None, None, 1, N, C,
                                         0, 0.1, 0.0, 0, 0, L.ptr(P[b]), None, L.ptr(losses), L.ptr(ws), ws.numel(), L.stream()))
        per[b] = losses[0]
    return per.mean(), per, torch.argmax(seg_pred, 2)
