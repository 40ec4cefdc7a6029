"""Drop-in for Util/Loss.py (reference :5-195): focal / class-weighted CE / self-entropy / "Overwhelm" losses.

The reference imports this module (ShapeNet_DGCNN_trainer.py:5) but never calls any of it: it is
API surface only and off the data-parallel hot path, so the functions are plain tensor expressions with the same
names, argument order and return values as the reference definitions."""
from __future__ import annotations

import torch


def focal_loss(prediction_tensor, target_tensor, weights=None, alpha=0.25, gamma=2):
    """FL = -alpha (z-p)^gamma log(p) - (1-alpha) p^gamma log(1-p), p = sigmoid(x)   (Util/Loss.py:5-33).
    Returns the per-entry tensor like the reference."""
    p = torch.sigmoid(prediction_tensor)
    zeros = torch.zeros_like(p)
    pos = torch.where(target_tensor > zeros, target_tensor - p, zeros)
    neg = torch.where(target_tensor > zeros, zeros, p)
    return (-alpha * pos ** gamma * torch.log(p.clamp(1e-8, 1.0))
            - (1 - alpha) * neg ** gamma * torch.log((1.0 - p).clamp(1e-8, 1.0)))


def focal_loss_v1(prediction_tensor, target_tensor, alpha=None, weights=None, gamma=2):
    """as focal_loss with a per-entry alpha tensor (Util/Loss.py:36-69)"""
    if alpha is None:
        alpha = 0.25 * torch.ones_like(prediction_tensor)
    return focal_loss(prediction_tensor, target_tensor, weights, alpha, gamma)


def class_weighted_CE_loss(pred, gt, posWeight, negWeight):
    '''Weighted sigmoid cross entropy (Util/Loss.py:72-84): pred, gt B*1*K'''
    p = torch.sigmoid(pred)
    return -(posWeight * gt * torch.log(p.clamp(1e-8, 1.0)) + negWeight * (1 - gt) * torch.log((1 - p).clamp(1e-8, 1.0)))


def SelfEntropy(Z):
    '''sum_k softmax(Z) log(softmax(Z) + 1e-5) over the last axis -> B*N   (Util/Loss.py:86-98)'''
    Z_hat = torch.softmax(Z, dim=-1)
    return (Z_hat * torch.log(Z_hat + 1e-5)).sum(-1)


def OverwhelmLoss_v1(L, Y):
    '''max logit of positive class j1 should exceed the min logit of positive class j2 (Util/Loss.py:100-124)'''
    K = Y.shape[-1]
    L_max = L.amax(dim=1).unsqueeze(-1).expand(-1, K, K)
    L_min = L.amin(dim=1, keepdim=True).expand(-1, K, K)
    pen = torch.clamp(L_min - L_max, min=0)
    Ym = Y.unsqueeze(-1)
    Mask = torch.einsum('ijk,ilk->ijl', Ym, Ym) - torch.diag_embed(Y)
    return (pen * Mask).mean(dim=(-1, -2)).mean()


def OverwhelmLoss_v2(L, Y):
    '''per-class positive / negative gap penalties (Util/Loss.py:127-170) -> loss, loss_full_pos, loss_full_neg'''
    K = L.shape[-1]
    pos, neg = [], []
    for k in range(K):
        L_k = L[..., k]
        others = torch.cat([L[..., :k], L[..., k + 1:]], dim=-1).amax(dim=-1)
        pos.append(Y[..., k] * torch.clamp((others - L_k).amin(dim=1), min=0))
        neg.append((1 - Y[..., k]) * torch.clamp((L_k - others).amax(dim=1), min=0))
    loss_full_pos, loss_full_neg = torch.stack(pos, -1), torch.stack(neg, -1)
    return (loss_full_pos + loss_full_neg).mean(), loss_full_pos, loss_full_neg


def OverwhelmLoss(L, Y):
    '''for at least one point the logit of a positive class should dominate (Util/Loss.py:173-195) -> loss, loss_full'''
    max_j = L.amax(dim=-1, keepdim=True)
    gap = torch.clamp((max_j - L).amin(dim=1), min=0)
    loss_full = Y * gap
    return loss_full.sum(-1).mean(), loss_full
