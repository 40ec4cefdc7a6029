"""Drop-in for the reference trainer class `ShapeNet_Trainer` (ShapeNet/ShapeNet_DGCNN_trainer.py:19-644).

Same method names and schedules (DECAY_STEP = 16881*20, :31); one `train_batch` == one
`sess.run([solver, loss, loss_siamese, loss_inexact, loss_smooth, Z_prob], feed_dict)` of TrainOneEpoch_Full
(:308-314) on the fused CUDA executor ShapeNetEngine."""
from __future__ import annotations

import math
import sys

import numpy as np
import torch

from . import Tool
from .engine_shapenet import LAYERS, ShapeNetEngine
from .S3DIS_DGCNN_trainer import S3DIS_Trainer, xavier_params


class ShapeNet_Trainer(S3DIS_Trainer):

    def __init__(self, device=None, seed=None):
        super().__init__(test_area=None, device=device, seed=seed)

    def SetLearningRate(self, LearningRate=1e-3, BatchSize=12):
        super().SetLearningRate(LearningRate, BatchSize)
        self.DECAY_STEP = 16881 * 20                       # (:31)
        self.BN_DECAY_DECAY_STEP = float(self.DECAY_STEP * 2)

    def defineNetwork(self, batch_size, point_num=2048, style='Full', rampup=101, params=None):
        self.rampup = rampup
        self.style = style
        if style not in ('Plain', 'Full'):
            sys.exit('Loss {} is not defined!'.format(style))
        if not hasattr(self, 'BATCH_SIZE'):
            self.SetLearningRate(1e-3, max(batch_size // 2, 1))
        if params is None:
            params = xavier_params(LAYERS, self.seed, shapenet=True)
        self.engine = ShapeNetEngine(params, batch_size, point_num, device=self.device)
        self.epoch = 0
        self.weak_gate = (style == 'Full') and (self.epoch >= self.rampup)   # frozen at build time (:92,:100)
        self.pinned = {}
        return True

    def train_batch(self, data_feed, label_onehot_feed, seg_onehot_feed, Mask_bin_feed, fetch_prob=True, dropout_masks=None):
        """feed order of the reference: X_ph, Label_ph (category one-hot), Y_ph, Mask_ph (:308-314)"""
        eng = self.engine
        X = self._to_device('X', data_feed)
        Lb = self._to_device('Label', label_onehot_feed)
        Y = self._to_device('Y', seg_onehot_feed)
        M = self._to_device('Mask', Mask_bin_feed)
        lr, decay = self.get_learning_rate(), self.get_bn_decay()
        full = self.style == 'Full'
        eng.forward(X, Lb, True, decay, dropout_masks)
        if full and not self.weak_gate:
            eng.losses_and_grad(Y, M, full=True, want_grad=False)
            weak = self._fetch_losses()
            eng.losses_and_grad(Y, M, full=False, want_grad=True)
        else:
            weak = None
            eng.losses_and_grad(Y, M, full=full, want_grad=True)
        eng.backward()
        self._allreduce_and_step(lr)
        zp = self._fetch_prob() if fetch_prob else None
        l = self._fetch_losses()
        if weak is not None:
            return float(l[0]), float(weak[1]), float(weak[2]), float(weak[3]), zp
        return float(l[4]), float(l[1]), float(l[2]), float(l[3]), zp

    def eval_batch(self, data_feed, label_onehot_feed, seg_onehot_feed, Mask_bin_feed):
        eng = self.engine
        X = self._to_device('X', data_feed)
        Lb = self._to_device('Label', label_onehot_feed)
        Y = self._to_device('Y', seg_onehot_feed)
        M = self._to_device('Mask', Mask_bin_feed)
        eng.forward(X, Lb, False, None)
        full = self.style == 'Full' and eng.B % 2 == 0
        eng.losses_and_grad(Y, M, full=full, want_grad=False)
        zp = self._fetch_prob()
        l = self._fetch_losses()
        return (float(l[4]) if (full and self.weak_gate) else float(l[0])), zp.copy()

    def defLabelPropSolver(self, alpha=1e0, beta=1e0, K=10):
        """(:136-140) the arguments are ignored by the reference as well"""
        from . import ProbLabelPropagation as PLP
        self.LPSolver = PLP.LabelPropagation_TF(alpha=1e0, beta=1e0, K=10)
        self.TFComp = {'Lmat': Tool.TF_Computation.LaplacianMatSym_XYZRGB_DirectComp()}
