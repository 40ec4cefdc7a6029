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
from .S3DIS_DGCNN_trainer import S3DIS_Trainer, prefetched, xavier_params


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
        if self.seed is not None:
            self.engine.seed = 4321 + 7919 * int(self.seed)
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
        gate_closed = full and not self.weak_gate
        eng.losses_and_grad(Y, M, full=2 if gate_closed else full, want_grad=True)       # gate closed: one pass, see S3DIS trainer
        eng.backward()
        self._allreduce_and_step(lr)
        zp = self._fetch_prob() if fetch_prob else None
        l = self._fetch_losses()
        if gate_closed:
            return float(l[0]), float(l[1]), float(l[2]), float(l[3]), zp
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

    # ------------------------------------------------------------------ epoch loops --------------
    @staticmethod
    def _restrict_to_category(Z_prob, iou_oids):
        """the prediction is taken among the part ids of the shape's category (reference :319-323, :486-489)"""
        z = np.array(Z_prob, copy=True)
        z[:, list(iou_oids)] += 1
        return np.argmax(z, axis=-1)

    def TrainOneEpoch(self, Loader, file_idx_list, data_idx_list, pts_idx_list=None):
        """Plain-style epoch (ShapeNet_DGCNN_trainer.py:143-218): the same loop without the Siamese partner."""
        return self._train_epoch(Loader, file_idx_list, data_idx_list, pts_idx_list, siamese=False)

    def TrainOneEpoch_Full(self, Loader, file_idx_list, data_idx_list, pts_idx_list=None):
        return self._train_epoch(Loader, file_idx_list, data_idx_list, pts_idx_list, siamese=True)

    def _train_epoch(self, Loader, file_idx_list, data_idx_list, pts_idx_list, siamese):
        """One training epoch (ShapeNet_DGCNN_trainer.py:220-341).  `Loader` follows ShapeNetIO.NextBatch_TrainSet
        (DataIO_ShapeNet.py:145-193): (flag, data (b,N,3), label (b,1), seg (b,N), weak_seg_onehot, mb_size, file_idx,
        data_idx) and exposes NUM_CATEGORIES, objcats, object2setofoid.  Mask from (file_idx_list, data_idx_list,
        pts_idx_list) as in :243-251; Siamese partner = jitter 2e-3 * extent + random mirror of axis 2 once the ramp-up
        epoch is reached (:260-275), a plain copy before; rows interleaved [sample, partner]."""
        batch_cnt, data_cnt, avg_loss, avg_acc = 1, 0, 0., 0.
        rep = 2 if siamese else 1
        feeds = prefetched(self._train_batches(Loader, file_idx_list, data_idx_list, pts_idx_list, siamese))
        for data_feed, label_onehot_feed, seg_onehot_feed, Mask_bin_feed, label, seg, mb_size in feeds:
            loss_mb, _, _, _, Z_prob_mb = self.train_batch(data_feed, label_onehot_feed, seg_onehot_feed, Mask_bin_feed)
            pred = np.stack([self._restrict_to_category(Z_prob_mb[rep * b_i],
                                                        Loader.object2setofoid[Loader.objcats[label[b_i, 0]]])
                             for b_i in range(mb_size)])
            avg_loss = (avg_loss * data_cnt + loss_mb * mb_size) / (data_cnt + mb_size)
            avg_acc = (avg_acc * data_cnt + float(np.mean(pred == seg)) * mb_size) / (data_cnt + mb_size)
            data_cnt += mb_size
            print('\rBatch {:d} TrainedSamp {:d}  Avg Loss {:.4f} Avg Acc {:.2f}%'.format(batch_cnt, data_cnt, avg_loss,
                                                                                         100 * avg_acc), end='')
            batch_cnt += 1
        self.epoch += 1
        return avg_loss, avg_acc

    def _train_batches(self, Loader, file_idx_list, data_idx_list, pts_idx_list, siamese):
        """Generator of one epoch's feeds (host numpy only; runs one batch ahead of the device step under `prefetched`):
        (data_feed, label_onehot_feed, seg_onehot_feed, Mask_bin_feed, label (b,1), seg (b,N), mb_size)."""
        rng = np.random.default_rng(1000 + self.epoch)
        file_idx_list = None if file_idx_list is None else np.asarray(file_idx_list)
        data_idx_list = None if data_idx_list is None else np.asarray(data_idx_list)
        while True:
            SuccessFlag, data, label, seg, _, mb_size, file_idx, data_idx = Loader.NextBatch_TrainSet(shuffle_flag=True)
            if not SuccessFlag:
                return
            if mb_size < (self.engine.B // 2 if siamese else self.engine.B):   # short batches are skipped (:239-240)
                continue
            data = np.asarray(data, np.float32)
            label = np.asarray(label).astype(np.int64).reshape(mb_size, -1)
            seg = np.asarray(seg).astype(np.int64)
            N = data.shape[1]
            mask = np.zeros((mb_size, N), np.float32)
            if pts_idx_list is not None:
                for b_i in range(mb_size):
                    hit = np.where((file_idx_list == file_idx[b_i]) & (data_idx_list == data_idx[b_i]))[0]
                    mask[b_i, np.asarray(pts_idx_list[hit][0]).astype(np.int64)] = 1
            partner = data.copy()
            if siamese and self.epoch >= self.rampup:
                extent = data.max(axis=1, keepdims=True) - data.min(axis=1, keepdims=True)
                partner = data + (2e-3 * extent * rng.standard_normal(data.shape)).astype(np.float32)
                flip = rng.integers(0, 2, mb_size).astype(bool)
                partner[flip, :, 2] *= -1
            rep = 2 if siamese else 1
            data_feed = np.empty((rep * mb_size, N, 3), np.float32)
            data_feed[0::rep] = data
            if siamese:
                data_feed[1::2] = partner
            label_onehot_feed = np.repeat(Tool.OnehotEncode(label[:, 0], Loader.NUM_CATEGORIES, np.float32), rep, axis=0)
            seg_onehot_feed = Tool.OnehotEncode(np.repeat(seg, rep, axis=0), 50, np.float32)
            yield data_feed, label_onehot_feed, seg_onehot_feed, np.repeat(mask, rep, axis=0), label, seg, mb_size

    def EvalOneEpoch(self, Loader, Eval):
        """Plain-style validation (ShapeNet_DGCNN_trainer.py:343-413): no Siamese duplication."""
        return self._eval_epoch(Loader, Eval, siamese=False)

    def EvalOneEpoch_Full(self, Loader, Eval):
        return self._eval_epoch(Loader, Eval, siamese=True)

    def _eval_epoch(self, Loader, Eval, siamese):
        """Validation pass (ShapeNet_DGCNN_trainer.py:417-507): short batches are padded with sample 0, every sample is
        duplicated to fill the Siamese graph, Is_Training=False, Z_prob[0::2] is scored with Eval.EvalIoU over the part ids
        of the shape's category.  Returns (avg_loss, avg_acc, perdata_miou, pershape_miou)."""
        data_cnt = 0
        shape_cnt = np.zeros(Loader.NUM_CATEGORIES)
        pershape_miou = np.zeros(Loader.NUM_CATEGORIES)
        avg_loss = avg_acc = perdata_miou = 0.
        while True:
            SuccessFlag, data, label, seg, _, mb_size, _, _ = Loader.NextBatch_ValSet()
            if not SuccessFlag:
                break
            data = np.asarray(data, np.float32)
            label = np.asarray(label).astype(np.int64).reshape(mb_size, -1)
            seg = np.asarray(seg).astype(np.int64)
            rep = 2 if siamese else 1
            pad = self.engine.B // rep - mb_size
            if pad > 0:
                data_f = np.concatenate([data, np.repeat(data[0:1], pad, 0)], 0)
                seg_f = np.concatenate([seg, np.repeat(seg[0:1], pad, 0)], 0)
                label_f = np.concatenate([label, np.repeat(label[0:1], pad, 0)], 0)
            else:
                data_f, seg_f, label_f = data, seg, label
            N = data_f.shape[1]
            loss_mb, Z_prob_mb = self.eval_batch(np.repeat(data_f, rep, axis=0),
                                                 np.repeat(Tool.OnehotEncode(label_f[:, 0], Loader.NUM_CATEGORIES, np.float32), rep, axis=0),
                                                 Tool.OnehotEncode(np.repeat(seg_f, rep, axis=0), 50, np.float32),
                                                 np.ones((rep * data_f.shape[0], N), np.float32))
            Z_prob_mb = Z_prob_mb[0:rep * mb_size:rep]
            for b_i in range(mb_size):
                shape_label = int(label[b_i, 0])
                iou_oids = Loader.object2setofoid[Loader.objcats[shape_label]]
                pred = self._restrict_to_category(Z_prob_mb[b_i], iou_oids)
                avg_iou = Eval.EvalIoU(pred, seg[b_i], iou_oids)
                perdata_miou = (perdata_miou * data_cnt + avg_iou) / (data_cnt + 1)
                pershape_miou[shape_label] = (pershape_miou[shape_label] * shape_cnt[shape_label] + avg_iou) / \
                    (shape_cnt[shape_label] + 1)
                avg_acc = (avg_acc * data_cnt + float(np.mean(pred == seg[b_i]))) / (data_cnt + 1)
                avg_loss = (avg_loss * data_cnt + loss_mb) / (data_cnt + 1)
                data_cnt += 1
                shape_cnt[shape_label] += 1
        return avg_loss, avg_acc, perdata_miou, pershape_miou

    def Test(self, Loader, Eval, style='Full'):
        """Test-time pass with label propagation (ShapeNet_DGCNN_trainer.py:511-596).  Shapes come one at a time from
        `Loader.NextSamp_TestSet`; each is resampled with replacement to the graph's point count with the reference's
        np.random calls in the reference's order (:531-534).  `engine.B` shapes then share one inference pass
        (Is_Training=False: population batch-norm statistics, so a shape's logits do not depend on its batch mates; the
        reference's graph has batch 1), and their Laplacians (RGB := XYZ, :551) and closed-form LP solves run
        concurrently on the device; the propagated probabilities of the ORIGINAL points are scored.  `loss_mb` is the
        segmentation loss of the batch (its mean over shapes equals the mean of the per-shape losses)."""
        from . import ops
        eng = self.engine
        EB = eng.B
        data_cnt = 0
        shape_cnt = np.zeros(Loader.NUM_CATEGORIES)
        pershape_miou = np.zeros(Loader.NUM_CATEGORIES)
        avg_loss = avg_acc = perdata_miou = 0.
        self.test_lp_info = []
        exhausted = False
        while not exhausted:
            pend = []
            while len(pend) < EB:
                SuccessFlag, data, label, seg, _, mb_size, _, _ = Loader.NextSamp_TestSet()
                if not SuccessFlag:
                    exhausted = True
                    break
                data = np.asarray(data, np.float32)
                label = np.asarray(label).astype(np.int64).reshape(mb_size, -1)
                seg = np.asarray(seg).astype(np.int64)
                n0 = data.shape[1]
                assert mb_size == 1 and n0 <= eng.N, "Test feeds one shape of at most num_point points per call (:524-534)"
                idx = np.concatenate([np.arange(n0), np.random.choice(np.arange(n0), eng.N - n0, True)]).astype(np.int64)  # (:531-533)
                pend.append((data[0, idx, :], seg[0, idx], int(label[0, 0]), seg[0], n0))
            if not pend:
                break
            n = len(pend)
            pad = [pend[-1]] * (EB - n)
            data_feed = np.stack([q[0] for q in pend + pad]).astype(np.float32)
            seg_feed = np.stack([q[1] for q in pend + pad])
            label_feed = Tool.OnehotEncode(np.array([q[2] for q in pend + pad]), Loader.NUM_CATEGORIES, np.float32)
            mask = np.zeros((EB, eng.N), np.float32)
            mask[:n] = 1                                         # padding clouds carry no loss
            X = self._to_device('X', data_feed)
            Lb = self._to_device('Label', label_feed)
            Y = self._to_device('Y', Tool.OnehotEncode(seg_feed, 50, np.float32))
            M = self._to_device('Mask', mask)
            eng.forward(X, Lb, False, None)
            eng.losses_and_grad(Y, M, full=False, want_grad=False)
            Xn = eng.X[:n].contiguous()
            _, Yp, _, info = ops.lp_blocks(Xn, Xn, eng.Zp[:n].contiguous(), 1.0, 1.0)            # (:551-552)
            Yp_h = Yp.cpu().numpy()
            loss_mb = float(self._fetch_losses()[0])
            conv, its, res = (info[k_].cpu().numpy() for k_ in ("converged", "iters", "resid"))
            self.test_lp_info.append((data_cnt, its, res))
            if not conv.all():
                print('warning: label propagation stopped at {} iterations for {} shape(s) (relative residual up to '
                      '{:.1e})'.format(int(its.max()), int((conv == 0).sum()), float(res.max())))
            for j, (_, _, shape_label, seg0, n0) in enumerate(pend):
                Z_prob = Yp_h[j, :n0]
                iou_oids = Loader.object2setofoid[Loader.objcats[shape_label]]
                pred = self._restrict_to_category(Z_prob, iou_oids)
                avg_iou = Eval.EvalIoU(pred, seg0, iou_oids)
                perdata_miou = (perdata_miou * data_cnt + avg_iou) / (data_cnt + 1)
                pershape_miou[shape_label] = (pershape_miou[shape_label] * shape_cnt[shape_label] + avg_iou) / \
                    (shape_cnt[shape_label] + 1)
                avg_acc = (avg_acc * data_cnt + float(np.mean(pred == seg0))) / (data_cnt + 1)
                avg_loss = (avg_loss * data_cnt + loss_mb) / (data_cnt + 1)
                data_cnt += 1
                shape_cnt[shape_label] += 1
        return avg_loss, avg_acc, perdata_miou, pershape_miou

