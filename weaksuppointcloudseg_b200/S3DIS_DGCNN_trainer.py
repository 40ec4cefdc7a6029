"""Drop-in for the reference trainer class `S3DIS_Trainer` (S3DIS/S3DIS_DGCNN_trainer.py:17-629).

Same method names, argument meaning and attribute names (`X_ph`-style placeholders collapse to
arguments; `sess` is None).  The TF graph/session protocol is replaced by the fused CUDA executor
`S3DISEngine`: one `train_batch` call == one `sess.run([solver, loss, loss_siamese, loss_inexact,
loss_smooth, Z_prob], feed_dict)` of TrainOneEpoch_Full (:317-323).
"""
from __future__ import annotations

import math
import os
import sys

import numpy as np
import torch

from . import Tool
from .engine_s3dis import S3DISEngine


def xavier_params(layers, seed=None, shapenet=False):
    """Variable initialisation of the reference graph: Xavier-uniform weights, zero biases, gamma=1,
    beta=0, pop_mean=0, pop_var=1 (tf_util.py:43-47,160-161,513-519)."""
    from collections import OrderedDict
    rng = np.random.default_rng(seed)
    p = OrderedDict()
    for scope, cin, cout, has_bn in layers:
        lim = math.sqrt(6.0 / (cin + cout))
        p[f"{scope}/weights"] = rng.uniform(-lim, lim, (cin, cout)).astype(np.float32)
        p[f"{scope}/biases"] = np.zeros((cout,), np.float32)
        if has_bn:
            p[f"{scope}/bn/beta"] = np.zeros((cout,), np.float32)
            p[f"{scope}/bn/gamma"] = np.ones((cout,), np.float32)
            p[f"{scope}/bn/pop_mean"] = np.zeros((cout,), np.float32)
            p[f"{scope}/bn/pop_var"] = np.ones((cout,), np.float32)
    if shapenet:
        p["transform_net1/transform_XYZ/weights"] = np.zeros((256, 9), np.float32)
        p["transform_net1/transform_XYZ/biases"] = np.zeros((9,), np.float32)
    return p


class S3DIS_Trainer():

    def __init__(self, test_area=5, device=None, seed=None):
        self.bestValCorrect = 0.    # initial best validation performance      (:21)
        self.test_area = test_area
        self.device = torch.device(device if device is not None else "cuda:0")
        self.seed = seed
        self.sess = None            # no session: kept for signature compatibility
        self.dist = None            # (rank, world, all_reduce fn) set by parallel.attach()

    # ------------------------------------------------------------------ schedules (:25-54) -------
    def SetLearningRate(self, LearningRate, BatchSize):
        self.BASE_LEARNING_RATE = LearningRate
        self.BATCH_SIZE = BatchSize
        self.BN_INIT_DECAY = 0.5
        self.BN_DECAY_DECAY_RATE = 0.5
        self.DECAY_STEP = 300000
        self.DECAY_RATE = 0.5
        self.BN_DECAY_DECAY_STEP = float(self.DECAY_STEP * 2)
        self.BN_DECAY_CLIP = 0.99

    def get_learning_rate(self):
        step = self.engine.vs.step   # == self.batch, incremented by the optimiser (:110)
        lr = self.BASE_LEARNING_RATE * self.DECAY_RATE ** math.floor(step * self.BATCH_SIZE / self.DECAY_STEP)
        return max(lr, 1e-5)  # CLIP THE LEARNING RATE (:43)

    def get_bn_decay(self):
        step = self.engine.vs.step
        bn_momentum = self.BN_INIT_DECAY * self.BN_DECAY_DECAY_RATE ** math.floor(
            step * self.BATCH_SIZE / self.BN_DECAY_DECAY_STEP)
        return min(self.BN_DECAY_CLIP, 1 - bn_momentum)

    @property
    def batch(self):
        return self.engine.vs.step

    # ------------------------------------------------------------------ graph (:56-118) ----------
    def defineNetwork(self, batch_size, num_points, style='Full', rampup=101, params=None):
        '''
        define DGCNN network for incomplete labels as supervision
        Args:
            batch_size: batchsize for training network (network clouds, i.e. 2x the CLI batch in Full style)
            num_points: number of points for each point cloud sample
            style: model style, use full model or plain model
            rampup: rampup epoch for training
        '''
        self.rampup = rampup
        self.style = style
        if style not in ('Plain', 'Full'):
            sys.exit('Loss {} is not defined!'.format(style))     # (:104)
        if not hasattr(self, 'BATCH_SIZE'):
            self.SetLearningRate(1e-3, max(batch_size // 2, 1))
        if params is None:
            from .engine_s3dis import LAYERS
            params = xavier_params(LAYERS, self.seed)
        self.engine = S3DISEngine(params, batch_size, num_points, device=self.device)
        self.epoch = 0
        # The reference evaluates `epoch >= rampup` once, at graph-build time (:93,:101): the gate is a
        # constant of the graph.  Same here.
        self.weak_gate = (style == 'Full') and (self.epoch >= self.rampup)
        self.pinned = {}
        return True

    # ------------------------------------------------------------------ one sess.run -------------
    def _copy_stream(self):
        """side stream for the feed/fetch copies that are not on the critical path (labels in, probabilities out)"""
        if getattr(self, '_cstream', None) is None:
            self._cstream = torch.cuda.Stream(device=self.device)
        return self._cstream

    def _to_device(self, name, arr, side=False):
        """feed_dict H2D: host arrays go through a persistent pinned staging buffer.  side=True issues the copy on the
        copy stream (overlapping the forward pass); the caller makes the compute stream wait before the first use."""
        if torch.is_tensor(arr) and arr.is_cuda:
            return arr.contiguous()
        t = torch.as_tensor(arr)
        if t.dtype != torch.float32:
            t = t.to(torch.float32)
        stage = self.pinned.get(name)
        if stage is None or stage[0].shape != t.shape:
            stage = (torch.empty(t.shape, dtype=torch.float32, pin_memory=True),
                     torch.empty(t.shape, dtype=torch.float32, device=self.device))
            self.pinned[name] = stage
        if not (t.is_pinned() and t.is_contiguous()):
            stage[0].copy_(t)
            t = stage[0]
        if side:
            cs = self._copy_stream()
            cs.wait_stream(torch.cuda.current_stream())      # the previous step's readers of this buffer are done
            with torch.cuda.stream(cs):
                stage[1].copy_(t, non_blocking=True)
        else:
            stage[1].copy_(t, non_blocking=True)
        return stage[1]

    def train_batch(self, data_feed, seg_onehot_feed, Mask_bin_feed, fetch_prob=True, dropout_mask=None):
        """== sess.run([solver, loss, loss_siamese, loss_inexact, loss_smooth, Z_prob], feed_dict) with
        Is_Training_ph=True (:317-323).  Returns (loss, loss_siamese, loss_inexact, loss_smooth, Z_prob)."""
        eng = self.engine
        X = self._to_device('X', data_feed)
        Y = self._to_device('Y', seg_onehot_feed, side=True)       # only the loss block reads the labels: the 27 MB copy
        M = self._to_device('Mask', Mask_bin_feed, side=True)      # overlaps the forward pass
        lr, decay = self.get_learning_rate(), self.get_bn_decay()
        full = self.style == 'Full'
        eng.forward(X, True, decay, dropout_mask)
        torch.cuda.current_stream().wait_stream(self._copy_stream())
        if full and not self.weak_gate:
            # Full graph, gate closed: the weak terms are evaluated (and printed) but multiplied by 0 (:100-102)
            eng.losses_and_grad(Y, M, full=True, want_grad=False)
            weak = self._fetch_losses()
            eng.losses_and_grad(Y, M, full=False, want_grad=True)
        else:
            weak = None
            eng.losses_and_grad(Y, M, full=full, want_grad=True)
        zp = self._fetch_prob(side=True) if fetch_prob else None    # D2H of Z_prob overlaps the backward pass
        eng.backward()
        self._allreduce_and_step(lr)
        l = self._fetch_losses()          # synchronises the stream: the step is complete on return
        if fetch_prob:
            self._copy_stream().synchronize()
        if weak is not None:
            return float(l[0]), float(weak[1]), float(weak[2]), float(weak[3]), zp
        return float(l[4]), float(l[1]), float(l[2]), float(l[3]), zp

    def _fetch_losses(self):
        """D2H of the five loss scalars through pinned memory (the sess.run fetch list)."""
        if 'losses' not in self.pinned:
            self.pinned['losses'] = torch.empty(5, dtype=torch.float32, pin_memory=True)
        h = self.pinned['losses']
        h.copy_(self.engine.losses, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return h.numpy().copy()

    def _fetch_prob(self, side=False):
        eng = self.engine
        if 'Zp' not in self.pinned:
            self.pinned['Zp'] = torch.empty(tuple(eng.Zp.shape), dtype=torch.float32, pin_memory=True)
        if side:      # Z_prob is final once the loss block has run; nothing in the backward pass writes it
            cs = self._copy_stream()
            cs.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(cs):
                self.pinned['Zp'].copy_(eng.Zp, non_blocking=True)
        else:
            self.pinned['Zp'].copy_(eng.Zp, non_blocking=True)
        return self.pinned['Zp'].numpy()

    def _allreduce_and_step(self, lr):
        gscale = 1.0
        if self.dist is not None:
            self.dist.all_reduce(self.engine.vs.grad)
            gscale = 1.0 / self.dist.world_size
        self.engine.vs.adam_step(lr, gscale=gscale)

    def eval_batch(self, data_feed, seg_onehot_feed, Mask_bin_feed):
        """== sess.run([loss, Z_prob], Is_Training_ph=False) (:457-462, :538-540)."""
        eng = self.engine
        X = self._to_device('X', data_feed)
        Y = self._to_device('Y', seg_onehot_feed)
        M = self._to_device('Mask', Mask_bin_feed)
        eng.forward(X, False, None)
        full = self.style == 'Full' and eng.B % 2 == 0
        eng.losses_and_grad(Y, M, full=full, want_grad=False)
        zp = self._fetch_prob()
        l = self._fetch_losses()
        loss = float(l[4]) if (full and self.weak_gate) else float(l[0])
        return loss, zp.copy()

    def defLabelPropSolver(self, alpha=1e0, beta=1e0, K=10):
        """(:139-143) Define Label Propagation Solver; the reference ignores the arguments too"""
        from . import ProbLabelPropagation as PLP
        self.LPSolver = PLP.LabelPropagation_TF(alpha=1e0, beta=1e0, K=10)
        self.TFComp = {}
        self.TFComp['Lmat'] = Tool.TF_Computation.LaplacianMatSym_XYZRGB_DirectComp()

    def Test(self, Loader, PRED_PATH=None):
        """Inference + label propagation over test rooms (Test, :499-584): per block one forward pass
        (Is_Training=False), the symmetric Laplacian of the block (xyz, rgb) and the closed-form LP solve —
        all on the device, without the reference's 64 MB D2H/H2D round trip per block.  `Loader` follows
        S3DIS_Test.LoadNextTestRoomData_v1 (DataIO_S3DIS.py:288-299): returns (flag, blocks (nb,4096,9),
        labels (nb,4096)).  Returns overall accuracy and per-class IoU with and without LP."""
        from . import ops
        eng = self.engine
        C = eng.C
        stat = {k: dict(tp=np.zeros(C), fp=np.zeros(C), fn=np.zeros(C)) for k in ('net', 'lp')}
        while True:
            out = Loader.LoadNextTestRoomData_v1()
            if not out[0]:
                break
            blocks, labels = np.asarray(out[1], np.float32), np.asarray(out[2]).astype(np.int64)
            for bi in range(blocks.shape[0]):
                X = torch.from_numpy(blocks[bi:bi + 1]).to(self.device)
                Xf = X.expand(eng.B, -1, -1).contiguous()              # graph batch is static, like the reference
                eng.forward(Xf, False, None)
                Yz = torch.zeros((eng.B, eng.N, C), device=self.device)
                eng.losses_and_grad(Yz, torch.ones((eng.B, eng.N), device=self.device), full=False, want_grad=False)
                G = eng.Zp[0].contiguous()
                Lm = ops.laplacian_sym(X[:, :, 0:3].contiguous(), X[:, :, 3:6].contiguous())   # (:543)
                _, Yp, _ = ops.lp_solve(Lm[0], G, 1.0, 1.0)                                   # (:544)
                for key, prob in (('net', G), ('lp', Yp)):
                    pred = prob.argmax(-1).cpu().numpy()
                    gt = labels[bi]
                    for c in range(C):                                   # vectorised (:552-555)
                        stat[key]['tp'][c] += np.sum((pred == c) & (gt == c))
                        stat[key]['fp'][c] += np.sum((pred == c) & (gt != c))
                        stat[key]['fn'][c] += np.sum((pred != c) & (gt == c))
        res = {}
        for key, s in stat.items():
            iou = s['tp'] / np.maximum(s['tp'] + s['fp'] + s['fn'], 1)
            res[key] = dict(acc=float(s['tp'].sum() / max((s['tp'] + s['fn']).sum(), 1)), iou=iou, miou=float(iou.mean()))
        return res

    # ------------------------------------------------------------------ epoch loops --------------
    def TrainOneEpoch_Full(self, Loader, pts_idx_list):
        '''
        Function to train one epoch (TrainOneEpoch_Full, :221-349).  `Loader` follows the contract of
        S3DIS_IO.NextBatch_TrainSet_v1 (DataIO_S3DIS.py:127-154); the mini-batch assembly (mask from
        pts_idx_list, Siamese partner, interleaving, one-hot) is vectorised instead of the per-point loops.
        '''
        batch_cnt = 1
        data_cnt = 0
        avg_loss = 0.
        avg_acc = 0.
        B = self.engine.B
        rng = np.random.default_rng(self.epoch)
        while True:
            SuccessFlag, data, seg, weak_seg_onehot, mb_size, data_idx = Loader.NextBatch_TrainSet_v1()
            if not SuccessFlag or mb_size < B // 2:      # short last batch is dropped (:242-243)
                break
            data = np.asarray(data, np.float32)
            N = data.shape[1]
            mask = np.zeros((mb_size, N), np.float32)
            for b_i in range(mb_size):                   # (:246-252)
                mask[b_i, np.asarray(pts_idx_list[data_idx[b_i]]).reshape(-1).astype(np.int64)] = 1
            aug = data.copy()                            # a true copy (the reference aliases, App. C-2)
            if self.epoch >= self.rampup:                # host-side augmentation switch (:261)
                for b_i in range(mb_size):
                    mode = rng.integers(0, 8)
                    if mode & 1:
                        aug[b_i][:, [0, 1]] = aug[b_i][:, [1, 0]]
                        aug[b_i][:, [6, 7]] = aug[b_i][:, [7, 6]]
                    if mode & 2:
                        aug[b_i][:, 0] *= -1
                    if mode & 4:
                        aug[b_i][:, 1] *= -1
            data_feed = np.empty((2 * mb_size, N, 9), np.float32)
            data_feed[0::2], data_feed[1::2] = data, aug
            seg_feed = np.repeat(np.asarray(seg).astype(np.int64), 2, axis=0)
            seg_onehot_feed = Tool.OnehotEncode(seg_feed, 13)
            Mask_bin_feed = np.repeat(mask, 2, axis=0)
            loss_mb, loss_siam, loss_inex, loss_smooth, Z_prob_mb = self.train_batch(data_feed, seg_onehot_feed,
                                                                                    Mask_bin_feed)
            pred = np.argmax(Z_prob_mb[0::2], axis=-1)   # (:326-339)
            acc = float(np.mean(pred == np.asarray(seg)))
            avg_loss = (avg_loss * data_cnt + loss_mb * mb_size) / (data_cnt + mb_size)
            avg_acc = (avg_acc * data_cnt + acc * mb_size) / (data_cnt + mb_size)
            data_cnt += mb_size
            print('\rBatch {}  loss {:.4f} siam {:.4f} inexact {:.4f} smooth {:.5f} acc {:.3f}'.format(
                batch_cnt, loss_mb, loss_siam, loss_inex, loss_smooth, acc), end='')
            batch_cnt += 1
        self.epoch += 1
        return avg_loss, avg_acc

    def EvalOneEpoch_Full(self, Loader, pts_idx_list=None):
        '''Validation pass (EvalOneEpoch_Full, :401-497): samples are duplicated to fill the 2B Siamese
        graph (:445-455), evaluated with Is_Training=False, and Z_prob[0::2] is scored.'''
        B = self.engine.B
        inter = np.zeros(13)
        union = np.zeros(13)
        correct = 0
        total = 0
        avg_loss, cnt = 0., 0
        while True:
            SuccessFlag, data, seg, weak_seg_onehot, mb_size = Loader.NextBatch_ValSet()[:5]
            if not SuccessFlag:
                break
            data = np.asarray(data, np.float32)
            seg = np.asarray(seg).astype(np.int64)
            if mb_size < B // 2:                         # pad by repeating sample 0 (:429-436)
                pad = B // 2 - mb_size
                data = np.concatenate([data, np.repeat(data[0:1], pad, 0)], 0)
                seg = np.concatenate([seg, np.repeat(seg[0:1], pad, 0)], 0)
            data_feed = np.repeat(data, 2, axis=0)
            seg_feed = np.repeat(seg, 2, axis=0)
            N = data.shape[1]
            loss_mb, Z_prob_mb = self.eval_batch(data_feed, Tool.OnehotEncode(seg_feed, 13),
                                                 np.ones((2 * data.shape[0], N), np.float32))
            pred = np.argmax(Z_prob_mb[0:2 * mb_size:2], axis=-1)
            gt = seg[:mb_size]
            correct += int(np.sum(pred == gt))
            total += gt.size
            for c in range(13):                          # vectorised IoU counters (:473-477)
                inter[c] += np.sum((pred == c) & (gt == c))
                union[c] += np.sum((pred == c) | (gt == c))
            avg_loss = (avg_loss * cnt + loss_mb * mb_size) / (cnt + mb_size)
            cnt += mb_size
        iou = inter / np.maximum(union, 1)
        return avg_loss, correct / max(total, 1), float(np.mean(iou)), iou

    # ------------------------------------------------------------------ checkpoints (:586-629) ---
    def SaveCheckPoint(self, save_filepath, best_filename=None, eval_avg_correct_rate=None):
        """Variables are stored under their TF names (`<scope>/weights`, `<scope>/bn/pop_mean`, ...) plus
        `Variable` (global step) and the Adam slots, as `<save_filepath>.npz`."""
        os.makedirs(os.path.dirname(os.path.abspath(save_filepath)), exist_ok=True)
        vs = self.engine.vs
        blob = dict(vs.export())
        blob['Variable'] = np.asarray(vs.step, np.int64)
        blob['__adam_m'] = vs.adam_m.cpu().numpy()
        blob['__adam_v'] = vs.adam_v.cpu().numpy()
        np.savez(save_filepath + '.npz', **blob)
        if best_filename is not None and eval_avg_correct_rate is not None and \
                eval_avg_correct_rate > self.bestValCorrect:
            self.bestValCorrect = eval_avg_correct_rate
            np.savez(best_filename + '.npz', **blob)

    def RestoreCheckPoint(self, filepath):
        blob = np.load(filepath if filepath.endswith('.npz') else filepath + '.npz')
        vs = self.engine.vs
        vs.load({k: blob[k] for k in vs.trainable_names + vs.state_names})
        vs.step = int(blob['Variable'])
        vs.adam_m.copy_(torch.from_numpy(blob['__adam_m']))
        vs.adam_v.copy_(torch.from_numpy(blob['__adam_v']))
