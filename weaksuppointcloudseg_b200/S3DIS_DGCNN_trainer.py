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


def prefetched(gen, enabled=None):
    """Iterate `gen` one item ahead on a worker thread: the host-side assembly of mini-batch i+1 (tens of ms of numpy) runs
    while the device executes step i (the main thread then sits in a stream synchronise with the GIL released).  Items come
    out in order; an exception in the generator is re-raised at the consumer.  WSPC_PREFETCH=0 turns it off."""
    if enabled is None:
        enabled = os.environ.get('WSPC_PREFETCH', '1') != '0'
    if not enabled:
        yield from gen
        return
    import queue
    import threading
    q = queue.Queue(maxsize=1)
    end = object()
    stop = threading.Event()

    def hand_over(msg):
        while not stop.is_set():              # a consumer that went away (exception, early break) releases the worker
            try:
                q.put(msg, timeout=0.1)
                return True
            except queue.Full:
                pass
        return False

    def work():
        try:
            for item in gen:
                if not hand_over((None, item)):
                    return
            hand_over((None, end))
        except BaseException as exc:      # noqa: BLE001  (handed to the consumer)
            hand_over((exc, None))

    worker = threading.Thread(target=work, name='wspc-prefetch', daemon=True)
    worker.start()
    try:
        while True:
            exc, item = q.get()
            if exc is not None:
                raise exc
            if item is end:
                break
            yield item
    finally:
        stop.set()
        worker.join(timeout=5)


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
    def defineNetwork(self, batch_size, num_points, style='Full', rampup=101, params=None, k=None):
        '''
        define DGCNN network for incomplete labels as supervision
        Args:
            batch_size: batchsize for training network (network clouds, i.e. 2x the CLI batch in Full style)
            num_points: number of points for each point cloud sample
            style: model style, use full model or plain model
            rampup: rampup epoch for training
            k: neighbours per point of the EdgeConv graphs (the reference hard-codes 20, DGCNN_S3DIS.py:30; BASELINE's
               stress configuration uses 40)
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
        self.engine = S3DISEngine(params, batch_size, num_points, device=self.device, **({} if k is None else {'k': int(k)}))
        if self.seed is not None:      # tf.nn.dropout draws from the graph-level seed: one mask sequence per trainer seed
            self.engine.seed = 1234 + 7919 * int(self.seed)
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

    def _to_device(self, name, arr, side=False, after_forward=False):
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
            if not after_forward:  # (after_forward: the only work in flight is this step's forward pass, which never touches
                                   #  the label buffers; their previous readers finished before the last step returned --
                                   #  every step ends with the stream synchronisation of _fetch_losses)
                cs.wait_stream(torch.cuda.current_stream())
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
        lr, decay = self.get_learning_rate(), self.get_bn_decay()
        full = self.style == 'Full'
        eng.forward(X, True, decay, dropout_mask)                  # enqueued; the device is busy from here on
        # only the loss block reads the labels: their pageable -> pinned staging (host memcpy, 29 MB at cfg-3) and the H2D
        # copy on the side stream both run under the forward pass
        Y = self._to_device('Y', seg_onehot_feed, side=True, after_forward=True)
        M = self._to_device('Mask', Mask_bin_feed, side=True, after_forward=True)
        torch.cuda.current_stream().wait_stream(self._copy_stream())
        gate_closed = full and not self.weak_gate
        # Full graph, gate closed: the weak terms are evaluated (and printed) but multiplied by 0 (:100-102) -- one head pass
        # that returns all four values and differentiates the segmentation term only
        eng.losses_and_grad(Y, M, full=2 if gate_closed else full, want_grad=True)
        zp = self._fetch_prob(side=True) if fetch_prob else None    # D2H of Z_prob overlaps the backward pass
        eng.backward()
        self._allreduce_and_step(lr)
        l = self._fetch_losses()          # synchronises the stream: the step is complete on return
        if fetch_prob:
            self._copy_stream().synchronize()
        if gate_closed:
            return float(l[0]), float(l[1]), float(l[2]), float(l[3]), zp
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
        """D2H of Z_prob into one of TWO pinned buffers used in turn: the array handed back by a step stays valid while the
        next step runs (the epoch loops read it there) and is overwritten by the step after that."""
        eng = self.engine
        if 'Zp' not in self.pinned:
            self.pinned['Zp'] = [torch.empty(tuple(eng.Zp.shape), dtype=torch.float32, pin_memory=True) for _ in range(2)]
            self._zp_turn = 0
        self._zp_turn ^= 1
        dst = self.pinned['Zp'][self._zp_turn]
        if side:      # Z_prob is final once the loss block has run; nothing in the backward pass writes it
            cs = self._copy_stream()
            cs.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(cs):
                dst.copy_(eng.Zp, non_blocking=True)
        else:
            dst.copy_(eng.Zp, non_blocking=True)
        return dst.numpy()

    def _allreduce_and_step(self, lr):
        gscale = 1.0
        if self.dist is not None:
            work = getattr(self, '_tail_work', None)
            if work is not None:      # the tail slice has been in flight since the head gradients were complete
                self.dist.all_reduce(self.engine.vs.grad[:self._tail_off])
                work.wait()
                self._tail_work = None
            else:
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

    # ------------------------------------------------------------------ test rooms (:499-584) ----
    @staticmethod
    def _count_classes(pred, gt, C, positive, true_positive, gt_count):
        """The reference's per-point python counters (:473-477, :552-555) as three bincounts."""
        pred, gt = np.asarray(pred).reshape(-1).astype(np.int64), np.asarray(gt).reshape(-1).astype(np.int64)
        positive += np.bincount(pred, minlength=C)[:C]
        gt_count += np.bincount(gt, minlength=C)[:C]
        true_positive += np.bincount(gt[pred == gt], minlength=C)[:C]

    def Test(self, Loader, PRED_PATH=None):
        """Inference + label propagation over the test rooms (Test, :499-584).  `Loader` is a DataIO_S3DIS.S3DIS_Test:
        `LoadNextTestRoomData_v1()` -> (blocks (nb,N,9), labels (nb,N), room_path), blocks None after the last room.
        The reference runs three `sess.run`s per block (:530-544); here the blocks of a room go through the network
        `engine.B` at a time (inference-mode batch norm uses the population statistics, so a block's logits do not depend on
        which blocks share its batch), and the Laplacians + label-propagation solves of those blocks run concurrently on the
        device (`ops.lp_blocks`, convergence decided on the device).  The LP-propagated prediction is scored; per room
        `<room>_pred_gt.mat` {data, pred, gt} is written to PRED_PATH (:571-580) when given.  Returns
        (true_positive_classes, positive_classes, gt_classes) like the reference; the network-only counters (no LP) are kept
        in `self.test_stats_net`, the LP solver's per-block iteration counts / residuals in `self.test_lp_info`."""
        from . import ops
        eng = self.engine
        C, EB = eng.C, eng.B
        true_positive_classes, positive_classes, gt_classes = np.zeros(C), np.zeros(C), np.zeros(C)
        net = [np.zeros(C), np.zeros(C), np.zeros(C)]
        total_correct = total_seen = 0.
        room_cnt = 0
        ones = torch.ones((EB, eng.N), device=self.device)
        Yz = torch.zeros((EB, eng.N, C), device=self.device)
        self.test_lp_info = []
        while True:
            data, label, room_path = Loader.LoadNextTestRoomData_v1()
            if data is None:
                break
            blocks, labels = np.asarray(data, np.float32), np.asarray(label).astype(np.int64)
            nb = blocks.shape[0]
            allPred = []
            for b0 in range(0, nb, EB):
                n = min(EB, nb - b0)
                chunk = blocks[b0:b0 + n]
                if n < EB:          # the graph batch is static: the tail of a room is padded with its last block
                    chunk = np.concatenate([chunk, np.repeat(chunk[-1:], EB - n, axis=0)], 0)
                X = self._to_device('X', np.ascontiguousarray(chunk))
                eng.forward(X, False, None)
                eng.losses_and_grad(Yz, ones, full=False, want_grad=False)
                G = eng.Zp[:n].contiguous()
                _, Yp, _, info = ops.lp_blocks(X[:n, :, 0:3], X[:n, :, 3:6], G, 1.0, 1.0)       # (:543-544)
                both = torch.stack([Yp.argmax(-1), G.argmax(-1)]).cpu().numpy()
                conv, its, res = (info[k_].cpu().numpy() for k_ in ("converged", "iters", "resid"))
                self.test_lp_info.append((room_cnt, b0, its, res))
                if not conv.all():
                    print('\nwarning: label propagation stopped at {} iterations for {} block(s) of room {} '
                          '(relative residual up to {:.1e})'.format(int(its.max()), int((conv == 0).sum()), room_cnt,
                                                                    float(res.max())))
                for j in range(n):
                    bi = b0 + j
                    pred = both[0][j]
                    self._count_classes(pred, labels[bi], C, positive_classes, true_positive_classes, gt_classes)
                    self._count_classes(both[1][j], labels[bi], C, net[1], net[0], net[2])
                    total_correct += float(np.sum(pred == labels[bi]))
                    total_seen += labels[bi].size
                    allPred.append(pred)
                iou = true_positive_classes / (gt_classes + positive_classes - true_positive_classes + 1e-5)
                print('\rroom {:d}  acc {:.2f}%  iou: {:.2f}%'.format(room_cnt, 100 * total_correct / total_seen,
                                                                      100 * np.mean(iou)), end='')
            if PRED_PATH is not None:
                import scipy.io as scio
                room_name = os.path.basename(room_path).split('.')[0]
                scio.savemat(os.path.join(PRED_PATH, '{}_pred_gt.mat'.format(room_name)),
                             {'data': data, 'pred': np.concatenate(allPred, 0), 'gt': labels.reshape(-1)})
            room_cnt += 1
        self.test_stats_net = tuple(net)
        return true_positive_classes, positive_classes, gt_classes

    # ------------------------------------------------------------------ epoch loops --------------
    @staticmethod
    def _mask_from_idx(pts_idx_list, data_idx, mb_size, N):
        """Mask_bin (mb,N): 1 at the labelled points of each sample, all zero without a list (:246-252)."""
        mask = np.zeros((mb_size, N), np.float32)
        if pts_idx_list is not None:
            for b_i in range(mb_size):
                mask[b_i, np.asarray(pts_idx_list[data_idx[b_i]]).reshape(-1).astype(np.int64)] = 1
        return mask

    # aug_choice -> (swap x/y, mirror x, mirror y), the eight cases of :265-296
    _AUG = ((0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (0, 1, 1), (1, 1, 0), (1, 0, 1), (1, 1, 1))

    @classmethod
    def _augment(cls, aug):
        """One of the 8 axis-swap / mirror augmentations per block, drawn with the reference's `np.random.choice` call.
        Channels 6:8 (room-normalised xy) follow: swapped with the axes, mirrored as 1 - v."""
        for b_i in range(aug.shape[0]):
            swap, mx, my = cls._AUG[int(np.random.choice([0, 1, 2, 3, 4, 5, 6, 7], 1)[0])]
            blk = aug[b_i]
            if swap:
                blk[:, [0, 1, 6, 7]] = blk[:, [1, 0, 7, 6]]
            if mx:
                blk[:, 0] = -blk[:, 0]
                blk[:, 6] = 1 - blk[:, 6]
            if my:
                blk[:, 1] = -blk[:, 1]
                blk[:, 7] = 1 - blk[:, 7]
        return aug

    def _end_train_epoch(self, Loader):
        if hasattr(Loader, 'ResetLoader_TrainSet'):
            Loader.ResetLoader_TrainSet()                # (:213, :343)
        self.epoch += 1

    def _plain_batches(self, Loader, pts_idx_list):
        while True:
            SuccessFlag, data, seg, weak_seg_onehot, mb_size, data_idx = Loader.NextBatch_TrainSet_v1()
            if not SuccessFlag or mb_size < self.engine.B:   # the short last batch is dropped (:167-168); static graph batch
                return
            data = np.asarray(data, np.float32)
            seg = np.asarray(seg).astype(np.int64)
            yield (data, Tool.OnehotEncode(seg, 13, np.float32), self._mask_from_idx(pts_idx_list, data_idx, mb_size, data.shape[1]),
                   seg, mb_size)

    def TrainOneEpoch(self, Loader, pts_idx_list=None, batch_size=12):
        """Plain-style epoch (TrainOneEpoch, :146-219): `batch_size` clouds per step, segmentation loss on the labelled
        points only.  -> (avg_loss, avg_acc)."""
        batch_cnt, data_cnt, avg_loss, avg_acc = 1, 0, 0., 0.
        for data, seg_onehot_feed, Mask_bin_feed, seg, mb_size in prefetched(self._plain_batches(Loader, pts_idx_list)):
            loss_mb, _, _, _, Z_prob_mb = self.train_batch(data, seg_onehot_feed, Mask_bin_feed)
            acc = float(np.mean(np.argmax(Z_prob_mb, axis=-1) == seg))
            avg_loss = (avg_loss * data_cnt + loss_mb * mb_size) / (data_cnt + mb_size)
            avg_acc = (avg_acc * data_cnt + acc * mb_size) / (data_cnt + mb_size)
            data_cnt += mb_size
            print('\rBatch {:d} TrainedSamp {:d}  mbLoss {:.4f} Avg Acc {:.2f}%'.format(batch_cnt, data_cnt, loss_mb,
                                                                                      100 * avg_acc), end='')
            batch_cnt += 1
        self._end_train_epoch(Loader)
        return avg_loss, avg_acc

    def _full_batches(self, Loader, pts_idx_list):
        """Generator of the Full-style feeds of one epoch: (data_feed (2b,N,9), seg_onehot_feed (2b,N,13) fp32,
        Mask_bin_feed (2b,N), seg (b,N), mb_size), rows interleaved [sample, Siamese partner] (:237-315)."""
        half = self.engine.B // 2
        while True:
            SuccessFlag, data, seg, weak_seg_onehot, mb_size, data_idx = Loader.NextBatch_TrainSet_v1()
            if not SuccessFlag or mb_size < half:        # short last batch dropped (:242-243); the graph batch is static
                return
            data = np.asarray(data, np.float32)
            seg = np.asarray(seg).astype(np.int64)
            N = data.shape[1]
            mask = self._mask_from_idx(pts_idx_list, data_idx, mb_size, N)
            data_feed = np.empty((2 * mb_size, N, 9), np.float32)
            data_feed[0::2] = data
            # The partner is an augmented COPY once the ramp-up epoch is reached (:261); the reference augments in place
            # through an alias, so both rows of its pair end up augmented — consciously not kept.
            data_feed[1::2] = self._augment(data.copy()) if self.epoch >= self.rampup else data
            seg_onehot_feed = Tool.OnehotEncode(np.repeat(seg, 2, axis=0), 13, np.float32)
            yield data_feed, seg_onehot_feed, np.repeat(mask, 2, axis=0), seg, mb_size

    def TrainOneEpoch_Full(self, Loader, pts_idx_list=None, batch_size=12):
        '''
        Function to train one epoch (TrainOneEpoch_Full, :221-349).  `Loader` follows S3DIS_IO.NextBatch_TrainSet_v1
        (DataIO_S3DIS.py:127-154); the mini-batch assembly (mask from pts_idx_list, Siamese partner, interleaving, one-hot)
        is vectorised instead of the per-point loops and runs one batch ahead on a worker thread, under the device step of
        the previous batch (`prefetched`).  `batch_size` = samples per step (half the graph's clouds).
        '''
        batch_cnt, data_cnt, avg_loss, avg_acc = 1, 0, 0., 0.
        for data_feed, seg_onehot_feed, Mask_bin_feed, seg, mb_size in prefetched(self._full_batches(Loader, pts_idx_list)):
            loss_mb, loss_siam, loss_inex, loss_smooth, Z_prob_mb = self.train_batch(data_feed, seg_onehot_feed,
                                                                                    Mask_bin_feed)
            pred = np.argmax(Z_prob_mb[0::2], axis=-1)   # (:326-339)
            acc = float(np.mean(pred == seg))
            avg_loss = (avg_loss * data_cnt + loss_mb * mb_size) / (data_cnt + mb_size)
            avg_acc = (avg_acc * data_cnt + acc * mb_size) / (data_cnt + mb_size)
            data_cnt += mb_size
            print('\rBatch {:d} TrainedSamp {:d}  mbLoss {:.4f} SiamLoss {:.3f} MILLoss {:.3f} SmoothLoss {:.3f} '
                  'Avg Acc {:.2f}%'.format(batch_cnt, data_cnt, loss_mb, loss_siam, loss_inex, loss_smooth, 100 * avg_acc),
                  end='')
            batch_cnt += 1
        self._end_train_epoch(Loader)
        return avg_loss, avg_acc

    def _eval_epoch(self, Loader, siamese):
        """Shared validation loop over `Loader.NextBatch_TestSet()` (the held-out area).  A short last batch is padded by
        repeating its first block (:429-436); `siamese` duplicates every block to fill the 2B graph (:445-455) and scores
        Z_prob[0::2].  Running averages follow the reference's formulas, including its loss average that adds the batch
        loss unweighted (:488).  -> (avg_loss, avg_correct_rate, mean IoU over the 13 classes)."""
        C = 13
        per_step = self.engine.B // 2 if siamese else self.engine.B
        true_positive, positive, gt_count = np.zeros(C), np.zeros(C), np.zeros(C)
        batch_cnt, samp_cnt, avg_loss, avg_correct_rate = 1, 0, 0., 0.
        iou = np.zeros(C)
        while True:
            SuccessFlag, data, seg_mb, weak_seg_onehot, mb_size = Loader.NextBatch_TestSet()[:5]
            if not SuccessFlag:
                break
            data = np.asarray(data, np.float32)
            seg_mb = np.asarray(seg_mb).astype(np.int64)
            if mb_size > per_step:
                raise ValueError("loader batch %d exceeds the graph's %d samples per step" % (mb_size, per_step))
            pad = per_step - mb_size
            data_feed = np.concatenate([data, np.repeat(data[0:1], pad, 0)], 0) if pad else data
            seg_feed = np.concatenate([seg_mb, np.repeat(seg_mb[0:1], pad, 0)], 0) if pad else seg_mb
            if siamese:
                data_feed, seg_feed = np.repeat(data_feed, 2, axis=0), np.repeat(seg_feed, 2, axis=0)
            loss_mb, Z_prob_mb = self.eval_batch(data_feed, Tool.OnehotEncode(seg_feed, C, np.float32),
                                                 np.ones(seg_feed.shape, np.float32))
            Z_prob_mb = Z_prob_mb[0:2 * mb_size:2] if siamese else Z_prob_mb[0:mb_size]
            pred_mb = np.argmax(Z_prob_mb, axis=-1)
            acc = float(np.mean(pred_mb == seg_mb))
            self._count_classes(pred_mb, seg_mb, C, positive, true_positive, gt_count)
            iou = true_positive / (gt_count + positive - true_positive + 1e-5)
            avg_loss = (avg_loss * samp_cnt + loss_mb) / (samp_cnt + mb_size)
            avg_correct_rate = (avg_correct_rate * samp_cnt + acc * mb_size) / (samp_cnt + mb_size)
            samp_cnt += mb_size
            print('\rBatch {:d} EvaluatedSamp {:d}  Avg Loss {:.4f}  Avg Correct Rate {:.3f}%  mIoU {:.3f}%'.format(
                batch_cnt, samp_cnt, avg_loss, 100 * avg_correct_rate, 100 * np.mean(iou)), end='')
            batch_cnt += 1
        if hasattr(Loader, 'ResetLoader_TestSet'):
            Loader.ResetLoader_TestSet()
        self.eval_iou = iou                              # per-class IoU of the pass
        return avg_loss, avg_correct_rate, float(np.mean(iou))

    def EvalOneEpoch(self, Loader):
        """Plain-style validation (EvalOneEpoch, :351-399).  The reference's version feeds integer labels to the one-hot
        placeholder and Is_Training=True; here it is the Full loop without the Siamese duplication (inference mode)."""
        return self._eval_epoch(Loader, siamese=False)

    def EvalOneEpoch_Full(self, Loader):
        '''Validation pass (EvalOneEpoch_Full, :401-497) -> (avg_loss, avg_correct_rate, mIoU).'''
        return self._eval_epoch(Loader, siamese=True)

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
                np.mean(eval_avg_correct_rate) > self.bestValCorrect:
            self.bestValCorrect = np.mean(eval_avg_correct_rate)
            # the best copy lives next to the checkpoint, under `best_filename` (:592-601)
            np.savez(os.path.join(os.path.dirname(os.path.abspath(save_filepath)), best_filename + '.npz'), **blob)

    def RestoreCheckPoint(self, filepath, weights_only=False):
        """`<filepath>.npz` written by SaveCheckPoint, or -- when `<filepath>.index` / `.data-00000-of-00001` exist -- a
        checkpoint written by the reference's tf.train.Saver (same variable names; tf_checkpoint.py).  A TF checkpoint
        without global step / Adam slots is refused unless `weights_only` (the schedules would silently restart)."""
        from . import tf_checkpoint
        vs = self.engine.vs
        if not os.path.exists(filepath if filepath.endswith('.npz') else filepath + '.npz') and tf_checkpoint.exists(filepath):
            shapes = {k: tuple(vs.get(k).shape) for k in vs.trainable_names + vs.state_names}
            blob = tf_checkpoint.to_store_blob(tf_checkpoint.read(filepath), vs.trainable_names, vs.state_names, shapes,
                                               allow_missing_optimizer_state=weights_only)
        else:
            blob = np.load(filepath if filepath.endswith('.npz') else filepath + '.npz')
        vs.load({k: blob[k] for k in vs.trainable_names + vs.state_names})
        vs.step = int(blob['Variable'])
        vs.adam_m.copy_(torch.from_numpy(blob['__adam_m']))
        vs.adam_v.copy_(torch.from_numpy(blob['__adam_v']))
