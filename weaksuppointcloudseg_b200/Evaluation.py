"""Drop-in for Util/Evaluation.py (reference :7-36): part-IoU of one shape, averaged over the part ids of its category
(an id absent from both prediction and ground truth counts as IoU 1).  Pure numpy -- host-side bookkeeping, not on the
device path."""
from __future__ import annotations

import numpy as np


class Eval():
    AP = 0

    def EvalIoU(self, pred, seg_gt, iou_oids):
        pred, seg_gt = np.asarray(pred), np.asarray(seg_gt)
        oids = np.asarray(list(iou_oids))
        p = pred[None, :] == oids[:, None]
        g = seg_gt[None, :] == oids[:, None]
        inter = (p & g).sum(1).astype(np.float64)
        union = (p | g).sum(1).astype(np.float64)
        return float(np.where(union == 0, 1.0, inter / np.maximum(union, 1)).mean())
