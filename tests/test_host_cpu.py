"""CPU: host-side logic of the drop-in layer — ABI surface, schedules, synthetic loaders' tensor contract,
vectorised host helpers, data-parallel sharding over gloo (world_size 2)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shared_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "wspc.h")).read()
    names = sorted(set(re.findall(r"\b(wspc_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 25
    lib = ctypes.CDLL(os.path.join(ROOT, "weaksuppointcloudseg_b200", "libwspc.so"))
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    lib.wspc_version.restype = ctypes.c_int
    assert lib.wspc_version() >= 100
    # workspace queries are pure host functions (no GPU needed)
    lib.wspc_knn_workspace_bytes.restype = ctypes.c_size_t
    assert lib.wspc_knn_workspace_bytes(128, 4096, 64) >= 128 * 4096 * 64 * 4


def test_compute_entry_points_fail_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from weaksuppointcloudseg_b200 import ops
    from weaksuppointcloudseg_b200._lib import WspcError
    with pytest.raises(WspcError):
        ops.knn_fused(torch.zeros((1, 64, 3)), 8)          # CPU tensor: no fallback path


def test_schedules_match_reference_formulas():
    from oracle import dgcnn as od
    from weaksuppointcloudseg_b200.S3DIS_DGCNN_trainer import S3DIS_Trainer

    class FakeVS:
        step = 0

    class FakeEngine:
        vs = FakeVS()
    tr = S3DIS_Trainer.__new__(S3DIS_Trainer)
    tr.SetLearningRate(1e-3, 14)
    tr.engine = FakeEngine()
    for step in (0, 1, 21428, 21429, 50000, 400000, 5_000_000):
        tr.engine.vs.step = step
        assert tr.get_learning_rate() == pytest.approx(od.learning_rate(step, 1e-3, 14, 300000))
        assert tr.get_bn_decay() == pytest.approx(od.bn_decay(step, 14, 300000))
    tr.engine.vs.step = 10 ** 9
    assert tr.get_learning_rate() == 1e-5 and tr.get_bn_decay() == 0.99     # clips (:43, :53)


def test_synthetic_tensor_contract():
    from weaksuppointcloudseg_b200 import synthetic as syn
    X, Y, M, seg = syn.s3dis_batch(3, N=512, n_labelled=40, seed=0)
    assert X.shape == (6, 512, 9) and Y.shape == (6, 512, 13) and M.shape == (6, 512)
    assert np.all(M.sum(1) == 40) and np.array_equal(M[0::2], M[1::2]) and np.array_equal(seg[0::2], seg[1::2])
    assert np.all(Y.sum(-1) == 1) and np.array_equal(Y.argmax(-1), seg)
    # duplicated points exist (short-block padding)
    assert len(np.unique(X[0], axis=0)) < 512
    Xs, lab, Ys, Ms, segs = syn.shapenet_batch(2, N=256, n_labelled=26, seed=1)
    assert Xs.shape == (4, 256, 3) and lab.shape == (4, 16) and Ys.shape == (4, 256, 50)
    assert np.all(np.abs(np.linalg.norm(Xs[0::2], axis=-1).max(1) - 1) < 1e-5)      # pc_normalize
    for b in range(4):
        lo, hi = syn.CAT_PART_RANGES[int(lab[b].argmax())]
        assert segs[b].min() >= lo and segs[b].max() < hi


def test_tool_helpers_match_reference_loops():
    from weaksuppointcloudseg_b200 import Tool
    rng = np.random.default_rng(0)
    Yl = rng.integers(0, 13, (3, 50))
    ref = np.zeros((3, 50, 13))
    for b in range(3):
        for r in range(50):
            ref[b, r, Yl[b, r]] = 1                      # Util/Tool.py:14-17
    assert np.array_equal(Tool.OnehotEncode(Yl, 13), ref)
    pred = rng.integers(0, 4, (2, 100))
    gt = rng.integers(0, 4, (2, 100))
    iou = Tool.IoU(pred, gt, 4)
    for b in range(2):
        for k in range(4):
            inter = np.sum((pred[b] == k) & (gt[b] == k))
            union = np.sum((pred[b] == k) | (gt[b] == k))
            assert iou[b, k] == pytest.approx(inter / (union + 1e-6))


def test_shard_pairs():
    from weaksuppointcloudseg_b200.parallel import shard_pairs
    assert [shard_pairs(512, r, 8) for r in (0, 7)] == [(0, 64), (448, 512)]
    with pytest.raises(ValueError):
        shard_pairs(10, 0, 4)


_WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import torch
from weaksuppointcloudseg_b200.parallel import DataParallel, shard_pairs
dp = DataParallel(backend="gloo")
g = torch.full((1000,), float(dp.rank + 1))
dp.all_reduce(g)
assert torch.all(g == 3.0), g[:3]
assert dp.max_over_ranks(float(dp.rank)) == 1.0 and dp.sum_over_ranks(1.0) == 2.0
flat = torch.arange(10.0) * (dp.rank + 1)
dp.broadcast_params(flat)
assert torch.equal(flat, torch.arange(10.0))
lo, hi = shard_pairs(8, dp.rank, dp.world_size)
assert (lo, hi) == (4 * dp.rank, 4 * dp.rank + 4)
dp.barrier(); dp.shutdown()
print("rank", dp.rank, "ok")
'''


def test_data_parallel_gloo_world_size_2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", str(script)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.count("ok") == 2


def test_evaluation_iou_matches_reference_loop():
    """Evaluation.Eval.EvalIoU (Util/Evaluation.py:13-36): per-part IoU averaged over the category's part ids, absent
    parts count as 1 -- checked against the reference's scalar loop restated inline."""
    from weaksuppointcloudseg_b200.Evaluation import Eval
    rng = np.random.default_rng(0)
    for oids in ([0, 1, 2, 3], [12, 13, 14, 15], [47, 48, 49]):
        pred = rng.choice(oids + [5], 300)
        gt = rng.choice(oids[:-1], 300)          # the last part never occurs in the ground truth
        total = 0.0
        for oid in oids:
            n_pred, n_gt = np.sum(pred == oid), np.sum(gt == oid)
            n_int = np.sum((gt == oid) & (pred == gt))
            n_uni = n_pred + n_gt - n_int
            total += 1.0 if n_uni == 0 else n_int / n_uni
        assert abs(Eval().EvalIoU(pred, gt, oids) - total / len(oids)) < 1e-12


def test_trainer_minibatch_helpers_match_reference_loops():
    """The vectorised pieces of the epoch loops against the reference's per-sample / per-point python
    (S3DIS_DGCNN_trainer.py:246-252 mask, :265-296 augmentation cases, :473-477 class counters)."""
    from weaksuppointcloudseg_b200.S3DIS_DGCNN_trainer import S3DIS_Trainer as T

    rng = np.random.default_rng(3)
    # labelled-point mask
    pts_idx_list = np.stack([rng.choice(32, 5, replace=False) for _ in range(10)])
    data_idx = np.array([7, 2, 9])
    mask = T._mask_from_idx(pts_idx_list, data_idx, 3, 32)
    for b in range(3):
        assert sorted(np.flatnonzero(mask[b])) == sorted(pts_idx_list[data_idx[b]])
    assert T._mask_from_idx(None, data_idx, 3, 32).sum() == 0
    # the eight augmentation cases: (swap xy, mirror x, mirror y); normalised channels follow (swap / 1 - v)
    blk = rng.random((1, 16, 9)).astype(np.float32)

    def expect(choice):
        d = blk[0].copy()
        swap = choice in (1, 5, 6, 7)
        mx = choice in (2, 4, 5, 7)
        my = choice in (3, 4, 6, 7)
        if swap:
            d[:, 0], d[:, 1] = blk[0][:, 1], blk[0][:, 0]
            d[:, 6], d[:, 7] = blk[0][:, 7], blk[0][:, 6]
        if mx:
            d[:, 0] = -d[:, 0]
            d[:, 6] = -d[:, 6] + 1
        if my:
            d[:, 1] = -d[:, 1]
            d[:, 7] = -d[:, 7] + 1
        return d

    seen = set()
    for seed in range(40):
        np.random.seed(seed)
        choice = int(np.random.choice([0, 1, 2, 3, 4, 5, 6, 7], 1)[0])
        np.random.seed(seed)                                       # the helper draws with the same call
        got = T._augment(blk.copy())[0]
        assert np.allclose(got, expect(choice), atol=1e-7), choice
        assert np.array_equal(got[:, [2, 3, 4, 5, 8]], blk[0][:, [2, 3, 4, 5, 8]])
        seen.add(choice)
    assert seen == set(range(8))
    # class counters
    pred, gt = rng.integers(0, 13, (3, 50)), rng.integers(0, 13, (3, 50))
    pos, tp, cnt = np.zeros(13), np.zeros(13), np.zeros(13)
    T._count_classes(pred, gt, 13, pos, tp, cnt)
    pos_r, tp_r, cnt_r = np.zeros(13), np.zeros(13), np.zeros(13)
    for p_row, l_row in zip(pred, gt):
        for i in range(50):
            pos_r[p_row[i]] += 1
            tp_r[l_row[i]] += float(p_row[i] == l_row[i])
            cnt_r[l_row[i]] += 1
    assert np.array_equal(pos, pos_r) and np.array_equal(tp, tp_r) and np.array_equal(cnt, cnt_r)


def test_tool_numpy_helpers():
    from weaksuppointcloudseg_b200 import Tool
    rng = np.random.default_rng(4)
    X = rng.standard_normal((20, 3))
    D = Tool.pdist_np(X)
    ref = np.sqrt(((X[:, None] - X[None]) ** 2).sum(-1))
    assert np.allclose(D, ref, atol=1e-6) and np.all(D >= 0)
    v = rng.random(7) + 0.1
    assert np.isclose(np.abs(Tool.L1NormVec(v)).sum(), 1.0)
    assert np.allclose(Tool.L2NormVec(v), np.sqrt(v / np.sum(v ** 2)))
    for target, replace in ((12, False), (33, True), (20, None)):
        np.random.seed(9)
        pts, idx = Tool.ResamplePointCloud(X, target)
        assert pts.shape == (target, 3) and np.array_equal(pts, X[idx])
        if replace is False:
            assert len(set(idx.tolist())) == target
        if replace is not None:
            np.random.seed(9)
            assert np.array_equal(idx, np.random.choice(np.arange(0, 20), target, replace))
