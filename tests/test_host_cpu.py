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
# overlapped gradient exchange: the head (seg/*, adj_conv7) slice is the contiguous tail of the flat buffer and is reduced
# asynchronously while the rest is still being produced; both slices must end up summed over the ranks
import numpy as np
from collections import OrderedDict
from weaksuppointcloudseg_b200.runtime import VariableStore
names = ["adj_conv1/weights", "adj_conv1/biases", "adj_conv1/bn/beta", "adj_conv2/weights", "adj_conv7/weights",
         "adj_conv7/bn/gamma", "seg/conv1/weights", "seg/conv3/biases"]
params = OrderedDict((n, np.zeros((5, 3) if n.endswith("weights") else (3,), np.float32)) for n in names)
vs = VariableStore(params, "cpu")
tail = vs.tail_offset(("adj_conv7/", "seg/"))
assert tail == vs._toffs["adj_conv7/weights"][0] and 0 < tail < vs.grad.numel()
assert vs.tail_offset(("adj_conv2/",)) is None            # not a tail: no overlap is attempted
vs.grad[:] = float(dp.rank + 1)
work = dp.all_reduce_async(vs.grad[tail:])
vs.grad[:tail] *= 2.0                                      # "the rest of the backward pass"
dp.all_reduce(vs.grad[:tail])
work.wait()
assert torch.all(vs.grad[tail:] == 3.0) and torch.all(vs.grad[:tail] == 6.0)
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


def test_train_epoch_loop_feeds_with_and_without_prefetch(tmp_path, monkeypatch):
    """TrainOneEpoch_Full on the real S3DIS_IO loader with the device step replaced by a recorder: the worker-thread
    prefetch delivers exactly the feeds of the in-line loop (same numpy random stream), in order; fp32 one-hot feeds;
    interleaved [sample, partner] rows; an exception in the assembly surfaces in the caller."""
    import types
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import dataio_fab as fab
    from weaksuppointcloudseg_b200 import S3DIS_DGCNN_trainer as T
    from weaksuppointcloudseg_b200.DataIO_S3DIS import S3DIS_IO

    root = fab.make_s3dis(str(tmp_path / 's3dis'), n_point=32)
    pts_idx_list = np.stack([np.random.default_rng(i).choice(32, 4, replace=False) for i in range(13)])

    def run(prefetch):
        monkeypatch.setenv('WSPC_PREFETCH', '1' if prefetch else '0')
        ld = S3DIS_IO(root, 13, batchsize=2, NUM_POINT=32)
        ld.LoadS3DIS_AllData()
        ld.CreateDataSplit(5)
        tr = T.S3DIS_Trainer(5, device='cpu', seed=0)
        tr.engine = types.SimpleNamespace(B=4)
        tr.epoch, tr.rampup = 3, 0                                  # augmentation on
        feeds = []

        def fake_train_batch(data_feed, seg_onehot_feed, Mask_bin_feed):
            feeds.append((data_feed.copy(), seg_onehot_feed.copy(), Mask_bin_feed.copy()))
            z = np.zeros(seg_onehot_feed.shape, np.float32)
            z[..., 3] = 1                                           # predicts class 3 everywhere
            return 1.0 + len(feeds), 0.1, 0.2, 0.3, z

        tr.train_batch = fake_train_batch
        np.random.seed(11)
        ld.Shuffle_TrainSet()
        loss, acc = tr.TrainOneEpoch_Full(ld, pts_idx_list, 2)
        assert tr.epoch == 4 and ld.train_samp_ptr == 0
        return feeds, loss, acc

    f1, l1, a1 = run(True)
    f0, l0, a0 = run(False)
    assert len(f1) == len(f0) == 5 and l1 == l0 and a1 == a0
    for (d1, y1, m1), (d0, y0, m0) in zip(f1, f0):
        assert np.array_equal(d1, d0) and np.array_equal(y1, y0) and np.array_equal(m1, m0)
        assert y1.dtype == np.float32 and d1.dtype == np.float32 and d1.shape == (4, 32, 9) and y1.shape == (4, 32, 13)
        assert np.array_equal(y1[0::2], y1[1::2]) and np.array_equal(m1[0::2], m1[1::2]) and m1.sum() == 4 * 4
        assert np.array_equal(d1[0::2, :, 2:6], d1[1::2, :, 2:6])    # z and rgb are untouched by the augmentation
    assert any(not np.array_equal(d[0::2], d[1::2]) for d, _, _ in f1)

    def boom():
        yield 1
        raise RuntimeError("assembly failed")

    got = []
    with pytest.raises(RuntimeError, match="assembly failed"):
        for x in T.prefetched(boom(), enabled=True):
            got.append(x)
    assert got == [1]


def test_shapenet_train_epoch_loop_feeds(tmp_path, monkeypatch):
    """ShapeNet `_train_epoch` (Full and Plain) on the real ShapeNetIO with a recording device step: prefetched == in-line,
    the tail batch is skipped, feeds are fp32 with the Siamese rows interleaved."""
    import types
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import dataio_fab as fab
    from weaksuppointcloudseg_b200.DataIO_ShapeNet import ShapeNetIO
    from weaksuppointcloudseg_b200.ShapeNet_DGCNN_trainer import ShapeNet_Trainer

    root = fab.make_shapenet(str(tmp_path / 'shapenet'), n_point=24)
    pts_idx_list = np.stack([np.random.default_rng(i).choice(24, 3, replace=False) for i in range(11)])
    file_idx_list, data_idx_list = np.zeros(11, np.int64), np.arange(11)

    def run(prefetch, siamese):
        monkeypatch.setenv('WSPC_PREFETCH', '1' if prefetch else '0')
        ld = ShapeNetIO(root, batchsize=2)
        ld.LoadTrainValFiles()
        tr = ShapeNet_Trainer(device='cpu', seed=0)
        tr.engine = types.SimpleNamespace(B=4 if siamese else 2)
        tr.epoch, tr.rampup = 2, 0
        feeds = []

        def fake_train_batch(data_feed, label_onehot_feed, seg_onehot_feed, Mask_bin_feed):
            feeds.append(tuple(np.array(a) for a in (data_feed, label_onehot_feed, seg_onehot_feed, Mask_bin_feed)))
            return 2.0, 0., 0., 0., np.zeros(seg_onehot_feed.shape, np.float32)

        tr.train_batch = fake_train_batch
        np.random.seed(4)
        ld.Shuffle_TrainSet()
        fn = tr.TrainOneEpoch_Full if siamese else tr.TrainOneEpoch
        loss, acc = fn(ld, file_idx_list, data_idx_list, pts_idx_list)
        assert tr.epoch == 3 and loss == 2.0
        return feeds, acc

    for siamese in (True, False):
        f1, a1 = run(True, siamese)
        f0, a0 = run(False, siamese)
        rep = 2 if siamese else 1
        assert len(f1) == len(f0) == 5 and a1 == a0                  # 11 shapes, batches of 2, the tail of 1 is skipped
        for x1, x0 in zip(f1, f0):
            assert all(np.array_equal(p, q) and p.dtype == np.float32 for p, q in zip(x1, x0))
            d, lab, y, m = x1
            assert d.shape == (2 * rep, 24, 3) and lab.shape == (2 * rep, 16) and y.shape == (2 * rep, 24, 50)
            assert m.sum() == 2 * rep * 3
            if siamese:
                assert np.array_equal(y[0::2], y[1::2]) and np.array_equal(lab[0::2], lab[1::2])
                assert np.abs(np.abs(d[1::2]) - np.abs(d[0::2])).max() < 0.05     # jitter (+ optional mirror of axis 2)


def test_ctypes_signatures_match_the_header():
    """ABI drift guard: every prototype of include/wspc.h has ctypes argtypes of the same arity and scalar kinds
    (int / long long / float / double / size_t / uint64_t / pointer) in weaksuppointcloudseg_b200/_lib.py."""
    import ctypes as C
    from weaksuppointcloudseg_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    h = open(os.path.join(root, "include", "wspc.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    h = re.sub(r"//[^\n]*", "", h)
    protos = re.findall(r"\b(wspc_\w+)\s*\(([^;{]*?)\)\s*;", h)
    assert len(protos) >= 40
    lib = _lib.lib()

    def kind_of_c(param):
        p = " ".join(param.split())
        if "*" in p or p.startswith("wspc_stream_t"):
            return "ptr"
        for key, kind in (("long long", "i64"), ("uint64_t", "u64"), ("size_t", "size"), ("double", "f64"), ("float", "f32"),
                          ("int32_t", "i32"), ("int", "i32")):
            if re.search(r"\b%s\b" % key, p):
                return kind
        raise AssertionError("unclassified parameter: " + p)

    def kind_of_ctypes(t):
        if t in (C.c_void_p, C.c_char_p) or hasattr(t, "contents") or (isinstance(t, type) and issubclass(t, C._Pointer)):
            return "ptr"
        return {C.c_int: "i32", C.c_longlong: "i64", C.c_uint64: "u64", C.c_size_t: "size", C.c_double: "f64",
                C.c_float: "f32"}[t]

    same_width = {frozenset(("u64", "size"))}                      # c_uint64 and c_size_t are one ctypes class on LP64
    for name, params in protos:
        fn = getattr(lib, name)
        plist = [p for p in (q.strip() for q in params.split(",")) if p and p != "void"]
        if fn.argtypes is None:
            assert not plist, "%s takes %d parameters but has no argtypes" % (name, len(plist))
            continue
        assert len(fn.argtypes) == len(plist), (name, len(fn.argtypes), len(plist))
        for i, (p, t) in enumerate(zip(plist, fn.argtypes)):
            a, b = kind_of_c(p), kind_of_ctypes(t)
            assert a == b or frozenset((a, b)) in same_width, (name, i, p, t)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU oracle arm) on a tiny sample: one JSON line with the keys the driver reads."""
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-clouds", "2", "--points", "128"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "clouds/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["steps"] == 1 and line["n_gpus"] == 1 and line["data"] == "synthetic"
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"] == line["e2e"]["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"] and line["vs_baseline"] is None
