"""Dev tool: repeats the un-forced ShapeNet / S3DIS step-vs-fixture comparisons and prints the spread of the errors
(run-to-run differences come from the summation order of the fp64 atomics of the BN statistics: last-bit changes of the
BN scale flip a few ReLU / arg-max decisions)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_golden_gpu as tg  # noqa: E402

rc = tg.rc
cuda = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
from weaksuppointcloudseg_b200.engine_shapenet import ShapeNetEngine  # noqa: E402

f = np.load(os.path.join(tg.G, "ref_shapenet_step.npz"))
params0 = rc.xavier_params(rc.SHAPENET_LAYERS, int(f["param_seed"][0]), tnet_seed=int(f["param_seed"][1]))
B, N = f["X"].shape[:2]
bs = int(f["batch_size"][0])
X, lab, M = (torch.from_numpy(f[k]).to(cuda) for k in ("X", "label", "Mask"))
Y = torch.from_numpy(f["Y"].astype(np.float32)).to(cuda)
gmax = max(np.abs(f[k]).max() for k in f.files if k.startswith("grad/"))
worst_all = []
for it in range(n):
    eng = ShapeNetEngine(params0, B, N, device=cuda)
    ov = {k: tg._i32(f[k], cuda) for k in ("knn1", "knn2", "knn3")}
    sg = tg._smooth_graph(X, tg._i32(f["knn_smooth"], cuda), cuda)
    masks = [tg._keep(f[k], cuda) for k in ("dropout_keep1", "dropout_keep2")]
    losses = eng.train_step(X, lab, Y, M, lr=1e-3, bn_decay=0.5, dropout_masks=masks, knn_override=ov, smooth_graph=sg, apply=False)
    torch.cuda.synchronize()
    Z = eng.Z.cpu().numpy()
    ez = np.abs(Z - f["Z"]).max() / np.abs(f["Z"]).max()
    got = eng.vs.grads()
    worst, wk, worst_t, wkt = 0.0, "", 0.0, ""
    for k in f.files:
        if not k.startswith("grad/"):
            continue
        a, b = rc.subsample(got[k[len("grad/"):]])[0].astype(np.float64), f[k].astype(np.float64)
        if np.abs(b).max() < 1e-6 * gmax:
            continue
        e = np.linalg.norm(a - b) / np.linalg.norm(b)
        if "transform_net1/" in k:
            if e > worst_t:
                worst_t, wkt = e, k
        elif e > worst:
            worst, wk = e, k
    print(f"run {it}: Z {ez:.2e}  grads {worst:.3e} ({wk})  tnet {worst_t:.3e} ({wkt})")
    worst_all.append((ez, worst, worst_t))
w = np.array(worst_all)
print("max over runs: Z %.2e grads %.3e tnet %.3e" % tuple(w.max(0)))

# ---- the S3DIS reference-code step (tests/test_golden_gpu.py::test_s3dis_step_against_reference_code, bound 3e-2) ----------
from weaksuppointcloudseg_b200.engine_s3dis import S3DISEngine  # noqa: E402

f = np.load(os.path.join(tg.G, "ref_s3dis_step.npz"))
params0 = rc.xavier_params(rc.S3DIS_LAYERS, int(f["param_seed"][0]))
B, N = f["X"].shape[:2]
X, M = (torch.from_numpy(f[k]).to(cuda) for k in ("X", "Mask"))
Y = torch.from_numpy(f["Y"].astype(np.float32)).to(cuda)
gmax = max(np.abs(f[k]).max() for k in f.files if k.startswith("grad/"))
ws = []
for it in range(n):
    eng = S3DISEngine(params0, B, N, device=cuda)
    ov = {k: tg._i32(f[k], cuda) for k in ("knn2", "knn3")}
    sg = tg._smooth_graph(X[:, :, 0:6], tg._i32(f["knn_smooth"], cuda), cuda)
    eng.train_step(X, Y, M, lr=1e-3, bn_decay=0.5, dropout_mask=tg._keep(f["dropout_keep"], cuda), knn_override=ov, smooth_graph=sg,
                   apply=False)
    torch.cuda.synchronize()
    got = eng.vs.grads()
    worst, wk = 0.0, ""
    for k in f.files:
        if not k.startswith("grad/"):
            continue
        a, b = rc.subsample(got[k[len("grad/"):]])[0].astype(np.float64), f[k].astype(np.float64)
        if np.abs(b).max() < 1e-6 * gmax:
            continue
        e = np.linalg.norm(a - b) / np.linalg.norm(b)
        if e > worst:
            worst, wk = e, k
    ws.append(worst)
print("S3DIS step vs reference-code fixture, L2 error of the worst tensor over %d runs: min %.3e median %.3e max %.3e" % (
    n, min(ws), float(np.median(ws)), max(ws)))
