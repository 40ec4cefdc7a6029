"""Dev tool: the fused EdgeConv backward kernel on the engine's own data (block 1 runs last, so the shared scratch still holds
its MS / TS after a step) against an fp64 torch evaluation of the same formulas."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import dgcnn as od
from weaksuppointcloudseg_b200 import synthetic as syn
from weaksuppointcloudseg_b200.engine_s3dis import S3DISEngine
import routing
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 1
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
cuda = torch.device("cuda:0")
X, Y, M, _ = syn.s3dis_batch(ns, N=N, n_labelled=40, seed=121)
B = 2 * ns
params = od.init_params(od.S3DIS_LAYERS, seed=122)
mask = np.floor(0.7 + np.random.default_rng(123).random((B, N, 256))).astype(np.float32)
eng = S3DISEngine(params, B, N, device=cuda)
eng.train_step(torch.from_numpy(X).to(cuda), torch.from_numpy(Y).to(cuda), torch.from_numpy(M).to(cuda), lr=1e-3,
               bn_decay=od.bn_decay(0, ns, 300000), dropout_mask=torch.from_numpy(mask).to(cuda), apply=False)
torch.cuda.synchronize()
rel = lambda a, b: float((a.double() - b.double()).abs().max() / b.double().abs().max())
k, P = eng.k, eng.P
l1, l2 = eng.layers["adj_conv1"], eng.layers["adj_conv2"]
UV = eng.eb[0].UV
u, v = UV[:, :64], UV[:, 64:]
base = (torch.arange(B, device=cuda) * N).view(B, 1, 1)
gi = (eng.idx[0].long() + base).reshape(P, k)
t = l1.b * l1.sc + l1.sh
pre1 = v[gi] * l1.sc + (u.unsqueeze(1) * l1.sc + t)
a1 = torch.relu(pre1).double()
y2 = a1 @ l2.W.double() + l2.b.double()
bn2 = y2 * l2.sc.double() + l2.sh.double()
out = eng.cat[:, 0:64].double()
dout = eng.dcat[:, 0:64].double()
MS = eng.ef.MS
print("MS[:, :64] vs (out>0?out:-1)", rel(MS[:, :64], torch.where(out > 0, out, -torch.ones_like(out))), " MS[:,64:] vs gated dout",
      rel(MS[:, 64:], torch.where(out > 0, dout, torch.zeros_like(dout))))
w = routing._maxk_weights(bn2, out)
G = w * MS[:, 64:].double().unsqueeze(1)
print("bstats sumG", rel(l2.bstats[0], G.sum((0, 1))), "sumGy", rel(l2.bstats[1], (G * y2).sum((0, 1))))
dy2 = l2.c1.double() * G + l2.c2.double() + l2.c3.double() * y2
dW2 = torch.einsum("pkc,pkd->cd", a1, dy2)
print("dW2 (adj_conv2/weights) err", rel(l2.dW, dW2), " |dW2| max", float(dW2.abs().max()))
# decomposition of dW2: sparse part and dense part
dWs = torch.einsum("pkc,pkd->cd", a1, l2.c1.double() * G)
dWd = torch.einsum("pkc,pkd->cd", a1, l2.c2.double() + l2.c3.double() * y2)
print("  sparse part max", float(dWs.abs().max()), "dense part max", float(dWd.abs().max()))
da1 = (dy2 @ l2.W.double().t()) * (a1 > 0)
SG = da1.sum(1)
TG = torch.zeros((P, 64), dtype=torch.float64, device=cuda)
TG.index_add_(0, gi.reshape(-1), da1.reshape(-1, 64))
TS = eng.ef.TS
print("TS SG err", rel(TS[:, :64], SG), " TS TG err", rel(TS[:, 64:], TG))
print("c1..c3 l2", float(l2.c1.abs().max()), float(l2.c2.abs().max()), float(l2.c3.abs().max()))
# BN-1 closed form
g1 = da1
y1 = (u + l1.b).unsqueeze(1).double() + v[gi].double()
print("bstats1 sum g", rel(l1.bstats[0], g1.sum((0, 1))), "sum g y1", rel(l1.bstats[1], (g1 * y1).sum((0, 1))))
dy1 = l1.c1.double() * g1 + l1.c2.double() + l1.c3.double() * y1
du = dy1.sum(1)
dv = torch.zeros((P, 64), dtype=torch.float64, device=cuda)
dv.index_add_(0, gi.reshape(-1), dy1.reshape(-1, 64))
DUV = eng.ef.DUV
print("DUV du err", rel(DUV[:, :64], du), " dv err", rel(DUV[:, 64:], dv))
