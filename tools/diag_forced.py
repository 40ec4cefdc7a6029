"""Dev tool: stage-by-stage comparison of the engine's forward pass with the routing-forced fp64 oracle."""
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
rng = np.random.default_rng(123)
mask = np.floor(0.7 + rng.random((B, N, 256))).astype(np.float32)
eng = S3DISEngine(params, B, N, device=cuda)
from weaksuppointcloudseg_b200 import runtime as rt
rt.ROUTING = {}
eng.train_step(torch.from_numpy(X).to(cuda), torch.from_numpy(Y).to(cuda), torch.from_numpy(M).to(cuda), lr=1e-3,
               bn_decay=od.bn_decay(0, ns, 300000), dropout_mask=torch.from_numpy(mask).to(cuda), apply=False)
torch.cuda.synchronize()
route = routing.export_s3dis(eng, rt.ROUTING)
rt.ROUTING = None
ov = {f"knn{i + 1}": eng.idx[i].cpu().long() for i in range(3)}
rel = lambda a, b: float((a.double() - b.double()).abs().max() / b.double().abs().max())
for forced in (False, True):
    p = od.to_torch(params, dtype=torch.float64)
    opt = od.AdamTF(p, od.trainable_names(p))
    rec = {}
    ctx = od.forced_routing(route if forced else None)
    with ctx:
        ref = od.train_step_s3dis(p, opt, torch.from_numpy(X).double(), torch.from_numpy(Y).double(), torch.from_numpy(M).double(),
                                  step=0, dropout_mask=torch.from_numpy(mask).double(), knn_override=ov, rec=rec)
    cat = eng.cat.cpu().view(B, N, 192)
    print("forced" if forced else "free  ", "net1 %.2e net2 %.2e net3 %.2e g %.2e ys1 %.2e ys2 %.2e Z %.2e" % (
        rel(cat[..., :64], rec["net_1"].detach()), rel(cat[..., 64:128], rec["net_2"].detach()),
        rel(cat[..., 128:], rec["net_3"].detach()), rel(eng.g.cpu(), rec["out_max"].detach()),
        rel(eng.ys1.cpu().view(B, N, -1), rec["seg/conv1/pre"].detach()), rel(eng.ys2.cpu().view(B, N, -1), rec["seg/conv2/pre"].detach()),
        rel(eng.Z.cpu(), ref["Z"].detach())))
    got = eng.vs.grads()
    gmax = max(float(g.abs().max()) for g in ref["grads"].values())
    errs = {n: rel(torch.from_numpy(got[n]), g) for n, g in ref["grads"].items() if float(g.abs().max()) > 1e-9 * gmax}
    print("   grads:", {k_: "%.1e" % v for k_, v in errs.items()})

raise SystemExit
# ---- finer's per-edge quantities against the free fp64 oracle run
p = od.to_torch(params, dtype=torch.float64)
opt = od.AdamTF(p, od.trainable_names(p))
rec = {}
od.train_step_s3dis(p, opt, torch.from_numpy(X).double(), torch.from_numpy(Y).double(), torch.from_numpy(M).double(),
                    step=0, dropout_mask=torch.from_numpy(mask).double(), knn_override=ov, rec=rec)
k, P = eng.k, eng.P
base = (torch.arange(B, device=cuda) * N).view(B, 1, 1)
for i, (s1, s2) in enumerate((("adj_conv1", "adj_conv2"), ("adj_conv3", "adj_conv4"))):
    l1, l2 = eng.layers[s1], eng.layers[s2]
    UV = eng.eb[i].UV
    u, v = UV[:, :64], UV[:, 64:]
    gi = (eng.idx[i].long() + base).reshape(P, k)
    y1 = (u + l1.b).unsqueeze(1) + v[gi]
    y1o = rec[s1 + "/pre"].detach().reshape(P, k, 64)
    print(s1, "y1 err", rel(y1.cpu(), y1o))
    m = y1o.reshape(-1, 64).mean(0); va = ((y1o.reshape(-1, 64) - m) ** 2).mean(0)
    sco = p[s1 + "/bn/gamma"].detach() * torch.rsqrt(va + 1e-3); sho = p[s1 + "/bn/beta"].detach() - m * sco
    print("   sc err", rel(l1.sc.cpu(), sco), "sh err", rel(l1.sh.cpu(), sho))
    t = l1.b * l1.sc + l1.sh
    pre1 = v[gi] * l1.sc + (u.unsqueeze(1) * l1.sc + t)
    pre1o = y1o * sco + sho
    print("   pre1 err", rel(pre1.cpu(), pre1o), "mask mismatch", float(((pre1.cpu() > 0) != (pre1o > 0)).double().mean()))
    y2o = rec[s2 + "/pre"].detach().reshape(P, k, 64)
    a1 = torch.relu(pre1).double()
    y2 = a1 @ l2.W.double() + l2.b.double()
    print("   y2 err", rel(y2.cpu(), y2o))
    m2 = y2o.reshape(-1, 64).mean(0); va2 = ((y2o.reshape(-1, 64) - m2) ** 2).mean(0)
    sc2o = p[s2 + "/bn/gamma"].detach() * torch.rsqrt(va2 + 1e-3); sh2o = p[s2 + "/bn/beta"].detach() - m2 * sc2o
    print("   sc2 err", rel(l2.sc.cpu(), sc2o), "sh2 err", rel(l2.sh.cpu(), sh2o))
    val = torch.relu(y2 * l2.sc.double() + l2.sh.double())
    out = eng.cat[:, 64 * i:64 * i + 64].double()
    print("   max val vs out", rel(val.max(1).values.cpu(), out.cpu()))
    w = routing._maxk_weights(val, out)
    print("   weight row sums min/max", float(w.sum(1).min()), float(w.sum(1).max()), "sum(w*val) vs out", rel((w * val).sum(1).cpu(), out.cpu()))
    wv = route[f"maxk/knn{i + 1}"].reshape(P, k, 64)
    print("   route vs w", float((wv - w.cpu()).abs().max()))

# ---- forced vs free oracle, stage by stage
recs = {}
for forced in (False, True):
    p = od.to_torch(params, dtype=torch.float64)
    opt = od.AdamTF(p, od.trainable_names(p))
    r = {}
    with od.forced_routing(route if forced else None):
        od.train_step_s3dis(p, opt, torch.from_numpy(X).double(), torch.from_numpy(Y).double(), torch.from_numpy(M).double(),
                            step=0, dropout_mask=torch.from_numpy(mask).double(), knn_override=ov, rec=r)
    recs[forced] = r
for key in ("adj_conv1/pre", "adj_conv2/pre", "net_1", "adj_conv3/pre", "adj_conv4/pre", "net_2", "adj_conv5/pre", "net_3"):
    print(key, "forced vs free", rel(recs[True][key].detach(), recs[False][key].detach()))
m = route["relu/adj_conv1"]
print("mask dtype", m.dtype, m.shape, "true frac", float(m.double().mean()))
w = route["maxk/knn1"]
print("w dtype", w.dtype, w.shape, "sum over k min/max", float(w.sum(2).min()), float(w.sum(2).max()))
y2 = recs[False]["adj_conv2/pre"].detach()
print("y2 shape", y2.shape)

# ---- the routed selection itself
p0 = od.to_torch(params, dtype=torch.float64)
y2 = recs[False]["adj_conv2/pre"].detach()
bn2 = od.batch_norm(y2, {k_: v.clone() for k_, v in p0.items()}, "adj_conv2", True, None, (0, 1, 2)).detach()
n1 = recs[False]["net_1"].detach()
print("amax relu(bn2) vs net_1 free", rel(torch.relu(bn2).amax(2), n1))
sel = (w * bn2).sum(2)
print("sum w*bn2 vs net_1 free", rel(sel, n1), " forced net_1 vs this", rel(recs[True]["net_1"].detach(), sel))
bad = ((sel - n1).abs() > 1e-3 * n1.abs().max())
print("bad fraction", float(bad.double().mean()), "of which engine out==0:", float((eng.cat[:, :64].cpu().view(B, N, 64)[bad] == 0).double().mean()) if bad.any() else None)
bi = bad.nonzero()[:5]
for b_, n_, c_ in bi.tolist():
    print(" ex", b_, n_, c_, "w", w[b_, n_, :, c_].tolist(), "bn2", [round(float(x), 5) for x in bn2[b_, n_, :, c_]], "engine out", float(eng.cat.view(B, N, 192)[b_, n_, c_]), "free", float(n1[b_, n_, c_]))
