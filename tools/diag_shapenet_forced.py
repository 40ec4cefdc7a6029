import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
import routing
from oracle import dgcnn as od
from weaksuppointcloudseg_b200 import runtime as rt, synthetic as syn
from weaksuppointcloudseg_b200.engine_shapenet import ShapeNetEngine
cuda = torch.device("cuda:0")
def rel(a, b): return float(np.abs(a - b).max() / np.abs(b).max())
ns, N = (int(sys.argv[1]) if len(sys.argv) > 1 else 2), (int(sys.argv[2]) if len(sys.argv) > 2 else 2048)
X, lab, Y, M, _ = syn.shapenet_batch(ns, N=N, n_labelled=204, seed=131)
B = 2 * ns
params = od.init_params(od.SHAPENET_LAYERS, seed=132, shapenet=True)
rng = np.random.default_rng(133)
params["transform_net1/transform_XYZ/weights"] = rng.normal(0, 0.02, (256, 9)).astype(np.float32)
params["transform_net1/transform_XYZ/biases"] = rng.normal(0, 0.05, (9,)).astype(np.float32)
masks = [np.floor(0.6 + rng.random((B, N, 256))).astype(np.float32) for _ in range(2)]
eng = ShapeNetEngine(params, B, N, device=cuda)
rt.ROUTING = {}
eng.train_step(*(torch.from_numpy(a).to(cuda) for a in (X, lab, Y, M)), lr=1e-3, bn_decay=od.bn_decay(0, ns, 16881 * 20),
               dropout_masks=[torch.from_numpy(m).to(cuda) for m in masks], apply=False)
torch.cuda.synchronize()
route = routing.export_shapenet(eng, rt.ROUTING)
rt.ROUTING = None
ov = {f"knn{i}": eng.idx[i].cpu().long() for i in range(4)}
sg = (eng.idxS.cpu().long(), torch.exp(-eng.dS.cpu().double() / 0.1))
for dt in (torch.float64, torch.float32):
    p = od.to_torch(params, dtype=dt)
    opt = od.AdamTF(p, od.trainable_names(p))
    rec = {}
    with od.forced_routing(route):
        ref = od.train_step_shapenet(p, opt, torch.from_numpy(X).to(dt), torch.from_numpy(lab).to(dt), torch.from_numpy(Y).to(dt),
                                     torch.from_numpy(M).to(dt), step=0, dropout_masks=[torch.from_numpy(m).to(dt) for m in masks],
                                     knn_override=ov, smooth_graph_=(sg[0], sg[1].to(dt)), rec=rec)
    if dt == torch.float64:
        ref64, rec64 = ref, rec
        T = eng.Tm.cpu().numpy().reshape(-1, 3, 3) + np.eye(3, dtype=np.float32)
        print("engine vs fp64: T", rel(T, rec["transform"].detach().numpy()), "Xt", rel(eng.Xt.cpu().numpy(), rec["pct"].detach().numpy()),
              "tmax-in (tconv2 pre)", rel(eng.yt2.cpu().numpy().reshape(-1, 128), rec["transform_net1/tconv2/pre"].detach().numpy().reshape(-1, 128)),
              "yt3", rel(eng.yt3.cpu().numpy(), rec["transform_net1/tconv3/pre"].detach().numpy().reshape(-1, 1024)),
              "yf1", rel(eng.yf1.cpu().numpy(), rec["transform_net1/tfc1/pre"].detach().numpy()),
              "yf2", rel(eng.yf2.cpu().numpy(), rec["transform_net1/tfc2/pre"].detach().numpy()))
        cat = eng.cat.cpu().numpy().reshape(B, N, 192)
        for i, nm in enumerate(("net_1", "net_2", "net_3")):
            print("  ", nm, rel(cat[..., 64 * i:64 * i + 64], rec[nm].detach().numpy()))
        print("   Z", rel(eng.Z.cpu().numpy(), ref["Z"].detach().numpy()))
    else:
        print("fp32 oracle vs fp64 oracle: T", rel(rec["transform"].detach().numpy(), rec64["transform"].detach().numpy()),
              "Z", rel(ref["Z"].detach().numpy(), ref64["Z"].detach().numpy()))
        w = {}
        for name, g in ref64["grads"].items():
            b = g.numpy(); a = ref["grads"][name].numpy().astype(np.float64)
            if np.abs(b).max() > 0: w[name] = rel(a, b)
        print("fp32 oracle vs fp64 oracle grads: tnet max", max(v for k, v in w.items() if "transform_net1" in k), "trunk max", max(v for k, v in w.items() if "transform_net1" not in k))
