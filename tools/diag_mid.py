import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import dgcnn as od
from weaksuppointcloudseg_b200 import synthetic as syn
from weaksuppointcloudseg_b200.engine_s3dis import S3DISEngine
def rel(a,b):
    a,b=a.double().cpu(),b.double().cpu(); return float((a-b).abs().max()/b.abs().max().clamp_min(1e-30))
n_samples,N=3,384
X,Y,M,_=syn.s3dis_batch(n_samples,N=N,n_labelled=12,seed=77); B=2*n_samples; P=B*N
params=od.init_params(od.S3DIS_LAYERS,seed=5)
rng=np.random.default_rng(9)
for k_ in params:
    if k_.endswith("gamma"): params[k_]=rng.uniform(0.5,1.5,params[k_].shape).astype(np.float32)
    if k_.endswith("beta"): params[k_]=rng.uniform(-0.2,0.2,params[k_].shape).astype(np.float32)
mask=np.floor(0.7+rng.random((B,N,256))).astype(np.float32)
dt=torch.float64
p=od.to_torch(params,dtype=dt); rec={}
Xt,Yt,Mt=[torch.from_numpy(a).to(dt) for a in (X,Y,M)]
Z=od.get_model_s3dis(p,Xt,True,bn_decay=0.5,dropout_mask=torch.from_numpy(mask).to(dt),rec=rec)
Ls=od.weak_sup_losses(Z,Xt[:,:,0:6],Yt,Mt,10.0)
names=["seg/conv2/pre","seg/conv1/pre","adj_conv7/pre","adj_conv5/pre"]
gr=torch.autograd.grad(Ls["loss"],[Z]+[rec[n] for n in names]+[rec["net_1"],rec["net_2"],rec["net_3"]],retain_graph=True)
eng=S3DISEngine(params,B,N,device="cuda:0")
ov={f"knn{i}":rec[f"knn{i}/idx"].to(torch.int32).cuda() for i in (2,3)}
eng.train_step(torch.from_numpy(X).cuda(),torch.from_numpy(Y).cuda(),torch.from_numpy(M).cuda(),lr=1e-3,bn_decay=0.5,dropout_mask=torch.from_numpy(mask).cuda(),knn_override=ov,apply=False)
torch.cuda.synchronize()
L=eng.layers
print("Z", rel(eng.Z, Z.detach()), "dZ", rel(eng.dZ, gr[0]))
print("Zp", rel(eng.Zp, Ls["Z_prob"].detach()))
def dy(G,y,l): return l.c1.double()*G.double()+l.c2.double()+l.c3.double()*y.double()
print("ys2", rel(eng.ys2, rec["seg/conv2/pre"].detach().reshape(P,-1)), "dy2", rel(dy(eng.Gs2,eng.ys2,L["seg/conv2"]), gr[1].reshape(P,-1)))
print("ys1", rel(eng.ys1, rec["seg/conv1/pre"].detach().reshape(P,-1)), "dy1", rel(dy(eng.Gs1,eng.ys1,L["seg/conv1"]), gr[2].reshape(P,-1)))
print("y7", rel(eng.y7, rec["adj_conv7/pre"].detach().reshape(P,-1)))
dcat_ref=torch.cat([gr[5],gr[6],gr[7]],-1).reshape(P,192)
print("y5", rel(eng.y[4], rec["adj_conv5/pre"].detach().reshape(-1,64)))
gw=torch.autograd.grad(Ls["loss"],[p[n] for n in od.trainable_names(p)],retain_graph=True)
got=eng.vs.grads()
for n,g in zip(od.trainable_names(p),gw):
    if 'seg/' in n or 'conv7' in n: print(n, rel(torch.from_numpy(got[n]),g))
# per-term dZ
for nm in ["loss_seg","loss_siamese","loss_inexact","loss_smooth"]:
    g,=torch.autograd.grad(Ls[nm],[Z],retain_graph=True); print(nm, float(Ls[nm]), "|dZ|max", float(g.abs().max()))
print("losses eng", eng.losses.cpu().numpy())
