import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import dgcnn as od
from weaksuppointcloudseg_b200 import synthetic as syn
from weaksuppointcloudseg_b200.engine_s3dis import S3DISEngine
def rel(a,b):
    a,b=a.double(),b.double(); return float((a-b).abs().max()/b.abs().max().clamp_min(1e-30))
n_samples,N=3,384
X,Y,M,_=syn.s3dis_batch(n_samples,N=N,n_labelled=12,seed=77); B=2*n_samples
params=od.init_params(od.S3DIS_LAYERS,seed=5)
rng=np.random.default_rng(9)
for k_ in params:
    if k_.endswith("gamma"): params[k_]=rng.uniform(0.5,1.5,params[k_].shape).astype(np.float32)
    if k_.endswith("beta"): params[k_]=rng.uniform(-0.2,0.2,params[k_].shape).astype(np.float32)
mask=np.floor(0.7+rng.random((B,N,256))).astype(np.float32)
eng=S3DISEngine(params,B,N,device="cuda:0")
eng.train_step(torch.from_numpy(X).cuda(),torch.from_numpy(Y).cuda(),torch.from_numpy(M).cuda(),lr=1e-3,bn_decay=0.5,dropout_mask=torch.from_numpy(mask).cuda(),apply=False)
torch.cuda.synchronize()
L=eng.layers; s1,s2,s3=L["seg/conv1"],L["seg/conv2"],L["seg/conv3"]
P=eng.P
dZ=eng.dZ.reshape(P,13).double(); W3=s3.W.double()
a2=torch.relu(eng.ys2.double()*s2.sc.double()+s2.sh.double())
dm=torch.from_numpy(mask).cuda().reshape(P,256).double()
G2=(dZ@W3.T)*dm/0.7*(a2>0)
print("Gs2 elem", rel(eng.Gs2,G2))
print("bstats2 sumG", rel(s2.bstats[0], G2.sum(0)), "sumGy", rel(s2.bstats[1],(G2*eng.ys2.double()).sum(0)))
print("dbeta2", rel(s2.dbeta, G2.sum(0)))
xh=(eng.ys2.double()-s2.mean.double())*s2.invstd.double()
dgam=(G2*xh).sum(0); print("dgamma2", rel(s2.dgamma,dgam))
R=P
dy2=(s2.gamma.double()*s2.invstd.double()/R)*(R*G2-G2.sum(0)-xh*dgam)
print("gamma view", s2.gamma[:4].cpu().numpy(), params["seg/conv2/bn/gamma"][:4], "sc/invstd", (s2.sc/s2.invstd)[:4].cpu().numpy())
dy2_eng=s2.c1.double()*eng.Gs2.double()+s2.c2.double()+s2.c3.double()*eng.ys2.double()
print("dy2", rel(dy2_eng,dy2))
a1=torch.relu(eng.ys1.double()*s1.sc.double()+s1.sh.double())
print("dW2", rel(s2.dW, a1.T@dy2), "db2 abs", float(s2.db.abs().max()), float(dy2.sum(0).abs().max()))
G1=(dy2@s2.W.double().T)*(a1>0)
print("Gs1 elem", rel(eng.Gs1,G1))
print("mean2", rel(s2.mean, eng.ys2.double().mean(0)), "var", rel(1/s2.invstd.double()**2-1e-3, eng.ys2.double().var(0,unbiased=False)))
