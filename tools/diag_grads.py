import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import dgcnn as od
from weaksuppointcloudseg_b200 import synthetic as syn
from weaksuppointcloudseg_b200.engine_s3dis import S3DISEngine
def rel(a,b):
    a,b=np.asarray(a,np.float64),np.asarray(b,np.float64); return np.abs(a-b).max()/max(np.abs(b).max(),1e-30)
n_samples,N=3,384
X,Y,M,_=syn.s3dis_batch(n_samples,N=N,n_labelled=12,seed=77); B=2*n_samples
params=od.init_params(od.S3DIS_LAYERS,seed=5)
rng=np.random.default_rng(9)
for k in params:
    if k.endswith("gamma"): params[k]=rng.uniform(0.5,1.5,params[k].shape).astype(np.float32)
    if k.endswith("beta"): params[k]=rng.uniform(-0.2,0.2,params[k].shape).astype(np.float32)
mask=np.floor(0.7+rng.random((B,N,256))).astype(np.float32)
outs={}
for dt in (torch.float32, torch.float64):
    p=od.to_torch(params,dtype=dt); opt=od.AdamTF(p,od.trainable_names(p)); rec={}
    kov=None if dt==torch.float32 else {k:v for k,v in ov_cpu.items()}
    out=od.train_step_s3dis(p,opt,torch.from_numpy(X).to(dt),torch.from_numpy(Y).to(dt),torch.from_numpy(M).to(dt),step=0,dropout_mask=torch.from_numpy(mask).to(dt),rec=rec,knn_override=kov, smooth_graph_=None if dt==torch.float32 else sg)
    if dt==torch.float32:
        ov_cpu={f"knn{i}":rec[f"knn{i}/idx"] for i in (1,2,3)}
        sg=od.smooth_graph(torch.from_numpy(X[:,:,0:6]))
    outs[dt]=out
eng=S3DISEngine(params,B,N,device="cuda:0")
ov={f"knn{i}":ov_cpu[f"knn{i}"].to(torch.int32).cuda() for i in (2,3)}
eng.train_step(torch.from_numpy(X).cuda(),torch.from_numpy(Y).cuda(),torch.from_numpy(M).cuda(),lr=1e-3,bn_decay=od.bn_decay(0,n_samples,300000),dropout_mask=torch.from_numpy(mask).cuda(),knn_override=ov)
torch.cuda.synchronize()
got=eng.vs.grads()
print("logits: eng-vs-f64 %.2e  f32-vs-f64 %.2e"%(rel(eng.Z.cpu().numpy(),outs[torch.float64]["Z"].detach().numpy()),rel(outs[torch.float32]["Z"].detach().numpy(),outs[torch.float64]["Z"].detach().numpy())))
for name in got:
    g64=outs[torch.float64]["grads"][name].numpy(); g32=outs[torch.float32]["grads"][name].numpy()
    print("%-28s eng-vs-f64 %.2e   orc32-vs-f64 %.2e   eng-vs-orc32 %.2e  |g|max %.2e"%(name,rel(got[name],g64),rel(g32,g64),rel(got[name],g32),np.abs(g64).max()))
