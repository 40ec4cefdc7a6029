import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import dgcnn as od
from weaksuppointcloudseg_b200 import synthetic as syn
from weaksuppointcloudseg_b200.engine_shapenet import ShapeNetEngine
def rel(a,b):
    a,b=np.asarray(a,np.float64),np.asarray(b,np.float64); return np.abs(a-b).max()/max(np.abs(b).max(),1e-30), np.linalg.norm(a-b)/max(np.linalg.norm(b),1e-30)
n_samples,N=3,320
X,lab,Y,M,_=syn.shapenet_batch(n_samples,N=N,n_labelled=32,seed=21); B=2*n_samples
params=od.init_params(od.SHAPENET_LAYERS,seed=8,shapenet=True)
rng=np.random.default_rng(2)
params["transform_net1/transform_XYZ/weights"]=rng.normal(0,0.02,(256,9)).astype(np.float32)
params["transform_net1/transform_XYZ/biases"]=rng.normal(0,0.05,(9,)).astype(np.float32)
masks=[np.floor(0.6+rng.random((B,N,256))).astype(np.float32) for _ in range(2)]
res={}; ov=None; sg=None
for dt in (torch.float32,torch.float64):
    p=od.to_torch(params,dtype=dt); opt=od.AdamTF(p,od.trainable_names(p)); rec={}
    out=od.train_step_shapenet(p,opt,torch.from_numpy(X).to(dt),torch.from_numpy(lab).to(dt),torch.from_numpy(Y).to(dt),torch.from_numpy(M).to(dt),dropout_masks=[torch.from_numpy(m).to(dt) for m in masks],rec=rec,knn_override=ov,smooth_graph_=sg)
    if ov is None:
        ov={f"knn{i}":rec[f"knn{i}/idx"] for i in (0,1,2,3)}; sg=od.smooth_graph(torch.from_numpy(X))
    res[dt]=out
eng=ShapeNetEngine(params,B,N,device="cuda:0")
ovd={f"knn{i}":ov[f"knn{i}"].to(torch.int32).cuda() for i in (1,2,3)}
eng.train_step(torch.from_numpy(X).cuda(),torch.from_numpy(lab).cuda(),torch.from_numpy(Y).cuda(),torch.from_numpy(M).cuda(),lr=1e-3,bn_decay=0.5,dropout_masks=[torch.from_numpy(m).cuda() for m in masks],knn_override=ovd,apply=False)
torch.cuda.synchronize()
got=eng.vs.grads()
print("Z eng-vs-f64", rel(eng.Z.cpu().numpy(),res[torch.float64]["Z"].detach().numpy())[0], "orc32-vs-f64", rel(res[torch.float32]["Z"].detach().numpy(),res[torch.float64]["Z"].detach().numpy())[0])
for n in od.trainable_names(params):
    g64=res[torch.float64]["grads"][n].numpy(); g32=res[torch.float32]["grads"][n].numpy()
    if 'biases' in n and 'conv4' not in n and 'XYZ' not in n: continue
    print("%-38s eng-f64 L2 %.2e  orc32-f64 L2 %.2e   |g| %.1e"%(n,rel(got[n],g64)[1],rel(g32,g64)[1],np.abs(g64).max()))
