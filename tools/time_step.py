"""Dev tool: per-phase CUDA-event timing of one S3DIS train step."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import dgcnn as od
from weaksuppointcloudseg_b200 import synthetic as syn
from weaksuppointcloudseg_b200.engine_s3dis import S3DISEngine
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
X, Y, M, _ = syn.s3dis_batch(ns, N=N, n_labelled=40)
B = 2 * ns
eng = S3DISEngine(od.init_params(od.S3DIS_LAYERS), B, N)
Xd, Yd, Md = (torch.from_numpy(a).cuda() for a in (X, Y, M))
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e
for it in range(3):
    e0 = ev(); eng.forward(Xd, True, 0.5); e1 = ev(); eng.losses_and_grad(Yd, Md, True, True); e2 = ev(); eng.backward(); e3 = ev(); eng.vs.adam_step(1e-3); e4 = ev()
    torch.cuda.synchronize()
    print(json.dumps(dict(B=B, N=N, fwd_ms=e0.elapsed_time(e1), loss_ms=e1.elapsed_time(e2), bwd_ms=e2.elapsed_time(e3), adam_ms=e3.elapsed_time(e4), total_ms=e0.elapsed_time(e4), clouds_per_s=B / e0.elapsed_time(e4) * 1e3)))
print("losses", eng.losses.cpu().numpy(), "mem GB", torch.cuda.max_memory_allocated() / 1e9)
