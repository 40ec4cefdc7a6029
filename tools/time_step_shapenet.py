"""Dev tool: per-phase CUDA-event timing of one ShapeNet part-seg train step (BASELINE cfg-2: 32 samples = 64 clouds, N=2048)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import dgcnn as od
from weaksuppointcloudseg_b200 import synthetic as syn
from weaksuppointcloudseg_b200.engine_shapenet import ShapeNetEngine
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 32
N = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
X, lab, Y, M, _ = syn.shapenet_batch(ns, N=N, n_labelled=204)
B = 2 * ns
eng = ShapeNetEngine(od.init_params(od.SHAPENET_LAYERS, shapenet=True), B, N)
Xd, Ld, Yd, Md = (torch.from_numpy(a).cuda() for a in (X, lab, Y, M))
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e
for it in range(4):
    e0 = ev(); eng.forward(Xd, Ld, True, 0.5); e1 = ev(); eng.losses_and_grad(Yd, Md, True, True); e2 = ev(); eng.backward(); e3 = ev(); eng.vs.adam_step(1e-3); e4 = ev()
    torch.cuda.synchronize()
    print(json.dumps(dict(model="ShapeNet", B=B, N=N, fwd_ms=e0.elapsed_time(e1), loss_ms=e1.elapsed_time(e2), bwd_ms=e2.elapsed_time(e3), total_ms=e0.elapsed_time(e4), clouds_per_s=B / e0.elapsed_time(e4) * 1e3)))
print("losses", eng.losses.cpu().numpy(), "mem GB", torch.cuda.max_memory_allocated() / 1e9)
