"""Times the batched test-time stage: Laplacians of B blocks, then the batched label-propagation solve.
usage: time_lp_blocks.py [blocks] [N] [sharpness]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from weaksuppointcloudseg_b200 import _lib as L, ops, synthetic as syn  # noqa: E402

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
sharp = float(sys.argv[3]) if len(sys.argv) > 3 else 2.0
dev = torch.device("cuda:0")
X, _, _, _ = syn.s3dis_batch(nb, N=N, n_labelled=40, seed=5)
X = torch.from_numpy(X[0::2]).to(dev)
xyz, rgb = X[:, :, 0:3].contiguous(), X[:, :, 3:6].contiguous()
G = torch.softmax(sharp * torch.randn(nb, N, 13, device=dev, generator=torch.Generator(dev).manual_seed(1)), -1)
for rep in range(3):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    buf = L.workspace(nb * N * N * 4 + nb * N * 4, dev, "lp_laplacian")
    Lm = buf[:nb * N * N * 4].view(torch.float32).view(nb, N, N)
    deg = buf[nb * N * N * 4:nb * N * N * 4 + nb * N * 4].view(torch.float32)
    L.check(L.lib().wspc_laplacian_sym(L.ptr(xyz), L.ptr(rgb), nb, N, 3, 3, 1e3, 1e1, L.ptr(deg), L.ptr(Lm), L.stream()))
    ev[1].record()
    Y, Yp, w, info = ops.lp_blocks_on_graph(Lm, G, 1.0, 1.0)
    ev[2].record()
    torch.cuda.synchronize()
    it = info["iters"].cpu()
    print("blocks %d N %d: laplacian %.3f ms (%.3f / block), solve %.3f ms (%.3f / block), iters min/mean/max %d/%.1f/%d, "
          "sum iters %d -> %.1f us per block-iteration, mean w %.3f" % (
              nb, N, ev[0].elapsed_time(ev[1]), ev[0].elapsed_time(ev[1]) / nb, ev[1].elapsed_time(ev[2]),
              ev[1].elapsed_time(ev[2]) / nb, int(it.min()), float(it.float().mean()), int(it.max()), int(it.sum()),
              1e3 * ev[1].elapsed_time(ev[2]) / float(it.sum()), float(w.mean())))
