"""Times the test-time label-propagation stage per block: Laplacian, solve, CG iterations.  usage: time_lp.py [blocks] [N]"""
import sys
import os
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from weaksuppointcloudseg_b200 import ops, synthetic as syn  # noqa: E402

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 8
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
dev = torch.device("cuda:0")
X, _, _, _ = syn.s3dis_batch(nb, N=N, n_labelled=40, seed=5)
X = torch.from_numpy(X[0::2]).to(dev)
for sharp in (0.2, 2.0, 8.0):                     # how peaked the network's probabilities are (w = 1 - entropy)
    G = torch.softmax(sharp * torch.randn(nb, N, 13, device=dev, generator=torch.Generator(dev).manual_seed(1)), -1)
    for b in range(X.shape[0]):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        Lm = ops.laplacian_sym(X[b:b + 1, :, 0:3].contiguous(), X[b:b + 1, :, 3:6].contiguous())
        ev[1].record()
        t0 = time.perf_counter()
        Y, Yp, w = ops.lp_solve(Lm[0], G[b].contiguous(), 1.0, 1.0)
        ev[2].record()
        torch.cuda.synchronize()
        if b:
            print("sharp %.1f block %d: laplacian %.3f ms, solve %.3f ms (host %.3f ms), iters %d, mean w %.3f" % (
                sharp, b, ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), 1e3 * (time.perf_counter() - t0),
                int(ops.lp_solve.last_info['iters'][0]), float(w.mean())))
