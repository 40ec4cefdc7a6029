"""Dev tool: run the fused EdgeConv kernels alone at a cfg-3-like shape (for ncu captures / CUDA-event timings).
usage: python tools/prof_edgeconv.py [B] [N] [k] [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from weaksuppointcloudseg_b200 import _lib as L  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
k = int(sys.argv[3]) if len(sys.argv) > 3 else 20
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
dev = torch.device("cuda:0")
P = B * N
g = torch.Generator(device="cpu").manual_seed(0)
idx = torch.randint(0, N, (B, N, k), generator=g, dtype=torch.int32).to(dev)
UV = torch.randn((P, 128), device=dev)
vec = lambda s=1.0, o=0.0: (torch.randn(64, device=dev) * s + o)   # noqa: E731
b1, sc1, sh1, b2, sc2, sh2 = vec(0.1), vec(0.1, 1.0), vec(0.3), vec(0.1), vec(0.1, 1.0), vec(0.3)
c1, c2, c3 = vec(0.1, 1.0), vec(1e-3), vec(1e-2)
W2 = torch.randn((64, 64), device=dev) * 0.2
dout = torch.randn((P, 64), device=dev)
stats = torch.zeros((2, 64), dtype=torch.float64, device=dev)
bst = torch.zeros((2, 64), dtype=torch.float64, device=dev)
MM, SS, TS, MS, DUV = (torch.zeros((P, 128), device=dev) for _ in range(5))
deg = torch.zeros((P,), device=dev)
out = torch.empty((P, 64), device=dev)
dW2 = torch.empty((64, 64), device=dev)
lib = L.lib()
ws = torch.empty(lib.wspc_edgeconv2_bwd_workspace_bytes(), dtype=torch.uint8, device=dev)


def timed(name, fn):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:28s} {e0.elapsed_time(e1) / reps:8.3f} ms")


S = L.stream()
timed("edge_gather_stats(all)", lambda: L.check(lib.wspc_edge_gather_stats(
    L.ptr(UV), 128, L.ptr(idx), L.ptr(b1), P, k, N, 64, L.ptr(stats), L.ptr(MM), L.ptr(SS), L.ptr(deg), S)))
timed("edge_gather_stats(stats)", lambda: L.check(lib.wspc_edge_gather_stats(
    L.ptr(UV), 128, L.ptr(idx), L.ptr(b1), P, k, N, 64, L.ptr(stats), None, None, None, S)))
timed("edgeconv2_fwd(stats)", lambda: L.check(lib.wspc_edgeconv2_fwd(
    L.ptr(UV), 128, L.ptr(idx), L.ptr(b1), L.ptr(sc1), L.ptr(sh1), L.ptr(W2), L.ptr(b2), P, k, N, 64, 64, L.ptr(stats),
    L.ptr(MM), S)))
timed("edgeconv2_fwd(infer)", lambda: L.check(lib.wspc_edgeconv2_fwd(
    L.ptr(UV), 128, L.ptr(idx), L.ptr(b1), L.ptr(sc1), L.ptr(sh1), L.ptr(W2), L.ptr(b2), P, k, N, 64, 64, None, L.ptr(MM), S)))
L.check(lib.wspc_maxk_from_extrema(L.ptr(MM), L.ptr(sc2), L.ptr(sh2), P, 64, L.ptr(out), 64, S))
timed("maxk_extrema_bwd_prep", lambda: L.check(lib.wspc_maxk_extrema_bwd_prep(
    L.ptr(MM), L.ptr(sc2), L.ptr(out), 64, L.ptr(dout), 64, P, 64, L.ptr(MS), L.ptr(bst), S)))
timed("edgeconv2_bwd", lambda: L.check(lib.wspc_edgeconv2_bwd(
    L.ptr(UV), 128, L.ptr(idx), L.ptr(b1), L.ptr(sc1), L.ptr(sh1), L.ptr(W2), L.ptr(b2), L.ptr(sc2), L.ptr(sh2), L.ptr(c1),
    L.ptr(c2), L.ptr(c3), L.ptr(MS), P, k, N, 64, 64, L.ptr(TS), L.ptr(dW2), L.ptr(ws), ws.numel(), S)))
timed("edge1_bwd", lambda: L.check(lib.wspc_edge1_bwd(
    L.ptr(UV), 128, L.ptr(idx), L.ptr(b1), L.ptr(sc1), L.ptr(sh1), L.ptr(out), 64, L.ptr(dout), 64, P, k, N, 64, L.ptr(TS), S)))
timed("edge_bwd_stats", lambda: L.check(lib.wspc_edge_bwd_stats(L.ptr(TS), L.ptr(UV), 128, L.ptr(b1), P, 64, L.ptr(bst), S)))
timed("edge_bwd_finalize", lambda: L.check(lib.wspc_edge_bwd_finalize(
    L.ptr(TS), L.ptr(SS), L.ptr(deg), L.ptr(UV), 128, L.ptr(b1), L.ptr(c1), L.ptr(c2), L.ptr(c3), P, k, 64, L.ptr(DUV), 128, S)))
