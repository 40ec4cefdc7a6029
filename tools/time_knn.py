"""Dev tool: CUDA-event timing of the fused kNN kernel at benchmark shapes."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from weaksuppointcloudseg_b200 import ops

def bench(B, N, D, k, flavour, iters=5, coff=None):
    if coff is not None:   # S3DIS-like block (9 channels, far from the origin in the normalised-xyz window)
        from weaksuppointcloudseg_b200 import synthetic as syn
        x = torch.from_numpy(syn.s3dis_batch(max(B // 2, 1), N=N, n_labelled=8)[0]).cuda()
        _f = ops.knn_fused
        ops_knn = lambda x_, k_, fl_: _f(x_, k_, fl_, coff=coff, D=D)
    else:
        x = torch.relu(torch.randn((B, N, D), device="cuda")) if D >= 16 else torch.rand((B, N, D), device="cuda")
        ops_knn = ops.knn_fused
    for _ in range(2):
        ops_knn(x, k, flavour)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        ops_knn(x, k, flavour)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))
    ms = ts[len(ts) // 2]
    eq_bytes = B * (2 * N * N * 4 + N * D * 4 + N * k * 4)
    flops = B * (2.0 * N * N * D)
    import ctypes
    from weaksuppointcloudseg_b200 import _lib as L
    fb = ctypes.c_int(0)
    ws = L.workspace(1, x.device, "knn")
    L.check(L.lib().wspc_knn_fallback_rows(L.ptr(ws), B, N, D, ctypes.byref(fb)))
    print(json.dumps(dict(B=B, N=N, D=D, k=k, flavour=flavour, s3dis_window=coff, fallback_rows=fb.value, ms=round(ms, 3), equiv_GBs=round(eq_bytes / ms / 1e6, 1),
                          fma_TFLOPs=round(flops / ms / 1e9, 2))))

if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    if len(sys.argv) > 2 and sys.argv[2] == "ncu":      # two shapes only, for ncu captures (3 launches each: 2 warm-up + 1)
        bench(B, 4096, 64, 20, 0, iters=1)
        bench(max(B // 8, 1), 8192, 64, 40, 0, iters=1)
        sys.exit(0)
    bench(B, 4096, 3, 20, 0)
    bench(B, 4096, 6, 10, 1)
    bench(B, 4096, 3, 20, 0, coff=6)
    bench(B, 4096, 6, 10, 1, coff=0)
    bench(B, 4096, 64, 20, 0)
    bench(max(B // 8, 1), 8192, 64, 40, 0)
    bench(B, 2048, 64, 20, 0)
