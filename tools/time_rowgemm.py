"""Times the per-point row GEMMs of a cfg-3 step (P = 128 clouds x 4096 points) through wspc_conv1x1_rows_ws.
WSPC_ROWGEMM_KERNEL=serial selects the single-role kernel for an A/B run.  Usage: python tools/time_rowgemm.py [clouds]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from weaksuppointcloudseg_b200 import _lib as L, runtime as rt  # noqa: E402


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    clouds = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    Np = 4096
    M = clouds * Np
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cuda").manual_seed(1)
    rows = []
    for name, K, N, kind in (("adj_conv7 fwd", 192, 1024, "plain"), ("seg/conv1 fwd", 192, 512, "plain_rb"),
                             ("seg/conv2 fwd", 512, 256, "bnrelu"), ("seg/conv2 dgrad", 256, 512, "dy_mask"),
                             ("seg/conv1 dgrad", 512, 192, "dy_store")):
        out = torch.empty((M, N), device=dev)
        stats = torch.zeros((2, N), dtype=torch.float64, device=dev)
        a = torch.randn((M, K), device=dev, generator=g)
        keep = [a, out, stats]
        if kind in ("plain", "plain_rb", "bnrelu"):
            W = torch.randn((K, N), device=dev, generator=g) * 0.1
            b = torch.randn(N, device=dev, generator=g)
            rb = torch.randn((clouds, N), device=dev, generator=g)
            sc = torch.rand(K, device=dev, generator=g) + 0.5
            sh = torch.randn(K, device=dev, generator=g) * 0.2
            keep += [W, b, rb, sc, sh]
            if kind == "bnrelu":
                A = (L.Operand(p=a.data_ptr(), ld=K, C=K, sc=sc.data_ptr(), sh=sh.data_ptr()), L.OP_BNRELU)
            else:
                A = (L.Operand(p=a.data_ptr(), ld=K, C=K), L.OP_PLAIN)
            epi = L.Epilogue(out=out.data_ptr(), ldo=N, bias=b.data_ptr(), rowbias=rb.data_ptr() if kind == "plain_rb" else 0,
                             rb_rows=Np, ldrb=N, stats=stats.data_ptr())
            fn = lambda A=A, W=W, N=N, K=K, epi=epi: rt.rows_gemm(A, W, N, 0, M, N, K, epi, L.EPI_STORE_STATS)  # noqa: E731
            nbytes = 4 * M * (K + N)
        else:
            y = torch.randn((M, K), device=dev, generator=g)
            c1, c2, c3 = (torch.randn(K, device=dev, generator=g) * 0.5 for _ in range(3))
            Wl = torch.randn((N, K), device=dev, generator=g) * 0.1
            A = (L.Operand(p=a.data_ptr(), ld=K, C=K, y=y.data_ptr(), ldy=K, c1=c1.data_ptr(), c2=c2.data_ptr(), c3=c3.data_ptr()),
                 L.OP_DY)
            keep += [y, c1, c2, c3, Wl]
            if kind == "dy_mask":
                yprev = torch.randn((M, N), device=dev, generator=g)
                scp = torch.rand(N, device=dev, generator=g) + 0.5
                shp = torch.randn(N, device=dev, generator=g) * 0.3
                keep += [yprev, scp, shp]
                epi = L.Epilogue(out=out.data_ptr(), ldo=N, stats=stats.data_ptr(), yprev=yprev.data_ptr(), ldyp=N,
                                 scp=scp.data_ptr(), shp=shp.data_ptr(), dscale=1.0)
                fn = lambda A=A, Wl=Wl, N=N, K=K, epi=epi: rt.rows_gemm(A, Wl, K, 1, M, N, K, epi, L.EPI_RELUMASK_STATS)  # noqa: E731
                nbytes = 4 * M * (2 * K + 2 * N)
            else:
                epi = L.Epilogue(out=out.data_ptr(), ldo=N)
                fn = lambda A=A, Wl=Wl, N=N, K=K, epi=epi: rt.rows_gemm(A, Wl, K, 1, M, N, K, epi, L.EPI_STORE)  # noqa: E731
                nbytes = 4 * M * (2 * K + N)
        ms = timed(fn)
        tf = 3 * 2.0 * M * K * N / ms / 1e9
        rows.append((name, K, N, ms, nbytes / ms / 1e6, tf))
        del keep
        torch.cuda.empty_cache()
    # adj_conv7 + max over points in one pass, and the image-operand variants (warp-specialised kernel only)
    if L.lib().wspc_conv1x1_pool_supported(M, 1024, 192, Np):
        import ctypes
        a = torch.randn((M, 192), device=dev, generator=g)
        img = rt.RowImage(M, 192, dev)
        rows.append(("rows_image 192", 192, 0, timed(lambda: img.build(a, 192)), 0.0, 0.0))
        W = torch.randn((192, 1024), device=dev, generator=g) * 0.1
        b = torch.randn(1024, device=dev, generator=g)
        gamma = torch.randn(1024, device=dev, generator=g)
        stats = torch.zeros((2, 1024), dtype=torch.float64, device=dev)
        keys = torch.empty((clouds, 1024), dtype=torch.int64, device=dev)
        ws = torch.empty(L.lib().wspc_conv1x1_rows_workspace_bytes(1024, 192), dtype=torch.uint8, device=dev)
        for tag, (A, mode) in (("conv7+pool plain", (L.Operand(p=a.data_ptr(), ld=192, C=192), L.OP_PLAIN)),
                               ("conv7+pool image", img.operand())):
            fn = lambda A=A, mode=mode: L.check(L.lib().wspc_conv1x1_pool_fwd(  # noqa: E731
                ctypes.byref(A), mode, L.ptr(W), 1024, M, 1024, 192, Np, L.ptr(b), L.ptr(gamma), L.ptr(stats), L.ptr(keys),
                L.ptr(ws), ws.numel(), L.stream()))
            ms = timed(fn)
            rows.append((tag, 192, 1024, ms, 4 * M * 192 / ms / 1e6, 3 * 2.0 * M * 192 * 1024 / ms / 1e9))
        W5 = torch.randn((192, 512), device=dev, generator=g) * 0.1
        out = torch.empty((M, 512), device=dev)
        st5 = torch.zeros((2, 512), dtype=torch.float64, device=dev)
        epi = L.Epilogue(out=out.data_ptr(), ldo=512, bias=b.data_ptr(), stats=st5.data_ptr())
        ms = timed(lambda: rt.rows_gemm(img.operand(), W5, 512, 0, M, 512, 192, epi, L.EPI_STORE_STATS))
        rows.append(("seg/conv1 fwd image", 192, 512, ms, 4 * M * (192 + 512) / ms / 1e6, 3 * 2.0 * M * 192 * 512 / ms / 1e9))
    print(f"kernel = {os.environ.get('WSPC_ROWGEMM_KERNEL', 'ws')}, rows = {M}")
    for name, K, N, ms, gbs, tf in rows:
        print(f"{name:20s} K={K:4d} N={N:4d}  {ms:7.3f} ms  {gbs:7.0f} GB/s (compulsory)  {tf:6.0f} TFLOP/s executed (3 bf16 passes)")


if __name__ == "__main__":
    main()
