"""Kernel-level parity: every MLP / pooling / loss kernel against fp64 torch formulas on identical inputs
(no ReLU/arg-max discontinuity in play, so the bound is tight).  The GEMMs are checked on both device
paths: tcgen05 (auto) and CUDA-core (wspc_set_gemm_path(1)).

Reference semantics: tf_util.conv2d / batch_norm_dist_template / get_edge_feature
(Networks/dgcnn/utils/tf_util.py:115-173,502-535,674-706) and SURVEY App. E for the gradients.
"""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.fixture(params=[0, 1], ids=["tcgen05", "cuda-core"])
def gemm_path(request, cuda):
    from weaksuppointcloudseg_b200 import _lib as L
    old = L.lib().wspc_set_gemm_path(request.param)
    yield request.param
    L.lib().wspc_set_gemm_path(old)


def _tol(path):
    # bf16x3 split on tensor cores carries ~2^-16 relative error per product; fp32 FMA chains ~1e-6
    return 2e-4 if path == 0 else 2e-5


def test_fwd_bnrelu_store_stats(cuda, gemm_path):
    from weaksuppointcloudseg_b200 import _lib as L, runtime as rt
    g = torch.Generator(device="cuda").manual_seed(0)
    M, K, N = 128 * 37 + 53, 64, 64
    y = torch.randn((M, K), device=cuda, generator=g)
    sc = torch.rand(K, device=cuda, generator=g) + 0.5
    sh = torch.randn(K, device=cuda, generator=g) * 0.2
    W = torch.randn((K, N), device=cuda, generator=g) * 0.2
    b = torch.randn(N, device=cuda, generator=g)
    out = torch.empty((M, N), device=cuda)
    stats = torch.zeros((2, N), dtype=torch.float64, device=cuda)
    A = (L.Operand(p=y.data_ptr(), ld=K, C=K, sc=sc.data_ptr(), sh=sh.data_ptr(), dscale=1.0), L.OP_BNRELU)
    epi = L.Epilogue(out=out.data_ptr(), ldo=N, bias=b.data_ptr(), stats=stats.data_ptr())
    rt.rows_gemm(A, W, N, 0, M, N, K, epi, L.EPI_STORE_STATS)
    ref = torch.relu(y.double() * sc.double() + sh.double()) @ W.double() + b.double()
    assert rel(out, ref) <= _tol(gemm_path)
    assert rel(stats[0], out.double().sum(0)) <= 1e-6
    assert rel(stats[1], (out.double() ** 2).sum(0)) <= 1e-6


@pytest.mark.parametrize("Cx", [64, 16])
def test_fwd_edge(cuda, gemm_path, Cx):
    from weaksuppointcloudseg_b200 import _lib as L, runtime as rt
    g = torch.Generator(device="cuda").manual_seed(1)
    B, Np, k, N = 3, 200, 20, 64
    x = torch.randn((B * Np, 192), device=cuda, generator=g)
    idx = torch.randint(0, Np, (B, Np, k), device=cuda, generator=g, dtype=torch.int32)
    W = torch.randn((2 * Cx, N), device=cuda, generator=g) * 0.1
    b = torch.zeros(N, device=cuda)
    R = B * Np * k
    out = torch.empty((R, N), device=cuda)
    stats = torch.zeros((2, N), dtype=torch.float64, device=cuda)
    A = (L.Operand(p=x.data_ptr() + 4 * 64, ld=192, C=2 * Cx, idx=idx.data_ptr(), k=k, npts=Np), L.OP_EDGE)
    epi = L.Epilogue(out=out.data_ptr(), ldo=N, bias=b.data_ptr(), stats=stats.data_ptr())
    rt.rows_gemm(A, W, N, 0, R, N, 2 * Cx, epi, L.EPI_STORE_STATS)
    xs = x[:, 64:64 + Cx].double().view(B, Np, Cx)
    gidx = (idx.long() + torch.arange(B, device=cuda).view(B, 1, 1) * Np).reshape(-1)
    nb = xs.reshape(B * Np, Cx)[gidx].view(B, Np, k, Cx)
    ctr = xs.unsqueeze(2).expand(B, Np, k, Cx)
    ref = (torch.cat([ctr, nb - ctr], -1) @ W.double()).reshape(R, N)
    assert rel(out, ref) <= _tol(gemm_path)


def test_bwd_dy_relumask_and_scatter(cuda, gemm_path):
    from weaksuppointcloudseg_b200 import _lib as L, runtime as rt
    g = torch.Generator(device="cuda").manual_seed(2)
    B, Np, k = 2, 150, 20
    R, C = B * Np * k, 64
    G = torch.randn((R, C), device=cuda, generator=g)
    y = torch.randn((R, C), device=cuda, generator=g)
    c1, c2, c3 = (torch.randn(C, device=cuda, generator=g) * 0.5 for _ in range(3))
    dy = c1.double() * G.double() + c2.double() + c3.double() * y.double()
    A = (L.Operand(p=G.data_ptr(), ld=C, C=C, y=y.data_ptr(), ldy=C, c1=c1.data_ptr(), c2=c2.data_ptr(),
                   c3=c3.data_ptr()), L.OP_DY)
    # (a) previous layer is conv+BN+ReLU: masked gradient + BN-backward sums
    W = torch.randn((64, C), device=cuda, generator=g) * 0.2          # layer weight (Cin=64, Cout=C)
    yprev = torch.randn((R, 64), device=cuda, generator=g)
    scp = torch.rand(64, device=cuda, generator=g) + 0.5
    shp = torch.randn(64, device=cuda, generator=g) * 0.3
    out = torch.empty((R, 64), device=cuda)
    stats = torch.zeros((2, 64), dtype=torch.float64, device=cuda)
    epi = L.Epilogue(out=out.data_ptr(), ldo=64, stats=stats.data_ptr(), yprev=yprev.data_ptr(), ldyp=64,
                     scp=scp.data_ptr(), shp=shp.data_ptr(), dscale=1.0)
    rt.rows_gemm(A, W, C, 1, R, 64, C, epi, L.EPI_RELUMASK_STATS)
    on = torch.addcmul(shp, yprev, scp) > 0           # same fp32 expression class as the kernel's fmaf
    ref = (dy @ W.double().T) * on
    bad = (out.double() - ref).abs() > _tol(gemm_path) * ref.abs().max()
    # elements whose activation is within rounding of 0 may legitimately flip the mask
    near0 = (yprev.double() * scp.double() + shp.double()).abs() < 1e-6
    assert int((bad & ~near0).sum()) == 0
    assert rel(stats[0], out.double().sum(0)) <= 1e-6
    assert rel(stats[1], (out.double() * yprev.double()).sum(0)) <= 1e-6
    # (b) previous "layer" is the edge-feature gather: scatter-add gradient (tf_util.py:696-705 backward)
    We = torch.randn((128, C), device=cuda, generator=g) * 0.2
    idx = torch.randint(0, Np, (B, Np, k), device=cuda, generator=g, dtype=torch.int32)
    dx = torch.zeros((B * Np, 192), device=cuda)
    epi = L.Epilogue(dx=dx.data_ptr() + 4 * 64, lddx=192, idx=idx.data_ptr(), k=k, npts=Np)
    rt.rows_gemm(A, We, C, 1, R, 128, C, epi, L.EPI_EDGE_SCATTER)
    dE = (dy @ We.double().T).view(B, Np, k, 128)
    ref_dx = torch.zeros((B * Np, 64), dtype=torch.float64, device=cuda)
    ref_dx += (dE[..., :64] - dE[..., 64:]).sum(2).reshape(B * Np, 64)
    gidx = (idx.long() + torch.arange(B, device=cuda).view(B, 1, 1) * Np).reshape(-1)
    ref_dx.index_add_(0, gidx, dE[..., 64:].reshape(-1, 64))
    assert rel(dx[:, 64:128], ref_dx) <= 5 * _tol(gemm_path)
    assert float(dx[:, :64].abs().max()) == 0 and float(dx[:, 128:].abs().max()) == 0


@pytest.mark.parametrize("K1,K2,mode", [(64, 64, "bnrelu"), (128, 64, "edge"), (192, 1024, "plain"), (512, 256, "bnrelu"),
                                        (18, 64, "edge"), (256, 13, "bnrelu")])
def test_wgrad_matches_fp64(cuda, gemm_path, K1, K2, mode):
    from weaksuppointcloudseg_b200 import _lib as L, runtime as rt
    g = torch.Generator(device="cuda").manual_seed(3)
    B, Np, k = 2, 100, 20
    M = B * Np * k + (0 if mode == "edge" else 37)
    G = torch.randn((M, K2), device=cuda, generator=g)
    y = torch.randn((M, K2), device=cuda, generator=g)
    c1, c2, c3 = (torch.randn(K2, device=cuda, generator=g) * 0.5 for _ in range(3))
    dy = c1.double() * G.double() + c2.double() + c3.double() * y.double()
    Gop = (L.Operand(p=G.data_ptr(), ld=K2, C=K2, y=y.data_ptr(), ldy=K2, c1=c1.data_ptr(), c2=c2.data_ptr(),
                     c3=c3.data_ptr()), L.OP_DY)
    if mode == "edge":
        Cx = K1 // 2
        x = torch.randn((B * Np, Cx), device=cuda, generator=g)
        idx = torch.randint(0, Np, (B, Np, k), device=cuda, generator=g, dtype=torch.int32)
        A = (L.Operand(p=x.data_ptr(), ld=Cx, C=K1, idx=idx.data_ptr(), k=k, npts=Np), L.OP_EDGE)
        gidx = (idx.long() + torch.arange(B, device=cuda).view(B, 1, 1) * Np).reshape(-1)
        ctr = x.double().view(B, Np, 1, Cx).expand(B, Np, k, Cx)
        nb = x.double()[gidx].view(B, Np, k, Cx)
        act = torch.cat([ctr, nb - ctr], -1).reshape(M, K1)
    else:
        a = torch.randn((M, K1), device=cuda, generator=g)
        if mode == "bnrelu":
            sc = torch.rand(K1, device=cuda, generator=g) + 0.5
            sh = torch.randn(K1, device=cuda, generator=g) * 0.2
            A = (L.Operand(p=a.data_ptr(), ld=K1, C=K1, sc=sc.data_ptr(), sh=sh.data_ptr(), dscale=1.0), L.OP_BNRELU)
            act = torch.relu(a.double() * sc.double() + sh.double())
        else:
            A = (L.Operand(p=a.data_ptr(), ld=K1, C=K1), L.OP_PLAIN)
            act = a.double()
    dW = torch.empty((K1, K2), device=cuda)
    db = torch.empty(K2, device=cuda)
    rt.wgrad(A, Gop, M, dW, db, cuda)
    assert rel(dW, act.T @ dy) <= _tol(gemm_path)
    assert rel(db, dy.sum(0)) <= 2e-5


@pytest.mark.parametrize("K1,K2,ldx", [(9, 128, 9), (3, 128, 3), (12, 256, 16), (9, 64, 9)])
def test_wgrad_narrow_inputs(cuda, K1, K2, ldx):
    """weight gradient of the factored first EdgeConv layer (X (P, 9 | 3) against the (P, 128) gradient of [u | v]): the
    narrow-input kernel, ragged row count, X read with a row pitch, bias gradient."""
    from weaksuppointcloudseg_b200 import _lib as L, runtime as rt
    g = torch.Generator(device="cuda").manual_seed(31)
    M = 70001
    xw = torch.randn((M, ldx), device=cuda, generator=g)
    G = torch.randn((M, K2), device=cuda, generator=g)
    A = (L.Operand(p=xw.data_ptr(), ld=ldx, C=K1), L.OP_PLAIN)
    Gop = (L.Operand(p=G.data_ptr(), ld=K2, C=K2), L.OP_DY)
    dW = torch.empty((K1, K2), device=cuda)
    db = torch.empty(K2, device=cuda)
    rt.wgrad(A, Gop, M, dW, db, cuda)
    assert rel(dW, xw[:, :K1].double().T @ G.double()) <= 2e-5
    assert rel(db, G.double().sum(0)) <= 2e-5


def test_maxk_fwd_bwd(cuda):
    from weaksuppointcloudseg_b200 import _lib as L
    g = torch.Generator(device="cuda").manual_seed(4)
    P, k, C = 700, 20, 64
    y = torch.randn((P, k, C), device=cuda, generator=g)
    y[:, 3] = y[:, 7]                                   # exact ties -> equal split of the gradient
    sc = torch.rand(C, device=cuda, generator=g) + 0.5
    sh = torch.randn(C, device=cuda, generator=g) * 0.5
    out = torch.zeros((P, 192), device=cuda)
    L.check(L.lib().wspc_maxk_bnrelu_fwd(L.ptr(y), L.ptr(sc), L.ptr(sh), P, k, C, ctypes.c_void_p(out.data_ptr() + 256),
                                         192, L.stream()))
    act = torch.relu(torch.addcmul(sh, y, sc))
    ref = act.amax(1)
    assert torch.equal(out[:, 64:128], ref)
    dout = torch.randn((P, 192), device=cuda, generator=g)
    G = torch.empty((P, k, C), device=cuda)
    stats = torch.zeros((2, C), dtype=torch.float64, device=cuda)
    L.check(L.lib().wspc_maxk_bnrelu_bwd(L.ptr(y), L.ptr(sc), L.ptr(sh), ctypes.c_void_p(out.data_ptr() + 256), 192,
                                         ctypes.c_void_p(dout.data_ptr() + 256), 192, P, k, C, L.ptr(G), L.ptr(stats),
                                         L.stream()))
    hit = (act == ref.unsqueeze(1)) & (ref.unsqueeze(1) > 0)
    cnt = hit.sum(1, keepdim=True).clamp_min(1)
    Gref = hit * dout[:, 64:128].unsqueeze(1) / cnt
    assert rel(G, Gref) <= 1e-6
    assert rel(stats[0], G.double().sum((0, 1))) <= 1e-6
    assert rel(stats[1], (G.double() * y.double()).sum((0, 1))) <= 1e-6


def test_maxn_and_bn_kernels(cuda):
    from weaksuppointcloudseg_b200 import _lib as L
    g = torch.Generator(device="cuda").manual_seed(5)
    B, N, C = 3, 333, 96
    y = torch.randn((B, N, C), device=cuda, generator=g)
    y[:, 100] = y[:, 50]                                # ties -> first index wins (tf max_pool grad)
    stats = torch.stack([y.double().sum((0, 1)), (y.double() ** 2).sum((0, 1))])
    gamma = torch.rand(C, device=cuda, generator=g) + 0.5
    beta = torch.randn(C, device=cuda, generator=g) * 0.1
    pm, pv = torch.zeros(C, device=cuda), torch.ones(C, device=cuda)
    sc, sh, mean, inv = (torch.empty(C, device=cuda) for _ in range(4))
    L.check(L.lib().wspc_bn_finalize(L.ptr(stats), C, float(B * N), L.ptr(gamma), L.ptr(beta), 1e-3, 0.5, 1, L.ptr(pm),
                                     L.ptr(pv), L.ptr(sc), L.ptr(sh), L.ptr(mean), L.ptr(inv), L.stream()))
    m = y.double().mean((0, 1))
    v = y.double().var((0, 1), unbiased=False)
    assert rel(mean, m) <= 1e-6 and rel(sc, gamma.double() * torch.rsqrt(v + 1e-3)) <= 1e-6
    assert rel(pm, 0.5 * m) <= 1e-6 and rel(pv, 0.5 + 0.5 * v) <= 1e-6
    gmax = torch.empty((B, C), device=cuda)
    amax = torch.empty((B, C), dtype=torch.int32, device=cuda)
    L.check(L.lib().wspc_maxn_bnrelu_fwd(L.ptr(y), L.ptr(sc), L.ptr(sh), B, N, C, L.ptr(gmax), L.ptr(amax), L.stream()))
    act = torch.relu(torch.addcmul(sh, y, sc))
    assert torch.equal(gmax, act.amax(1))
    first = (act == act.amax(1, keepdim=True)).to(torch.uint8).argmax(1)
    assert torch.equal(amax.long(), first)


def test_adam_tf_and_dropout_mask(cuda):
    from weaksuppointcloudseg_b200 import _lib as L
    g = torch.Generator(device="cuda").manual_seed(6)
    n = 100003
    p = torch.randn(n + 1, device=cuda, generator=g)[:n]
    p = p.clone()
    gr = torch.randn(n, device=cuda, generator=g)
    m, v = torch.zeros(n, device=cuda), torch.zeros(n, device=cuda)
    p0 = p.double().clone()
    lr_t = 1e-3 * (1 - 0.999) ** 0.5 / (1 - 0.9)
    L.check(L.lib().wspc_adam_tf(L.ptr(p), L.ptr(gr), L.ptr(m), L.ptr(v), n, lr_t, 0.9, 0.999, 1e-8, 0.5, L.stream()))
    gg = gr.double() * 0.5
    mr, vr = 0.1 * gg, 0.001 * gg * gg
    assert rel(p, p0 - lr_t * mr / (vr.sqrt() + 1e-8)) <= 1e-6
    mask = torch.empty(1 << 20, device=cuda)
    L.check(L.lib().wspc_dropout_mask(L.ptr(mask), mask.numel(), 0.7, 1234, 0, L.stream()))
    assert set(mask.unique().tolist()) <= {0.0, 1.0}
    assert abs(float(mask.mean()) - 0.7) < 5e-3


@pytest.mark.parametrize("K,N,kind", [(192, 1024, "plain_rowbias"), (512, 256, "bnrelu_drop"), (512, 192, "dy_store"),
                                      (1024, 192, "sparse_accum"), (256, 16, "bnrelu_drop"), (1024, 512, "plain_small")])
def test_rows_gemm_chunked_shapes(cuda, gemm_path, K, N, kind):
    """per-point layer shapes: K > 128 is accumulated in TMEM over 64-channel chunks, N > 128 is tiled"""
    from weaksuppointcloudseg_b200 import _lib as L, runtime as rt
    g = torch.Generator(device="cuda").manual_seed(7)
    Bc, Np = 3, 173
    M = Bc * Np if kind != "plain_small" else 5
    W = torch.randn((K, N), device=cuda, generator=g) * 0.1
    out = torch.randn((M, N), device=cuda, generator=g)
    out0 = out.double().clone()
    if kind in ("plain_rowbias", "plain_small"):
        a = torch.randn((M, K), device=cuda, generator=g)
        b = torch.randn(N, device=cuda, generator=g)
        rb = torch.randn((Bc, N), device=cuda, generator=g)
        stats = torch.zeros((2, N), dtype=torch.float64, device=cuda)
        A = (L.Operand(p=a.data_ptr(), ld=K, C=K), L.OP_PLAIN)
        use_rb = kind == "plain_rowbias"
        epi = L.Epilogue(out=out.data_ptr(), ldo=N, bias=b.data_ptr(), rowbias=rb.data_ptr() if use_rb else 0, rb_rows=Np,
                         ldrb=N, stats=stats.data_ptr())
        rt.rows_gemm(A, W, N, 0, M, N, K, epi, L.EPI_STORE_STATS)
        ref = a.double() @ W.double() + b.double()
        if use_rb:
            ref = ref + rb.double().repeat_interleave(Np, 0)
        assert rel(out, ref) <= _tol(gemm_path)
        assert rel(stats[0], out.double().sum(0)) <= 1e-6 and rel(stats[1], (out.double() ** 2).sum(0)) <= 1e-6
    elif kind == "bnrelu_drop":
        y = torch.randn((M, K), device=cuda, generator=g)
        sc = torch.rand(K, device=cuda, generator=g) + 0.5
        sh = torch.randn(K, device=cuda, generator=g) * 0.2
        dm = torch.floor(0.7 + torch.rand((M, K), device=cuda, generator=g))
        b = torch.randn(N, device=cuda, generator=g)
        A = (L.Operand(p=y.data_ptr(), ld=K, C=K, sc=sc.data_ptr(), sh=sh.data_ptr(), dmask=dm.data_ptr(), dscale=1 / 0.7),
             L.OP_BNRELU)
        rt.rows_gemm(A, W, N, 0, M, N, K, L.Epilogue(out=out.data_ptr(), ldo=N, bias=b.data_ptr()), L.EPI_STORE)
        ref = (torch.relu(y.double() * sc.double() + sh.double()) * dm.double() / 0.7) @ W.double() + b.double()
        assert rel(out, ref) <= _tol(gemm_path)
    elif kind == "dy_store":
        G = torch.randn((M, K), device=cuda, generator=g)
        y = torch.randn((M, K), device=cuda, generator=g)
        c1, c2, c3 = (torch.randn(K, device=cuda, generator=g) * 0.5 for _ in range(3))
        Wl = torch.randn((N, K), device=cuda, generator=g) * 0.1          # layer weight (Cin=N, Cout=K)
        A = (L.Operand(p=G.data_ptr(), ld=K, C=K, y=y.data_ptr(), ldy=K, c1=c1.data_ptr(), c2=c2.data_ptr(),
                       c3=c3.data_ptr()), L.OP_DY)
        rt.rows_gemm(A, Wl, K, 1, M, N, K, L.Epilogue(out=out.data_ptr(), ldo=N), L.EPI_STORE)
        dy = c1.double() * G.double() + c2.double() + c3.double() * y.double()
        assert rel(out, dy @ Wl.double().T) <= _tol(gemm_path)
    else:  # sparse_accum: gradient of max_pool2d + BN folded, accumulated onto dcat
        y = torch.randn((M, K), device=cuda, generator=g)
        c1, c2, c3 = (torch.randn(K, device=cuda, generator=g) * 0.5 for _ in range(3))
        dg = torch.randn((Bc, K), device=cuda, generator=g)
        amax = torch.randint(0, Np, (Bc, K), device=cuda, generator=g, dtype=torch.int32)
        Wl = torch.randn((N, K), device=cuda, generator=g) * 0.1
        A = (L.Operand(p=0, ld=0, C=K, y=y.data_ptr(), ldy=K, c1=c1.data_ptr(), c2=c2.data_ptr(), c3=c3.data_ptr(),
                       dg=dg.data_ptr(), amax=amax.data_ptr(), npts=Np), L.OP_DY_SPARSE)
        rt.rows_gemm(A, Wl, K, 1, M, N, K, L.Epilogue(out=out.data_ptr(), ldo=N), L.EPI_ACCUM)
        Gd = torch.zeros((Bc, Np, K), dtype=torch.float64, device=cuda)
        Gd.scatter_(1, amax.long().unsqueeze(1), dg.double().unsqueeze(1))
        dy = c1.double() * Gd.reshape(M, K) + c2.double() + c3.double() * y.double()
        assert rel(out, out0 + dy @ Wl.double().T) <= _tol(gemm_path)


@pytest.mark.parametrize("K,N,kind", [(192, 1024, "plain_stats_rowbias"), (192, 512, "plain_stats_rowbias"),
                                      (512, 256, "bnrelu_stats"), (256, 512, "dy_relumask_drop"), (512, 192, "dy_store"),
                                      (160, 320, "bnrelu_stats"), (256, 256, "dy_plain_store")])
def test_rows_gemm_warp_specialised(cuda, K, N, kind):
    """per-point layers at training size (>= 592 row tiles, K > 128 in 32-channel chunks): the warp-specialised kernel
    (cp.async operand ring, MMA issue off the epilogue warps, two TMEM accumulators).  Ragged last tile, several column
    tiles, every operand map / epilogue it takes over from the serial kernel."""
    from weaksuppointcloudseg_b200 import _lib as L, runtime as rt
    g = torch.Generator(device="cuda").manual_seed(11)
    Bc, Np = 19, 4001
    M = Bc * Np                                                      # 76019 rows: 593.9 tiles
    W = torch.randn((K, N), device=cuda, generator=g) * 0.1
    out = torch.full((M, N), 3.0, device=cuda)
    tol = 2e-4                                                       # bf16 x 3 split, as _tol(0)
    if kind == "plain_stats_rowbias":
        a = torch.randn((M, K), device=cuda, generator=g)
        b = torch.randn(N, device=cuda, generator=g)
        rb = torch.randn((Bc, N), device=cuda, generator=g)
        stats = torch.zeros((2, N), dtype=torch.float64, device=cuda)
        A = (L.Operand(p=a.data_ptr(), ld=K, C=K), L.OP_PLAIN)
        epi = L.Epilogue(out=out.data_ptr(), ldo=N, bias=b.data_ptr(), rowbias=rb.data_ptr(), rb_rows=Np, ldrb=N,
                         stats=stats.data_ptr())
        rt.rows_gemm(A, W, N, 0, M, N, K, epi, L.EPI_STORE_STATS)
        ref = a.double() @ W.double() + b.double() + rb.double().repeat_interleave(Np, 0)
        assert rel(out, ref) <= tol
        assert rel(stats[0], out.double().sum(0)) <= 1e-6 and rel(stats[1], (out.double() ** 2).sum(0)) <= 1e-6
    elif kind == "bnrelu_stats":
        y = torch.randn((M, K), device=cuda, generator=g)
        sc = torch.rand(K, device=cuda, generator=g) + 0.5
        sh = torch.randn(K, device=cuda, generator=g) * 0.2
        b = torch.randn(N, device=cuda, generator=g)
        stats = torch.zeros((2, N), dtype=torch.float64, device=cuda)
        A = (L.Operand(p=y.data_ptr(), ld=K, C=K, sc=sc.data_ptr(), sh=sh.data_ptr()), L.OP_BNRELU)
        rt.rows_gemm(A, W, N, 0, M, N, K, L.Epilogue(out=out.data_ptr(), ldo=N, bias=b.data_ptr(), stats=stats.data_ptr()),
                     L.EPI_STORE_STATS)
        ref = torch.relu(y.double() * sc.double() + sh.double()) @ W.double() + b.double()
        assert rel(out, ref) <= tol
        assert rel(stats[0], out.double().sum(0)) <= 1e-6 and rel(stats[1], (out.double() ** 2).sum(0)) <= 1e-6
    else:
        G = torch.randn((M, K), device=cuda, generator=g)
        y = torch.randn((M, K), device=cuda, generator=g)
        Wl = torch.randn((N, K), device=cuda, generator=g) * 0.1          # layer weight (Cin=N, Cout=K)
        if kind == "dy_plain_store":
            A = (L.Operand(p=G.data_ptr(), ld=K, C=K), L.OP_DY)
            dy = G.double()
        else:
            c1, c2, c3 = (torch.randn(K, device=cuda, generator=g) * 0.5 for _ in range(3))
            A = (L.Operand(p=G.data_ptr(), ld=K, C=K, y=y.data_ptr(), ldy=K, c1=c1.data_ptr(), c2=c2.data_ptr(),
                           c3=c3.data_ptr()), L.OP_DY)
            dy = c1.double() * G.double() + c2.double() + c3.double() * y.double()
        if kind == "dy_relumask_drop":
            yprev = torch.randn((M, N), device=cuda, generator=g)
            scp = torch.rand(N, device=cuda, generator=g) + 0.5
            shp = torch.randn(N, device=cuda, generator=g) * 0.3
            dm = torch.floor(0.7 + torch.rand((M, N), device=cuda, generator=g))
            stats = torch.zeros((2, N), dtype=torch.float64, device=cuda)
            epi = L.Epilogue(out=out.data_ptr(), ldo=N, stats=stats.data_ptr(), yprev=yprev.data_ptr(), ldyp=N,
                             scp=scp.data_ptr(), shp=shp.data_ptr(), dmask=dm.data_ptr(), dscale=1 / 0.7)
            rt.rows_gemm(A, Wl, K, 1, M, N, K, epi, L.EPI_RELUMASK_STATS)
            on = torch.addcmul(shp, yprev, scp) > 0
            ref = (dy @ Wl.double().T) * on * dm.double() / 0.7
            bad = (out.double() - ref).abs() > tol * ref.abs().max()
            near0 = (yprev.double() * scp.double() + shp.double()).abs() < 1e-6
            assert int((bad & ~near0).sum()) == 0
            assert rel(stats[0], out.double().sum(0)) <= 1e-6
            assert rel(stats[1], (out.double() * yprev.double()).sum(0)) <= 1e-6
        else:
            rt.rows_gemm(A, Wl, K, 1, M, N, K, L.Epilogue(out=out.data_ptr(), ldo=N), L.EPI_STORE)
            assert rel(out, dy @ Wl.double().T) <= tol


@pytest.mark.parametrize("Bc,Np,K,N", [(20, 3840, 192, 1024), (38, 2048, 256, 320)])
def test_conv_pool_forward_fused(cuda, Bc, Np, K, N):
    """adj_conv7 + max_pool2d([N,1]) in one pass (wspc_conv1x1_pool_fwd + wspc_maxn_from_keys): BN sums, pooled activation,
    arg-max rows and the pre-BN value at the arg-max against the unfused definition, with negative BN scales (the pooled row is
    then the column MINIMUM) and a tie (duplicated points: the first row wins, as max_pool2d's gradient does)."""
    import ctypes
    from weaksuppointcloudseg_b200 import _lib as L
    g = torch.Generator(device="cuda").manual_seed(5)
    M = Bc * Np
    assert L.lib().wspc_conv1x1_pool_supported(M, N, K, Np) == 1
    assert L.lib().wspc_conv1x1_pool_supported(M, N, K, Np - 64) == 0          # tiles would straddle clouds
    a = torch.randn((M, K), device=cuda, generator=g)
    a[5 * Np + 77] = a[5 * Np + 3]                                            # duplicated point inside cloud 5
    W = torch.randn((K, N), device=cuda, generator=g) * 0.1
    b = torch.randn(N, device=cuda, generator=g)
    gamma = torch.randn(N, device=cuda, generator=g)
    gamma[::7] = -gamma[::7].abs()
    stats = torch.zeros((2, N), dtype=torch.float64, device=cuda)
    keys = torch.empty((Bc, N), dtype=torch.int64, device=cuda)
    nbytes = L.lib().wspc_conv1x1_rows_workspace_bytes(N, K)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=cuda)
    A = L.Operand(p=a.data_ptr(), ld=K, C=K)
    L.check(L.lib().wspc_conv1x1_pool_fwd(ctypes.byref(A), L.OP_PLAIN, L.ptr(W), N, M, N, K, Np, L.ptr(b), L.ptr(gamma),
                                          L.ptr(stats), L.ptr(keys), L.ptr(ws), ws.numel(), L.stream()))
    y = (a.double() @ W.double() + b.double())
    assert rel(stats[0], y.sum(0)) <= 2e-5 and rel(stats[1], (y ** 2).sum(0)) <= 2e-5
    mean, var = y.mean(0), y.var(0, unbiased=False)
    sc = (gamma.double() / torch.sqrt(var + 1e-3)).float()
    sh = (-mean * sc.double()).float()
    gout = torch.empty((Bc, N), device=cuda)
    amax = torch.empty((Bc, N), dtype=torch.int32, device=cuda)
    ymax = torch.empty((Bc, N), device=cuda)
    L.check(L.lib().wspc_maxn_from_keys(L.ptr(keys), L.ptr(gamma), L.ptr(sc), L.ptr(sh), Bc, N, L.ptr(gout), L.ptr(amax),
                                        L.ptr(ymax), L.stream()))
    y3 = y.view(Bc, Np, N)
    act = torch.relu(y3 * sc.double() + sh.double())
    assert rel(gout, act.max(1).values) <= 2e-4
    assert int(amax.min()) >= 0 and int(amax.max()) < Np
    picked = torch.gather(y3, 1, amax.long().unsqueeze(1)).squeeze(1)
    assert rel(ymax, picked) <= 2e-4                                          # the reported value is the one at the reported row
    ext = torch.where(gamma >= 0, y3.max(1).values, y3.min(1).values)
    assert float(((picked - ext).abs() / y.abs().max()).max()) <= 2e-4        # ... and it is the column extreme
    # tie: wherever the duplicated pair holds the extreme, the earlier row is reported
    dup = (amax[5] == 77)
    assert int(dup.sum()) == 0
    with pytest.raises(Exception):
        L.check(L.lib().wspc_conv1x1_pool_fwd(ctypes.byref(A), L.OP_PLAIN, L.ptr(W), N, M, N, K, Np - 64, L.ptr(b), L.ptr(gamma),
                                              L.ptr(stats), L.ptr(keys), L.ptr(ws), ws.numel(), L.stream()))


def test_rows_image_operand(cuda):
    """WSPC_OP_IMG: the pre-split image of a matrix (wspc_rows_image) gives the same products as the fp32 operand -- store +
    statistics + per-cloud row bias (seg/conv1) and the pooled epilogue (adj_conv7) -- with a ragged last tile; ineligible
    shapes and the CUDA-core path refuse it loudly."""
    import ctypes
    from weaksuppointcloudseg_b200 import _lib as L, runtime as rt
    g = torch.Generator(device="cuda").manual_seed(21)
    Bc, Np, K, N = 19, 4001, 192, 512
    M = Bc * Np
    assert L.lib().wspc_rows_image_supported(M, N, K) == 1 and L.lib().wspc_rows_image_supported(1000, N, K) == 0
    wide = torch.randn((M, 256), device=cuda, generator=g)
    a = wide[:, 32:32 + K]                                                   # a channel window: ld != K
    img = rt.RowImage(M, K, cuda)
    img.build(a, 256)
    W = torch.randn((K, N), device=cuda, generator=g) * 0.1
    b = torch.randn(N, device=cuda, generator=g)
    rb = torch.randn((Bc, N), device=cuda, generator=g)
    outs, stats = [], []
    for A in (img.operand(), (L.Operand(p=a.data_ptr(), ld=256, C=K), L.OP_PLAIN)):
        out = torch.empty((M, N), device=cuda)
        st = torch.zeros((2, N), dtype=torch.float64, device=cuda)
        epi = L.Epilogue(out=out.data_ptr(), ldo=N, bias=b.data_ptr(), rowbias=rb.data_ptr(), rb_rows=Np, ldrb=N, stats=st.data_ptr())
        rt.rows_gemm(A, W, N, 0, M, N, K, epi, L.EPI_STORE_STATS)
        outs.append(out)
        stats.append(st)
    assert torch.equal(outs[0], outs[1])                                     # same split, same MMAs, same order
    ref = a.double() @ W.double() + b.double() + rb.double().repeat_interleave(Np, 0)
    assert rel(outs[0], ref) <= 2e-4 and rel(stats[0], stats[1]) <= 1e-9
    # pooled epilogue on the image (whole clouds per tile set: 20 x 3840)
    Bc2, Np2 = 20, 3840
    M2 = Bc2 * Np2
    a2 = torch.randn((M2, K), device=cuda, generator=g)
    img2 = rt.RowImage(M2, K, cuda)
    img2.build(a2, K)
    W2 = torch.randn((K, 1024), device=cuda, generator=g) * 0.1
    b2 = torch.randn(1024, device=cuda, generator=g)
    gamma = torch.randn(1024, device=cuda, generator=g)
    ws = torch.empty(L.lib().wspc_conv1x1_rows_workspace_bytes(1024, K), dtype=torch.uint8, device=cuda)
    keys = []
    for (A, mode) in (img2.operand(), (L.Operand(p=a2.data_ptr(), ld=K, C=K), L.OP_PLAIN)):
        kk = torch.empty((Bc2, 1024), dtype=torch.int64, device=cuda)
        st = torch.zeros((2, 1024), dtype=torch.float64, device=cuda)
        L.check(L.lib().wspc_conv1x1_pool_fwd(ctypes.byref(A), mode, L.ptr(W2), 1024, M2, 1024, K, Np2, L.ptr(b2), L.ptr(gamma),
                                              L.ptr(st), L.ptr(kk), L.ptr(ws), ws.numel(), L.stream()))
        keys.append(kk)
    assert torch.equal(keys[0], keys[1])
    # refusals
    small = torch.randn((1000, K), device=cuda, generator=g)
    simg = rt.RowImage(1000, K, cuda)
    simg.build(small, K)
    out = torch.empty((1000, N), device=cuda)
    with pytest.raises(Exception):
        rt.rows_gemm(simg.operand(), W, N, 0, 1000, N, K, L.Epilogue(out=out.data_ptr(), ldo=N), L.EPI_STORE)
    prev = L.lib().wspc_set_gemm_path(1)
    try:
        with pytest.raises(Exception):
            rt.rows_gemm(img.operand(), W, N, 0, M, N, K, L.Epilogue(out=outs[0].data_ptr(), ldo=N), L.EPI_STORE)
    finally:
        L.lib().wspc_set_gemm_path(prev)


# ---- narrow heads (seg/conv3 of the S3DIS net: 256 -> 13, DGCNN_S3DIS.py:100-101) ------------------------------------
@pytest.mark.parametrize("K,N,dropout", [(256, 13, True), (128, 16, False), (64, 9, False)])
def test_narrow_head_forward(cuda, gemm_path, K, N, dropout):
    """warp-per-row kernel (auto path) / generic CUDA-core kernel: relu(bn(y)) (* dropout) @ W + b for N <= 16."""
    from weaksuppointcloudseg_b200 import _lib as L, runtime as rt
    g = torch.Generator(device="cuda").manual_seed(7)
    M = 1000 + 37
    y = torch.randn((M, K), device=cuda, generator=g)
    sc = torch.rand(K, device=cuda, generator=g) + 0.5
    sh = torch.randn(K, device=cuda, generator=g) * 0.2
    W = torch.randn((K, N), device=cuda, generator=g) * 0.2
    b = torch.randn(N, device=cuda, generator=g)
    mask = (torch.rand((M, K), device=cuda, generator=g) < 0.7).float() if dropout else None
    out = torch.empty((M, N), device=cuda)
    A = (L.Operand(p=y.data_ptr(), ld=K, C=K, sc=sc.data_ptr(), sh=sh.data_ptr(), dmask=L.dptr(mask), dscale=1.0 / 0.7), L.OP_BNRELU)
    rt.rows_gemm(A, W, N, 0, M, N, K, L.Epilogue(out=out.data_ptr(), ldo=N, bias=b.data_ptr()), L.EPI_STORE)
    a = torch.relu(y.double() * sc.double() + sh.double())
    if dropout:
        a = a * mask.double() / 0.7
    assert rel(out, a @ W.double() + b.double()) <= 2e-5


def test_narrow_head_data_gradient(cuda, gemm_path):
    """dA (M,256) = dZ (M,13) W^T with the ReLU mask / dropout of the producing layer and its BN-backward sums."""
    from weaksuppointcloudseg_b200 import _lib as L, runtime as rt
    g = torch.Generator(device="cuda").manual_seed(8)
    M, K, N = 777, 13, 256
    dZ = torch.randn((M, K), device=cuda, generator=g)
    W = torch.randn((N, K), device=cuda, generator=g) * 0.2           # the layer's own (256, 13) weight, used transposed
    yprev = torch.randn((M, N), device=cuda, generator=g)
    sc = torch.rand(N, device=cuda, generator=g) + 0.5
    sh = torch.randn(N, device=cuda, generator=g) * 0.2
    mask = (torch.rand((M, N), device=cuda, generator=g) < 0.7).float()
    out = torch.empty((M, N), device=cuda)
    stats = torch.zeros((2, N), dtype=torch.float64, device=cuda)
    A = (L.Operand(p=dZ.data_ptr(), ld=K, C=K), L.OP_DY)
    epi = L.Epilogue(out=out.data_ptr(), ldo=N, stats=stats.data_ptr(), yprev=yprev.data_ptr(), ldyp=N, scp=sc.data_ptr(),
                     shp=sh.data_ptr(), dmask=mask.data_ptr(), dscale=1.0 / 0.7)
    rt.rows_gemm(A, W, K, 1, M, N, K, epi, L.EPI_RELUMASK_STATS)
    on = (yprev.double() * sc.double() + sh.double()) > 0
    ref = (dZ.double() @ W.double().t()) * on * mask.double() / 0.7
    assert rel(out, ref) <= 2e-5
    assert rel(stats[0], ref.sum(0)) <= 1e-5
    assert rel(stats[1], (ref * yprev.double()).sum(0)) <= 1e-5


@pytest.mark.parametrize("K1,K2", [(64, 64), (64, 128)])
def test_fused_weight_and_data_gradient(cuda, K1, K2):
    """wspc_conv1x1_bwd_fused == wspc_conv1x1_wgrad + wspc_conv1x1_rows(RELUMASK_STATS) on the same operands."""
    from weaksuppointcloudseg_b200 import _lib as L, runtime as rt
    g = torch.Generator(device="cuda").manual_seed(21 + K2)
    M = 128 * 11 + 77
    ya = torch.randn((M, K1), device=cuda, generator=g)          # pre-BN output of the previous layer
    sca = torch.rand(K1, device=cuda, generator=g) + 0.5
    sha = torch.randn(K1, device=cuda, generator=g) * 0.3
    G = torch.randn((M, K2), device=cuda, generator=g)
    yb = torch.randn((M, K2), device=cuda, generator=g)
    c1, c2, c3 = (torch.randn(K2, device=cuda, generator=g) * 0.5 for _ in range(3))
    W = torch.randn((K1, K2), device=cuda, generator=g) * 0.2
    A = (L.Operand(p=ya.data_ptr(), ld=K1, C=K1, sc=sca.data_ptr(), sh=sha.data_ptr(), dscale=1.0), L.OP_BNRELU)
    Gop = (L.Operand(p=G.data_ptr(), ld=K2, C=K2, y=yb.data_ptr(), ldy=K2, c1=c1.data_ptr(), c2=c2.data_ptr(), c3=c3.data_ptr()),
           L.OP_DY)
    outs = []
    for fused in (False, True):
        dW, db = torch.empty((K1, K2), device=cuda), torch.empty(K2, device=cuda)
        dA = torch.empty((M, K1), device=cuda)
        st = torch.zeros((2, K1), dtype=torch.float64, device=cuda)
        epi = L.Epilogue(out=dA.data_ptr(), ldo=K1, stats=st.data_ptr(), yprev=ya.data_ptr(), ldyp=K1, scp=sca.data_ptr(),
                         shp=sha.data_ptr(), dscale=1.0)
        if fused:
            ws = L.workspace(L.lib().wspc_conv1x1_wgrad_workspace_bytes(K1, K2), cuda, "wgrad")
            L.check(L.lib().wspc_conv1x1_bwd_fused(ctypes.byref(A[0]), A[1], ctypes.byref(Gop[0]), Gop[1], M, L.ptr(W), K2,
                                                   ctypes.byref(epi), L.ptr(dW), L.ptr(db), L.ptr(ws), ws.numel(), L.stream()))
        else:
            rt.wgrad(A, Gop, M, dW, db, cuda)
            rt.rows_gemm(Gop, W, K2, 1, M, K1, K2, epi, L.EPI_RELUMASK_STATS)
        torch.cuda.synchronize()
        outs.append((dW, db, dA, st))
    # fp64 reference
    a = torch.relu(ya.double() * sca.double() + sha.double())
    dy = c1.double() * G.double() + c2.double() + c3.double() * yb.double()
    dA_ref = (dy @ W.double().t()) * (a > 0)
    for dW, db, dA, st in outs:
        assert rel(dW, a.t() @ dy) <= 2e-4
        assert rel(db, dy.sum(0)) <= 2e-4
        assert rel(dA, dA_ref) <= 2e-4
        assert rel(st[0], dA_ref.sum(0)) <= 2e-4
        assert rel(st[1], (dA_ref * ya.double()).sum(0)) <= 2e-4
    assert rel(outs[1][2], outs[0][2].double()) <= 1e-5      # same tensor-core arithmetic in both formulations
