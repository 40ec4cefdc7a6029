"""tf.reduce_max(axis=-2) backward without materialising the (P*k, C) gradient: the statistics-only kernel +
WSPC_OP_DY_MAXK operand (G synthesised on load) must reproduce the materialised path (wspc_maxk_bnrelu_bwd + WSPC_OP_DY)
in the weight gradient, the data gradient (+ ReLU-mask epilogue sums) and the factored EdgeConv backward.
Reference semantics: DGCNN_S3DIS.py:46,62,78 (reduce_max over k; gradient split equally among ties [TF])."""
from collections import OrderedDict

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _layer(rt, cuda, name, cin, cout, seed):
    g = torch.Generator().manual_seed(seed)
    params = OrderedDict()
    params[name + "/weights"] = (torch.randn((cin, cout), generator=g) * 0.2).numpy()
    params[name + "/biases"] = np.zeros((cout,), np.float32)
    for n, v in (("beta", 0.0), ("gamma", 1.0), ("pop_mean", 0.0), ("pop_var", 1.0)):
        params[name + "/bn/" + n] = np.full((cout,), v, np.float32)
    vs = rt.VariableStore(params, cuda)
    ly = rt.Layer(vs, name, cin, cout, True)
    ly.sc.copy_(torch.rand(cout, generator=g) + 0.5)
    ly.sh.copy_(torch.randn(cout, generator=g) * 0.3)
    for t in (ly.c1, ly.c2, ly.c3):
        t.copy_(torch.randn(cout, generator=g) * 0.5)
    return ly


@pytest.mark.parametrize("B,N,k,C", [(2, 80, 20, 64), (1, 50, 7, 128)])
def test_synthesised_maxk_gradient_matches_materialised(cuda, B, N, k, C):
    from weaksuppointcloudseg_b200 import _lib as L, runtime as rt
    P, R = B * N, B * N * k
    g = torch.Generator().manual_seed(C + k)
    lb = _layer(rt, cuda, "b", 64, C, 1)          # the layer whose activation is max-pooled
    la = _layer(rt, cuda, "a", 64, 64, 2)         # the layer before it (its pre-BN output is the ReLU mask / operand)
    y = torch.randn((R, C), generator=g)
    y[::3] = y[1::3][: len(y[::3])]               # exact ties inside a point's k rows
    y = y.to(cuda).contiguous()
    ya = torch.randn((R, 64), generator=g).to(cuda)
    out = torch.zeros((P, C), device=cuda)
    dout = torch.randn((P, C), generator=g).to(cuda)
    rt.maxk_fwd(lb, y, P, k, out.data_ptr(), C)
    # materialised
    G = torch.empty((R, C), device=cuda)
    rt.maxk_bwd(lb, y, P, k, out.data_ptr(), C, dout.data_ptr(), C, G)
    st_ref = lb.bstats.clone()
    dW_ref, db_ref = torch.empty((64, C), device=cuda), torch.empty(C, device=cuda)
    rt.wgrad(rt.op_bnrelu(ya, la), rt.op_dy(G, C, y, C, lb, C), R, dW_ref, db_ref, cuda)
    Ga_ref = torch.empty((R, 64), device=cuda)
    e, m = rt.epi_relumask(Ga_ref, la, ya)
    rt.rows_gemm(rt.op_dy(G, C, y, C, lb, C), lb.W, C, 1, R, 64, C, e, m)
    sta_ref = la.bstats.clone()
    # synthesised
    MS = torch.empty((P, 2 * C), device=cuda)
    rt.maxk_bwd_stats(lb, y, P, k, out.data_ptr(), C, dout.data_ptr(), C, MS)
    torch.cuda.synchronize()
    assert torch.allclose(lb.bstats, st_ref, rtol=1e-5, atol=1e-4)
    dW, db = torch.empty((64, C), device=cuda), torch.empty(C, device=cuda)
    Gm = rt.op_dy_maxk(lb, y, MS, k, N)
    rt.wgrad(rt.op_bnrelu(ya, la), Gm, R, dW, db, cuda)
    Ga = torch.empty((R, 64), device=cuda)
    e, m = rt.epi_relumask(Ga, la, ya)
    rt.rows_gemm(Gm, lb.W, C, 1, R, 64, C, e, m)
    torch.cuda.synchronize()
    scale = lambda t: float(t.abs().max())   # noqa: E731
    assert float((dW - dW_ref).abs().max()) <= 1e-5 * scale(dW_ref)
    assert float((db - db_ref).abs().max()) <= 1e-5 * scale(db_ref)
    assert float((Ga - Ga_ref).abs().max()) <= 1e-5 * scale(Ga_ref)
    assert torch.allclose(la.bstats, sta_ref, rtol=1e-5, atol=1e-4)
    if C == 64:   # factored EdgeConv backward with the synthesised gradient
        idx = torch.randint(0, N, (B, N, k), generator=g, dtype=torch.int32).to(cuda)
        D1, D2 = torch.zeros((P, 128), device=cuda), torch.zeros((P, 128), device=cuda)
        L.check(L.lib().wspc_edge_combine_bwd(L.ptr(G), L.ptr(y), L.ptr(lb.c1), L.ptr(lb.c2), L.ptr(lb.c3), L.ptr(idx), P, k, N,
                                              64, L.ptr(D1), 128, L.stream()))
        L.check(L.lib().wspc_edge_combine_bwd_maxk(L.ptr(y), L.ptr(lb.c1), L.ptr(lb.c2), L.ptr(lb.c3), L.ptr(lb.sc),
                                                   L.ptr(lb.sh), L.ptr(MS), L.ptr(idx), P, k, N, 64, L.ptr(D2), 128, L.stream()))
        torch.cuda.synchronize()
        assert float((D1 - D2).abs().max()) <= 1e-5 * scale(D1)
