"""Backward of conv2d 1x1 -> BN -> ReLU -> max over N through the Gram identity (csrc/poolconv.cu, runtime.PoolConv)
against the direct formulation dy = c1*G + c2 + c3*y; dA = dy W^T, dW = A^T dy, db = 1^T dy in fp64
(adj_conv7 + maxpool, DGCNN_S3DIS.py:80-85; tconv3 + tmaxpool, transform_nets.py:29-34)."""
from collections import OrderedDict

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,N,cin,cout", [(4, 96, 192, 1024), (3, 70, 128, 1024), (2, 40, 64, 256)])
def test_poolconv_gram_identity(cuda, B, N, cin, cout):
    from weaksuppointcloudseg_b200 import runtime as rt
    g = torch.Generator().manual_seed(B * 1000 + cin)
    P = B * N
    A = torch.relu(torch.randn((P, cin), generator=g)) + 0.1
    params = OrderedDict()
    params["l/weights"] = (torch.randn((cin, cout), generator=g) * 0.1).numpy()
    params["l/biases"] = (torch.randn((cout,), generator=g) * 0.1).numpy()
    for n, v in (("beta", 0.0), ("gamma", 1.0), ("pop_mean", 0.0), ("pop_var", 1.0)):
        params["l/bn/" + n] = np.full((cout,), v, np.float32)
    vs = rt.VariableStore(params, cuda)
    layer = rt.Layer(vs, "l", cin, cout, True)
    c1, c2, c3 = (torch.randn(cout, generator=g) * s for s in (1.0, 0.01, 0.01))
    layer.c1.copy_(c1), layer.c2.copy_(c2), layer.c3.copy_(c3)
    dg = torch.randn((B, cout), generator=g)
    dg[torch.rand((B, cout), generator=g) < 0.3] = 0.0          # ReLU-gated entries
    amax = torch.randint(0, N, (B, cout), generator=g, dtype=torch.int32)
    dx0 = torch.randn((P, cin), generator=g)                    # dA already holds other contributions
    # ---- direct formulation, fp64
    Ad, Wd, bd = A.double(), torch.from_numpy(params["l/weights"]).double(), torch.from_numpy(params["l/biases"]).double()
    y = Ad @ Wd + bd
    G = torch.zeros((P, cout), dtype=torch.float64)
    rows = (torch.arange(B).view(B, 1) * N + amax.long())
    G[rows, torch.arange(cout).expand(B, cout)] = dg.double()
    dy = c1.double() * G + c2.double() + c3.double() * y
    dA_ref, dW_ref, db_ref = dx0.double() + dy @ Wd.t(), Ad.t() @ dy, dy.sum(0)
    # ---- device path
    Ag, dgg, amg, dx = A.to(cuda), dg.to(cuda), amax.to(cuda), dx0.to(cuda).clone()
    pc = rt.PoolConv(layer, cuda)
    r0 = pc.prepare()
    dx += r0                                                     # the engines fold r0 into the GEMM that first writes dA
    pc.backward(Ag, cin, P, B, N, dgg, amg, dx.data_ptr(), cin)
    torch.cuda.synchronize()
    rel = lambda a, b: float((a.cpu().double() - b).abs().max() / b.abs().max())   # noqa: E731
    # 2e-3: the dense terms (A^T A) W diag(c3) and (A^T 1) t^T are large and partly cancel (bf16x3 products, 2^-16)
    assert rel(dx, dA_ref) <= 2e-3
    assert rel(layer.dW, dW_ref) <= 2e-3
    assert rel(layer.db, db_ref) <= 2e-3
