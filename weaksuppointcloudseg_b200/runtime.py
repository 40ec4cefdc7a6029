"""Host-side runtime shared by the two DGCNN engines: the scope-keyed variable store (TF variable names,
SURVEY §5 checkpoint row), per-layer scratch, and the layer-level launch helpers that translate
"conv2d 1x1 -> BN -> ReLU" (reference tf_util.conv2d, Networks/dgcnn/utils/tf_util.py:115-173) into
calls of the C ABI (include/wspc.h).  No arithmetic happens in Python/PyTorch here: torch only owns
device memory and the stream.
"""
from __future__ import annotations

import ctypes
import os
import math
from collections import OrderedDict

import numpy as np
import torch

from . import _lib as L

BN_EPS = 1e-3  # tf_util.py:527


class VariableStore:
    """All trainable variables live in ONE flat fp32 buffer (theta) with a same-shaped gradient buffer and
    Adam slots, so the optimiser and the data-parallel all-reduce are one call each.  Population BN
    statistics (non-trainable) live in a second flat buffer.  Views are keyed by TF variable names."""

    def __init__(self, params: "OrderedDict[str, np.ndarray]", device):
        self.device = device
        tn = [k for k in params if not (k.endswith("pop_mean") or k.endswith("pop_var"))]
        sn = [k for k in params if k not in tn]
        self.trainable_names, self.state_names = tn, sn

        def pack(names):
            offs, n = {}, 0
            for k in names:
                offs[k] = (n, params[k].shape)
                n += (int(np.prod(params[k].shape)) + 3) // 4 * 4   # keep every view 16-byte aligned
            return offs, n

        self._toffs, nt = pack(tn)
        self._soffs, ns = pack(sn)
        self.theta = torch.zeros(nt, dtype=torch.float32, device=device)
        self.grad = torch.zeros(nt, dtype=torch.float32, device=device)
        self.adam_m = torch.zeros(nt, dtype=torch.float32, device=device)
        self.adam_v = torch.zeros(nt, dtype=torch.float32, device=device)
        self.state = torch.zeros(max(ns, 4), dtype=torch.float32, device=device)
        self.step = 0  # tf global_step (S3DIS_DGCNN_trainer.py:78)
        self.load(params)

    def _view(self, buf, offs, name):
        o, shape = offs[name]
        return buf[o:o + int(np.prod(shape))].view(*shape)

    def tail_offset(self, prefixes):
        """first element of the flat trainable buffer that belongs to a variable whose name starts with one of `prefixes`,
        provided those variables form the contiguous tail of the buffer (else None)"""
        offs = {k: o for k, (o, _) in self._toffs.items()}
        inside = [o for k, o in offs.items() if k.startswith(tuple(prefixes))]
        if not inside:
            return None
        t = min(inside)
        ok = all((o >= t) == k.startswith(tuple(prefixes)) for k, o in offs.items())
        return int(t) if ok and t > 0 else None

    def p(self, name):
        return self._view(self.theta, self._toffs, name)

    def g(self, name):
        return self._view(self.grad, self._toffs, name)

    def s(self, name):
        return self._view(self.state, self._soffs, name)

    def get(self, name):
        return self.p(name) if name in self._toffs else self.s(name)

    def load(self, params):
        for k in self.trainable_names:
            self.p(k).copy_(torch.as_tensor(np.asarray(params[k]), dtype=torch.float32))
        for k in self.state_names:
            self.s(k).copy_(torch.as_tensor(np.asarray(params[k]), dtype=torch.float32))

    def export(self) -> "OrderedDict[str, np.ndarray]":
        out = OrderedDict()
        for k in self.trainable_names:
            out[k] = self.p(k).detach().cpu().numpy().copy()
        for k in self.state_names:
            out[k] = self.s(k).detach().cpu().numpy().copy()
        return out

    def grads(self) -> "OrderedDict[str, np.ndarray]":
        return OrderedDict((k, self.g(k).detach().cpu().numpy().copy()) for k in self.trainable_names)

    def num_trainable(self) -> int:
        return sum(int(np.prod(s)) for _, s in self._toffs.values())

    def adam_step(self, lr, b1=0.9, b2=0.999, eps=1e-8, gscale=1.0):
        """tf.train.AdamOptimizer.apply_gradients [TF] on the flat buffers (one launch)."""
        self.step += 1
        t = self.step
        lr_t = lr * math.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t)
        L.check(L.lib().wspc_adam_tf(L.ptr(self.theta), L.ptr(self.grad), L.ptr(self.adam_m), L.ptr(self.adam_v),
                                     self.theta.numel(), lr_t, b1, b2, eps, gscale, L.stream()))


class Layer:
    """One conv2d-1x1 (+BN) layer: views of its variables/gradients plus per-channel scratch."""

    def __init__(self, vs: VariableStore, scope: str, cin: int, cout: int, has_bn: bool):
        self.scope, self.cin, self.cout, self.has_bn = scope, cin, cout, has_bn
        dev = vs.device
        self.W, self.b = vs.p(f"{scope}/weights"), vs.p(f"{scope}/biases")
        self.dW, self.db = vs.g(f"{scope}/weights"), vs.g(f"{scope}/biases")
        if has_bn:
            self.gamma, self.beta = vs.p(f"{scope}/bn/gamma"), vs.p(f"{scope}/bn/beta")
            self.dgamma, self.dbeta = vs.g(f"{scope}/bn/gamma"), vs.g(f"{scope}/bn/beta")
            self.pop_mean, self.pop_var = vs.s(f"{scope}/bn/pop_mean"), vs.s(f"{scope}/bn/pop_var")
            f = lambda: torch.empty(cout, dtype=torch.float32, device=dev)  # noqa: E731
            self.sc, self.sh, self.mean, self.invstd = f(), f(), f(), f()
            self.c1, self.c2, self.c3 = f(), f(), f()
            self.stats = torch.zeros((2, cout), dtype=torch.float64, device=dev)    # fwd: sum, sum^2
            self.bstats = torch.zeros((2, cout), dtype=torch.float64, device=dev)   # bwd: sum G, sum G*y


# ---- operand / epilogue builders ------------------------------------------------------------------
def op_plain(t, ld, C):
    return L.Operand(p=L.dptr(t), ld=ld, C=C), L.OP_PLAIN


# WSPC_ROWIMG=off keeps the fp32 operand loaders for the layers that read the concatenated EdgeConv features (A/B tests)
ROW_IMAGE = os.environ.get("WSPC_ROWIMG", "on") != "off"


class RowImage:
    """Pre-split (bf16 hi / lo, tensor-core tile layout) copy of a (M, K) matrix that several GEMMs read: written once per
    step by wspc_rows_image, consumed as WSPC_OP_IMG with one bulk copy per chunk (no CUDA-core operand work)."""

    def __init__(self, M, K, device):
        self.M, self.K = M, K
        self.buf = torch.empty(L.lib().wspc_rows_image_bytes(M, K), dtype=torch.uint8, device=device)

    @staticmethod
    def usable(M, K, couts):
        return ROW_IMAGE and all(L.lib().wspc_rows_image_supported(M, n, K) for n in couts)

    def build(self, x, ld):
        L.check(L.lib().wspc_rows_image(L.ptr(x), ld, self.M, self.K, L.ptr(self.buf), L.stream()))

    def operand(self):
        return L.Operand(p=L.dptr(self.buf), ld=self.K, C=self.K), L.OP_IMG


def op_bnrelu(y, layer: Layer, dmask=None, keep=1.0):
    return (L.Operand(p=L.dptr(y), ld=layer.cout, C=layer.cout, sc=L.dptr(layer.sc), sh=L.dptr(layer.sh),
                      dmask=L.dptr(dmask), dscale=1.0 / keep), L.OP_BNRELU)


def op_edge(x, ld, cx, idx, k, npts):
    return L.Operand(p=L.dptr(x), ld=ld, C=2 * cx, idx=L.dptr(idx), k=k, npts=npts), L.OP_EDGE


def op_dy(G, ldg, y, ldy, layer: Layer | None, C):
    """upstream gradient through the layer's BN backward (layer=None or no BN: identity)."""
    if layer is not None and layer.has_bn:
        return (L.Operand(p=L.dptr(G), ld=ldg, C=C, y=L.dptr(y), ldy=ldy, c1=L.dptr(layer.c1), c2=L.dptr(layer.c2),
                          c3=L.dptr(layer.c3)), L.OP_DY)
    return L.Operand(p=L.dptr(G), ld=ldg, C=C), L.OP_DY


def op_dy_sparse(y, layer: Layer, dg, amax, npts):
    return (L.Operand(p=0, ld=0, C=layer.cout, y=L.dptr(y), ldy=layer.cout, c1=L.dptr(layer.c1), c2=L.dptr(layer.c2),
                      c3=L.dptr(layer.c3), dg=L.dptr(dg), amax=L.dptr(amax), npts=npts), L.OP_DY_SPARSE)


def _addr(t, col_off=0):
    return 0 if t is None else t.data_ptr() + 4 * col_off


def rows_gemm(A, Bm, ldb, bT, M, N, K, epi: L.Epilogue, epi_mode):
    a, amode = A
    nbytes = L.lib().wspc_conv1x1_rows_workspace_bytes(N, K)       # pre-split weight image for chunked K (K > 128)
    ws = L.workspace(nbytes, torch.cuda.current_device(), "gemm_w") if nbytes else None
    L.check(L.lib().wspc_conv1x1_rows_ws(ctypes.byref(a), amode, L.ptr(Bm) if torch.is_tensor(Bm) else Bm, ldb, bT, M, N, K,
                                         ctypes.byref(epi), epi_mode, L.ptr(ws), ws.numel() if ws is not None else 0,
                                         L.stream()))


def wgrad(A, G, M, dW, db, device):
    a, amode = A
    g, gmode = G
    nbytes = L.lib().wspc_conv1x1_wgrad_workspace_bytes(a.C, g.C)
    ws = L.workspace(nbytes, device, "wgrad")
    L.check(L.lib().wspc_conv1x1_wgrad(ctypes.byref(a), amode, ctypes.byref(g), gmode, M,
                                       dW if isinstance(dW, ctypes.c_void_p) else L.ptr(dW), L.ptr(db), L.ptr(ws),
                                       ws.numel(), L.stream()))


# WSPC_BWD=split runs the weight gradient and the data gradient of a conv2d as two kernels (A/B tests)
BWD_FUSED = os.environ.get("WSPC_BWD", "fused") != "split"


def conv_bwd(A, G, M, layer: Layer, epi_pair, device):
    """weight gradient (layer.dW, layer.db) and data gradient (through the RELUMASK epilogue `epi_pair` from epi_relumask)
    of one conv2d: A = the layer's input operand, G = gradient w.r.t. its pre-BN output."""
    e, m = epi_pair
    if not BWD_FUSED:
        wgrad(A, G, M, layer.dW, layer.db, device)
        rows_gemm(G, layer.W, layer.cout, 1, M, layer.cin, layer.cout, e, m)
        return
    a, amode = A
    g, gmode = G
    nbytes = L.lib().wspc_conv1x1_wgrad_workspace_bytes(a.C, g.C)
    ws = L.workspace(nbytes, device, "wgrad")
    L.check(L.lib().wspc_conv1x1_bwd_fused(ctypes.byref(a), amode, ctypes.byref(g), gmode, M, L.ptr(layer.W), layer.cout,
                                           ctypes.byref(e), L.ptr(layer.dW), L.ptr(layer.db), L.ptr(ws), ws.numel(), L.stream()))


def zero_(t):
    L.check(L.lib().wspc_zero(L.ptr(t), t.numel() * t.element_size(), L.stream()))


def conv_forward(layer: Layer, A, M, y_out, ldo, training, decay, rowbias=None, rb_rows=1, Wview=None):
    """y_out = A @ W + b (+ per-cloud row bias); accumulates BN batch statistics and folds them into
    (sc, sh) for the consumer (tf_util.py:160-169, 521-530)."""
    W = layer.W if Wview is None else Wview
    K = A[0].C
    epi = L.Epilogue(out=L.dptr(y_out), ldo=ldo, bias=L.dptr(layer.b), rowbias=L.dptr(rowbias), rb_rows=rb_rows,
                     ldrb=layer.cout)
    if layer.has_bn and training:
        zero_(layer.stats)
        epi.stats = L.dptr(layer.stats)
        rows_gemm(A, W, layer.cout, 0, M, layer.cout, K, epi, L.EPI_STORE_STATS)
    else:
        rows_gemm(A, W, layer.cout, 0, M, layer.cout, K, epi, L.EPI_STORE)
    if layer.has_bn:
        bn_finalize(layer, M, training, decay)


# WSPC_POOL=unfused keeps the (P, 1024) adj_conv7 output in HBM and pools it with wspc_maxn_bnrelu_fwd (A/B tests)
POOL_FUSED = os.environ.get("WSPC_POOL", "fused") != "unfused"


def pool_fusable(M, N, K, npts):
    return POOL_FUSED and bool(L.lib().wspc_conv1x1_pool_supported(M, N, K, npts))


def conv_pool_forward(layer: Layer, A, M, npts, keys, g, amax, ymax, training, decay):
    """conv2d 1x1 -> BN -> ReLU -> max over the npts points of every cloud (DGCNN_S3DIS.py:80-85) without the (M, cout)
    intermediate: BN sums and per-(cloud, channel) extreme rows in one GEMM pass, then g / arg-max / pre-BN value at the
    arg-max (what the backward gate needs) from the keys."""
    a, amode = A
    K, N = a.C, layer.cout
    zero_(layer.stats)
    nbytes = L.lib().wspc_conv1x1_rows_workspace_bytes(N, K)
    ws = L.workspace(nbytes, torch.cuda.current_device(), "gemm_w")
    L.check(L.lib().wspc_conv1x1_pool_fwd(ctypes.byref(a), amode, L.ptr(layer.W), N, M, N, K, npts, L.ptr(layer.b),
                                          L.ptr(layer.gamma), L.ptr(layer.stats), L.ptr(keys), L.ptr(ws), ws.numel(), L.stream()))
    bn_finalize(layer, M, training, decay)
    L.check(L.lib().wspc_maxn_from_keys(L.ptr(keys), L.ptr(layer.gamma), L.ptr(layer.sc), L.ptr(layer.sh), M // npts, N,
                                        L.ptr(g), L.ptr(amax), L.ptr(ymax), L.stream()))


def bn_finalize(layer: Layer, rows, training, decay):
    d = 0.9 if decay is None else decay   # tf_util.py:523
    L.check(L.lib().wspc_bn_finalize(L.ptr(layer.stats), layer.cout, float(rows), L.ptr(layer.gamma), L.ptr(layer.beta),
                                     BN_EPS, d, 1 if training else 0, L.ptr(layer.pop_mean), L.ptr(layer.pop_var),
                                     L.ptr(layer.sc), L.ptr(layer.sh), L.ptr(layer.mean), L.ptr(layer.invstd),
                                     L.stream()))


def bn_bwd_coeffs(layer: Layer, rows):
    L.check(L.lib().wspc_bn_bwd_coeffs(L.ptr(layer.bstats), layer.cout, float(rows), L.ptr(layer.gamma),
                                       L.ptr(layer.mean), L.ptr(layer.invstd), L.ptr(layer.c1), L.ptr(layer.c2),
                                       L.ptr(layer.c3), L.ptr(layer.dgamma), L.ptr(layer.dbeta), L.stream()))


def epi_relumask(out, prev: Layer, yprev, dmask=None, keep=1.0):
    """dA -> gradient w.r.t. the producing layer's BN output (ReLU mask, dropout) + its BN-backward sums."""
    zero_(prev.bstats)
    return (L.Epilogue(out=L.dptr(out), ldo=prev.cout, stats=L.dptr(prev.bstats), yprev=L.dptr(yprev), ldyp=prev.cout,
                       scp=L.dptr(prev.sc), shp=L.dptr(prev.sh), dmask=L.dptr(dmask), dscale=1.0 / keep),
            L.EPI_RELUMASK_STATS)


def epi_scatter(dx_addr, lddx, idx, k, npts):
    return L.Epilogue(dx=dx_addr, lddx=lddx, idx=L.dptr(idx), k=k, npts=npts), L.EPI_EDGE_SCATTER


def maxk_fwd(layer: Layer, y, P, k, out_addr, ldo):
    L.check(L.lib().wspc_maxk_bnrelu_fwd(L.ptr(y), L.ptr(layer.sc), L.ptr(layer.sh), P, k, layer.cout,
                                         ctypes.c_void_p(out_addr), ldo, L.stream()))


# Gradient of the max over k: by default the (P*k, C) tensor G is materialised (wspc_maxk_bnrelu_bwd).  WSPC_MAXK=synth
# keeps statistics only and lets the consumers synthesise G on load (WSPC_OP_DY_MAXK / wspc_edge_combine_bwd_maxk): 8 GB less
# HBM traffic per step at cfg-3 but more loader work -- measured equal (46.7 vs 46.5 ms), so it stays an option
# (tests/test_maxk_synth_gpu.py holds the two paths together).
MAXK_SYNTH = os.environ.get("WSPC_MAXK", "materialise") == "synth"


def maxk_bwd_stats(layer: Layer, y, P, k, out_addr, ldo, dout_addr, lddo, MS):
    """tf.reduce_max(axis=-2) backward without materialising G: MS (P, 2C) = [pooled max | dout / #ties] + BN sums."""
    zero_(layer.bstats)
    L.check(L.lib().wspc_maxk_bnrelu_bwd_stats(L.ptr(y), L.ptr(layer.sc), L.ptr(layer.sh), ctypes.c_void_p(out_addr), ldo,
                                               ctypes.c_void_p(dout_addr), lddo, P, k, layer.cout, L.ptr(MS),
                                               L.ptr(layer.bstats), L.stream()))


def op_dy_maxk(layer: Layer, y, MS, k, npts):
    """gradient w.r.t. the layer's pre-BN output when its activation feeds a max over k (after bn_bwd_coeffs)."""
    C = layer.cout
    return (L.Operand(p=L.dptr(MS), ld=2 * C, C=C, sc=L.dptr(layer.sc), sh=L.dptr(layer.sh), k=k, npts=npts, y=L.dptr(y), ldy=C,
                      c1=L.dptr(layer.c1), c2=L.dptr(layer.c2), c3=L.dptr(layer.c3)), L.OP_DY_MAXK)


def maxk_bwd(layer: Layer, y, P, k, out_addr, ldo, dout_addr, lddo, G):
    zero_(layer.bstats)
    L.check(L.lib().wspc_maxk_bnrelu_bwd(L.ptr(y), L.ptr(layer.sc), L.ptr(layer.sh), ctypes.c_void_p(out_addr), ldo,
                                         ctypes.c_void_p(dout_addr), lddo, P, k, layer.cout, L.ptr(G),
                                         L.ptr(layer.bstats), L.stream()))


# ---- factored first layer of an EdgeConv block (csrc/edge.cu) ---------------------------------------
# WSPC_EDGE=gemm keeps the gathered-operand GEMM formulation (another device kernel, used for A/B tests)
EDGE_FACTORED = os.environ.get("WSPC_EDGE", "factored") != "gemm"


class EdgeSplit:
    """Scratch shared by the first conv2d of every EdgeConv block of one engine:
    y_ij = x_i (W1 - W2) + x_j W2 + b  (tf_util.py:696-705 feeding tf_util.conv2d, e.g. DGCNN_S3DIS.py:34-39)."""

    def __init__(self, P, device, max_cx=64):
        f32 = dict(dtype=torch.float32, device=device)
        self.P = P
        self.UV = torch.empty((P, 128), **f32)       # [u | v] forward
        self.DUV = torch.empty((P, 128), **f32)      # [du | dv] backward
        self.dWc = torch.empty((max_cx, 128), **f32)
        self.dbc = torch.empty(128, **f32)
        self.Wc = {}                                 # per layer: [W1 - W2 | W2] of the current step

    def weights(self, layer: Layer, cx):
        w = self.Wc.get(layer.scope)
        if w is None:
            w = self.Wc[layer.scope] = torch.empty((cx, 128), dtype=torch.float32, device=self.UV.device)
        L.check(L.lib().wspc_edge_split_weights(L.ptr(layer.W), cx, layer.cout, L.ptr(w), L.stream()))
        return w


def edge_first_forward(es: EdgeSplit, layer: Layer, x, ld, cx, idx, k, npts, P, y_out, training, decay, pool_out=None,
                       pool_ld=0):
    """x: tensor or raw address of the (P, ld) point features (cx channels used).  Writes pre-BN y_out (P*k, 64) and
    the layer's batch-norm scale/shift, exactly like conv_forward(op_edge(...)).  pool_out (address, leading dimension
    pool_ld): also write max over k of relu(bn(y)) there (a block whose only conv feeds tf.reduce_max: no second read of y)."""
    assert layer.cout == 64 and layer.cin == 2 * cx
    Wc = es.weights(layer, cx)
    xa = x if isinstance(x, int) else x.data_ptr()
    A = (L.Operand(p=xa, ld=ld, C=cx), L.OP_PLAIN)
    rows_gemm(A, Wc, 128, 0, P, 128, cx, L.Epilogue(out=L.dptr(es.UV), ldo=128), L.EPI_STORE)
    stats = None
    if layer.has_bn and training:
        zero_(layer.stats)
        stats = layer.stats
    mm = es.DUV if pool_out is not None else None        # (P, 128) scratch, free during the forward pass
    L.check(L.lib().wspc_edge_combine_fwd_extrema(L.ptr(es.UV), 128, L.ptr(idx), L.ptr(layer.b), P, k, npts, 64, L.ptr(y_out),
                                                  L.ptr(stats), L.ptr(mm), L.stream()))
    if layer.has_bn:
        bn_finalize(layer, P * k, training, decay)
    if pool_out is not None:
        L.check(L.lib().wspc_maxk_from_extrema(L.ptr(mm), L.ptr(layer.sc), L.ptr(layer.sh), P, 64, ctypes.c_void_p(pool_out),
                                               pool_ld, L.stream()))


def edge_first_backward(es: EdgeSplit, layer: Layer, x, ld, cx, idx, k, npts, P, G, y, dx_addr=None, lddx=0, MS=None):
    """Gradients of the factored layer: layer.dW / layer.db, and (if dx_addr) dX accumulated into (P, lddx) at dx_addr.
    G is the gradient w.r.t. the BN output (bn_bwd_coeffs(layer, P*k) must have run), y the saved pre-BN output;
    with MS (from maxk_bwd_stats) G is None and synthesised from the max over k."""
    xa = x if isinstance(x, int) else x.data_ptr()
    Wc = es.Wc[layer.scope]
    zero_(es.DUV)
    bn = layer.has_bn
    if MS is not None:
        L.check(L.lib().wspc_edge_combine_bwd_maxk(L.ptr(y), L.ptr(layer.c1), L.ptr(layer.c2), L.ptr(layer.c3), L.ptr(layer.sc),
                                                   L.ptr(layer.sh), L.ptr(MS), L.ptr(idx), P, k, npts, 64, L.ptr(es.DUV), 128,
                                                   L.stream()))
    else:
        L.check(L.lib().wspc_edge_combine_bwd(L.ptr(G), L.ptr(y) if bn else None, L.ptr(layer.c1) if bn else None,
                                              L.ptr(layer.c2) if bn else None, L.ptr(layer.c3) if bn else None, L.ptr(idx), P,
                                              k, npts, 64, L.ptr(es.DUV), 128, L.stream()))
    D = (L.Operand(p=L.dptr(es.DUV), ld=128, C=128), L.OP_DY)
    A = (L.Operand(p=xa, ld=ld, C=cx), L.OP_PLAIN)
    wgrad(A, D, P, es.dWc, es.dbc, es.UV.device)
    L.check(L.lib().wspc_edge_merge_wgrad(L.ptr(es.dWc), L.ptr(es.dbc), cx, 64, L.ptr(layer.dW), L.ptr(layer.db), L.stream()))
    if dx_addr is not None:
        rows_gemm(D, Wc, 128, 1, P, cx, 128, L.Epilogue(out=dx_addr, ldo=lddx), L.EPI_ACCUM)


# ---- backward of conv -> BN -> ReLU -> max over N through the Gram identity (csrc/poolconv.cu) ----------
# WSPC_POOLCONV=dense keeps the (P, cout) formulation (dy synthesised from y through OP_DY_SPARSE) for A/B tests
POOLCONV_GRAM = os.environ.get("WSPC_POOLCONV", "gram") != "dense"


class PoolConv:
    """Scratch for one `conv2d 1x1 -> BN -> ReLU -> max_pool2d([N,1])` layer (adj_conv7 / tconv3)."""

    def __init__(self, layer: Layer, device):
        cin, cout = layer.cin, layer.cout
        f32 = dict(dtype=torch.float32, device=device)
        self.layer = layer
        self.t, self.sdb = torch.empty(cout, **f32), torch.empty(cout, **f32)
        self.Wsc, self.T, self.sW = (torch.empty((cin, cout), **f32) for _ in range(3))
        self.M, self.gram = torch.empty((cin, cin), **f32), torch.empty((cin, cin), **f32)
        self.r0, self.colsum = torch.empty(cin, **f32), torch.empty(cin, **f32)

    def prepare(self):
        """after bn_bwd_coeffs(layer): t, W diag(c3), the constant row r0 = W t of dA and M = W diag(c3) W^T.
        Returns r0 (cin) -- the caller adds it to every row of dA (e.g. as the bias of the GEMM that first writes dA)."""
        ly = self.layer
        cin, cout = ly.cin, ly.cout
        L.check(L.lib().wspc_poolconv_coeffs(L.ptr(ly.W), L.ptr(ly.b), L.ptr(ly.c2), L.ptr(ly.c3), cin, cout, L.ptr(self.t),
                                             L.ptr(self.Wsc), L.stream()))
        rows_gemm(op_plain(self.t, cout, cout), ly.W, cout, 1, 1, cin, cout, L.Epilogue(out=L.dptr(self.r0), ldo=cin),
                  L.EPI_STORE)
        rows_gemm(op_plain(self.Wsc, cout, cout), ly.W, cout, 1, cin, cin, cout, L.Epilogue(out=L.dptr(self.M), ldo=cin),
                  L.EPI_STORE)
        return self.r0

    def backward(self, A, lda, P, B, N, dg, amax, dx_addr, lddx):
        """A: (P, lda) layer input (tensor or address).  Accumulates A M + sparse(c1 G) W^T into dx (r0 excluded, see
        prepare) and writes layer.dW / layer.db."""
        ly = self.layer
        cin, cout = ly.cin, ly.cout
        aa = A if isinstance(A, int) else A.data_ptr()
        Aop = (L.Operand(p=aa, ld=lda, C=cin), L.OP_PLAIN)
        dev = self.t.device
        rows_gemm(Aop, self.M, cin, 0, P, cin, cin, L.Epilogue(out=dx_addr, ldo=lddx), L.EPI_ACCUM)
        L.check(L.lib().wspc_poolconv_sparse(L.ptr(dg), L.ptr(amax), L.ptr(ly.c1), L.ptr(ly.W), ctypes.c_void_p(aa), lda, B, N,
                                             cin, cout, ctypes.c_void_p(dx_addr), lddx, L.ptr(self.sW), L.ptr(self.sdb),
                                             L.stream()))
        wgrad(Aop, (L.Operand(p=aa, ld=lda, C=cin), L.OP_DY), P, self.gram, self.colsum, dev)           # A^T A, A^T 1
        rows_gemm(op_plain(self.gram, cin, cin), self.Wsc, cout, 0, cin, cout, cin, L.Epilogue(out=L.dptr(self.T), ldo=cout),
                  L.EPI_STORE)
        L.check(L.lib().wspc_poolconv_finalize(L.ptr(self.T), L.ptr(self.colsum), L.ptr(self.t), L.ptr(self.sW), L.ptr(self.sdb),
                                               L.ptr(self.Wsc), cin, cout, float(P), L.ptr(ly.dW), L.ptr(ly.db), L.stream()))


# ---- fused EdgeConv blocks (csrc/edgeconv.cu): no (P*k, C) tensor in HBM ------------------------------------
# WSPC_EDGECONV=unfused keeps the round-1 formulation (pre-BN outputs y1..y5 and the max-over-k gradient materialised)
# as a second device path for A/B tests; both need the factored first layer.
EDGE_FUSED = EDGE_FACTORED and os.environ.get("WSPC_EDGECONV", "fused") != "unfused"


class EdgeBlockState:
    """What one EdgeConv block keeps between its forward and its backward pass (all per point, never per edge)."""

    def __init__(self, P, device):
        f32 = dict(dtype=torch.float32, device=device)
        self.UV = torch.empty((P, 128), **f32)      # [u | v] = X [W1 - W2 | W2]
        self.SS = torch.empty((P, 128), **f32)      # [S_i = sum_j v_j | SU_p = sum_{i->p} u_i]
        self.MM = torch.empty((P, 128), **f32)      # [max_j y | min_j y] of the block's last pre-BN output
        self.deg = torch.empty((P,), **f32)         # in-degree of every point in the kNN graph


class EdgeFused:
    """Scratch shared by the EdgeConv blocks of one engine."""

    def __init__(self, P, device, max_cx=64):
        f32 = dict(dtype=torch.float32, device=device)
        self.P, self.device = P, device
        self.TS = torch.empty((P, 128), **f32)      # [SG_i = sum_j g_ij | TG_p = sum_{i->p} g_ij]
        self.MS = torch.empty((P, 128), **f32)      # [out > 0 ? out : -1 | gated dout] of the max over k
        self.DUV = torch.empty((P, 128), **f32)     # [du | dv]
        self.dWc = torch.empty((max_cx, 128), **f32)
        self.dbc = torch.empty(128, **f32)
        self.Wc = {}

    def weights(self, layer: Layer, cx):
        w = self.Wc.get(layer.scope)
        if w is None:
            w = self.Wc[layer.scope] = torch.empty((cx, 128), dtype=torch.float32, device=self.device)
        L.check(L.lib().wspc_edge_split_weights(L.ptr(layer.W), cx, layer.cout, L.ptr(w), L.stream()))
        return w


PROF = None      # optional list of (tag, start_event, end_event): bench.py times the tcgen05 EdgeConv kernels inside the step
ROUTING = None   # optional dict (parity tests): the backward kernels export which rows attained each max over k, keyed by the
                 # scope of the block's last conv layer ((P*k, 2) int32 bit words for two-conv blocks, (P, 64) int64 for one-conv)


class _timed:
    """CUDA events around one launch on the current stream when PROF is a list; free otherwise."""

    def __init__(self, tag):
        self.tag = tag

    def __enter__(self):
        self.e0 = None
        if PROF is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *exc):
        if self.e0 is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            PROF.append((self.tag, self.e0, e1))
        return False


def _zero_cols(t, col0, ncols):
    L.check(L.lib().wspc_zero_cols(L.ptr(t), t.shape[1], col0, ncols, t.shape[0], L.stream()))


def edgeblock_forward(ef: EdgeFused, st: EdgeBlockState, l1: Layer, l2, x, ld, cx, idx, k, npts, P, training, decay,
                      out_addr, out_ld):
    """get_edge_feature -> conv2d(l1) [-> conv2d(l2)] -> reduce_max over k, written to (P, out_ld) at out_addr.
    x: tensor or raw address of the (P, ld) point features (cx channels used); l2 = None for a single-conv block."""
    assert l1.cout == 64 and l1.cin == 2 * cx and l1.has_bn and (l2 is None or (l2.cin == 64 and l2.cout == 64 and l2.has_bn))
    lib = L.lib()
    Wc = ef.weights(l1, cx)
    xa = x if isinstance(x, int) else x.data_ptr()
    rows_gemm((L.Operand(p=xa, ld=ld, C=cx), L.OP_PLAIN), Wc, 128, 0, P, 128, cx, L.Epilogue(out=L.dptr(st.UV), ldo=128),
              L.EPI_STORE)
    R = P * k
    single = l2 is None
    if training:
        zero_(l1.stats)
        _zero_cols(st.SS, 64, 64)
        zero_(st.deg)
        L.check(lib.wspc_edge_gather_stats(L.ptr(st.UV), 128, L.ptr(idx), L.ptr(l1.b), P, k, npts, 64, L.ptr(l1.stats),
                                           L.ptr(st.MM) if single else None, L.ptr(st.SS), L.ptr(st.deg), L.stream()))
    elif single:
        L.check(lib.wspc_edge_gather_stats(L.ptr(st.UV), 128, L.ptr(idx), L.ptr(l1.b), P, k, npts, 64, None, L.ptr(st.MM), None,
                                           None, L.stream()))
    bn_finalize(l1, R, training, decay)
    last = l1
    if not single:
        if training:
            zero_(l2.stats)
        with _timed("edgeconv2_fwd"):
            L.check(lib.wspc_edgeconv2_fwd(L.ptr(st.UV), 128, L.ptr(idx), L.ptr(l1.b), L.ptr(l1.sc), L.ptr(l1.sh), L.ptr(l2.W),
                                           L.ptr(l2.b), P, k, npts, 64, 64, L.ptr(l2.stats) if training else None,
                                           L.ptr(st.MM), L.stream()))
        bn_finalize(l2, R, training, decay)
        last = l2
    L.check(lib.wspc_maxk_from_extrema(L.ptr(st.MM), L.ptr(last.sc), L.ptr(last.sh), P, 64, ctypes.c_void_p(out_addr), out_ld,
                                       L.stream()))


def edgeblock_backward(ef: EdgeFused, st: EdgeBlockState, l1: Layer, l2, x, ld, cx, idx, k, npts, P, out_addr, out_ld,
                       dout_addr, dout_ld, dx_addr=None, lddx=0):
    """Gradients of every variable of the block (l1, l2: dW, db, dgamma, dbeta) and, if dx_addr, dX accumulated into
    (P, lddx) at dx_addr.  (out, dout): the block's pooled output and the gradient w.r.t. it."""
    lib = L.lib()
    R = P * k
    xa = x if isinstance(x, int) else x.data_ptr()
    out_p, dout_p = ctypes.c_void_p(out_addr), ctypes.c_void_p(dout_addr)
    _zero_cols(ef.TS, 64, 64)
    if l2 is None:
        rout = None
        if ROUTING is not None:
            rout = ROUTING[l1.scope] = torch.zeros((P, 64), dtype=torch.int64, device=ef.device)
        L.check(lib.wspc_edge1_bwd_ex(L.ptr(st.UV), 128, L.ptr(idx), L.ptr(l1.b), L.ptr(l1.sc), L.ptr(l1.sh), out_p, out_ld, dout_p,
                                      dout_ld, P, k, npts, 64, L.ptr(ef.TS), L.ptr(rout), L.stream()))
    else:
        zero_(l2.bstats)
        L.check(lib.wspc_maxk_extrema_bwd_prep(L.ptr(st.MM), L.ptr(l2.sc), out_p, out_ld, dout_p, dout_ld, P, 64, L.ptr(ef.MS),
                                               L.ptr(l2.bstats), L.stream()))
        bn_bwd_coeffs(l2, R)
        L.check(lib.wspc_bn_bias_grad(L.ptr(l2.stats), L.ptr(l2.bstats), L.ptr(l2.c1), L.ptr(l2.c2), L.ptr(l2.c3), 64, float(R),
                                      L.ptr(l2.db), L.stream()))
        nbytes = lib.wspc_edgeconv2_bwd_workspace_bytes()
        ws = L.workspace(nbytes, ef.device, "edgeconv_bwd")
        rout = None
        if ROUTING is not None:
            rout = ROUTING[l2.scope] = torch.zeros((P * k, 2), dtype=torch.int32, device=ef.device)
        with _timed("edgeconv2_bwd"):
            L.check(lib.wspc_edgeconv2_bwd_ex(L.ptr(st.UV), 128, L.ptr(idx), L.ptr(l1.b), L.ptr(l1.sc), L.ptr(l1.sh), L.ptr(l2.W),
                                              L.ptr(l2.b), L.ptr(l2.sc), L.ptr(l2.sh), L.ptr(l2.c1), L.ptr(l2.c2), L.ptr(l2.c3),
                                              L.ptr(ef.MS), P, k, npts, 64, 64, L.ptr(ef.TS), L.ptr(l2.dW), L.ptr(rout), L.ptr(ws),
                                              ws.numel(), L.stream()))
    zero_(l1.bstats)
    L.check(lib.wspc_edge_bwd_stats(L.ptr(ef.TS), L.ptr(st.UV), 128, L.ptr(l1.b), P, 64, L.ptr(l1.bstats), L.stream()))
    bn_bwd_coeffs(l1, R)
    L.check(lib.wspc_edge_bwd_finalize(L.ptr(ef.TS), L.ptr(st.SS), L.ptr(st.deg), L.ptr(st.UV), 128, L.ptr(l1.b), L.ptr(l1.c1),
                                       L.ptr(l1.c2), L.ptr(l1.c3), P, k, 64, L.ptr(ef.DUV), 128, L.stream()))
    Wc = ef.Wc[l1.scope]
    D = (L.Operand(p=L.dptr(ef.DUV), ld=128, C=128), L.OP_DY)
    A = (L.Operand(p=xa, ld=ld, C=cx), L.OP_PLAIN)
    wgrad(A, D, P, ef.dWc, ef.dbc, ef.device)
    L.check(lib.wspc_edge_merge_wgrad(L.ptr(ef.dWc), L.ptr(ef.dbc), cx, 64, L.ptr(l1.dW), L.ptr(l1.db), L.stream()))
    if dx_addr is not None:
        rows_gemm(D, Wc, 128, 1, P, cx, 128, L.Epilogue(out=dx_addr, ldo=lddx), L.EPI_ACCUM)
