"""Fused forward / loss / backward / Adam executor for the ShapeNet part-segmentation DGCNN.

Mirrors DGCNN_ShapeNet.get_model (ShapeNet/DGCNN_ShapeNet.py:15-113) with its spatial transformer
(Networks/dgcnn/models/transform_nets.py:10-56), the loss block of ShapeNet_DGCNN_trainer.defineNetwork /
WeakSupLoss (ShapeNet/ShapeNet_DGCNN_trainer.py:85-100,115-133) and tf.train.AdamOptimizer (:105).
Same kernel set and HBM layout conventions as engine_s3dis.py; additionally
  * the T-net (EdgeConv 6->64->128, max-k, 128->1024, max-N, FC 512/256, 256->9 + I) and X' = X T,
  * the one-hot category branch (16->64, BN over the batch) whose output joins the 1024-wide global feature;
    the tiled 1088-wide vector is folded into seg/conv1 as a per-cloud bias:
      seg/conv1 = cat @ W[1088:] + (g @ W[:1024] + lab @ W[1024:1088])[cloud]        (:92-101)
  * two dropouts (keep 0.6) fused into the operand loads of seg/conv2 and seg/conv3 (:102,:105).
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib as L
from . import runtime as rt
from .runtime import Layer, VariableStore

K_NEIGH = 20
SMOOTH_KNN = 10
SMOOTH_GAMMA = 0.1
SIAMESE_W = 1.0        # ShapeNet_DGCNN_trainer.py:123-124
KEEP = 0.6             # DGCNN_ShapeNet.py:102,105

LAYERS = [("transform_net1/tconv1", 6, 64, True), ("transform_net1/tconv2", 64, 128, True),
          ("transform_net1/tconv3", 128, 1024, True), ("transform_net1/tfc1", 1024, 512, True),
          ("transform_net1/tfc2", 512, 256, True),
          ("adj_conv1", 6, 64, True), ("adj_conv2", 64, 64, True), ("adj_conv3", 128, 64, True),
          ("adj_conv4", 64, 64, True), ("adj_conv5", 128, 64, True), ("adj_conv7", 192, 1024, True),
          ("one_hot_label_expand", 16, 64, True),
          ("seg/conv1", 1280, 256, True), ("seg/conv2", 256, 256, True), ("seg/conv3", 256, 128, True),
          ("seg/conv4", 128, 50, False)]


class ShapeNetEngine:
    def __init__(self, params, B, N, device="cuda:0", k=K_NEIGH, cat_num=16, part_num=50):
        self.B, self.N, self.k, self.C = B, N, k, part_num
        self.P, self.R = B * N, B * N * k
        self.dev = torch.device(device)
        self.vs = VariableStore(params, self.dev)
        vs = self.vs
        self.layers = {s: Layer(vs, s, ci, co, bn) for s, ci, co, bn in LAYERS}
        # transform_XYZ: 256 -> 9, no BN (transform_nets.py:42-55)
        self.tx = Layer.__new__(Layer)
        self.tx.scope, self.tx.cin, self.tx.cout, self.tx.has_bn = "transform_net1/transform_XYZ", 256, 9, False
        self.tx.W, self.tx.b = vs.p("transform_net1/transform_XYZ/weights"), vs.p("transform_net1/transform_XYZ/biases")
        self.tx.dW, self.tx.db = vs.g("transform_net1/transform_XYZ/weights"), vs.g("transform_net1/transform_XYZ/biases")
        f32 = dict(dtype=torch.float32, device=self.dev)
        i32 = dict(dtype=torch.int32, device=self.dev)
        P, R = self.P, self.R
        self.idx = [torch.empty((B, N, k), **i32) for _ in range(4)]      # knn0 (X), knn1 (X'), knn2, knn3
        self.idxS, self.dS = torch.empty((B, N, SMOOTH_KNN), **i32), torch.empty((B, N, SMOOTH_KNN), **f32)
        self.yt1, self.yt2 = torch.empty((R, 64), **f32), torch.empty((R, 128), **f32)
        self.Gt2 = torch.empty((R, 128), **f32)
        self.tmax, self.dtmax = torch.empty((P, 128), **f32), torch.empty((P, 128), **f32)
        self.yt3 = torch.empty((P, 1024), **f32)
        self.tg, self.dtg_in, self.dtg = (torch.empty((B, 1024), **f32) for _ in range(3))
        self.tamax = torch.empty((B, 1024), **i32)
        self.yf1, self.Gf1 = torch.empty((B, 512), **f32), torch.empty((B, 512), **f32)
        self.yf2, self.Gf2 = torch.empty((B, 256), **f32), torch.empty((B, 256), **f32)
        self.Tm, self.dTm = torch.empty((B, 9), **f32), torch.empty((B, 9), **f32)
        self.Xt, self.dXt = torch.empty((B, N, 3), **f32), torch.empty((B, N, 3), **f32)
        # the three EdgeConv blocks of the segmentation trunk run fused (csrc/edgeconv.cu, per-point state only); the T-net's
        # 64 -> 128 block keeps the materialised formulation (its second layer is 128 wide)
        self.fused = rt.EDGE_FUSED
        if self.fused:
            self.ef = rt.EdgeFused(P, self.dev)
            self.eb = [rt.EdgeBlockState(P, self.dev) for _ in range(3)]
            self.y, self.Gb = None, None
        else:
            self.y = [torch.empty((R, 64), **f32) for _ in range(5)]
            self.Gb = torch.empty((R, 64), **f32)
        self.Ga = torch.empty((R, 64), **f32)
        self.cat, self.dcat = torch.empty((P, 192), **f32), torch.empty((P, 192), **f32)
        self.pool_fused = rt.POOLCONV_GRAM and rt.pool_fusable(P, 1024, 192, N)
        self.cat_img = rt.RowImage(P, 192, self.dev) if rt.RowImage.usable(P, 192, (1024, 256)) else None
        if self.pool_fused:
            self.y7 = None
            self.pool_keys = torch.empty((B, 1024), dtype=torch.int64, device=self.dev)
            self.y7max = torch.empty((B, 1024), **f32)
            self.amax0 = torch.zeros((B, 1024), **i32)
        else:
            self.y7 = torch.empty((P, 1024), **f32)
        self.g, self.dg_in, self.dg = (torch.empty((B, 1024), **f32) for _ in range(3))
        self.amax = torch.empty((B, 1024), **i32)
        self.ylab, self.Glab = torch.empty((B, 64), **f32), torch.empty((B, 64), **f32)
        self.gW, self.S = torch.empty((B, 256), **f32), torch.empty((B, 256), **f32)
        self.ys1, self.Gs1 = torch.empty((P, 256), **f32), torch.empty((P, 256), **f32)
        self.ys2, self.Gs2 = torch.empty((P, 256), **f32), torch.empty((P, 256), **f32)
        self.ys3, self.Gs3 = torch.empty((P, 128), **f32), torch.empty((P, 128), **f32)
        self.dmask1, self.dmask2 = torch.empty((P, 256), **f32), torch.empty((P, 256), **f32)
        self.Z, self.Zp, self.dZ = (torch.empty((B, N, part_num), **f32) for _ in range(3))
        self.losses = torch.zeros(5, **f32)
        self.seed = 4321
        self.es = rt.EdgeSplit(self.P, self.dev) if rt.EDGE_FACTORED else None   # factored first EdgeConv layers (csrc/edge.cu)
        # conv -> BN -> ReLU -> max over N layers: backward through the Gram identity (csrc/poolconv.cu)
        self.pc7 = rt.PoolConv(self.layers["adj_conv7"], self.dev) if rt.POOLCONV_GRAM else None
        self.pct3 = rt.PoolConv(self.layers["transform_net1/tconv3"], self.dev) if rt.POOLCONV_GRAM else None
        self.prof = None

    def _knn(self, i, src_addr, ld, coff, D, ov, tag):
        B, N, k = self.B, self.N, self.k
        if ov.get(tag) is not None:
            self.idx[i].copy_(ov[tag])
            return
        ws = L.workspace(L.lib().wspc_knn_workspace_bytes(B, N, D), self.dev, "knn")
        with rt._timed(f"knn_D{D}_k{k}"):
            L.check(L.lib().wspc_knn_fused(ctypes.c_void_p(src_addr), B, N, ld, coff, D, k, L.DIST_TFUTIL, L.ptr(self.idx[i]),
                                           None, L.ptr(ws), ws.numel(), L.stream()))

    # -------------------------------------------------------------------------------- forward ----
    def forward(self, X, label_onehot, is_training, bn_decay=None, dropout_masks=None, knn_override=None):
        """DGCNN_ShapeNet.get_model. X (B,N,3), label_onehot (B,16) fp32 CUDA -> logits (B,N,50)."""
        B, N, k, P, R = self.B, self.N, self.k, self.P, self.R
        Ly, ov = self.layers, (knn_override or {})
        assert X.shape == (B, N, 3) and X.is_cuda and X.is_contiguous()
        self.X, self.label = X, label_onehot.contiguous()
        T = "transform_net1/"
        cat_a = self.cat.data_ptr()
        tr, d = is_training, bn_decay
        # ---- T-net on the edge feature of the raw cloud                      (:23-28, transform_nets.py)
        self._knn(0, X.data_ptr(), 3, 0, 3, ov, "knn0")
        if self.es is not None:
            rt.edge_first_forward(self.es, Ly[T + "tconv1"], X, 3, 3, self.idx[0], k, N, P, self.yt1, tr, d)
        else:
            rt.conv_forward(Ly[T + "tconv1"], rt.op_edge(X, 3, 3, self.idx[0], k, N), R, self.yt1, 64, tr, d)
        rt.conv_forward(Ly[T + "tconv2"], rt.op_bnrelu(self.yt1, Ly[T + "tconv1"]), R, self.yt2, 128, tr, d)
        rt.maxk_fwd(Ly[T + "tconv2"], self.yt2, P, k, self.tmax.data_ptr(), 128)
        t3 = Ly[T + "tconv3"]
        rt.conv_forward(t3, rt.op_plain(self.tmax, 128, 128), P, self.yt3, 1024, tr, d)
        L.check(L.lib().wspc_maxn_bnrelu_fwd(L.ptr(self.yt3), L.ptr(t3.sc), L.ptr(t3.sh), B, N, 1024, L.ptr(self.tg),
                                             L.ptr(self.tamax), L.stream()))
        f1, f2 = Ly[T + "tfc1"], Ly[T + "tfc2"]
        rt.conv_forward(f1, rt.op_plain(self.tg, 1024, 1024), B, self.yf1, 512, tr, d)
        rt.conv_forward(f2, rt.op_bnrelu(self.yf1, f1), B, self.yf2, 256, tr, d)
        rt.conv_forward(self.tx, rt.op_bnrelu(self.yf2, f2), B, self.Tm, 9, tr, None)
        L.check(L.lib().wspc_transform_points_fwd(L.ptr(X), L.ptr(self.Tm), B, N, 1, L.ptr(self.Xt), L.stream()))  # :29
        # ---- EdgeConv blocks on the transformed cloud                        (:31-78)
        if self.fused:
            c = [Ly[f"adj_conv{i}"] for i in (1, 2, 3, 4, 5)]
            self._knn(1, self.Xt.data_ptr(), 3, 0, 3, ov, "knn1")
            rt.edgeblock_forward(self.ef, self.eb[0], c[0], c[1], self.Xt, 3, 3, self.idx[1], k, N, P, tr, d, cat_a, 192)
            self._knn(2, cat_a, 192, 0, 64, ov, "knn2")
            rt.edgeblock_forward(self.ef, self.eb[1], c[2], c[3], cat_a, 192, 64, self.idx[2], k, N, P, tr, d, cat_a + 4 * 64, 192)
            self._knn(3, cat_a, 192, 64, 64, ov, "knn3")
            rt.edgeblock_forward(self.ef, self.eb[2], c[4], None, cat_a + 4 * 64, 192, 64, self.idx[3], k, N, P, tr, d,
                                 cat_a + 4 * 128, 192)
            return self._forward_head(tr, d, dropout_masks)
        self._knn(1, self.Xt.data_ptr(), 3, 0, 3, ov, "knn1")
        if self.es is not None:
            rt.edge_first_forward(self.es, Ly["adj_conv1"], self.Xt, 3, 3, self.idx[1], k, N, P, self.y[0], tr, d)
        else:
            rt.conv_forward(Ly["adj_conv1"], rt.op_edge(self.Xt, 3, 3, self.idx[1], k, N), R, self.y[0], 64, tr, d)
        rt.conv_forward(Ly["adj_conv2"], rt.op_bnrelu(self.y[0], Ly["adj_conv1"]), R, self.y[1], 64, tr, d)
        rt.maxk_fwd(Ly["adj_conv2"], self.y[1], P, k, cat_a, 192)
        self._knn(2, cat_a, 192, 0, 64, ov, "knn2")
        if self.es is not None:
            rt.edge_first_forward(self.es, Ly["adj_conv3"], cat_a, 192, 64, self.idx[2], k, N, P, self.y[2], tr, d)
        else:
            rt.conv_forward(Ly["adj_conv3"], rt.op_edge(self.cat, 192, 64, self.idx[2], k, N), R, self.y[2], 64, tr, d)
        rt.conv_forward(Ly["adj_conv4"], rt.op_bnrelu(self.y[2], Ly["adj_conv3"]), R, self.y[3], 64, tr, d)
        rt.maxk_fwd(Ly["adj_conv4"], self.y[3], P, k, cat_a + 4 * 64, 192)
        self._knn(3, cat_a, 192, 64, 64, ov, "knn3")
        if self.es is not None:
            rt.edge_first_forward(self.es, Ly["adj_conv5"], cat_a + 4 * 64, 192, 64, self.idx[3], k, N, P, self.y[4], tr, d,
                                  pool_out=cat_a + 4 * 128, pool_ld=192)
        else:
            e3 = L.Operand(p=cat_a + 4 * 64, ld=192, C=128, idx=L.dptr(self.idx[3]), k=k, npts=N), L.OP_EDGE
            rt.conv_forward(Ly["adj_conv5"], e3, R, self.y[4], 64, tr, d)
            rt.maxk_fwd(Ly["adj_conv5"], self.y[4], P, k, cat_a + 4 * 128, 192)
        return self._forward_head(tr, d, dropout_masks)

    def _forward_head(self, tr, d, dropout_masks):
        B, N, P = self.B, self.N, self.P
        Ly = self.layers
        l7 = Ly["adj_conv7"]
        cat_op = rt.op_plain(self.cat, 192, 192)
        if self.cat_img is not None:     # adj_conv7 and seg/conv1 both read the concatenated features: split them once
            self.cat_img.build(self.cat, 192)
            cat_op = self.cat_img.operand()
        if self.pool_fused:     # :80-85 in one pass, the (P, 1024) pre-BN tensor is never written
            rt.conv_pool_forward(l7, cat_op, P, N, self.pool_keys, self.g, self.amax, self.y7max, tr, d)
        else:
            rt.conv_forward(l7, cat_op, P, self.y7, 1024, tr, d)                                          # :80-83
            L.check(L.lib().wspc_maxn_bnrelu_fwd(L.ptr(self.y7), L.ptr(l7.sc), L.ptr(l7.sh), B, N, 1024, L.ptr(self.g),
                                                 L.ptr(self.amax), L.stream()))                           # :85
        # ---- category branch + folded global feature                         (:87-101)
        lab = Ly["one_hot_label_expand"]
        rt.conv_forward(lab, rt.op_plain(self.label, 16, 16), B, self.ylab, 64, tr, d)
        s1, s2, s3, s4 = (Ly[f"seg/conv{i}"] for i in (1, 2, 3, 4))
        rt.rows_gemm(rt.op_plain(self.g, 1024, 1024), s1.W, 256, 0, B, 256, 1024, L.Epilogue(out=L.dptr(self.gW), ldo=256),
                     L.EPI_STORE)
        rt.rows_gemm(rt.op_bnrelu(self.ylab, lab), s1.W[1024:], 256, 0, B, 256, 64,
                     L.Epilogue(out=L.dptr(self.gW), ldo=256), L.EPI_ACCUM)
        rt.conv_forward(s1, cat_op, P, self.ys1, 256, tr, d, rowbias=self.gW, rb_rows=N,
                        Wview=s1.W[1088:])
        # ---- dropouts fused into the next layer's operand load               (:102-109)
        self._m1 = self._m2 = None
        keep = KEEP if tr else 1.0
        if tr:
            if dropout_masks is not None:
                self.dmask1.copy_(dropout_masks[0].reshape(P, 256))
                self.dmask2.copy_(dropout_masks[1].reshape(P, 256))
            else:
                off = self.vs.step * (P * 128)
                L.check(L.lib().wspc_dropout_mask(L.ptr(self.dmask1), P * 256, KEEP, self.seed, off, L.stream()))
                L.check(L.lib().wspc_dropout_mask(L.ptr(self.dmask2), P * 256, KEEP, self.seed, off + P * 64, L.stream()))
            self._m1, self._m2 = self.dmask1, self.dmask2
        rt.conv_forward(s2, rt.op_bnrelu(self.ys1, s1, self._m1, keep), P, self.ys2, 256, tr, d)
        rt.conv_forward(s3, rt.op_bnrelu(self.ys2, s2, self._m2, keep), P, self.ys3, 128, tr, d)
        rt.conv_forward(s4, rt.op_bnrelu(self.ys3, s3), P, self.Z, self.C, tr, None)
        return self.Z

    # ---------------------------------------------------------------------------------- losses ---
    def losses_and_grad(self, Y, Mask, full=True, want_grad=True, smooth_graph=None):
        B, N, C = self.B, self.N, self.C
        if full:
            if smooth_graph is not None:
                self.idxS.copy_(smooth_graph[0])
                self.dS.copy_(smooth_graph[1])
            else:   # smooth term on X_ph xyz (ShapeNet_DGCNN_trainer.py:133)
                ws = L.workspace(L.lib().wspc_knn_workspace_bytes(B, N, 3), self.dev, "knn")
                with rt._timed(f"knn_D3_k{SMOOTH_KNN}"):
                    L.check(L.lib().wspc_knn_fused(L.ptr(self.X), B, N, 3, 0, 3, SMOOTH_KNN, L.DIST_SMOOTH, L.ptr(self.idxS),
                                                   L.ptr(self.dS), L.ptr(ws), ws.numel(), L.stream()))
        ws = L.workspace(L.lib().wspc_head_losses_workspace_bytes(B, N, C), self.dev, "head")
        L.check(L.lib().wspc_head_losses(L.ptr(self.Z), L.ptr(Y), L.ptr(Mask), L.ptr(self.idxS) if full else None,
                                         L.ptr(self.dS) if full else None, B, N, C, SMOOTH_KNN, SMOOTH_GAMMA, SIAMESE_W,
                                         int(full), 1 if want_grad else 0, L.ptr(self.Zp), L.ptr(self.dZ),
                                         L.ptr(self.losses), L.ptr(ws), ws.numel(), L.stream()))
        return self.losses

    # -------------------------------------------------------------------------------- backward ---
    def backward(self):
        B, N, k, P, R, C = self.B, self.N, self.k, self.P, self.R, self.C
        Ly, dev = self.layers, self.dev
        T = "transform_net1/"
        c1, c2, c3, c4, c5, c7 = (Ly[f"adj_conv{i}"] for i in (1, 2, 3, 4, 5, 7))
        s1, s2, s3, s4 = (Ly[f"seg/conv{i}"] for i in (1, 2, 3, 4))
        lab = Ly["one_hot_label_expand"]
        cat_a, dcat_a = self.cat.data_ptr(), self.dcat.data_ptr()
        # seg/conv4 (no BN)
        G4 = rt.op_dy(self.dZ, C, None, 0, None, C)
        rt.wgrad(rt.op_bnrelu(self.ys3, s3), G4, P, s4.dW, s4.db, dev)
        e, m = rt.epi_relumask(self.Gs3, s3, self.ys3)
        rt.rows_gemm(G4, s4.W, C, 1, P, 128, C, e, m)
        # seg/conv3
        rt.bn_bwd_coeffs(s3, P)
        G3 = rt.op_dy(self.Gs3, 128, self.ys3, 128, s3, 128)
        rt.wgrad(rt.op_bnrelu(self.ys2, s2, self._m2, KEEP), G3, P, s3.dW, s3.db, dev)
        e, m = rt.epi_relumask(self.Gs2, s2, self.ys2, self._m2, KEEP)
        rt.rows_gemm(G3, s3.W, 128, 1, P, 256, 128, e, m)
        # seg/conv2
        rt.bn_bwd_coeffs(s2, P)
        G2 = rt.op_dy(self.Gs2, 256, self.ys2, 256, s2, 256)
        rt.wgrad(rt.op_bnrelu(self.ys1, s1, self._m1, KEEP), G2, P, s2.dW, s2.db, dev)
        e, m = rt.epi_relumask(self.Gs1, s1, self.ys1, self._m1, KEEP)
        rt.rows_gemm(G2, s2.W, 256, 1, P, 256, 256, e, m)
        # seg/conv1: point part + folded [global | category] part
        rt.bn_bwd_coeffs(s1, P)
        G1 = rt.op_dy(self.Gs1, 256, self.ys1, 256, s1, 256)
        rt.wgrad(rt.op_plain(self.cat, 192, 192), G1, P, s1.dW[1088:], s1.db, dev)
        L.check(L.lib().wspc_cloud_colsum(ctypes.byref(G1[0]), B, N, L.ptr(self.S), L.stream()))
        GS = rt.op_dy(self.S, 256, None, 0, None, 256)
        rt.wgrad(rt.op_plain(self.g, 1024, 1024), GS, B, s1.dW[:1024], None, dev)
        rt.wgrad(rt.op_bnrelu(self.ylab, lab), GS, B, s1.dW[1024:1088], None, dev)
        rt.rows_gemm(GS, s1.W, 256, 1, B, 1024, 256, L.Epilogue(out=L.dptr(self.dg_in), ldo=1024), L.EPI_STORE)
        e, m = rt.epi_relumask(self.Glab, lab, self.ylab)
        rt.rows_gemm(GS, s1.W[1024:1088], 256, 1, B, 64, 256, e, m)
        self._G1 = G1     # the (P,192) data gradient is written after adj_conv7's coefficients are known (constant row r0)
        # category branch
        rt.bn_bwd_coeffs(lab, B)
        rt.wgrad(rt.op_plain(self.label, 16, 16), rt.op_dy(self.Glab, 64, self.ylab, 64, lab, 64), B, lab.dW, lab.db, dev)
        # max over points + adj_conv7
        rt.zero_(c7.bstats)
        y7_rows, y7_n, y7_amax = (self.y7max, 1, self.amax0) if self.pool_fused else (self.y7, N, self.amax)
        L.check(L.lib().wspc_maxn_bwd_gate(L.ptr(self.g), L.ptr(self.dg_in), L.ptr(y7_amax), L.ptr(y7_rows), B, y7_n, 1024,
                                           L.ptr(self.dg), L.ptr(c7.bstats), L.stream()))
        rt.bn_bwd_coeffs(c7, P)
        G1 = self._G1
        if self.pc7 is not None:
            r0 = self.pc7.prepare()
            rt.rows_gemm(G1, s1.W[1088:], 256, 1, P, 192, 256, L.Epilogue(out=dcat_a, ldo=192, bias=L.dptr(r0)), L.EPI_STORE)
            self.pc7.backward(cat_a, 192, P, B, N, self.dg, self.amax, dcat_a, 192)
        else:
            rt.rows_gemm(G1, s1.W[1088:], 256, 1, P, 192, 256, L.Epilogue(out=dcat_a, ldo=192), L.EPI_STORE)
            G7 = rt.op_dy_sparse(self.y7, c7, self.dg, self.amax, N)
            rt.wgrad(rt.op_plain(self.cat, 192, 192), G7, P, c7.dW, c7.db, dev)
            rt.rows_gemm(G7, c7.W, 1024, 1, P, 192, 1024, L.Epilogue(out=dcat_a, ldo=192), L.EPI_ACCUM)
        if self.fused:
            ef, eb = self.ef, self.eb
            rt.edgeblock_backward(ef, eb[2], c5, None, cat_a + 4 * 64, 192, 64, self.idx[3], k, N, P, cat_a + 4 * 128, 192,
                                  dcat_a + 4 * 128, 192, dcat_a + 4 * 64, 192)
            rt.edgeblock_backward(ef, eb[1], c3, c4, cat_a, 192, 64, self.idx[2], k, N, P, cat_a + 4 * 64, 192,
                                  dcat_a + 4 * 64, 192, dcat_a, 192)
            rt.zero_(self.dXt)
            rt.edgeblock_backward(ef, eb[0], c1, c2, self.Xt, 3, 3, self.idx[1], k, N, P, cat_a, 192, dcat_a, 192,
                                  self.dXt.data_ptr(), 3)
            return self._backward_tnet()
        # block 3
        rt.maxk_bwd(c5, self.y[4], P, k, cat_a + 4 * 128, 192, dcat_a + 4 * 128, 192, self.Ga)
        rt.bn_bwd_coeffs(c5, R)
        if self.es is not None:
            rt.edge_first_backward(self.es, c5, cat_a + 4 * 64, 192, 64, self.idx[3], k, N, P, self.Ga, self.y[4],
                                   dcat_a + 4 * 64, 192)
        else:
            G5 = rt.op_dy(self.Ga, 64, self.y[4], 64, c5, 64)
            A5 = L.Operand(p=cat_a + 4 * 64, ld=192, C=128, idx=L.dptr(self.idx[3]), k=k, npts=N), L.OP_EDGE
            rt.wgrad(A5, G5, R, c5.dW, c5.db, dev)
            e, m = rt.epi_scatter(dcat_a + 4 * 64, 192, self.idx[3], k, N)
            rt.rows_gemm(G5, c5.W, 64, 1, R, 128, 64, e, m)
        # block 2
        rt.maxk_bwd(c4, self.y[3], P, k, cat_a + 4 * 64, 192, dcat_a + 4 * 64, 192, self.Ga)
        rt.bn_bwd_coeffs(c4, R)
        G4e = rt.op_dy(self.Ga, 64, self.y[3], 64, c4, 64)
        rt.conv_bwd(rt.op_bnrelu(self.y[2], c3), G4e, R, c4, rt.epi_relumask(self.Gb, c3, self.y[2]), dev)
        rt.bn_bwd_coeffs(c3, R)
        if self.es is not None:
            rt.edge_first_backward(self.es, c3, cat_a, 192, 64, self.idx[2], k, N, P, self.Gb, self.y[2], dcat_a, 192)
        else:
            G3e = rt.op_dy(self.Gb, 64, self.y[2], 64, c3, 64)
            rt.wgrad(rt.op_edge(self.cat, 192, 64, self.idx[2], k, N), G3e, R, c3.dW, c3.db, dev)
            e, m = rt.epi_scatter(dcat_a, 192, self.idx[2], k, N)
            rt.rows_gemm(G3e, c3.W, 64, 1, R, 128, 64, e, m)
        # block 1: gradient continues into the transformed cloud (dX')
        rt.maxk_bwd(c2, self.y[1], P, k, cat_a, 192, dcat_a, 192, self.Ga)
        rt.bn_bwd_coeffs(c2, R)
        G2e = rt.op_dy(self.Ga, 64, self.y[1], 64, c2, 64)
        rt.conv_bwd(rt.op_bnrelu(self.y[0], c1), G2e, R, c2, rt.epi_relumask(self.Gb, c1, self.y[0]), dev)
        rt.bn_bwd_coeffs(c1, R)
        rt.zero_(self.dXt)
        if self.es is not None:
            rt.edge_first_backward(self.es, c1, self.Xt, 3, 3, self.idx[1], k, N, P, self.Gb, self.y[0], self.dXt.data_ptr(), 3)
        else:
            G1e = rt.op_dy(self.Gb, 64, self.y[0], 64, c1, 64)
            rt.wgrad(rt.op_edge(self.Xt, 3, 3, self.idx[1], k, N), G1e, R, c1.dW, c1.db, dev)
            e, m = rt.epi_scatter(self.dXt.data_ptr(), 3, self.idx[1], k, N)
            rt.rows_gemm(G1e, c1.W, 64, 1, R, 6, 64, e, m)
        return self._backward_tnet()

    def _backward_tnet(self):
        B, N, k, P, R = self.B, self.N, self.k, self.P, self.R
        Ly, dev = self.layers, self.dev
        T = "transform_net1/"
        # ---- T-net                                                          (DGCNN_ShapeNet.py:29 backward)
        L.check(L.lib().wspc_transform_points_bwd(L.ptr(self.X), L.ptr(self.dXt), B, N, L.ptr(self.dTm), L.stream()))
        t1, t2, t3, f1, f2 = [Ly[T + n] for n in ("tconv1", "tconv2", "tconv3", "tfc1", "tfc2")]
        tx = self.tx
        GT = rt.op_dy(self.dTm, 9, None, 0, None, 9)
        rt.wgrad(rt.op_bnrelu(self.yf2, f2), GT, B, tx.dW, tx.db, dev)
        e, m = rt.epi_relumask(self.Gf2, f2, self.yf2)
        rt.rows_gemm(GT, tx.W, 9, 1, B, 256, 9, e, m)
        rt.bn_bwd_coeffs(f2, B)
        Gf2 = rt.op_dy(self.Gf2, 256, self.yf2, 256, f2, 256)
        rt.wgrad(rt.op_bnrelu(self.yf1, f1), Gf2, B, f2.dW, f2.db, dev)
        e, m = rt.epi_relumask(self.Gf1, f1, self.yf1)
        rt.rows_gemm(Gf2, f2.W, 256, 1, B, 512, 256, e, m)
        rt.bn_bwd_coeffs(f1, B)
        Gf1 = rt.op_dy(self.Gf1, 512, self.yf1, 512, f1, 512)
        rt.wgrad(rt.op_plain(self.tg, 1024, 1024), Gf1, B, f1.dW, f1.db, dev)
        rt.rows_gemm(Gf1, f1.W, 512, 1, B, 1024, 512, L.Epilogue(out=L.dptr(self.dtg_in), ldo=1024), L.EPI_STORE)
        rt.zero_(t3.bstats)
        L.check(L.lib().wspc_maxn_bwd_gate(L.ptr(self.tg), L.ptr(self.dtg_in), L.ptr(self.tamax), L.ptr(self.yt3), B, N,
                                           1024, L.ptr(self.dtg), L.ptr(t3.bstats), L.stream()))
        rt.bn_bwd_coeffs(t3, P)
        if self.pct3 is not None:
            r0 = self.pct3.prepare()
            L.check(L.lib().wspc_fill_rows(L.ptr(self.dtmax), 128, L.ptr(r0), P, 128, L.stream()))     # dtmax = 1 r0^T
            self.pct3.backward(self.tmax, 128, P, B, N, self.dtg, self.tamax, self.dtmax.data_ptr(), 128)
        else:
            Gt3 = rt.op_dy_sparse(self.yt3, t3, self.dtg, self.tamax, N)
            rt.wgrad(rt.op_plain(self.tmax, 128, 128), Gt3, P, t3.dW, t3.db, dev)
            rt.rows_gemm(Gt3, t3.W, 1024, 1, P, 128, 1024, L.Epilogue(out=L.dptr(self.dtmax), ldo=128), L.EPI_STORE)
        rt.maxk_bwd(t2, self.yt2, P, k, self.tmax.data_ptr(), 128, self.dtmax.data_ptr(), 128, self.Gt2)
        rt.bn_bwd_coeffs(t2, R)
        Gt2 = rt.op_dy(self.Gt2, 128, self.yt2, 128, t2, 128)
        rt.conv_bwd(rt.op_bnrelu(self.yt1, t1), Gt2, R, t2, rt.epi_relumask(self.Ga, t1, self.yt1), dev)
        rt.bn_bwd_coeffs(t1, R)
        if self.es is not None:
            rt.edge_first_backward(self.es, t1, self.X, 3, 3, self.idx[0], k, N, P, self.Ga, self.yt1)
        else:
            rt.wgrad(rt.op_edge(self.X, 3, 3, self.idx[0], k, N), rt.op_dy(self.Ga, 64, self.yt1, 64, t1, 64), R, t1.dW,
                     t1.db, dev)

    def train_step(self, X, label, Y, Mask, lr, bn_decay, full=True, dropout_masks=None, knn_override=None,
                   smooth_graph=None, apply=True, gscale=1.0, allreduce=None):
        self.forward(X, label, True, bn_decay, dropout_masks, knn_override)
        self.losses_and_grad(Y, Mask, full, True, smooth_graph)
        self.backward()
        if allreduce is not None:
            allreduce(self.vs.grad)
        if apply:
            self.vs.adam_step(lr, gscale=gscale)
        return self.losses
