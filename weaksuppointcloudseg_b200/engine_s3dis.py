"""Fused forward / loss / backward / Adam executor for the S3DIS segmentation DGCNN.

Mirrors, launch by launch, what one `sess.run([solver, loss, ...])` of the reference executes
(S3DIS/S3DIS_DGCNN_trainer.py:317-323): the graph of DGCNN_S3DIS.get_model (S3DIS/DGCNN_S3DIS.py:24-104),
the loss block of defineNetwork/WeakSupLoss (:85-102, :120-137) and tf.train.AdamOptimizer (:110).
Data layout in HBM (P = B*N points, R = P*k edges, fp32):
  y1..y5  (R,64)   pre-BN outputs of adj_conv1..5 (kept for the backward pass; 180 GB HBM makes
                   recomputation unnecessary)
  cat     (P,192)  [net_1 | net_2 | net_3]; each max-over-k writes its slice, kNN 2/3 read theirs in place
  y7      (P,1024) pre-BN adj_conv7;  g (B,1024) max over N;  the tiled global feature is never
                   materialised: seg/conv1 = cat @ W[1024:] + (g @ W[:1024])[cloud]   (DGCNN_S3DIS.py:87-96)
  ys1 (P,512), ys2 (P,256), Z (P,13)
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib as L
from . import ops
from . import runtime as rt
from .runtime import Layer, VariableStore

K_NEIGH = 20           # DGCNN_S3DIS.py:30
SMOOTH_KNN = 10        # SmoothConstraint.py:130 default
SMOOTH_GAMMA = 0.1
SIAMESE_W = 10.0       # S3DIS_DGCNN_trainer.py:128
KEEP = 0.7             # DGCNN_S3DIS.py:99

LAYERS = [("adj_conv1", 18, 64, True), ("adj_conv2", 64, 64, True), ("adj_conv3", 128, 64, True),
          ("adj_conv4", 64, 64, True), ("adj_conv5", 128, 64, True), ("adj_conv7", 192, 1024, True),
          ("seg/conv1", 1216, 512, True), ("seg/conv2", 512, 256, True), ("seg/conv3", 256, 13, False)]


class S3DISEngine:
    def __init__(self, params, B, N, device="cuda:0", k=K_NEIGH, num_classes=13, unnorm_xyz=False):
        self.B, self.N, self.k, self.C = B, N, k, num_classes
        self.P, self.R = B * N, B * N * k
        self.dev = torch.device(device)
        self.knn_coff = 0 if unnorm_xyz else 6      # get_model_unnormXYZ (:114) vs get_model (:32)
        self.vs = VariableStore(params, self.dev)
        self.layers = {s: Layer(self.vs, s, ci, co, bn) for s, ci, co, bn in LAYERS}
        f32 = dict(dtype=torch.float32, device=self.dev)
        i32 = dict(dtype=torch.int32, device=self.dev)
        P, R = self.P, self.R
        self.idx = [torch.empty((B, N, k), **i32) for _ in range(3)]
        self.idxS = torch.empty((B, N, SMOOTH_KNN), **i32)
        self.dS = torch.empty((B, N, SMOOTH_KNN), **f32)
        # fused EdgeConv blocks (csrc/edgeconv.cu) keep per-point state only; WSPC_EDGECONV=unfused materialises the
        # pre-BN outputs y1..y5 and the max-over-k gradient (round-1 formulation, kept as a second device path)
        self.fused = rt.EDGE_FUSED
        if self.fused:
            self.ef = rt.EdgeFused(P, self.dev)
            self.eb = [rt.EdgeBlockState(P, self.dev) for _ in range(3)]
            self.y, self.Ga, self.Gb = None, None, None
        else:
            self.y = [torch.empty((R, 64), **f32) for _ in range(5)]
            self.Ga, self.Gb = torch.empty((R, 64), **f32), torch.empty((R, 64), **f32)
        self.cat, self.dcat = torch.empty((P, 192), **f32), torch.empty((P, 192), **f32)
        # adj_conv7 + max over points in one pass (no (P,1024) tensor) when the Gram-identity backward is on and the shape fits
        self.pool_fused = rt.POOLCONV_GRAM and rt.pool_fusable(P, 1024, 192, N)
        self.cat_img = rt.RowImage(P, 192, self.dev) if rt.RowImage.usable(P, 192, (1024, 512)) else None
        if self.pool_fused:
            self.y7 = None
            self.pool_keys = torch.empty((B, 1024), dtype=torch.int64, device=self.dev)
            self.y7max = torch.empty((B, 1024), **f32)                 # pre-BN value at the arg-max row
            self.amax0 = torch.zeros((B, 1024), **i32)
        else:
            self.y7 = torch.empty((P, 1024), **f32)
        self.g, self.dg_in, self.dg = (torch.empty((B, 1024), **f32) for _ in range(3))
        self.amax = torch.empty((B, 1024), **i32)
        self.gW, self.S = torch.empty((B, 512), **f32), torch.empty((B, 512), **f32)
        self.ys1, self.Gs1 = torch.empty((P, 512), **f32), torch.empty((P, 512), **f32)
        self.ys2, self.Gs2 = torch.empty((P, 256), **f32), torch.empty((P, 256), **f32)
        self.dmask = torch.empty((P, 256), **f32)
        self.Z, self.Zp, self.dZ = (torch.empty((B, N, num_classes), **f32) for _ in range(3))
        self.losses = torch.zeros(5, **f32)
        self.zero_bias = torch.zeros(512, **f32)
        self.seed = 1234
        # first conv2d of each EdgeConv block: factored (csrc/edge.cu) unless WSPC_EDGE=gemm asks for the gathered GEMM
        self.es = rt.EdgeSplit(P, self.dev) if (rt.EDGE_FACTORED and not self.fused) else None
        self.MS = torch.empty((P, 128), **f32) if (rt.MAXK_SYNTH and not self.fused) else None   # [pooled max | dout / #ties]
        self.pc7 = rt.PoolConv(self.layers["adj_conv7"], self.dev) if rt.POOLCONV_GRAM else None
        self.prof = None   # optional list of (tag, start_event, end_event) filled around the kNN launches
        self.on_head_grads_ready = None   # data parallelism: called once the seg/* and adj_conv7 gradients are enqueued

    def _tick(self):
        if self.prof is None:
            return None
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def _tock(self, tag, t0):
        if t0 is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self.prof.append((tag, t0, e))

    # -------------------------------------------------------------------------------- forward ----
    def forward(self, X, is_training, bn_decay=None, dropout_mask=None, knn_override=None):
        """DGCNN_S3DIS.get_model. X (B,N,9) fp32 CUDA. Returns logits (B,N,13) (a view of engine memory)."""
        B, N, k, P, R = self.B, self.N, self.k, self.P, self.R
        Ly = self.layers
        assert X.shape == (B, N, 9) and X.is_cuda and X.is_contiguous()
        self.X = X
        ov = knn_override or {}
        cat_a = self.cat.data_ptr()

        def knn_into(i, src, ld, coff, D):
            if ov.get(f"knn{i + 1}") is not None:
                self.idx[i].copy_(ov[f"knn{i + 1}"])
                return
            ws = L.workspace(L.lib().wspc_knn_workspace_bytes(B, N, D), self.dev, "knn")
            t0 = self._tick()
            L.check(L.lib().wspc_knn_fused(ctypes.c_void_p(src), B, N, ld, coff, D, k, L.DIST_TFUTIL,
                                           L.ptr(self.idx[i]), None, L.ptr(ws), ws.numel(), L.stream()))
            self._tock(f"knn_D{D}_k{k}", t0)

        if self.fused:
            c = [Ly[f"adj_conv{i}"] for i in (1, 2, 3, 4, 5)]
            knn_into(0, X.data_ptr(), 9, self.knn_coff, 3)                                    # block 1 (:32-46)
            rt.edgeblock_forward(self.ef, self.eb[0], c[0], c[1], X, 9, 9, self.idx[0], k, N, P, is_training, bn_decay,
                                 cat_a, 192)
            knn_into(1, cat_a, 192, 0, 64)                                                    # block 2 (:48-62)
            rt.edgeblock_forward(self.ef, self.eb[1], c[2], c[3], cat_a, 192, 64, self.idx[1], k, N, P, is_training, bn_decay,
                                 cat_a + 4 * 64, 192)
            knn_into(2, cat_a, 192, 64, 64)                                                   # block 3 (:64-78)
            rt.edgeblock_forward(self.ef, self.eb[2], c[4], None, cat_a + 4 * 64, 192, 64, self.idx[2], k, N, P, is_training,
                                 bn_decay, cat_a + 4 * 128, 192)
            return self._forward_head(X, is_training, bn_decay, dropout_mask)
        # block 1: kNN on normalised xyz (ch 6:9), edge feature of all 9 channels      (:32-46)
        knn_into(0, X.data_ptr(), 9, self.knn_coff, 3)
        if self.es is not None:
            rt.edge_first_forward(self.es, Ly["adj_conv1"], X, 9, 9, self.idx[0], k, N, P, self.y[0], is_training, bn_decay)
        else:
            rt.conv_forward(Ly["adj_conv1"], rt.op_edge(X, 9, 9, self.idx[0], k, N), R, self.y[0], 64, is_training, bn_decay)
        rt.conv_forward(Ly["adj_conv2"], rt.op_bnrelu(self.y[0], Ly["adj_conv1"]), R, self.y[1], 64, is_training, bn_decay)
        rt.maxk_fwd(Ly["adj_conv2"], self.y[1], P, k, cat_a, 192)
        # block 2                                                                       (:48-62)
        knn_into(1, cat_a, 192, 0, 64)
        if self.es is not None:
            rt.edge_first_forward(self.es, Ly["adj_conv3"], cat_a, 192, 64, self.idx[1], k, N, P, self.y[2], is_training,
                                  bn_decay)
        else:
            rt.conv_forward(Ly["adj_conv3"], rt.op_edge(self.cat, 192, 64, self.idx[1], k, N), R, self.y[2], 64, is_training,
                            bn_decay)
        rt.conv_forward(Ly["adj_conv4"], rt.op_bnrelu(self.y[2], Ly["adj_conv3"]), R, self.y[3], 64, is_training, bn_decay)
        rt.maxk_fwd(Ly["adj_conv4"], self.y[3], P, k, cat_a + 4 * 64, 192)
        # block 3                                                                       (:64-78)
        knn_into(2, cat_a, 192, 64, 64)
        if self.es is not None:
            rt.edge_first_forward(self.es, Ly["adj_conv5"], cat_a + 4 * 64, 192, 64, self.idx[2], k, N, P, self.y[4],
                                  is_training, bn_decay, pool_out=cat_a + 4 * 128, pool_ld=192)
        else:
            e3 = L.Operand(p=cat_a + 4 * 64, ld=192, C=128, idx=L.dptr(self.idx[2]), k=k, npts=N), L.OP_EDGE
            rt.conv_forward(Ly["adj_conv5"], e3, R, self.y[4], 64, is_training, bn_decay)
            rt.maxk_fwd(Ly["adj_conv5"], self.y[4], P, k, cat_a + 4 * 128, 192)
        return self._forward_head(X, is_training, bn_decay, dropout_mask)

    def _forward_head(self, X, is_training, bn_decay, dropout_mask):
        B, N, P = self.B, self.N, self.P
        Ly = self.layers
        # adj_conv7 + max over points                                                   (:80-85)
        l7 = Ly["adj_conv7"]
        cat_op = rt.op_plain(self.cat, 192, 192)
        if self.cat_img is not None:     # adj_conv7 and seg/conv1 both read the concatenated features: split them once
            self.cat_img.build(self.cat, 192)
            cat_op = self.cat_img.operand()
        if self.pool_fused:     # the (P, 1024) pre-BN tensor is never written: BN sums + arg-max rows in the GEMM epilogue
            rt.conv_pool_forward(l7, cat_op, P, N, self.pool_keys, self.g, self.amax, self.y7max, is_training, bn_decay)
        else:
            rt.conv_forward(l7, cat_op, P, self.y7, 1024, is_training, bn_decay)
            L.check(L.lib().wspc_maxn_bnrelu_fwd(L.ptr(self.y7), L.ptr(l7.sc), L.ptr(l7.sh), B, N, 1024, L.ptr(self.g),
                                                 L.ptr(self.amax), L.stream()))
        # seg/conv1 with the tiled global feature folded into a per-cloud bias          (:87-96)
        s1, s2, s3 = Ly["seg/conv1"], Ly["seg/conv2"], Ly["seg/conv3"]
        epi = L.Epilogue(out=L.dptr(self.gW), ldo=512)
        rt.rows_gemm(rt.op_plain(self.g, 1024, 1024), s1.W, 512, 0, B, 512, 1024, epi, L.EPI_STORE)
        rt.conv_forward(s1, cat_op, P, self.ys1, 512, is_training, None, rowbias=self.gW, rb_rows=N, Wview=s1.W[1024:])
        rt.conv_forward(s2, rt.op_bnrelu(self.ys1, s1), P, self.ys2, 256, is_training, None)       # (:97-98)
        # dropout (:99) fused into seg/conv3's operand load (:100-101)
        self._mask = None
        if is_training:
            if dropout_mask is not None:
                self.dmask.copy_(dropout_mask.reshape(P, 256))
            else:
                L.check(L.lib().wspc_dropout_mask(L.ptr(self.dmask), P * 256, KEEP, self.seed, self.vs.step * (P * 64),
                                                  L.stream()))
            self._mask = self.dmask
        rt.conv_forward(s3, rt.op_bnrelu(self.ys2, s2, self._mask, KEEP if is_training else 1.0), P, self.Z, self.C,
                        is_training, None)
        return self.Z

    # ---------------------------------------------------------------------------------- losses ---
    def losses_and_grad(self, Y, Mask, full=True, want_grad=True, smooth_graph=None):
        """loss block of defineNetwork + WeakSupLoss; fills self.Zp, self.dZ, self.losses (device)."""
        B, N, C = self.B, self.N, self.C
        if full:
            if smooth_graph is not None:
                self.idxS.copy_(smooth_graph[0])
                self.dS.copy_(smooth_graph[1])
            else:   # SmoothConstraint.py:141-154 on X[:, :, 0:6]  (S3DIS_DGCNN_trainer.py:137)
                ws = L.workspace(L.lib().wspc_knn_workspace_bytes(B, N, 6), self.dev, "knn")
                t0 = self._tick()
                L.check(L.lib().wspc_knn_fused(L.ptr(self.X), B, N, 9, 0, 6, SMOOTH_KNN, L.DIST_SMOOTH, L.ptr(self.idxS),
                                               L.ptr(self.dS), L.ptr(ws), ws.numel(), L.stream()))
                self._tock(f"knn_D6_k{SMOOTH_KNN}", t0)
        nbytes = L.lib().wspc_head_losses_workspace_bytes(B, N, C)
        ws = L.workspace(nbytes, self.dev, "head")
        L.check(L.lib().wspc_head_losses(L.ptr(self.Z), L.ptr(Y), L.ptr(Mask), L.ptr(self.idxS) if full else None,
                                         L.ptr(self.dS) if full else None, B, N, C, SMOOTH_KNN, SMOOTH_GAMMA, SIAMESE_W,
                                         int(full), 1 if want_grad else 0, L.ptr(self.Zp), L.ptr(self.dZ),
                                         L.ptr(self.losses), L.ptr(ws), ws.numel(), L.stream()))
        return self.losses

    # -------------------------------------------------------------------------------- backward ---
    def backward(self):
        """TF autodiff of the graph above into the flat gradient buffer."""
        B, N, k, P, R = self.B, self.N, self.k, self.P, self.R
        Ly, dev = self.layers, self.dev
        c1, c2, c3, c4, c5, c7 = (Ly[f"adj_conv{i}"] for i in (1, 2, 3, 4, 5, 7))
        s1, s2, s3 = Ly["seg/conv1"], Ly["seg/conv2"], Ly["seg/conv3"]
        cat_a, dcat_a = self.cat.data_ptr(), self.dcat.data_ptr()
        C = self.C
        # seg/conv3 (no BN)
        G3 = rt.op_dy(self.dZ, C, None, 0, None, C)
        A3 = rt.op_bnrelu(self.ys2, s2, self._mask, KEEP)
        rt.wgrad(A3, G3, P, s3.dW, s3.db, dev)
        e, m = rt.epi_relumask(self.Gs2, s2, self.ys2, self._mask, KEEP)
        rt.rows_gemm(G3, s3.W, C, 1, P, 256, C, e, m)
        # seg/conv2
        rt.bn_bwd_coeffs(s2, P)
        G2 = rt.op_dy(self.Gs2, 256, self.ys2, 256, s2, 256)
        rt.wgrad(rt.op_bnrelu(self.ys1, s1), G2, P, s2.dW, s2.db, dev)
        e, m = rt.epi_relumask(self.Gs1, s1, self.ys1)
        rt.rows_gemm(G2, s2.W, 256, 1, P, 512, 256, e, m)
        # seg/conv1: point part, folded global part, and the gradient of the tiled global feature
        rt.bn_bwd_coeffs(s1, P)
        G1 = rt.op_dy(self.Gs1, 512, self.ys1, 512, s1, 512)
        rt.wgrad(rt.op_plain(self.cat, 192, 192), G1, P, s1.dW[1024:], s1.db, dev)
        L.check(L.lib().wspc_cloud_colsum(ctypes.byref(G1[0]), B, N, L.ptr(self.S), L.stream()))
        GS = rt.op_dy(self.S, 512, None, 0, None, 512)
        rt.wgrad(rt.op_plain(self.g, 1024, 1024), GS, B, s1.dW[:1024], None, dev)
        rt.rows_gemm(GS, s1.W, 512, 1, B, 1024, 512, L.Epilogue(out=L.dptr(self.dg_in), ldo=1024), L.EPI_STORE)
        # max over points: ReLU gate + BN-backward sums of the sparse gradient, then adj_conv7
        rt.zero_(c7.bstats)
        y7_rows, y7_n, y7_amax = (self.y7max, 1, self.amax0) if self.pool_fused else (self.y7, N, self.amax)
        L.check(L.lib().wspc_maxn_bwd_gate(L.ptr(self.g), L.ptr(self.dg_in), L.ptr(y7_amax), L.ptr(y7_rows), B, y7_n, 1024,
                                           L.ptr(self.dg), L.ptr(c7.bstats), L.stream()))
        rt.bn_bwd_coeffs(c7, P)
        if self.pc7 is not None:     # Gram identity (csrc/poolconv.cu): no (P,1024) operand, y7 is not read
            r0 = self.pc7.prepare()
            rt.rows_gemm(G1, s1.W[1024:], 512, 1, P, 192, 512, L.Epilogue(out=dcat_a, ldo=192, bias=L.dptr(r0)), L.EPI_STORE)
            self.pc7.backward(cat_a, 192, P, B, N, self.dg, self.amax, dcat_a, 192)
        else:
            rt.rows_gemm(G1, s1.W[1024:], 512, 1, P, 192, 512, L.Epilogue(out=dcat_a, ldo=192), L.EPI_STORE)
            G7 = rt.op_dy_sparse(self.y7, c7, self.dg, self.amax, N)
            rt.wgrad(rt.op_plain(self.cat, 192, 192), G7, P, c7.dW, c7.db, dev)
            rt.rows_gemm(G7, c7.W, 1024, 1, P, 192, 1024, L.Epilogue(out=dcat_a, ldo=192), L.EPI_ACCUM)
        if self.on_head_grads_ready is not None:
            self.on_head_grads_ready()
        if self.fused:
            ef, eb = self.ef, self.eb
            rt.edgeblock_backward(ef, eb[2], c5, None, cat_a + 4 * 64, 192, 64, self.idx[2], k, N, P, cat_a + 4 * 128, 192,
                                  dcat_a + 4 * 128, 192, dcat_a + 4 * 64, 192)
            rt.edgeblock_backward(ef, eb[1], c3, c4, cat_a, 192, 64, self.idx[1], k, N, P, cat_a + 4 * 64, 192,
                                  dcat_a + 4 * 64, 192, dcat_a, 192)
            rt.edgeblock_backward(ef, eb[0], c1, c2, self.X, 9, 9, self.idx[0], k, N, P, cat_a, 192, dcat_a, 192)
            return
        # block 3
        synth = self.MS is not None
        if synth and self.es is not None:
            rt.maxk_bwd_stats(c5, self.y[4], P, k, cat_a + 4 * 128, 192, dcat_a + 4 * 128, 192, self.MS)
            rt.bn_bwd_coeffs(c5, R)
            rt.edge_first_backward(self.es, c5, cat_a + 4 * 64, 192, 64, self.idx[2], k, N, P, None, self.y[4],
                                   dcat_a + 4 * 64, 192, MS=self.MS)
        else:
            rt.maxk_bwd(c5, self.y[4], P, k, cat_a + 4 * 128, 192, dcat_a + 4 * 128, 192, self.Ga)
            rt.bn_bwd_coeffs(c5, R)
            if self.es is not None:
                rt.edge_first_backward(self.es, c5, cat_a + 4 * 64, 192, 64, self.idx[2], k, N, P, self.Ga, self.y[4],
                                       dcat_a + 4 * 64, 192)
            else:
                G5 = rt.op_dy(self.Ga, 64, self.y[4], 64, c5, 64)
                A5 = L.Operand(p=cat_a + 4 * 64, ld=192, C=128, idx=L.dptr(self.idx[2]), k=k, npts=N), L.OP_EDGE
                rt.wgrad(A5, G5, R, c5.dW, c5.db, dev)
                e, m = rt.epi_scatter(dcat_a + 4 * 64, 192, self.idx[2], k, N)
                rt.rows_gemm(G5, c5.W, 64, 1, R, 128, 64, e, m)
        # block 2
        if synth:
            rt.maxk_bwd_stats(c4, self.y[3], P, k, cat_a + 4 * 64, 192, dcat_a + 4 * 64, 192, self.MS)
            rt.bn_bwd_coeffs(c4, R)
            G4 = rt.op_dy_maxk(c4, self.y[3], self.MS, k, N)
        else:
            rt.maxk_bwd(c4, self.y[3], P, k, cat_a + 4 * 64, 192, dcat_a + 4 * 64, 192, self.Ga)
            rt.bn_bwd_coeffs(c4, R)
            G4 = rt.op_dy(self.Ga, 64, self.y[3], 64, c4, 64)
        rt.conv_bwd(rt.op_bnrelu(self.y[2], c3), G4, R, c4, rt.epi_relumask(self.Gb, c3, self.y[2]), dev)
        rt.bn_bwd_coeffs(c3, R)
        if self.es is not None:
            rt.edge_first_backward(self.es, c3, cat_a, 192, 64, self.idx[1], k, N, P, self.Gb, self.y[2], dcat_a, 192)
        else:
            G3e = rt.op_dy(self.Gb, 64, self.y[2], 64, c3, 64)
            rt.wgrad(rt.op_edge(self.cat, 192, 64, self.idx[1], k, N), G3e, R, c3.dW, c3.db, dev)
            e, m = rt.epi_scatter(dcat_a, 192, self.idx[1], k, N)
            rt.rows_gemm(G3e, c3.W, 64, 1, R, 128, 64, e, m)
        # block 1 (no gradient w.r.t. the input cloud)
        if synth:
            rt.maxk_bwd_stats(c2, self.y[1], P, k, cat_a, 192, dcat_a, 192, self.MS)
            rt.bn_bwd_coeffs(c2, R)
            G2e = rt.op_dy_maxk(c2, self.y[1], self.MS, k, N)
        else:
            rt.maxk_bwd(c2, self.y[1], P, k, cat_a, 192, dcat_a, 192, self.Ga)
            rt.bn_bwd_coeffs(c2, R)
            G2e = rt.op_dy(self.Ga, 64, self.y[1], 64, c2, 64)
        rt.conv_bwd(rt.op_bnrelu(self.y[0], c1), G2e, R, c2, rt.epi_relumask(self.Gb, c1, self.y[0]), dev)
        rt.bn_bwd_coeffs(c1, R)
        if self.es is not None:
            rt.edge_first_backward(self.es, c1, self.X, 9, 9, self.idx[0], k, N, P, self.Gb, self.y[0])
        else:
            G1e = rt.op_dy(self.Gb, 64, self.y[0], 64, c1, 64)
            rt.wgrad(rt.op_edge(self.X, 9, 9, self.idx[0], k, N), G1e, R, c1.dW, c1.db, dev)

    # ------------------------------------------------------------------------------ train step ---
    def train_step(self, X, Y, Mask, lr, bn_decay, full=True, dropout_mask=None, knn_override=None, smooth_graph=None,
                   apply=True, gscale=1.0, allreduce=None):
        """forward + losses + backward (+ gradient all-reduce) + Adam; returns the device tensor of
        [seg, siamese, inexact, smooth, total] losses (no host sync)."""
        self.forward(X, True, bn_decay, dropout_mask, knn_override)
        self.losses_and_grad(Y, Mask, full, True, smooth_graph)
        self.backward()
        if allreduce is not None:
            allreduce(self.vs.grad)
        if apply:
            self.vs.adam_step(lr, gscale=gscale)
        return self.losses
