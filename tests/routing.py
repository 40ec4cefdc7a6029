"""Test helper: collects from a (fused) S3DISEngine, after one training step run with `runtime.ROUTING = {}`, every discrete
decision its pass took -- ReLU masks, the rows that attain each max over k, the arg-max rows of the max over points, the
points that attain the per-cloud logit maxima of the inexact loss -- in the form oracle/dgcnn.py's `forced_routing` consumes.

  first conv of a block   the fused engine never stores an edge tensor: a1 = relu(fma(v_j, sc, fma(u_i, sc, fma(b, sc, sh)))) is
                          re-evaluated from the stored UV = [u | v] and the BN affine with the kernels' own fused
                          multiply-adds (_fma32), so the masks are the kernels' masks bit for bit
  max over k              EXACT: the backward kernels export which rows they routed the gradient to (bit masks,
                          wspc_edgeconv2_bwd_ex / wspc_edge1_bwd_ex); tied rows share it equally as tf.reduce_max does; no row
                          where the pooled value is 0 (clipped by the ReLU)
"""
import torch


def _maxk_weights(pre, out):
    """pre (P,k,C) candidate PRE-ReLU values, out (P,C) pooled post-ReLU output of the engine -> (P,k,C) weights.
    Where out > 0 the row that produced it is the one whose value is closest to it (a value the engine saw as +1e-7 may be
    -1e-5 here: comparing before the ReLU still finds its row); where out == 0 the ReLU clipped the maximum: no row."""
    diff = (pre - out.unsqueeze(1)).abs()
    best = diff.min(dim=1, keepdim=True).values
    scale = out.abs().max().clamp_min(1e-30)
    sel = (diff <= best + 1e-7 * scale).to(torch.float64)
    w = sel / sel.sum(dim=1, keepdim=True)
    return w * (out > 0).to(torch.float64).unsqueeze(1)


def _weights_from_bits(sel, out):
    """sel (P,k,C) 0/1 rows attaining the maximum, out (P,C) pooled output -> (P,k,C) weights (equal split among ties)"""
    sel = sel.to(torch.float64)
    n = sel.sum(dim=1, keepdim=True)
    assert bool(((n > 0).squeeze(1) == (out > 0)).all()), "a positive pooled value must be attained by a row (and only then)"
    return sel / n.clamp_min(1.0)


def _fma32(a, b, c):
    """fmaf(a, b, c) of fp32 tensors as the kernels evaluate it (one rounding): the product of two fp32 values is exact in
    fp64, and the fp64 sum rounds to fp32 like the fused operation except for double-rounding ties (~1e-9 of the elements)"""
    return (a.double() * b.double() + c.double()).float()


def _export_blocks(eng, routing, route, idx_of_block):
    """the three fused EdgeConv blocks of the segmentation trunk (both engines); idx_of_block: engine.idx entries of knn1..3"""
    B, N, k, P = eng.B, eng.N, eng.k, eng.P
    Ly = eng.layers
    cat = eng.cat
    base = (torch.arange(B, device=cat.device) * N).view(B, 1, 1)
    bits32 = torch.arange(32, device=cat.device, dtype=torch.int32)
    for i, (s1, s2, col) in enumerate((("adj_conv1", "adj_conv2", 0), ("adj_conv3", "adj_conv4", 64), ("adj_conv5", None, 128))):
        l1 = Ly[s1]
        out = cat[:, col:col + 64]
        if s2 is None:
            words = routing[s1]                                               # (P, 64) int64, bit j = row j
            sel = ((words.unsqueeze(1) >> torch.arange(k, device=cat.device, dtype=torch.int64).view(1, k, 1)) & 1)
        else:
            UV = eng.eb[i].UV
            u, v = UV[:, :64], UV[:, 64:]
            gi = (idx_of_block[i].long() + base).reshape(P, k)
            t = _fma32(l1.b, l1.sc, l1.sh)
            pre1 = _fma32(v[gi], l1.sc, _fma32(u, l1.sc, t).unsqueeze(1))    # (P,k,64): csrc/edgeconv.cu's nested fmaf, bit for bit
            route[f"relu/{s1}"] = (pre1 > 0).view(B, N, k, 64).cpu()
            del pre1
            words = routing[s2].view(P, k, 2)                                 # bit c of word h = channel 32 h + c
            sel = ((words.unsqueeze(-1) >> bits32) & 1).reshape(P, k, 64)
        route[f"maxk/knn{i + 1}"] = _weights_from_bits(sel, out).view(B, N, k, 64).cpu()
    route["maxn/adj_conv7"] = (eng.amax.long().cpu(), (eng.g > 0).cpu())
    Z = eng.Z.double()
    sel = (Z == Z.max(dim=1, keepdim=True).values).double()
    route["inexact"] = (sel / sel.sum(dim=1, keepdim=True)).cpu()


def _relu_mask(y, layer):
    """the decision of the kernels' fmaf(y, sc, sh) > 0"""
    return _fma32(y, layer.sc, layer.sh) > 0


def export_shapenet(eng, routing):
    """ShapeNetEngine after one training step run with runtime.ROUTING = {}: the trunk's blocks as above, plus the T-net (its
    64 -> 128 EdgeConv block runs the materialised kernels: ReLU masks from the stored pre-BN tensors, the max-over-k split
    EXACTLY as the backward kernel made it: Gt2 / dtmax), the two max-over-points stages, the FC layers, the label branch and
    the four seg layers."""
    B, N, k, P = eng.B, eng.N, eng.k, eng.P
    Ly = eng.layers
    T = "transform_net1/"
    route = {}
    _export_blocks(eng, routing, route, [eng.idx[1], eng.idx[2], eng.idx[3]])
    t1, t2 = Ly[T + "tconv1"], Ly[T + "tconv2"]
    route["relu/" + T + "tconv1"] = _relu_mask(eng.yt1, t1).view(B, N, k, 64).cpu()
    G = eng.Gt2.view(P, k, 128).double()
    d = eng.dtmax.view(P, 1, 128).double()
    w = torch.where(d != 0, G / torch.where(d != 0, d, torch.ones_like(d)), torch.zeros_like(G))
    pre2 = _fma32(eng.yt2, t2.sc, t2.sh).view(P, k, 128)
    w_fallback = _maxk_weights(pre2.double(), eng.tmax.double())
    w = torch.where((d != 0).expand_as(w), w, w_fallback)
    # a positive pooled value is attained by at least one row (and only then)
    assert bool(((w.sum(1) > 0.5) == (eng.tmax > 0)).all())
    route["maxk/tnet"] = w.view(B, N, k, 128).cpu()
    route["maxn/" + T + "tconv3"] = (eng.tamax.long().cpu(), (eng.tg > 0).cpu())
    route["relu/" + T + "tfc1"] = _relu_mask(eng.yf1, Ly[T + "tfc1"]).cpu()
    route["relu/" + T + "tfc2"] = _relu_mask(eng.yf2, Ly[T + "tfc2"]).cpu()
    route["relu/one_hot_label_expand"] = _relu_mask(eng.ylab, Ly["one_hot_label_expand"]).cpu()
    for i, y in ((1, eng.ys1), (2, eng.ys2), (3, eng.ys3)):
        route[f"relu/seg/conv{i}"] = _relu_mask(y, Ly[f"seg/conv{i}"]).view(B, N, -1).cpu()
    return route


def export_s3dis(eng, routing):
    """routing: the dict the step filled through runtime.ROUTING"""
    B, N = eng.B, eng.N
    Ly = eng.layers
    route = {}
    _export_blocks(eng, routing, route, [eng.idx[0], eng.idx[1], eng.idx[2]])
    s1, s2 = Ly["seg/conv1"], Ly["seg/conv2"]
    route["relu/seg/conv1"] = _relu_mask(eng.ys1, s1).view(B, N, -1).cpu()
    route["relu/seg/conv2"] = _relu_mask(eng.ys2, s2).view(B, N, -1).cpu()
    return route
