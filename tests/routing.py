"""Test helper: recovers from a (fused) S3DISEngine, after one training step, every discrete decision its forward pass took --
ReLU masks, the rows that attain each max over k, the arg-max rows of the max over points, the points that attain the
per-cloud logit maxima of the inexact loss -- in the form oracle/dgcnn.py's `forced_routing` consumes.

The fused engine never stores an edge tensor, so the per-edge decisions are re-derived from what it does keep:
  first conv of a block   a1 = relu(v_j*sc + (u_i*sc + (b*sc + sh))) from the stored UV = [u | v] and the folded BN affine
                          (the expression csrc/edgeconv.cu evaluates; plain torch rounds the products separately from the
                          fused multiply-adds, which can only matter for values within an ulp of zero)
  max over k              the row(s) whose pre-ReLU value is closest to the pooled output the engine wrote (equally close rows
                          = exact ties, e.g. duplicated neighbours, share the gradient as tf.reduce_max does); no row where
                          the pooled value is 0 (clipped by the ReLU)
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


def export_s3dis(eng):
    B, N, k, P = eng.B, eng.N, eng.k, eng.P
    Ly = eng.layers
    route = {}
    cat = eng.cat
    base = (torch.arange(B, device=cat.device) * N).view(B, 1, 1)
    for i, (s1, s2, col) in enumerate((("adj_conv1", "adj_conv2", 0), ("adj_conv3", "adj_conv4", 64), ("adj_conv5", None, 128))):
        l1 = Ly[s1]
        UV = eng.eb[i].UV
        u, v = UV[:, :64], UV[:, 64:]
        gi = (eng.idx[i].long() + base).reshape(P, k)
        t = l1.b * l1.sc + l1.sh
        pre1 = v[gi] * l1.sc + (u.unsqueeze(1) * l1.sc + t)                  # (P,k,64) fp32
        out = cat[:, col:col + 64].double()
        if s2 is None:
            y = ((u + l1.b).unsqueeze(1) + v[gi])                             # pre-BN y1 as edge_gather_stats forms it
            val = (y * l1.sc + l1.sh).double()
            route[f"maxk/knn{i + 1}"] = _maxk_weights(val, out).view(B, N, k, 64).cpu()
        else:
            l2 = Ly[s2]
            route[f"relu/{s1}"] = (pre1 > 0).view(B, N, k, 64).cpu()
            a1 = torch.relu(pre1).double()
            y2 = a1 @ l2.W.double() + l2.b.double()
            val = y2 * l2.sc.double() + l2.sh.double()
            route[f"maxk/knn{i + 1}"] = _maxk_weights(val, out).view(B, N, k, 64).cpu()
            del a1, y2
        del pre1, val
    route["maxn/adj_conv7"] = (eng.amax.long().cpu(), (eng.g > 0).cpu())
    s1, s2 = Ly["seg/conv1"], Ly["seg/conv2"]
    route["relu/seg/conv1"] = ((eng.ys1 * s1.sc + s1.sh) > 0).view(B, N, -1).cpu()
    route["relu/seg/conv2"] = ((eng.ys2 * s2.sc + s2.sh) > 0).view(B, N, -1).cpu()
    Z = eng.Z.double()
    sel = (Z == Z.max(dim=1, keepdim=True).values).double()
    route["inexact"] = (sel / sel.sum(dim=1, keepdim=True)).cpu()
    return route
