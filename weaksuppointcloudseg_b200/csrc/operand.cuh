// Operand / epilogue descriptors shared by the GEMM kernels.
//
// The EdgeConv / per-point MLP layers of the reference are `conv2d 1x1 -> BN -> ReLU`
// chains (tf_util.py:115-173) whose inputs are either gathered edge features
// (tf_util.py:674-706) or the previous layer's activation.  Instead of
// materialising edge features / normalised activations in HBM, every GEMM reads
// its A operand through a loader that synthesises the element on the fly:
//   OP_PLAIN      a[row, c]
//   OP_BNRELU     relu(y[row,c]*sc[c] + sh[c]) (* dropout mask * 1/keep)
//   OP_EDGE       [x_i | x_j - x_i] for row = (point i, neighbour slot r)
//   OP_DY         c1[c]*G[row,c] + c2[c] + c3[c]*y[row,c]   (BN backward folded to an affine map)
//   OP_DY_SPARSE  same, with G given as the per-cloud arg-max scatter of max_pool2d's gradient
//   OP_DY_MAXK    same, with G synthesised from the max over k: G = (relu(bn(y)) == max) ? dout / #ties : 0
//   OP_IMG        a[row, c] given as the pre-split bf16 hi / lo tile image of wspc_rows_image (no loader work at all)
#pragma once
#include "common.cuh"

namespace wspc {

enum OpMode : int { OP_PLAIN = WSPC_OP_PLAIN, OP_BNRELU = WSPC_OP_BNRELU, OP_EDGE = WSPC_OP_EDGE, OP_DY = WSPC_OP_DY,
                    OP_DY_SPARSE = WSPC_OP_DY_SPARSE, OP_DY_MAXK = WSPC_OP_DY_MAXK, OP_IMG = WSPC_OP_IMG };

// gradient of tf.reduce_max over k at one element: a = relu(y*sc+sh) is the activation, m the pooled maximum, share = dout/#ties
__device__ __forceinline__ float maxk_grad(float y, float sc, float sh, float m, float share) {
  return (m > 0.f && fmaxf(fmaf(y, sc, sh), 0.f) == m) ? share : 0.f;
}

using Operand = wspc_operand_t;   // declared in include/wspc.h

// Loads channels [c0, c0+8) of logical row `row` (zero beyond C).  `row` must be valid.
template <int MODE>
__device__ __forceinline__ void load8(const Operand& o, long long row, int c0, float (&v)[8]) {
  if (MODE == OP_PLAIN) {
    const float* r = o.p + row * o.ld;
    if (c0 + 8 <= o.C && ((o.ld & 3) == 0) && ((c0 & 3) == 0)) {
      const float4 a = *reinterpret_cast<const float4*>(r + c0);
      const float4 b = *reinterpret_cast<const float4*>(r + c0 + 4);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = (c0 + i < o.C) ? r[c0 + i] : 0.f;
    }
  } else if (MODE == OP_BNRELU) {
    const float* r = o.p + row * o.ld;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = c0 + i;
      float t = 0.f;
      if (c < o.C) {
        t = fmaxf(fmaf(r[c], o.sc[c], o.sh[c]), 0.f);
        if (o.dmask) t *= o.dmask[row * o.C + c] * o.dscale;
      }
      v[i] = t;
    }
  } else if (MODE == OP_EDGE) {
    const int Cx = o.C >> 1;
    const long long pt = row / o.k;
    const long long base = (pt / o.npts) * o.npts;
    const long long nb = base + o.idx[row];
    const float* xi = o.p + pt * o.ld;
    const float* xj = o.p + nb * o.ld;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = c0 + i;
      float t = 0.f;
      if (c < Cx) t = xi[c];
      else if (c < o.C) t = xj[c - Cx] - xi[c - Cx];
      v[i] = t;
    }
  } else if (MODE == OP_DY) {
    const float* g = o.p + row * o.ld;
    const float* y = o.y + row * o.ldy;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = c0 + i;
      float t = 0.f;
      if (c < o.C) t = o.c1 ? fmaf(o.c1[c], g[c], fmaf(o.c3[c], y[c], o.c2[c])) : g[c];
      v[i] = t;
    }
  } else if (MODE == OP_DY_MAXK) {
    const long long pt = row / o.k;
    const float* y = o.y + row * o.ldy;
    const float* ms = o.p + pt * o.ld;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = c0 + i;
      float t = 0.f;
      if (c < o.C) {
        const float g = maxk_grad(y[c], o.sc[c], o.sh[c], ms[c], ms[o.C + c]);
        t = fmaf(o.c1[c], g, fmaf(o.c3[c], y[c], o.c2[c]));
      }
      v[i] = t;
    }
  } else {  // OP_DY_SPARSE
    const long long cloud = row / o.npts;
    const int n = (int)(row - cloud * o.npts);
    const float* y = o.y + row * o.ldy;
    const float* dg = o.dg + cloud * o.C;
    const int32_t* am = o.amax + cloud * o.C;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = c0 + i;
      float t = 0.f;
      if (c < o.C) {
        const float g = (am[c] == n) ? dg[c] : 0.f;
        t = fmaf(o.c1[c], g, fmaf(o.c3[c], y[c], o.c2[c]));
      }
      v[i] = t;
    }
  }
}

enum EpiMode : int {
  EPI_STORE = 0,           // out = acc + bias (+ rowbias)
  EPI_STORE_STATS = 1,     // + per-column (sum, sum of squares) in double  -> BN batch statistics
  EPI_RELUMASK_STATS = 2,  // out = acc * [prev activation > 0] (* dropout); stats = (sum G, sum G*y_prev)
  EPI_ACCUM = 3,           // out += acc
  EPI_EDGE_SCATTER = 4,    // acc = [dE_c | dE_d] -> atomics into point gradients (tf_util.py:700-705 backward)
  EPI_STATS_POOL = 5       // nothing stored: BN sums + per (cloud, column) extreme row of acc + bias (conv2d -> BN -> ReLU -> max over the
                           // cloud's points; internal to wspc_conv1x1_pool_fwd: scp = gamma, dx = packed keys, npts = rows per cloud)
};

using Epilogue = wspc_epilogue_t; // declared in include/wspc.h

}  // namespace wspc
