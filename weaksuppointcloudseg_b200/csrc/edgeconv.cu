// Fused EdgeConv blocks: nothing of shape (B*N*k, C) ever reaches HBM.
//
// Reference: one EdgeConv block of the DGCNN graphs is
//     get_edge_feature -> conv2d 1x1 + BN + ReLU [-> conv2d 1x1 + BN + ReLU] -> tf.reduce_max over k
// (tf_util.py:674-706, 115-173, 502-535; S3DIS/DGCNN_S3DIS.py:32-46, 48-62, 64-78; ShapeNet/DGCNN_ShapeNet.py:32-78).
// With the factored first layer y1_ij = u_i + v_j + b1 ([u | v] = X [W1 - W2 | W2], see edge.cu) every edge row is a
// function of two rows of the L2-resident (P, 128) matrix UV, so both passes RECOMPUTE the edge tensors on chip instead of
// storing them (round 1 wrote y1..y5 and the max-over-k gradient: ~78 GB of HBM traffic per cfg-3 step):
//
//   forward   edge_gather_stats    BN moments of y1 (+ extrema for a single-conv block, + the gradient-independent sums
//                                  S_i = sum_j v_j, SU_p = sum_{i->p} u_i, deg_p that the backward pass needs)
//             edgeconv2_fwd        tile of 128/k whole points: a1 = relu(bn1(u_i + v_j + b1)) -> bf16 hi/lo operand image in
//                                  shared memory -> tcgen05 GEMM with W2 (3 bf16 passes, fp32 accumulate in TMEM) -> per-point
//                                  max / min over k and BN-2 moments.  max_k relu(bn2(y2)) follows from the extrema
//                                  (wspc_maxk_from_extrema): y2 is never written.
//   backward  maxk_extrema_bwd_prep  BN-2 backward sums from per-point data (the max-over-k gradient lives on the extremal rows)
//             edgeconv2_bwd        same tile: recompute a1 and y2 (bit-identical MMA sequence), G = tie-split max-over-k
//                                  gradient, dy2 = c1 G + c2 + c3 y2 -> operand image; dW2 += a1^T dy2 (accumulated in TMEM
//                                  over the CTA's tiles), da1 = dy2 W2^T, ReLU mask -> per-point row sums SG and
//                                  neighbour scatter TG of the masked gradient (red.global.add.v4).
//             edge1_bwd            single-conv block: the same from one gather sweep (no GEMM).
//             edge_bwd_stats / edge_bwd_finalize   BN-1 backward is affine, dy1 = c1 g + c2 + c3 y1, and y1 is linear in u, v:
//                                  du_i = c1 SG_i + k (c2 + c3 (u_i + b1)) + c3 S_i
//                                  dv_p = c1 TG_p + deg_p (c2 + c3 (b1 + v_p)) + c3 SU_p
//                                  sum g      = sum_i SG_i,      sum g y1 = sum_i (u_i + b1) SG_i + v_i TG_i
//                                  so neither g nor y1 is needed per edge once SG / TG are known.
#include "operand.cuh"
#include "tc.cuh"
#include <math_constants.h>
#include <stdlib.h>

namespace wspc {
void count_launch(int n = 1);
namespace {

constexpr int CO = 64;                   // channels of every EdgeConv layer handled here
constexpr int C4 = CO / 4;
constexpr int EC_THREADS = 256;
constexpr int TILE_M = 128;
constexpr int AGB = TILE_M * 16 + 16;    // bytes of one 8-channel group of a 128-row operand image (padded)
constexpr int WGB = CO * 16 + 16;        // bytes of one 8-channel group of a 64-row weight image
constexpr int STAGE_LD = 68;             // staging row pitch (floats)
constexpr int STAGE_BYTES = TILE_M * STAGE_LD * 4;
constexpr int IMG_BYTES = 2 * 8 * AGB;   // hi + lo image of a (128 x 64) operand

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 f4add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4max(float4 a, float4 b) {
  return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
}
__device__ __forceinline__ float4 f4min(float4 a, float4 b) {
  return make_float4(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z), fminf(a.w, b.w));
}
__device__ __forceinline__ float4 f4fma(float4 a, float4 b, float4 c) {
  return make_float4(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z), fmaf(a.w, b.w, c.w));
}
__device__ __forceinline__ float4 f4shfl16(float4 v) {
  return make_float4(__shfl_xor_sync(0xffffffffu, v.x, 16), __shfl_xor_sync(0xffffffffu, v.y, 16),
                     __shfl_xor_sync(0xffffffffu, v.z, 16), __shfl_xor_sync(0xffffffffu, v.w, 16));
}

// block-level reduction of per-thread (float4 column quad) partial sums into two rows of fp64 global accumulators
__device__ __forceinline__ void flush_col_stats(double* __restrict__ stats, const double (&a)[4], const double (&b)[4], int c4,
                                                bool owner, double* red /* shared [2][CO] */) {
  if (threadIdx.x < 2 * CO) red[threadIdx.x] = 0.0;
  __syncthreads();
  if (owner) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      atomicAdd(&red[c4 * 4 + j], a[j]);
      atomicAdd(&red[CO + c4 * 4 + j], b[j]);
    }
  }
  __syncthreads();
  if (threadIdx.x < 2 * CO) atomicAdd(stats + threadIdx.x, red[threadIdx.x]);
}

// ------------------------------------------------------------------ gather statistics (forward) ---
// thread = (point, float4 column); y_ij = (u_i + b) + v_j.
__global__ void __launch_bounds__(256)
edge_gather_stats_kernel(const float* __restrict__ UV, long long ldu, const int32_t* __restrict__ idx,
                         const float* __restrict__ bias, long long P, int k, int npts, double* __restrict__ stats,
                         float* __restrict__ MM, float* __restrict__ SS, float* __restrict__ deg) {
  __shared__ double red[2 * CO];
  const int c4 = threadIdx.x & (C4 - 1), pl = threadIdx.x >> 4;
  const float4 b4 = bias ? ld4(bias + c4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  double a0[4] = {0.0, 0.0, 0.0, 0.0}, a1[4] = {0.0, 0.0, 0.0, 0.0};
  const float* vbase = UV + CO + c4 * 4;
  for (long long i = (long long)blockIdx.x * 16 + pl; i < P; i += (long long)gridDim.x * 16) {
    const long long cloud0 = (long long)((uint32_t)i / (uint32_t)npts) * npts;
    const float4 u = ld4(UV + i * ldu + c4 * 4);
    const float4 ub = f4add(u, b4);
    const int32_t* ip = idx + i * k;
    float4 S = make_float4(0.f, 0.f, 0.f, 0.f), s1 = S, s2 = S;
    float4 mx = make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
    float4 mn = make_float4(CUDART_INF_F, CUDART_INF_F, CUDART_INF_F, CUDART_INF_F);
    for (int j0 = 0; j0 < k; j0 += 4) {
      long long nb[4];
      float4 v[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) nb[t] = (j0 + t < k) ? cloud0 + ip[j0 + t] : -1;
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (nb[t] >= 0) v[t] = ld4(vbase + nb[t] * ldu);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        if (nb[t] < 0) continue;
        const float4 y = f4add(ub, v[t]);
        S = f4add(S, v[t]);
        s1 = f4add(s1, y);
        s2 = f4fma(y, y, s2);
        mx = f4max(mx, y);
        mn = f4min(mn, y);
        if (SS) {   // SU_p += u_i for every edge i -> p; deg_p += 1
          tc::red_add_v4(SS + nb[t] * (2 * CO) + CO + c4 * 4, u.x, u.y, u.z, u.w);
          if (c4 == 0) atomicAdd(deg + nb[t], 1.f);
        }
      }
    }
    if (SS) st4(SS + i * (2 * CO) + c4 * 4, S);
    if (MM) {
      st4(MM + i * (2 * CO) + c4 * 4, mx);
      st4(MM + i * (2 * CO) + CO + c4 * 4, mn);
    }
    a0[0] += (double)s1.x; a0[1] += (double)s1.y; a0[2] += (double)s1.z; a0[3] += (double)s1.w;
    a1[0] += (double)s2.x; a1[1] += (double)s2.y; a1[2] += (double)s2.z; a1[3] += (double)s2.w;
  }
  if (stats) flush_col_stats(stats, a0, a1, c4, true, red);
}

// ------------------------------------------------------- single-conv block: backward gather sweep ---
// G_ij = (relu(bn(y_ij)) == out_i && out_i > 0) ? dout_i / #ties : 0 ;  TS[i, 0:64] = SG_i = sum_j G_ij,
// TS[p, 64:128] += G_ij for every edge i -> p (sparse: only the extremal rows).  k <= 64.
__global__ void __launch_bounds__(256)
edge1_bwd_kernel(const float* __restrict__ UV, long long ldu, const int32_t* __restrict__ idx, const float* __restrict__ bias,
                 const float* __restrict__ sc, const float* __restrict__ sh, const float* __restrict__ out, long long ldo,
                 const float* __restrict__ dout, long long lddo, long long P, int k, int npts, float* __restrict__ TS,
                 unsigned long long* __restrict__ flags_out) {
  const int c4 = threadIdx.x & (C4 - 1), pl = threadIdx.x >> 4;
  const long long i = (long long)blockIdx.x * 16 + pl;
  if (i >= P) return;
  const float4 b4 = bias ? ld4(bias + c4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 s4 = ld4(sc + c4 * 4), h4 = ld4(sh + c4 * 4);
  const long long cloud0 = (long long)((uint32_t)i / (uint32_t)npts) * npts;
  const float4 ub = f4add(ld4(UV + i * ldu + c4 * 4), b4);
  float4 o = ld4(out + i * ldo + c4 * 4);
  const float4 d = ld4(dout + i * lddo + c4 * 4);
  // a >= 0 always, so -1 never matches: no gradient through a pooled value that the ReLU clipped
  o.x = o.x > 0.f ? o.x : -1.f; o.y = o.y > 0.f ? o.y : -1.f; o.z = o.z > 0.f ? o.z : -1.f; o.w = o.w > 0.f ? o.w : -1.f;
  const int32_t* ip = idx + i * k;
  const float* vbase = UV + CO + c4 * 4;
  unsigned long long e0 = 0, e1 = 0, e2 = 0, e3 = 0;   // bit j: row j attains the pooled maximum (per channel)
  for (int j0 = 0; j0 < k; j0 += 4) {
    long long nb[4];
    float4 v[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) nb[t] = (j0 + t < k) ? cloud0 + ip[j0 + t] : -1;
#pragma unroll
    for (int t = 0; t < 4; ++t)
      if (nb[t] >= 0) v[t] = ld4(vbase + nb[t] * ldu);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      if (nb[t] < 0) continue;
      const float4 y = f4add(ub, v[t]);
      const unsigned long long bit = 1ull << (j0 + t);
      if (fmaxf(fmaf(y.x, s4.x, h4.x), 0.f) == o.x) e0 |= bit;
      if (fmaxf(fmaf(y.y, s4.y, h4.y), 0.f) == o.y) e1 |= bit;
      if (fmaxf(fmaf(y.z, s4.z, h4.z), 0.f) == o.z) e2 |= bit;
      if (fmaxf(fmaf(y.w, s4.w, h4.w), 0.f) == o.w) e3 |= bit;
    }
  }
  if (flags_out) {   // routing export (tests): bit j of word (point, channel) = row j attains the pooled maximum
    flags_out[i * CO + c4 * 4 + 0] = e0; flags_out[i * CO + c4 * 4 + 1] = e1;
    flags_out[i * CO + c4 * 4 + 2] = e2; flags_out[i * CO + c4 * 4 + 3] = e3;
  }
  const int n0 = __popcll(e0), n1 = __popcll(e1), n2 = __popcll(e2), n3 = __popcll(e3);
  float4 share;   // tf.reduce_max splits the gradient equally among tied maxima [TF _MinOrMaxGrad]
  share.x = n0 ? d.x / (float)n0 : 0.f; share.y = n1 ? d.y / (float)n1 : 0.f;
  share.z = n2 ? d.z / (float)n2 : 0.f; share.w = n3 ? d.w / (float)n3 : 0.f;
  st4(TS + i * (2 * CO) + c4 * 4, make_float4(n0 ? d.x : 0.f, n1 ? d.y : 0.f, n2 ? d.z : 0.f, n3 ? d.w : 0.f));
  unsigned long long any = e0 | e1 | e2 | e3;
  while (any) {
    const int j = __ffsll((long long)any) - 1;
    any &= any - 1;
    const long long nb = cloud0 + ip[j];
    tc::red_add_v4(TS + nb * (2 * CO) + CO + c4 * 4, ((e0 >> j) & 1) ? share.x : 0.f, ((e1 >> j) & 1) ? share.y : 0.f,
                   ((e2 >> j) & 1) ? share.z : 0.f, ((e3 >> j) & 1) ? share.w : 0.f);
  }
}

// ------------------------------------------------ BN-1 backward sums and the (P, 128) gradient of [u | v] ---
__global__ void __launch_bounds__(256)
edge_bwd_stats_kernel(const float* __restrict__ TS, const float* __restrict__ UV, long long ldu, const float* __restrict__ bias,
                      long long P, double* __restrict__ bstats) {
  __shared__ double red[2 * CO];
  const int c4 = threadIdx.x & (C4 - 1), pl = threadIdx.x >> 4;
  const float4 b4 = bias ? ld4(bias + c4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
  double a0[4] = {0.0, 0.0, 0.0, 0.0}, a1[4] = {0.0, 0.0, 0.0, 0.0};
  int n = 0;
  for (long long i = (long long)blockIdx.x * 16 + pl; i < P; i += (long long)gridDim.x * 16) {
    const float4 sg = ld4(TS + i * (2 * CO) + c4 * 4), tg = ld4(TS + i * (2 * CO) + CO + c4 * 4);
    const float4 ub = f4add(ld4(UV + i * ldu + c4 * 4), b4), v = ld4(UV + i * ldu + CO + c4 * 4);
    s0 = f4add(s0, sg);
    s1 = f4fma(ub, sg, f4fma(v, tg, s1));
    if (++n == 16) {   // fp32 partials over 16 points, fp64 beyond
      a0[0] += (double)s0.x; a0[1] += (double)s0.y; a0[2] += (double)s0.z; a0[3] += (double)s0.w;
      a1[0] += (double)s1.x; a1[1] += (double)s1.y; a1[2] += (double)s1.z; a1[3] += (double)s1.w;
      s0 = make_float4(0.f, 0.f, 0.f, 0.f); s1 = s0; n = 0;
    }
  }
  a0[0] += (double)s0.x; a0[1] += (double)s0.y; a0[2] += (double)s0.z; a0[3] += (double)s0.w;
  a1[0] += (double)s1.x; a1[1] += (double)s1.y; a1[2] += (double)s1.z; a1[3] += (double)s1.w;
  flush_col_stats(bstats, a0, a1, c4, true, red);
}

__global__ void __launch_bounds__(256)
edge_bwd_finalize_kernel(const float* __restrict__ TS, const float* __restrict__ SS, const float* __restrict__ deg,
                         const float* __restrict__ UV, long long ldu, const float* __restrict__ bias,
                         const float* __restrict__ c1, const float* __restrict__ c2, const float* __restrict__ c3, long long P,
                         int k, float* __restrict__ DUV, long long ldd) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= P * C4) return;
  const long long i = t >> 4;
  const int c = (int)(t & 15) * 4;
  const float4 b4 = bias ? ld4(bias + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 k1 = ld4(c1 + c), k2 = ld4(c2 + c), k3 = ld4(c3 + c);
  const float4 sg = ld4(TS + i * (2 * CO) + c), tg = ld4(TS + i * (2 * CO) + CO + c);
  const float4 S = ld4(SS + i * (2 * CO) + c), SU = ld4(SS + i * (2 * CO) + CO + c);
  const float4 ub = f4add(ld4(UV + i * ldu + c), b4), vb = f4add(ld4(UV + i * ldu + CO + c), b4);
  const float dg = deg[i], kf = (float)k;
  float4 du, dv;
  du.x = fmaf(k1.x, sg.x, fmaf(kf, fmaf(k3.x, ub.x, k2.x), k3.x * S.x));
  du.y = fmaf(k1.y, sg.y, fmaf(kf, fmaf(k3.y, ub.y, k2.y), k3.y * S.y));
  du.z = fmaf(k1.z, sg.z, fmaf(kf, fmaf(k3.z, ub.z, k2.z), k3.z * S.z));
  du.w = fmaf(k1.w, sg.w, fmaf(kf, fmaf(k3.w, ub.w, k2.w), k3.w * S.w));
  dv.x = fmaf(k1.x, tg.x, fmaf(dg, fmaf(k3.x, vb.x, k2.x), k3.x * SU.x));
  dv.y = fmaf(k1.y, tg.y, fmaf(dg, fmaf(k3.y, vb.y, k2.y), k3.y * SU.y));
  dv.z = fmaf(k1.z, tg.z, fmaf(dg, fmaf(k3.z, vb.z, k2.z), k3.z * SU.z));
  dv.w = fmaf(k1.w, tg.w, fmaf(dg, fmaf(k3.w, vb.w, k2.w), k3.w * SU.w));
  st4(DUV + i * ldd + c, du);
  st4(DUV + i * ldd + CO + c, dv);
}

// ------------------------------------- BN-2 backward sums of the max-over-k gradient from per-point data ---
// MS[i, 0:64] = out > 0 ? out : -1 (the value a row must reproduce to receive gradient), MS[i, 64:128] = out > 0 ? dout : 0;
// stats += (sum_i MS[i,64+c], sum_i MS[i,64+c] * yext[i,c]) with yext the pre-BN extremum that produced out.
__global__ void __launch_bounds__(256)
maxk_extrema_bwd_prep_kernel(const float* __restrict__ MM, const float* __restrict__ sc, const float* __restrict__ out,
                             long long ldo, const float* __restrict__ dout, long long lddo, long long P,
                             float* __restrict__ MS, double* __restrict__ stats) {
  __shared__ double red[2 * CO];
  const int c4 = threadIdx.x & (C4 - 1), pl = threadIdx.x >> 4;
  const float4 s4 = ld4(sc + c4 * 4);
  double a0[4] = {0.0, 0.0, 0.0, 0.0}, a1[4] = {0.0, 0.0, 0.0, 0.0};
  float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
  int n = 0;
  for (long long i = (long long)blockIdx.x * 16 + pl; i < P; i += (long long)gridDim.x * 16) {
    const float4 mx = ld4(MM + i * (2 * CO) + c4 * 4), mn = ld4(MM + i * (2 * CO) + CO + c4 * 4);
    float4 o = ld4(out + i * ldo + c4 * 4), d = ld4(dout + i * lddo + c4 * 4);
    const float4 ye = make_float4(s4.x >= 0.f ? mx.x : mn.x, s4.y >= 0.f ? mx.y : mn.y, s4.z >= 0.f ? mx.z : mn.z,
                                  s4.w >= 0.f ? mx.w : mn.w);
    if (!(o.x > 0.f)) { o.x = -1.f; d.x = 0.f; }
    if (!(o.y > 0.f)) { o.y = -1.f; d.y = 0.f; }
    if (!(o.z > 0.f)) { o.z = -1.f; d.z = 0.f; }
    if (!(o.w > 0.f)) { o.w = -1.f; d.w = 0.f; }
    st4(MS + i * (2 * CO) + c4 * 4, o);
    st4(MS + i * (2 * CO) + CO + c4 * 4, d);
    s0 = f4add(s0, d);
    s1 = f4fma(d, ye, s1);
    if (++n == 16) {
      a0[0] += (double)s0.x; a0[1] += (double)s0.y; a0[2] += (double)s0.z; a0[3] += (double)s0.w;
      a1[0] += (double)s1.x; a1[1] += (double)s1.y; a1[2] += (double)s1.z; a1[3] += (double)s1.w;
      s0 = make_float4(0.f, 0.f, 0.f, 0.f); s1 = s0; n = 0;
    }
  }
  a0[0] += (double)s0.x; a0[1] += (double)s0.y; a0[2] += (double)s0.z; a0[3] += (double)s0.w;
  a1[0] += (double)s1.x; a1[1] += (double)s1.y; a1[2] += (double)s1.z; a1[3] += (double)s1.w;
  flush_col_stats(stats, a0, a1, c4, true, red);
}

// bias gradient of a conv followed by batch norm: db = sum_rows (c1 G + c2 + c3 y) = c1 sum G + rows c2 + c3 sum y
// (analytically zero; TF's BiasAddGrad produces the same rounding-level residue from the materialised tensor)
__global__ void bn_bias_grad_kernel(const double* __restrict__ fstats, const double* __restrict__ bstats,
                                    const float* __restrict__ c1, const float* __restrict__ c2, const float* __restrict__ c3,
                                    int C, double rows, float* __restrict__ db) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) db[c] = (float)((double)c1[c] * bstats[c] + rows * (double)c2[c] + (double)c3[c] * fstats[c]);
}

// =============================================================== tcgen05 tiles =====================
// weight image for a K-major B operand with N = 64 rows: element (n, kk) -> group kk/8, row n, slot kk%8.
// transposed == 0: B(n, kk) = W[kk][n]  (forward, y = a W);   transposed == 1: B(n, kk) = W[n][kk]  (data gradient, dy W^T)
__device__ __forceinline__ void load_weight_image(const float* __restrict__ W, int transposed, unsigned char* sHi,
                                                  unsigned char* sLo, int tid) {
  for (int e = tid; e < CO * 8; e += EC_THREADS) {
    const int n = e & (CO - 1), g = e >> 6;
    float w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = transposed ? W[n * CO + g * 8 + i] : W[(g * 8 + i) * CO + n];
    uint4 hi, lo;
    tc::split8(w, hi, lo);
    *reinterpret_cast<uint4*>(sHi + g * WGB + n * 16) = hi;
    *reinterpret_cast<uint4*>(sLo + g * WGB + n * 16) = lo;
  }
}

struct TileGeom {
  long long pt0;   // first point of the tile
  int npt, rows;   // whole points in the tile, rows = npt * k (<= 128)
};
__device__ __forceinline__ TileGeom tile_geom(int tile, int PT, int k, long long P) {
  TileGeom g;
  g.pt0 = (long long)tile * PT;
  const long long left = P - g.pt0;
  g.npt = left < PT ? (int)left : PT;
  g.rows = g.npt * k;
  return g;
}

// global point index of the neighbour of tile row `r` (thread r of the first 128 threads)
__device__ __forceinline__ int fetch_neighbour(const int32_t* __restrict__ idx, int tile, int num_tiles, int PT, int k,
                                               uint32_t kinv, int npts, long long P, int r) {
  if (tile >= num_tiles) return 0;
  const TileGeom g = tile_geom(tile, PT, k, P);
  if (r >= g.rows) return 0;
  const uint32_t i = (uint32_t)g.pt0 + __umulhi((uint32_t)r, kinv);
  const uint32_t cb = (i / (uint32_t)npts) * (uint32_t)npts;
  return (int)(cb + (uint32_t)idx[g.pt0 * k + r]);
}

// a1 tile: a1[r, c] = relu(bn1(u_i + v_j + b1)) evaluated as max(fma(v, sc, fma(u, sc, fma(b1, sc, sh))), 0); rows >= g.rows
// are zero.  thread = (channel group tid & 7, rows tid >> 3 + 32 s); written as the bf16 hi / lo K-major operand image.
__device__ __forceinline__ void build_a1_tile(const float* __restrict__ UV, long long ldu, const int* s_nb, const TileGeom& g,
                                              uint32_t kinv, const float (&scg)[8], const float (&tg)[8], unsigned char* sAhi,
                                              unsigned char* sAlo, int tid) {
  const int grp = tid & 7, rr = tid >> 3;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    float u[2][8], v[2][8];
    bool ok[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int r = rr + 32 * (2 * half + s);
      ok[s] = r < g.rows;
      if (ok[s]) {
        const long long i = g.pt0 + __umulhi((uint32_t)r, kinv);
        const long long nb = s_nb[r];
        tc::ld8(UV + i * ldu + grp * 8, u[s]);
        tc::ld8(UV + nb * ldu + CO + grp * 8, v[s]);
      }
    }
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int r = rr + 32 * (2 * half + s);
      uint4 hi = make_uint4(0, 0, 0, 0), lo = hi;
      if (ok[s]) {
        float a[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) a[e] = fmaxf(fmaf(v[s][e], scg[e], fmaf(u[s][e], scg[e], tg[e])), 0.f);
        tc::split8(a, hi, lo);
      }
      *reinterpret_cast<uint4*>(sAhi + grp * AGB + r * 16) = hi;
      *reinterpret_cast<uint4*>(sAlo + grp * AGB + r * 16) = lo;
    }
  }
}

// acc (128 rows x 64) = A (128 x 64, K-major image at a_hi / a_lo) * B (64 x 64 K-major weight image): hi*hi + lo*hi + hi*lo.
// The sequence is fixed so that the forward and the backward kernel obtain bit-identical y2.
__device__ __forceinline__ void issue_rows_gemm(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo,
                                                uint32_t idesc) {
  uint32_t accum = 0;
#pragma unroll 1
  for (int pass = 0; pass < 3; ++pass) {
    const uint32_t ab = (pass == 1) ? a_lo : a_hi;
    const uint32_t bb = (pass == 2) ? b_lo : b_hi;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const uint64_t ad = tc::smem_desc(ab + (uint32_t)(2 * kk) * AGB, AGB, 128);
      const uint64_t bd = tc::smem_desc(bb + (uint32_t)(2 * kk) * WGB, WGB, 128);
      tc::mma_bf16(d_tmem, ad, bd, idesc, accum);
      accum = 1;
    }
  }
}

// ------------------------------------------------------------------ forward: conv1 -> conv2 -> extrema ---
constexpr int FWD_SMEM = 2 * 8 * WGB + STAGE_BYTES + 512 + (EC_THREADS / 32) * 2 * CO * 8 + 64;

template <bool STATS>
__global__ void __launch_bounds__(EC_THREADS, 3)
edgeconv2_fwd_kernel(const float* __restrict__ UV, long long ldu, const int32_t* __restrict__ idx,
                     const float* __restrict__ b1, const float* __restrict__ sc1, const float* __restrict__ sh1,
                     const float* __restrict__ W2, const float* __restrict__ b2, long long P, int k, int npts, int PT,
                     int num_tiles, double* __restrict__ stats2, float* __restrict__ MM) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* sWhi = smem;
  unsigned char* sWlo = sWhi + 8 * WGB;
  unsigned char* sAhi = sWlo + 8 * WGB;
  unsigned char* sAlo = sAhi + 8 * AGB;
  float* stage = reinterpret_cast<float*>(sAhi);           // aliases the a1 image (dead once its MMAs have completed)
  int* s_nb = reinterpret_cast<int*>(sAhi + STAGE_BYTES);
  double* s_st = reinterpret_cast<double*>(sAhi + STAGE_BYTES + 512);     // [warp][2][CO] BN-2 moment accumulators
  uint64_t* mma_bar = reinterpret_cast<uint64_t*>(sAhi + STAGE_BYTES + 512 + (EC_THREADS / 32) * 2 * CO * 8);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_bar + 2);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (warp == 0) tc::alloc(tmem_slot, 64);
  if (tid == 32) { mbar_init(mma_bar, 1); mbar_fence_init(); }
  load_weight_image(W2, 0, sWhi, sWlo, tid);
  for (int i = tid; i < (EC_THREADS / 32) * 2 * CO; i += EC_THREADS) s_st[i] = 0.0;
  fence_proxy_async_smem();
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t idesc = tc::idesc_kmajor(CO);
  const uint32_t kinv = (uint32_t)((1ull << 32) / (uint32_t)k + 1);

  const int grp = tid & 7;
  float scg[8], tg[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    scg[e] = sc1[grp * 8 + e];
    tg[e] = fmaf(b1 ? b1[grp * 8 + e] : 0.f, scg[e], sh1[grp * 8 + e]);
  }
  const int lq = warp & 3, ch = warp >> 2;           // TMEM lane quadrant / 32-column half
  const int trow = lq * 32 + lane;
  const int c4 = lane & 15, hh = lane >> 4;          // reduction: float4 column, interleaved half of a point's rows
  const float4 bias2 = b2 ? ld4(b2 + c4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  double* my_st = s_st + warp * 2 * CO + c4 * 4;     // slot owned by lane (c4, hh == 0) of this warp

  uint32_t phase = 0;
  int nb_mine = (tid < TILE_M) ? fetch_neighbour(idx, blockIdx.x, num_tiles, PT, k, kinv, npts, P, tid) : 0;
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const TileGeom g = tile_geom(tile, PT, k, P);
    if (tid < TILE_M) s_nb[tid] = nb_mine;
    __syncthreads();
    if (tid < TILE_M) nb_mine = fetch_neighbour(idx, tile + gridDim.x, num_tiles, PT, k, kinv, npts, P, tid);
    build_a1_tile(UV, ldu, s_nb, g, kinv, scg, tg, sAhi, sAlo, tid);
    fence_proxy_async_smem();
    tc::fence_before();
    __syncthreads();
    if (tid == 0) {
      tc::fence_after();
      issue_rows_gemm(tmem_base, smem_u32(sAhi), smem_u32(sAlo), smem_u32(sWhi), smem_u32(sWlo), idesc);
      tc::commit(mma_bar);
    }
    mbar_wait(mma_bar, phase);
    phase ^= 1;
    tc::fence_after();
    {
      float v[32];
      tc::ld32(tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(ch * 32), v);
      float* srow = stage + trow * STAGE_LD + ch * 32;
#pragma unroll
      for (int i = 0; i < 32; i += 4) st4(srow + i, make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]));
    }
    __syncthreads();
    // per-point extrema over the k rows (+ BN-2 moments): warp <-> point, lanes = (float4 column, row parity)
    for (int p = warp; p < g.npt; p += EC_THREADS / 32) {
      float4 mx = make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
      float4 mn = make_float4(CUDART_INF_F, CUDART_INF_F, CUDART_INF_F, CUDART_INF_F);
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
      const float* base = stage + (p * k) * STAGE_LD + c4 * 4;
      for (int j = hh; j < k; j += 2) {
        const float4 y = f4add(ld4(base + j * STAGE_LD), bias2);
        mx = f4max(mx, y);
        mn = f4min(mn, y);
        if (STATS) { s = f4add(s, y); q = f4fma(y, y, q); }
      }
      mx = f4max(mx, f4shfl16(mx));
      mn = f4min(mn, f4shfl16(mn));
      if (STATS) { s = f4add(s, f4shfl16(s)); q = f4add(q, f4shfl16(q)); }
      if (hh == 0) {
        const long long i = g.pt0 + p;
        st4(MM + i * (2 * CO) + c4 * 4, mx);
        st4(MM + i * (2 * CO) + CO + c4 * 4, mn);
        if (STATS) {
          my_st[0] += (double)s.x; my_st[1] += (double)s.y; my_st[2] += (double)s.z; my_st[3] += (double)s.w;
          my_st[CO + 0] += (double)q.x; my_st[CO + 1] += (double)q.y; my_st[CO + 2] += (double)q.z; my_st[CO + 3] += (double)q.w;
        }
      }
    }
    tc::fence_before();
    __syncthreads();     // staging (= a1 image) and the accumulator are free for the next tile
    tc::fence_after();
  }
  if (STATS) {     // (the loop's trailing barrier ordered every my_st update)
    if (tid < 2 * CO) {
      double t = 0.0;
      for (int w = 0; w < EC_THREADS / 32; ++w) t += s_st[w * 2 * CO + tid];
      atomicAdd(stats2 + tid, t);
    }
  }
  __syncthreads();
  if (warp == 0) tc::dealloc(tmem_base, 64);
}

// ----------------------------------------------------------------- backward of the two-conv block ---
// shared memory: [W2 fwd image hi|lo][W2^T image hi|lo][a1 image hi|lo][dy2 image hi|lo (+pad; aliased by the staging tile)]
constexpr int BWD_OFF_WT = 2 * 8 * WGB;
constexpr int BWD_OFF_A = 2 * BWD_OFF_WT;
constexpr int BWD_OFF_G = BWD_OFF_A + IMG_BYTES;
constexpr int BWD_G_REGION = (STAGE_BYTES > IMG_BYTES) ? STAGE_BYTES : IMG_BYTES;
constexpr int BWD_OFF_CF = BWD_OFF_G + BWD_G_REGION;       // [6][64] floats: b2, sc2, sh2, c1, c2, c3
constexpr int BWD_OFF_MASK = BWD_OFF_CF + 6 * CO * 4;      // [128][2] words
constexpr int BWD_OFF_NB = BWD_OFF_MASK + TILE_M * 2 * 4;
constexpr int BWD_OFF_MISC = BWD_OFF_NB + TILE_M * 4;
constexpr int BWD_SMEM = BWD_OFF_MISC + 64;

__device__ __noinline__ float tie_count(const uint32_t* s_mask, int row0, int k, int ch, int c) {
  int n = 0;
  for (int j = 0; j < k; ++j) n += (s_mask[(row0 + j) * 2 + ch] >> c) & 1u;
  return (float)n;
}

__global__ void __launch_bounds__(EC_THREADS, 2)
edgeconv2_bwd_kernel(const float* __restrict__ UV, long long ldu, const int32_t* __restrict__ idx,
                     const float* __restrict__ b1, const float* __restrict__ sc1, const float* __restrict__ sh1,
                     const float* __restrict__ W2, const float* __restrict__ b2, const float* __restrict__ sc2,
                     const float* __restrict__ sh2, const float* __restrict__ c1, const float* __restrict__ c2,
                     const float* __restrict__ c3, const float* __restrict__ MS, long long P, int k, int npts, int PT,
                     int num_tiles, float* __restrict__ TS, float* __restrict__ partial) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* sWhi = smem;
  unsigned char* sWlo = sWhi + 8 * WGB;
  unsigned char* sWThi = smem + BWD_OFF_WT;
  unsigned char* sWTlo = sWThi + 8 * WGB;
  unsigned char* sAhi = smem + BWD_OFF_A;
  unsigned char* sAlo = sAhi + 8 * AGB;
  unsigned char* sGhi = smem + BWD_OFF_G;
  unsigned char* sGlo = sGhi + 8 * AGB;
  float* stage = reinterpret_cast<float*>(sGhi);           // aliases the dy2 image once its MMAs have completed
  float* s_cf = reinterpret_cast<float*>(smem + BWD_OFF_CF);
  uint32_t* s_mask = reinterpret_cast<uint32_t*>(smem + BWD_OFF_MASK);
  int* s_nb = reinterpret_cast<int*>(smem + BWD_OFF_NB);
  uint64_t* mma_bar = reinterpret_cast<uint64_t*>(smem + BWD_OFF_MISC);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + BWD_OFF_MISC + 16);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (warp == 0) tc::alloc(tmem_slot, 128);
  if (tid == 32) { mbar_init(mma_bar, 1); mbar_fence_init(); }
  load_weight_image(W2, 0, sWhi, sWlo, tid);
  load_weight_image(W2, 1, sWThi, sWTlo, tid);
  if (tid < CO) {
    s_cf[0 * CO + tid] = b2 ? b2[tid] : 0.f;
    s_cf[1 * CO + tid] = sc2[tid];
    s_cf[2 * CO + tid] = sh2[tid];
    s_cf[3 * CO + tid] = c1[tid];
    s_cf[4 * CO + tid] = c2[tid];
    s_cf[5 * CO + tid] = c3[tid];
  }
  fence_proxy_async_smem();
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t acc_y = tmem_base, acc_w = tmem_base + 64;      // y2 / da1 tile;  dW2 (persistent over the CTA's tiles)
  const uint32_t idesc = tc::idesc_kmajor(CO), idesc_mn = tc::idesc_mnmajor(CO);
  const uint32_t kinv = (uint32_t)((1ull << 32) / (uint32_t)k + 1);

  const int grp = tid & 7;
  float scg[8], tg[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    scg[e] = sc1[grp * 8 + e];
    tg[e] = fmaf(b1 ? b1[grp * 8 + e] : 0.f, scg[e], sh1[grp * 8 + e]);
  }
  const int lq = warp & 3, ch = warp >> 2;
  const int trow = lq * 32 + lane;
  const int c4 = lane & 15, hh = lane >> 4;
  const uint32_t lane_addr = (uint32_t)(lq * 32) << 16;

  uint32_t phase = 0, accum_w = 0;
  int nb_mine = (tid < TILE_M) ? fetch_neighbour(idx, blockIdx.x, num_tiles, PT, k, kinv, npts, P, tid) : 0;
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const TileGeom g = tile_geom(tile, PT, k, P);
    if (tid < TILE_M) s_nb[tid] = nb_mine;
    __syncthreads();
    if (tid < TILE_M) nb_mine = fetch_neighbour(idx, tile + gridDim.x, num_tiles, PT, k, kinv, npts, P, tid);
    build_a1_tile(UV, ldu, s_nb, g, kinv, scg, tg, sAhi, sAlo, tid);
    fence_proxy_async_smem();
    tc::fence_before();
    __syncthreads();
    if (tid == 0) {   // y2 = a1 W2 (same instruction sequence as the forward kernel)
      tc::fence_after();
      issue_rows_gemm(acc_y, smem_u32(sAhi), smem_u32(sAlo), smem_u32(sWhi), smem_u32(sWlo), idesc);
      tc::commit(mma_bar);
    }
    const bool valid = trow < g.rows;
    const int pl = (int)__umulhi((uint32_t)trow, kinv);           // tile-local point of this thread's row
    const float* msrow = MS + (g.pt0 + pl) * (2 * CO) + ch * 32;
    mbar_wait(mma_bar, phase);
    phase ^= 1;
    tc::fence_after();

    // ---- epilogue 1: y2 -> arg-max flags -> tie split -> dy2 = c1 G + c2 + c3 y2 -> operand image
    float y[32];
    tc::ld32(acc_y + lane_addr + (uint32_t)(ch * 32), y);
    uint32_t mine = 0;
    if (valid) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 o = ld4(msrow + 4 * q);
        const float4 bb = ld4(s_cf + 0 * CO + ch * 32 + 4 * q), ss = ld4(s_cf + 1 * CO + ch * 32 + 4 * q),
                     hs = ld4(s_cf + 2 * CO + ch * 32 + 4 * q);
        y[4 * q + 0] += bb.x; y[4 * q + 1] += bb.y; y[4 * q + 2] += bb.z; y[4 * q + 3] += bb.w;
        if (fmaxf(fmaf(y[4 * q + 0], ss.x, hs.x), 0.f) == o.x) mine |= 1u << (4 * q + 0);
        if (fmaxf(fmaf(y[4 * q + 1], ss.y, hs.y), 0.f) == o.y) mine |= 1u << (4 * q + 1);
        if (fmaxf(fmaf(y[4 * q + 2], ss.z, hs.z), 0.f) == o.z) mine |= 1u << (4 * q + 2);
        if (fmaxf(fmaf(y[4 * q + 3], ss.w, hs.w), 0.f) == o.w) mine |= 1u << (4 * q + 3);
      }
    }
    s_mask[trow * 2 + ch] = mine;
    __syncthreads();
    uint32_t dup = 0;
    if (valid && mine) {      // columns of this row whose maximum is shared with another row of the same point
      uint32_t seen = 0;
      const int row0 = pl * k;
      for (int j = 0; j < k; ++j) {
        const uint32_t m = s_mask[(row0 + j) * 2 + ch];
        dup |= seen & m;
        seen |= m;
      }
      dup &= mine;
    }
#pragma unroll
    for (int gg = 0; gg < 4; ++gg) {
      float dy[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) dy[e] = 0.f;
      if (valid) {
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          const int c0 = gg * 8 + h2 * 4;
          const float4 d = ld4(msrow + CO + c0);
          const float4 k1 = ld4(s_cf + 3 * CO + ch * 32 + c0), k2 = ld4(s_cf + 4 * CO + ch * 32 + c0),
                       k3 = ld4(s_cf + 5 * CO + ch * 32 + c0);
          const float dd[4] = {d.x, d.y, d.z, d.w}, kk1[4] = {k1.x, k1.y, k1.z, k1.w}, kk2[4] = {k2.x, k2.y, k2.z, k2.w},
                      kk3[4] = {k3.x, k3.y, k3.z, k3.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int c = c0 + e;
            float gv = ((mine >> c) & 1u) ? dd[e] : 0.f;
            if ((dup >> c) & 1u) gv = dd[e] / tie_count(s_mask, pl * k, k, ch, c);
            dy[h2 * 4 + e] = fmaf(kk1[e], gv, fmaf(kk3[e], y[c], kk2[e]));
          }
        }
      }
      uint4 hi, lo;
      tc::split8(dy, hi, lo);
      *reinterpret_cast<uint4*>(sGhi + (ch * 4 + gg) * AGB + trow * 16) = hi;
      *reinterpret_cast<uint4*>(sGlo + (ch * 4 + gg) * AGB + trow * 16) = lo;
    }
    fence_proxy_async_smem();
    tc::fence_before();
    __syncthreads();
    if (tid == 0) {
      tc::fence_after();
      const uint32_t ah = smem_u32(sAhi), gh = smem_u32(sGhi), gl = smem_u32(sGlo);
      // dW2 += a1^T dy2: both images read MN-major (reduction over the 128 rows).  M = 128 spans the 16 channel groups that
      // start at the hi image, i.e. [a1_hi ; a1_lo]: accumulator lanes 0..63 = a1_hi^T dy2, lanes 64..127 = a1_lo^T dy2.
#pragma unroll 1
      for (int pass = 0; pass < 2; ++pass) {
        const uint32_t gb = pass ? gl : gh;
#pragma unroll
        for (int j = 0; j < TILE_M / 16; ++j) {
          const uint64_t ad = tc::smem_desc(ah + (uint32_t)j * 256, 128, AGB);
          const uint64_t gd = tc::smem_desc(gb + (uint32_t)j * 256, 128, AGB);
          tc::mma_bf16(acc_w, ad, gd, idesc_mn, accum_w);
          accum_w = 1;
        }
      }
      // da1 = dy2 W2^T (the dy2 image read K-major) into the tile accumulator (y2 has been consumed)
      issue_rows_gemm(acc_y, gh, gl, smem_u32(sWThi), smem_u32(sWTlo), idesc);
      tc::commit(mma_bar);
    }
    mbar_wait(mma_bar, phase);
    phase ^= 1;
    tc::fence_after();

    // ---- epilogue 2: ReLU mask of a1, neighbour scatter (TG) and per-point row sums (SG)
    {
      float da[32];
      tc::ld32(acc_y + lane_addr + (uint32_t)(ch * 32), da);
#pragma unroll
      for (int gg = 0; gg < 4; ++gg) {
        const uint4 h = *reinterpret_cast<const uint4*>(sAhi + (ch * 4 + gg) * AGB + trow * 16);
        const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {      // a1 > 0 <=> its bf16 hi part is non-zero (a1 >= 0, round-to-nearest keeps the sign)
          if ((hw[e] & 0x0000ffffu) == 0u) da[gg * 8 + 2 * e] = 0.f;
          if ((hw[e] & 0xffff0000u) == 0u) da[gg * 8 + 2 * e + 1] = 0.f;
        }
      }
      if (valid) {
        float* dst = TS + (long long)s_nb[trow] * (2 * CO) + CO + ch * 32;
#pragma unroll
        for (int i = 0; i < 32; i += 4) tc::red_add_v4(dst + i, da[i], da[i + 1], da[i + 2], da[i + 3]);
      }
      float* srow = stage + trow * STAGE_LD + ch * 32;
#pragma unroll
      for (int i = 0; i < 32; i += 4) st4(srow + i, make_float4(da[i], da[i + 1], da[i + 2], da[i + 3]));
    }
    __syncthreads();
    for (int p = warp; p < g.npt; p += EC_THREADS / 32) {
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      const float* base = stage + (p * k) * STAGE_LD + c4 * 4;
      for (int j = hh; j < k; j += 2) s = f4add(s, ld4(base + j * STAGE_LD));
      s = f4add(s, f4shfl16(s));
      if (hh == 0) st4(TS + (g.pt0 + p) * (2 * CO) + c4 * 4, s);
    }
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
  }
  // dW2 partial of this CTA: slab 2*cta = a1_hi^T dy2, slab 2*cta + 1 = a1_lo^T dy2 (summed in fp64 by ec_slab_reduce)
  {
    float v[32];
    tc::ld32(acc_w + lane_addr + (uint32_t)(ch * 32), v);
    const int half = trow >> 6, cin = trow & 63;
    float* dst = partial + (((size_t)(2 * blockIdx.x + half) * CO + cin) * CO) + ch * 32;
#pragma unroll
    for (int i = 0; i < 32; i += 4) st4(dst + i, make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]));
  }
  tc::fence_before();
  __syncthreads();
  if (warp == 0) tc::dealloc(tmem_base, 128);
}

// =========================================================== warp-specialised backward (k >= 8) =======================
// The serial kernels above run build -> MMA -> epilogue one after the other inside a CTA and rely on 2-3 co-resident CTAs for
// overlap.  Here ONE persistent CTA per SM splits the backward tile by role:
//   warps 0-7    epilogue   TMEM -> registers -> (arg-max routing, BN backward, operand image | ReLU mask, scatter, row sums)
//   warps 8-15   builders   the TMA engine gathers the tile's raw rows: one 1-D bulk copy (cp.async.bulk) per neighbour row,
//                           issued by that row's thread, completing on an mbarrier; a1 = relu(bn1(u_i + v_j + b1)) is then
//                           formed from shared memory into the bf16 hi / lo operand image; one builder thread also issues
//                           the tcgen05.mma / tcgen05.commit (a 17th warp would cost the register file of 20: allocation
//                           is per four warps)
// Buffers: two a1 operand images, two dy2 images, two y2 and two da1 accumulators (tile parity), one raw staging tile; the
// forward weight image doubles as W2^T (read MN-major).  Roles talk through mbarriers (full / empty per buffer, parity = use
// count); each group also has a named barrier.  The a1 expression and the MMA sequence of y2 are those of the serial kernels:
// forward and backward obtain bit-identical y2.
//
// Measured (cfg-3, B200, profiles/r2_edgeconv_summary.md): the block is GATHER-bound, not tensor-bound.  One launch gathers
// 10.5 M rows of 256 B; with everything else removed the bulk-copy gather alone takes 0.71 ms per launch (~18 cycles per
// copy and SM), per-thread vector loads are bound by L1 wavefronts / miss latency at a similar level, and the tensor work
// of the tile (1280 cycles of tcgen05.mma for the backward, 384 for the forward) is 10-15 % of the tile time in every
// variant tried (serial 3.82 ms, per-thread gathers + 3 operand buffers 3.42 ms, this kernel 3.65 ms per backward launch;
// forward: serial 1.12-1.21 ms, warp-specialised 1.4-2.4 ms -> the forward stays on the serial kernel).
constexpr int WS_EPI_THREADS = 256, WS_BUILD_THREADS = 256;
constexpr int WS_THREADS = WS_EPI_THREADS + WS_BUILD_THREADS;
constexpr int WS_PTMAX = 16;                                            // points per tile: k >= 8
constexpr int WSB_NA = 2;                                                // a1 operand images in flight
constexpr int WS_VBYTES = TILE_M * 2 * CO * 2, WS_UBYTES = WS_PTMAX * 2 * CO * 2;   // raw v rows / u rows of one tile (256 B each)

__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void bar_sync_epi() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
// A K-major (rows x channels image), B MN-major: the forward weight image [cin group][cout row][8 cin] read with the
// reduction over its rows is W2^T -- one image serves y2 = a1 W2 and da1 = dy2 W2^T
__device__ __forceinline__ uint32_t idesc_a_k_b_mn(int N) { return tc::idesc_kmajor(N) | (1u << 16); }

// Gather by the TMA engine: builder thread r < 128 owns tile row r and issues ONE 1-D bulk copy (cp.async.bulk, SASS UBLKCP)
// of its neighbour's 256-byte v row into the raw staging tile; threads 128.. copy the u rows of the tile's points.  The copies
// complete on an mbarrier (expect_tx = total bytes); nothing is held in registers and the L1 miss queue is not involved
// (with per-thread loads the kernel was bound by gather latency / L1 wavefronts: ncu, 49 % long-scoreboard stalls).
// Afterwards builder thread = (16-byte piece c16 of a row, rows rr, rr + 16, ..): a1 = relu(bn1(u_i + v_j + b1)) from shared
// memory -> bf16 hi / lo operand image.
__device__ __forceinline__ void ws_issue_gather(const float* __restrict__ UV, long long ldu, const TileGeom& g, int k, uint32_t kinv,
                                                int nb, int bt, unsigned char* sV, unsigned char* sU, uint64_t* fullV) {
  if (bt < TILE_M) {
    if (bt < g.rows) bulk_g2s(sV + bt * (2 * CO * 2), UV + (long long)nb * ldu + CO, 2 * CO * 2, fullV);
  } else if (bt - TILE_M < g.npt) {
    bulk_g2s(sU + (bt - TILE_M) * (2 * CO * 2), UV + (g.pt0 + (bt - TILE_M)) * ldu, 2 * CO * 2, fullV);
  } else if (bt == WS_BUILD_THREADS - 1) {
    mbar_expect_tx(fullV, (uint32_t)(g.rows + g.npt) * (2 * CO * 2));
  }
}
__device__ __forceinline__ void ws_build_rows(const unsigned char* sV, const unsigned char* sU, int rows, uint32_t kinv, float4 sc4,
                                              float4 tg4, unsigned char* sAhi, unsigned char* sAlo, int rr, int c16) {
  const int off = (c16 >> 1) * AGB + (c16 & 1) * 8;
#pragma unroll
  for (int s = 0; s < TILE_M / 16; ++s) {
    const int r = rr + 16 * s;
    uint2 hi = make_uint2(0, 0), lo = hi;
    if (r < rows) {
      const float4 u = *reinterpret_cast<const float4*>(sU + __umulhi((uint32_t)r, kinv) * (2 * CO * 2) + c16 * 16);
      const float4 v = *reinterpret_cast<const float4*>(sV + r * (2 * CO * 2) + c16 * 16);
      const float a0 = fmaxf(fmaf(v.x, sc4.x, fmaf(u.x, sc4.x, tg4.x)), 0.f);
      const float a1 = fmaxf(fmaf(v.y, sc4.y, fmaf(u.y, sc4.y, tg4.y)), 0.f);
      const float a2 = fmaxf(fmaf(v.z, sc4.z, fmaf(u.z, sc4.z, tg4.z)), 0.f);
      const float a3 = fmaxf(fmaf(v.w, sc4.w, fmaf(u.w, sc4.w, tg4.w)), 0.f);
      const __nv_bfloat162 h01 = __floats2bfloat162_rn(a0, a1), h23 = __floats2bfloat162_rn(a2, a3);
      const uint32_t b01 = *reinterpret_cast<const uint32_t*>(&h01), b23 = *reinterpret_cast<const uint32_t*>(&h23);
      const __nv_bfloat162 l01 = __floats2bfloat162_rn(a0 - __uint_as_float(b01 << 16), a1 - __uint_as_float(b01 & 0xffff0000u));
      const __nv_bfloat162 l23 = __floats2bfloat162_rn(a2 - __uint_as_float(b23 << 16), a3 - __uint_as_float(b23 & 0xffff0000u));
      hi = make_uint2(b01, b23);
      lo = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
    }
    *reinterpret_cast<uint2*>(sAhi + off + r * 16) = hi;
    *reinterpret_cast<uint2*>(sAlo + off + r * 16) = lo;
  }
}
__device__ __forceinline__ void bar_sync_build() { asm volatile("bar.sync 2, 256;" ::: "memory"); }

// ------------------------------------------------------------------------------ backward ---
constexpr int WSB_OFF_A = 2 * 8 * WGB;
constexpr int WSB_OFF_G = WSB_OFF_A + WSB_NA * IMG_BYTES;
constexpr int WSB_OFF_CF = WSB_OFF_G + 2 * BWD_G_REGION;                // [8][64] floats: b2, sc2, sh2, c1, c2, c3, sc1, tg1
constexpr int WSB_OFF_MS = WSB_OFF_CF + 8 * CO * 4;                     // [WSB_NA][WS_PTMAX][128] floats
constexpr int WSB_OFF_CNT = WSB_OFF_MS + WSB_NA * WS_PTMAX * 2 * CO * 4; // [2][WS_PTMAX][64] ints
constexpr int WSB_OFF_NB = WSB_OFF_CNT + 2 * WS_PTMAX * CO * 4;         // [WSB_NA][128] ints
constexpr int WSB_OFF_V = WSB_OFF_NB + WSB_NA * TILE_M * 4;
constexpr int WSB_OFF_U = WSB_OFF_V + WS_VBYTES;
constexpr int WSB_OFF_BAR = WSB_OFF_U + WS_UBYTES;
constexpr int WSB_SMEM = WSB_OFF_BAR + 128;
static_assert(WSB_SMEM <= 227 * 1024, "edgeconv2_bwd_ws_kernel: shared memory budget");
static_assert(WSB_NA == 2, "the MMA issue order of edgeconv2_bwd_ws_kernel assumes two operand images");

// da1 (128 x 64) = dy2 (K-major image) * W2^T (the forward weight image read MN-major): hi*hi + lo*hi + hi*lo
__device__ __forceinline__ void issue_da_gemm(uint32_t d_tmem, uint32_t g_hi, uint32_t g_lo, uint32_t w_hi, uint32_t w_lo,
                                              uint32_t idesc_mix) {
  uint32_t accum = 0;
#pragma unroll 1
  for (int pass = 0; pass < 3; ++pass) {
    const uint32_t ab = (pass == 1) ? g_lo : g_hi;
    const uint32_t bb = (pass == 2) ? w_lo : w_hi;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {      // 16 output channels of layer 2 per step: two channel groups of dy2, 16 rows of the W image
      const uint64_t ad = tc::smem_desc(ab + (uint32_t)(2 * kk) * AGB, AGB, 128);
      const uint64_t bd = tc::smem_desc(bb + (uint32_t)kk * 256, 128, WGB);
      tc::mma_bf16(d_tmem, ad, bd, idesc_mix, accum);
      accum = 1;
    }
  }
}

__global__ void __launch_bounds__(WS_THREADS, 1)
edgeconv2_bwd_ws_kernel(const float* __restrict__ UV, long long ldu, const int32_t* __restrict__ idx,
                        const float* __restrict__ b1, const float* __restrict__ sc1, const float* __restrict__ sh1,
                        const float* __restrict__ W2, const float* __restrict__ b2, const float* __restrict__ sc2,
                        const float* __restrict__ sh2, const float* __restrict__ c1, const float* __restrict__ c2,
                        const float* __restrict__ c3, const float* __restrict__ MS, long long P, int k, int npts, int PT,
                        int num_tiles, float* __restrict__ TS, float* __restrict__ partial, uint32_t* __restrict__ flags_out) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* sWhi = smem;
  unsigned char* sWlo = sWhi + 8 * WGB;
  float* s_cf = reinterpret_cast<float*>(smem + WSB_OFF_CF);
  float* s_ms = reinterpret_cast<float*>(smem + WSB_OFF_MS);
  int* s_cnt = reinterpret_cast<int*>(smem + WSB_OFF_CNT);
  int* s_nb = reinterpret_cast<int*>(smem + WSB_OFF_NB);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WSB_OFF_BAR);
  uint64_t *fullA = bars, *emptyA = bars + WSB_NA, *fullG = bars + 2 * WSB_NA, *barY = fullG + 2, *barD = barY + 2, *fullV = barD + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (warp == 0) tc::alloc(tmem_slot, 512);
  if (tid == 32) {
    for (int i = 0; i < WSB_NA; ++i) {
      mbar_init(fullA + i, WS_BUILD_THREADS);
      mbar_init(emptyA + i, WS_EPI_THREADS);
    }
    mbar_init(fullV, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(fullG + i, WS_EPI_THREADS);
      mbar_init(barY + i, 1);
      mbar_init(barD + i, 1);
    }
    mbar_fence_init();
  }
  for (int e = tid; e < CO * 8; e += WS_THREADS) {
    const int n = e & (CO - 1), g = e >> 6;
    float w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = W2[(g * 8 + i) * CO + n];
    uint4 hi, lo;
    tc::split8(w, hi, lo);
    *reinterpret_cast<uint4*>(sWhi + g * WGB + n * 16) = hi;
    *reinterpret_cast<uint4*>(sWlo + g * WGB + n * 16) = lo;
  }
  if (tid < CO) {
    s_cf[0 * CO + tid] = b2 ? b2[tid] : 0.f;
    s_cf[1 * CO + tid] = sc2[tid];
    s_cf[2 * CO + tid] = sh2[tid];
    s_cf[3 * CO + tid] = c1[tid];
    s_cf[4 * CO + tid] = c2[tid];
    s_cf[5 * CO + tid] = c3[tid];
    s_cf[6 * CO + tid] = sc1[tid];
    s_cf[7 * CO + tid] = fmaf(b1 ? b1[tid] : 0.f, sc1[tid], sh1[tid]);
  }
  for (int i = tid; i < 2 * WS_PTMAX * CO; i += WS_THREADS) s_cnt[i] = 0;
  fence_proxy_async_smem();
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // tensor memory: y2 of tile parity 0/1 at columns 0/64, da1 at 128/192, dW2 (persistent over the CTA's tiles) at 256
  const uint32_t acc_w = tmem_base + 256;
  const uint32_t idesc = tc::idesc_kmajor(CO), idesc_mn = tc::idesc_mnmajor(CO), idesc_mix = idesc_a_k_b_mn(CO);
  const uint32_t kinv = (uint32_t)((1ull << 32) / (uint32_t)k + 1);
  const int n_my = (num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  // tcgen05.mma issue: ONE builder thread (the builders wait for the epilogue most of the time; on an epilogue thread the 28
  // MMAs + descriptors per tile delayed all sixteen warps by ~5 500 cycles per tile, measured).  Order in the tensor queue:
  // y2(t), then the gradients of tile t-1.
  uint32_t accum_w = 0;
  auto mma1 = [&](int t) {
    const int s = t % WSB_NA, sg = t & 1;
    mbar_wait(fullA + s, (uint32_t)(t / WSB_NA) & 1u);
    tc::fence_after();
    const uint32_t a_hi = smem_u32(smem + WSB_OFF_A + s * IMG_BYTES);
    issue_rows_gemm(tmem_base + (uint32_t)(sg * CO), a_hi, a_hi + 8 * AGB, smem_u32(sWhi), smem_u32(sWlo), idesc);
    tc::commit(barY + sg);
  };
  auto mma2 = [&](int t) {
    const int s = t % WSB_NA, sg = t & 1;
    mbar_wait(fullG + sg, (uint32_t)(t >> 1) & 1u);
    tc::fence_after();
    const uint32_t ah = smem_u32(smem + WSB_OFF_A + s * IMG_BYTES);
    const uint32_t gh = smem_u32(smem + WSB_OFF_G + sg * BWD_G_REGION), gl = gh + 8 * AGB;
    // dW2 += a1^T dy2: both images read MN-major (reduction over the 128 rows); M = 128 spans [a1_hi ; a1_lo]
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
      const uint32_t gb = pass ? gl : gh;
#pragma unroll
      for (int j = 0; j < TILE_M / 16; ++j) {
        const uint64_t ad = tc::smem_desc(ah + (uint32_t)j * 256, 128, AGB);
        const uint64_t gd = tc::smem_desc(gb + (uint32_t)j * 256, 128, AGB);
        tc::mma_bf16(acc_w, ad, gd, idesc_mn, accum_w);
        accum_w = 1;
      }
    }
    issue_da_gemm(tmem_base + (uint32_t)(2 * CO + sg * CO), gh, gl, smem_u32(sWhi), smem_u32(sWlo), idesc_mix);
    tc::commit(barD + sg);
  };

  if (warp < 8) {
    const int lq = warp & 3, ch = warp >> 2;
    const int trow = lq * 32 + lane;
    const int c4 = lane & 15, hh = lane >> 4;
    const uint32_t lane_addr = (uint32_t)(lq * 32) << 16;
    const int pl = (int)__umulhi((uint32_t)trow, kinv);           // tile-local point of this thread's row

    // ---- y2 -> arg-max flags -> tie split -> dy2 = c1 G + c2 + c3 y2 -> operand image G[t & 1]
    auto epi1 = [&](int t) {
      const int s = t % WSB_NA, sg = t & 1;
      const TileGeom g = tile_geom((int)blockIdx.x + t * (int)gridDim.x, PT, k, P);
      unsigned char* sGhi = smem + WSB_OFF_G + sg * BWD_G_REGION;
      unsigned char* sGlo = sGhi + 8 * AGB;
      mbar_wait(fullA + s, (uint32_t)(t / WSB_NA) & 1u);          // the builders' MS rows of this tile are visible
      mbar_wait(barY + sg, (uint32_t)(t >> 1) & 1u);
      tc::fence_after();
      float y[32];
      tc::ld32(tmem_base + (uint32_t)(sg * CO) + lane_addr + (uint32_t)(ch * 32), y);
      const bool valid = trow < g.rows;
      const float* ms = s_ms + (s * WS_PTMAX + (valid ? pl : 0)) * (2 * CO) + ch * 32;
      int* cnt = s_cnt + (sg * WS_PTMAX + (valid ? pl : 0)) * CO + ch * 32;
      uint32_t mine = 0;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 o = ld4(ms + 4 * q);
        const float4 bb = ld4(s_cf + 0 * CO + ch * 32 + 4 * q), ss = ld4(s_cf + 1 * CO + ch * 32 + 4 * q),
                     hs = ld4(s_cf + 2 * CO + ch * 32 + 4 * q);
        y[4 * q + 0] += bb.x; y[4 * q + 1] += bb.y; y[4 * q + 2] += bb.z; y[4 * q + 3] += bb.w;
        if (fmaxf(fmaf(y[4 * q + 0], ss.x, hs.x), 0.f) == o.x) mine |= 1u << (4 * q + 0);
        if (fmaxf(fmaf(y[4 * q + 1], ss.y, hs.y), 0.f) == o.y) mine |= 1u << (4 * q + 1);
        if (fmaxf(fmaf(y[4 * q + 2], ss.z, hs.z), 0.f) == o.z) mine |= 1u << (4 * q + 2);
        if (fmaxf(fmaf(y[4 * q + 3], ss.w, hs.w), 0.f) == o.w) mine |= 1u << (4 * q + 3);
      }
      if (!valid) mine = 0;
      if (flags_out && valid) flags_out[((size_t)g.pt0 * k + trow) * 2 + ch] = mine;
      for (uint32_t m = mine; m; m &= m - 1) atomicAdd(cnt + (__ffs((int)m) - 1), 1);   // rows attaining the maximum, per channel
      bar_sync_epi();
      // every epilogue thread has left the previous tile's epi1: its counters may be cleared for the tile after this one
      // (t == 0: the other buffer is still zero from the prologue, and tile 1 may already be counting into it)
      if (t > 0) reinterpret_cast<int4*>(s_cnt + (sg ^ 1) * WS_PTMAX * CO)[tid] = make_int4(0, 0, 0, 0);
      uint32_t dup = 0;
      for (uint32_t m = mine; m; m &= m - 1) {
        const int c = __ffs((int)m) - 1;
        if (cnt[c] > 1) dup |= 1u << c;
      }
#pragma unroll
      for (int gg = 0; gg < 4; ++gg) {
        float dy[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) dy[e] = 0.f;
        if (valid) {
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            const int c0 = gg * 8 + h2 * 4;
            const float4 d = ld4(ms + CO + c0);
            const float4 k1 = ld4(s_cf + 3 * CO + ch * 32 + c0), k2 = ld4(s_cf + 4 * CO + ch * 32 + c0),
                         k3 = ld4(s_cf + 5 * CO + ch * 32 + c0);
            const float dd[4] = {d.x, d.y, d.z, d.w}, kk1[4] = {k1.x, k1.y, k1.z, k1.w}, kk2[4] = {k2.x, k2.y, k2.z, k2.w},
                        kk3[4] = {k3.x, k3.y, k3.z, k3.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int c = c0 + e;
              const float gv = ((mine >> c) & 1u) ? dd[e] : 0.f;
              dy[h2 * 4 + e] = fmaf(kk1[e], gv, fmaf(kk3[e], y[c], kk2[e]));
            }
          }
          if ((dup >> (gg * 8)) & 0xffu) {          // tf.reduce_max: equal split among tied maxima (rare: duplicated neighbours)
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int c = gg * 8 + e;
              if ((dup >> c) & 1u)
                dy[e] = fmaf(s_cf[3 * CO + ch * 32 + c], ms[CO + c] / (float)cnt[c],
                             fmaf(s_cf[5 * CO + ch * 32 + c], y[c], s_cf[4 * CO + ch * 32 + c]));
            }
          }
        }
        uint4 hi, lo;
        tc::split8(dy, hi, lo);
        *reinterpret_cast<uint4*>(sGhi + (ch * 4 + gg) * AGB + trow * 16) = hi;
        *reinterpret_cast<uint4*>(sGlo + (ch * 4 + gg) * AGB + trow * 16) = lo;
      }
      fence_proxy_async_smem();
      tc::fence_before();
      mbar_arrive(fullG + sg);
    };

    // ---- da1 -> ReLU mask of a1 -> neighbour scatter (TG) and per-point row sums (SG)
    auto epi2 = [&](int t) {
      const int s = t % WSB_NA, sg = t & 1;
      const TileGeom g = tile_geom((int)blockIdx.x + t * (int)gridDim.x, PT, k, P);
      const unsigned char* sAhi = smem + WSB_OFF_A + s * IMG_BYTES;
      float* stage = reinterpret_cast<float*>(smem + WSB_OFF_G + sg * BWD_G_REGION);   // the dy2 image is dead: its MMAs are done
      mbar_wait(barD + sg, (uint32_t)(t >> 1) & 1u);
      tc::fence_after();
      float da[32];
      tc::ld32(tmem_base + (uint32_t)(2 * CO + sg * CO) + lane_addr + (uint32_t)(ch * 32), da);
#pragma unroll
      for (int gg = 0; gg < 4; ++gg) {
        const uint4 h = *reinterpret_cast<const uint4*>(sAhi + (ch * 4 + gg) * AGB + trow * 16);
        const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {      // a1 > 0 <=> its bf16 hi part is non-zero
          if ((hw[e] & 0x0000ffffu) == 0u) da[gg * 8 + 2 * e] = 0.f;
          if ((hw[e] & 0xffff0000u) == 0u) da[gg * 8 + 2 * e + 1] = 0.f;
        }
      }
      float* srow = stage + trow * STAGE_LD + ch * 32;
#pragma unroll
      for (int i = 0; i < 32; i += 4) st4(srow + i, make_float4(da[i], da[i + 1], da[i + 2], da[i + 3]));
      bar_sync_epi();
      // neighbour scatter from the staged tile: sixteen lanes add one contiguous 256-byte row (lane = row would spread every
      // reduction instruction over 32 cache lines)
#pragma unroll
      for (int pass = 0; pass < TILE_M / 16; ++pass) {
        const int row = pass * 16 + (tid >> 4);
        if (row < g.rows) {
          const float4 v4 = ld4(stage + row * STAGE_LD + c4 * 4);
          tc::red_add_v4(TS + (long long)s_nb[s * TILE_M + row] * (2 * CO) + CO + c4 * 4, v4.x, v4.y, v4.z, v4.w);
        }
      }
      tc::fence_before();
      mbar_arrive(emptyA + s);                        // the operand image, its neighbour list and MS rows may be overwritten
      for (int p = warp; p < g.npt; p += 8) {
        float4 sm = make_float4(0.f, 0.f, 0.f, 0.f);
        const float* base = stage + (p * k) * STAGE_LD + c4 * 4;
        for (int j = hh; j < k; j += 2) sm = f4add(sm, ld4(base + j * STAGE_LD));
        sm = f4add(sm, f4shfl16(sm));
        if (hh == 0) st4(TS + (g.pt0 + p) * (2 * CO) + c4 * 4, sm);
      }
    };

    if (n_my > 0) epi1(0);
    for (int t = 0; t < n_my; ++t) {
      if (t + 1 < n_my) epi1(t + 1);                  // runs while the tensor pipe works on tile t's gradients
      epi2(t);
    }
    // dW2 partial of this CTA: slab 2*cta = a1_hi^T dy2, slab 2*cta + 1 = a1_lo^T dy2 (summed in fp64 by ec_slab_reduce)
    {
      float v[32];
      tc::ld32(acc_w + lane_addr + (uint32_t)(ch * 32), v);
      const int half = trow >> 6, cin = trow & 63;
      float* dst = partial + (((size_t)(2 * blockIdx.x + half) * CO + cin) * CO) + ch * 32;
#pragma unroll
      for (int i = 0; i < 32; i += 4) st4(dst + i, make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]));
    }
  } else {
    const int bt = tid - WS_EPI_THREADS, c16 = bt & 15, rr = bt >> 4;
    const float4 sc4 = ld4(s_cf + 6 * CO + c16 * 4), tg4 = ld4(s_cf + 7 * CO + c16 * 4);
    unsigned char* sV = smem + WSB_OFF_V;
    unsigned char* sU = smem + WSB_OFF_U;
    int tile = blockIdx.x;
    TileGeom g = tile_geom(tile, PT, k, P);
    int nb = (bt < TILE_M) ? fetch_neighbour(idx, tile, num_tiles, PT, k, kinv, npts, P, bt) : 0;
    if (n_my > 0) ws_issue_gather(UV, ldu, g, k, kinv, nb, bt, sV, sU, fullV);
    int nb_next = (bt < TILE_M) ? fetch_neighbour(idx, tile + gridDim.x, num_tiles, PT, k, kinv, npts, P, bt) : 0;
    for (int t = 0; t < n_my; ++t) {
      const int s = t % WSB_NA, use = t / WSB_NA;
      if (use >= 1) mbar_wait(emptyA + s, (uint32_t)(use - 1) & 1u);
      if (bt < TILE_M) s_nb[s * TILE_M + bt] = nb;
      for (int e = bt; e < g.npt * 32; e += WS_BUILD_THREADS)     // MS rows of the tile's points: npt * 32 16-byte pieces
        cp_async16(s_ms + (size_t)s * WS_PTMAX * 2 * CO + e * 4, MS + g.pt0 * (2 * CO) + e * 4);
      mbar_wait(fullV, (uint32_t)t & 1u);            // the tile's raw rows have landed
      unsigned char* sAhi = smem + WSB_OFF_A + s * IMG_BYTES;
      ws_build_rows(sV, sU, g.rows, kinv, sc4, tg4, sAhi, sAhi + 8 * AGB, rr, c16);
      cp_async_wait_all();
      fence_proxy_async_smem();
      mbar_arrive(fullA + s);
      bar_sync_build();                              // every builder has consumed the staging tile: refill it
      tile += gridDim.x;
      nb = nb_next;
      if (t + 1 < n_my) {
        g = tile_geom(tile, PT, k, P);
        ws_issue_gather(UV, ldu, g, k, kinv, nb, bt, sV, sU, fullV);
      }
      nb_next = (bt < TILE_M) ? fetch_neighbour(idx, tile + gridDim.x, num_tiles, PT, k, kinv, npts, P, bt) : 0;
      if (bt == WS_BUILD_THREADS - 1) {
        mma1(t);
        if (t > 0) mma2(t - 1);
      }
    }
    if (bt == WS_BUILD_THREADS - 1 && n_my > 0) mma2(n_my - 1);
  }
  tc::fence_before();
  __syncthreads();
  if (warp == 0) tc::dealloc(tmem_base, 512);
}

// p[row, col0 : col0 + 4*n4] = 0 for a (rows, ld) matrix (the scatter halves of SS / TS)
__global__ void zero_cols_kernel(float* __restrict__ p, long long ld, int col0, int n4, long long rows) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rows * n4) return;
  const long long r = t / n4;
  const int c = (int)(t - r * n4) * 4;
  st4(p + r * ld + col0 + c, make_float4(0.f, 0.f, 0.f, 0.f));
}

__global__ void ec_slab_reduce_kernel(const float* __restrict__ partial, int S, float* __restrict__ dW) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= CO * CO) return;
  double s = 0.0;
  for (int z = 0; z < S; ++z) s += (double)partial[(size_t)z * CO * CO + i];
  dW[i] = (float)s;
}

int ec_tiles(long long P, int k, int* PT_out) {
  const int PT = TILE_M / k;
  *PT_out = PT;
  return (int)((P + PT - 1) / PT);
}

// k >= 8 (<= 16 points per 128-row tile): persistent warp-specialised kernels; smaller k or WSPC_EDGECONV_KERNEL=serial: the
// serial-phase kernels (kept as the second device path for A/B runs)
bool ec_use_ws(int k) {
  static const bool serial = [] {
    const char* e = getenv("WSPC_EDGECONV_KERNEL");
    return e && strcmp(e, "serial") == 0;
  }();
  return !serial && k >= 8;
}

bool ec_shape_ok(long long P, int k, int npts, const char* who) {
  if (P < 1 || k < 1 || k > TILE_M || npts < 1 || P % npts != 0 || P * (long long)k >= (1ll << 31)) {
    set_error("%s: bad shape P=%lld k=%d npts=%d (1 <= k <= 128, P %% npts == 0, P*k < 2^31)", who, P, k, npts);
    return false;
  }
  return true;
}

}  // namespace
}  // namespace wspc

using namespace wspc;

extern "C" int wspc_edge_gather_stats(const float* UV, long long ldu, const int32_t* idx, const float* bias, long long P, int k,
                                      int npts, int Cout, double* stats, float* MM, float* SS, float* deg,
                                      wspc_stream_t stream) {
  if (int rc = check_arch()) return rc;
  WSPC_REQUIRE(UV && idx, "edge_gather_stats: null pointer");
  WSPC_REQUIRE(Cout == CO && ldu >= 2 * CO && (ldu & 3) == 0, "edge_gather_stats: Cout=%d ldu=%lld (Cout must be %d)", Cout, ldu, CO);
  if (!ec_shape_ok(P, k, npts, "edge_gather_stats")) return WSPC_ERR_INVALID;
  WSPC_REQUIRE((SS == nullptr) == (deg == nullptr), "edge_gather_stats: SS and deg go together");
  WSPC_REQUIRE(aligned16(UV) && (!bias || aligned16(bias)) && (!MM || aligned16(MM)) && (!SS || aligned16(SS)),
               "edge_gather_stats: pointers must be 16-byte aligned");
  const long long chunks = (P + 15) / 16;
  const unsigned grid = (unsigned)(chunks < 8LL * kNumSM ? chunks : 8LL * kNumSM);
  edge_gather_stats_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(UV, ldu, idx, bias, P, k, npts, stats, MM, SS,
                                                                                      deg);
  count_launch();
  WSPC_LAUNCH_CHECK("edge_gather_stats_kernel");
  return WSPC_OK;
}

extern "C" int wspc_edge1_bwd(const float* UV, long long ldu, const int32_t* idx, const float* bias, const float* sc,
                              const float* sh, const float* out, long long ldo, const float* dout, long long lddo, long long P,
                              int k, int npts, int Cout, float* TS, wspc_stream_t stream) {
  return wspc_edge1_bwd_ex(UV, ldu, idx, bias, sc, sh, out, ldo, dout, lddo, P, k, npts, Cout, TS, nullptr, stream);
}

extern "C" int wspc_edge1_bwd_ex(const float* UV, long long ldu, const int32_t* idx, const float* bias, const float* sc,
                                 const float* sh, const float* out, long long ldo, const float* dout, long long lddo,
                                 long long P, int k, int npts, int Cout, float* TS, uint64_t* routing_out,
                                 wspc_stream_t stream) {
  if (int rc = check_arch()) return rc;
  WSPC_REQUIRE(UV && idx && sc && sh && out && dout && TS, "edge1_bwd: null pointer");
  WSPC_REQUIRE(Cout == CO && ldu >= 2 * CO && (ldu & 3) == 0 && (ldo & 3) == 0 && (lddo & 3) == 0,
               "edge1_bwd: Cout=%d ldu=%lld ldo=%lld lddo=%lld", Cout, ldu, ldo, lddo);
  if (!ec_shape_ok(P, k, npts, "edge1_bwd")) return WSPC_ERR_INVALID;
  WSPC_REQUIRE(k <= 64, "edge1_bwd: k=%d > 64", k);
  WSPC_REQUIRE(aligned16(UV) && (!bias || aligned16(bias)) && aligned16(sc) && aligned16(sh) && aligned16(out) &&
               aligned16(dout) && aligned16(TS), "edge1_bwd: pointers must be 16-byte aligned");
  edge1_bwd_kernel<<<(unsigned)((P + 15) / 16), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      UV, ldu, idx, bias, sc, sh, out, ldo, dout, lddo, P, k, npts, TS, reinterpret_cast<unsigned long long*>(routing_out));
  count_launch();
  WSPC_LAUNCH_CHECK("edge1_bwd_kernel");
  return WSPC_OK;
}

extern "C" int wspc_edge_bwd_stats(const float* TS, const float* UV, long long ldu, const float* bias, long long P, int Cout,
                                   double* bstats, wspc_stream_t stream) {
  if (int rc = check_arch()) return rc;
  WSPC_REQUIRE(TS && UV && bstats && P >= 1, "edge_bwd_stats: null pointer / empty");
  WSPC_REQUIRE(Cout == CO && ldu >= 2 * CO && (ldu & 3) == 0, "edge_bwd_stats: Cout=%d ldu=%lld", Cout, ldu);
  WSPC_REQUIRE(aligned16(TS) && aligned16(UV) && (!bias || aligned16(bias)), "edge_bwd_stats: pointers must be 16-byte aligned");
  const long long chunks = (P + 15) / 16;
  const unsigned grid = (unsigned)(chunks < 8LL * kNumSM ? chunks : 8LL * kNumSM);
  edge_bwd_stats_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(TS, UV, ldu, bias, P, bstats);
  count_launch();
  WSPC_LAUNCH_CHECK("edge_bwd_stats_kernel");
  return WSPC_OK;
}

extern "C" int wspc_edge_bwd_finalize(const float* TS, const float* SS, const float* deg, const float* UV, long long ldu,
                                      const float* bias, const float* c1, const float* c2, const float* c3, long long P, int k,
                                      int Cout, float* DUV, long long ldd, wspc_stream_t stream) {
  if (int rc = check_arch()) return rc;
  WSPC_REQUIRE(TS && SS && deg && UV && c1 && c2 && c3 && DUV && P >= 1 && k >= 1, "edge_bwd_finalize: null pointer / empty");
  WSPC_REQUIRE(Cout == CO && ldu >= 2 * CO && (ldu & 3) == 0 && ldd >= 2 * CO && (ldd & 3) == 0,
               "edge_bwd_finalize: Cout=%d ldu=%lld ldd=%lld", Cout, ldu, ldd);
  WSPC_REQUIRE(aligned16(TS) && aligned16(SS) && aligned16(UV) && aligned16(DUV) && (!bias || aligned16(bias)) && aligned16(c1) &&
               aligned16(c2) && aligned16(c3), "edge_bwd_finalize: pointers must be 16-byte aligned");
  edge_bwd_finalize_kernel<<<(unsigned)((P * C4 + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      TS, SS, deg, UV, ldu, bias, c1, c2, c3, P, k, DUV, ldd);
  count_launch();
  WSPC_LAUNCH_CHECK("edge_bwd_finalize_kernel");
  return WSPC_OK;
}

extern "C" int wspc_maxk_extrema_bwd_prep(const float* MM, const float* sc, const float* out, long long ldo, const float* dout,
                                          long long lddo, long long P, int C, float* MS, double* stats, wspc_stream_t stream) {
  if (int rc = check_arch()) return rc;
  WSPC_REQUIRE(MM && sc && out && dout && MS && stats && P >= 1, "maxk_extrema_bwd_prep: null pointer / empty");
  WSPC_REQUIRE(C == CO && (ldo & 3) == 0 && (lddo & 3) == 0, "maxk_extrema_bwd_prep: C=%d ldo=%lld lddo=%lld", C, ldo, lddo);
  WSPC_REQUIRE(aligned16(MM) && aligned16(sc) && aligned16(out) && aligned16(dout) && aligned16(MS),
               "maxk_extrema_bwd_prep: pointers must be 16-byte aligned");
  const long long chunks = (P + 15) / 16;
  const unsigned grid = (unsigned)(chunks < 8LL * kNumSM ? chunks : 8LL * kNumSM);
  maxk_extrema_bwd_prep_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(MM, sc, out, ldo, dout, lddo, P, MS, stats);
  count_launch();
  WSPC_LAUNCH_CHECK("maxk_extrema_bwd_prep_kernel");
  return WSPC_OK;
}

extern "C" int wspc_bn_bias_grad(const double* fstats, const double* bstats, const float* c1, const float* c2, const float* c3,
                                 int C, double rows, float* db, wspc_stream_t stream) {
  if (int rc = check_arch()) return rc;
  WSPC_REQUIRE(fstats && bstats && c1 && c2 && c3 && db && C >= 1, "bn_bias_grad: null pointer");
  bn_bias_grad_kernel<<<(C + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(fstats, bstats, c1, c2, c3, C, rows, db);
  count_launch();
  WSPC_LAUNCH_CHECK("bn_bias_grad_kernel");
  return WSPC_OK;
}

extern "C" int wspc_edgeconv2_fwd(const float* UV, long long ldu, const int32_t* idx, const float* bias1, const float* sc1,
                                  const float* sh1, const float* W2, const float* bias2, long long P, int k, int npts, int C1,
                                  int C2, double* stats2, float* MM, wspc_stream_t stream) {
  if (int rc = check_arch()) return rc;
  WSPC_REQUIRE(UV && idx && sc1 && sh1 && W2 && MM, "edgeconv2_fwd: null pointer");
  WSPC_REQUIRE(C1 == CO && C2 == CO && ldu >= 2 * CO && (ldu & 3) == 0, "edgeconv2_fwd: C1=%d C2=%d ldu=%lld (channels must be %d)",
               C1, C2, ldu, CO);
  if (!ec_shape_ok(P, k, npts, "edgeconv2_fwd")) return WSPC_ERR_INVALID;
  WSPC_REQUIRE(k >= 2, "edgeconv2_fwd: k=%d < 2", k);
  WSPC_REQUIRE(aligned16(UV) && aligned16(MM) && (!bias2 || aligned16(bias2)), "edgeconv2_fwd: pointers must be 16-byte aligned");
  int PT;
  const int tiles = ec_tiles(P, k, &PT);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int grid = tiles < 3 * kNumSM ? tiles : 3 * kNumSM;
  if (stats2) {
    auto kern = edgeconv2_fwd_kernel<true>;
    WSPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM));
    kern<<<grid, EC_THREADS, FWD_SMEM, st>>>(UV, ldu, idx, bias1, sc1, sh1, W2, bias2, P, k, npts, PT, tiles, stats2, MM);
  } else {
    auto kern = edgeconv2_fwd_kernel<false>;
    WSPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM));
    kern<<<grid, EC_THREADS, FWD_SMEM, st>>>(UV, ldu, idx, bias1, sc1, sh1, W2, bias2, P, k, npts, PT, tiles, stats2, MM);
  }
  count_launch();
  WSPC_LAUNCH_CHECK("edgeconv2_fwd_kernel");
  return WSPC_OK;
}

extern "C" size_t wspc_edgeconv2_bwd_workspace_bytes(void) { return (size_t)2 * (2 * kNumSM) * CO * CO * sizeof(float); }

extern "C" int wspc_edgeconv2_bwd(const float* UV, long long ldu, const int32_t* idx, const float* bias1, const float* sc1,
                                  const float* sh1, const float* W2, const float* bias2, const float* sc2, const float* sh2,
                                  const float* c1, const float* c2, const float* c3, const float* MS, long long P, int k, int npts,
                                  int C1, int C2, float* TS, float* dW2, void* workspace, size_t workspace_bytes,
                                  wspc_stream_t stream) {
  return wspc_edgeconv2_bwd_ex(UV, ldu, idx, bias1, sc1, sh1, W2, bias2, sc2, sh2, c1, c2, c3, MS, P, k, npts, C1, C2, TS, dW2,
                               nullptr, workspace, workspace_bytes, stream);
}

extern "C" int wspc_edgeconv2_bwd_ex(const float* UV, long long ldu, const int32_t* idx, const float* bias1, const float* sc1,
                                     const float* sh1, const float* W2, const float* bias2, const float* sc2, const float* sh2,
                                     const float* c1, const float* c2, const float* c3, const float* MS, long long P, int k,
                                     int npts, int C1, int C2, float* TS, float* dW2, uint32_t* routing_out, void* workspace,
                                     size_t workspace_bytes, wspc_stream_t stream) {
  if (int rc = check_arch()) return rc;
  WSPC_REQUIRE(UV && idx && sc1 && sh1 && W2 && sc2 && sh2 && c1 && c2 && c3 && MS && TS && dW2 && workspace,
               "edgeconv2_bwd: null pointer");
  WSPC_REQUIRE(C1 == CO && C2 == CO && ldu >= 2 * CO && (ldu & 3) == 0, "edgeconv2_bwd: C1=%d C2=%d ldu=%lld (channels must be %d)",
               C1, C2, ldu, CO);
  if (!ec_shape_ok(P, k, npts, "edgeconv2_bwd")) return WSPC_ERR_INVALID;
  WSPC_REQUIRE(k >= 2, "edgeconv2_bwd: k=%d < 2", k);
  WSPC_REQUIRE(aligned16(UV) && aligned16(MS) && aligned16(TS) && aligned16(workspace),
               "edgeconv2_bwd: pointers must be 16-byte aligned");
  if (workspace_bytes < wspc_edgeconv2_bwd_workspace_bytes()) {
    set_error("edgeconv2_bwd: workspace %zu < required %zu", workspace_bytes, wspc_edgeconv2_bwd_workspace_bytes());
    return WSPC_ERR_WORKSPACE;
  }
  int PT;
  const int tiles = ec_tiles(P, k, &PT);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* partial = static_cast<float*>(workspace);
  if (ec_use_ws(k)) {
    const int wgrid = tiles < kNumSM ? tiles : kNumSM;
    auto kern = edgeconv2_bwd_ws_kernel;
    WSPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, WSB_SMEM));
    kern<<<wgrid, WS_THREADS, WSB_SMEM, st>>>(UV, ldu, idx, bias1, sc1, sh1, W2, bias2, sc2, sh2, c1, c2, c3, MS, P, k, npts, PT,
                                              tiles, TS, partial, routing_out);
    count_launch();
    WSPC_LAUNCH_CHECK("edgeconv2_bwd_ws_kernel");
    ec_slab_reduce_kernel<<<(CO * CO + 255) / 256, 256, 0, st>>>(partial, 2 * wgrid, dW2);
    count_launch();
    WSPC_LAUNCH_CHECK("ec_slab_reduce_kernel");
    return WSPC_OK;
  }
  WSPC_REQUIRE(routing_out == nullptr, "edgeconv2_bwd: the routing export needs the warp-specialised kernel (k >= 8)");
  const int grid = tiles < 2 * kNumSM ? tiles : 2 * kNumSM;
  auto kern = edgeconv2_bwd_kernel;
  WSPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM));
  kern<<<grid, EC_THREADS, BWD_SMEM, st>>>(UV, ldu, idx, bias1, sc1, sh1, W2, bias2, sc2, sh2, c1, c2, c3, MS, P, k, npts, PT, tiles,
                                           TS, partial);
  count_launch();
  WSPC_LAUNCH_CHECK("edgeconv2_bwd_kernel");
  ec_slab_reduce_kernel<<<(CO * CO + 255) / 256, 256, 0, st>>>(partial, 2 * grid, dW2);
  count_launch();
  WSPC_LAUNCH_CHECK("ec_slab_reduce_kernel");
  return WSPC_OK;
}

extern "C" int wspc_zero_cols(float* p, long long ld, int col0, int ncols, long long rows, wspc_stream_t stream) {
  if (int rc = check_arch()) return rc;
  WSPC_REQUIRE(p && rows >= 1 && ncols >= 4 && (ncols & 3) == 0 && (col0 & 3) == 0 && (ld & 3) == 0 && col0 + ncols <= ld &&
               aligned16(p), "zero_cols: bad arguments (16-byte aligned, multiples of 4)");
  const long long total = rows * (ncols / 4);
  zero_cols_kernel<<<(unsigned)((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p, ld, col0, ncols / 4, rows);
  count_launch();
  WSPC_LAUNCH_CHECK("zero_cols_kernel");
  return WSPC_OK;
}
