// Factored first layer of an EdgeConv block.
//
// The reference feeds the edge feature e_ij = [x_i | x_j - x_i] (tf_util.get_edge_feature, tf_util.py:674-706) to a
// 1x1 conv2d with W = [W1; W2] (2*Cx, 64) (DGCNN_S3DIS.py:36-39, :52-55, :68-71).  Algebraically
//     y_ij = x_i W1 + (x_j - x_i) W2 + b = x_i (W1 - W2) + x_j W2 + b = u_i + v_j + b,
// so the R = P*k-row GEMM collapses into one P-row GEMM [u | v] = X [W1 - W2 | W2] (20x fewer flops, no gathered
// operand) followed by the streaming gather-add below; the backward pass is its transpose:
//     du_i = sum_j dy_ij,   dv_p = sum_{(i,j): idx_ij = p} dy_ij,   dX = [du | dv] [W1 - W2 | W2]^T,
//     d[W1 - W2 | W2] = X^T [du | dv]   =>   dW1 = dWa,  dW2 = dWb - dWa.
// Both kernels are HBM-bound on the (R, 64) tensors (y written once forward; G and y read once backward).
#include "operand.cuh"

namespace wspc {
void count_launch(int n = 1);
namespace {

constexpr int CO = 64;                 // output channels of every EdgeConv layer in both models
constexpr int C4 = CO / 4;             // float4 columns per row

__device__ __forceinline__ void red_add_v4(float* addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Wc (Cx, 2*CO) = [W1 - W2 | W2]
__global__ void edge_split_weights_kernel(const float* __restrict__ W, int Cx, float* __restrict__ Wc) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= Cx * CO) return;
  const int c = t / CO, o = t - c * CO;
  const float w1 = W[(size_t)c * CO + o], w2 = W[(size_t)(Cx + c) * CO + o];
  Wc[(size_t)c * 2 * CO + o] = w1 - w2;
  Wc[(size_t)c * 2 * CO + CO + o] = w2;
}

// dW (2*Cx, CO) from dWc (Cx, 2*CO);  db = first half of the column sums
__global__ void edge_merge_wgrad_kernel(const float* __restrict__ dWc, const float* __restrict__ dbc, int Cx,
                                        float* __restrict__ dW, float* __restrict__ db) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < Cx * CO) {
    const int c = t / CO, o = t - c * CO;
    const float a = dWc[(size_t)c * 2 * CO + o], b = dWc[(size_t)c * 2 * CO + CO + o];
    dW[(size_t)c * CO + o] = a;
    dW[(size_t)(Cx + c) * CO + o] = b - a;
  }
  if (db && dbc && t < CO) db[t] = dbc[t];
}

// y[(i,j), :] = U[i, :] + V[cloud(i)*npts + idx[i,j], :] + bias;  per-channel sum / sum of squares for the batch norm.
// block 256 = 16 points x 16 float4 columns: a thread owns one point (u loaded once, the k neighbour rows fetched four
// at a time so the idx -> V gathers overlap); persistent grid-stride loop so the moments stay in registers.
__global__ void __launch_bounds__(256)
edge_combine_fwd_kernel(const float* __restrict__ UV, long long ldu, const int32_t* __restrict__ idx,
                        const float* __restrict__ bias, long long P, int k, int npts, float* __restrict__ y,
                        double* __restrict__ stats, float* __restrict__ MM) {
  __shared__ float red[2][CO];
  const int c4 = threadIdx.x & (C4 - 1), pl = threadIdx.x >> 4;
  const float4 b4 = bias ? *reinterpret_cast<const float4*>(bias + c4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
  const float* vbase = UV + CO + c4 * 4;
  for (long long i = (long long)blockIdx.x * 16 + pl; i < P; i += (long long)gridDim.x * 16) {
    const long long cloud0 = (i / npts) * npts;
    float4 u = *reinterpret_cast<const float4*>(UV + i * ldu + c4 * 4);
    u.x += b4.x; u.y += b4.y; u.z += b4.z; u.w += b4.w;
    const int32_t* ip = idx + i * k;
    float* yp = y + i * k * CO + c4 * 4;
    float4 mx = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY), mn = make_float4(INFINITY, INFINITY, INFINITY, INFINITY);
    int j = 0;
    for (; j + 4 <= k; j += 4) {
      int nb[4];
      float4 v[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) nb[t] = ip[j + t];
#pragma unroll
      for (int t = 0; t < 4; ++t) v[t] = *reinterpret_cast<const float4*>(vbase + (cloud0 + nb[t]) * ldu);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        float4 o;
        o.x = u.x + v[t].x; o.y = u.y + v[t].y; o.z = u.z + v[t].z; o.w = u.w + v[t].w;
        __stcs(reinterpret_cast<float4*>(yp + (size_t)(j + t) * CO), o);
        mx.x = fmaxf(mx.x, o.x); mx.y = fmaxf(mx.y, o.y); mx.z = fmaxf(mx.z, o.z); mx.w = fmaxf(mx.w, o.w);
        mn.x = fminf(mn.x, o.x); mn.y = fminf(mn.y, o.y); mn.z = fminf(mn.z, o.z); mn.w = fminf(mn.w, o.w);
        s1.x += o.x; s1.y += o.y; s1.z += o.z; s1.w += o.w;
        s2.x = fmaf(o.x, o.x, s2.x); s2.y = fmaf(o.y, o.y, s2.y); s2.z = fmaf(o.z, o.z, s2.z); s2.w = fmaf(o.w, o.w, s2.w);
      }
    }
    for (; j < k; ++j) {
      const float4 v = *reinterpret_cast<const float4*>(vbase + (cloud0 + ip[j]) * ldu);
      float4 o;
      o.x = u.x + v.x; o.y = u.y + v.y; o.z = u.z + v.z; o.w = u.w + v.w;
      __stcs(reinterpret_cast<float4*>(yp + (size_t)j * CO), o);
      mx.x = fmaxf(mx.x, o.x); mx.y = fmaxf(mx.y, o.y); mx.z = fmaxf(mx.z, o.z); mx.w = fmaxf(mx.w, o.w);
      mn.x = fminf(mn.x, o.x); mn.y = fminf(mn.y, o.y); mn.z = fminf(mn.z, o.z); mn.w = fminf(mn.w, o.w);
      s1.x += o.x; s1.y += o.y; s1.z += o.z; s1.w += o.w;
      s2.x = fmaf(o.x, o.x, s2.x); s2.y = fmaf(o.y, o.y, s2.y); s2.z = fmaf(o.z, o.z, s2.z); s2.w = fmaf(o.w, o.w, s2.w);
    }
    if (MM) {   // per-point max / min over the k rows of the pre-BN output: the max over k of relu(bn(.)) follows from them
      *reinterpret_cast<float4*>(MM + i * 2 * CO + c4 * 4) = mx;
      *reinterpret_cast<float4*>(MM + i * 2 * CO + CO + c4 * 4) = mn;
    }
  }
  if (!stats) return;
  if (threadIdx.x < 2 * CO) (&red[0][0])[threadIdx.x] = 0.f;
  __syncthreads();
  atomicAdd(&red[0][c4 * 4 + 0], s1.x); atomicAdd(&red[0][c4 * 4 + 1], s1.y);
  atomicAdd(&red[0][c4 * 4 + 2], s1.z); atomicAdd(&red[0][c4 * 4 + 3], s1.w);
  atomicAdd(&red[1][c4 * 4 + 0], s2.x); atomicAdd(&red[1][c4 * 4 + 1], s2.y);
  atomicAdd(&red[1][c4 * 4 + 2], s2.z); atomicAdd(&red[1][c4 * 4 + 3], s2.w);
  __syncthreads();
  if (threadIdx.x < CO) {
    atomicAdd(stats + threadIdx.x, (double)red[0][threadIdx.x]);
    atomicAdd(stats + CO + threadIdx.x, (double)red[1][threadIdx.x]);
  }
}

// out[p, c] = max_r relu(sc*y_r + sh) from the per-point extrema of y: fmaf(y, sc, sh) is monotone in y (rounding preserves
// monotonicity), so the maximum is attained at max_r y (sc >= 0) or min_r y (sc < 0) -- bit-identical to maxk_fwd_kernel.
__global__ void maxk_from_extrema_kernel(const float* __restrict__ MM, const float* __restrict__ sc, const float* __restrict__ sh,
                                         long long P, float* __restrict__ out, long long ldo) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= P * CO) return;
  const long long p = t / CO;
  const int c = (int)(t - p * CO);
  const float s = sc[c];
  const float ye = s >= 0.f ? MM[p * 2 * CO + c] : MM[p * 2 * CO + CO + c];
  out[p * ldo + c] = fmaxf(fmaf(ye, s, sh[c]), 0.f);
}

// dy = c1*G + c2 + c3*y (batch-norm backward folded to an affine map; c1 == NULL: dy = G);
// DUV[i, 0:64] = sum_j dy[(i,j), :];  DUV[cloud(i)*npts + idx[i,j], 64:128] += dy[(i,j), :]  (vector reductions).
// block 256 = 16 points x 16 float4 columns.
// MAXK: G is not read but synthesised from the max over k (WSPC_OP_DY_MAXK): G = (relu(y*sc+sh) == MS[i,c]) ? MS[i,64+c] : 0.
template <bool MAXK>
__global__ void __launch_bounds__(256)
edge_combine_bwd_kernel(const float* __restrict__ G, const float* __restrict__ y, const float* __restrict__ c1,
                        const float* __restrict__ c2, const float* __restrict__ c3, const int32_t* __restrict__ idx,
                        long long P, int k, int npts, float* __restrict__ DUV, long long ldd,
                        const float* __restrict__ sc, const float* __restrict__ sh, const float* __restrict__ MS) {
  const int c4 = threadIdx.x & (C4 - 1), pl = threadIdx.x >> 4;
  const long long i = (long long)blockIdx.x * 16 + pl;
  if (i >= P) return;
  float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f), h4 = s4, m4 = s4, q4 = s4;
  if (MAXK) {
    s4 = *reinterpret_cast<const float4*>(sc + c4 * 4);
    h4 = *reinterpret_cast<const float4*>(sh + c4 * 4);
    m4 = *reinterpret_cast<const float4*>(MS + i * 2 * CO + c4 * 4);
    q4 = *reinterpret_cast<const float4*>(MS + i * 2 * CO + CO + c4 * 4);
  }
  float4 a1 = make_float4(1.f, 1.f, 1.f, 1.f), a2 = make_float4(0.f, 0.f, 0.f, 0.f), a3 = a2;
  if (c1) {
    a1 = *reinterpret_cast<const float4*>(c1 + c4 * 4);
    a2 = *reinterpret_cast<const float4*>(c2 + c4 * 4);
    a3 = *reinterpret_cast<const float4*>(c3 + c4 * 4);
  }
  const long long cloud0 = (i / npts) * npts;
  const float* gp = G + i * k * CO + c4 * 4;
  const float* yp = y + i * k * CO + c4 * 4;
  const int32_t* ip = idx + i * k;
  float* dv = DUV + CO + c4 * 4;
  float4 du = make_float4(0.f, 0.f, 0.f, 0.f);
  int j = 0;
  for (; j + 4 <= k; j += 4) {
    float4 g[4], yy[4];
    int nb[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      yy[u] = c1 ? __ldcs(reinterpret_cast<const float4*>(yp + (size_t)(j + u) * CO)) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (MAXK) {
        g[u] = make_float4(maxk_grad(yy[u].x, s4.x, h4.x, m4.x, q4.x), maxk_grad(yy[u].y, s4.y, h4.y, m4.y, q4.y),
                           maxk_grad(yy[u].z, s4.z, h4.z, m4.z, q4.z), maxk_grad(yy[u].w, s4.w, h4.w, m4.w, q4.w));
      } else {
        g[u] = __ldcs(reinterpret_cast<const float4*>(gp + (size_t)(j + u) * CO));
      }
      nb[u] = ip[j + u];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float4 d;
      d.x = fmaf(a3.x, yy[u].x, fmaf(a1.x, g[u].x, a2.x));
      d.y = fmaf(a3.y, yy[u].y, fmaf(a1.y, g[u].y, a2.y));
      d.z = fmaf(a3.z, yy[u].z, fmaf(a1.z, g[u].z, a2.z));
      d.w = fmaf(a3.w, yy[u].w, fmaf(a1.w, g[u].w, a2.w));
      du.x += d.x; du.y += d.y; du.z += d.z; du.w += d.w;
      red_add_v4(dv + (cloud0 + nb[u]) * ldd, d);
    }
  }
  for (; j < k; ++j) {
    const float4 yy = c1 ? __ldcs(reinterpret_cast<const float4*>(yp + (size_t)j * CO)) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 g = MAXK ? make_float4(maxk_grad(yy.x, s4.x, h4.x, m4.x, q4.x), maxk_grad(yy.y, s4.y, h4.y, m4.y, q4.y),
                                        maxk_grad(yy.z, s4.z, h4.z, m4.z, q4.z), maxk_grad(yy.w, s4.w, h4.w, m4.w, q4.w))
                          : __ldcs(reinterpret_cast<const float4*>(gp + (size_t)j * CO));
    float4 d;
    d.x = fmaf(a3.x, yy.x, fmaf(a1.x, g.x, a2.x));
    d.y = fmaf(a3.y, yy.y, fmaf(a1.y, g.y, a2.y));
    d.z = fmaf(a3.z, yy.z, fmaf(a1.z, g.z, a2.z));
    d.w = fmaf(a3.w, yy.w, fmaf(a1.w, g.w, a2.w));
    du.x += d.x; du.y += d.y; du.z += d.z; du.w += d.w;
    red_add_v4(dv + (cloud0 + ip[j]) * ldd, d);
  }
  *reinterpret_cast<float4*>(DUV + i * ldd + c4 * 4) = du;
}

}  // namespace
}  // namespace wspc

using namespace wspc;

extern "C" int wspc_edge_split_weights(const float* W, int Cx, int Cout, float* Wc, wspc_stream_t stream) {
  if (int rc = check_arch()) return rc;
  WSPC_REQUIRE(W && Wc, "edge_split_weights: null pointer");
  WSPC_REQUIRE(Cx >= 1 && Cout == CO, "edge_split_weights: Cx=%d Cout=%d (Cout must be %d)", Cx, Cout, CO);
  edge_split_weights_kernel<<<(Cx * CO + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(W, Cx, Wc);
  count_launch();
  WSPC_LAUNCH_CHECK("edge_split_weights_kernel");
  return WSPC_OK;
}

extern "C" int wspc_edge_merge_wgrad(const float* dWc, const float* dbc, int Cx, int Cout, float* dW, float* db,
                                     wspc_stream_t stream) {
  if (int rc = check_arch()) return rc;
  WSPC_REQUIRE(dWc && dW, "edge_merge_wgrad: null pointer");
  WSPC_REQUIRE(Cx >= 1 && Cout == CO, "edge_merge_wgrad: Cx=%d Cout=%d (Cout must be %d)", Cx, Cout, CO);
  edge_merge_wgrad_kernel<<<(Cx * CO + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(dWc, dbc, Cx, dW, db);
  count_launch();
  WSPC_LAUNCH_CHECK("edge_merge_wgrad_kernel");
  return WSPC_OK;
}

extern "C" int wspc_edge_combine_fwd(const float* UV, long long ldu, const int32_t* idx, const float* bias, long long P,
                                     int k, int npts, int Cout, float* y, double* stats, wspc_stream_t stream) {
  return wspc_edge_combine_fwd_extrema(UV, ldu, idx, bias, P, k, npts, Cout, y, stats, nullptr, stream);
}

extern "C" int wspc_maxk_from_extrema(const float* MM, const float* sc, const float* sh, long long P, int C, float* out,
                                      long long ldo, wspc_stream_t stream) {
  if (int rc = check_arch()) return rc;
  WSPC_REQUIRE(MM && sc && sh && out && P >= 1 && C == CO && ldo >= C, "maxk_from_extrema: bad arguments (C must be %d)", CO);
  maxk_from_extrema_kernel<<<(unsigned)((P * CO + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(MM, sc, sh, P, out,
                                                                                                              ldo);
  count_launch();
  WSPC_LAUNCH_CHECK("maxk_from_extrema_kernel");
  return WSPC_OK;
}

extern "C" int wspc_edge_combine_fwd_extrema(const float* UV, long long ldu, const int32_t* idx, const float* bias, long long P,
                                             int k, int npts, int Cout, float* y, double* stats, float* MM,
                                             wspc_stream_t stream) {
  if (int rc = check_arch()) return rc;
  WSPC_REQUIRE(UV && idx && y, "edge_combine_fwd: null pointer");
  WSPC_REQUIRE(Cout == CO && ldu >= 2 * CO && (ldu & 3) == 0, "edge_combine_fwd: Cout=%d ldu=%lld", Cout, ldu);
  WSPC_REQUIRE(P >= 1 && k >= 1 && npts >= 1 && P % npts == 0, "edge_combine_fwd: bad shape P=%lld k=%d npts=%d", P, k, npts);
  WSPC_REQUIRE(aligned16(UV) && aligned16(y) && (!bias || aligned16(bias)), "edge_combine_fwd: pointers must be 16-byte aligned");
  const long long chunks = (P + 15) / 16;
  const unsigned grid = (unsigned)(chunks < 8LL * kNumSM ? chunks : 8LL * kNumSM);
  WSPC_REQUIRE(!MM || aligned16(MM), "edge_combine_fwd: MM must be 16-byte aligned");
  edge_combine_fwd_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(UV, ldu, idx, bias, P, k, npts, y, stats, MM);
  count_launch();
  WSPC_LAUNCH_CHECK("edge_combine_fwd_kernel");
  return WSPC_OK;
}

extern "C" int wspc_edge_combine_bwd(const float* G, const float* y, const float* c1, const float* c2, const float* c3,
                                     const int32_t* idx, long long P, int k, int npts, int Cout, float* DUV, long long ldd,
                                     wspc_stream_t stream) {
  if (int rc = check_arch()) return rc;
  WSPC_REQUIRE(G && idx && DUV, "edge_combine_bwd: null pointer");
  WSPC_REQUIRE(!c1 || (y && c2 && c3), "edge_combine_bwd: c1 given without y/c2/c3");
  WSPC_REQUIRE(Cout == CO && ldd >= 2 * CO && (ldd & 3) == 0, "edge_combine_bwd: Cout=%d ldd=%lld", Cout, ldd);
  WSPC_REQUIRE(P >= 1 && k >= 1 && npts >= 1 && P % npts == 0, "edge_combine_bwd: bad shape P=%lld k=%d npts=%d", P, k, npts);
  WSPC_REQUIRE(aligned16(G) && aligned16(DUV) && (!y || aligned16(y)) && (!c1 || (aligned16(c1) && aligned16(c2) && aligned16(c3))),
               "edge_combine_bwd: pointers must be 16-byte aligned");
  edge_combine_bwd_kernel<false><<<(unsigned)((P + 15) / 16), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      G, y, c1, c2, c3, idx, P, k, npts, DUV, ldd, nullptr, nullptr, nullptr);
  count_launch();
  WSPC_LAUNCH_CHECK("edge_combine_bwd_kernel");
  return WSPC_OK;
}

extern "C" int wspc_edge_combine_bwd_maxk(const float* y, const float* c1, const float* c2, const float* c3, const float* sc,
                                          const float* sh, const float* MS, const int32_t* idx, long long P, int k, int npts,
                                          int Cout, float* DUV, long long ldd, wspc_stream_t stream) {
  if (int rc = check_arch()) return rc;
  WSPC_REQUIRE(y && c1 && c2 && c3 && sc && sh && MS && idx && DUV, "edge_combine_bwd_maxk: null pointer");
  WSPC_REQUIRE(Cout == CO && ldd >= 2 * CO && (ldd & 3) == 0, "edge_combine_bwd_maxk: Cout=%d ldd=%lld", Cout, ldd);
  WSPC_REQUIRE(P >= 1 && k >= 1 && npts >= 1 && P % npts == 0, "edge_combine_bwd_maxk: bad shape P=%lld k=%d npts=%d", P, k, npts);
  WSPC_REQUIRE(aligned16(y) && aligned16(DUV) && aligned16(MS) && aligned16(c1) && aligned16(c2) && aligned16(c3) &&
               aligned16(sc) && aligned16(sh), "edge_combine_bwd_maxk: pointers must be 16-byte aligned");
  edge_combine_bwd_kernel<true><<<(unsigned)((P + 15) / 16), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      nullptr, y, c1, c2, c3, idx, P, k, npts, DUV, ldd, sc, sh, MS);
  count_launch();
  WSPC_LAUNCH_CHECK("edge_combine_bwd_kernel<maxk>");
  return WSPC_OK;
}
