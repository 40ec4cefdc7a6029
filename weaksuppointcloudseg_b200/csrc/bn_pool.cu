// Batch-norm bookkeeping, max-over-k / max-over-N pooling (forward + backward) and small reductions.
//
// Reference semantics:
//   batch_norm_dist_template           Networks/dgcnn/utils/tf_util.py:502-535 (biased var, eps 1e-3, pop update)
//   tf.reduce_max(axis=-2)             S3DIS/DGCNN_S3DIS.py:46,62,78   (gradient split equally among ties [TF])
//   tf_util.max_pool2d([N,1])          tf_util.py:357-380, DGCNN_S3DIS.py:85 (gradient to the first arg-max [TF])
#include "operand.cuh"
#include <math_constants.h>

namespace wspc {
void count_launch(int n = 1);
namespace {

// ------------------------------------------------------------ BN finalize ---
__global__ void bn_finalize_kernel(const double* __restrict__ stats, int C, double rows, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float eps, float decay, int training,
                                   float* __restrict__ pop_mean, float* __restrict__ pop_var, float* __restrict__ sc,
                                   float* __restrict__ sh, float* __restrict__ save_mean,
                                   float* __restrict__ save_invstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float mean, var;
  if (training) {
    const double m = stats[c] / rows;
    double v = stats[C + c] / rows - m * m;   // biased variance (tf.nn.moments)
    if (v < 0.0) v = 0.0;
    mean = (float)m;
    var = (float)v;
    pop_mean[c] = pop_mean[c] * decay + mean * (1.f - decay);   // tf_util.py:524
    pop_var[c] = pop_var[c] * decay + var * (1.f - decay);      // tf_util.py:525
  } else {
    mean = pop_mean[c];
    var = pop_var[c];
  }
  const float invstd = rsqrtf(var + eps);
  const float s = invstd * gamma[c];
  sc[c] = s;
  sh[c] = beta[c] - mean * s;
  if (save_mean) save_mean[c] = mean;
  if (save_invstd) save_invstd[c] = invstd;
}

// stats = (sum G, sum G*y) over the rows the layer normalised over.
__global__ void bn_bwd_coeffs_kernel(const double* __restrict__ stats, int C, double rows,
                                     const float* __restrict__ gamma, const float* __restrict__ mean,
                                     const float* __restrict__ invstd, float* __restrict__ c1, float* __restrict__ c2,
                                     float* __restrict__ c3, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double s = invstd[c], mu = mean[c], g = gamma[c];
  const double dB = stats[c];
  const double dG = s * (stats[C + c] - mu * dB);
  const double k1 = g * s;
  const double k3 = -g * s * s * dG / rows;
  const double k2 = -g * s * dB / rows - k3 * mu;
  c1[c] = (float)k1;
  c2[c] = (float)k2;
  c3[c] = (float)k3;
  dgamma[c] = (float)dG;
  dbeta[c] = (float)dB;
}

// ------------------------------------------------------------- max over k ---
// one thread per (point, 4 channels): float4 rows => every warp instruction moves 512 contiguous bytes
__device__ __forceinline__ float4 bnrelu4(float4 v, float4 s, float4 h) {
  return make_float4(fmaxf(fmaf(v.x, s.x, h.x), 0.f), fmaxf(fmaf(v.y, s.y, h.y), 0.f), fmaxf(fmaf(v.z, s.z, h.z), 0.f),
                     fmaxf(fmaf(v.w, s.w, h.w), 0.f));
}

__global__ void __launch_bounds__(256)
maxk_fwd_kernel(const float* __restrict__ y, const float* __restrict__ sc, const float* __restrict__ sh,
                long long P, int k, int C, float* __restrict__ out, long long ldo) {
  const int C4 = C >> 2;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= P * C4) return;
  const long long p = t / C4;
  const int c = (int)(t - p * C4) * 4;
  const float4 s = *reinterpret_cast<const float4*>(sc + c), h = *reinterpret_cast<const float4*>(sh + c);
  const float* yp = y + p * k * C + c;
  float4 m = make_float4(0.f, 0.f, 0.f, 0.f);  // relu output is >= 0
  int r = 0;
  for (; r + 4 <= k; r += 4) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = __ldcs(reinterpret_cast<const float4*>(yp + (size_t)(r + u) * C));
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float4 a = bnrelu4(v[u], s, h);
      m.x = fmaxf(m.x, a.x); m.y = fmaxf(m.y, a.y); m.z = fmaxf(m.z, a.z); m.w = fmaxf(m.w, a.w);
    }
  }
  for (; r < k; ++r) {
    const float4 a = bnrelu4(__ldcs(reinterpret_cast<const float4*>(yp + (size_t)r * C)), s, h);
    m.x = fmaxf(m.x, a.x); m.y = fmaxf(m.y, a.y); m.z = fmaxf(m.z, a.z); m.w = fmaxf(m.w, a.w);
  }
  *reinterpret_cast<float4*>(out + p * ldo + c) = m;
}

__global__ void __launch_bounds__(256)
maxk_bwd_kernel(const float* __restrict__ y, const float* __restrict__ sc, const float* __restrict__ sh,
                const float* __restrict__ out, long long ldo, const float* __restrict__ dout,
                long long lddo, long long P, int k, int C, float* __restrict__ G,
                double* __restrict__ stats) {
  __shared__ float red[2][128];
  const int tid = threadIdx.x;
  const int C4 = C >> 2;
  // block covers (256 / C4) points x C channels (C4 divides 256, C <= 128)
  const long long t = (long long)blockIdx.x * blockDim.x + tid;
  const bool valid = t < P * C4;
  const long long p = valid ? t / C4 : 0;
  const int c = valid ? (int)(t - p * C4) * 4 : 0;
  (&red[0][0])[tid] = 0.f;  // blockDim.x == 256 == 2*128
  __syncthreads();
  float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
  if (valid) {
    const float4 s = *reinterpret_cast<const float4*>(sc + c), h = *reinterpret_cast<const float4*>(sh + c);
    const float4 m4 = *reinterpret_cast<const float4*>(out + p * ldo + c);
    const float4 go4 = *reinterpret_cast<const float4*>(dout + p * lddo + c);
    const float m[4] = {m4.x, m4.y, m4.z, m4.w}, go[4] = {go4.x, go4.y, go4.z, go4.w};
    const float* yp = y + p * k * C + c;
    float* gp = G + p * k * C + c;
    int cnt[4] = {0, 0, 0, 0};
    {   // (keeping the k rows in registers for a single read of y was measured slower: 96 live registers halve the occupancy)
      for (int r = 0; r < k; ++r) {   // tie count (reduce_max splits the gradient equally among tied maxima [TF])
        const float4 a = bnrelu4(*reinterpret_cast<const float4*>(yp + (size_t)r * C), s, h);
        cnt[0] += (a.x == m[0]) ? 1 : 0; cnt[1] += (a.y == m[1]) ? 1 : 0;
        cnt[2] += (a.z == m[2]) ? 1 : 0; cnt[3] += (a.w == m[3]) ? 1 : 0;
      }
      float share[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) share[i] = (m[i] > 0.f && cnt[i] > 0) ? go[i] / (float)cnt[i] : 0.f;
      for (int r = 0; r < k; ++r) {   // second visit of the same k rows: L1/L2 hits
        const float4 yv = *reinterpret_cast<const float4*>(yp + (size_t)r * C);
        const float4 a = bnrelu4(yv, s, h);
        float4 g;
        g.x = (m[0] > 0.f && a.x == m[0]) ? share[0] : 0.f;
        g.y = (m[1] > 0.f && a.y == m[1]) ? share[1] : 0.f;
        g.z = (m[2] > 0.f && a.z == m[2]) ? share[2] : 0.f;
        g.w = (m[3] > 0.f && a.w == m[3]) ? share[3] : 0.f;
        __stcs(reinterpret_cast<float4*>(gp + (size_t)r * C), g);
        s0[0] += g.x; s0[1] += g.y; s0[2] += g.z; s0[3] += g.w;
        s1[0] += g.x * yv.x; s1[1] += g.y * yv.y; s1[2] += g.z * yv.z; s1[3] += g.w * yv.w;
      }
    }
  }
  // channels repeat with period C inside the block (256 % C4 == 0 guaranteed by the wrapper)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    atomicAdd(&red[0][c + i], s0[i]);
    atomicAdd(&red[1][c + i], s1[i]);
  }
  __syncthreads();
  if (tid < C) {
    atomicAdd(stats + tid, (double)red[0][tid]);
    atomicAdd(stats + C + tid, (double)red[1][tid]);
  }
}

// Statistics-only backward of the max over k: instead of materialising G (P*k, C) it writes, per point, the pooled maximum
// and the tie-split share  MS[p, 0:C] = max_r a,  MS[p, C:2C] = dout / #ties  (0 when the maximum is not positive), from
// which the GEMM loaders synthesise G on the fly (WSPC_OP_DY_MAXK), and accumulates the BN-backward sums
// stats[0] += sum G, stats[1] += sum G*y.  One read of y.
__global__ void __launch_bounds__(256)
maxk_bwd_stats_kernel(const float* __restrict__ y, const float* __restrict__ sc, const float* __restrict__ sh,
                      const float* __restrict__ out, long long ldo, const float* __restrict__ dout, long long lddo,
                      long long P, int k, int C, float* __restrict__ MS, double* __restrict__ stats) {
  __shared__ float red[2][128];
  const int tid = threadIdx.x;
  const int C4 = C >> 2;
  const long long t = (long long)blockIdx.x * blockDim.x + tid;
  const bool valid = t < P * C4;
  const long long p = valid ? t / C4 : 0;
  const int c = valid ? (int)(t - p * C4) * 4 : 0;
  (&red[0][0])[tid] = 0.f;
  __syncthreads();
  float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
  if (valid) {
    const float4 s = *reinterpret_cast<const float4*>(sc + c), h = *reinterpret_cast<const float4*>(sh + c);
    const float4 m4 = *reinterpret_cast<const float4*>(out + p * ldo + c);
    const float4 go4 = *reinterpret_cast<const float4*>(dout + p * lddo + c);
    const float m[4] = {m4.x, m4.y, m4.z, m4.w}, go[4] = {go4.x, go4.y, go4.z, go4.w};
    const float* yp = y + p * k * C + c;
    float cnt[4] = {0.f, 0.f, 0.f, 0.f}, ys[4] = {0.f, 0.f, 0.f, 0.f};
    int r = 0;
    for (; r + 4 <= k; r += 4) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = __ldcs(reinterpret_cast<const float4*>(yp + (size_t)(r + u) * C));
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float4 a = bnrelu4(v[u], s, h);
        if (a.x == m[0]) { cnt[0] += 1.f; ys[0] += v[u].x; }
        if (a.y == m[1]) { cnt[1] += 1.f; ys[1] += v[u].y; }
        if (a.z == m[2]) { cnt[2] += 1.f; ys[2] += v[u].z; }
        if (a.w == m[3]) { cnt[3] += 1.f; ys[3] += v[u].w; }
      }
    }
    for (; r < k; ++r) {
      const float4 v = __ldcs(reinterpret_cast<const float4*>(yp + (size_t)r * C));
      const float4 a = bnrelu4(v, s, h);
      if (a.x == m[0]) { cnt[0] += 1.f; ys[0] += v.x; }
      if (a.y == m[1]) { cnt[1] += 1.f; ys[1] += v.y; }
      if (a.z == m[2]) { cnt[2] += 1.f; ys[2] += v.z; }
      if (a.w == m[3]) { cnt[3] += 1.f; ys[3] += v.w; }
    }
    float share[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      share[i] = (m[i] > 0.f && cnt[i] > 0.f) ? go[i] / cnt[i] : 0.f;
      s0[i] = share[i] * cnt[i];
      s1[i] = share[i] * ys[i];
    }
    *reinterpret_cast<float4*>(MS + p * 2 * C + c) = m4;
    *reinterpret_cast<float4*>(MS + p * 2 * C + C + c) = make_float4(share[0], share[1], share[2], share[3]);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    atomicAdd(&red[0][c + i], s0[i]);
    atomicAdd(&red[1][c + i], s1[i]);
  }
  __syncthreads();
  if (tid < C) {
    atomicAdd(stats + tid, (double)red[0][tid]);
    atomicAdd(stats + C + tid, (double)red[1][tid]);
  }
}

// ------------------------------------------------------------- max over N ---
// grid (C/32, B); block 32 x 8: lane = channel, 8 point groups
__global__ void maxn_fwd_kernel(const float* __restrict__ y, const float* __restrict__ sc, const float* __restrict__ sh,
                                int N, int C, float* __restrict__ g, int32_t* __restrict__ amax) {
  __shared__ float sv[8][32];
  __shared__ int si[8][32];
  const int b = blockIdx.y;
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int grp = threadIdx.y;
  float best = -1.f;
  int bi = 0x7fffffff;
  if (c < C) {
    const float s = sc[c], h = sh[c];
    const float* yb = y + (size_t)b * N * C + c;
    for (int n = grp; n < N; n += 8) {
      const float a = fmaxf(fmaf(yb[(size_t)n * C], s, h), 0.f);
      if (a > best) { best = a; bi = n; }   // strict > keeps the first index within this group
    }
  }
  sv[grp][threadIdx.x] = best;
  si[grp][threadIdx.x] = bi;
  __syncthreads();
  if (grp == 0 && c < C) {
    for (int q = 1; q < 8; ++q) {
      const float v = sv[q][threadIdx.x];
      const int i = si[q][threadIdx.x];
      if (v > best || (v == best && i < bi)) { best = v; bi = i; }
    }
    g[(size_t)b * C + c] = best;
    amax[(size_t)b * C + c] = bi;
  }
}

// gate the incoming gradient by the ReLU (g > 0) and accumulate the BN-backward statistics of the sparse G
__global__ void maxn_bwd_gate_kernel(const float* __restrict__ g, const float* __restrict__ dgin,
                                     const int32_t* __restrict__ amax, const float* __restrict__ y, int B, int N, int C,
                                     float* __restrict__ dg, double* __restrict__ stats) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s0 = 0.0, s1 = 0.0;
  for (int b = 0; b < B; ++b) {
    const size_t o = (size_t)b * C + c;
    const float v = (g[o] > 0.f) ? dgin[o] : 0.f;
    dg[o] = v;
    s0 += v;
    s1 += (double)v * (double)y[((size_t)b * N + amax[o]) * C + c];
  }
  stats[c] = s0;
  stats[C + c] = s1;
}

// ------------------------------------------------------ per-cloud col sums --
// S[b, c] = sum_n dY[b, n, c] with dY read through an OP_DY operand.  grid (C/32, B), block 32 x 8
template <int GMODE>
__global__ void cloud_colsum_kernel(const Operand G, int N, float* __restrict__ S) {
  __shared__ float red[8][32];
  const int b = blockIdx.y;
  const int c = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (c < G.C) {
    for (int n = threadIdx.y; n < N; n += 8) {
      const long long row = (long long)b * N + n;
      const float g = G.p[row * G.ld + c];
      s += G.c1 ? fmaf(G.c1[c], g, fmaf(G.c3[c], G.y[row * G.ldy + c], G.c2[c])) : g;
    }
  }
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < G.C) {
    float t = 0.f;
    for (int q = 0; q < 8; ++q) t += red[q][threadIdx.x];
    S[(size_t)b * G.C + c] = t;
  }
}

// float4 variant (C % 4 == 0, 16-byte aligned rows): grid (C/128, B), block 32 x 16; four rows of loads in flight per thread
__global__ void __launch_bounds__(512)
cloud_colsum4_kernel(const Operand G, int N, float* __restrict__ S) {
  __shared__ float4 red[16][32];
  const int b = blockIdx.y;
  const int c = (blockIdx.x * 32 + threadIdx.x) * 4;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < G.C) {
    float4 k1 = make_float4(1.f, 1.f, 1.f, 1.f), k2 = make_float4(0.f, 0.f, 0.f, 0.f), k3 = k2;
    if (G.c1) {
      k1 = *reinterpret_cast<const float4*>(G.c1 + c);
      k2 = *reinterpret_cast<const float4*>(G.c2 + c);
      k3 = *reinterpret_cast<const float4*>(G.c3 + c);
    }
    const float* gp = G.p + (long long)b * N * G.ld + c;
    const float* yp = G.c1 ? G.y + (long long)b * N * G.ldy + c : nullptr;
    for (int n0 = threadIdx.y; n0 < N; n0 += 64) {
      float4 g[4], y[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int n = n0 + 16 * u;
        g[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        y[u] = g[u];
        if (n < N) {
          g[u] = __ldcs(reinterpret_cast<const float4*>(gp + (long long)n * G.ld));
          if (yp) y[u] = __ldcs(reinterpret_cast<const float4*>(yp + (long long)n * G.ldy));
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (n0 + 16 * u < N) {
          if (yp) {
            s.x += fmaf(k1.x, g[u].x, fmaf(k3.x, y[u].x, k2.x));
            s.y += fmaf(k1.y, g[u].y, fmaf(k3.y, y[u].y, k2.y));
            s.z += fmaf(k1.z, g[u].z, fmaf(k3.z, y[u].z, k2.z));
            s.w += fmaf(k1.w, g[u].w, fmaf(k3.w, y[u].w, k2.w));
          } else {
            s.x += g[u].x; s.y += g[u].y; s.z += g[u].z; s.w += g[u].w;
          }
        }
      }
    }
  }
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < G.C) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int q = 0; q < 16; ++q) { t.x += red[q][threadIdx.x].x; t.y += red[q][threadIdx.x].y; t.z += red[q][threadIdx.x].z; t.w += red[q][threadIdx.x].w; }
    *reinterpret_cast<float4*>(S + (size_t)b * G.C + c) = t;
  }
}

}  // namespace
}  // namespace wspc

using namespace wspc;

extern "C" int wspc_bn_finalize(const double* stats, int C, double rows, const float* gamma, const float* beta,
                                float eps, float decay, int training, float* pop_mean, float* pop_var, float* sc,
                                float* sh, float* save_mean, float* save_invstd, wspc_stream_t stream) {
  if (int rc = check_arch()) return rc;
  WSPC_REQUIRE(gamma && beta && pop_mean && pop_var && sc && sh, "bn_finalize: null pointer");
  WSPC_REQUIRE(!training || stats, "bn_finalize: training needs batch statistics");
  WSPC_REQUIRE(C >= 1 && rows >= 1, "bn_finalize: bad shape");
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      stats, C, rows, gamma, beta, eps, decay, training, pop_mean, pop_var, sc, sh, save_mean, save_invstd);
  count_launch();
  WSPC_LAUNCH_CHECK("bn_finalize_kernel");
  return WSPC_OK;
}

extern "C" int wspc_bn_bwd_coeffs(const double* stats, int C, double rows, const float* gamma, const float* mean,
                                  const float* invstd, float* c1, float* c2, float* c3, float* dgamma, float* dbeta,
                                  wspc_stream_t stream) {
  if (int rc = check_arch()) return rc;
  WSPC_REQUIRE(stats && gamma && mean && invstd && c1 && c2 && c3 && dgamma && dbeta, "bn_bwd_coeffs: null pointer");
  bn_bwd_coeffs_kernel<<<(C + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      stats, C, rows, gamma, mean, invstd, c1, c2, c3, dgamma, dbeta);
  count_launch();
  WSPC_LAUNCH_CHECK("bn_bwd_coeffs_kernel");
  return WSPC_OK;
}

extern "C" int wspc_maxk_bnrelu_fwd(const float* y, const float* sc, const float* sh, long long P, int k, int C,
                                    float* out, long long ldo, wspc_stream_t stream) {
  if (int rc = check_arch()) return rc;
  WSPC_REQUIRE(y && sc && sh && out, "maxk_fwd: null pointer");
  WSPC_REQUIRE(P >= 1 && k >= 1 && C >= 4 && (C & 3) == 0 && ldo >= C && (ldo & 3) == 0, "maxk_fwd: bad shape (C, ldo multiples of 4)");
  WSPC_REQUIRE(aligned16(y) && aligned16(sc) && aligned16(sh) && aligned16(out), "maxk_fwd: pointers must be 16-byte aligned");
  const long long total = P * (C >> 2);
  maxk_fwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(y, sc, sh, P, k,
                                                                                                     C, out, ldo);
  count_launch();
  WSPC_LAUNCH_CHECK("maxk_fwd_kernel");
  return WSPC_OK;
}

extern "C" int wspc_maxk_bnrelu_bwd(const float* y, const float* sc, const float* sh, const float* out, long long ldo,
                                    const float* dout, long long lddo, long long P, int k, int C, float* G,
                                    double* stats, wspc_stream_t stream) {
  if (int rc = check_arch()) return rc;
  WSPC_REQUIRE(y && sc && sh && out && dout && G && stats, "maxk_bwd: null pointer");
  WSPC_REQUIRE(C >= 4 && C <= 128 && (C & 3) == 0 && (256 % (C >> 2)) == 0 && (ldo & 3) == 0 && (lddo & 3) == 0,
               "maxk_bwd: C=%d must be a multiple of 4, <= 128, with C/4 dividing 256; ldo/lddo multiples of 4", C);
  WSPC_REQUIRE(aligned16(y) && aligned16(sc) && aligned16(sh) && aligned16(out) && aligned16(dout) && aligned16(G),
               "maxk_bwd: pointers must be 16-byte aligned");
  const long long total = P * (C >> 2);
  maxk_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      y, sc, sh, out, ldo, dout, lddo, P, k, C, G, stats);
  count_launch();
  WSPC_LAUNCH_CHECK("maxk_bwd_kernel");
  return WSPC_OK;
}

extern "C" int wspc_maxk_bnrelu_bwd_stats(const float* y, const float* sc, const float* sh, const float* out, long long ldo,
                                          const float* dout, long long lddo, long long P, int k, int C, float* MS,
                                          double* stats, wspc_stream_t stream) {
  if (int rc = check_arch()) return rc;
  WSPC_REQUIRE(y && sc && sh && out && dout && MS && stats, "maxk_bwd_stats: null pointer");
  WSPC_REQUIRE(C >= 4 && C <= 128 && (C & 3) == 0 && (256 % (C >> 2)) == 0 && (ldo & 3) == 0 && (lddo & 3) == 0,
               "maxk_bwd_stats: C=%d must be a multiple of 4, <= 128, with C/4 dividing 256; ldo/lddo multiples of 4", C);
  WSPC_REQUIRE(aligned16(y) && aligned16(sc) && aligned16(sh) && aligned16(out) && aligned16(dout) && aligned16(MS),
               "maxk_bwd_stats: pointers must be 16-byte aligned");
  const long long total = P * (C >> 2);
  maxk_bwd_stats_kernel<<<(unsigned)((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      y, sc, sh, out, ldo, dout, lddo, P, k, C, MS, stats);
  count_launch();
  WSPC_LAUNCH_CHECK("maxk_bwd_stats_kernel");
  return WSPC_OK;
}

extern "C" int wspc_maxn_bnrelu_fwd(const float* y, const float* sc, const float* sh, int B, int N, int C, float* g,
                                    int32_t* amax, wspc_stream_t stream) {
  if (int rc = check_arch()) return rc;
  WSPC_REQUIRE(y && sc && sh && g && amax, "maxn_fwd: null pointer");
  dim3 grid((C + 31) / 32, B), block(32, 8);
  maxn_fwd_kernel<<<grid, block, 0, reinterpret_cast<cudaStream_t>(stream)>>>(y, sc, sh, N, C, g, amax);
  count_launch();
  WSPC_LAUNCH_CHECK("maxn_fwd_kernel");
  return WSPC_OK;
}

extern "C" int wspc_maxn_bwd_gate(const float* g, const float* dgin, const int32_t* amax, const float* y, int B, int N,
                                  int C, float* dg, double* stats, wspc_stream_t stream) {
  if (int rc = check_arch()) return rc;
  WSPC_REQUIRE(g && dgin && amax && y && dg && stats, "maxn_bwd_gate: null pointer");
  maxn_bwd_gate_kernel<<<(C + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(g, dgin, amax, y, B, N, C,
                                                                                           dg, stats);
  count_launch();
  WSPC_LAUNCH_CHECK("maxn_bwd_gate_kernel");
  return WSPC_OK;
}

extern "C" int wspc_cloud_colsum(const wspc_operand_t* G, int B, int N, float* S, wspc_stream_t stream) {
  if (int rc = check_arch()) return rc;
  WSPC_REQUIRE(G && G->p && S, "cloud_colsum: null pointer");
  const bool vec = (G->C % 4) == 0 && (G->ld % 4) == 0 && aligned16(G->p) && aligned16(S) &&
                   (!G->c1 || ((G->ldy % 4) == 0 && aligned16(G->y) && aligned16(G->c1) && aligned16(G->c2) && aligned16(G->c3)));
  if (vec) {
    dim3 grid((G->C / 4 + 31) / 32, B), block(32, 16);
    cloud_colsum4_kernel<<<grid, block, 0, reinterpret_cast<cudaStream_t>(stream)>>>(*G, N, S);
  } else {
    dim3 grid((G->C + 31) / 32, B), block(32, 8);
    cloud_colsum_kernel<OP_DY><<<grid, block, 0, reinterpret_cast<cudaStream_t>(stream)>>>(*G, N, S);
  }
  count_launch();
  WSPC_LAUNCH_CHECK("cloud_colsum_kernel");
  return WSPC_OK;
}
