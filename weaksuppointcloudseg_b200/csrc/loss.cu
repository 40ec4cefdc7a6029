// Segmentation head losses of the weakly supervised trainers, forward and backward fused.
//
// Reference (S3DIS/S3DIS_DGCNN_trainer.py, same block in ShapeNet/ShapeNet_DGCNN_trainer.py):
//   :86      Z_prob = softmax(Z)
//   :89-90   loss_seg = sum(Mask * softmax_cross_entropy_with_logits(Y, Z)) / sum(Mask)
//   :128     loss_siamese = w * mean_{pair,point} sum_c (P[0::2] - P[1::2])^2        (w = 10 S3DIS, 1 ShapeNet)
//   :131-134 L_gt = max_n Y ; L = max_n Z ; loss_inexact = mean sigmoid_cross_entropy_with_logits(L_gt, L)
//   :137     loss_smooth = Loss_SpatialColorSmooth_add_SelfContain(Z_prob, X)   (Util/SmoothConstraint.py:130-167)
// Gradients follow SURVEY.md App. E (TF autodiff): reduce_max splits equally among ties.
#include "common.cuh"
#include <math_constants.h>

namespace wspc {
void count_launch(int n = 1);
namespace {

// monotone float <-> uint mapping so that atomicMax works on floats of either sign
__device__ __forceinline__ unsigned enc_f(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float dec_f(unsigned e) {
  const unsigned u = (e & 0x80000000u) ? (e & 0x7fffffffu) : ~e;
  return __uint_as_float(u);
}

constexpr int MAXC = 64;  // classes per point (13 S3DIS, 50 ShapeNet)

// acc: [0] sum mask*CE  [1] sum mask  [2] sum siamese d^2  [3] sum inexact terms  [4] sum smooth terms
// pass 1: softmax, masked CE, per-(cloud,class) max of logits and of labels. One warp per point.
__global__ void __launch_bounds__(256)
head_softmax_kernel(const float* __restrict__ Z, const float* __restrict__ Y, const float* __restrict__ Mask, int B, int N,
                    int C, float* __restrict__ P, unsigned* __restrict__ colmaxZ, unsigned* __restrict__ colmaxY,
                    double* __restrict__ acc) {
  __shared__ unsigned smz[MAXC], smy[MAXC];
  __shared__ float sred[2][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  if (threadIdx.x < MAXC) { smz[threadIdx.x] = 0u; smy[threadIdx.x] = 0u; }
  __syncthreads();
  float ce_sum = 0.f, m_sum = 0.f;
  unsigned mz0 = 0u, mz1 = 0u, my0 = 0u, my1 = 0u;
  for (int n = blockIdx.x * 8 + warp; n < N; n += gridDim.x * 8) {
    const size_t row = ((size_t)b * N + n) * C;
    const float z0 = (lane < C) ? Z[row + lane] : -CUDART_INF_F;
    const float z1 = (lane + 32 < C) ? Z[row + lane + 32] : -CUDART_INF_F;
    const float mx = warp_max(fmaxf(z0, z1));
    const float e0 = (lane < C) ? expf(z0 - mx) : 0.f;
    const float e1 = (lane + 32 < C) ? expf(z1 - mx) : 0.f;
    const float se = warp_sum(e0 + e1);
    const float lse = mx + logf(se);
    const float inv = 1.f / se;
    float ce = 0.f;
    if (lane < C) {
      P[row + lane] = e0 * inv;
      const float y = Y[row + lane];
      ce += y * (lse - z0);
      mz0 = max(mz0, enc_f(z0));
      my0 = max(my0, enc_f(y));
    }
    if (lane + 32 < C) {
      P[row + lane + 32] = e1 * inv;
      const float y = Y[row + lane + 32];
      ce += y * (lse - z1);
      mz1 = max(mz1, enc_f(z1));
      my1 = max(my1, enc_f(y));
    }
    ce = warp_sum(ce);
    const float mk = Mask[(size_t)b * N + n];
    ce_sum += mk * ce;
    m_sum += mk;
  }
  if (lane < C) { atomicMax(&smz[lane], mz0); atomicMax(&smy[lane], my0); }
  if (lane + 32 < C) { atomicMax(&smz[lane + 32], mz1); atomicMax(&smy[lane + 32], my1); }
  if (lane == 0) { sred[0][warp] = ce_sum; sred[1][warp] = m_sum; }
  __syncthreads();
  if (threadIdx.x < C) {
    atomicMax(&colmaxZ[(size_t)b * C + threadIdx.x], smz[threadIdx.x]);
    atomicMax(&colmaxY[(size_t)b * C + threadIdx.x], smy[threadIdx.x]);
  }
  if (threadIdx.x == 0) {
    float a = 0.f, m = 0.f;
    for (int w = 0; w < 8; ++w) { a += sred[0][w]; m += sred[1][w]; }
    atomicAdd(&acc[0], (double)a);
    atomicAdd(&acc[1], (double)m);
  }
}

// pass 2: number of points attaining the per-(cloud,class) logit max (reduce_max tie split)
__global__ void head_tiecount_kernel(const float* __restrict__ Z, int B, int N, int C,
                                     const unsigned* __restrict__ colmaxZ, int* __restrict__ tiecnt) {
  __shared__ int sc[MAXC];
  const int b = blockIdx.y;
  if (threadIdx.x < MAXC) sc[threadIdx.x] = 0;
  __syncthreads();
  const long long total = (long long)N * C;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(t % C);
    if (Z[(size_t)b * N * C + t] == dec_f(colmaxZ[(size_t)b * C + c])) atomicAdd(&sc[c], 1);
  }
  __syncthreads();
  if (threadIdx.x < C && sc[threadIdx.x]) atomicAdd(&tiecnt[(size_t)b * C + threadIdx.x], sc[threadIdx.x]);
}

// Smoothness term on the kNN graph of the input points; one warp per point, lanes = classes.
// W = exp(-d / gamma) (SmoothConstraint.py:157); loss = mean_{b,n,r} W * mean_c (P_i - P_j)^2 (:161-163).
// dP_i += g, dP_j -= g with g = 2 W (P_i - P_j) / (C * B*N*knn)   (App. E)
// Small class counts (C <= 16, knn <= 16; the S3DIS head has 13 classes and 10 neighbours): lane = (neighbour slot, half of the
// channels), so a point's neighbours are fetched and differenced in parallel instead of one after the other with 13 of 32 lanes
// busy.  Same quantities and options as smooth_kernel below.
__global__ void __launch_bounds__(256)
smooth_small_kernel(const float* __restrict__ P, const int32_t* __restrict__ idx, const float* __restrict__ dist, int B, int N,
                    int C, int knn, float gamma, float gscale, float* __restrict__ dP, double* __restrict__ acc, float cinv,
                    const int32_t* __restrict__ idx_match, int weights_direct, int want_global) {
  __shared__ float sred[8], sred_w[8], sred_s[8];
  if (cinv < 0.f) cinv = 1.f / (float)C;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long pt = (long long)blockIdx.x * 8 + warp;
  const long long total = (long long)B * N;
  float lsum = 0.f, wsum = 0.f, sssum = 0.f;
  if (pt < total) {
    const int slot = lane >> 1, half = lane & 1, c0 = half * 8;
    const long long base = (pt / N) * N;
    const bool act = slot < knn;
    long long j = pt;
    float w = 0.f;
    if (act) {
      const int nb = idx[pt * knn + slot];
      j = base + nb;
      w = weights_direct ? dist[pt * knn + slot] : expf((-dist[pt * knn + slot]) / gamma);
      if (idx_match && idx_match[pt * knn + slot] != nb) w = 0.f;   // knn_mask of SmoothConstraint.py:113
    }
    float d[8], ss = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = c0 + i;
      d[i] = (act && c < C) ? P[pt * C + c] - P[j * C + c] : 0.f;
      ss = fmaf(d[i], d[i], ss);
    }
    ss += __shfl_xor_sync(0xffffffffu, ss, 1);                       // both channel halves of the slot
    const bool once = act && half == 0;
    lsum = warp_sum(once ? w * (ss * cinv) : 0.f);
    wsum = warp_sum(once ? w : 0.f);
    sssum = warp_sum(once ? ss : 0.f);
    if (dP) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = c0 + i;
        const float g = 2.f * w * d[i] * gscale;
        if (act && j != pt && c < C) atomicAdd(dP + j * C + c, -g);
        float gi = g;                                                // the point's own share: sum over the slots
        gi += __shfl_xor_sync(0xffffffffu, gi, 2);
        gi += __shfl_xor_sync(0xffffffffu, gi, 4);
        gi += __shfl_xor_sync(0xffffffffu, gi, 8);
        gi += __shfl_xor_sync(0xffffffffu, gi, 16);
        if (slot == 0 && c < C) atomicAdd(dP + pt * C + c, gi);
      }
    }
  }
  if (lane == 0) { sred[warp] = lsum; sred_w[warp] = wsum; sred_s[warp] = sssum; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, aw = 0.f, as = 0.f;
    for (int w8 = 0; w8 < 8; ++w8) { a += sred[w8]; aw += sred_w[w8]; as += sred_s[w8]; }
    atomicAdd(&acc[4], (double)a);
    if (want_global) {
      atomicAdd(&acc[5], (double)aw);
      atomicAdd(&acc[6], (double)as);
    }
  }
}

__global__ void __launch_bounds__(256)
smooth_kernel(const float* __restrict__ P, const int32_t* __restrict__ idx, const float* __restrict__ dist, int B, int N,
              int C, int knn, float gamma, float gscale, float* __restrict__ dP, double* __restrict__ acc,
              float cinv /* 1/C (mean over channels) or 1 (sum) */ = -1.f, const int32_t* __restrict__ idx_match = nullptr,
              int weights_direct = 0, int want_global = 0) {
  __shared__ float sred[8], sred_w[8], sred_s[8];
  if (cinv < 0.f) cinv = 1.f / (float)C;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long pt = (long long)blockIdx.x * 8 + warp;
  const long long total = (long long)B * N;
  float lsum = 0.f, wsum = 0.f, sssum = 0.f;
  if (pt < total) {
    const long long base = (pt / N) * N;
    const float* pi = P + pt * C;
    const float zi0 = (lane < C) ? pi[lane] : 0.f;
    const float zi1 = (lane + 32 < C) ? pi[lane + 32] : 0.f;
    float gi0 = 0.f, gi1 = 0.f;
    for (int r = 0; r < knn; ++r) {
      const long long j = base + idx[pt * knn + r];
      float w = weights_direct ? dist[pt * knn + r] : expf((-dist[pt * knn + r]) / gamma);
      // knn_mask of SmoothConstraint.py:113: the slot counts only where a second graph holds the same neighbour
      if (idx_match && idx_match[pt * knn + r] != idx[pt * knn + r]) w = 0.f;
      const float* pj = P + j * C;
      const float d0 = (lane < C) ? zi0 - pj[lane] : 0.f;
      const float d1 = (lane + 32 < C) ? zi1 - pj[lane + 32] : 0.f;
      const float ss = warp_sum(d0 * d0 + d1 * d1);
      lsum += w * (ss * cinv);
      wsum += w;
      sssum += ss;
      if (dP) {
        const float g0 = 2.f * w * d0 * gscale, g1 = 2.f * w * d1 * gscale;
        gi0 += g0;
        gi1 += g1;
        if (j != pt) {
          if (lane < C) atomicAdd(dP + j * C + lane, -g0);
          if (lane + 32 < C) atomicAdd(dP + j * C + lane + 32, -g1);
        }
      }
    }
    if (dP) {
      if (lane < C) atomicAdd(dP + pt * C + lane, gi0);
      if (lane + 32 < C) atomicAdd(dP + pt * C + lane + 32, gi1);
    }
  }
  if (lane == 0) { sred[warp] = lsum; sred_w[warp] = wsum; sred_s[warp] = sssum; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, aw = 0.f, as = 0.f;
    for (int w = 0; w < 8; ++w) { a += sred[w]; aw += sred_w[w]; as += sred_s[w]; }
    atomicAdd(&acc[4], (double)a);
    if (want_global) {   // SmoothConstraint.py:64-65 multiplies every weight by the sum over ALL edges and channels
      atomicAdd(&acc[5], (double)aw);
      atomicAdd(&acc[6], (double)as);
    }
  }
}

// pass 3: Siamese term + every gradient -> dZ.  One warp per Siamese pair point (handles rows 2b' and 2b'+1).
// dP may be NULL (no smooth term).  `full` = 0 -> only the seg term is differentiated (Plain style).
__global__ void __launch_bounds__(256)
head_grad_kernel(const float* __restrict__ Z, const float* __restrict__ P, const float* __restrict__ Y,
                 const float* __restrict__ Mask, const float* __restrict__ dP, const unsigned* __restrict__ colmaxZ,
                 const unsigned* __restrict__ colmaxY, const int* __restrict__ tiecnt, int B, int N, int C,
                 float siam_w, int full, float* __restrict__ dZ, double* __restrict__ acc) {
  __shared__ float sred[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int npair = B / 2;
  const long long total = (long long)npair * N;
  const long long t = (long long)blockIdx.x * 8 + warp;
  const float msum = (float)acc[1];
  const float siam_scale = full ? siam_w * 2.f / (float)((double)npair * N) : 0.f;
  const float inex_scale = full ? 1.f / (float)((double)B * C) : 0.f;
  float d2 = 0.f;
  if (t < total) {
    const int bp = (int)(t / N), n = (int)(t % N);
    float pv[2][2], dpv[2][2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const size_t row = ((size_t)(2 * bp + h) * N + n) * C;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int c = lane + 32 * q;
        pv[h][q] = (c < C) ? P[row + c] : 0.f;
        dpv[h][q] = (c < C && dP && full) ? dP[row + c] : 0.f;
      }
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const float d = pv[0][q] - pv[1][q];
      d2 += d * d;
      dpv[0][q] += siam_scale * d;
      dpv[1][q] -= siam_scale * d;
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int b = 2 * bp + h;
      const size_t row = ((size_t)b * N + n) * C;
      const float dot = warp_sum(dpv[h][0] * pv[h][0] + dpv[h][1] * pv[h][1]);
      const float mk = Mask[(size_t)b * N + n] / msum;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int c = lane + 32 * q;
        if (c < C) {
          float g = pv[h][q] * (dpv[h][q] - dot);            // softmax backward
          g += mk * (pv[h][q] - Y[row + c]);                 // masked CE
          if (full) {
            const float L = dec_f(colmaxZ[(size_t)b * C + c]);
            if (Z[row + c] == L) {                           // inexact: arg-max points share the gradient
              const float Lgt = dec_f(colmaxY[(size_t)b * C + c]);
              const float sig = 1.f / (1.f + expf(-L));
              g += (sig - Lgt) * inex_scale / (float)tiecnt[(size_t)b * C + c];
            }
          }
          dZ[row + c] = g;
        }
      }
    }
  }
  d2 = warp_sum(d2);
  if (lane == 0) sred[warp] = d2;
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f;
    for (int w = 0; w < 8; ++w) a += sred[w];
    atomicAdd(&acc[2], (double)a);
  }
}

// single-thread-block finish: inexact loss value + normalisation of the five scalars
__global__ void head_finalize_kernel(const unsigned* __restrict__ colmaxZ, const unsigned* __restrict__ colmaxY, int B,
                                     int N, int C, int knn, float siam_w, int full, double* __restrict__ acc,
                                     float* __restrict__ losses) {
  __shared__ double red[256];
  double s = 0.0;
  if (full) {
    for (int i = threadIdx.x; i < B * C; i += blockDim.x) {
      const float L = dec_f(colmaxZ[i]), Lgt = dec_f(colmaxY[i]);
      s += (double)(fmaxf(L, 0.f) - L * Lgt + log1pf(expf(-fabsf(L))));   // sigmoid CE [TF]
    }
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    acc[3] = red[0];
    const double seg = acc[0] / acc[1];
    const double siam = full ? siam_w * acc[2] / ((double)(B / 2) * N) : 0.0;
    const double inex = full ? acc[3] / ((double)B * C) : 0.0;
    const double smooth = (full && knn > 0) ? acc[4] / ((double)B * N * knn) : 0.0;
    losses[0] = (float)seg;
    losses[1] = (float)siam;
    losses[2] = (float)inex;
    losses[3] = (float)smooth;
    losses[4] = (float)(seg + siam + inex + smooth);
  }
}

}  // namespace
}  // namespace wspc

using namespace wspc;

extern "C" size_t wspc_head_losses_workspace_bytes(int B, int N, int C) {
  // colmaxZ, colmaxY (uint), tiecnt (int): B*C each ; acc: 8 doubles ; dP: B*N*C floats
  return align_up((size_t)3 * B * C * 4, 256) + 256 + align_up((size_t)B * N * C * 4, 256);
}

extern "C" int wspc_head_losses(const float* Z, const float* Y, const float* Mask, const int32_t* sm_idx,
                                const float* sm_dist, int B, int N, int C, int knn, float gamma, float siam_w, int full,
                                int want_grad, float* P, float* dZ, float* losses, void* workspace,
                                size_t workspace_bytes, wspc_stream_t stream) {
  if (int rc = check_arch()) return rc;
  WSPC_REQUIRE(Z && Y && Mask && P && losses && workspace, "head_losses: null pointer");
  WSPC_REQUIRE(B >= 1 && N >= 1 && C >= 1 && C <= MAXC, "head_losses: bad shape (C <= %d)", MAXC);
  WSPC_REQUIRE(!full || (B % 2) == 0, "head_losses: Siamese pairs need an even batch (B=%d)", B);
  WSPC_REQUIRE(!want_grad || dZ, "head_losses: dZ is null");
  WSPC_REQUIRE(full >= 0 && full <= 2, "head_losses: full must be 0 (Plain), 1 (Full) or 2 (Full values, seg-only gradient)");
  // full == 2: the Full graph with its ramp-up gate closed (S3DIS_DGCNN_trainer.py:100-102): the weak terms are evaluated (the
  // training loop prints them) but multiplied by 0, so only the segmentation term is differentiated -- in ONE pass
  const bool grad_full = full == 1;
  const bool smooth = full && sm_idx && sm_dist && knn > 0;
  if (workspace_bytes < wspc_head_losses_workspace_bytes(B, N, C)) {
    set_error("head_losses: workspace too small");
    return WSPC_ERR_WORKSPACE;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  char* w = static_cast<char*>(workspace);
  const size_t small = align_up((size_t)3 * B * C * 4, 256);
  unsigned* colmaxZ = reinterpret_cast<unsigned*>(w);
  unsigned* colmaxY = colmaxZ + (size_t)B * C;
  int* tiecnt = reinterpret_cast<int*>(colmaxY + (size_t)B * C);
  double* acc = reinterpret_cast<double*>(w + small);
  float* dP = reinterpret_cast<float*>(w + small + 256);
  WSPC_CUDA(cudaMemsetAsync(w, 0, small + 256, st));
  if (smooth && want_grad && grad_full) WSPC_CUDA(cudaMemsetAsync(dP, 0, (size_t)B * N * C * 4, st));

  const int nblk = (N + 7) / 8 < 64 ? (N + 7) / 8 : 64;
  head_softmax_kernel<<<dim3(nblk, B), 256, 0, st>>>(Z, Y, Mask, B, N, C, P, colmaxZ, colmaxY, acc);
  count_launch();
  WSPC_LAUNCH_CHECK("head_softmax_kernel");
  if (full) {
    head_tiecount_kernel<<<dim3(32, B), 256, 0, st>>>(Z, B, N, C, colmaxZ, tiecnt);
    count_launch();
    WSPC_LAUNCH_CHECK("head_tiecount_kernel");
  }
  if (smooth) {
    const long long pts = (long long)B * N;
    const float gscale = 1.f / (float)((double)C * B * N * knn);
    if (C <= 16 && knn <= 16)
      smooth_small_kernel<<<(unsigned)((pts + 7) / 8), 256, 0, st>>>(P, sm_idx, sm_dist, B, N, C, knn, gamma, gscale,
                                                                    (want_grad && grad_full) ? dP : nullptr, acc, -1.f, nullptr, 0, 0);
    else
      smooth_kernel<<<(unsigned)((pts + 7) / 8), 256, 0, st>>>(P, sm_idx, sm_dist, B, N, C, knn, gamma, gscale,
                                                              (want_grad && grad_full) ? dP : nullptr, acc);
    count_launch();
    WSPC_LAUNCH_CHECK("smooth_kernel");
  }
  if (want_grad || full) {
    // the Siamese sum is produced by the gradient pass; it also runs (writing into dZ scratch) when only
    // the loss values are wanted in Full style
    WSPC_REQUIRE(dZ, "head_losses: Full style needs a dZ buffer (also used as scratch)");
    const long long total = (long long)(B / 2) * N;
    if (grad_full) {
      head_grad_kernel<<<(unsigned)((total + 7) / 8), 256, 0, st>>>(Z, P, Y, Mask, (smooth && want_grad) ? dP : nullptr,
                                                                   colmaxZ, colmaxY, tiecnt, B, N, C, siam_w, 1, dZ, acc);
    } else {
      // Plain style: pairs are not meaningful; treat consecutive rows as pairs only for work distribution
      WSPC_REQUIRE((B % 2) == 0, "head_losses: want_grad needs an even batch");
      head_grad_kernel<<<(unsigned)((total + 7) / 8), 256, 0, st>>>(Z, P, Y, Mask, nullptr, colmaxZ, colmaxY, tiecnt, B,
                                                                   N, C, 0.f, 0, dZ, acc);
    }
    count_launch();
    WSPC_LAUNCH_CHECK("head_grad_kernel");
  }
  head_finalize_kernel<<<1, 256, 0, st>>>(colmaxZ, colmaxY, B, N, C, smooth ? knn : 0, siam_w, full ? 1 : 0, acc, losses);
  count_launch();
  WSPC_LAUNCH_CHECK("head_finalize_kernel");
  return WSPC_OK;
}

namespace wspc {
namespace {
__global__ void smooth_finish_kernel(const double* __restrict__ acc, double denom, int global_form, float* __restrict__ loss) {
  loss[0] = (float)((global_form ? acc[5] * acc[6] : acc[4]) / denom);
}
}  // namespace
}  // namespace wspc

// Stand-alone manifold smoothness terms of Util/SmoothConstraint.py on a given kNN graph (see include/wspc.h).
// dZ (B,N,C) may be NULL; if given it must be zeroed by the caller and receives d loss / d Z.
extern "C" int wspc_smooth_loss_ex(const float* Z, const int32_t* idx, const float* dist, const int32_t* idx_match, int B, int N,
                                   int C, int knn, float gamma, int flags, float* dZ, float* loss, void* workspace,
                                   size_t workspace_bytes, wspc_stream_t stream) {
  if (int rc = check_arch()) return rc;
  WSPC_REQUIRE(Z && idx && dist && loss && workspace, "smooth_loss: null pointer");
  WSPC_REQUIRE(C >= 1 && C <= MAXC && knn >= 1 && B >= 1 && N >= 1, "smooth_loss: bad shape (C <= %d)", MAXC);
  WSPC_REQUIRE(workspace_bytes >= 64, "smooth_loss: workspace needs 64 bytes");
  WSPC_REQUIRE((flags & ~(WSPC_SMOOTH_SUM_C | WSPC_SMOOTH_WEIGHTS | WSPC_SMOOTH_GLOBAL_SS)) == 0, "smooth_loss: unknown flags %d", flags);
  const bool global_form = (flags & WSPC_SMOOTH_GLOBAL_SS) != 0;
  WSPC_REQUIRE(!(global_form && dZ), "smooth_loss: the WSPC_SMOOTH_GLOBAL_SS form is forward-only (dZ must be NULL)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  double* acc = static_cast<double*>(workspace);
  WSPC_CUDA(cudaMemsetAsync(acc, 0, 64, st));
  const long long pts = (long long)B * N;
  const float cinv = (flags & WSPC_SMOOTH_SUM_C) ? 1.f : 1.f / (float)C;
  const float gscale = (float)((double)cinv / ((double)B * N * knn));
  if (C <= 16 && knn <= 16)
    smooth_small_kernel<<<(unsigned)((pts + 7) / 8), 256, 0, st>>>(Z, idx, dist, B, N, C, knn, gamma, gscale, dZ, acc, cinv, idx_match,
                                                                  (flags & WSPC_SMOOTH_WEIGHTS) ? 1 : 0, global_form ? 1 : 0);
  else
    smooth_kernel<<<(unsigned)((pts + 7) / 8), 256, 0, st>>>(Z, idx, dist, B, N, C, knn, gamma, gscale, dZ, acc, cinv, idx_match,
                                                             (flags & WSPC_SMOOTH_WEIGHTS) ? 1 : 0, global_form ? 1 : 0);
  smooth_finish_kernel<<<1, 1, 0, st>>>(acc, (double)B * N * knn, global_form ? 1 : 0, loss);
  count_launch(2);
  WSPC_LAUNCH_CHECK("smooth_kernel");
  return WSPC_OK;
}

extern "C" int wspc_smooth_loss(const float* Z, const int32_t* idx, const float* dist, int B, int N, int C, int knn,
                                float gamma, float* dZ, float* loss, void* workspace, size_t workspace_bytes,
                                wspc_stream_t stream) {
  return wspc_smooth_loss_ex(Z, idx, dist, nullptr, B, N, C, knn, gamma, 0, dZ, loss, workspace, workspace_bytes, stream);
}
