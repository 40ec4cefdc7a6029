// Test-time label propagation: symmetric graph Laplacian + closed-form solve, plus small gather / BN-apply ops.
//
// Reference:
//   Tool.TF_Computation.LaplacianMatSym_XYZRGB_DirectComp   Util/Tool.py:435-468
//       W = exp(-1e3 d_xyz) * exp(-1e1 d_rgb), d = clamp((X2_i + X2_j) - 2 X_i.X_j, 0)
//       Lsym = D^-1/2 (diag(d + 1e-8) - W) D^-1/2,  d_i = sum_j W_ij
//       (the reference forms diag matrices and runs two dense N^3 matmuls, :461-465; here L is written directly)
//   LabelPropagation_TF                                        Util/ProbLabelPropagation.py:8-42
//       w = 1 - H_2(G)/log_2 K ; Y = beta (alpha L + beta diag(w) + 1e-5 I)^-1 diag(w) G ; Y_prob = Y / sum_k Y
//       (the reference calls tf.linalg.inv; the system is SPD, so it is solved here by Jacobi-preconditioned
//        conjugate gradients on all K right-hand sides at once — SURVEY App. A-11: any solver reaching 1e-3 on
//        Y_prob is acceptable)
//   Tool.batch_gather_v1 (Util/Tool.py:72-104), tf_util.get_edge_feature (tf_util.py:674-706)
#include "operand.cuh"
#include <vector>

namespace wspc {
void count_launch(int n = 1);
namespace {

__device__ __forceinline__ float sqdist_smooth(const float* a, const float* b, float sa, float sb, int D) {
  float dot = 0.f;
  for (int c = 0; c < D; ++c) dot = __fmaf_rn(a[c], b[c], dot);
  const float d = __fsub_rn(__fadd_rn(sa, sb), __fmul_rn(2.f, dot));
  return d > 0.f ? d : 0.f;
}

// Symmetric normalised Laplacian in two passes over the N x N pair kernel (it is never stored unnormalised):
// pass 1 (WRITE = false): degree d_i = sum_j W_ij;  pass 2: L_ij = ((i == j) (d_i + 1e-8) - W_ij) d_i^-1/2 d_j^-1/2.
// grid (N/16, B), block 256: a CTA owns 16 rows (2 per warp); 128 columns at a time are staged in shared memory and a
// lane evaluates columns lane, lane+32, lane+64, lane+96 of the tile -- four independent exp chains per thread, 16 warps
// per SM with two resident CTAs, and pass 2 stores 128 contiguous bytes per warp and row.  The row sums are reduced by a
// fixed-order butterfly, so the result does not depend on scheduling.
constexpr int LAP_ROWS = 16;
template <bool WRITE>
__global__ void __launch_bounds__(256)
laplacian_kernel(const float* __restrict__ X, const float* __restrict__ RGB, int N, int D1, int D2, float s1, float s2,
                 float* __restrict__ deg, float* __restrict__ Lout) {
  __shared__ float sx[128][3], sc[128][3], ssx[128], ssc[128], sdeg[128];
  const int b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* Xb = X + (size_t)b * N * D1;
  const float* Cb = RGB + (size_t)b * N * D2;
  float xi[2][3], ci[2][3], sxi[2], sci[2], di[2], acc[2];
  int row[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int i = blockIdx.x * LAP_ROWS + warp * 2 + r;
    row[r] = i;
    sxi[r] = 0.f; sci[r] = 0.f; di[r] = 1.f; acc[r] = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) { xi[r][c] = 0.f; ci[r][c] = 0.f; }
    if (i < N) {
      for (int c = 0; c < D1; ++c) { xi[r][c] = Xb[(size_t)i * D1 + c]; sxi[r] = __fmaf_rn(xi[r][c], xi[r][c], sxi[r]); }
      for (int c = 0; c < D2; ++c) { ci[r][c] = Cb[(size_t)i * D2 + c]; sci[r] = __fmaf_rn(ci[r][c], ci[r][c], sci[r]); }
      if (WRITE) di[r] = deg[(size_t)b * N + i];
    }
  }
  for (int j0 = 0; j0 < N; j0 += 128) {
    __syncthreads();
    if (tid < 128) {
      const int j = j0 + tid;
      float a = 0.f, c2 = 0.f;
      if (j < N) {
        for (int c = 0; c < D1; ++c) { const float v = Xb[(size_t)j * D1 + c]; sx[tid][c] = v; a = __fmaf_rn(v, v, a); }
        for (int c = 0; c < D2; ++c) { const float v = Cb[(size_t)j * D2 + c]; sc[tid][c] = v; c2 = __fmaf_rn(v, v, c2); }
        if (WRITE) sdeg[tid] = deg[(size_t)b * N + j];
      }
      ssx[tid] = a;
      ssc[tid] = c2;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      if (row[r] >= N) continue;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int jj = lane + 32 * t, j2 = j0 + jj;
        if (j2 >= N) continue;
        const float w = expf(-sqdist_smooth(xi[r], sx[jj], sxi[r], ssx[jj], D1) * s1) *
                        expf(-sqdist_smooth(ci[r], sc[jj], sci[r], ssc[jj], D2) * s2);   // Tool.py:449,457,459
        if (WRITE) {
          const float num = ((j2 == row[r]) ? (di[r] + 1e-8f) : 0.f) - w;                 // D - W  (:462,:464)
          Lout[((size_t)b * N + row[r]) * N + j2] = num * rsqrtf(di[r]) * rsqrtf(sdeg[jj]);   // D^-1/2 . D^-1/2 (:463,:465)
        } else {
          acc[r] += w;
        }
      }
    }
  }
  if (!WRITE) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      float v = acc[r];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0 && row[r] < N) deg[(size_t)b * N + row[r]] = v;
    }
  }
}

// w = 1 - H_2(G)/log_2 K ; A = alpha L + beta diag(w) + 1e-5 I (in place over a copy of L) ; rhs = beta w G ;
// dinv = 1 / diag(A)   (ProbLabelPropagation.py:19-22,38-40)
__global__ void lp_setup_kernel(const float* __restrict__ Lm, const float* __restrict__ G, int N, int K, int Kc,
                                float alpha, float beta, float* __restrict__ A, float* __restrict__ w,
                                float* __restrict__ rhs, float* __restrict__ dinv) {
  const int n = blockIdx.x;
  __shared__ float sw;
  if (threadIdx.x == 0) {
    float h = 0.f;
    for (int c = 0; c < K; ++c) {
      const float g = G[(size_t)n * K + c];
      h += g * logf(g + 1e-5f) / logf(2.f);
    }
    const float ww = 1.f - (-h) / (logf((float)K) / logf(2.f));
    sw = ww;
    w[n] = ww;
  }
  __syncthreads();
  const float ww = sw;
  for (int j = threadIdx.x; j < N; j += blockDim.x) {
    float a = alpha * Lm[(size_t)n * N + j];
    if (j == n) {
      a += beta * ww + 1e-5f;
      dinv[n] = 1.f / a;
    }
    A[(size_t)n * N + j] = a;
  }
  for (int c = threadIdx.x; c < Kc; c += blockDim.x) rhs[(size_t)n * Kc + c] = (c < K) ? beta * ww * G[(size_t)n * K + c] : 0.f;
}

// CG state vectors are (N, Kc) row-major; scal = [rz(Kc) | pq(Kc) | rz_new(Kc) | rr(Kc)] fp64
__global__ void cg_init_kernel(const float* __restrict__ rhs, const float* __restrict__ dinv, int N, int Kc,
                               float* __restrict__ x, float* __restrict__ r, float* __restrict__ z, float* __restrict__ p,
                               double* __restrict__ scal) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= N * Kc) return;
  const int n = t / Kc, c = t % Kc;
  const float rv = rhs[t], zv = rv * dinv[n];
  x[t] = 0.f; r[t] = rv; z[t] = zv; p[t] = zv;
  atomicAdd(&scal[c], (double)rv * (double)zv);
}
__global__ void cg_dot_kernel(const float* __restrict__ p, const float* __restrict__ q, int N, int Kc,
                              double* __restrict__ scal) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= N * Kc) return;
  atomicAdd(&scal[Kc + t % Kc], (double)p[t] * (double)q[t]);
}
__global__ void cg_update_kernel(const float* __restrict__ q, const float* __restrict__ dinv, int N, int Kc,
                                 float* __restrict__ x, float* __restrict__ r, float* __restrict__ z,
                                 const float* __restrict__ p, double* __restrict__ scal) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= N * Kc) return;
  const int n = t / Kc, c = t % Kc;
  const double pq = scal[Kc + c];
  const float alpha = (pq != 0.0) ? (float)(scal[c] / pq) : 0.f;
  x[t] += alpha * p[t];
  const float rv = r[t] - alpha * q[t];
  const float zv = rv * dinv[n];
  r[t] = rv;
  z[t] = zv;
  atomicAdd(&scal[2 * Kc + c], (double)rv * (double)zv);
  atomicAdd(&scal[3 * Kc + c], (double)rv * (double)rv);
}
// Same update for Kc dividing the block size (N * Kc a multiple of 256): the two dot products are reduced inside the block
// first -- 2 Kc fp64 atomics per block instead of 2 per element (4096 per address and iteration at N = 4096).
__global__ void __launch_bounds__(256)
cg_update_blockred_kernel(const float* __restrict__ q, const float* __restrict__ dinv, int Kc, float* __restrict__ x,
                          float* __restrict__ r, float* __restrict__ z, const float* __restrict__ p,
                          double* __restrict__ scal) {
  __shared__ double s_rz[256], s_rr[256];
  const int tid = threadIdx.x;
  const int t = blockIdx.x * 256 + tid;
  const int n = t / Kc, c = t % Kc;
  const double pq = scal[Kc + c];
  const float alpha = (pq != 0.0) ? (float)(scal[c] / pq) : 0.f;
  x[t] += alpha * p[t];
  const float rv = r[t] - alpha * q[t];
  const float zv = rv * dinv[n];
  r[t] = rv;
  z[t] = zv;
  s_rz[tid] = (double)rv * (double)zv;
  s_rr[tid] = (double)rv * (double)rv;
  __syncthreads();
  if (tid < Kc) {            // 256 % Kc == 0: entries tid, tid + Kc, ... share the column
    double a = 0.0, b = 0.0;
    for (int i = tid; i < 256; i += Kc) { a += s_rz[i]; b += s_rr[i]; }
    atomicAdd(&scal[2 * Kc + tid], a);
    atomicAdd(&scal[3 * Kc + tid], b);
  }
}
__global__ void cg_dir_kernel(const float* __restrict__ z, int N, int Kc, float* __restrict__ p, double* __restrict__ scal,
                              double* __restrict__ resid_out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < N * Kc) {
    const int c = t % Kc;
    const double rz = scal[c];
    const float beta = (rz != 0.0) ? (float)(scal[2 * Kc + c] / rz) : 0.f;
    p[t] = z[t] + beta * p[t];
  }
}
// roll the scalars after every thread of cg_dir has read them (separate launch => ordering by the stream)
__global__ void cg_roll_kernel(int Kc, double* __restrict__ scal, double* __restrict__ resid_out) {
  const int c = threadIdx.x;
  if (c < Kc) {
    scal[c] = scal[2 * Kc + c];
    if (resid_out) resid_out[c] = scal[3 * Kc + c];
    scal[Kc + c] = 0.0;
    scal[2 * Kc + c] = 0.0;
    scal[3 * Kc + c] = 0.0;
  }
}
// q = A p for the CG loop, A (N, N) row-major, p / q (N, 16) row-major, plus the dot products pq[c] = sum_n p[n,c] q[n,c]
// (fp64 atomics into scal_pq[0:16]).  A CTA owns 32 rows (4 per warp); 256 columns of A and the matching 256 rows of p are
// staged in shared memory while the next chunk's global loads are already in flight in registers.  A lane accumulates
// 4 rows x 16 columns over the columns j = lane (mod 32) of every chunk; a butterfly sums the lanes at the end.
// (The tcgen05 row GEMM runs this shape -- 4096 x 4096 x 16 -- on 32 CTAs with a serial chunk pipeline: 213 us; this
// kernel: every SM busy, A streamed once.)
constexpr int MV_ROWS = 32, MV_JC = 256, MV_PLD = 20;
constexpr size_t MV_SMEM = sizeof(float) * (MV_ROWS * MV_JC + MV_JC * MV_PLD + 8 * 16);
__global__ void __launch_bounds__(256, 1)
lp_matvec16_kernel(const float* __restrict__ A, const float* __restrict__ p, int N, float* __restrict__ q,
                   double* __restrict__ scal_pq) {
  extern __shared__ __align__(16) unsigned char mv_smem[];
  float (*sA)[MV_JC] = reinterpret_cast<float (*)[MV_JC]>(mv_smem);
  float (*sP)[MV_PLD] = reinterpret_cast<float (*)[MV_PLD]>(mv_smem + sizeof(float) * MV_ROWS * MV_JC);
  float (*sdot)[16] = reinterpret_cast<float (*)[16]>(mv_smem + sizeof(float) * (MV_ROWS * MV_JC + MV_JC * MV_PLD));
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n0 = blockIdx.x * MV_ROWS;
  float acc[4][16];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[r][c] = 0.f;
  float4 ra[8], rp[4];
  auto fetch = [&](int j0) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int f = tid + 256 * u, r = f >> 6, c4 = f & 63;
      const int n = n0 + r, j = j0 + c4 * 4;
      ra[u] = (n < N && j < N) ? *reinterpret_cast<const float4*>(A + (size_t)n * N + j) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int f = tid + 256 * u, jr = f >> 2, q4 = f & 3;
      const int j = j0 + jr;
      rp[u] = (j < N) ? *reinterpret_cast<const float4*>(p + (size_t)j * 16 + q4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  fetch(0);
  for (int j0 = 0; j0 < N; j0 += MV_JC) {
    __syncthreads();                                   // the previous chunk has been consumed
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int f = tid + 256 * u;
      *reinterpret_cast<float4*>(&sA[f >> 6][(f & 63) * 4]) = ra[u];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int f = tid + 256 * u;
      *reinterpret_cast<float4*>(&sP[f >> 2][(f & 3) * 4]) = rp[u];
    }
    __syncthreads();
    if (j0 + MV_JC < N) fetch(j0 + MV_JC);             // in flight during the FMAs below
#pragma unroll 2
    for (int i = 0; i < MV_JC / 32; ++i) {
      const int j = 32 * i + lane;
      float pv[16];
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        const float4 v = *reinterpret_cast<const float4*>(&sP[j][c4 * 4]);
        pv[c4 * 4] = v.x; pv[c4 * 4 + 1] = v.y; pv[c4 * 4 + 2] = v.z; pv[c4 * 4 + 3] = v.w;
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float a = sA[warp * 4 + r][j];
#pragma unroll
        for (int c = 0; c < 16; ++c) acc[r][c] = fmaf(a, pv[c], acc[r][c]);
      }
    }
  }
  // lanes -> one value per (row, column): fixed-order butterfly, then lane c keeps column c
  float dotc = 0.f;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    float mine = 0.f;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      float v = acc[r][c];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == c) mine = v;
    }
    const int n = n0 + warp * 4 + r;
    if (lane < 16 && n < N) {
      q[(size_t)n * 16 + lane] = mine;
      dotc = fmaf(mine, p[(size_t)n * 16 + lane], dotc);
    }
  }
  if (lane < 16) sdot[warp][lane] = dotc;
  __syncthreads();
  if (tid < 16) {
    double d = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) d += (double)sdot[w][tid];
    atomicAdd(scal_pq + tid, d);
  }
}

__global__ void lp_finish_kernel(const float* __restrict__ x, int N, int K, int Kc, float* __restrict__ Y,
                                 float* __restrict__ Yp) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float s = 0.f;
  for (int c = 0; c < K; ++c) s += x[(size_t)n * Kc + c];
  for (int c = 0; c < K; ++c) {
    const float v = x[(size_t)n * Kc + c];
    Y[(size_t)n * K + c] = v;
    Yp[(size_t)n * K + c] = v / s;          // ProbLabelPropagation.py:23
  }
}

// out[b,n,r,:] = X[b, idx[b,n,r], :]  (edge = 0)   or   [X[b,n,:] | X[b,idx,:] - X[b,n,:]]  (edge = 1)
__global__ void gather_kernel(const float* __restrict__ X, const int32_t* __restrict__ idx, long long R, int N, int k,
                              int C, long long ldx, int edge, float* __restrict__ out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int OC = edge ? 2 * C : C;
  if (t >= R * OC) return;
  const long long row = t / OC;
  const int c = (int)(t % OC);
  const long long pt = row / k;
  const long long nb = (pt / N) * N + idx[row];
  float v;
  if (!edge) v = X[nb * ldx + c];
  else v = (c < C) ? X[pt * ldx + c] : X[nb * ldx + c - C] - X[pt * ldx + c - C];
  out[t] = v;
}

__global__ void bn_apply_kernel(const float* __restrict__ y, const float* __restrict__ sc, const float* __restrict__ sh,
                                long long rows, int C, int relu, float* __restrict__ out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rows * C) return;
  const int c = (int)(t % C);
  const float v = fmaf(y[t], sc[c], sh[c]);
  out[t] = relu ? fmaxf(v, 0.f) : v;
}

}  // namespace
}  // namespace wspc

using namespace wspc;

extern "C" int wspc_laplacian_sym(const float* X, const float* RGB, int B, int N, int D1, int D2, float scale_xyz,
                                  float scale_rgb, float* deg_ws, float* Lout, wspc_stream_t stream) {
  if (int rc = check_arch()) return rc;
  WSPC_REQUIRE(X && RGB && deg_ws && Lout, "laplacian_sym: null pointer");
  WSPC_REQUIRE(B >= 1 && N >= 1 && D1 >= 1 && D1 <= 3 && D2 >= 1 && D2 <= 3, "laplacian_sym: bad shape (D <= 3)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid((N + LAP_ROWS - 1) / LAP_ROWS, B);
  laplacian_kernel<false><<<grid, 256, 0, st>>>(X, RGB, N, D1, D2, scale_xyz, scale_rgb, deg_ws, nullptr);
  laplacian_kernel<true><<<grid, 256, 0, st>>>(X, RGB, N, D1, D2, scale_xyz, scale_rgb, deg_ws, Lout);
  count_launch(2);
  WSPC_LAUNCH_CHECK("laplacian_kernel");
  return WSPC_OK;
}

extern "C" size_t wspc_lp_solve_workspace_bytes(int N, int K) {
  const int Kc = ((K + 3) / 4 * 4) < 16 ? 16 : (K + 3) / 4 * 4;
  return align_up((size_t)N * N * 4, 256) + 6 * align_up((size_t)N * Kc * 4, 256) + align_up((size_t)N * 4, 256) + 4096;
}

extern "C" int wspc_lp_solve(const float* Lm, const float* G, int N, int K, float alpha, float beta, int max_iter,
                             float tol, float* Y, float* Yprob, float* w, int* iters_out, void* workspace,
                             size_t workspace_bytes, wspc_stream_t stream) {
  if (int rc = check_arch()) return rc;
  WSPC_REQUIRE(Lm && G && Y && Yprob && w && workspace, "lp_solve: null pointer");
  WSPC_REQUIRE(N >= 1 && K >= 2 && K <= 64 && (N % 8) == 0, "lp_solve: need 2 <= K <= 64 and N %% 8 == 0");
  if (workspace_bytes < wspc_lp_solve_workspace_bytes(N, K)) {
    set_error("lp_solve: workspace too small");
    return WSPC_ERR_WORKSPACE;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int Kc = ((K + 3) / 4 * 4) < 16 ? 16 : (K + 3) / 4 * 4;
  char* wsp = static_cast<char*>(workspace);
  const size_t vb = align_up((size_t)N * Kc * 4, 256);
  float* A = reinterpret_cast<float*>(wsp); wsp += align_up((size_t)N * N * 4, 256);
  float* rhs = reinterpret_cast<float*>(wsp); wsp += vb;
  float* x = reinterpret_cast<float*>(wsp); wsp += vb;
  float* r = reinterpret_cast<float*>(wsp); wsp += vb;
  float* z = reinterpret_cast<float*>(wsp); wsp += vb;
  float* p = reinterpret_cast<float*>(wsp); wsp += vb;
  float* q = reinterpret_cast<float*>(wsp); wsp += vb;
  float* dinv = reinterpret_cast<float*>(wsp); wsp += align_up((size_t)N * 4, 256);
  double* scal = reinterpret_cast<double*>(wsp);           // 4*Kc doubles (Kc <= 64 -> 2 KB) + residuals (512 B)
  double* resid = scal + 4 * 64;
  static const cudaError_t mv_attr =
      cudaFuncSetAttribute(lp_matvec16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MV_SMEM);
  WSPC_CUDA(mv_attr);
  lp_setup_kernel<<<N, 128, 0, st>>>(Lm, G, N, K, Kc, alpha, beta, A, w, rhs, dinv);
  WSPC_CUDA(cudaMemsetAsync(scal, 0, 4096, st));
  const int nt = N * Kc;
  cg_init_kernel<<<(nt + 255) / 256, 256, 0, st>>>(rhs, dinv, N, Kc, x, r, z, p, scal);
  count_launch(2);
  wspc_operand_t Aop;
  memset(&Aop, 0, sizeof(Aop));
  Aop.p = A; Aop.ld = N; Aop.C = N;
  wspc_epilogue_t ep;
  memset(&ep, 0, sizeof(ep));
  ep.out = q; ep.ldo = Kc; ep.rb_rows = 1;
  double h_res[64], h_b2[64];
  int it = 0;
  bool have_b2 = false;
  const int check_every = 50;
  for (; it < max_iter; ++it) {
    if (Kc == 16) {   // q = A p and the p.q dots in one pass over A
      lp_matvec16_kernel<<<(N + MV_ROWS - 1) / MV_ROWS, 256, MV_SMEM, st>>>(A, p, N, q, scal + Kc);
    } else {
      if (int rc = wspc_conv1x1_rows(&Aop, WSPC_OP_PLAIN, p, Kc, 0, N, Kc, N, &ep, WSPC_EPI_STORE, stream)) return rc;
      cg_dot_kernel<<<(nt + 255) / 256, 256, 0, st>>>(p, q, N, Kc, scal);
    }
    if (256 % Kc == 0 && nt % 256 == 0) cg_update_blockred_kernel<<<nt / 256, 256, 0, st>>>(q, dinv, Kc, x, r, z, p, scal);
    else cg_update_kernel<<<(nt + 255) / 256, 256, 0, st>>>(q, dinv, N, Kc, x, r, z, p, scal);
    cg_dir_kernel<<<(nt + 255) / 256, 256, 0, st>>>(z, N, Kc, p, scal, nullptr);
    cg_roll_kernel<<<1, 64, 0, st>>>(Kc, scal, resid);
    count_launch(4);
    if ((it + 1) % check_every == 0 || (it + 1 <= check_every && (it + 1) % 10 == 0) || it + 1 == max_iter) {
      if (!have_b2) {   // ||b||^2 per column, once (host-side convergence control only)
        std::vector<float> hb((size_t)N * Kc);
        WSPC_CUDA(cudaMemcpyAsync(hb.data(), rhs, hb.size() * 4, cudaMemcpyDeviceToHost, st));
        WSPC_CUDA(cudaStreamSynchronize(st));
        for (int c = 0; c < Kc; ++c) h_b2[c] = 0.0;
        for (size_t i = 0; i < hb.size(); ++i) h_b2[i % Kc] += (double)hb[i] * hb[i];
        have_b2 = true;
      }
      WSPC_CUDA(cudaMemcpyAsync(h_res, resid, Kc * 8, cudaMemcpyDeviceToHost, st));
      WSPC_CUDA(cudaStreamSynchronize(st));
      bool done = true;
      for (int c = 0; c < K; ++c)
        if (h_res[c] > (double)tol * tol * (h_b2[c] > 0 ? h_b2[c] : 1.0)) done = false;
      if (done) { ++it; break; }
    }
  }
  lp_finish_kernel<<<(N + 127) / 128, 128, 0, st>>>(x, N, K, Kc, Y, Yprob);
  count_launch();
  WSPC_LAUNCH_CHECK("lp kernels");
  if (iters_out) *iters_out = it;
  return WSPC_OK;
}

extern "C" int wspc_gather(const float* X, const int32_t* idx, int B, int N, int k, int C, long long ldx, int edge,
                           float* out, wspc_stream_t stream) {
  if (int rc = check_arch()) return rc;
  WSPC_REQUIRE(X && idx && out && B >= 1 && N >= 1 && k >= 1 && C >= 1, "gather: bad argument");
  const long long R = (long long)B * N * k;
  const long long total = R * (edge ? 2 * C : C);
  gather_kernel<<<(unsigned)((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(X, idx, R, N, k, C,
                                                                                                     ldx, edge, out);
  count_launch();
  WSPC_LAUNCH_CHECK("gather_kernel");
  return WSPC_OK;
}

extern "C" int wspc_bn_apply(const float* y, const float* sc, const float* sh, long long rows, int C, int relu, float* out,
                             wspc_stream_t stream) {
  if (int rc = check_arch()) return rc;
  WSPC_REQUIRE(y && sc && sh && out && rows >= 1 && C >= 1, "bn_apply: bad argument");
  const long long total = rows * C;
  bn_apply_kernel<<<(unsigned)((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(y, sc, sh, rows, C,
                                                                                                      relu, out);
  count_launch();
  WSPC_LAUNCH_CHECK("bn_apply_kernel");
  return WSPC_OK;
}
