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

// channels beyond D are zero on both sides: fma(0, 0, dot) == dot exactly, so the chain is the D-term chain of the reference
// formula while the loop has a compile-time trip count (a run-time D put the row registers into local memory)
__device__ __forceinline__ float sqdist_smooth(const float (&a)[3], const float (&b)[3], float sa, float sb) {
  float dot = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) dot = __fmaf_rn(a[c], b[c], dot);
  const float d = __fsub_rn(__fadd_rn(sa, sb), __fmul_rn(2.f, dot));
  return d > 0.f ? d : 0.f;
}

// Symmetric normalised Laplacian in two passes over the N x N pair kernel (it is never stored unnormalised):
// pass 1 (WRITE = false): degree d_i = sum_j W_ij;  pass 2: L_ij = ((i == j) (d_i + 1e-8) - W_ij) d_i^-1/2 d_j^-1/2.
// grid (N/32, B), block 256: a CTA owns 32 rows (4 per warp); 512 columns at a time are staged in shared memory and a lane
// evaluates columns lane, lane+32, ... of the tile for its warp's four rows -- the column's coordinates are read from shared
// memory once per four pair weights, there are two block barriers per 512 columns (round 2 start: per 128, 2 rows per warp:
// 0.22 ms per 4096-point block), and pass 2 stores 128 contiguous bytes per warp and row.  A lane visits its columns in
// increasing order and the row sums are reduced by a fixed-order butterfly, so the result does not depend on scheduling.
constexpr int LAP_ROWS = 32;
constexpr int LAP_RPW = 4;        // rows per warp
constexpr int LAP_COLS = 512;     // columns staged per round
template <bool WRITE>
__global__ void __launch_bounds__(256)
laplacian_kernel(const float* __restrict__ X, const float* __restrict__ RGB, int N, int D1, int D2, float s1, float s2,
                 float* __restrict__ deg, float* __restrict__ Lout) {
  __shared__ float sx[LAP_COLS][3], sc[LAP_COLS][3], ssx[LAP_COLS], ssc[LAP_COLS], sdeg[LAP_COLS];
  const int b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* Xb = X + (size_t)b * N * D1;
  const float* Cb = RGB + (size_t)b * N * D2;
  float xi[LAP_RPW][3], ci[LAP_RPW][3], sxi[LAP_RPW], sci[LAP_RPW], di[LAP_RPW], acc[LAP_RPW];
  int row[LAP_RPW];
#pragma unroll
  for (int r = 0; r < LAP_RPW; ++r) {
    const int i = blockIdx.x * LAP_ROWS + warp * LAP_RPW + r;
    row[r] = i;
    sxi[r] = 0.f; sci[r] = 0.f; di[r] = 1.f; acc[r] = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) { xi[r][c] = 0.f; ci[r][c] = 0.f; }
    if (i < N) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        if (c < D1) { xi[r][c] = Xb[(size_t)i * D1 + c]; sxi[r] = __fmaf_rn(xi[r][c], xi[r][c], sxi[r]); }
        if (c < D2) { ci[r][c] = Cb[(size_t)i * D2 + c]; sci[r] = __fmaf_rn(ci[r][c], ci[r][c], sci[r]); }
      }
      if (WRITE) di[r] = deg[(size_t)b * N + i];
    }
  }
  float rdi[LAP_RPW];
#pragma unroll
  for (int r = 0; r < LAP_RPW; ++r) rdi[r] = rsqrtf(di[r]);
  for (int j0 = 0; j0 < N; j0 += LAP_COLS) {
    __syncthreads();
    for (int e = tid; e < LAP_COLS; e += 256) {
      const int j = j0 + e;
      float a = 0.f, c2 = 0.f;
      if (j < N) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float v = c < D1 ? Xb[(size_t)j * D1 + c] : 0.f;
          sx[e][c] = v;
          if (c < D1) a = __fmaf_rn(v, v, a);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float v = c < D2 ? Cb[(size_t)j * D2 + c] : 0.f;
          sc[e][c] = v;
          if (c < D2) c2 = __fmaf_rn(v, v, c2);
        }
        if (WRITE) sdeg[e] = rsqrtf(deg[(size_t)b * N + j]);
      }
      ssx[e] = a;
      ssc[e] = c2;
    }
    __syncthreads();
#pragma unroll 4
    for (int t = 0; t < LAP_COLS / 32; ++t) {
      const int jj = lane + 32 * t, j2 = j0 + jj;
      if (j2 >= N) break;
      const float xj[3] = {sx[jj][0], sx[jj][1], sx[jj][2]}, cj[3] = {sc[jj][0], sc[jj][1], sc[jj][2]};
      const float sxj = ssx[jj], scj = ssc[jj];
      const float rdj = WRITE ? sdeg[jj] : 0.f;
#pragma unroll
      for (int r = 0; r < LAP_RPW; ++r) {
        if (row[r] >= N) continue;
        const float w = expf(-sqdist_smooth(xi[r], xj, sxi[r], sxj) * s1) *
                        expf(-sqdist_smooth(ci[r], cj, sci[r], scj) * s2);             // Tool.py:449,457,459
        if (WRITE) {
          const float num = ((j2 == row[r]) ? (di[r] + 1e-8f) : 0.f) - w;                 // D - W  (:462,:464)
          Lout[((size_t)b * N + row[r]) * N + j2] = num * rdi[r] * rdj;                   // D^-1/2 . D^-1/2 (:463,:465)
        } else {
          acc[r] += w;
        }
      }
    }
  }
  if (!WRITE) {
#pragma unroll
    for (int r = 0; r < LAP_RPW; ++r) {
      float v = acc[r];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0 && row[r] < N) deg[(size_t)b * N + row[r]] = v;
    }
  }
}

// mat-vec tiling: a CTA owns 32 rows (4 per warp); 256 columns of L and the matching 256 rows of p are staged in shared
// memory while the next chunk's global loads are already in flight in registers.  A lane accumulates 4 rows x 16 columns over
// the columns j = lane (mod 32) of every chunk; a butterfly sums the lanes at the end.
constexpr int MV_ROWS = 32, MV_JC = 256, MV_PLD = 20;
constexpr size_t MV_SMEM = sizeof(float) * (MV_ROWS * MV_JC + MV_JC * MV_PLD + 8 * 16);

// ================================================= batched label propagation: every block of a room in flight =========
// One set of launches solves B independent systems (alpha L_b + diag(d_b)) Y_b = rhs_b; A is never formed (q = alpha L p + d p).
// Scalars live in per-iteration slots  scal[b][it][{rz, pq, rr}][64]  (zeroed once), so no kernel has to roll or re-zero them:
//   matvec(it)  pq[it]  += p.q
//   update(it)  alpha = rz[it] / pq[it];  x += alpha p;  r -= alpha q;  z = dinv r;   rz[it+1] += r.z;  rr[it] += r.r
//   dir(it)     converged (rr[it] <= tol^2 |b|^2 for every class) -> done[b] = 1, iters[b] = it + 1;  else p = z + (rz[it+1]/rz[it]) p
// Every kernel returns at once for a block whose done flag is set: convergence is decided on the device, per block.
constexpr int LPB_SC = 64;                       // scalar slot pitch (Kc <= 64)
__device__ __forceinline__ double* lpb_slot(double* scal, int b, int it, int which, int iters_alloc) {
  return scal + (((size_t)b * iters_alloc + it) * 3 + which) * LPB_SC;
}

// w, diagonal d = beta w + 1e-5, dinv = 1 / (alpha L_nn + d), rhs = beta w G, x = 0, r = rhs, z = p = dinv r, rz[0], |b|^2
__global__ void __launch_bounds__(128)
lpb_setup_kernel(const float* __restrict__ Lm, const float* __restrict__ G, int N, int K, int Kc, float alpha, float beta,
                 float* __restrict__ w, float* __restrict__ diag, float* __restrict__ dinv, float* __restrict__ x,
                 float* __restrict__ r, float* __restrict__ z, float* __restrict__ p, double* __restrict__ scal,
                 double* __restrict__ b2, int iters_alloc) {
  const int b = blockIdx.y;
  const int n = blockIdx.x * 8 + (threadIdx.x >> 4), c0 = threadIdx.x & 15;   // 16 threads per point
  __shared__ double s_rz[8][LPB_SC], s_b2[8][LPB_SC];
  const int pl = threadIdx.x >> 4;
  for (int c = c0; c < Kc; c += 16) { s_rz[pl][c] = 0.0; s_b2[pl][c] = 0.0; }
  const bool valid = n < N;
  const size_t row = (size_t)b * N + (valid ? n : 0);
  const float* g = G + row * K;
  float h = 0.f;
  if (valid)
    for (int c = c0; c < K; c += 16) h += g[c] * logf(g[c] + 1e-5f) / logf(2.f);       // ProbLabelPropagation.py:38-39
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o, 16);
  if (valid) {
    const float ww = 1.f - (-h) / (logf((float)K) / logf(2.f));                           // :40
    const float d = beta * ww + 1e-5f;                                                      // :21-22
    const float di = 1.f / (alpha * Lm[row * N + n] + d);
    if (c0 == 0) { w[row] = ww; diag[row] = d; dinv[row] = di; }
    for (int c = c0; c < Kc; c += 16) {
      const float rv = (c < K) ? beta * ww * g[c] : 0.f;
      const float zv = rv * di;
      const size_t t = row * Kc + c;
      x[t] = 0.f; r[t] = rv; z[t] = zv; p[t] = zv;
      s_rz[pl][c] = (double)rv * (double)zv;
      s_b2[pl][c] = (double)rv * (double)rv;
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < Kc; c += 128) {
    double a = 0.0, e = 0.0;
    for (int i = 0; i < 8; ++i) { a += s_rz[i][c]; e += s_b2[i][c]; }
    atomicAdd(lpb_slot(scal, b, 0, 0, iters_alloc) + c, a);
    atomicAdd(b2 + (size_t)b * LPB_SC + c, e);
  }
}

// q = alpha L p + d p for one 16-column slab of one block (grid: row tiles x slabs x blocks), dots pq[it] += p.q
__global__ void __launch_bounds__(256, 2)
lpb_matvec_kernel(const float* __restrict__ Lm, const float* __restrict__ p, const float* __restrict__ diag, float alpha, int N,
                  int Kc, int it, int iters_alloc, float* __restrict__ q, double* __restrict__ scal,
                  const int* __restrict__ done) {
  const int b = blockIdx.z, slab = blockIdx.y;
  if (done[b]) return;
  extern __shared__ __align__(16) unsigned char mv_smem[];
  float (*sA)[MV_JC] = reinterpret_cast<float (*)[MV_JC]>(mv_smem);
  float (*sP)[MV_PLD] = reinterpret_cast<float (*)[MV_PLD]>(mv_smem + sizeof(float) * MV_ROWS * MV_JC);
  float (*sdot)[16] = reinterpret_cast<float (*)[16]>(mv_smem + sizeof(float) * (MV_ROWS * MV_JC + MV_JC * MV_PLD));
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n0 = blockIdx.x * MV_ROWS;
  const float* A = Lm + (size_t)b * N * N;
  const float* pb = p + (size_t)b * N * Kc + slab * 16;
  const bool vec = (N & 3) == 0;
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
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n < N) {
        const float* src = A + (size_t)n * N + j;
        if (vec) {
          if (j < N) v = *reinterpret_cast<const float4*>(src);
        } else {
          if (j < N) v.x = src[0];
          if (j + 1 < N) v.y = src[1];
          if (j + 2 < N) v.z = src[2];
          if (j + 3 < N) v.w = src[3];
        }
      }
      ra[u] = v;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int f = tid + 256 * u, jr = f >> 2, q4 = f & 3;
      const int j = j0 + jr;
      rp[u] = (j < N) ? *reinterpret_cast<const float4*>(pb + (size_t)j * Kc + q4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
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
      const float pn = pb[(size_t)n * Kc + lane];
      const float qv = fmaf(alpha, mine, diag[(size_t)b * N + n] * pn);
      q[((size_t)b * N + n) * Kc + slab * 16 + lane] = qv;
      dotc = fmaf(qv, pn, dotc);
    }
  }
  if (lane < 16) sdot[warp][lane] = dotc;
  __syncthreads();
  if (tid < 16) {
    double d = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) d += (double)sdot[w][tid];
    atomicAdd(lpb_slot(scal, b, it, 1, iters_alloc) + slab * 16 + tid, d);
  }
}

// grid (ceil(N*Kc/256), B); Kc in {16, 32, 64} divides the block size, so entries tid, tid + Kc, ... share a column and the
// two dot products are reduced inside the block first (2 Kc fp64 atomics per block)
__global__ void __launch_bounds__(256)
lpb_update_kernel(const float* __restrict__ q, const float* __restrict__ dinv, int N, int Kc, int it, int iters_alloc,
                  float* __restrict__ x, float* __restrict__ r, float* __restrict__ z, const float* __restrict__ p,
                  double* __restrict__ scal, const int* __restrict__ done) {
  const int b = blockIdx.y;
  if (done[b]) return;
  __shared__ double s_rz[256], s_rr[256];
  const int tid = threadIdx.x;
  const int e = blockIdx.x * 256 + tid;
  double vrz = 0.0, vrr = 0.0;
  if (e < N * Kc) {
    const int n = e / Kc, c = e - n * Kc;
    const size_t t = (size_t)b * N * Kc + e;
    const double pq = lpb_slot(scal, b, it, 1, iters_alloc)[c];
    const float a = (pq != 0.0) ? (float)(lpb_slot(scal, b, it, 0, iters_alloc)[c] / pq) : 0.f;
    x[t] += a * p[t];
    const float rv = r[t] - a * q[t];
    const float zv = rv * dinv[(size_t)b * N + n];
    r[t] = rv;
    z[t] = zv;
    vrz = (double)rv * (double)zv;
    vrr = (double)rv * (double)rv;
  }
  s_rz[tid] = vrz;
  s_rr[tid] = vrr;
  __syncthreads();
  if (tid < Kc) {
    double a = 0.0, c2 = 0.0;
    for (int i = tid; i < 256; i += Kc) { a += s_rz[i]; c2 += s_rr[i]; }
    atomicAdd(lpb_slot(scal, b, it + 1, 0, iters_alloc) + tid, a);
    atomicAdd(lpb_slot(scal, b, it, 2, iters_alloc) + tid, c2);
  }
}

__device__ __forceinline__ bool lpb_converged(const double* rr, const double* b2, int K, float tol, float* worst) {
  bool ok = true;
  float wr = 0.f;
  for (int c = 0; c < K; ++c) {
    const double ref = b2[c] > 0.0 ? b2[c] : 1.0;
    if (rr[c] > (double)tol * tol * ref) ok = false;
    wr = fmaxf(wr, (float)sqrt(rr[c] / ref));
  }
  *worst = wr;
  return ok;
}

__global__ void __launch_bounds__(256)
lpb_dir_kernel(const float* __restrict__ z, int N, int K, int Kc, int it, int iters_alloc, float tol, float* __restrict__ p,
               const double* __restrict__ scal_c, const double* __restrict__ b2, const int* __restrict__ done) {
  const int b = blockIdx.y;
  if (done[b]) return;             // (set by an earlier launch; this launch's own decision is taken below by every CTA alike)
  double* scal = const_cast<double*>(scal_c);
  float worst;
  const bool conv = lpb_converged(lpb_slot(scal, b, it, 2, iters_alloc), b2 + (size_t)b * LPB_SC, K, tol, &worst);
  if (conv) return;                // every CTA of the block reads the same completed sums -> the same decision; the flag is
                                   // written by lpb_flag_kernel, a separate launch
  const int e = blockIdx.x * 256 + threadIdx.x;
  if (e < N * Kc) {
    const int c = e % Kc;
    const size_t t = (size_t)b * N * Kc + e;
    const double rz = lpb_slot(scal, b, it, 0, iters_alloc)[c];
    const float bt = (rz != 0.0) ? (float)(lpb_slot(scal, b, it + 1, 0, iters_alloc)[c] / rz) : 0.f;
    p[t] = z[t] + bt * p[t];
  }
}
// one thread per block: records convergence after dir(it) (a separate launch, so no CTA of dir(it) can see a flag that another
// CTA of the same launch has just set and skip its part of p)
__global__ void lpb_flag_kernel(int B, int K, int it, int iters_alloc, float tol, const double* __restrict__ scal_c,
                                const double* __restrict__ b2, int* __restrict__ done, int* __restrict__ iters,
                                float* __restrict__ resid, int* __restrict__ n_done, int last) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B || done[b]) return;
  double* scal = const_cast<double*>(scal_c);
  float worst;
  const bool conv = lpb_converged(lpb_slot(scal, b, it, 2, iters_alloc), b2 + (size_t)b * LPB_SC, K, tol, &worst);
  if (conv || last) {
    iters[b] = it + 1;
    resid[b] = worst;
    if (conv) {
      done[b] = 1;
      atomicAdd(n_done, 1);
    }
  }
}

__global__ void lpb_finish_kernel(const float* __restrict__ x, long long rows, int K, int Kc, float* __restrict__ Y,
                                  float* __restrict__ Yp) {
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= rows) return;
  float s = 0.f;
  for (int c = 0; c < K; ++c) s += x[n * Kc + c];
  for (int c = 0; c < K; ++c) {
    const float v = x[n * Kc + c];
    Y[n * K + c] = v;
    Yp[n * K + c] = v / s;          // ProbLabelPropagation.py:23
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

static int lpb_kc(int K) { return K <= 16 ? 16 : (K <= 32 ? 32 : 64); }   // divides the 256-thread blocks

extern "C" size_t wspc_lp_blocks_workspace_bytes(int B, int N, int K, int max_iter) {
  const int Kc = lpb_kc(K);
  const size_t vec = align_up((size_t)B * N * Kc * 4, 256);
  return 5 * vec + 2 * align_up((size_t)B * N * 4, 256) + align_up((size_t)B * (max_iter + 1) * 3 * LPB_SC * 8, 256) +
         align_up((size_t)B * LPB_SC * 8, 256) + align_up((size_t)B * 4, 256) + 256;
}

// All B systems advance together; nothing here blocks the host.  Iterations are enqueued in chunks of LPB_CHUNK; after each
// chunk the number of converged blocks is copied to pinned host memory, and before enqueuing a later chunk the host LOOKS at
// the copies that have already landed (cudaEventQuery, never a wait): once every block has converged it stops enqueuing.
// A host that runs ahead of the device simply enqueues up to max_iter iterations, which return at once for finished blocks.
extern "C" int wspc_lp_blocks(const float* Lm, const float* G, int B, int N, int K, float alpha, float beta, int max_iter,
                              float tol, float* Y, float* Yprob, float* w, int32_t* iters, float* resid, int32_t* done,
                              void* workspace, size_t workspace_bytes, wspc_stream_t stream) {
  if (int rc = check_arch()) return rc;
  WSPC_REQUIRE(Lm && G && Y && Yprob && w && iters && resid && done && workspace, "lp_blocks: null pointer");
  WSPC_REQUIRE(B >= 1 && B <= 65535 && N >= 1 && K >= 2 && K <= 64 && max_iter >= 1 && max_iter <= 100000,
               "lp_blocks: need 1 <= B <= 65535, 2 <= K <= 64, 1 <= max_iter <= 100000");
  WSPC_REQUIRE(aligned16(Lm), "lp_blocks: L must be 16-byte aligned");
  if (workspace_bytes < wspc_lp_blocks_workspace_bytes(B, N, K, max_iter)) {
    set_error("lp_blocks: workspace %zu < required %zu", workspace_bytes, wspc_lp_blocks_workspace_bytes(B, N, K, max_iter));
    return WSPC_ERR_WORKSPACE;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int Kc = lpb_kc(K), IA = max_iter + 1;
  char* wsp = static_cast<char*>(workspace);
  const size_t vb = align_up((size_t)B * N * Kc * 4, 256), nb = align_up((size_t)B * N * 4, 256);
  float* x = reinterpret_cast<float*>(wsp); wsp += vb;
  float* r = reinterpret_cast<float*>(wsp); wsp += vb;
  float* z = reinterpret_cast<float*>(wsp); wsp += vb;
  float* p = reinterpret_cast<float*>(wsp); wsp += vb;
  float* q = reinterpret_cast<float*>(wsp); wsp += vb;
  float* diag = reinterpret_cast<float*>(wsp); wsp += nb;
  float* dinv = reinterpret_cast<float*>(wsp); wsp += nb;
  double* scal = reinterpret_cast<double*>(wsp);
  const size_t scal_bytes = align_up((size_t)B * IA * 3 * LPB_SC * 8, 256);
  wsp += scal_bytes;
  double* b2 = reinterpret_cast<double*>(wsp); wsp += align_up((size_t)B * LPB_SC * 8, 256);
  int* n_done = reinterpret_cast<int*>(wsp);

  int dev = 0;
  WSPC_CUDA(cudaGetDevice(&dev));
  static bool attr_set[64] = {};
  if (dev < 64 && !attr_set[dev]) {       // the opt-in shared-memory size is a per-device attribute
    WSPC_CUDA(cudaFuncSetAttribute(lpb_matvec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MV_SMEM));
    attr_set[dev] = true;
  }
  constexpr int LPB_CHUNK = 16, LPB_SLOTS = 8;
  struct Poll { int* host; cudaEvent_t ev[LPB_SLOTS]; bool pending[LPB_SLOTS]; bool ok; };
  static thread_local Poll poll = {nullptr, {}, {}, false};
  if (!poll.host) {
    if (cudaHostAlloc(reinterpret_cast<void**>(&poll.host), LPB_SLOTS * sizeof(int), cudaHostAllocDefault) == cudaSuccess) {
      poll.ok = true;
      for (int i = 0; i < LPB_SLOTS; ++i) {
        poll.pending[i] = false;
        if (cudaEventCreateWithFlags(&poll.ev[i], cudaEventDisableTiming) != cudaSuccess) poll.ok = false;
      }
    }
    (void)cudaGetLastError();
  }
  bool can_poll = poll.ok;
  if (can_poll) {     // copies of an earlier call that are still in flight would be mistaken for this call's counters
    for (int i = 0; i < LPB_SLOTS; ++i)
      if (poll.pending[i]) {
        if (cudaEventQuery(poll.ev[i]) == cudaSuccess) poll.pending[i] = false;
        else { (void)cudaGetLastError(); can_poll = false; }
      }
    if (can_poll) for (int i = 0; i < LPB_SLOTS; ++i) poll.host[i] = 0;
  }

  WSPC_CUDA(cudaMemsetAsync(scal, 0, scal_bytes + align_up((size_t)B * LPB_SC * 8, 256) + 256, st));   // scal, b2, n_done
  WSPC_CUDA(cudaMemsetAsync(done, 0, (size_t)B * 4, st));
  WSPC_CUDA(cudaMemsetAsync(iters, 0, (size_t)B * 4, st));
  lpb_setup_kernel<<<dim3((N + 7) / 8, B), 128, 0, st>>>(Lm, G, N, K, Kc, alpha, beta, w, diag, dinv, x, r, z, p, scal, b2, IA);
  count_launch();
  const dim3 gmv((N + MV_ROWS - 1) / MV_ROWS, Kc / 16, B), gvec((N * Kc + 255) / 256, B);
  int chunk = 0;
  bool stop = false;
  for (int it = 0; it < max_iter && !stop; ++it) {
    const int last = (it + 1 == max_iter) ? 1 : 0;
    lpb_matvec_kernel<<<gmv, 256, MV_SMEM, st>>>(Lm, p, diag, alpha, N, Kc, it, IA, q, scal, done);
    lpb_update_kernel<<<gvec, 256, 0, st>>>(q, dinv, N, Kc, it, IA, x, r, z, p, scal, done);
    lpb_dir_kernel<<<gvec, 256, 0, st>>>(z, N, K, Kc, it, IA, tol, p, scal, b2, done);
    lpb_flag_kernel<<<(B + 127) / 128, 128, 0, st>>>(B, K, it, IA, tol, scal, b2, done, iters, resid, n_done, last);
    count_launch(4);
    if (!can_poll || last || (it + 1) % LPB_CHUNK != 0) continue;
    const int slot = chunk++ % LPB_SLOTS;
    if (poll.pending[slot]) {               // the host is a full ring ahead of the device: skip this sample
      if (cudaEventQuery(poll.ev[slot]) != cudaSuccess) { (void)cudaGetLastError(); continue; }
      poll.pending[slot] = false;
      if (poll.host[slot] >= B) { stop = true; continue; }
    }
    WSPC_CUDA(cudaMemcpyAsync(poll.host + slot, n_done, sizeof(int), cudaMemcpyDeviceToHost, st));
    WSPC_CUDA(cudaEventRecord(poll.ev[slot], st));
    poll.pending[slot] = true;
    for (int s2 = 0; s2 < LPB_SLOTS; ++s2) {   // look at whatever has landed, never wait (the counter only grows)
      if (!poll.pending[s2]) continue;
      if (cudaEventQuery(poll.ev[s2]) == cudaSuccess) {
        poll.pending[s2] = false;
        if (poll.host[s2] >= B) stop = true;
      } else {
        (void)cudaGetLastError();
      }
    }
  }
  lpb_finish_kernel<<<(unsigned)(((long long)B * N + 127) / 128), 128, 0, st>>>(x, (long long)B * N, K, Kc, Y, Yprob);
  count_launch();
  WSPC_LAUNCH_CHECK("lp_blocks kernels");
  return WSPC_OK;
}

// single system = a batch of one (kept for Util/ProbLabelPropagation.LabelPropagation_TF.SolveLabelProp callers)
extern "C" size_t wspc_lp_solve_workspace_bytes(int N, int K) {
  return wspc_lp_blocks_workspace_bytes(1, N, K, 4096) + 256;
}

extern "C" int wspc_lp_solve(const float* Lm, const float* G, int N, int K, float alpha, float beta, int max_iter,
                             float tol, float* Y, float* Yprob, float* w, int* iters_out, void* workspace,
                             size_t workspace_bytes, wspc_stream_t stream) {
  WSPC_REQUIRE(workspace, "lp_solve: null pointer");
  if (max_iter > 4096) max_iter = 4096;
  if (workspace_bytes < wspc_lp_solve_workspace_bytes(N, K)) {
    set_error("lp_solve: workspace too small");
    return WSPC_ERR_WORKSPACE;
  }
  int32_t* flags = static_cast<int32_t*>(workspace);          // [iters, done] + resid
  float* resid = reinterpret_cast<float*>(flags + 2);
  if (int rc = wspc_lp_blocks(Lm, G, 1, N, K, alpha, beta, max_iter, tol, Y, Yprob, w, flags, resid, flags + 1,
                              static_cast<char*>(workspace) + 256, workspace_bytes - 256, stream))
    return rc;
  if (iters_out) {     // the caller asked for a host-side count: one copy + wait at the very end (NULL keeps the call asynchronous)
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int32_t h = 0;
    WSPC_CUDA(cudaMemcpyAsync(&h, flags, sizeof(h), cudaMemcpyDeviceToHost, st));
    WSPC_CUDA(cudaStreamSynchronize(st));
    *iters_out = h;
  }
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
