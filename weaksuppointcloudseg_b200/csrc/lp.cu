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

// pass 1: degree d_i = sum_j W_ij.   grid (N/128, B), block 128: one thread per row i, columns staged in smem
template <bool WRITE>
__global__ void __launch_bounds__(128)
laplacian_kernel(const float* __restrict__ X, const float* __restrict__ RGB, int N, int D1, int D2, float s1, float s2,
                 float* __restrict__ deg, float* __restrict__ Lout) {
  __shared__ float sx[128][3], sc[128][3], ssx[128], ssc[128], sdeg[128];
  const int b = blockIdx.y;
  const int i = blockIdx.x * 128 + threadIdx.x;
  const float* Xb = X + (size_t)b * N * D1;
  const float* Cb = RGB + (size_t)b * N * D2;
  float xi[3] = {0.f, 0.f, 0.f}, ci[3] = {0.f, 0.f, 0.f};
  float sxi = 0.f, sci = 0.f, di = 0.f;
  if (i < N) {
    for (int c = 0; c < D1; ++c) { xi[c] = Xb[(size_t)i * D1 + c]; sxi = __fmaf_rn(xi[c], xi[c], sxi); }
    for (int c = 0; c < D2; ++c) { ci[c] = Cb[(size_t)i * D2 + c]; sci = __fmaf_rn(ci[c], ci[c], sci); }
    if (WRITE) di = deg[(size_t)b * N + i];
  }
  float acc = 0.f;
  for (int j0 = 0; j0 < N; j0 += 128) {
    const int j = j0 + threadIdx.x;
    __syncthreads();
    if (j < N) {
      float a = 0.f, c2 = 0.f;
      for (int c = 0; c < D1; ++c) { const float v = Xb[(size_t)j * D1 + c]; sx[threadIdx.x][c] = v; a = __fmaf_rn(v, v, a); }
      for (int c = 0; c < D2; ++c) { const float v = Cb[(size_t)j * D2 + c]; sc[threadIdx.x][c] = v; c2 = __fmaf_rn(v, v, c2); }
      ssx[threadIdx.x] = a;
      ssc[threadIdx.x] = c2;
      if (WRITE) sdeg[threadIdx.x] = deg[(size_t)b * N + j];
    }
    __syncthreads();
    if (i < N) {
      const int jn = (N - j0 < 128) ? N - j0 : 128;
      for (int jj = 0; jj < jn; ++jj) {
        const float w = expf(-sqdist_smooth(xi, sx[jj], sxi, ssx[jj], D1) * s1) *
                        expf(-sqdist_smooth(ci, sc[jj], sci, ssc[jj], D2) * s2);     // Tool.py:449,457,459
        if (WRITE) {
          const int j2 = j0 + jj;
          const float num = ((j2 == i) ? (di + 1e-8f) : 0.f) - w;                     // D - W  (:462,:464)
          Lout[((size_t)b * N + i) * N + j2] = num * rsqrtf(di) * rsqrtf(sdeg[jj]);   // D^-1/2 . D^-1/2 (:463,:465)
        } else {
          acc += w;
        }
      }
    }
  }
  if (!WRITE && i < N) deg[(size_t)b * N + i] = acc;
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
  dim3 grid((N + 127) / 128, B);
  laplacian_kernel<false><<<grid, 128, 0, st>>>(X, RGB, N, D1, D2, scale_xyz, scale_rgb, deg_ws, nullptr);
  laplacian_kernel<true><<<grid, 128, 0, st>>>(X, RGB, N, D1, D2, scale_xyz, scale_rgb, deg_ws, Lout);
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
    if (int rc = wspc_conv1x1_rows(&Aop, WSPC_OP_PLAIN, p, Kc, 0, N, Kc, N, &ep, WSPC_EPI_STORE, stream)) return rc;   // q = A p
    cg_dot_kernel<<<(nt + 255) / 256, 256, 0, st>>>(p, q, N, Kc, scal);
    cg_update_kernel<<<(nt + 255) / 256, 256, 0, st>>>(q, dinv, N, Kc, x, r, z, p, scal);
    cg_dir_kernel<<<(nt + 255) / 256, 256, 0, st>>>(z, N, Kc, p, scal, nullptr);
    cg_roll_kernel<<<1, 64, 0, st>>>(Kc, scal, resid);
    count_launch(4);
    if ((it + 1) % check_every == 0 || it + 1 == max_iter) {
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
