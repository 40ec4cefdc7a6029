// Backward pass of "conv2d 1x1 -> batch norm -> ReLU -> max over the N points of a cloud" without touching the
// (P, Cout) activation (adj_conv7 + maxpool, DGCNN_S3DIS.py:80-85 / DGCNN_ShapeNet.py:80-85; tconv3 + tmaxpool,
// transform_nets.py:29-34).
//
// max_pool2d sends its gradient to one point per (cloud, channel) [TF MaxPoolGrad: first arg-max], so the gradient G
// at the batch-norm output has <= B*Cout non-zeros.  The batch-norm backward is affine, dy = c1*G + c2 + c3*y
// (wspc_bn_bwd_coeffs), and y = A W + 1 b^T is itself linear in the layer input A (P, Cin).  Therefore
//     dA = dy W^T   = A (W diag(c3) W^T) + 1 (t^T W^T)            + sparse(c1*G) W^T
//     dW = A^T dy   = (A^T A) W diag(c3) + (A^T 1) t^T            + A^T sparse(c1*G)
//     db = 1^T dy   = c3 * ((A^T 1)^T W) + P t                    + 1^T sparse(c1*G),        t = c2 + c3*b
// The dense terms are a (P, Cin) x (Cin, Cin) GEMM and the Gram matrix A^T A (both through wspc_conv1x1_rows /
// wspc_conv1x1_wgrad): Cout/Cin = 5-8x fewer flops than the (P, Cout) formulation and no read of y (2.1 GB at cfg-3).
// This file holds the small glue kernels: coefficient vectors, the sparse terms, and the final combination.
#include "common.cuh"

namespace wspc {
void count_launch(int n = 1);
namespace {

// t[c] = c2[c] + c3[c] * b[c];   Wsc[r, c] = W[r, c] * c3[c]
__global__ void poolconv_coeffs_kernel(const float* __restrict__ W, const float* __restrict__ b, const float* __restrict__ c2,
                                       const float* __restrict__ c3, int cin, int cout, float* __restrict__ t,
                                       float* __restrict__ Wsc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cin * cout) {
    const int c = i % cout;
    Wsc[i] = W[i] * c3[c];
  }
  if (i < cout) t[i] = fmaf(c3[i], b[i], c2[i]);
}

// dx[cloud b, point amax[b,c], :] += c1[c] * dg[b,c] * W[:, c]     grid (cout/128, B), block 128 (thread = channel c)
__global__ void __launch_bounds__(128)
poolconv_sparse_dx_kernel(const float* __restrict__ dg, const int32_t* __restrict__ amax, const float* __restrict__ c1,
                          const float* __restrict__ W, int N, int cin, int cout, float* __restrict__ dx, long long lddx) {
  const int c = blockIdx.x * 128 + threadIdx.x, b = blockIdx.y;
  if (c >= cout) return;
  const float s = c1[c] * dg[(size_t)b * cout + c];
  if (s == 0.f) return;
  float* row = dx + ((size_t)b * N + amax[(size_t)b * cout + c]) * lddx;
  for (int r = 0; r < cin; ++r) atomicAdd(row + r, s * W[(size_t)r * cout + c]);
}

// sW[r, c] = sum_b c1[c] dg[b,c] A[b*N + amax[b,c], r];  sdb[c] = sum_b c1[c] dg[b,c]
// grid (cout/64, cin/8), block 64: thread = channel c, 8 input channels r0..r0+7; deterministic (b ascending)
__global__ void __launch_bounds__(64)
poolconv_sparse_dw_kernel(const float* __restrict__ dg, const int32_t* __restrict__ amax, const float* __restrict__ c1,
                          const float* __restrict__ A, long long lda, int B, int N, int cin, int cout,
                          float* __restrict__ sW, float* __restrict__ sdb) {
  const int c = blockIdx.x * 64 + threadIdx.x, r0 = blockIdx.y * 8;
  if (c >= cout) return;
  const float k1 = c1[c];
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float sb = 0.f;
  for (int b = 0; b < B; ++b) {
    const float s = k1 * dg[(size_t)b * cout + c];
    sb += s;
    if (s == 0.f) continue;
    const float* ar = A + ((size_t)b * N + amax[(size_t)b * cout + c]) * lda + r0;
    const float4 a0 = *reinterpret_cast<const float4*>(ar), a1 = *reinterpret_cast<const float4*>(ar + 4);
    acc[0] = fmaf(s, a0.x, acc[0]); acc[1] = fmaf(s, a0.y, acc[1]); acc[2] = fmaf(s, a0.z, acc[2]); acc[3] = fmaf(s, a0.w, acc[3]);
    acc[4] = fmaf(s, a1.x, acc[4]); acc[5] = fmaf(s, a1.y, acc[5]); acc[6] = fmaf(s, a1.z, acc[6]); acc[7] = fmaf(s, a1.w, acc[7]);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) sW[(size_t)(r0 + i) * cout + c] = acc[i];
  if (blockIdx.y == 0) sdb[c] = sb;
}

// dW[r,c] = T[r,c] + colsum[r] t[c] + sW[r,c];   db[c] = sdb[c] + P t[c] + sum_r colsum[r] Wsc[r,c]
__global__ void poolconv_finalize_kernel(const float* __restrict__ T, const float* __restrict__ colsum,
                                         const float* __restrict__ t, const float* __restrict__ sW,
                                         const float* __restrict__ sdb, const float* __restrict__ Wsc, int cin, int cout,
                                         double rows, float* __restrict__ dW, float* __restrict__ db) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cin * cout) {
    const int r = i / cout, c = i - r * cout;
    dW[i] = T[i] + fmaf(colsum[r], t[c], sW[i]);
  }
  if (db && i < cout) {
    double s = (double)sdb[i] + rows * (double)t[i];
    for (int r = 0; r < cin; ++r) s += (double)colsum[r] * (double)Wsc[(size_t)r * cout + i];
    db[i] = (float)s;
  }
}

// out[row, :] = v   (the constant row r0 of dA when no earlier GEMM writes dA)
__global__ void fill_rows_kernel(float* __restrict__ out, long long ldo, const float* __restrict__ v, long long rows, int C) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * C) return;
  const long long r = i / C;
  const int c = (int)(i - r * C);
  out[r * ldo + c] = v[c];
}

}  // namespace
}  // namespace wspc

using namespace wspc;

extern "C" int wspc_fill_rows(float* out, long long ldo, const float* v, long long rows, int C, wspc_stream_t stream) {
  if (int rc = check_arch()) return rc;
  WSPC_REQUIRE(out && v && rows >= 1 && C >= 1 && ldo >= C, "fill_rows: bad arguments");
  fill_rows_kernel<<<(unsigned)((rows * C + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(out, ldo, v, rows, C);
  count_launch();
  WSPC_LAUNCH_CHECK("fill_rows_kernel");
  return WSPC_OK;
}

extern "C" int wspc_poolconv_coeffs(const float* W, const float* b, const float* c2, const float* c3, int cin, int cout,
                                    float* t, float* Wsc, wspc_stream_t stream) {
  if (int rc = check_arch()) return rc;
  WSPC_REQUIRE(W && b && c2 && c3 && t && Wsc, "poolconv_coeffs: null pointer");
  WSPC_REQUIRE(cin >= 1 && cout >= 1, "poolconv_coeffs: bad shape");
  poolconv_coeffs_kernel<<<(cin * cout + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(W, b, c2, c3, cin, cout, t,
                                                                                                      Wsc);
  count_launch();
  WSPC_LAUNCH_CHECK("poolconv_coeffs_kernel");
  return WSPC_OK;
}

extern "C" int wspc_poolconv_sparse(const float* dg, const int32_t* amax, const float* c1, const float* W, const float* A,
                                    long long lda, int B, int N, int cin, int cout, float* dx, long long lddx, float* sW,
                                    float* sdb, wspc_stream_t stream) {
  if (int rc = check_arch()) return rc;
  WSPC_REQUIRE(dg && amax && c1 && W && A && sW && sdb, "poolconv_sparse: null pointer");
  WSPC_REQUIRE(B >= 1 && N >= 1 && cin >= 8 && (cin & 7) == 0 && cout >= 1 && (lda & 3) == 0 && aligned16(A),
               "poolconv_sparse: bad shape (cin multiple of 8, lda multiple of 4, A 16-byte aligned)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dx) {
    poolconv_sparse_dx_kernel<<<dim3((cout + 127) / 128, B), 128, 0, st>>>(dg, amax, c1, W, N, cin, cout, dx, lddx);
    count_launch();
  }
  poolconv_sparse_dw_kernel<<<dim3((cout + 63) / 64, cin / 8), 64, 0, st>>>(dg, amax, c1, A, lda, B, N, cin, cout, sW, sdb);
  count_launch();
  WSPC_LAUNCH_CHECK("poolconv_sparse kernels");
  return WSPC_OK;
}

extern "C" int wspc_poolconv_finalize(const float* T, const float* colsum, const float* t, const float* sW, const float* sdb,
                                      const float* Wsc, int cin, int cout, double rows, float* dW, float* db,
                                      wspc_stream_t stream) {
  if (int rc = check_arch()) return rc;
  WSPC_REQUIRE(T && colsum && t && sW && sdb && Wsc && dW, "poolconv_finalize: null pointer");
  poolconv_finalize_kernel<<<(cin * cout + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      T, colsum, t, sW, sdb, Wsc, cin, cout, rows, dW, db);
  count_launch();
  WSPC_LAUNCH_CHECK("poolconv_finalize_kernel");
  return WSPC_OK;
}
