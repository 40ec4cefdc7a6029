// Optimiser, dropout-mask generation and small utilities.
#include "common.cuh"

namespace wspc {
void count_launch(int n = 1);
namespace {

// tf.train.AdamOptimizer [TF] (S3DIS_DGCNN_trainer.py:110): epsilon is NOT bias-corrected:
//   m <- b1 m + (1-b1) g ; v <- b2 v + (1-b2) g^2 ; theta <- theta - lr_t * m / (sqrt(v) + eps)
// with lr_t = lr * sqrt(1-b2^t) / (1-b1^t) computed by the caller.  gscale folds the 1/world_size of the
// data-parallel gradient all-reduce into the same pass.
__global__ void adam_tf_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                               float* __restrict__ v, long long n, float lr_t, float b1, float b2, float eps,
                               float gscale) {
  const long long i4 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 + 3 < n) {
    float4 P = *reinterpret_cast<float4*>(p + i4);
    const float4 G = *reinterpret_cast<const float4*>(g + i4);
    float4 M = *reinterpret_cast<float4*>(m + i4);
    float4 V = *reinterpret_cast<float4*>(v + i4);
#define WSPC_ADAM1(x)                                   \
  {                                                     \
    const float gg = G.x * gscale;                      \
    M.x = b1 * M.x + (1.f - b1) * gg;                   \
    V.x = b2 * V.x + (1.f - b2) * gg * gg;              \
    P.x -= lr_t * M.x / (sqrtf(V.x) + eps);             \
  }
    WSPC_ADAM1(x) WSPC_ADAM1(y) WSPC_ADAM1(z) WSPC_ADAM1(w)
#undef WSPC_ADAM1
    *reinterpret_cast<float4*>(p + i4) = P;
    *reinterpret_cast<float4*>(m + i4) = M;
    *reinterpret_cast<float4*>(v + i4) = V;
  } else {
    for (long long i = i4; i < n; ++i) {
      const float gg = g[i] * gscale;
      const float mm = b1 * m[i] + (1.f - b1) * gg;
      const float vv = b2 * v[i] + (1.f - b2) * gg * gg;
      m[i] = mm;
      v[i] = vv;
      p[i] -= lr_t * mm / (sqrtf(vv) + eps);
    }
  }
}

// Philox-4x32-10 counter RNG (Salmon et al. 2011), one 128-bit block -> 4 uniforms
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
  c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}

// tf.nn.dropout [TF] (tf_util.py:631-635): keep element iff floor(keep + U[0,1)) == 1, i.e. U >= 1 - keep
__global__ void dropout_mask_kernel(float* __restrict__ mask, long long n, float keep, uint64_t seed, uint64_t offset) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long i4 = t * 4;
  if (i4 >= n) return;
  uint32_t c[4] = {(uint32_t)(t + offset), (uint32_t)((t + offset) >> 32), 0u, 0u};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (i4 + j < n) {
      const float u = (float)(c[j] >> 8) * (1.0f / 16777216.0f);   // [0,1)
      mask[i4 + j] = (floorf(keep + u) >= 1.f) ? 1.f : 0.f;
    }
  }
}

}  // namespace
}  // namespace wspc

using namespace wspc;

extern "C" int wspc_adam_tf(float* p, const float* g, float* m, float* v, long long n, float lr_t, float b1, float b2,
                            float eps, float gscale, wspc_stream_t stream) {
  if (int rc = check_arch()) return rc;
  WSPC_REQUIRE(p && g && m && v && n >= 1, "adam_tf: bad argument");
  WSPC_REQUIRE(aligned16(p) && aligned16(g) && aligned16(m) && aligned16(v), "adam_tf: buffers must be 16-byte aligned");
  const long long nthread = (n + 3) / 4;
  adam_tf_kernel<<<(unsigned)((nthread + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      p, g, m, v, n, lr_t, b1, b2, eps, gscale);
  count_launch();
  WSPC_LAUNCH_CHECK("adam_tf_kernel");
  return WSPC_OK;
}

extern "C" int wspc_dropout_mask(float* mask, long long n, float keep, uint64_t seed, uint64_t offset,
                                 wspc_stream_t stream) {
  if (int rc = check_arch()) return rc;
  WSPC_REQUIRE(mask && n >= 1 && keep > 0.f && keep <= 1.f, "dropout_mask: bad argument");
  const long long nthread = (n + 3) / 4;
  dropout_mask_kernel<<<(unsigned)((nthread + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      mask, n, keep, seed, offset);
  count_launch();
  WSPC_LAUNCH_CHECK("dropout_mask_kernel");
  return WSPC_OK;
}

extern "C" int wspc_zero(void* ptr, size_t bytes, wspc_stream_t stream) {
  WSPC_REQUIRE(ptr || bytes == 0, "zero: null pointer");
  if (bytes) WSPC_CUDA(cudaMemsetAsync(ptr, 0, bytes, reinterpret_cast<cudaStream_t>(stream)));
  return WSPC_OK;
}

// ---- input transform (T-net): X' = X (T + I)    ShapeNet/DGCNN_ShapeNet.py:29, transform_nets.py:51-55 --------
namespace wspc {
namespace {
__global__ void transform_fwd_kernel(const float* __restrict__ X, const float* __restrict__ T, int N, int add_eye,
                                     float* __restrict__ Xt) {
  const int b = blockIdx.y;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float t[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) t[i] = T[b * 9 + i] + ((add_eye && (i % 4) == 0) ? 1.f : 0.f);
  const float* x = X + ((size_t)b * N + n) * 3;
  float* o = Xt + ((size_t)b * N + n) * 3;
  const float x0 = x[0], x1 = x[1], x2 = x[2];
#pragma unroll
  for (int j = 0; j < 3; ++j) o[j] = fmaf(x2, t[6 + j], fmaf(x1, t[3 + j], x0 * t[j]));   // row-vector times matrix
}

// dT[b] = X[b]^T dXt[b]  (3x3 per cloud; tf.matmul gradient w.r.t. the second operand)
__global__ void transform_bwd_kernel(const float* __restrict__ X, const float* __restrict__ dXt, int N,
                                     float* __restrict__ dT) {
  __shared__ float red[9][8];
  const int b = blockIdx.x;
  float acc[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) acc[i] = 0.f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const float* x = X + ((size_t)b * N + n) * 3;
    const float* g = dXt + ((size_t)b * N + n) * 3;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) acc[i * 3 + j] = fmaf(x[i], g[j], acc[i * 3 + j]);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const float v = warp_sum(acc[i]);
    if (lane == 0) red[i][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < 9) {
    float s = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[threadIdx.x][w];
    dT[b * 9 + threadIdx.x] = s;
  }
}
}  // namespace
}  // namespace wspc

extern "C" int wspc_transform_points_fwd(const float* X, const float* T, int B, int N, int add_eye, float* Xt,
                                         wspc_stream_t stream) {
  if (int rc = check_arch()) return rc;
  WSPC_REQUIRE(X && T && Xt && B >= 1 && N >= 1, "transform_points_fwd: bad argument");
  transform_fwd_kernel<<<dim3((N + 255) / 256, B), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(X, T, N, add_eye, Xt);
  count_launch();
  WSPC_LAUNCH_CHECK("transform_fwd_kernel");
  return WSPC_OK;
}

extern "C" int wspc_transform_points_bwd(const float* X, const float* dXt, int B, int N, float* dT, wspc_stream_t stream) {
  if (int rc = check_arch()) return rc;
  WSPC_REQUIRE(X && dXt && dT && B >= 1 && N >= 1, "transform_points_bwd: bad argument");
  transform_bwd_kernel<<<B, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(X, dXt, N, dT);
  count_launch();
  WSPC_LAUNCH_CHECK("transform_bwd_kernel");
  return WSPC_OK;
}
