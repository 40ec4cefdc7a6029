// fp32 CUDA-core GEMMs for the shared-MLP (1x1 conv) layers: generic in operand / epilogue mode.
//
// These kernels serve every layer shape of the two DGCNN variants (Cin 6..1280, Cout 13..1024),
// the small per-cloud products and the weight gradients.  (The large 64-wide EdgeConv layers are
// additionally served by the tcgen05 path in gemm_tc.cu.)
//
//   rowgemm : out(M,N) = A(M,K) * Bm(K,N)          forward and data-gradient
//   colgemm : dW(K1,K2) = sum_rows A^T dY           weight-gradient, deterministic slab reduction
#include "operand.cuh"

#include <stdlib.h>

namespace wspc {
void count_launch(int n = 1);
int wgrad_tc_slabs(int K1, int K2);
int bwd_fused_tc_dispatch(const Operand& A, int amode, const Operand& G, int gmode, long long M, int S, int K1p, int K2p,
                          float* partial, float* partial_b, const float* Wf, long long ldw, int Nf, const Epilogue& E,
                          cudaStream_t st);
int wgrad_tc_dispatch(const Operand& A, int amode, const Operand& G, int gmode, long long M, int S, int K1p, int K2p,
                      float* partial, float* partial_b, cudaStream_t st);
int rowgemm_tc_dispatch(const Operand& A, int amode, const float* Bm, long long ldb, int bT, long long M, int N, int K,
                        const Epilogue& E, int emode, void* ws, size_t ws_bytes, cudaStream_t st);
size_t rowgemm_tc_workspace_bytes(int N, int K);
static int g_gemm_path = 0;   // 0 = auto (tcgen05 where eligible), 1 = CUDA-core kernels only
namespace {

constexpr int BM = 128, BN = 64, BK = 16, ALD = BM + 4;
constexpr int RG_THREADS = 256;

template <int AMODE, int EMODE>
__global__ void __launch_bounds__(RG_THREADS)
rowgemm_kernel(const Operand A, const float* __restrict__ Bm, long long ldb, int bT, long long M, int N, int K,
               const Epilogue E) {
  __shared__ __align__(16) float As[2][BK][ALD];
  __shared__ __align__(16) float Bs[2][BK][BN];
  __shared__ float red[2][BN];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const long long row0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int nk = (K + BK - 1) / BK;

  // loader mapping
  const int a_lr = tid & 127, a_h = tid >> 7;        // A: row a_lr, channels kc + a_h*8 .. +8
  const long long a_row = row0 + a_lr;
  const bool a_valid = a_row < M;
  const int b_kk = tid >> 4, b_nn = (tid & 15) * 4;  // B: row b_kk, cols b_nn..+3

  float areg[8], breg[4];
  auto gload = [&](int kc) {
    if (a_valid) load8<AMODE>(A, a_row, kc * BK + a_h * 8, areg);
    else {
#pragma unroll
      for (int i = 0; i < 8; ++i) areg[i] = 0.f;
    }
    const int kg = kc * BK + b_kk;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + b_nn + j;
      float v = 0.f;
      if (kg < K && n < N) v = bT ? Bm[(long long)n * ldb + kg] : Bm[(long long)kg * ldb + n];
      breg[j] = v;
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 8; ++i) As[buf][a_h * 8 + i][a_lr] = areg[i];
    *reinterpret_cast<float4*>(&Bs[buf][b_kk][b_nn]) = make_float4(breg[0], breg[1], breg[2], breg[3]);
  };

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  gload(0);
  sstore(0);
  __syncthreads();
  for (int kc = 0; kc < nk; ++kc) {
    const int cur = kc & 1;
    if (kc + 1 < nk) gload(kc + 1);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[cur][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[cur][kk][64 + ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[cur][kk][tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kc + 1 < nk) {
      sstore(cur ^ 1);
      __syncthreads();
    }
  }

  // ------------------------------------------------------------ epilogue ---
  const int col0 = n0 + tx * 4;
  constexpr bool kStats = (EMODE == EPI_STORE_STATS || EMODE == EPI_RELUMASK_STATS);
  float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
  if (kStats) {
    if (tid < 2 * BN) (&red[0][0])[tid] = 0.f;
    __syncthreads();
  }

  if (EMODE == EPI_EDGE_SCATTER) {
    const int Cx = N >> 1;
    long long cur_pt = -1;
    float csum[4] = {0.f, 0.f, 0.f, 0.f};
    auto flush = [&]() {
      if (cur_pt >= 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int col = col0 + j;
          if (col < N) atomicAdd(E.dx + cur_pt * E.lddx + (col < Cx ? col : col - Cx), csum[j]);
        }
      }
    };
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const long long row = row0 + ((i < 4) ? (ty * 4 + i) : (64 + ty * 4 + i - 4));
      if (row < M) {
        const long long pt = row / E.k;
        if (pt != cur_pt) {
          flush();
          cur_pt = pt;
#pragma unroll
          for (int j = 0; j < 4; ++j) csum[j] = 0.f;
        }
        const long long nb = (pt / E.npts) * E.npts + E.idx[row];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int col = col0 + j;
          if (col < Cx) csum[j] += acc[i][j];
          else if (col < N) {
            csum[j] -= acc[i][j];
            atomicAdd(E.dx + nb * E.lddx + (col - Cx), acc[i][j]);
          }
        }
      }
    }
    flush();
  } else {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long row = row0 + ((i < 4) ? (ty * 4 + i) : (64 + ty * 4 + i - 4));
    if (row >= M) continue;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = acc[i][j];
    if (EMODE == EPI_STORE || EMODE == EPI_STORE_STATS) {
      const float* rb = E.rowbias ? E.rowbias + (row / E.rb_rows) * E.ldrb : nullptr;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = col0 + j;
        if (col < N) {
          if (E.bias) v[j] += E.bias[col];
          if (rb) v[j] += rb[col];
        }
      }
    } else if (EMODE == EPI_RELUMASK_STATS) {
      const float* yp = E.yprev + row * E.ldyp;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = col0 + j;
        if (col < N) {
          const float y = yp[col];
          const bool on = fmaf(y, E.scp[col], E.shp[col]) > 0.f;
          float g = on ? v[j] : 0.f;
          if (E.dmask) g *= E.dmask[row * N + col] * E.dscale;
          v[j] = g;
          s0[j] += g;
          s1[j] += g * y;
        }
      }
    } else if (EMODE == EPI_ACCUM) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = col0 + j;
        if (col < N) v[j] += E.out[row * E.ldo + col];
      }
    }
    if (EMODE == EPI_STORE_STATS) {
#pragma unroll
      for (int j = 0; j < 4; ++j) { s0[j] += v[j]; s1[j] += v[j] * v[j]; }
    }
    float* o = E.out + row * E.ldo + col0;
    if (col0 + 3 < N && ((E.ldo & 3) == 0) && aligned16(E.out)) {
      *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (col0 + j < N) o[j] = v[j];
    }
  }
  }

  if (kStats) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      atomicAdd(&red[0][tx * 4 + j], s0[j]);
      atomicAdd(&red[1][tx * 4 + j], s1[j]);
    }
    __syncthreads();
    if (tid < 2 * BN) {
      const int which = tid / BN, c = tid % BN;
      if (n0 + c < N) atomicAdd(E.stats + (size_t)which * N + n0 + c, (double)red[which][c]);
    }
  }
}

// ------------------------------------------------------------- colgemm -----
constexpr int CT = 64;    // output tile (K1 x K2)
constexpr int CR = 16;    // rows per smem chunk

template <int AMODE, int GMODE>
__global__ void __launch_bounds__(256)
colgemm_kernel(const Operand A, const Operand G, long long M, long long rows_per_slab, int K1p, int K2p,
               float* __restrict__ partial, float* __restrict__ partial_b) {
  __shared__ __align__(16) float As[2][CR][CT];
  __shared__ __align__(16) float Gs[2][CR][CT];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int k1_0 = blockIdx.x * CT, k2_0 = blockIdx.y * CT;
  const long long r_begin = (long long)blockIdx.z * rows_per_slab;
  const long long r_end = (r_begin + rows_per_slab < M) ? r_begin + rows_per_slab : M;

  const bool isA = tid < 128;
  const int lt = tid & 127;
  const int l_r = lt >> 3, l_c = (lt & 7) * 8;
  float reg[8];
  auto gload = [&](long long rbase) {
    const long long row = rbase + l_r;
    if (row < r_end) {
      if (isA) load8<AMODE>(A, row, k1_0 + l_c, reg);
      else     load8<GMODE>(G, row, k2_0 + l_c, reg);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) reg[i] = 0.f;
    }
  };
  auto sstore = [&](int buf) {
    float* dst = isA ? &As[buf][l_r][l_c] : &Gs[buf][l_r][l_c];
    *reinterpret_cast<float4*>(dst) = make_float4(reg[0], reg[1], reg[2], reg[3]);
    *reinterpret_cast<float4*>(dst + 4) = make_float4(reg[4], reg[5], reg[6], reg[7]);
  };

  float acc[4][4];
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  if (r_begin < r_end) {
    gload(r_begin);
    sstore(0);
    __syncthreads();
    int cur = 0;
    for (long long rb = r_begin; rb < r_end; rb += CR) {
      const bool more = rb + CR < r_end;
      if (more) gload(rb + CR);
#pragma unroll
      for (int r = 0; r < CR; ++r) {
        const float4 a = *reinterpret_cast<const float4*>(&As[cur][r][ty * 4]);
        const float4 g = *reinterpret_cast<const float4*>(&Gs[cur][r][tx * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w};
        const float gv[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], gv[j], acc[i][j]);
        if (ty == 0) {
#pragma unroll
          for (int j = 0; j < 4; ++j) bsum[j] += gv[j];
        }
      }
      if (more) {
        sstore(cur ^ 1);
        __syncthreads();
        cur ^= 1;
      }
    }
  }
  float* out = partial + (size_t)blockIdx.z * K1p * K2p;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k1 = k1_0 + ty * 4 + i;
    *reinterpret_cast<float4*>(out + (size_t)k1 * K2p + k2_0 + tx * 4) =
        make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
  }
  if (blockIdx.x == 0 && ty == 0) {
    *reinterpret_cast<float4*>(partial_b + (size_t)blockIdx.z * K2p + k2_0 + tx * 4) =
        make_float4(bsum[0], bsum[1], bsum[2], bsum[3]);
  }
}

__global__ void slab_reduce_kernel(const float* __restrict__ partial, const float* __restrict__ partial_b, int S,
                                   int K1, int K2, int K1p, int K2p, float* __restrict__ dW,
                                   float* __restrict__ db) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = K1 * K2;
  if (i < total) {
    const int k1 = i / K2, k2 = i % K2;
    double s = 0.0;
    for (int z = 0; z < S; ++z) s += (double)partial[((size_t)z * K1p + k1) * K2p + k2];
    dW[i] = (float)s;
  } else if (db && i < total + K2) {
    const int k2 = i - total;
    double s = 0.0;
    for (int z = 0; z < S; ++z) s += (double)partial_b[(size_t)z * K2p + k2];
    db[k2] = (float)s;
  }
}

struct WgradPlan {
  int K1p, K2p, t1, t2, S, S_tc;
  size_t partial_bytes, bias_bytes;
};
WgradPlan wgrad_plan(int K1, int K2) {
  WgradPlan p;
  p.t1 = (K1 + CT - 1) / CT;
  p.t2 = (K2 + CT - 1) / CT;
  p.K1p = p.t1 * CT;
  p.K2p = p.t2 * CT;
  const int tiles = p.t1 * p.t2;
  p.S = (4 * kNumSM + tiles - 1) / tiles;  // ~4 CTAs per SM in flight
  if (p.S < 1) p.S = 1;
  p.S_tc = wgrad_tc_slabs(K1, K2);
  const int smax = p.S > p.S_tc ? p.S : p.S_tc;
  p.partial_bytes = align_up((size_t)smax * p.K1p * p.K2p * sizeof(float), 256);
  p.bias_bytes = align_up((size_t)smax * p.K2p * sizeof(float), 256);
  return p;
}

template <int AMODE>
int launch_rows_e(const Operand& A, const float* Bm, long long ldb, int bT, long long M, int N, int K,
                  const Epilogue& E, int emode, cudaStream_t st) {
  dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((N + BN - 1) / BN));
#define WSPC_RG(EM)                                                                          \
  case EM:                                                                                   \
    rowgemm_kernel<AMODE, EM><<<grid, RG_THREADS, 0, st>>>(A, Bm, ldb, bT, M, N, K, E);      \
    break;
  switch (emode) {
    WSPC_RG(EPI_STORE)
    WSPC_RG(EPI_STORE_STATS)
    WSPC_RG(EPI_RELUMASK_STATS)
    WSPC_RG(EPI_ACCUM)
    WSPC_RG(EPI_EDGE_SCATTER)
    default:
      set_error("conv1x1_rows: bad epilogue mode %d", emode);
      return WSPC_ERR_INVALID;
  }
#undef WSPC_RG
  count_launch();
  WSPC_LAUNCH_CHECK("rowgemm_kernel");
  return WSPC_OK;
}

template <int AMODE>
int launch_wgrad_g(const Operand& A, const Operand& G, int gmode, long long M, const WgradPlan& p, float* partial,
                   float* partial_b, cudaStream_t st) {
  long long rps = (M + p.S - 1) / p.S;
  rps = (rps + CR - 1) / CR * CR;
  dim3 grid(p.t1, p.t2, p.S);
  if (gmode == OP_DY)
    colgemm_kernel<AMODE, OP_DY><<<grid, 256, 0, st>>>(A, G, M, rps, p.K1p, p.K2p, partial, partial_b);
  else if (gmode == OP_DY_SPARSE)
    colgemm_kernel<AMODE, OP_DY_SPARSE><<<grid, 256, 0, st>>>(A, G, M, rps, p.K1p, p.K2p, partial, partial_b);
  else if (gmode == OP_DY_MAXK)
    colgemm_kernel<AMODE, OP_DY_MAXK><<<grid, 256, 0, st>>>(A, G, M, rps, p.K1p, p.K2p, partial, partial_b);
  else {
    set_error("conv1x1_wgrad: bad gradient operand mode %d", gmode);
    return WSPC_ERR_INVALID;
  }
  count_launch();
  WSPC_LAUNCH_CHECK("colgemm_kernel");
  return WSPC_OK;
}


// ---------------------------------------------------------------- narrow inputs (K1 <= 12) ---
// Weight gradient of a layer with a handful of input channels (the factored first EdgeConv layer: X (P, 9 or 3) against the
// (P, 128) gradient of [u | v], tf_util.py:160-165 backward): dW(K1, K2) = X^T G.  The general kernels tile K1 to 64 and spend
// their time on padding; here a thread owns 4 gradient columns and all K1 rows of dW, streams G once with float4 loads
// (the only HBM traffic that matters: K2 * 4 B per row) and reads the row's inputs from a staged shared-memory tile.
constexpr int NK1_ROWS = 256;      // rows of X staged per round
__global__ void __launch_bounds__(256)
wgrad_narrowk1_kernel(const float* __restrict__ X, long long ldx, int K1, const float* __restrict__ G, long long ldg, int K2,
                      long long M, long long rps, int K1p, int K2p, float* __restrict__ partial, float* __restrict__ partial_b) {
  __shared__ __align__(16) float sx[NK1_ROWS * 12];
  __shared__ __align__(16) float red[256 * 4];
  const int tid = threadIdx.x;
  const int nq = K2 >> 2;                   // column quads
  const int cq = tid % nq, rp = tid / nq, RP = 256 / nq;
  const long long r0 = (long long)blockIdx.x * rps;
  const long long r1 = (r0 + rps < M) ? r0 + rps : M;
  float acc[12][4], accb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 12; ++j) { acc[j][0] = 0.f; acc[j][1] = 0.f; acc[j][2] = 0.f; acc[j][3] = 0.f; }
  for (long long base = r0; base < r1; base += NK1_ROWS) {
    for (int e = tid; e < NK1_ROWS * 12; e += 256) {
      const int rr = e / 12, j = e - rr * 12;
      const long long row = base + rr;
      sx[e] = (j < K1 && row < r1) ? X[row * ldx + j] : 0.f;
    }
    __syncthreads();
    const int nrow = (r1 - base < NK1_ROWS) ? (int)(r1 - base) : NK1_ROWS;
#pragma unroll 4
    for (int rr = rp; rr < nrow; rr += RP) {
      const float4 g = __ldcs(reinterpret_cast<const float4*>(G + (base + rr) * ldg + cq * 4));
      const float4* xr = reinterpret_cast<const float4*>(sx + rr * 12);
      const float4 x0 = xr[0], x1 = xr[1], x2 = xr[2];
      const float xv[12] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w, x2.x, x2.y, x2.z, x2.w};
#pragma unroll
      for (int j = 0; j < 12; ++j) {
        acc[j][0] = fmaf(xv[j], g.x, acc[j][0]); acc[j][1] = fmaf(xv[j], g.y, acc[j][1]);
        acc[j][2] = fmaf(xv[j], g.z, acc[j][2]); acc[j][3] = fmaf(xv[j], g.w, acc[j][3]);
      }
      accb[0] += g.x; accb[1] += g.y; accb[2] += g.z; accb[3] += g.w;
    }
    __syncthreads();
  }
  // the RP row phases of a column quad are summed through shared memory, one dW row (or the bias row) per round
#pragma unroll 1
  for (int j = 0; j <= K1; ++j) {
    float4 v;
    if (j < K1) {
      // select row j without dynamic register indexing
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
      for (int q = 0; q < 12; ++q)
        if (q == j) { a0 = acc[q][0]; a1 = acc[q][1]; a2 = acc[q][2]; a3 = acc[q][3]; }
      v = make_float4(a0, a1, a2, a3);
    } else {
      v = make_float4(accb[0], accb[1], accb[2], accb[3]);
    }
    *reinterpret_cast<float4*>(red + tid * 4) = v;
    __syncthreads();
    if (tid < K2) {
      const int q = tid >> 2, i = tid & 3;
      float t = 0.f;
      for (int ph = 0; ph < RP; ++ph) t += red[(ph * nq + q) * 4 + i];
      if (j < K1) partial[((size_t)blockIdx.x * K1p + j) * K2p + tid] = t;
      else if (partial_b) partial_b[(size_t)blockIdx.x * K2p + tid] = t;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------- narrow outputs (N <= 16) ---
// seg/conv3 of the S3DIS net (256 -> 13, DGCNN_S3DIS.py:100-101) and similar heads: one warp per row, the lane owns
// 8 input channels per 256-channel slice (coalesced operand load through load8), the transposed weights sit in
// shared memory, and the <= 16 partial dot products are reduced with warp shuffles.  out = acc + bias (EPI_STORE).
template <int AMODE>
__global__ void __launch_bounds__(256)
rows_narrow_kernel(const Operand A, const float* __restrict__ Bm, long long ldb, int bT, long long M, int N, int K,
                   const Epilogue E) {
  extern __shared__ float Ws[];                 // [N][K + 4]  (row pitch keeps float4 reads conflict-light)
  const int pitch = K + 4;
  for (int e = threadIdx.x; e < N * K; e += blockDim.x) {
    const int n = e / K, k = e - n * K;
    Ws[n * pitch + k] = bT ? Bm[(long long)n * ldb + k] : Bm[(long long)k * ldb + n];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long wstride = (long long)gridDim.x * (blockDim.x >> 5);
  const bool fast = aligned16(A.p) && (A.ld & 3) == 0 && (AMODE == OP_PLAIN || (aligned16(A.sc) && aligned16(A.sh) &&
                    (!A.dmask || (aligned16(A.dmask) && (A.C & 3) == 0))));
  // the operand of channels [c0, c0 + 8) of one row
  auto load_v = [&](long long row, int c0, float (&v)[8]) {
    if (fast) {   // 16-byte aligned rows: two float4 loads per tensor instead of eight scalar ones
      const float4 a0 = __ldcs(reinterpret_cast<const float4*>(A.p + row * A.ld + c0));
      const float4 a1 = __ldcs(reinterpret_cast<const float4*>(A.p + row * A.ld + c0 + 4));
      v[0] = a0.x; v[1] = a0.y; v[2] = a0.z; v[3] = a0.w; v[4] = a1.x; v[5] = a1.y; v[6] = a1.z; v[7] = a1.w;
      if (AMODE == OP_BNRELU) {
        const float4 s0 = *reinterpret_cast<const float4*>(A.sc + c0), s1 = *reinterpret_cast<const float4*>(A.sc + c0 + 4);
        const float4 h0 = *reinterpret_cast<const float4*>(A.sh + c0), h1 = *reinterpret_cast<const float4*>(A.sh + c0 + 4);
        const float sc8[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
        const float sh8[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = fmaxf(fmaf(v[i], sc8[i], sh8[i]), 0.f);
        if (A.dmask) {
          const float4 m0 = __ldcs(reinterpret_cast<const float4*>(A.dmask + row * A.C + c0));
          const float4 m1 = __ldcs(reinterpret_cast<const float4*>(A.dmask + row * A.C + c0 + 4));
          const float m8[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] *= m8[i] * A.dscale;
        }
      }
    } else {
      load8<AMODE>(A, row, c0, v);
    }
  };
  // 16 values x 32 lanes -> lane n holds the total of value n (halve the number of live values at every step), then store
  auto reduce_store = [&](float (&acc)[16], long long row) {
#pragma unroll
    for (int n = 0; n < 8; ++n) {     // step 16: lanes with bit 4 clear keep values 0..7, the others 8..15
      const float send = (lane & 16) ? acc[n] : acc[n + 8];
      const float keep = (lane & 16) ? acc[n + 8] : acc[n];
      acc[n] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      const float send = (lane & 8) ? acc[n] : acc[n + 4];
      const float keep = (lane & 8) ? acc[n + 4] : acc[n];
      acc[n] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int n = 0; n < 2; ++n) {
      const float send = (lane & 4) ? acc[n] : acc[n + 2];
      const float keep = (lane & 4) ? acc[n + 2] : acc[n];
      acc[n] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    {
      const float send = (lane & 2) ? acc[0] : acc[1];
      const float keep = (lane & 2) ? acc[1] : acc[0];
      acc[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], 1);
    // lane l now holds value n(l) = 8*b4 + 4*b3 + 2*b2 + b1 (bits of l); lanes with bit 0 clear write
    const int n = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
    if ((lane & 1) == 0 && n < N) E.out[row * E.ldo + n] = acc[0] + (E.bias ? E.bias[n] : 0.f);
  };
  // four rows per trip: every weight vector read from shared memory feeds all of them (the kernel is bound by those reads:
  // one row per trip 0.57 ms at cfg-3, two rows 0.50 ms)
  constexpr int RT = 4;
  for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + warp; row < M; row += RT * wstride) {
    float acc[RT][16];
#pragma unroll
    for (int r = 0; r < RT; ++r)
#pragma unroll
      for (int n = 0; n < 16; ++n) acc[r][n] = 0.f;
    for (int c0 = lane * 8; c0 < K; c0 += 256) {
      float v[RT][8];
#pragma unroll
      for (int r = 0; r < RT; ++r) {
        if (row + r * wstride < M) {
          load_v(row + r * wstride, c0, v[r]);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) v[r][i] = 0.f;
        }
      }
#pragma unroll
      for (int n = 0; n < 16; ++n) {
        if (n < N) {
          const float4 w0 = *reinterpret_cast<const float4*>(Ws + n * pitch + c0);
          const float4 w1 = *reinterpret_cast<const float4*>(Ws + n * pitch + c0 + 4);
#pragma unroll
          for (int r = 0; r < RT; ++r) {
            float a = acc[r][n];
            a = fmaf(v[r][0], w0.x, a); a = fmaf(v[r][1], w0.y, a); a = fmaf(v[r][2], w0.z, a); a = fmaf(v[r][3], w0.w, a);
            a = fmaf(v[r][4], w1.x, a); a = fmaf(v[r][5], w1.y, a); a = fmaf(v[r][6], w1.z, a); a = fmaf(v[r][7], w1.w, a);
            acc[r][n] = a;
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < RT; ++r)
      if (row + r * wstride < M) reduce_store(acc[r], row + r * wstride);
  }
}


// ---------------------------------------------------------------- narrow reductions (K <= 16) ---
// Data gradient of a narrow head (seg/conv3: dA (P,256) = dZ (P,13) W^T, DGCNN_S3DIS.py:100-101 backward) fused with
// the ReLU mask / dropout of the producing layer and its BN-backward sums (EPI_RELUMASK_STATS).  Thread = 4 output
// columns of a row; the K <= 16 gradient values of the row are a broadcast load; W^T sits in shared memory.
__global__ void __launch_bounds__(256)
rows_narrowk_relumask_kernel(const float* __restrict__ G, long long ldg, const float* __restrict__ Bm, long long ldb, int bT,
                             long long M, int N, int K, const Epilogue E) {
  extern __shared__ float sm[];
  float* Wt = sm;                                 // [K][N]
  float* red = sm + (size_t)K * N;                // [2][N]
  for (int e = threadIdx.x; e < K * N; e += blockDim.x) {
    const int k = e / N, n = e - k * N;
    Wt[e] = bT ? Bm[(long long)n * ldb + k] : Bm[(long long)k * ldb + n];
  }
  for (int e = threadIdx.x; e < 2 * N; e += blockDim.x) red[e] = 0.f;
  __syncthreads();
  const int n4 = N >> 2;
  const int c = (threadIdx.x % n4) * 4;
  const int rpb = blockDim.x / n4;                // rows per block sweep
  const float4 sc = *reinterpret_cast<const float4*>(E.scp + c), sh = *reinterpret_cast<const float4*>(E.shp + c);
  float f0[4] = {0.f, 0.f, 0.f, 0.f}, f1[4] = {0.f, 0.f, 0.f, 0.f};
  const float scv[4] = {sc.x, sc.y, sc.z, sc.w}, shv[4] = {sh.x, sh.y, sh.z, sh.w};
  // ReLU mask / dropout of the producing layer, its BN-backward sums, store
  auto finish = [&](long long row, float (&o)[4], const float4& y4, const float4& m4) {
    const float yv[4] = {y4.x, y4.y, y4.z, y4.w};
    const float dm[4] = {m4.x * E.dscale, m4.y * E.dscale, m4.z * E.dscale, m4.w * E.dscale};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool on = fmaf(yv[j], scv[j], shv[j]) > 0.f;
      o[j] = on ? (E.dmask ? o[j] * dm[j] : o[j]) : 0.f;
      f0[j] += o[j];
      f1[j] = fmaf(o[j], yv[j], f1[j]);
    }
    *reinterpret_cast<float4*>(E.out + row * E.ldo + c) = make_float4(o[0], o[1], o[2], o[3]);
  };
  // two rows per trip: the epilogue operands of both are requested before the products, and every weight vector read from
  // shared memory feeds both rows
  const long long rstride = (long long)gridDim.x * rpb;
  const float4 one4 = make_float4(1.f, 1.f, 1.f, 1.f);
  for (long long row = (long long)blockIdx.x * rpb + threadIdx.x / n4; row < M; row += 2 * rstride) {
    const long long rowb = row + rstride;
    const bool hasb = rowb < M;
    const float4 ya = *reinterpret_cast<const float4*>(E.yprev + row * E.ldyp + c);
    const float4 ma = E.dmask ? *reinterpret_cast<const float4*>(E.dmask + row * N + c) : one4;
    const float4 yb = hasb ? *reinterpret_cast<const float4*>(E.yprev + rowb * E.ldyp + c) : one4;
    const float4 mb = (hasb && E.dmask) ? *reinterpret_cast<const float4*>(E.dmask + rowb * N + c) : one4;
    const float* ga = G + row * ldg;
    const float* gb = G + (hasb ? rowb : row) * ldg;
    float oa[4] = {0.f, 0.f, 0.f, 0.f}, ob[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < K; ++k) {
      const float va = ga[k], vb = gb[k];
      const float4 w = *reinterpret_cast<const float4*>(Wt + (size_t)k * N + c);
      oa[0] = fmaf(va, w.x, oa[0]); oa[1] = fmaf(va, w.y, oa[1]); oa[2] = fmaf(va, w.z, oa[2]); oa[3] = fmaf(va, w.w, oa[3]);
      ob[0] = fmaf(vb, w.x, ob[0]); ob[1] = fmaf(vb, w.y, ob[1]); ob[2] = fmaf(vb, w.z, ob[2]); ob[3] = fmaf(vb, w.w, ob[3]);
    }
    finish(row, oa, ya, ma);
    if (hasb) finish(rowb, ob, yb, mb);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    atomicAdd(&red[c + j], f0[j]);
    atomicAdd(&red[N + c + j], f1[j]);
  }
  __syncthreads();
  for (int e = threadIdx.x; e < N; e += blockDim.x) {
    atomicAdd(E.stats + e, (double)red[e]);
    atomicAdd(E.stats + N + e, (double)red[N + e]);
  }
}

template <int AMODE>
int launch_rows_narrow(const Operand& A, const float* Bm, long long ldb, int bT, long long M, int N, int K, const Epilogue& E,
                       cudaStream_t st) {
  const size_t smem = (size_t)N * (K + 4) * sizeof(float);
  long long blocks = (M + 7) / 8;
  if (blocks > 8LL * kNumSM) blocks = 8LL * kNumSM;
  rows_narrow_kernel<AMODE><<<(unsigned)blocks, 256, smem, st>>>(A, Bm, ldb, bT, M, N, K, E);
  count_launch();
  WSPC_LAUNCH_CHECK("rows_narrow_kernel");
  return WSPC_OK;
}

}  // namespace
}  // namespace wspc

using namespace wspc;

extern "C" int wspc_set_gemm_path(int path) {
  const int old = wspc::g_gemm_path;
  wspc::g_gemm_path = path;
  return old;
}

extern "C" size_t wspc_conv1x1_rows_workspace_bytes(int N, int K) {
  if (N < 1 || K < 1) return 0;
  return rowgemm_tc_workspace_bytes(N, K);
}

extern "C" int wspc_conv1x1_rows_ws(const wspc_operand_t* A, int a_mode, const float* Bm, long long ldb,
                                    int b_transposed, long long M, int N, int K, const wspc_epilogue_t* epi,
                                    int epi_mode, void* workspace, size_t workspace_bytes, wspc_stream_t stream) {
  if (int rc = check_arch()) return rc;
  WSPC_REQUIRE(A && Bm && epi, "conv1x1_rows: null argument");
  WSPC_REQUIRE(M >= 1 && N >= 1 && K >= 1, "conv1x1_rows: bad shape M=%lld N=%d K=%d", M, N, K);
  WSPC_REQUIRE(A->p || a_mode == OP_DY_SPARSE, "conv1x1_rows: operand pointer is null");
  WSPC_REQUIRE(a_mode != OP_DY_MAXK || (A->y && A->c1 && A->c2 && A->c3 && A->sc && A->sh && A->k >= 1 && A->ld >= 2 * A->C),
               "conv1x1_rows: incomplete DY_MAXK operand");
  WSPC_REQUIRE(A->C == K, "conv1x1_rows: operand channels %d != K %d", A->C, K);
  if (epi_mode == EPI_EDGE_SCATTER) WSPC_REQUIRE(epi->dx && epi->idx && epi->k > 0 && epi->npts > 0 && (N % 2) == 0,
                                                 "conv1x1_rows: incomplete scatter epilogue");
  else WSPC_REQUIRE(epi->out, "conv1x1_rows: output pointer is null");
  if (epi_mode == EPI_STORE_STATS || epi_mode == EPI_RELUMASK_STATS)
    WSPC_REQUIRE(epi->stats, "conv1x1_rows: stats pointer is null");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // tensor-core (tcgen05) path for eligible shapes; WSPC_GEMM=simt forces the CUDA-core kernels (A/B testing)
  static const bool env_simt = []() { const char* e = getenv("WSPC_GEMM"); return e && strcmp(e, "simt") == 0; }();
  // narrow reductions (K <= 16): data gradient of a narrow head fused with the ReLU-mask / BN-sum epilogue.  Checked BEFORE the
  // tensor-core dispatch, which also accepts the shape but runs its element-wise operand loader for K = 13 (0.87 ms at cfg-3)
  if (!env_simt && g_gemm_path == 0 && epi_mode == EPI_RELUMASK_STATS && a_mode == OP_DY && !A->c1 && K <= 16 && N % 4 == 0 &&
      N <= 1024 && 256 % (N / 4) == 0 && aligned16(epi->out) && (epi->ldo % 4) == 0 && aligned16(epi->yprev) &&
      (epi->ldyp % 4) == 0 && aligned16(epi->scp) && aligned16(epi->shp) && (!epi->dmask || aligned16(epi->dmask))) {
    const size_t smem = ((size_t)K * N + 2 * (size_t)N) * sizeof(float);
    const int rpb = 256 / (N / 4);
    long long blocks = (M + rpb - 1) / rpb;
    if (blocks > 8LL * kNumSM) blocks = 8LL * kNumSM;
    rows_narrowk_relumask_kernel<<<(unsigned)blocks, 256, smem, st>>>(A->p, A->ld, Bm, ldb, b_transposed, M, N, K, *epi);
    count_launch();
    WSPC_LAUNCH_CHECK("rows_narrowk_relumask_kernel");
    return WSPC_OK;
  }
  // (Sending the products over a handful of rows -- the T-net's FC layers, the folded global feature: M = clouds -- to the fp32
  // CUDA-core kernel was measured: the 3x3 transform's error against fp64 moved from 6.8e-5 to 5.9e-5 only, because the
  // batch norm over the clouds amplifies the error of the layers in FRONT of them, and it cost 0.2-0.5 ms per step.)
  if (!env_simt && g_gemm_path == 0) {
    const int rc = rowgemm_tc_dispatch(*A, a_mode, Bm, ldb, b_transposed, M, N, K, *epi, epi_mode, workspace, workspace_bytes, st);
    if (rc != 0) return rc < 0 ? rc : WSPC_OK;
  }
  WSPC_REQUIRE(a_mode != OP_IMG, "conv1x1_rows: WSPC_OP_IMG is an operand of the tensor-core path only");
  // narrow outputs (N <= 16, K a multiple of 8): warp-per-row kernel
  if (!env_simt && g_gemm_path == 0 && epi_mode == EPI_STORE && N <= 16 && K % 8 == 0 && K >= 64 && K <= 2048 &&
      !epi->rowbias && (a_mode == OP_PLAIN || a_mode == OP_BNRELU) && (size_t)N * (K + 4) * 4 <= 48 * 1024) {
    if (a_mode == OP_PLAIN) return launch_rows_narrow<OP_PLAIN>(*A, Bm, ldb, b_transposed, M, N, K, *epi, st);
    return launch_rows_narrow<OP_BNRELU>(*A, Bm, ldb, b_transposed, M, N, K, *epi, st);
  }
  switch (a_mode) {
    case OP_PLAIN: return launch_rows_e<OP_PLAIN>(*A, Bm, ldb, b_transposed, M, N, K, *epi, epi_mode, st);
    case OP_BNRELU: return launch_rows_e<OP_BNRELU>(*A, Bm, ldb, b_transposed, M, N, K, *epi, epi_mode, st);
    case OP_EDGE: return launch_rows_e<OP_EDGE>(*A, Bm, ldb, b_transposed, M, N, K, *epi, epi_mode, st);
    case OP_DY: return launch_rows_e<OP_DY>(*A, Bm, ldb, b_transposed, M, N, K, *epi, epi_mode, st);
    case OP_DY_SPARSE: return launch_rows_e<OP_DY_SPARSE>(*A, Bm, ldb, b_transposed, M, N, K, *epi, epi_mode, st);
    case OP_DY_MAXK: return launch_rows_e<OP_DY_MAXK>(*A, Bm, ldb, b_transposed, M, N, K, *epi, epi_mode, st);
  }
  set_error("conv1x1_rows: bad operand mode %d", a_mode);
  return WSPC_ERR_INVALID;
}

extern "C" int wspc_conv1x1_rows(const wspc_operand_t* A, int a_mode, const float* Bm, long long ldb,
                                 int b_transposed, long long M, int N, int K, const wspc_epilogue_t* epi,
                                 int epi_mode, wspc_stream_t stream) {
  return wspc_conv1x1_rows_ws(A, a_mode, Bm, ldb, b_transposed, M, N, K, epi, epi_mode, nullptr, 0, stream);
}

extern "C" size_t wspc_conv1x1_wgrad_workspace_bytes(int K1, int K2) {
  if (K1 < 1 || K2 < 1) return 0;
  const WgradPlan p = wgrad_plan(K1, K2);
  return p.partial_bytes + p.bias_bytes;
}

extern "C" int wspc_conv1x1_wgrad(const wspc_operand_t* A, int a_mode, const wspc_operand_t* G, int g_mode,
                                  long long M, float* dW, float* db, void* workspace, size_t workspace_bytes,
                                  wspc_stream_t stream) {
  if (int rc = check_arch()) return rc;
  WSPC_REQUIRE(A && G && dW && workspace, "conv1x1_wgrad: null argument");
  WSPC_REQUIRE(M >= 1 && A->C >= 1 && G->C >= 1, "conv1x1_wgrad: bad shape");
  const WgradPlan p = wgrad_plan(A->C, G->C);
  if (workspace_bytes < p.partial_bytes + p.bias_bytes) {
    set_error("conv1x1_wgrad: workspace %zu < required %zu", workspace_bytes, p.partial_bytes + p.bias_bytes);
    return WSPC_ERR_WORKSPACE;
  }
  float* partial = static_cast<float*>(workspace);
  float* partial_b = reinterpret_cast<float*>(static_cast<char*>(workspace) + p.partial_bytes);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int rc;
  int S_used = p.S;
  static const bool env_simt = []() { const char* e = getenv("WSPC_GEMM"); return e && strcmp(e, "simt") == 0; }();
  if (a_mode == OP_PLAIN && A->C <= 12 && g_mode == OP_DY && !G->c1 && (G->C % 4) == 0 && G->C <= 256 && 256 % (G->C / 4) == 0 &&
      aligned16(G->p) && (G->ld % 4) == 0) {
    const long long rps = (M + p.S - 1) / p.S;
    wgrad_narrowk1_kernel<<<p.S, 256, 0, st>>>(A->p, A->ld, A->C, G->p, G->ld, G->C, M, rps, p.K1p, p.K2p, partial,
                                              db ? partial_b : nullptr);
    count_launch();
    WSPC_LAUNCH_CHECK("wgrad_narrowk1_kernel");
    goto reduce;
  }
  if (!env_simt && g_gemm_path == 0) {
    rc = wgrad_tc_dispatch(*A, a_mode, *G, g_mode, M, p.S_tc, p.K1p, p.K2p, partial, db ? partial_b : nullptr, st);
    if (rc < 0) return rc;
    if (rc == 1) { S_used = p.S_tc; goto reduce; }
  }
  switch (a_mode) {
    case OP_PLAIN: rc = launch_wgrad_g<OP_PLAIN>(*A, *G, g_mode, M, p, partial, partial_b, st); break;
    case OP_BNRELU: rc = launch_wgrad_g<OP_BNRELU>(*A, *G, g_mode, M, p, partial, partial_b, st); break;
    case OP_EDGE: rc = launch_wgrad_g<OP_EDGE>(*A, *G, g_mode, M, p, partial, partial_b, st); break;
    default:
      set_error("conv1x1_wgrad: bad operand mode %d", a_mode);
      return WSPC_ERR_INVALID;
  }
  if (rc) return rc;
reduce:
  const int total = A->C * G->C + (db ? G->C : 0);
  slab_reduce_kernel<<<(total + 255) / 256, 256, 0, st>>>(partial, partial_b, S_used, A->C, G->C, p.K1p, p.K2p, dW, db);
  count_launch();
  WSPC_LAUNCH_CHECK("slab_reduce_kernel");
  return WSPC_OK;
}

// Weight gradient AND data gradient of one conv2d in a single pass over the rows (tcgen05 path; shapes it does not cover
// run the two separate kernels).  dW(K1,K2), db(K2) as wspc_conv1x1_wgrad; dA(M, K1) = dY W^T through `epi`
// (WSPC_EPI_RELUMASK_STATS: ReLU mask / dropout of the producing layer + its BN-backward sums).
extern "C" int wspc_conv1x1_bwd_fused(const wspc_operand_t* A, int a_mode, const wspc_operand_t* G, int g_mode, long long M,
                                      const float* W, long long ldw, const wspc_epilogue_t* epi, float* dW, float* db,
                                      void* workspace, size_t workspace_bytes, wspc_stream_t stream) {
  if (int rc = check_arch()) return rc;
  WSPC_REQUIRE(A && G && W && epi && dW && workspace, "conv1x1_bwd_fused: null argument");
  WSPC_REQUIRE(M >= 1 && A->C >= 1 && G->C >= 1, "conv1x1_bwd_fused: bad shape");
  WSPC_REQUIRE(epi->out && epi->stats && epi->yprev && epi->scp && epi->shp, "conv1x1_bwd_fused: incomplete epilogue");
  const WgradPlan p = wgrad_plan(A->C, G->C);
  if (workspace_bytes < p.partial_bytes + p.bias_bytes) {
    set_error("conv1x1_bwd_fused: workspace %zu < required %zu", workspace_bytes, p.partial_bytes + p.bias_bytes);
    return WSPC_ERR_WORKSPACE;
  }
  float* partial = static_cast<float*>(workspace);
  float* partial_b = reinterpret_cast<float*>(static_cast<char*>(workspace) + p.partial_bytes);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  static const bool env_simt = []() { const char* e = getenv("WSPC_GEMM"); return e && strcmp(e, "simt") == 0; }();
  if (!env_simt && g_gemm_path == 0) {
    const int rc = bwd_fused_tc_dispatch(*A, a_mode, *G, g_mode, M, p.S_tc, p.K1p, p.K2p, partial, db ? partial_b : nullptr, W, ldw,
                                         A->C, *epi, st);
    if (rc < 0) return rc;
    if (rc == 1) {
      const int total = A->C * G->C + (db ? G->C : 0);
      slab_reduce_kernel<<<(total + 255) / 256, 256, 0, st>>>(partial, partial_b, p.S_tc, A->C, G->C, p.K1p, p.K2p, dW, db);
      count_launch();
      WSPC_LAUNCH_CHECK("slab_reduce_kernel");
      return WSPC_OK;
    }
  }
  // not eligible: the two separate kernels
  if (int rc = wspc_conv1x1_wgrad(A, a_mode, G, g_mode, M, dW, db, workspace, workspace_bytes, stream)) return rc;
  return wspc_conv1x1_rows(G, g_mode, W, ldw, 1, M, A->C, G->C, epi, EPI_RELUMASK_STATS, stream);
}
