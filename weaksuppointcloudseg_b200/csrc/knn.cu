// Fused pairwise-distance + k-nearest-neighbour selection for sm_100a.
//
// Replaces tf_util.pairwise_distance + tf_util.knn
// (reference Networks/dgcnn/utils/tf_util.py:638-671) and the Dmat/top_k block
// of Util/SmoothConstraint.py:141-154.  The N x N matrix is never written:
// each CTA owns 128 query rows, streams 128-column candidate tiles of the
// transposed cloud through shared memory with 1-D bulk async copies (TMA
// engine, mbarrier completion), accumulates the dot products as sequential
// fp32 FMA chains in registers (the canonical arithmetic of SURVEY.md App. A-1,
// so indices are bit-exact against oracle/knn_oracle.c), parks the 128x128
// distance tile in shared memory and lets each warp keep the sorted k-best
// lists of its 16 rows in registers (one list entry per lane).
#include "common.cuh"
#include <math_constants.h>
#include <limits.h>

namespace wspc {
void count_launch(int n = 1);
bool knn_tc_eligible(int N, int D, int k);
size_t knn_tc_workspace_bytes(int B, int N, int D);
int knn_tc_run(const float* x, int B, int N, int ldx, int coff, int D, int k, int flavour, int32_t* idx, float* dist,
               void* ws, size_t ws_bytes, cudaStream_t st);
int knn_tc_fallback_rows(const void* ws, int B, int N, int D, int* out);
void knn_tc_clear_fallback_rows(void* ws, int B, int N, int D, cudaStream_t st);
static int g_knn_path = 0;   // 0 = auto (tcgen05 distances + exact re-scoring for wide features), 1 = CUDA-core kernel only

namespace {

constexpr int TM = 64;       // query rows per CTA (two CTAs per SM overlap GEMM and selection phases)
constexpr int TN = 128;      // candidate columns per tile
constexpr int DLD = TN + 4;  // distance-tile leading dimension (floats)
constexpr int NSTAGE = 3;    // bulk-copy pipeline depth
constexpr int KNN_THREADS = TM * 4;   // TM/8 warps; each warp: 8 rows x 128 columns of accumulators

// ------------------------------------------------------------------ prep ---
// x (B,N,ldx)[coff:coff+D] -> xT (B,Dp,Npad) zero padded, sq (B,Npad):
// sq is the same fmaf chain as the dot product so that d_ii == 0 exactly.
__global__ void knn_prep_kernel(const float* __restrict__ x, int N, int ldx, int coff, int D, int Dp,
                                int Npad, float* __restrict__ xT, float* __restrict__ sq) {
  const int b = blockIdx.y;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= Npad) return;
  float* xTb = xT + (size_t)b * Dp * Npad;
  float acc = 0.f;
  if (n < N) {
    const float* xr = x + ((size_t)b * N + n) * ldx + coff;
    for (int c = 0; c < D; ++c) {
      const float v = xr[c];
      xTb[(size_t)c * Npad + n] = v;
      acc = __fmaf_rn(v, v, acc);
    }
  } else {
    for (int c = 0; c < D; ++c) xTb[(size_t)c * Npad + n] = 0.f;
  }
  for (int c = D; c < Dp; ++c) xTb[(size_t)c * Npad + n] = 0.f;
  sq[(size_t)b * Npad + n] = acc;
}

// ------------------------------------------------------ sorted list in regs --
// A warp holds, per row, a lexicographically sorted (dist, idx) list of
// 32*KSLOT entries: entry e lives in lane e%32, register e/32.  Sentinels
// (+inf, INT_MAX) fill the tail.
template <int KSLOT>
__device__ __forceinline__ void list_insert(float (&ld)[KSLOT], int (&li)[KSLOT], float cd, int cj,
                                            int lane) {
  int pos = 0;
#pragma unroll
  for (int s = 0; s < KSLOT; ++s) {
    const bool lt = (ld[s] < cd) || (ld[s] == cd && li[s] < cj);
    pos += __popc(__ballot_sync(0xffffffffu, lt));
  }
  float carry_d = 0.f;
  int carry_i = 0;
#pragma unroll
  for (int s = 0; s < KSLOT; ++s) {
    float up_d = __shfl_up_sync(0xffffffffu, ld[s], 1);
    int up_i = __shfl_up_sync(0xffffffffu, li[s], 1);
    const float last_d = __shfl_sync(0xffffffffu, ld[s], 31);
    const int last_i = __shfl_sync(0xffffffffu, li[s], 31);
    if (lane == 0) { up_d = carry_d; up_i = carry_i; }
    const int e = s * 32 + lane;
    if (e > pos) { ld[s] = up_d; li[s] = up_i; }
    else if (e == pos) { ld[s] = cd; li[s] = cj; }
    carry_d = last_d;
    carry_i = last_i;
  }
}

// Insert when the candidate's index is known to exceed every stored index (columns are visited in
// strictly ascending order per row): (d', j') <lex (d, j)  <=>  d' <= d, so no index compare is needed.
template <int KSLOT>
__device__ __forceinline__ void list_insert_ordered(float (&ld)[KSLOT], int (&li)[KSLOT], float cd, int cj,
                                                    int lane) {
  if (KSLOT == 1) {
    const int pos = __popc(__ballot_sync(0xffffffffu, ld[0] <= cd));
    const float up_d = __shfl_up_sync(0xffffffffu, ld[0], 1);
    const int up_i = __shfl_up_sync(0xffffffffu, li[0], 1);
    const bool gt = lane > pos;
    ld[0] = gt ? up_d : ld[0];
    li[0] = gt ? up_i : li[0];
    if (lane == pos) { ld[0] = cd; li[0] = cj; }
  } else {
    int pos = 0;
#pragma unroll
    for (int s = 0; s < KSLOT; ++s) pos += __popc(__ballot_sync(0xffffffffu, ld[s] <= cd));
    float carry_d = 0.f;
    int carry_i = 0;
#pragma unroll
    for (int s = 0; s < KSLOT; ++s) {
      float up_d = __shfl_up_sync(0xffffffffu, ld[s], 1);
      int up_i = __shfl_up_sync(0xffffffffu, li[s], 1);
      const float last_d = __shfl_sync(0xffffffffu, ld[s], 31);
      const int last_i = __shfl_sync(0xffffffffu, li[s], 31);
      if (lane == 0) { up_d = carry_d; up_i = carry_i; }
      const int e = s * 32 + lane;
      if (e > pos) { ld[s] = up_d; li[s] = up_i; }
      else if (e == pos) { ld[s] = cd; li[s] = cj; }
      carry_d = last_d;
      carry_i = last_i;
    }
  }
}

template <int KSLOT>
__device__ __forceinline__ float list_tau_d(const float (&ld)[KSLOT], int k) {
  const int e = k - 1;
  return __shfl_sync(0xffffffffu, (KSLOT == 1 || e < 32) ? ld[0] : ld[KSLOT - 1], e & 31);
}

template <int KSLOT>
__device__ __forceinline__ void list_tau(const float (&ld)[KSLOT], const int (&li)[KSLOT], int k,
                                         float& td, int& ti) {
  const int e = k - 1;
  if (KSLOT == 1 || e < 32) {
    td = __shfl_sync(0xffffffffu, ld[0], e & 31);
    ti = __shfl_sync(0xffffffffu, li[0], e & 31);
  } else {
    td = __shfl_sync(0xffffffffu, ld[KSLOT - 1], e & 31);
    ti = __shfl_sync(0xffffffffu, li[KSLOT - 1], e & 31);
  }
}

__device__ __forceinline__ float dist_value(int flavour, float sqi, float sqj, float dot) {
  if (flavour == WSPC_DIST_TFUTIL) {
    // (sq_i + (-2*dot)) + sq_j       tf_util.py:654-657
    return __fadd_rn(__fadd_rn(sqi, __fmul_rn(-2.f, dot)), sqj);
  }
  // (X2_i + Y2_j) - 2*XY, negatives clamped to 0   SmoothConstraint.py:147-148
  const float d = __fsub_rn(__fadd_rn(sqi, sqj), __fmul_rn(2.f, dot));
  return d > 0.f ? d : 0.f;
}

// ---------------------------------------------------------------- main ------
// MODE 0: fused top-k (idx/dist out).  MODE 1: write the full adjacency.
//
// TM*4 threads = TM/8 warps.  GEMM phase: thread (tx = tid%32, ty = tid/32) owns rows {ty*4+i, TM/2+ty*4+i}
// x cols {tx*4+j} (8x4 accumulators, one sequential fmaf chain each).  Selection phase: warp w owns
// rows w*8..w*8+7; lane l looks at columns {l, l+32, l+64, l+96} of the distance tile.  For each
// 32-column group the warp ballots "beats the current k-th" per row and inserts the survivors with
// warp-shuffle insertion; four rows are inserted in lock-step (independent dependency chains, padded
// with sentinel no-op inserts) so the shuffle/ballot latency of one row hides behind the others.
template <int KC, int KSLOT, int MODE>
__global__ void __launch_bounds__(KNN_THREADS, 2)
knn_tile_kernel(const float* __restrict__ xT, const float* __restrict__ sq, int N, int Npad, int Dp,
                int k, int flavour, int32_t* __restrict__ idx_out, float* __restrict__ dist_out,
                float* __restrict__ adj_out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* As = reinterpret_cast<float*>(smem_raw);        // [Dp][128]
  float* Bs = As + (size_t)Dp * TM;                       // [NSTAGE][KC][128]
  float* Ds = Bs + (size_t)NSTAGE * KC * TN;              // [128][DLD]  (MODE 0 only)
  uint64_t* bars = reinterpret_cast<uint64_t*>(Ds + (MODE == 0 ? TM * DLD : 0));  // full[NSTAGE], abar

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int tx = lane;
  const int ty = warp;
  const int b = blockIdx.y;
  const int i0 = blockIdx.x * TM;
  const float* xTb = xT + (size_t)b * Dp * Npad;
  const float* sqb = sq + (size_t)b * Npad;
  const int nchunk = Dp / KC;
  const int ntile = Npad / TN;
  const int total = ntile * nchunk;

  auto issue = [&](int t) {  // thread 0 only: stream chunk t into stage t % NSTAGE
    const int s = t % NSTAGE;
    const int jt = t / nchunk, c = t - jt * nchunk;
    float* dst = Bs + (size_t)s * KC * TN;
    const float* src = xTb + (size_t)(c * KC) * Npad + (size_t)jt * TN;
    mbar_expect_tx(&bars[s], KC * TN * 4);
#pragma unroll 4
    for (int r = 0; r < KC; ++r) bulk_g2s(dst + r * TN, src + (size_t)r * Npad, TN * 4, &bars[s]);
  };

  if (tid == 0) {
    for (int s = 0; s <= NSTAGE; ++s) mbar_init(&bars[s], 1);
    mbar_fence_init();
    mbar_expect_tx(&bars[NSTAGE], Dp * TM * 4);
    for (int r = 0; r < Dp; ++r) bulk_g2s(As + r * TM, xTb + (size_t)r * Npad + i0, TM * 4, &bars[NSTAGE]);
    for (int t = 0; t < NSTAGE && t < total; ++t) issue(t);
  }
  __syncthreads();

  // per-thread row norms (rows ty*4+i and TM/2+ty*4+i)
  float sqa[8];
  {
    const float4 s0 = *reinterpret_cast<const float4*>(sqb + i0 + ty * 4);
    const float4 s1 = *reinterpret_cast<const float4*>(sqb + i0 + TM / 2 + ty * 4);
    sqa[0] = s0.x; sqa[1] = s0.y; sqa[2] = s0.z; sqa[3] = s0.w;
    sqa[4] = s1.x; sqa[5] = s1.y; sqa[6] = s1.z; sqa[7] = s1.w;
  }

  // sorted lists: this warp owns rows warp*8 .. warp*8+7; taud/taui = current k-th entry (all lanes)
  float ld[8][KSLOT];
  int li[8][KSLOT];
  float taud[8];
  if (MODE == 0) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int s = 0; s < KSLOT; ++s) { ld[r][s] = CUDART_INF_F; li[r][s] = INT_MAX; }
      taud[r] = CUDART_INF_F;
    }
  }

  mbar_wait(&bars[NSTAGE], 0);

  float acc[8][4];
  for (int t = 0; t < total; ++t) {
    const int s = t % NSTAGE;
    const int jt = t / nchunk, c = t - jt * nchunk;
    if (c == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    }
    mbar_wait(&bars[s], (t / NSTAGE) & 1);

    const float* Ap = As + (size_t)(c * KC) * TM;
    const float* Bp = Bs + (size_t)s * KC * TN;
#pragma unroll(KC < 8 ? KC : 8)
    for (int kk = 0; kk < KC; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(Ap + kk * TM + ty * 4);
      const float4 a1 = *reinterpret_cast<const float4*>(Ap + kk * TM + TM / 2 + ty * 4);
      const float4 b0 = *reinterpret_cast<const float4*>(Bp + kk * TN + tx * 4);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[4] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = __fmaf_rn(av[i], bv[j], acc[i][j]);
    }

    const bool last = (c == nchunk - 1);
    const int col0 = jt * TN;
    if (last) {
      const float4 q0 = *reinterpret_cast<const float4*>(sqb + col0 + tx * 4);
      const float sqj[4] = {q0.x, q0.y, q0.z, q0.w};
      if (MODE == 0) {
        __syncthreads();  // every warp has finished scanning the previous distance tile
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = (i < 4) ? (ty * 4 + i) : (TM / 2 + ty * 4 + i - 4);
          float4 d0;
          d0.x = dist_value(flavour, sqa[i], sqj[0], acc[i][0]);
          d0.y = dist_value(flavour, sqa[i], sqj[1], acc[i][1]);
          d0.z = dist_value(flavour, sqa[i], sqj[2], acc[i][2]);
          d0.w = dist_value(flavour, sqa[i], sqj[3], acc[i][3]);
          if (col0 + tx * 4 + 3 >= N) {   // ragged last tile: padded columns never pass the filter
            const float qnan = __int_as_float(0x7fc00000);
            if (col0 + tx * 4 + 0 >= N) d0.x = qnan;
            if (col0 + tx * 4 + 1 >= N) d0.y = qnan;
            if (col0 + tx * 4 + 2 >= N) d0.z = qnan;
            d0.w = qnan;
          }
          *reinterpret_cast<float4*>(Ds + row * DLD + tx * 4) = d0;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = i0 + ((i < 4) ? (ty * 4 + i) : (TM / 2 + ty * 4 + i - 4));
          if (row < N) {
            float* orow = adj_out + ((size_t)b * N + row) * N;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int col = col0 + tx * 4 + j;
              if (col < N) orow[col] = dist_value(flavour, sqa[i], sqj[j], acc[i][j]);
            }
          }
        }
      }
    }
    __syncthreads();  // stage s fully consumed; distance tile visible
    if (tid == 0 && t + NSTAGE < total) issue(t + NSTAGE);

    if (MODE == 0 && last) {
      // this warp's 8 rows x 4 column groups (lane + 32*q): conflict-free LDS.32
      float v[8][4];
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const float* drow = Ds + (warp * 8 + r) * DLD + lane;
#pragma unroll
        for (int q = 0; q < 4; ++q) v[r][q] = drow[32 * q];
      }
      // Columns are visited in strictly ascending order for every row (tiles, then 32-column groups,
      // then lanes), so "beats the current k-th" is the strict test d < tau_d and ties keep the lower
      // index by construction (tf.nn.top_k's rule).
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        unsigned m[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) m[r] = __ballot_sync(0xffffffffu, v[r][q] < taud[r]);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          if (m[r] == 0u) continue;      // warp-uniform
          unsigned mm = m[r];
          do {
            const int src = __ffs(mm) - 1;
            mm &= mm - 1u;
            const float cd = __shfl_sync(0xffffffffu, v[r][q], src);
            list_insert_ordered<KSLOT>(ld[r], li[r], cd, col0 + 32 * q + src, lane);
          } while (mm);
          taud[r] = list_tau_d<KSLOT>(ld[r], k);
        }
      }
    }
  }

  if (MODE == 0) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int row = i0 + warp * 8 + r;
      if (row < N) {
#pragma unroll
        for (int s = 0; s < KSLOT; ++s) {
          const int e = s * 32 + lane;
          if (e < k) {
            const size_t o = ((size_t)b * N + row) * k + e;
            idx_out[o] = li[r][s];
            if (dist_out) dist_out[o] = ld[r][s];
          }
        }
      }
    }
  }
}

// ------------------------------------------------------- unfused top-k ------
template <int KSLOT>
__global__ void __launch_bounds__(256)
topk_rows_kernel(const float* __restrict__ adj, long long rows, int ncols, int k,
                 int32_t* __restrict__ idx_out, float* __restrict__ val_out) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* a = adj + (size_t)row * ncols;
  float ld[KSLOT];
  int li[KSLOT];
#pragma unroll
  for (int s = 0; s < KSLOT; ++s) { ld[s] = CUDART_INF_F; li[s] = INT_MAX; }
  for (int base = 0; base < ncols; base += 32) {
    const int j = base + lane;
    const float v = (j < ncols) ? a[j] : CUDART_INF_F;
    float td; int ti;
    list_tau<KSLOT>(ld, li, k, td, ti);
    const bool pass = (j < ncols) && ((v < td) || (v == td && j < ti));
    unsigned m = __ballot_sync(0xffffffffu, pass);
    while (m) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      const float cd = __shfl_sync(0xffffffffu, v, src);
      list_insert<KSLOT>(ld, li, cd, base + src, lane);
    }
  }
#pragma unroll
  for (int s = 0; s < KSLOT; ++s) {
    const int e = s * 32 + lane;
    if (e < k) {
      idx_out[(size_t)row * k + e] = li[s];
      if (val_out) val_out[(size_t)row * k + e] = ld[s];
    }
  }
}

// ---------------------------------------------------------------- host ------
struct KnnPlan {
  int KC, Dp, Npad;
  size_t xT_bytes, sq_bytes;
};
KnnPlan make_plan(int B, int N, int D) {
  KnnPlan p;
  p.KC = (D <= 4) ? 4 : (D <= 8) ? 8 : (D <= 16) ? 16 : 32;
  p.Dp = (D + p.KC - 1) / p.KC * p.KC;
  p.Npad = (N + TN - 1) / TN * TN;
  p.xT_bytes = align_up((size_t)B * p.Dp * p.Npad * sizeof(float), 256);
  p.sq_bytes = align_up((size_t)B * p.Npad * sizeof(float), 256);
  return p;
}

template <int KC, int KSLOT, int MODE>
int launch_tile(const KnnPlan& p, int B, int N, int k, int flavour, const float* xT, const float* sq,
                int32_t* idx, float* dist, float* adj, cudaStream_t st) {
  const size_t smem = ((size_t)p.Dp * TM + (size_t)NSTAGE * KC * TN + (MODE == 0 ? TM * DLD : 0)) * sizeof(float) +
                      (NSTAGE + 1) * sizeof(uint64_t);
  auto kern = knn_tile_kernel<KC, KSLOT, MODE>;
  // set on every launch: the attribute is per device, and a host thread may drive more than one
  WSPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(p.Npad / TM, B);
  kern<<<grid, KNN_THREADS, smem, st>>>(xT, sq, N, p.Npad, p.Dp, k, flavour, idx, dist, adj);
  count_launch();
  WSPC_LAUNCH_CHECK("knn_tile_kernel");
  return WSPC_OK;
}

template <int MODE>
int dispatch_tile(const KnnPlan& p, int B, int N, int k, int flavour, const float* xT, const float* sq,
                  int32_t* idx, float* dist, float* adj, cudaStream_t st) {
  const int kslot = (MODE == 0 && k > 32) ? 2 : 1;
#define WSPC_KNN_CASE(KC_)                                                                         \
  case KC_:                                                                                        \
    return kslot == 2 ? launch_tile<KC_, 2, MODE>(p, B, N, k, flavour, xT, sq, idx, dist, adj, st) \
                      : launch_tile<KC_, 1, MODE>(p, B, N, k, flavour, xT, sq, idx, dist, adj, st);
  switch (p.KC) {
    WSPC_KNN_CASE(4)
    WSPC_KNN_CASE(8)
    WSPC_KNN_CASE(16)
    WSPC_KNN_CASE(32)
  }
#undef WSPC_KNN_CASE
  set_error("knn: unsupported chunk %d", p.KC);
  return WSPC_ERR_INVALID;
}

int run_prep(const KnnPlan& p, const float* x, int B, int N, int ldx, int coff, int D, float* xT, float* sq,
             cudaStream_t st) {
  dim3 grid((p.Npad + 255) / 256, B);
  knn_prep_kernel<<<grid, 256, 0, st>>>(x, N, ldx, coff, D, p.Dp, p.Npad, xT, sq);
  count_launch();
  WSPC_LAUNCH_CHECK("knn_prep_kernel");
  return WSPC_OK;
}

int validate_common(const float* x, int B, int N, int ldx, int coff, int D, const void* ws) {
  WSPC_REQUIRE(x && ws, "knn: null pointer");
  WSPC_REQUIRE(B >= 1 && N >= 1, "knn: bad shape B=%d N=%d", B, N);
  WSPC_REQUIRE(D >= 1 && D <= 128, "knn: D=%d outside [1,128]", D);
  WSPC_REQUIRE(coff >= 0 && coff + D <= ldx, "knn: channel window [%d,%d) outside ldx=%d", coff, coff + D, ldx);
  WSPC_REQUIRE(aligned16(ws), "knn: workspace not 16-byte aligned");
  return WSPC_OK;
}

}  // namespace
}  // namespace wspc

using namespace wspc;

extern "C" size_t wspc_knn_workspace_bytes(int B, int N, int D) {
  if (B < 1 || N < 1 || D < 1 || D > 128) return 0;
  const KnnPlan p = make_plan(B, N, D);
  const size_t exact = p.xT_bytes + p.sq_bytes;
  const size_t tc = knn_tc_eligible(N, D, 1) ? knn_tc_workspace_bytes(B, N, D) : 0;
  return exact > tc ? exact : tc;
}

extern "C" int wspc_knn_fallback_rows(const void* workspace, int B, int N, int D, int* rows_out) {
  WSPC_REQUIRE(workspace && rows_out, "knn_fallback_rows: null pointer");
  if (!knn_tc_eligible(N, D, 1) || g_knn_path != 0) { *rows_out = 0; return WSPC_OK; }
  return knn_tc_fallback_rows(workspace, B, N, D, rows_out);
}

extern "C" int wspc_set_knn_path(int path) {
  const int old = wspc::g_knn_path;
  wspc::g_knn_path = path;
  return old;
}

extern "C" int wspc_knn_fused(const float* x, int B, int N, int ldx, int coff, int D, int k, int flavour,
                              int32_t* idx, float* dist, void* workspace, size_t workspace_bytes,
                              wspc_stream_t stream) {
  if (int rc = check_arch()) return rc;
  if (int rc = validate_common(x, B, N, ldx, coff, D, workspace)) return rc;
  WSPC_REQUIRE(idx, "knn_fused: idx is null");
  WSPC_REQUIRE(k >= 1 && k <= 64 && k <= N, "knn_fused: k=%d outside [1,min(64,N=%d)]", k, N);
  WSPC_REQUIRE(flavour == WSPC_DIST_TFUTIL || flavour == WSPC_DIST_SMOOTH, "knn_fused: bad flavour %d", flavour);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (g_knn_path == 0 && knn_tc_eligible(N, D, k))
    return knn_tc_run(x, B, N, ldx, coff, D, k, flavour, idx, dist, workspace, workspace_bytes, st);
  const KnnPlan p = make_plan(B, N, D);
  if (workspace_bytes < p.xT_bytes + p.sq_bytes) {
    set_error("knn_fused: workspace %zu < required %zu", workspace_bytes, p.xT_bytes + p.sq_bytes);
    return WSPC_ERR_WORKSPACE;
  }
  float* xT = static_cast<float*>(workspace);
  float* sq = reinterpret_cast<float*>(static_cast<char*>(workspace) + p.xT_bytes);
  if (int rc = run_prep(p, x, B, N, ldx, coff, D, xT, sq, st)) return rc;
  if (int rc = dispatch_tile<0>(p, B, N, k, flavour, xT, sq, idx, dist, nullptr, st)) return rc;
  if (knn_tc_eligible(N, D, 1) && workspace_bytes >= knn_tc_workspace_bytes(B, N, D))
    knn_tc_clear_fallback_rows(workspace, B, N, D, st);   // wspc_knn_fallback_rows stays meaningful on this path
  return WSPC_OK;
}

extern "C" int wspc_pairwise_distance(const float* x, int B, int N, int ldx, int coff, int D, int flavour,
                                      float* adj, void* workspace, size_t workspace_bytes,
                                      wspc_stream_t stream) {
  if (int rc = check_arch()) return rc;
  if (int rc = validate_common(x, B, N, ldx, coff, D, workspace)) return rc;
  WSPC_REQUIRE(adj, "pairwise_distance: adj is null");
  WSPC_REQUIRE(flavour == WSPC_DIST_TFUTIL || flavour == WSPC_DIST_SMOOTH, "pairwise_distance: bad flavour %d",
               flavour);
  const KnnPlan p = make_plan(B, N, D);
  if (workspace_bytes < p.xT_bytes + p.sq_bytes) {
    set_error("pairwise_distance: workspace %zu < required %zu", workspace_bytes, p.xT_bytes + p.sq_bytes);
    return WSPC_ERR_WORKSPACE;
  }
  float* xT = static_cast<float*>(workspace);
  float* sq = reinterpret_cast<float*>(static_cast<char*>(workspace) + p.xT_bytes);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (int rc = run_prep(p, x, B, N, ldx, coff, D, xT, sq, st)) return rc;
  return dispatch_tile<1>(p, B, N, 1, flavour, xT, sq, nullptr, nullptr, adj, st);
}

extern "C" int wspc_topk_rows(const float* adj, long long rows, int ncols, int k, int32_t* idx, float* vals,
                              wspc_stream_t stream) {
  if (int rc = check_arch()) return rc;
  WSPC_REQUIRE(adj && idx, "topk_rows: null pointer");
  WSPC_REQUIRE(rows >= 1 && ncols >= 1, "topk_rows: bad shape");
  WSPC_REQUIRE(k >= 1 && k <= 64 && k <= ncols, "topk_rows: k=%d outside [1,min(64,ncols=%d)]", k, ncols);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int wpb = 8;
  const unsigned grid = (unsigned)((rows + wpb - 1) / wpb);
  if (k > 32) topk_rows_kernel<2><<<grid, wpb * 32, 0, st>>>(adj, rows, ncols, k, idx, vals);
  else        topk_rows_kernel<1><<<grid, wpb * 32, 0, st>>>(adj, rows, ncols, k, idx, vals);
  count_launch();
  WSPC_LAUNCH_CHECK("topk_rows_kernel");
  return WSPC_OK;
}
