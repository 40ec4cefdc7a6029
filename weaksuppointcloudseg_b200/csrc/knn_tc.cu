// Tensor-core kNN (D <= 64, k <= 64): tcgen05 distances, two-pass threshold selection, exact fp32 re-scoring.
//
// The canonical (bit-exact) distance arithmetic is a sequential fp32 FMA chain per pair, which
// caps the CUDA-core kernel in knn.cu at the FP32 pipe, and any per-element sorted-list selection costs tens of
// instructions per surviving candidate.  Here the N x N products run on the 5th-gen tensor cores, the selection
// needs ~1 instruction per matrix element, and exactness is restored afterwards:
//
//   prep   : split every feature into bf16 hi + lo and lay 128-point tiles out in the UMMA K-major core-matrix
//            image ([channel group][row][8]; hi groups then lo groups) so that one 1-D bulk copy (UBLKCP) stages a
//            whole operand tile.  The approximate pass works on coordinates CENTRED on the cloud mean (distances are
//            translation invariant; the bf16 split error scales with the norms, so clouds far from the origin would
//            otherwise get a margin wider than their neighbour spacing); the exact pass uses the original data.  Three extra "hi" channels carry a 3-way bf16 split of -sq_j/2 (canonical row
//            norm); the query-side copy of the tile has those channels patched to 1, so the accumulator holds
//            v_ij = x_i.x_j - sq_j/2 directly and d~_ij = sq_i - 2 v_ij: the largest v of a row are its nearest.
//   main   : CTA = 256 query rows (two M=128 accumulators, double buffered = all 512 TMEM columns), 16 selection
//            warps + 1 producer warp (one thread issues the bulk copies and the tcgen05.mma's:
//            hi*hi + lo*hi + hi*lo, fp32 accumulate).  Thread = (row, 64-column half of the tile).
//            pass 1: every thread keeps the running maximum of each of its 64 column classes (column mod 128) in
//                    registers -- one FMNMX per element.  The k-th largest of a row's 128 class maxima, v_k, is
//                    reached by k distinct columns, so the exact k-th distance is <= d~(v_k) + eps.
//            pass 2: the MMAs are replayed and every column with v >= v_k - eps_i is appended to the row's
//                    candidate list (<= 32 entries).  eps_i bounds |d~ - d_exact| for every pair of row i:
//                    eps_i = 2^-13 sqrt(sq'_i smax') + 2^-18 (sq'_i + smax')      [centred norms: bf16x3 split error
//                            <= 6*2^-18 |x'||y'| = 2^-15.4, tensor-core fp32 accumulation, centring roundings]
//                          + (D+3) 2^-23 (|x_i| + sqrt(smax))^2                    [worst-case rounding of the canonical
//                            fp32 chain sq_i - 2 dot + sq_j on the ORIGINAL coordinates], hence
//                    every exact top-k member t has d~_t <= d_t + eps <= d~(v_k) + 2 eps, i.e. v_t >= v_k - eps.
//                    More than 32 candidates (heavy ties / clustered classes) flags the row.
//   refine : each lane recomputes the canonical fp32 distance of its candidate (same fmaf chain as the oracle),
//            the warp sorts the <= 32 (d, j) pairs lexicographically and writes the first k: bit-exact indices and
//            distances.  d <= d~(threshold) + eps is verified on every candidate; a violation flags the row.
//   fallback: flagged rows are recomputed by a warp-per-row exact kernel (device-side list, no host sync).
#include "common.cuh"
#include <cuda_bf16.h>
#include <math_constants.h>
#include <limits.h>

namespace wspc {
void count_launch(int n = 1);
namespace {

constexpr int QT = 128;                    // points per operand tile (MMA M and N)
constexpr int RB = 256;                    // query rows per CTA (two M blocks)
constexpr int GROUP_BYTES = QT * 16 + 16;  // one 8-channel group of a tile image (padded, see gemm_tc.cu)
constexpr int SEL_WARPS = 16;              // selection warps: (TMEM lane quadrant, M block, column half)
constexpr int KTC_THREADS = (SEL_WARPS + 1) * 32;
constexpr int MAXC = 32;                   // candidate slots per row for k <= 24 (one per lane in the refine phase)
constexpr int MAXC_BIG = 96;               // ... for 24 < k <= 64 (three per lane, ranked instead of sorted)
constexpr int KSORT = 24;                  // largest k of the register-sort threshold / one-candidate-per-lane refine path
constexpr int MAX_NST = 6;                 // candidate-tile ring depth (chosen per D from the smem budget)
constexpr int NAUG = 3;                    // extra hi channels: 3-way bf16 split of -sq/2
constexpr int SMEM_BUDGET = 227 * 1024;

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(d_tmem),
      "l"(ad), "l"(bd), "r"(idesc), "r"(accum), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ float exact_dist(int flavour, float sqi, float sqj, float dot) {
  if (flavour == WSPC_DIST_TFUTIL) return __fadd_rn(__fadd_rn(sqi, __fmul_rn(-2.f, dot)), sqj);
  const float d = __fsub_rn(__fadd_rn(sqi, sqj), __fmul_rn(2.f, dot));
  return d > 0.f ? d : 0.f;
}

__host__ __device__ inline int tc_maxc(int k) { return k <= KSORT ? MAXC : MAXC_BIG; }
__host__ __device__ inline size_t tc_scratch_bytes(int k) {
  const size_t kk = (size_t)(k < 1 ? 1 : k);
  if (k <= KSORT) return kk * 2 * RB * 4 > (size_t)RB * MAXC * 2 ? kk * 2 * RB * 4 : (size_t)RB * MAXC * 2;
  return (size_t)RB * MAXC_BIG * 2 + (size_t)SEL_WARPS * MAXC_BIG * 8;
}

struct TcPlan {
  int Dp, Dl, Npad, ntile, ghi, glo, nst;
  uint32_t tile_bytes;
  size_t img_bytes, smem;
};
__host__ __device__ inline int round16(int v) { return (v + 15) / 16 * 16; }
TcPlan make_tc_plan(int B, int N, int D, int k) {
  TcPlan p;
  p.Dp = round16(D + NAUG);
  p.Dl = round16(D);
  p.Npad = (N + QT - 1) / QT * QT;
  p.ntile = p.Npad / QT;
  p.ghi = p.Dp / 8;
  p.glo = p.Dl / 8;
  p.tile_bytes = (uint32_t)(p.ghi + p.glo) * GROUP_BYTES;
  p.img_bytes = align_up((size_t)B * p.ntile * p.tile_bytes, 256);
  // selection scratch.  k <= 24: top-k exchange [k][2][RB] fp32, later aliased by the candidate lists [RB][MAXC] u16.
  // k > 24: count exchange [2][2][RB] int (bisection), aliased by the lists [RB][MAXC_BIG] u16, + per-warp refine scratch
  const size_t scratch = tc_scratch_bytes(k);
  const size_t fixed = 2 * (size_t)p.tile_bytes + scratch + RB * 4 /*cnt*/ + RB * 4 /*thr*/ + 256 /*barriers*/;
  long nst = ((long)SMEM_BUDGET - (long)fixed) / (long)p.tile_bytes;
  if (nst > MAX_NST) nst = MAX_NST;
  p.nst = (int)nst;
  p.smem = fixed + (size_t)(p.nst > 0 ? p.nst : 0) * p.tile_bytes;
  if (p.smem < 120 * 1024) p.smem = 120 * 1024;   // one CTA per SM: each CTA owns all 512 TMEM columns
  return p;
}

// ------------------------------------------------------------------ prep ---
// per-cloud centre = channel sums of the window (divided by N in the prep kernel; any vector would do, it only
// conditions the approximate pass).  grid (8, B), block 256: lane = channel, warps stride over the rows of the chunk.
__global__ void __launch_bounds__(256)
knn_tc_centre_kernel(const float* __restrict__ x, int N, int ldx, int coff, int D, float* __restrict__ centre_sum) {
  __shared__ float red[8][64];
  const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int per = (N + gridDim.x - 1) / gridDim.x;
  const int n0 = blockIdx.x * per, n1 = min(N, n0 + per);
  const float* xb = x + (size_t)b * N * ldx + coff;
  for (int c = lane; c < 64; c += 32) {
    float s = 0.f;
    if (c < D)
      for (int n = n0 + warp; n < n1; n += 8) s += xb[(size_t)n * ldx + c];
    red[warp][c] = s;
  }
  __syncthreads();
  if (threadIdx.x < D) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    atomicAdd(centre_sum + b * 64 + threadIdx.x, s);
  }
}

// grid (Npad/128, B), block 128: thread = point.  img per (cloud, tile): [hi: ghi groups][lo: glo groups] x [128][8] bf16
__global__ void __launch_bounds__(128)
knn_tc_prep_kernel(const float* __restrict__ x, int N, int ldx, int coff, int D, int ghi, int glo, int Npad,
                   const float* __restrict__ centre, unsigned char* __restrict__ img, float* __restrict__ sq,
                   float* __restrict__ sqc, unsigned* __restrict__ smax_bits) {
  const int b = blockIdx.y, tile = blockIdx.x, r = threadIdx.x;
  const int n = tile * QT + r;
  const size_t tile_bytes = (size_t)(ghi + glo) * GROUP_BYTES;
  unsigned char* base = img + ((size_t)b * (Npad / QT) + tile) * tile_bytes;
  // the tile's rows are fetched coalesced (float4, 16 lanes per 64-channel row) into shared memory; thread = point then reads
  // its own row from there (pitch D+1: conflict-free).  Row-per-thread global loads ran this kernel at 0.5 TB/s.
  __shared__ float sx[QT * 65];
  const bool staged = (D % 4 == 0) && (ldx % 4 == 0) && (coff % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15u) == 0);
  if (staged) {
    const int d4 = D >> 2;
    for (int e = r; e < QT * d4; e += QT) {
      const int rr = e / d4, c4 = e - rr * d4;
      const int nn = tile * QT + rr;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (nn < N) v = *reinterpret_cast<const float4*>(x + ((size_t)b * N + nn) * ldx + coff + c4 * 4);
      float* d = sx + rr * (D + 1) + c4 * 4;
      d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
    __syncthreads();
  }
  const float* xr = staged ? sx + r * (D + 1) : x + ((size_t)b * N + (n < N ? n : 0)) * ldx + coff;
  const float* cb = centre + b * 64;       // channel sums
  const float inv_n = 1.f / (float)N;
  float acc = 0.f, accc = 0.f;
  for (int c = 0; c < D; ++c) {             // canonical chain on the original data; centred norm for the approximation
    const float v = n < N ? xr[c] : 0.f;
    acc = __fmaf_rn(v, v, acc);
    const float vc = n < N ? v - cb[c] * inv_n : 0.f;
    accc = __fmaf_rn(vc, vc, accc);
  }
  // 3-way bf16 split of -sq'/2 (exact to 2^-24); padded points get a huge negative value: they never win
  const float s = n < N ? -0.5f * accc : -1.0e30f;
  const __nv_bfloat16 a1 = __float2bfloat16_rn(s);
  const float r1 = s - __bfloat162float(a1);
  const __nv_bfloat16 a2 = __float2bfloat16_rn(r1);
  const __nv_bfloat16 a3 = __float2bfloat16_rn(r1 - __bfloat162float(a2));
  for (int g = 0; g < ghi; ++g) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      unsigned short hh[2], ll[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = g * 8 + 2 * i + e;
        const float v = (n < N && c < D) ? xr[c] - cb[c] * inv_n : 0.f;
        __nv_bfloat16 hb = __float2bfloat16_rn(v);
        const __nv_bfloat16 lb = __float2bfloat16_rn(v - __bfloat162float(hb));
        if (c == D) hb = a1;
        if (c == D + 1) hb = a2;
        if (c == D + 2) hb = a3;
        hh[e] = __bfloat16_as_ushort(hb);
        ll[e] = __bfloat16_as_ushort(lb);
      }
      h[i] = (uint32_t)hh[0] | ((uint32_t)hh[1] << 16);
      l[i] = (uint32_t)ll[0] | ((uint32_t)ll[1] << 16);
    }
    *reinterpret_cast<uint4*>(base + (size_t)g * GROUP_BYTES + r * 16) = make_uint4(h[0], h[1], h[2], h[3]);
    if (g < glo)
      *reinterpret_cast<uint4*>(base + (size_t)(ghi + g) * GROUP_BYTES + r * 16) = make_uint4(l[0], l[1], l[2], l[3]);
  }
  if (r == 0) {   // the 16 pad bytes of every group are copied by the bulk copy: keep them defined
    for (int g = 0; g < ghi + glo; ++g) *reinterpret_cast<uint4*>(base + (size_t)g * GROUP_BYTES + QT * 16) = make_uint4(0, 0, 0, 0);
  }
  sq[(size_t)b * Npad + n] = acc;
  sqc[(size_t)b * Npad + n] = accc;
  // per-cloud maximum norms (norms >= 0: uint order == float order): one atomic per warp instead of one per point
  const unsigned m0 = __reduce_max_sync(0xffffffffu, n < N ? __float_as_uint(acc) : 0u);
  const unsigned m1 = __reduce_max_sync(0xffffffffu, n < N ? __float_as_uint(accc) : 0u);
  if ((r & 31) == 0) {
    atomicMax(smax_bits + 2 * b, m0);
    atomicMax(smax_bits + 2 * b + 1, m1);
  }
}

// error model shared by the selection and the refine phase (see the header comment)
__device__ __forceinline__ float knn_eps(float sq_i, float smax, float sqc_i, float smaxc, int D) {
  const float t = sqrtf(sq_i) + sqrtf(smax);
  return 1.2207e-4f * sqrtf(sqc_i * smaxc) + 3.8147e-6f * (sqc_i + smaxc) + (float)(D + 3) * 1.1921e-7f * t * t;
}

// ---------------------------------------------------------- register lists ---
__device__ __forceinline__ void list_insert32(float& ld, int& li, float cd, int cj, int lane) {
  const int pos = __popc(__ballot_sync(0xffffffffu, ld <= cd));
  const float up_d = __shfl_up_sync(0xffffffffu, ld, 1);
  const int up_i = __shfl_up_sync(0xffffffffu, li, 1);
  const bool gt = lane > pos;
  ld = gt ? up_d : ld;
  li = gt ? up_i : li;
  if (lane == pos) { ld = cd; li = cj; }
}

__device__ __forceinline__ bool lex_less(float da, int ja, float db, int jb) {
  return (da < db) || (da == db && ja < jb);
}
// ascending bitonic sort of one (d, j) pair per lane
__device__ __forceinline__ void warp_sort_pairs(float& d, int& j, int lane) {
#pragma unroll
  for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      const float od = __shfl_xor_sync(0xffffffffu, d, stride);
      const int oj = __shfl_xor_sync(0xffffffffu, j, stride);
      const bool up = (lane & size) == 0;
      const bool lower = (lane & stride) == 0;
      const bool other_less = lex_less(od, oj, d, j);
      const bool take = (lower == up) ? other_less : !other_less && !(od == d && oj == j);
      if (take) { d = od; j = oj; }
    }
  }
}

// descending bitonic sorting network over 64 registers (fully unrolled: every index is a compile-time constant)
__device__ __forceinline__ void sort64_desc(float (&a)[64]) {
#pragma unroll
  for (int kk = 2; kk <= 64; kk <<= 1) {
#pragma unroll
    for (int j = kk >> 1; j > 0; j >>= 1) {
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        const int l = i ^ j;
        if (l > i) {
          const float hi = fmaxf(a[i], a[l]), lo = fminf(a[i], a[l]);
          const bool desc = (i & kk) == 0;
          a[i] = desc ? hi : lo;
          a[l] = desc ? lo : hi;
        }
      }
    }
  }
}

// ------------------------------------------------------------------ main ---
template <bool BIGK>
__global__ void __launch_bounds__(KTC_THREADS, 1)
knn_tc_kernel(const unsigned char* __restrict__ img, const float* __restrict__ sq, const float* __restrict__ sqc,
              const unsigned* __restrict__ smax_bits, const float* __restrict__ x, int N, int Npad, int ldx, int coff, int D, int ghi, int glo, int nst, int k,
              int flavour, int32_t* __restrict__ idx_out, float* __restrict__ dist_out, int* __restrict__ flag_count,
              int* __restrict__ flag_rows) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t tile_bytes = (uint32_t)(ghi + glo) * GROUP_BYTES;
  const size_t scratch_bytes = tc_scratch_bytes(k);
  constexpr int maxc = BIGK ? MAXC_BIG : MAXC;
  unsigned char* sA = smem;                                       // two query tile images
  unsigned char* sB = sA + 2 * (size_t)tile_bytes;                // nst candidate tile images
  float* exch = reinterpret_cast<float*>(sB + (size_t)nst * tile_bytes);            // [k][2][RB] top-k exchange
  unsigned short* cand = reinterpret_cast<unsigned short*>(exch);                   // [RB][MAXC], aliases exch
  int* cnt = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(exch) + scratch_bytes);   // [RB]
  float* thr_s = reinterpret_cast<float*>(cnt + RB);                                // [RB] pass threshold (v space)
  uint64_t* bars = reinterpret_cast<uint64_t*>(thr_s + RB);
  uint64_t* full = bars;                    // [MAX_NST] candidate tile landed
  uint64_t* bfree = bars + MAX_NST;         // [MAX_NST] MMAs reading the stage retired
  uint64_t* accf = bars + 2 * MAX_NST;      // [2] accumulator pair ready
  uint64_t* acce = accf + 2;                // [2] accumulator pair drained by the 16 selection warps
  uint64_t* abar = acce + 2;                // query tiles landed
  uint64_t* aready = abar + 1;              // query tiles patched
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aready + 1);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.y, i0 = blockIdx.x * RB;
  const int ntile = Npad / QT;
  const int nm = (blockIdx.x * 2 + 1 < ntile) ? 2 : 1;            // M blocks present in this row block
  const unsigned char* imgb = img + (size_t)b * ntile * tile_bytes;
  const float* sqb = sq + (size_t)b * Npad;
  const float* sqcb = sqc + (size_t)b * Npad;
  const int total = 2 * ntile;                                    // tile visits: pass 1 then pass 2

  if (warp == SEL_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int s = 0; s < 2 * MAX_NST; ++s) mbar_init(&bars[s], 1);
    mbar_init(&accf[0], 1);
    mbar_init(&accf[1], 1);
    mbar_init(&acce[0], SEL_WARPS);
    mbar_init(&acce[1], SEL_WARPS);
    mbar_init(abar, 1);
    mbar_init(aready, RB);
    mbar_fence_init();
  }
  if (tid < RB) cnt[tid] = 0;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == SEL_WARPS) {
    // ================================================================ producer: bulk copies + MMA issue
    if (lane == 0) {
      auto issue_load = [&](int u) {
        const int s = u % nst, t = u % ntile;
        mbar_expect_tx(&full[s], tile_bytes);
        bulk_g2s(sB + (size_t)s * tile_bytes, imgb + (size_t)t * tile_bytes, tile_bytes, &full[s]);
      };
      mbar_expect_tx(abar, (uint32_t)nm * tile_bytes);
      for (int m = 0; m < nm; ++m)
        bulk_g2s(sA + (size_t)m * tile_bytes, imgb + (size_t)(blockIdx.x * 2 + m) * tile_bytes, tile_bytes, abar);
      for (int u = 0; u < nst && u < total; ++u) issue_load(u);
      // instruction descriptor: D=f32, A=B=bf16, K-major both, M=128, N=128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(QT >> 3) << 17) | ((uint32_t)(QT >> 4) << 24);
      const int khi = ghi / 2, klo = glo / 2;                     // K=16 steps per pass
      mbar_wait(aready, 0);
      for (int u = 0; u < total; ++u) {
        const int s = u % nst, a = u & 1;
        mbar_wait(&full[s], (u / nst) & 1);
        if (u >= 2) mbar_wait(&acce[a], ((u >> 1) - 1) & 1);
        tc_fence_after();
        const uint32_t b_hi = smem_u32(sB + (size_t)s * tile_bytes), b_lo = b_hi + (uint32_t)ghi * GROUP_BYTES;
        for (int m = 0; m < nm; ++m) {
          const uint32_t a_hi = smem_u32(sA + (size_t)m * tile_bytes), a_lo = a_hi + (uint32_t)ghi * GROUP_BYTES;
          const uint32_t acc = tmem_base + (uint32_t)(a * 2 + m) * QT;
          uint32_t accum = 0;
          for (int kk = 0; kk < khi; ++kk) {                       // hi * hi (includes the -sq_j/2 channels)
            tc_mma(acc, umma_desc(a_hi + (uint32_t)(2 * kk) * GROUP_BYTES, GROUP_BYTES, 128),
                   umma_desc(b_hi + (uint32_t)(2 * kk) * GROUP_BYTES, GROUP_BYTES, 128), idesc, accum);
            accum = 1;
          }
          for (int kk = 0; kk < klo; ++kk)                         // lo * hi
            tc_mma(acc, umma_desc(a_lo + (uint32_t)(2 * kk) * GROUP_BYTES, GROUP_BYTES, 128),
                   umma_desc(b_hi + (uint32_t)(2 * kk) * GROUP_BYTES, GROUP_BYTES, 128), idesc, 1u);
          for (int kk = 0; kk < klo; ++kk)                         // hi * lo
            tc_mma(acc, umma_desc(a_hi + (uint32_t)(2 * kk) * GROUP_BYTES, GROUP_BYTES, 128),
                   umma_desc(b_lo + (uint32_t)(2 * kk) * GROUP_BYTES, GROUP_BYTES, 128), idesc, 1u);
        }
        tc_commit(&accf[a]);
        tc_commit(&bfree[s]);
        if (u >= 1 && u - 1 + nst < total) {                      // refill the stage the previous visit used
          const int sp = (u - 1) % nst;
          mbar_wait(&bfree[sp], ((u - 1) / nst) & 1);
          issue_load(u - 1 + nst);
        }
      }
    }
  } else {
    // ================================================================ selection warps
    const int q = warp & 3, m = (warp >> 2) & 1, h = warp >> 3;
    const int rloc = m * QT + q * 32 + lane;                      // row within the CTA (TMEM lane q*32+lane of block m)
    const int row = i0 + rloc;
    const bool m_ok = m < nm;
    const bool row_ok = row < N;
    const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(m * QT + h * 64);

    // ---- patch the query-side copies: channels D..D+2 of the hi image become 1 (the candidate side keeps -sq/2)
    mbar_wait(abar, 0);
    if (h == 0 && m_ok) {
      unsigned char* rowp = sA + (size_t)m * tile_bytes + (size_t)(q * 32 + lane) * 16;
#pragma unroll
      for (int i = 0; i < NAUG; ++i) {
        const int c = D + i;
        *reinterpret_cast<unsigned short*>(rowp + (size_t)(c >> 3) * GROUP_BYTES + (c & 7) * 2) = 0x3F80;   // bf16 1.0
      }
      fence_proxy_async_smem();
    }
    if (h == 0) mbar_arrive(aready);

    const float smax = __uint_as_float(smax_bits[2 * b]), smaxc = __uint_as_float(smax_bits[2 * b + 1]);
    const float eps = knn_eps(sqb[m_ok ? row : i0], smax, sqcb[m_ok ? row : i0], smaxc, D);

    // ---- pass 1: running maximum of each column class
    float acc[64];
#pragma unroll
    for (int c = 0; c < 64; ++c) acc[c] = -CUDART_INF_F;
    for (int u = 0; u < ntile; ++u) {
      const int a = u & 1;
      mbar_wait(&accf[a], (u >> 1) & 1);
      tc_fence_after();
      if (m_ok) {
        float v[32];
        tc_ld32(tbase + (uint32_t)(a * 2 * QT), v);
#pragma unroll
        for (int c = 0; c < 32; ++c) acc[c] = fmaxf(acc[c], v[c]);
        tc_ld32(tbase + (uint32_t)(a * 2 * QT + 32), v);
#pragma unroll
        for (int c = 0; c < 32; ++c) acc[32 + c] = fmaxf(acc[32 + c], v[c]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acce[a]);
    }

    // ---- k-th largest class maximum of the row (both halves), pass threshold in v space
    float thr = CUDART_INF_F;
    if (!BIGK) {
      sort64_desc(acc);
#pragma unroll
      for (int s = 0; s < KSORT; ++s)
        if (s < k) exch[(s * 2 + h) * RB + rloc] = acc[s];
      named_bar_sync(1, SEL_WARPS * 32);
      int ia = 0, ib = 0;
      float vk = 0.f;
      for (int s = 0; s < k; ++s) {
        const float va = exch[(ia * 2 + 0) * RB + rloc], vb = exch[(ib * 2 + 1) * RB + rloc];
        if (va >= vb) { vk = va; ++ia; } else { vk = vb; ++ib; }
      }
      if (row_ok && m_ok) thr = vk - eps;
    } else {
      // No sort and no list exchange (64 values x 2 halves x 256 rows would not fit): bit-wise bisection on the order-preserving
      // integer image of the floats.  T grows from the top bit down while at least k of the row's 128 class maxima are >= T, so
      // it ends on the k-th largest value exactly; every step exchanges one count per half through shared memory.
      int* cx = reinterpret_cast<int*>(exch);                        // [2 parities][2 halves][RB]
      unsigned key[64];
#pragma unroll
      for (int c = 0; c < 64; ++c) {
        const unsigned u = __float_as_uint(acc[c]);
        key[c] = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
      }
      unsigned T = 0u;
#pragma unroll 1
      for (int bit = 31; bit >= 0; --bit) {
        const unsigned cand_t = T | (1u << bit);
        int n = 0;
#pragma unroll
        for (int c = 0; c < 64; ++c) n += key[c] >= cand_t ? 1 : 0;
        int* slot = cx + (bit & 1) * 2 * RB;
        slot[h * RB + rloc] = n;
        named_bar_sync(1, SEL_WARPS * 32);
        if (slot[rloc] + slot[RB + rloc] >= k) T = cand_t;
      }
      const float vk = __uint_as_float((T & 0x80000000u) ? (T & 0x7fffffffu) : ~T);
      if (row_ok && m_ok) thr = vk - eps;
    }
    if (h == 0) thr_s[rloc] = thr;
    named_bar_sync(1, SEL_WARPS * 32);     // every thread has read the exchange area: the candidate lists may alias it

    // ---- pass 2: collect every column whose v reaches the threshold
    for (int u = ntile; u < total; ++u) {
      const int a = u & 1;
      mbar_wait(&accf[a], (u >> 1) & 1);
      tc_fence_after();
      if (m_ok) {
        const int col0 = (u - ntile) * QT + h * 64;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float v[32];
          tc_ld32(tbase + (uint32_t)(a * 2 * QT + half * 32), v);
          // branch-free miss mask: the sign bit of (v - thr) is shifted in, element c lands on bit 31 - c
          // (v == thr gives +0: a hit; thr = +inf for inactive rows gives -inf or NaN(-inf - -inf never occurs): a miss)
          unsigned miss = 0u;
#pragma unroll
          for (int c = 0; c < 32; ++c) miss = __funnelshift_l(__float_as_uint(v[c] - thr), miss, 1);
          unsigned hits = ~miss;
          while (hits) {
            const int c = __clz(hits);
            hits &= ~(0x80000000u >> c);
            const int slot = atomicAdd(&cnt[rloc], 1);
            if (slot < maxc) cand[rloc * maxc + slot] = (unsigned short)(col0 + half * 32 + c);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acce[a]);
    }
    named_bar_sync(1, SEL_WARPS * 32);     // all candidate lists complete

    // ---- exact re-scoring: warp owns 16 rows, lane e owns candidate e
    const float* xb = x + (size_t)b * N * ldx + coff;
    const bool vec_ok = ((ldx & 3) == 0) && ((coff & 3) == 0) && ((D & 3) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
#pragma unroll 1
    for (int r = 0; r < RB / SEL_WARPS; ++r) {
      const int rl = warp * (RB / SEL_WARPS) + r;
      const int grow = i0 + rl;
      if (grow >= N) continue;                                    // warp-uniform
      const int c = cnt[rl];
      bool flag = (c > maxc) || (c < k);
      const float sq_r = sqb[grow], sqc_r = sqcb[grow];
      const float eps_r = knn_eps(sq_r, smax, sqc_r, smaxc, D);
      const float bound = (sqc_r - 2.f * thr_s[rl]) + 2.f * eps_r;  // d~ of the threshold + model error (+ rounding slack)
      if (BIGK) {
        // 24 < k <= 64: up to three candidates per lane.  Their canonical distances go to the warp's scratch; the position of a
        // candidate in the answer is its lexicographic (d, j) rank among the row's candidates (indices are distinct, so the
        // ranks are a permutation).
        float* sd = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(exch) + (size_t)RB * MAXC_BIG * 2) +
                    (size_t)warp * MAXC_BIG * 2;
        int* sj = reinterpret_cast<int*>(sd + MAXC_BIG);
        const int cc = c < maxc ? c : maxc;
        const float* xi = xb + (size_t)grow * ldx;
        float dm[MAXC_BIG / 32];
        int jm[MAXC_BIG / 32];
#pragma unroll
        for (int e = 0; e < MAXC_BIG / 32; ++e) {
          const int ci = lane + 32 * e;
          dm[e] = CUDART_INF_F;
          jm[e] = INT_MAX;
          if (ci < cc) {
            const int j = cand[rl * maxc + ci];
            const float* xj = xb + (size_t)j * ldx;
            float dot = 0.f;
            if (vec_ok) {
#pragma unroll 1
              for (int c0 = 0; c0 < D; c0 += 16) {
                float4 av[4], bq[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  if (c0 + 4 * u < D) {
                    av[u] = *reinterpret_cast<const float4*>(xi + c0 + 4 * u);
                    bq[u] = *reinterpret_cast<const float4*>(xj + c0 + 4 * u);
                  }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  if (c0 + 4 * u < D) {
                    dot = __fmaf_rn(av[u].x, bq[u].x, dot);
                    dot = __fmaf_rn(av[u].y, bq[u].y, dot);
                    dot = __fmaf_rn(av[u].z, bq[u].z, dot);
                    dot = __fmaf_rn(av[u].w, bq[u].w, dot);
                  }
                }
              }
            } else {
              for (int q2 = 0; q2 < D; ++q2) dot = __fmaf_rn(xi[q2], xj[q2], dot);
            }
            dm[e] = exact_dist(flavour, sq_r, sqb[j], dot);
            jm[e] = j;
            if (!(dm[e] <= bound)) flag = true;                      // error-model self check
            sd[ci] = dm[e];
            sj[ci] = j;
          }
        }
        flag = __any_sync(0xffffffffu, flag);
        __syncwarp();
#pragma unroll
        for (int e = 0; e < MAXC_BIG / 32; ++e) {
          if (lane + 32 * e < cc) {
            int rank = 0;
            for (int m2 = 0; m2 < cc; ++m2) rank += lex_less(sd[m2], sj[m2], dm[e], jm[e]) ? 1 : 0;
            if (rank < k) {
              const size_t o = ((size_t)b * N + grow) * k + rank;
              idx_out[o] = jm[e];
              if (dist_out) dist_out[o] = dm[e];
            }
          }
        }
        __syncwarp();                                                // the scratch is reused by the warp's next row
        if (flag && lane == 0) {
          const int slot = atomicAdd(flag_count, 1);
          flag_rows[slot] = b * N + grow;
        }
        continue;
      }
      float d = CUDART_INF_F;
      int jj = INT_MAX;
      if (lane < c && lane < MAXC) {
        const int j = cand[rl * maxc + lane];
        const float* xi = xb + (size_t)grow * ldx;
        const float* xj = xb + (size_t)j * ldx;
        float dot = 0.f;
        if (vec_ok) {   // 16-byte aligned rows: issue all loads first, then the canonical chain (c ascending from +0)
#pragma unroll 1
          for (int c0 = 0; c0 < D; c0 += 16) {
            float4 av[4], bq[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              if (c0 + 4 * u < D) {
                av[u] = *reinterpret_cast<const float4*>(xi + c0 + 4 * u);
                bq[u] = *reinterpret_cast<const float4*>(xj + c0 + 4 * u);
              }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              if (c0 + 4 * u < D) {
                dot = __fmaf_rn(av[u].x, bq[u].x, dot);
                dot = __fmaf_rn(av[u].y, bq[u].y, dot);
                dot = __fmaf_rn(av[u].z, bq[u].z, dot);
                dot = __fmaf_rn(av[u].w, bq[u].w, dot);
              }
            }
          }
        } else {
          for (int cc = 0; cc < D; ++cc) dot = __fmaf_rn(xi[cc], xj[cc], dot);
        }
        d = exact_dist(flavour, sq_r, sqb[j], dot);
        jj = j;
        if (!(d <= bound)) flag = true;                            // error-model self check
      }
      flag = __any_sync(0xffffffffu, flag);
      warp_sort_pairs(d, jj, lane);
      if (lane < k) {
        const size_t o = ((size_t)b * N + grow) * k + lane;
        idx_out[o] = jj == INT_MAX ? 0 : jj;
        if (dist_out) dist_out[o] = d;
      }
      if (flag && lane == 0) {
        const int slot = atomicAdd(flag_count, 1);
        flag_rows[slot] = b * N + grow;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == SEL_WARPS) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

// ---------------------------------------------------------------- fallback ---
// exact streaming kNN for the flagged rows: one CTA per row, each of the 8 warps scans an eighth of the columns
// (lane = candidate column within a 32-column group) and keeps its own sorted list; warp 0 merges the 8 lists in
// column order, so "equal distance -> lower index first" is preserved.
__global__ void __launch_bounds__(256)
knn_exact_rows_kernel(const float* __restrict__ x, const float* __restrict__ sq, int N, int Npad, int ldx, int coff, int D,
                      int k, int flavour, const int* __restrict__ flag_count, const int* __restrict__ flag_rows,
                      int32_t* __restrict__ idx_out, float* __restrict__ dist_out) {
  __shared__ float xs[64];
  __shared__ float sd[8][32];
  __shared__ int si[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nflag = *flag_count;
  const bool vec_ok = ((ldx & 3) == 0) && ((coff & 3) == 0) && ((D & 3) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  const int per = ((N + 7) / 8 + 31) / 32 * 32;       // columns per warp, a multiple of 32
  for (int f = blockIdx.x; f < nflag; f += gridDim.x) {
    const int grow = flag_rows[f];
    const int b = grow / N, row = grow - b * N;
    const float* xb = x + (size_t)b * N * ldx + coff;
    __syncthreads();                                   // previous row's shared state fully consumed
    if (threadIdx.x < D) xs[threadIdx.x] = xb[(size_t)row * ldx + threadIdx.x];
    __syncthreads();
    const float sqi = sq[(size_t)b * Npad + row];
    float ld = CUDART_INF_F;
    int li = INT_MAX;
    float tau = CUDART_INF_F;
    const int j0 = warp * per, j1 = min(N, j0 + per);
    for (int base = j0; base < j1; base += 32) {
      const int j = base + lane;
      float d = __int_as_float(0x7fc00000);
      if (j < j1) {
        const float* xj = xb + (size_t)j * ldx;
        float dot = 0.f;
        if (vec_ok) {
          for (int c0 = 0; c0 < D; c0 += 16) {
            float4 q[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (c0 + 4 * u < D) q[u] = *reinterpret_cast<const float4*>(xj + c0 + 4 * u);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              if (c0 + 4 * u < D) {
                dot = __fmaf_rn(xs[c0 + 4 * u + 0], q[u].x, dot);
                dot = __fmaf_rn(xs[c0 + 4 * u + 1], q[u].y, dot);
                dot = __fmaf_rn(xs[c0 + 4 * u + 2], q[u].z, dot);
                dot = __fmaf_rn(xs[c0 + 4 * u + 3], q[u].w, dot);
              }
            }
          }
        } else {
          for (int c = 0; c < D; ++c) dot = __fmaf_rn(xs[c], xj[c], dot);
        }
        d = exact_dist(flavour, sqi, sq[(size_t)b * Npad + j], dot);
      }
      unsigned m = __ballot_sync(0xffffffffu, d < tau);
      while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1u;
        list_insert32(ld, li, __shfl_sync(0xffffffffu, d, src), base + src, lane);
      }
      tau = __shfl_sync(0xffffffffu, ld, k - 1);
    }
    sd[warp][lane] = ld;
    si[warp][lane] = li;
    __syncthreads();
    if (warp == 0) {
      for (int w = 1; w < 8; ++w) {
        for (int e = 0; e < k; ++e) {
          const float cd = sd[w][e];
          if (!(cd < tau)) break;                      // that list is ascending: nothing further can enter
          list_insert32(ld, li, cd, si[w][e], lane);
          tau = __shfl_sync(0xffffffffu, ld, k - 1);
        }
      }
      if (lane < k) {
        const size_t o = (size_t)grow * k + lane;
        idx_out[o] = li;
        if (dist_out) dist_out[o] = ld;
      }
    }
  }
}

// the same for k > 32 (the register list above holds one entry per lane): one CTA per flagged row selects the answers one
// after the other -- the lexicographically smallest (d, j) above the previous one, canonical distances recomputed on the fly.
// Flagged rows are rare (rows with more than MAXC_BIG candidates or a violated error bound), so the k passes do not matter.
__global__ void __launch_bounds__(256)
knn_exact_rows_bigk_kernel(const float* __restrict__ x, const float* __restrict__ sq, int N, int Npad, int ldx, int coff, int D,
                           int k, int flavour, const int* __restrict__ flag_count, const int* __restrict__ flag_rows,
                           int32_t* __restrict__ idx_out, float* __restrict__ dist_out) {
  __shared__ float xs[64];
  __shared__ float rd[8];
  __shared__ int rj[8];
  __shared__ float best_d;
  __shared__ int best_j;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nflag = *flag_count;
  for (int f = blockIdx.x; f < nflag; f += gridDim.x) {
    const int grow = flag_rows[f];
    const int b = grow / N, row = grow - b * N;
    const float* xb = x + (size_t)b * N * ldx + coff;
    __syncthreads();
    if (threadIdx.x < D) xs[threadIdx.x] = xb[(size_t)row * ldx + threadIdx.x];
    __syncthreads();
    const float sqi = sq[(size_t)b * Npad + row];
    float pd = -CUDART_INF_F;
    int pj = -1;
    for (int sel = 0; sel < k; ++sel) {
      float md = CUDART_INF_F;
      int mj = INT_MAX;
      for (int j = threadIdx.x; j < N; j += 256) {
        const float* xj = xb + (size_t)j * ldx;
        float dot = 0.f;
        for (int c = 0; c < D; ++c) dot = __fmaf_rn(xs[c], xj[c], dot);
        const float d = exact_dist(flavour, sqi, sq[(size_t)b * Npad + j], dot);
        if (lex_less(pd, pj, d, j) && lex_less(d, j, md, mj)) { md = d; mj = j; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float od = __shfl_xor_sync(0xffffffffu, md, o);
        const int oj = __shfl_xor_sync(0xffffffffu, mj, o);
        if (lex_less(od, oj, md, mj)) { md = od; mj = oj; }
      }
      if (lane == 0) { rd[warp] = md; rj[warp] = mj; }
      __syncthreads();
      if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w)
          if (lex_less(rd[w], rj[w], md, mj)) { md = rd[w]; mj = rj[w]; }
        best_d = md;
        best_j = mj;
        const size_t o = (size_t)grow * k + sel;
        idx_out[o] = mj == INT_MAX ? 0 : mj;
        if (dist_out) dist_out[o] = md;
      }
      __syncthreads();
      pd = best_d;
      pj = best_j;
    }
  }
}

}  // namespace

// ---------------------------------------------------------------- host ------
bool knn_tc_eligible(int N, int D, int k) {
  if (!(D >= 1 && D <= 64 && k >= 1 && k <= 64 && N >= k && N <= 65535)) return false;
  return make_tc_plan(1, N, D, k).nst >= 2;
}

namespace {
struct TcWs { unsigned char* img; float* sq; float* sqc; float* centre; unsigned* smax; int* flag_count; int* flag_rows; size_t total; };
TcWs carve(void* ws, const TcPlan& p, int B, int N) {
  TcWs w;
  char* c = static_cast<char*>(ws);
  w.img = reinterpret_cast<unsigned char*>(c); c += p.img_bytes;
  w.sq = reinterpret_cast<float*>(c); c += align_up((size_t)B * p.Npad * 4, 256);
  w.sqc = reinterpret_cast<float*>(c); c += align_up((size_t)B * p.Npad * 4, 256);
  w.centre = reinterpret_cast<float*>(c); c += align_up((size_t)B * 64 * 4, 256);
  w.smax = reinterpret_cast<unsigned*>(c); c += align_up((size_t)B * 8, 256);
  w.flag_count = reinterpret_cast<int*>(c); c += 256;
  w.flag_rows = reinterpret_cast<int*>(c); c += align_up((size_t)B * N * 4, 256);
  w.total = (size_t)(c - static_cast<char*>(ws));
  return w;
}
}  // namespace

size_t knn_tc_workspace_bytes(int B, int N, int D) {
  const TcPlan p = make_tc_plan(B, N, D, 24);
  return carve(nullptr, p, B, N).total;
}

// telemetry: number of rows the last knn_tc_run on this workspace sent to the exact fallback (synchronises)
int knn_tc_fallback_rows(const void* ws, int B, int N, int D, int* out) {
  const TcPlan p = make_tc_plan(B, N, D, 24);
  const TcWs w = carve(const_cast<void*>(ws), p, B, N);
  WSPC_CUDA(cudaMemcpy(out, w.flag_count, sizeof(int), cudaMemcpyDeviceToHost));
  return WSPC_OK;
}

void knn_tc_clear_fallback_rows(void* ws, int B, int N, int D, cudaStream_t st) {
  const TcPlan p = make_tc_plan(B, N, D, 24);
  cudaMemsetAsync(carve(ws, p, B, N).flag_count, 0, sizeof(int), st);
}

int knn_tc_run(const float* x, int B, int N, int ldx, int coff, int D, int k, int flavour, int32_t* idx, float* dist,
               void* ws, size_t ws_bytes, cudaStream_t st) {
  const TcPlan p = make_tc_plan(B, N, D, k);
  if (ws_bytes < knn_tc_workspace_bytes(B, N, D)) {
    set_error("knn_fused: workspace %zu < required %zu", ws_bytes, knn_tc_workspace_bytes(B, N, D));
    return WSPC_ERR_WORKSPACE;
  }
  const TcWs w = carve(ws, p, B, N);
  // centre sums, max norms and the flagged-row counter are contiguous
  WSPC_CUDA(cudaMemsetAsync(w.centre, 0, align_up((size_t)B * 64 * 4, 256) + align_up((size_t)B * 8, 256) + 256, st));
  knn_tc_centre_kernel<<<dim3(8, B), 256, 0, st>>>(x, N, ldx, coff, D, w.centre);
  knn_tc_prep_kernel<<<dim3(p.ntile, B), 128, 0, st>>>(x, N, ldx, coff, D, p.ghi, p.glo, p.Npad, w.centre, w.img, w.sq, w.sqc,
                                                      w.smax);
  const int big = k > KSORT ? 1 : 0;
  auto kern = big ? knn_tc_kernel<true> : knn_tc_kernel<false>;
  // set on every launch: the attribute is per device, and a host thread may drive more than one
  WSPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
  kern<<<dim3((N + RB - 1) / RB, B), KTC_THREADS, p.smem, st>>>(w.img, w.sq, w.sqc, w.smax, x, N, p.Npad, ldx, coff, D, p.ghi, p.glo,
                                                               p.nst, k, flavour, idx, dist, w.flag_count, w.flag_rows);
  if (k <= 32)
    knn_exact_rows_kernel<<<2 * kNumSM, 256, 0, st>>>(x, w.sq, N, p.Npad, ldx, coff, D, k, flavour, w.flag_count, w.flag_rows,
                                                      idx, dist);
  else
    knn_exact_rows_bigk_kernel<<<2 * kNumSM, 256, 0, st>>>(x, w.sq, N, p.Npad, ldx, coff, D, k, flavour, w.flag_count,
                                                           w.flag_rows, idx, dist);
  count_launch(4);
  WSPC_LAUNCH_CHECK("knn_tc kernels");
  return WSPC_OK;
}

}  // namespace wspc
