// Tensor-core kNN for wide features (D >= 16): tcgen05 distances + exact fp32 re-scoring.
//
// The canonical (bit-exact) distance arithmetic is a sequential fp32 FMA chain per pair, which
// caps the CUDA-core kernel in knn.cu at the FP32 pipe.  Here the N x N dot products run on the 5th-gen tensor
// cores instead and exactness is restored afterwards:
//
//   prep   : split every feature into bf16 hi + lo and lay 128-point tiles out in the UMMA K-major core-matrix
//            image ([channel group][row][8], hi then lo) so that one 1-D bulk copy (UBLKCP) stages a whole operand
//            tile; row norms with the canonical chain; per-cloud max norm.
//   main   : CTA = 128 query rows.  For each 128-column tile: TMA ring -> 12 x tcgen05.mma (hi*hi + lo*hi + hi*lo,
//            fp32 accumulate in TMEM) -> tcgen05.ld -> d~ = (sq_i + sq_j) - 2 dot~ -> shared distance tile ->
//            per-row streaming selection of the 32 smallest d~ (register lists, one entry per lane).  The pass
//            threshold is d~_(k) + 2 eps_i, where eps_i bounds |d~ - d_exact| for every pair of row i
//            (eps_i = 2^-11.5 sqrt(sq_i smax) + 2^-20 (sq_i + smax): bf16x3 split error 2^-16 |x||y|, tensor-core
//            fp32 accumulation of 192 products, and the fp32 roundings of both formulas, with >4x head-room).
//            Every exact top-k member t satisfies d~_t <= d_t + eps <= d~_(k) + 2 eps, so it is in the list unless
//            more than 32 candidates fall inside the margin — detected (entry 31 inside the margin) and the row is
//            flagged.
//   refine : each lane recomputes the canonical fp32 distance of its candidate (same fmaf chain as the oracle),
//            the warp sorts the <= 32 (d, j) pairs lexicographically and writes the first k: bit-exact indices and
//            distances.  |d~ - d| <= eps is verified on every candidate; a violation flags the row.
//   fallback: flagged rows are recomputed by a warp-per-row exact kernel (device-side list, no host sync).
#include "common.cuh"
#include <cuda_bf16.h>
#include <math_constants.h>
#include <limits.h>

namespace wspc {
void count_launch(int n = 1);
namespace {

constexpr int QT = 128;                    // query rows / candidate columns per tile
constexpr int GROUP_BYTES = QT * 16 + 16;  // one 8-channel group of a tile image (padded, see gemm_tc.cu)
constexpr int DLD2 = QT + 4;
constexpr int KTC_THREADS = 1024;                 // 32 warps: 8 per scheduler hide the shuffle/ballot latency of the selection
constexpr int RW = QT / (KTC_THREADS / 32);       // rows owned by a warp in the selection phase
constexpr int ECOLS = QT / (KTC_THREADS / 128);   // columns read from TMEM by a warp in the epilogue
constexpr int NST = 2;                     // candidate-tile ring depth

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

__device__ __forceinline__ void tc_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ float exact_dist(int flavour, float sqi, float sqj, float dot) {
  if (flavour == WSPC_DIST_TFUTIL) return __fadd_rn(__fadd_rn(sqi, __fmul_rn(-2.f, dot)), sqj);
  const float d = __fsub_rn(__fadd_rn(sqi, sqj), __fmul_rn(2.f, dot));
  return d > 0.f ? d : 0.f;
}

// ------------------------------------------------------------------ prep ---
// grid (Npad/128, B), block 128: thread = point.  img: per (cloud, tile): [hi | lo] x [Dp/8 groups][128][8] bf16
__global__ void __launch_bounds__(128)
knn_tc_prep_kernel(const float* __restrict__ x, int N, int ldx, int coff, int D, int Dp, int Npad,
                   unsigned char* __restrict__ img, float* __restrict__ sq, unsigned* __restrict__ smax_bits) {
  const int b = blockIdx.y, tile = blockIdx.x, r = threadIdx.x;
  const int n = tile * QT + r;
  const int ngrp = Dp / 8;
  const size_t tile_bytes = (size_t)2 * ngrp * GROUP_BYTES;
  unsigned char* base = img + ((size_t)b * (Npad / QT) + tile) * tile_bytes;
  const float* xr = x + ((size_t)b * N + (n < N ? n : 0)) * ldx + coff;
  float acc = 0.f;
  for (int g = 0; g < ngrp; ++g) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float v0 = 0.f, v1 = 0.f;
      const int c = g * 8 + 2 * i;
      if (n < N && c < D) v0 = xr[c];
      if (n < N && c + 1 < D) v1 = xr[c + 1];
      acc = __fmaf_rn(v0, v0, acc);          // canonical chain (zeros beyond D do not change it)
      acc = __fmaf_rn(v1, v1, acc);
      const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
      const __nv_bfloat16 l0 = __float2bfloat16_rn(v0 - __bfloat162float(h0));
      const __nv_bfloat16 l1 = __float2bfloat16_rn(v1 - __bfloat162float(h1));
      h[i] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
      l[i] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    }
    *reinterpret_cast<uint4*>(base + (size_t)g * GROUP_BYTES + r * 16) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(base + (size_t)(ngrp + g) * GROUP_BYTES + r * 16) = make_uint4(l[0], l[1], l[2], l[3]);
  }
  if (r == 0) {   // the 16 pad bytes of every group are copied by the bulk copy: keep them defined
    for (int g = 0; g < 2 * ngrp; ++g) *reinterpret_cast<uint4*>(base + (size_t)g * GROUP_BYTES + QT * 16) = make_uint4(0, 0, 0, 0);
  }
  sq[(size_t)b * Npad + n] = acc;
  if (n < N) atomicMax(smax_bits + b, __float_as_uint(acc));   // acc >= 0: uint order == float order
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

// ------------------------------------------------------------------ main ---
__global__ void __launch_bounds__(KTC_THREADS, 1)
knn_tc_kernel(const unsigned char* __restrict__ img, const float* __restrict__ sq, const unsigned* __restrict__ smax_bits,
              const float* __restrict__ x, int N, int Npad, int ldx, int coff, int D, int Dp, int k, int flavour,
              int32_t* __restrict__ idx_out, float* __restrict__ dist_out, int* __restrict__ flag_count,
              int* __restrict__ flag_rows) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int ngrp = Dp / 8;
  const uint32_t tile_bytes = (uint32_t)2 * ngrp * GROUP_BYTES;
  unsigned char* sA = smem;                                  // query tile image (hi | lo)
  unsigned char* sB = sA + tile_bytes;                       // NST candidate tile images
  float* Ds = reinterpret_cast<float*>(sB + (size_t)NST * tile_bytes);   // [128][DLD2]
  float* sqB = Ds + QT * DLD2;                               // [NST][128] candidate norms
  uint64_t* bars = reinterpret_cast<uint64_t*>(sqB + NST * QT);   // full[NST], abar, mma_bar[2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NST + 3);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.y, i0 = blockIdx.x * QT;
  const int ntile = Npad / QT;
  const unsigned char* imgb = img + (size_t)b * ntile * tile_bytes;
  const float* sqb = sq + (size_t)b * Npad;

  auto issue = [&](int t) {   // thread 0: stage candidate tile t
    const int s = t % NST;
    mbar_expect_tx(&bars[s], tile_bytes + QT * 4);
    bulk_g2s(sB + (size_t)s * tile_bytes, imgb + (size_t)t * tile_bytes, tile_bytes, &bars[s]);
    bulk_g2s(sqB + s * QT, sqb + (size_t)t * QT, QT * 4, &bars[s]);
  };
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    for (int s = 0; s < NST + 3; ++s) mbar_init(&bars[s], 1);
    mbar_fence_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 0) {
    mbar_expect_tx(&bars[NST], tile_bytes);
    bulk_g2s(sA, imgb + (size_t)blockIdx.x * tile_bytes, tile_bytes, &bars[NST]);
    for (int t = 0; t < NST && t < ntile; ++t) issue(t);
  }
  // instruction descriptor: D=f32, A=B=bf16, K-major both, M=128, N=128
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(QT >> 3) << 17) | ((uint32_t)(QT >> 4) << 24);

  // epilogue role: TMEM lane quadrant q (rows q*32+lane), column block cb
  const int q = warp & 3, cb = warp >> 2;
  const int erow = q * 32 + lane;
  const float sq_e = sqb[i0 + erow];
  // selection role: this warp owns rows warp*RW .. +RW-1
  float ld[RW], tau[RW], eps[RW];
  int li[RW];
  const float smax = __uint_as_float(smax_bits[b]);
#pragma unroll
  for (int r = 0; r < RW; ++r) {
    ld[r] = CUDART_INF_F;
    li[r] = INT_MAX;
    tau[r] = CUDART_INF_F;
    const float sqi = sqb[i0 + warp * RW + r];
    eps[r] = 3.4527e-4f * sqrtf(sqi * smax) + 9.5367e-7f * (sqi + smax);   // 2^-11.5, 2^-20
  }

  auto issue_mma = [&](int t) {   // thread 0: tile t -> TMEM accumulator t & 1 (stage t % NST must have landed)
    const int s = t % NST;
    mbar_wait(&bars[s], (t / NST) & 1);
    tc_fence_after();
    const uint32_t a_hi = smem_u32(sA), a_lo = a_hi + ngrp * GROUP_BYTES;
    const uint32_t b_hi = smem_u32(sB + (size_t)s * tile_bytes), b_lo = b_hi + ngrp * GROUP_BYTES;
    const uint32_t acc = tmem_base + (uint32_t)(t & 1) * QT;
    uint32_t accum = 0;
#pragma unroll 1
    for (int pass = 0; pass < 3; ++pass) {
      const uint32_t ab = (pass == 1) ? a_lo : a_hi;
      const uint32_t bb = (pass == 2) ? b_lo : b_hi;
      for (int kk = 0; kk < Dp / 16; ++kk) {
        tc_mma(acc, umma_desc(ab + (uint32_t)(2 * kk) * GROUP_BYTES, GROUP_BYTES, 128),
               umma_desc(bb + (uint32_t)(2 * kk) * GROUP_BYTES, GROUP_BYTES, 128), idesc, accum);
        accum = 1;
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[NST + 1 + (t & 1)]))
                 : "memory");
  };

  mbar_wait(&bars[NST], 0);
  if (tid == 0) issue_mma(0);
  for (int t = 0; t < ntile; ++t) {
    const int s = t % NST;
    mbar_wait(&bars[NST + 1 + (t & 1)], (t >> 1) & 1);   // MMAs of tile t complete: accumulator t&1 ready, stage s consumed
    tc_fence_after();
    if (tid == 0) {
      // accumulator (t+1)&1 was drained by the epilogue of tile t-1 (ordered by the barriers of that iteration)
      if (t + 1 < ntile) issue_mma(t + 1);
    }
    mbar_wait(&bars[s], (t / NST) & 1);       // candidate norms of stage s (landed together with the tile image)

    // ---- TMEM -> approximate distances -> shared tile (previous tile's scan finished at the last barrier below)
    {
      float v[ECOLS];
      if (ECOLS == 32) tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((t & 1) * QT + cb * ECOLS), *reinterpret_cast<float(*)[32]>(v));
      else             tc_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((t & 1) * QT + cb * ECOLS), *reinterpret_cast<float(*)[16]>(v));
      const float* sj = sqB + s * QT + cb * ECOLS;
      float* drow = Ds + erow * DLD2 + cb * ECOLS;
      const int col0 = t * QT + cb * ECOLS;
#pragma unroll
      for (int i = 0; i < ECOLS; i += 4) {
        const float4 s4 = *reinterpret_cast<const float4*>(sj + i);
        float4 o;
        o.x = (sq_e + s4.x) - 2.f * v[i];
        o.y = (sq_e + s4.y) - 2.f * v[i + 1];
        o.z = (sq_e + s4.z) - 2.f * v[i + 2];
        o.w = (sq_e + s4.w) - 2.f * v[i + 3];
        if (col0 + i + 3 >= N) {            // ragged last tile: padded columns never pass
          const float qnan = __int_as_float(0x7fc00000);
          if (col0 + i + 0 >= N) o.x = qnan;
          if (col0 + i + 1 >= N) o.y = qnan;
          if (col0 + i + 2 >= N) o.z = qnan;
          o.w = qnan;
        }
        *reinterpret_cast<float4*>(drow + i) = o;
      }
    }
    tc_fence_before();
    __syncthreads();          // distance tile complete; candidate stage s and the accumulator are free
    tc_fence_after();
    if (tid == 0 && t + NST < ntile) issue(t + NST);

    // ---- streaming selection of the 32 smallest approximate distances per row
    {
      float v[RW][4];
#pragma unroll
      for (int r = 0; r < RW; ++r) {
        const float* drow = Ds + (warp * RW + r) * DLD2 + lane;
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) v[r][qq] = drow[32 * qq];
      }
      const int col0 = t * QT;
#pragma unroll
      for (int qq = 0; qq < 4; ++qq) {
        unsigned m[RW];
#pragma unroll
        for (int r = 0; r < RW; ++r) m[r] = __ballot_sync(0xffffffffu, v[r][qq] <= tau[r]);
#pragma unroll
        for (int r = 0; r < RW; ++r) {
          if (m[r] == 0u) continue;
          unsigned mm = m[r];
          do {
            const int src = __ffs(mm) - 1;
            mm &= mm - 1u;
            const float cd = __shfl_sync(0xffffffffu, v[r][qq], src);
            list_insert32(ld[r], li[r], cd, col0 + 32 * qq + src, lane);
          } while (mm);
          // pass threshold: k-th smallest so far + margin, but never beyond what the 32-slot list can hold
          const float kth = __shfl_sync(0xffffffffu, ld[r], k - 1);
          const float last = __shfl_sync(0xffffffffu, ld[r], 31);
          tau[r] = fminf(kth + 2.f * eps[r], last);
          // if `last` is the binding term the margin may have been truncated: resolved at the end (overflow flag)
        }
      }
    }
    __syncthreads();          // every warp finished scanning Ds before the next tile overwrites it
  }

  // ---- exact re-scoring: lane e owns candidate li[r] of row r
  const float* xb = x + (size_t)b * N * ldx + coff;
  const bool vec_ok = ((ldx & 3) == 0) && ((coff & 3) == 0) && ((D & 3) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
#pragma unroll 1
  for (int r = 0; r < RW; ++r) {
    const int row = i0 + warp * RW + r;
    if (row >= N) continue;                                   // warp-uniform
    const float kth = __shfl_sync(0xffffffffu, ld[r], k - 1);
    const float last_d = __shfl_sync(0xffffffffu, ld[r], 31);
    const int last_j = __shfl_sync(0xffffffffu, li[r], 31);
    const float thr = kth + 2.f * eps[r];
    bool flag = (last_j != INT_MAX) && (last_d <= thr);        // more than 32 candidates inside the margin
    const int j = li[r];
    const bool valid = (j != INT_MAX) && (ld[r] <= thr);
    float d = CUDART_INF_F;
    int jj = INT_MAX;
    if (valid) {
      const float* xi = xb + (size_t)row * ldx;
      const float* xj = xb + (size_t)j * ldx;
      float dot = 0.f;
      if (vec_ok) {   // 16-byte aligned rows: issue all loads first, then the canonical chain (c ascending from +0)
#pragma unroll 1
        for (int c0 = 0; c0 < D; c0 += 16) {
          float4 a[4], bq[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (c0 + 4 * u < D) {
              a[u] = *reinterpret_cast<const float4*>(xi + c0 + 4 * u);
              bq[u] = *reinterpret_cast<const float4*>(xj + c0 + 4 * u);
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (c0 + 4 * u < D) {
              dot = __fmaf_rn(a[u].x, bq[u].x, dot);
              dot = __fmaf_rn(a[u].y, bq[u].y, dot);
              dot = __fmaf_rn(a[u].z, bq[u].z, dot);
              dot = __fmaf_rn(a[u].w, bq[u].w, dot);
            }
          }
        }
      } else {
        for (int c = 0; c < D; ++c) dot = __fmaf_rn(xi[c], xj[c], dot);
      }
      d = exact_dist(flavour, sqb[row], sqb[j], dot);
      jj = j;
      if (!(fabsf(d - ld[r]) <= eps[r]) && flavour == WSPC_DIST_TFUTIL) flag = true;   // error-bound self check
      if (flavour != WSPC_DIST_TFUTIL) {
        // the clamp only moves negative values to 0: compare against the clamped approximation
        const float da = ld[r] > 0.f ? ld[r] : 0.f;
        if (!(fabsf(d - da) <= eps[r])) flag = true;
      }
    }
    flag = __any_sync(0xffffffffu, flag);
    warp_sort_pairs(d, jj, lane);
    if (lane < k) {
      const size_t o = ((size_t)b * N + row) * k + lane;
      idx_out[o] = jj;
      if (dist_out) dist_out[o] = d;
    }
    if (flag && lane == 0) {
      const int slot = atomicAdd(flag_count, 1);
      flag_rows[slot] = b * N + row;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
}

// ---------------------------------------------------------------- fallback ---
// exact streaming kNN for the flagged rows, one warp per row (lane = candidate column within a 32-column group)
__global__ void __launch_bounds__(256)
knn_exact_rows_kernel(const float* __restrict__ x, const float* __restrict__ sq, int N, int Npad, int ldx, int coff, int D,
                      int k, int flavour, const int* __restrict__ flag_count, const int* __restrict__ flag_rows,
                      int32_t* __restrict__ idx_out, float* __restrict__ dist_out) {
  const int lane = threadIdx.x & 31;
  const int nflag = *flag_count;
  for (int f = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); f < nflag; f += gridDim.x * (blockDim.x >> 5)) {
    const int grow = flag_rows[f];
    const int b = grow / N, row = grow - b * N;
    const float* xb = x + (size_t)b * N * ldx + coff;
    const float* xi = xb + (size_t)row * ldx;
    const float sqi = sq[(size_t)b * Npad + row];
    float ld = CUDART_INF_F;
    int li = INT_MAX;
    float tau = CUDART_INF_F;
    for (int base = 0; base < N; base += 32) {
      const int j = base + lane;
      float d = __int_as_float(0x7fc00000);
      if (j < N) {
        const float* xj = xb + (size_t)j * ldx;
        float dot = 0.f;
        for (int c = 0; c < D; ++c) dot = __fmaf_rn(xi[c], xj[c], dot);
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
    if (lane < k) {
      const size_t o = (size_t)grow * k + lane;
      idx_out[o] = li;
      if (dist_out) dist_out[o] = ld;
    }
  }
}

}  // namespace

// ---------------------------------------------------------------- host ------
bool knn_tc_eligible(int D, int k) { return D >= 16 && D <= 64 && k <= 24; }

size_t knn_tc_workspace_bytes(int B, int N, int D) {
  const int Dp = (D + 15) / 16 * 16, Npad = (N + QT - 1) / QT * QT;
  const size_t img = (size_t)B * (Npad / QT) * 2 * (Dp / 8) * GROUP_BYTES;
  return align_up(img, 256) + align_up((size_t)B * Npad * 4, 256) + align_up((size_t)B * 4, 256) + 256 +
         align_up((size_t)B * N * 4, 256);
}

// telemetry: number of rows the last knn_tc_run on this workspace sent to the exact fallback (synchronises)
int knn_tc_fallback_rows(const void* ws, int B, int N, int D, int* out) {
  const int Dp = (D + 15) / 16 * 16, Npad = (N + QT - 1) / QT * QT;
  const char* w = static_cast<const char*>(ws);
  w += align_up((size_t)B * (Npad / QT) * 2 * (Dp / 8) * GROUP_BYTES, 256) + align_up((size_t)B * Npad * 4, 256) +
       align_up((size_t)B * 4, 256);
  WSPC_CUDA(cudaMemcpy(out, w, sizeof(int), cudaMemcpyDeviceToHost));
  return WSPC_OK;
}

int knn_tc_run(const float* x, int B, int N, int ldx, int coff, int D, int k, int flavour, int32_t* idx, float* dist,
               void* ws, size_t ws_bytes, cudaStream_t st) {
  const int Dp = (D + 15) / 16 * 16, Npad = (N + QT - 1) / QT * QT;
  if (ws_bytes < knn_tc_workspace_bytes(B, N, D)) {
    set_error("knn_fused: workspace %zu < required %zu", ws_bytes, knn_tc_workspace_bytes(B, N, D));
    return WSPC_ERR_WORKSPACE;
  }
  char* w = static_cast<char*>(ws);
  const size_t img_bytes = align_up((size_t)B * (Npad / QT) * 2 * (Dp / 8) * GROUP_BYTES, 256);
  unsigned char* img = reinterpret_cast<unsigned char*>(w); w += img_bytes;
  float* sq = reinterpret_cast<float*>(w); w += align_up((size_t)B * Npad * 4, 256);
  unsigned* smax = reinterpret_cast<unsigned*>(w); w += align_up((size_t)B * 4, 256);
  int* flag_count = reinterpret_cast<int*>(w); w += 256;
  int* flag_rows = reinterpret_cast<int*>(w);
  WSPC_CUDA(cudaMemsetAsync(smax, 0, align_up((size_t)B * 4, 256) + 256, st));
  knn_tc_prep_kernel<<<dim3(Npad / QT, B), 128, 0, st>>>(x, N, ldx, coff, D, Dp, Npad, img, sq, smax);
  const size_t tile_bytes = (size_t)2 * (Dp / 8) * GROUP_BYTES;
  const size_t smem = (1 + NST) * tile_bytes + (size_t)QT * DLD2 * 4 + NST * QT * 4 + (NST + 3) * 8 + 16;
  static thread_local size_t configured = 0;
  if (smem > configured) {
    WSPC_CUDA(cudaFuncSetAttribute(knn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  knn_tc_kernel<<<dim3(Npad / QT, B), KTC_THREADS, smem, st>>>(img, sq, smax, x, N, Npad, ldx, coff, D, Dp, k, flavour, idx,
                                                              dist, flag_count, flag_rows);
  knn_exact_rows_kernel<<<2 * kNumSM, 256, 0, st>>>(x, sq, N, Npad, ldx, coff, D, k, flavour, flag_count, flag_rows, idx,
                                                    dist);
  count_launch(3);
  WSPC_LAUNCH_CHECK("knn_tc kernels");
  return WSPC_OK;
}

}  // namespace wspc
