// tcgen05 (5th-gen tensor core) path for the wide shared-MLP GEMMs:  out(M,N) = A(M,K) * Bm(K,N)
//
// Used for every 1x1-conv forward / data-gradient GEMM whose K is a multiple of 8 with 16-byte aligned rows
// (all EdgeConv layers except the 18-wide first one, and the per-point layers).  One CTA owns 128-row tiles
// (persistent loop).  Per tile:
//   1. all 8 warps synthesise the A operand on load (BN+ReLU of the previous layer, gathered edge feature,
//      or the BN-backward affine of the upstream gradient), split every fp32 value into bf16 hi + bf16 lo
//      and store both into shared memory in the UMMA no-swizzle K-major core-matrix layout
//      ([K/8][rows][8 elems], 16 B per row and 8-column group, +16 B pad per group against bank conflicts);
//   2. one thread issues tcgen05.mma (kind::f16, bf16 x bf16 -> fp32 in TMEM) three times over K:
//      A_hi*B_hi + A_lo*B_hi + A_hi*B_lo  (error ~2^-16 relative, i.e. fp32-class accuracy for the 1e-3
//      parity bar; a single bf16 pass would not hold it), then tcgen05.commit -> mbarrier;
//   3. all warps read the accumulator back with tcgen05.ld (one thread = one output row), stage it through
//      shared memory and write it out coalesced while folding the epilogue (bias / BN batch statistics /
//      ReLU-mask + BN-backward sums / edge-gradient scatter).
// The weight matrix is split and laid out once per CTA.  Two CTAs per SM overlap load, MMA and epilogue
// phases of different tiles when shared memory allows (K<=64).
#include "operand.cuh"
#include <cuda_bf16.h>
#include <type_traits>
#include <cstdlib>
#include <cstring>

namespace wspc {
void count_launch(int n = 1);
namespace {

constexpr int TC_THREADS = 256;
constexpr int TILE_M = 128;
constexpr int A_GROUP_BYTES = TILE_M * 16 + 16;   // one 8-column group of the A tile (LBO), padded
constexpr int STAGE_LD = 64 + 4;                  // staging row pitch (floats) for a 64-column pass

// ---------------------------------------------------------------- tcgen05 PTX ---
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_alloc(uint32_t* slot_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
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

// UMMA shared-memory descriptor, no swizzle, K-major: core matrix = 8 rows x 16 B stored contiguously (128 B);
// SBO = distance between 8-row groups, LBO = distance between the two 8-column halves of a K=16 slice.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// instruction descriptor: D=f32, A=B=bf16, both K-major, M=128
__device__ __forceinline__ uint32_t umma_idesc(int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
}

// x = hi + lo with hi, lo in bf16: packed conversions (F2FP.BF16.PACK_AB converts two floats per instruction);
// a bf16 is the upper half of the fp32 pattern, so the round trip back to fp32 is a shift / mask.
__device__ __forceinline__ void split8(const float (&v)[8], uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 hp = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    const uint32_t hb = *reinterpret_cast<const uint32_t*>(&hp);
    const float h0 = __uint_as_float(hb << 16), h1 = __uint_as_float(hb & 0xffff0000u);
    const __nv_bfloat162 lp = __floats2bfloat162_rn(v[2 * i] - h0, v[2 * i + 1] - h1);
    h[i] = hb;
    l[i] = *reinterpret_cast<const uint32_t*>(&lp);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// the same split for 4 values (same arithmetic per element as split8)
__device__ __forceinline__ void split4(const float (&v)[4], uint2& hi, uint2& lo) {
  uint32_t h[2], l[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const __nv_bfloat162 hp = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    const uint32_t hb = *reinterpret_cast<const uint32_t*>(&hp);
    const float h0 = __uint_as_float(hb << 16), h1 = __uint_as_float(hb & 0xffff0000u);
    const __nv_bfloat162 lp = __floats2bfloat162_rn(v[2 * i] - h0, v[2 * i + 1] - h1);
    h[i] = hb;
    l[i] = *reinterpret_cast<const uint32_t*>(&lp);
  }
  hi = make_uint2(h[0], h[1]);
  lo = make_uint2(l[0], l[1]);
}

__device__ __forceinline__ void ld8(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Rows of an edge tensor are (point, neighbour) pairs, row = pt*k + j.  The 64-bit divisions are done once per
// 128-row tile; the per-row part is one 32-bit multiply-high (exact while (k+128)*k < 2^32: k <= 32768).
struct RowMap {
  long long pt0, cb0;    // point of the tile's first row; first point of that point's cloud
  uint32_t rem0, npts;   // row0 - pt0*k
  uint64_t inv;          // floor(2^32/k)+1
};
__device__ __forceinline__ uint64_t rowmap_inv(int k) { return (1ull << 32) / (uint32_t)(k > 0 ? k : 1) + 1; }
__device__ __forceinline__ long long div_pos(long long a, int b) {
  return (a < (1ll << 31)) ? (long long)((uint32_t)a / (uint32_t)b) : a / b;
}
__device__ __forceinline__ RowMap rowmap_tile(long long row0, int k, int npts, uint64_t inv) {
  RowMap m;
  m.pt0 = div_pos(row0, k);
  m.rem0 = (uint32_t)(row0 - m.pt0 * k);
  m.cb0 = div_pos(m.pt0, npts) * npts;
  m.npts = (uint32_t)npts;
  m.inv = inv;
  return m;
}
__device__ __forceinline__ void rowmap_point(const RowMap& m, int r, long long& pt, long long& cb) {
  const uint32_t x = m.rem0 + (uint32_t)r;
  pt = m.pt0 + (uint32_t)(((uint64_t)x * m.inv) >> 32);
  cb = m.cb0;
  while (pt - cb >= m.npts) cb += m.npts;
}

// Loads the 8 channels [kg*8, kg*8+8) of logical row `row` of an operand (zeros if !valid).
// EDGE: `pt`/`cb` are the row's point and the first point of its cloud (rowmap_point).
// pc0..2 hold the thread's per-channel constants (BNRELU: sc, sh; DY: c1, c2, c3).
// `generic` selects the element-wise loader of operand.cuh (any alignment, any channel count).
template <int AMODE>
__device__ __forceinline__ void load_chunk(const Operand& A, long long row, int kg, bool valid, const float (&pc0)[8],
                                           const float (&pc1)[8], const float (&pc2)[8], float (&v)[8],
                                           bool generic = false, long long pt = 0, long long cb = 0) {
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = 0.f;
  if (!valid) return;
  if (generic) {
    load8<AMODE>(A, row, kg * 8, v);
    return;
  }
  if (AMODE == OP_PLAIN) {
    ld8(A.p + row * A.ld + kg * 8, v);
  } else if (AMODE == OP_BNRELU) {
    float y[8];
    ld8(A.p + row * A.ld + kg * 8, y);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = fmaxf(fmaf(y[i], pc0[i], pc1[i]), 0.f);
    if (A.dmask) {
      float m[8];
      ld8(A.dmask + row * A.C + kg * 8, m);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] *= m[i] * A.dscale;
    }
  } else if (AMODE == OP_EDGE) {
    const int Cx = A.C >> 1;
    if (kg * 8 < Cx) {
      ld8(A.p + pt * A.ld + kg * 8, v);
    } else {
      const long long nb = cb + A.idx[row];
      float xi[8], xj[8];
      ld8(A.p + pt * A.ld + kg * 8 - Cx, xi);
      ld8(A.p + nb * A.ld + kg * 8 - Cx, xj);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = xj[i] - xi[i];
    }
  } else if (AMODE == OP_DY) {
    float g[8];
    ld8(A.p + row * A.ld + kg * 8, g);
    if (A.c1) {
      float y[8];
      ld8(A.y + row * A.ldy + kg * 8, y);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = fmaf(pc0[i], g[i], fmaf(pc2[i], y[i], pc1[i]));
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = g[i];
    }
  } else if (AMODE == OP_DY_MAXK) {   // pt = point of the row
    float y[8], m[8], sh8[8], sc8[8], hh[8];
    ld8(A.y + row * A.ldy + kg * 8, y);
    ld8(A.p + pt * A.ld + kg * 8, m);
    ld8(A.p + pt * A.ld + A.C + kg * 8, sh8);
    ld8(A.sc + kg * 8, sc8);
    ld8(A.sh + kg * 8, hh);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = fmaf(pc0[i], maxk_grad(y[i], sc8[i], hh[i], m[i], sh8[i]), fmaf(pc2[i], y[i], pc1[i]));
  } else {  // OP_DY_SPARSE: pt = cloud of the row, cb = point index inside the cloud
    float y[8], dg[8];
    ld8(A.y + row * A.ldy + kg * 8, y);
    ld8(A.dg + pt * A.C + kg * 8, dg);
    const int4 m0 = *reinterpret_cast<const int4*>(A.amax + pt * A.C + kg * 8);
    const int4 m1 = *reinterpret_cast<const int4*>(A.amax + pt * A.C + kg * 8 + 4);
    const int am[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
    const int n = (int)cb;
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = fmaf(pc0[i], am[i] == n ? dg[i] : 0.f, fmaf(pc2[i], y[i], pc1[i]));
  }
}

// DY_SPARSE rows are points: (cloud, n) of tile-local row r from the tile's first row (one division per tile)
struct CloudMap { long long cloud0; uint32_t rem0, npts; };
__device__ __forceinline__ CloudMap cloudmap_tile(long long row0, int npts) {
  CloudMap m;
  m.cloud0 = div_pos(row0, npts);
  m.rem0 = (uint32_t)(row0 - m.cloud0 * npts);
  m.npts = (uint32_t)npts;
  return m;
}
__device__ __forceinline__ void cloudmap_point(const CloudMap& m, int r, long long& cloud, long long& n) {
  uint32_t x = m.rem0 + (uint32_t)r;
  cloud = m.cloud0;
  while (x >= m.npts) { x -= m.npts; ++cloud; }
  n = x;
}

template <int AMODE>
__device__ __forceinline__ void load_consts(const Operand& A, int kg, int K, float (&pc0)[8], float (&pc1)[8],
                                            float (&pc2)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = kg * 8 + i;
    pc0[i] = 0.f; pc1[i] = 0.f; pc2[i] = 0.f;
    if (AMODE == OP_BNRELU && c < K) { pc0[i] = A.sc[c]; pc1[i] = A.sh[c]; }
    if ((AMODE == OP_DY || AMODE == OP_DY_SPARSE || AMODE == OP_DY_MAXK) && A.c1 && c < K) {
      pc0[i] = A.c1[c]; pc1[i] = A.c2[c]; pc2[i] = A.c3[c];
    }
  }
}

// Split form of load_chunk for software-batched loops: fetch_chunk only issues the global loads of a chunk
// (so a batch of chunks has all its loads in flight before the first use), finish_chunk does the arithmetic.
struct RawChunk { float a[8], b[8]; };
template <int AMODE>
__device__ __forceinline__ void fetch_chunk(const Operand& A, long long row, int kg, bool valid, long long pt,
                                            long long nb, RawChunk& w) {
#pragma unroll
  for (int i = 0; i < 8; ++i) { w.a[i] = 0.f; w.b[i] = 0.f; }
  if (!valid) return;
  if (AMODE == OP_PLAIN || AMODE == OP_BNRELU) {
    ld8(A.p + row * A.ld + kg * 8, w.a);
  } else if (AMODE == OP_EDGE) {
    const int Cx = A.C >> 1;
    if (kg * 8 < Cx) {
      ld8(A.p + pt * A.ld + kg * 8, w.a);
    } else {
      ld8(A.p + nb * A.ld + kg * 8 - Cx, w.a);
      ld8(A.p + pt * A.ld + kg * 8 - Cx, w.b);
    }
  } else if (AMODE == OP_DY) {
    ld8(A.p + row * A.ld + kg * 8, w.a);
    if (A.c1) ld8(A.y + row * A.ldy + kg * 8, w.b);
  } else if (AMODE == OP_DY_MAXK) {
    ld8(A.p + pt * A.ld + kg * 8, w.a);            // pooled maximum of the row's point
    ld8(A.y + row * A.ldy + kg * 8, w.b);
  } else {  // OP_DY_SPARSE
    ld8(A.y + row * A.ldy + kg * 8, w.b);
  }
}
template <int AMODE>
__device__ __forceinline__ void finish_chunk(const Operand& A, long long row, int kg, bool valid, long long pt,
                                             long long cb, const float (&pc0)[8], const float (&pc1)[8],
                                             const float (&pc2)[8], const RawChunk& w, float (&v)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = 0.f;
  if (!valid) return;
  if (AMODE == OP_PLAIN) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = w.a[i];
  } else if (AMODE == OP_BNRELU) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = fmaxf(fmaf(w.a[i], pc0[i], pc1[i]), 0.f);
    if (A.dmask) {
      float m[8];
      ld8(A.dmask + row * A.C + kg * 8, m);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] *= m[i] * A.dscale;
    }
  } else if (AMODE == OP_EDGE) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = w.a[i] - w.b[i];     // centre groups: b = 0
  } else if (AMODE == OP_DY) {
    if (A.c1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = fmaf(pc0[i], w.a[i], fmaf(pc2[i], w.b[i], pc1[i]));
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = w.a[i];
    }
  } else if (AMODE == OP_DY_MAXK) {
    float sh8[8], sc8[8], hh[8];
    ld8(A.p + pt * A.ld + A.C + kg * 8, sh8);
    ld8(A.sc + kg * 8, sc8);
    ld8(A.sh + kg * 8, hh);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      v[i] = fmaf(pc0[i], maxk_grad(w.b[i], sc8[i], hh[i], w.a[i], sh8[i]), fmaf(pc2[i], w.b[i], pc1[i]));
  } else {  // OP_DY_SPARSE: pt = cloud, cb = point inside the cloud
    float dg[8];
    ld8(A.dg + pt * A.C + kg * 8, dg);
    const int4 m0 = *reinterpret_cast<const int4*>(A.amax + pt * A.C + kg * 8);
    const int4 m1 = *reinterpret_cast<const int4*>(A.amax + pt * A.C + kg * 8 + 4);
    const int am[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
    const int n = (int)cb;
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = fmaf(pc0[i], am[i] == n ? dg[i] : 0.f, fmaf(pc2[i], w.b[i], pc1[i]));
  }
}

struct TcSmem {
  int Kp, Npad, b_group_bytes;
  size_t off_bhi, off_blo, off_ahi, off_alo, off_stage, off_misc, total;
};
__host__ __device__ inline TcSmem tc_smem_plan(int K, int N, bool need_stage) {
  TcSmem s;
  s.Kp = (K + 15) / 16 * 16;
  s.Npad = (N + 15) / 16 * 16;
  s.b_group_bytes = s.Npad * 16 + 16;
  const size_t bbytes = (size_t)(s.Kp / 8) * s.b_group_bytes;
  const size_t abytes = (size_t)(s.Kp / 8) * A_GROUP_BYTES;
  size_t o = 0;
  s.off_bhi = o; o += (bbytes + 127) / 128 * 128;
  s.off_blo = o; o += (bbytes + 127) / 128 * 128;
  s.off_ahi = o; o += (abytes + 127) / 128 * 128;
  s.off_alo = o; o += (abytes + 127) / 128 * 128;
  // the epilogue staging tile aliases the A buffers: A is dead once the tile's MMAs have completed
  s.off_stage = s.off_ahi;
  const size_t stage_end = s.off_stage + (need_stage ? (size_t)TILE_M * STAGE_LD * 4 : 0);
  if (o < stage_end) o = (stage_end + 127) / 128 * 128;
  s.off_misc = o; o += 64;
  s.total = o;
  return s;
}

// ------------------------------------------------------------------ kernel ---
// K is processed in chunks of KC channels (KC = Kp when K <= 128, else 64) accumulated in TMEM; N is tiled
// by blockIdx.y (whole N when it fits one tile, else 128 columns per tile).  When there is a single chunk
// and a single N tile (all EdgeConv layers) the split weights stay resident in shared memory for the whole
// persistent loop; otherwise the CTA re-splits the (KC x Nt) weight block it needs per chunk (L2-resident).
template <int AMODE, int EMODE, int MINB, int MAXPASS>
__global__ void __launch_bounds__(TC_THREADS, MINB)
rowgemm_tc_kernel(const Operand A, const float* __restrict__ Bm, long long ldb, int bT, long long M, int N, int K,
                  const Epilogue E, int num_tiles, int tmem_cols, int KC, int NtMax, int generic,
                  const unsigned char* __restrict__ wimg) {
  extern __shared__ __align__(128) unsigned char smem[];
  const TcSmem sp = tc_smem_plan(KC, NtMax, true);
  unsigned char* sBhi = smem + sp.off_bhi;
  unsigned char* sBlo = smem + sp.off_blo;
  unsigned char* sAhi = smem + sp.off_ahi;
  unsigned char* sAlo = smem + sp.off_alo;
  float* stage = reinterpret_cast<float*>(smem + sp.off_stage);
  uint64_t* mma_bar = reinterpret_cast<uint64_t*>(smem + sp.off_misc);
  uint64_t* w_bar = reinterpret_cast<uint64_t*>(smem + sp.off_misc + 8);      // pre-split weight block landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + sp.off_misc + 16);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int k8c = sp.Kp / 8;                      // channel groups per chunk
  const int nkc = (K + KC - 1) / KC;
  const int n0 = (gridDim.y > 1) ? blockIdx.y * NtMax : 0;
  const int Nt = (N - n0 < NtMax) ? (N - n0) : NtMax;
  const int Ntp = (Nt + 15) / 16 * 16;
  const bool resident_w = (nkc == 1);

  if (warp == 0) tc_alloc(tmem_slot, (uint32_t)tmem_cols);
  if (tid == 32) {
    mbar_init(mma_bar, 1);
    mbar_init(w_bar, 1);
    mbar_fence_init();
  }
  // weights of chunk kc: element (n, k) of Bm^T -> group (k - kc*KC)/8, row n, slot k%8 ; zero padded
  auto load_w = [&](int kc) {
    for (int e = tid; e < sp.Npad * k8c; e += TC_THREADS) {
      const int n = e % sp.Npad, g = e / sp.Npad;
      float w[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = kc * KC + g * 8 + i;
        w[i] = (n < Nt && k < K) ? (bT ? Bm[(long long)(n0 + n) * ldb + k] : Bm[(long long)k * ldb + n0 + n]) : 0.f;
      }
      uint4 hi, lo;
      split8(w, hi, lo);
      *reinterpret_cast<uint4*>(sBhi + (size_t)g * sp.b_group_bytes + n * 16) = hi;
      *reinterpret_cast<uint4*>(sBlo + (size_t)g * sp.b_group_bytes + n * 16) = lo;
    }
  };
  if (resident_w) load_w(0);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t idesc = umma_idesc(Ntp);

  // this thread always handles channel group kg (within the chunk) of rows r0, r0+rstep, ...
  const int kg = tid % k8c;
  const int rstep = TC_THREADS / k8c;
  const int r0 = tid / k8c;
  float pc0[8], pc1[8], pc2[8];
  if (resident_w) load_consts<AMODE>(A, kg, K, pc0, pc1, pc2);

  const int e_c4 = tid & 15, e_r0 = tid >> 4;     // epilogue: 4 fixed columns, 16 rows per sweep
  constexpr bool kStats = (EMODE == EPI_STORE_STATS || EMODE == EPI_RELUMASK_STATS);
  double st0[kStats ? MAXPASS : 1][4], st1[kStats ? MAXPASS : 1][4];
  if (kStats) {
#pragma unroll
    for (int p = 0; p < MAXPASS; ++p)
#pragma unroll
      for (int j = 0; j < 4; ++j) { st0[p][j] = 0.0; st1[p][j] = 0.0; }
  }

  uint32_t phase = 0, wphase = 0;
  // chunked K with a pre-split weight image (wprep_kernel): one bulk copy per chunk stages [hi | lo] (contiguous in
  // shared memory for 8 channel groups) while the warps synthesise the A operand
  const uint32_t wblock_bytes = 2u * (uint32_t)k8c * (uint32_t)sp.b_group_bytes;
  const bool w_async = !resident_w && wimg != nullptr;
  const uint64_t a_inv = (AMODE == OP_EDGE || AMODE == OP_DY_MAXK) ? rowmap_inv(A.k) : 0;
  const uint64_t e_inv = (EMODE == EPI_EDGE_SCATTER) ? rowmap_inv(E.k) : 0;
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const long long row0 = (long long)tile * TILE_M;
    uint32_t accum = 0;
    RowMap arm, erm;
    CloudMap acm;
    if (AMODE == OP_EDGE || AMODE == OP_DY_MAXK) arm = rowmap_tile(row0, A.k, A.npts, a_inv);
    if (AMODE == OP_DY_SPARSE) acm = cloudmap_tile(row0, A.npts);
    if (EMODE == EPI_EDGE_SCATTER) erm = rowmap_tile(row0, E.k, E.npts, e_inv);

    for (int kc = 0; kc < nkc; ++kc) {

      // ------------------------------------------------------------ 1. operands -> smem ------
      const int cg = kc * (KC / 8) + kg;            // absolute channel group
      const bool kvalid = cg * 8 < K;
      if (!resident_w) {
        if (w_async) {
          if (tid == 0) {   // the previous chunk's MMAs (the last readers of sB) retired at the mma_bar wait below
            mbar_expect_tx(w_bar, wblock_bytes);
            bulk_g2s(sBhi, wimg + ((size_t)blockIdx.y * nkc + kc) * wblock_bytes, wblock_bytes, w_bar);
          }
        } else {
          load_w(kc);
        }
        load_consts<AMODE>(A, cg, K, pc0, pc1, pc2);
      }
      if (AMODE != OP_EDGE && AMODE != OP_DY_MAXK && !generic) {
        // batches of UB rows: all global loads of a batch are in flight before the first use (the kernel is
        // latency-bound on these loads: 16 resident warps per SM).  Measured: forward layers 18.8 -> 17.9 ms per
        // step; the gradient operands (two loads per row already) lose 1 ms with the same batching, so they keep
        // the plain loop below.
        constexpr int UB = (AMODE == OP_PLAIN || AMODE == OP_BNRELU) ? 4 : 2;   // gradient operands: two loads per row
#pragma unroll 1
        for (int rbase = r0; rbase < TILE_M; rbase += rstep * UB) {
          long long pt[UB], cb[UB];
          bool ok[UB];
          RawChunk w[UB];
#pragma unroll
          for (int u = 0; u < UB; ++u) {
            const int r = rbase + rstep * u;
            ok[u] = r < TILE_M && row0 + r < M && kvalid;
            pt[u] = 0; cb[u] = 0;
            if (AMODE == OP_DY_MAXK && r < TILE_M) rowmap_point(arm, r, pt[u], cb[u]);
            if (AMODE == OP_DY_SPARSE && r < TILE_M) cloudmap_point(acm, r, pt[u], cb[u]);
          }
#pragma unroll
          for (int u = 0; u < UB; ++u) fetch_chunk<AMODE>(A, row0 + rbase + rstep * u, cg, ok[u], pt[u], 0, w[u]);
#pragma unroll
          for (int u = 0; u < UB; ++u) {
            const int r = rbase + rstep * u;
            if (r < TILE_M) {
              float v[8];
              finish_chunk<AMODE>(A, row0 + r, cg, ok[u], pt[u], cb[u], pc0, pc1, pc2, w[u], v);
              uint4 hi, lo;
              split8(v, hi, lo);
              *reinterpret_cast<uint4*>(sAhi + (size_t)kg * A_GROUP_BYTES + r * 16) = hi;
              *reinterpret_cast<uint4*>(sAlo + (size_t)kg * A_GROUP_BYTES + r * 16) = lo;
            }
          }
        }
      } else {
#pragma unroll 4
        for (int r = r0; r < TILE_M; r += rstep) {
          const long long row = row0 + r;
          float v[8];
          long long pt = 0, cb = 0;
          if (AMODE == OP_EDGE || AMODE == OP_DY_MAXK) rowmap_point(arm, r, pt, cb);
          if (AMODE == OP_DY_SPARSE) cloudmap_point(acm, r, pt, cb);
          load_chunk<AMODE>(A, row, cg, row < M && kvalid, pc0, pc1, pc2, v, generic != 0, pt, cb);
          uint4 hi, lo;
          split8(v, hi, lo);
          *reinterpret_cast<uint4*>(sAhi + (size_t)kg * A_GROUP_BYTES + r * 16) = hi;
          *reinterpret_cast<uint4*>(sAlo + (size_t)kg * A_GROUP_BYTES + r * 16) = lo;
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncthreads();
      // ------------------------------------------------------------ 2. MMA issue --------------
      if (tid == 0) {
        if (w_async) mbar_wait(w_bar, wphase);
        tc_fence_after();
        const uint32_t a_hi = smem_u32(sAhi), a_lo = smem_u32(sAlo), b_hi = smem_u32(sBhi), b_lo = smem_u32(sBlo);
#pragma unroll 1
        for (int pass = 0; pass < 3; ++pass) {
          const uint32_t ab = (pass == 1) ? a_lo : a_hi;
          const uint32_t bb = (pass == 2) ? b_lo : b_hi;
          for (int kk = 0; kk < sp.Kp / 16; ++kk) {
            const uint64_t ad = umma_desc(ab + (uint32_t)(2 * kk) * A_GROUP_BYTES, A_GROUP_BYTES, 128);
            const uint64_t bd = umma_desc(bb + (uint32_t)(2 * kk) * sp.b_group_bytes, sp.b_group_bytes, 128);
            tc_mma_bf16(tmem_base, ad, bd, idesc, accum);
            accum = 1;
          }
        }
        tc_commit(mma_bar);
      }
      mbar_wait(mma_bar, phase);                    // MMAs of this chunk finished (smem reusable)
      phase ^= 1;
      wphase ^= 1;
    }
    tc_fence_after();

    // ---------------------------------------------------------------- 3. epilogue ---------------
    const int lq = warp & 3, ch = warp >> 2;        // TMEM lane quadrant / 32-column half of a 64-col pass
    const int trow = lq * 32 + lane;                // tile-local row owned in TMEM
    const int npass = (Nt + 63) / 64;
#pragma unroll
    for (int p = 0; p < MAXPASS; ++p) {
      if (p >= npass) break;
      float v[32];
      tc_ld32(tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(p * 64 + ch * 32), v);
      float* srow = stage + trow * STAGE_LD + ch * 32;
      if (EMODE == EPI_EDGE_SCATTER && p * 64 >= (N >> 1)) {
        // neighbour half [dE_d]: scatter to dx[idx], subtract from the staged centre sum
        const int Cx = N >> 1;
        const long long row = row0 + trow;
        if (row < M) {
          long long pt, cb;
          rowmap_point(erm, trow, pt, cb);
          const long long nb = cb + E.idx[row];
          float* dst = E.dx + nb * E.lddx + (p * 64 - Cx) + ch * 32;
#pragma unroll
          for (int i = 0; i < 32; i += 4) red_add_v4(dst + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          float4 c = *reinterpret_cast<float4*>(srow + i);
          c.x -= v[i]; c.y -= v[i + 1]; c.z -= v[i + 2]; c.w -= v[i + 3];
          *reinterpret_cast<float4*>(srow + i) = c;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(srow + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      }
      if (EMODE == EPI_EDGE_SCATTER && p + 1 < npass) { __syncwarp(); continue; }   // pass 0 = centre, pass 1 = neighbour
      __syncthreads();
      if (EMODE == EPI_EDGE_SCATTER) {
        // centre reduction: rows of one point are consecutive; thread = (column, quarter of the rows)
        const int col = tid & 63, q = tid >> 6;
        long long cur, cb;
        rowmap_point(erm, q * 32, cur, cb);
        int left = E.k - (int)(erm.rem0 + (uint32_t)(q * 32) - (uint32_t)(cur - erm.pt0) * (uint32_t)E.k);   // rows left of point `cur`
        float acc = 0.f;
        bool any = false;
        for (int r = q * 32; r < q * 32 + 32; ++r) {
          if (row0 + r >= M) break;
          if (left == 0) {
            atomicAdd(E.dx + cur * E.lddx + col, acc);
            ++cur;
            left = E.k;
            acc = 0.f;
          }
          acc += stage[r * STAGE_LD + col];
          --left;
          any = true;
        }
        if (any) atomicAdd(E.dx + cur * E.lddx + col, acc);
      } else {
        const int cl = p * 64 + e_c4 * 4;           // column within this N tile
        const int cbase = n0 + cl;                  // global column
        if (cl < Nt) {
          float bias[4] = {0.f, 0.f, 0.f, 0.f}, scp[4] = {0.f, 0.f, 0.f, 0.f}, shp[4] = {-1.f, -1.f, -1.f, -1.f};
          if ((EMODE == EPI_STORE || EMODE == EPI_STORE_STATS) && E.bias) {
#pragma unroll
            for (int j = 0; j < 4; ++j) bias[j] = E.bias[cbase + j];
          }
          if (EMODE == EPI_RELUMASK_STATS) {
#pragma unroll
            for (int j = 0; j < 4; ++j) { scp[j] = E.scp[cbase + j]; shp[j] = E.shp[cbase + j]; }
          }
          float f0[4] = {0.f, 0.f, 0.f, 0.f}, f1[4] = {0.f, 0.f, 0.f, 0.f};   // fp32 partials of this tile (8 rows)
          long long rb_cloud0 = 0;
          uint32_t rb_rem0 = 0;
          if ((EMODE == EPI_STORE || EMODE == EPI_STORE_STATS) && E.rowbias) {
            rb_cloud0 = div_pos(row0, E.rb_rows);
            rb_rem0 = (uint32_t)(row0 - rb_cloud0 * E.rb_rows);
          }
          // global operands of the epilogue (previous pre-activations / the accumulation target) for the thread's 8 rows:
          // all loads in flight before the first use instead of one dependent round trip per row
          float4 pre[8];
          if (EMODE == EPI_RELUMASK_STATS || EMODE == EPI_ACCUM) {
            const float* src = (EMODE == EPI_RELUMASK_STATS) ? E.yprev : E.out;
            const long long lds = (EMODE == EPI_RELUMASK_STATS) ? E.ldyp : E.ldo;
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              const long long row = row0 + e_r0 + 16 * it;
              pre[it] = (row < M) ? *reinterpret_cast<const float4*>(src + row * lds + cbase) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int r = e_r0 + 16 * it;
            const long long row = row0 + r;
            if (row >= M) break;
            const float4 s4 = *reinterpret_cast<const float4*>(stage + r * STAGE_LD + e_c4 * 4);
            float o[4] = {s4.x, s4.y, s4.z, s4.w};
            if (EMODE == EPI_STORE || EMODE == EPI_STORE_STATS) {
              const float* rb = nullptr;
              if (E.rowbias) {
                long long cl0 = rb_cloud0;
                uint32_t x = rb_rem0 + (uint32_t)r;
                if (x >= (uint32_t)E.rb_rows) { cl0 += x / (uint32_t)E.rb_rows; }
                rb = E.rowbias + cl0 * E.ldrb + cbase;
              }
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                o[j] += bias[j];
                if (rb) o[j] += rb[j];
              }
              if (EMODE == EPI_STORE_STATS) {
#pragma unroll
                for (int j = 0; j < 4; ++j) { f0[j] += o[j]; f1[j] = fmaf(o[j], o[j], f1[j]); }
              }
            } else if (EMODE == EPI_RELUMASK_STATS) {
              const float4 y4 = pre[it];
              const float yv[4] = {y4.x, y4.y, y4.z, y4.w};
              float dm[4] = {1.f, 1.f, 1.f, 1.f};
              if (E.dmask) {
                const float4 m4 = *reinterpret_cast<const float4*>(E.dmask + row * N + cbase);
                dm[0] = m4.x * E.dscale; dm[1] = m4.y * E.dscale; dm[2] = m4.z * E.dscale; dm[3] = m4.w * E.dscale;
              }
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const bool on = fmaf(yv[j], scp[j], shp[j]) > 0.f;
                o[j] = on ? o[j] * dm[j] : 0.f;
                f0[j] += o[j];
                f1[j] = fmaf(o[j], yv[j], f1[j]);
              }
            } else if (EMODE == EPI_ACCUM) {
              const float4 c4 = pre[it];
              o[0] += c4.x; o[1] += c4.y; o[2] += c4.z; o[3] += c4.w;
            }
            *reinterpret_cast<float4*>(E.out + row * E.ldo + cbase) = make_float4(o[0], o[1], o[2], o[3]);
          }
          if (kStats) {
#pragma unroll
            for (int j = 0; j < 4; ++j) { st0[p][j] += (double)f0[j]; st1[p][j] += (double)f1[j]; }
          }
        }
      }
      __syncthreads();   // staging free for the next pass / tile
    }
    tc_fence_before();
    __syncthreads();     // TMEM accumulator and A tile free for the next tile
    tc_fence_after();
  }

  // ---- flush BN statistics ---------------------------------------------------------------------
  if (kStats) {
#pragma unroll
    for (int p = 0; p < MAXPASS; ++p) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        double a = st0[p][j], b = st1[p][j];
        a += __shfl_xor_sync(0xffffffffu, a, 16);   // lanes l and l+16 own the same columns
        b += __shfl_xor_sync(0xffffffffu, b, 16);
        const int cl = p * 64 + e_c4 * 4 + j;
        if (lane < 16 && cl < Nt) {
          atomicAdd(E.stats + n0 + cl, a);
          atomicAdd(E.stats + N + n0 + cl, b);
        }
      }
    }
  }
  __syncthreads();
  if (warp == 0) tc_dealloc(tmem_base, (uint32_t)tmem_cols);
}

// Pre-split weight image for chunked K: per (N tile ny, chunk kc) one block [hi | lo], each [KC/8 groups][Npad rows][8]
// bf16 with the kernel's padded group stride, i.e. exactly the shared-memory image of load_w -- so the GEMM kernel
// stages a chunk with ONE bulk copy instead of re-splitting (KC x Nt) fp32 weights per 128-row tile.
__global__ void __launch_bounds__(256)
wprep_kernel(const float* __restrict__ Bm, long long ldb, int bT, int N, int K, int KC, int NtMax,
             unsigned char* __restrict__ img) {
  const TcSmem sp = tc_smem_plan(KC, NtMax, true);
  const int k8c = sp.Kp / 8;
  const int kc = blockIdx.x, ny = blockIdx.y, nkc = gridDim.x;
  const int n0 = ny * NtMax;
  const int Nt = (N - n0 < NtMax) ? (N - n0) : NtMax;
  const size_t bbytes = (size_t)k8c * sp.b_group_bytes;
  unsigned char* hi_blk = img + ((size_t)ny * nkc + kc) * 2 * bbytes;
  unsigned char* lo_blk = hi_blk + bbytes;
  for (int e = threadIdx.x; e < sp.Npad * k8c; e += blockDim.x) {
    const int n = e % sp.Npad, g = e / sp.Npad;
    float w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = kc * KC + g * 8 + i;
      w[i] = (n < Nt && k < K) ? (bT ? Bm[(long long)(n0 + n) * ldb + k] : Bm[(long long)k * ldb + n0 + n]) : 0.f;
    }
    uint4 hi, lo;
    split8(w, hi, lo);
    *reinterpret_cast<uint4*>(hi_blk + (size_t)g * sp.b_group_bytes + n * 16) = hi;
    *reinterpret_cast<uint4*>(lo_blk + (size_t)g * sp.b_group_bytes + n * 16) = lo;
  }
  for (int g = threadIdx.x; g < k8c; g += blockDim.x) {     // the 16 pad bytes of every group travel with the bulk copy
    *reinterpret_cast<uint4*>(hi_blk + (size_t)g * sp.b_group_bytes + sp.Npad * 16) = make_uint4(0, 0, 0, 0);
    *reinterpret_cast<uint4*>(lo_blk + (size_t)g * sp.b_group_bytes + sp.Npad * 16) = make_uint4(0, 0, 0, 0);
  }
}

struct TcPlan {
  int KC, NtMax, ntiles_n, npass, tmem_cols;
  size_t smem;
  bool two;
  int minb;
};
TcPlan tc_plan(int N, int K) {
  TcPlan p;
  const int Kp = (K + 15) / 16 * 16;
  p.KC = (K <= 128) ? Kp : 64;
  p.NtMax = N <= 256 ? N : 256;     // one MMA covers up to 256 columns: the A tile is synthesised once per 256
  p.ntiles_n = (N + p.NtMax - 1) / p.NtMax;
  p.npass = (p.NtMax + 63) / 64;
  p.tmem_cols = 64;
  while (p.tmem_cols < p.npass * 64) p.tmem_cols <<= 1;
  p.smem = tc_smem_plan(p.KC, p.NtMax, true).total;
  p.two = p.smem <= 110 * 1024;
  p.minb = p.smem <= 72 * 1024 ? 3 : (p.two ? 2 : 1);
  if (p.npass == 2 && p.minb > 2) p.minb = 2;   // matches the instantiations launched below
  if (p.npass > 2) p.minb = p.two ? 2 : 1;
  return p;
}

// bytes of the pre-split weight image (0 when the weights stay resident in shared memory: single K chunk)
size_t tc_wimg_bytes(int N, int K) {
  const TcPlan p = tc_plan(N, K);
  const int nkc = (K + p.KC - 1) / p.KC;
  if (nkc <= 1) return 0;
  const TcSmem sp = tc_smem_plan(p.KC, p.NtMax, true);
  return (size_t)p.ntiles_n * nkc * 2 * (sp.Kp / 8) * sp.b_group_bytes;
}

// aligned fast path for the A operand? (otherwise the element-wise loader is used inside the same kernel)
bool tc_operand_fast(const Operand& A, int amode, int K) {
  if (K % 8 != 0) return false;
  if (amode == OP_DY_SPARSE)
    return A.c1 && aligned16(A.y) && (A.ldy % 4) == 0 && aligned16(A.dg) && aligned16(A.amax) && (A.C % 8) == 0 && A.npts >= 1;
  if (amode == OP_DY_MAXK)
    return A.c1 && aligned16(A.y) && (A.ldy % 4) == 0 && aligned16(A.p) && (A.ld % 4) == 0 && (A.C % 8) == 0 &&
           aligned16(A.sc) && aligned16(A.sh);
  if (!aligned16(A.p) || (A.ld % 4) != 0) return false;
  if (amode == OP_EDGE && (A.C / 2) % 8 != 0) return false;
  if (amode == OP_DY && A.c1 && (!aligned16(A.y) || (A.ldy % 4) != 0)) return false;
  if (amode == OP_BNRELU && A.dmask && (!aligned16(A.dmask) || (A.C % 4) != 0)) return false;
  return true;
}

bool tc_supported(const Operand& A, int amode, const float* Bm, long long M, int N, int K, const Epilogue& E, int emode) {
  if (K < 12) return false;
  if ((amode == OP_EDGE || amode == OP_DY_MAXK) && (A.k < 1 || A.k > 32768 || A.npts < 1)) return false;   // RowMap range
  if (emode == EPI_EDGE_SCATTER && (E.k < 1 || E.k > 32768 || E.npts < 1)) return false;
  if (N % 4 != 0 || N < 16) return false;
  const TcPlan pl = tc_plan(N, K);
  const int k8c = ((pl.KC + 15) / 16 * 16) / 8;
  if (TC_THREADS % k8c != 0) return false;
  if (pl.smem > 200 * 1024) return false;
  if (emode == EPI_EDGE_SCATTER) {
    if (N != 128 || K > 128 || (E.lddx % 4) != 0 || !aligned16(E.dx)) return false;
  } else {
    if (!aligned16(E.out) || (E.ldo % 4) != 0) return false;
    if (emode == EPI_RELUMASK_STATS && (!aligned16(E.yprev) || (E.ldyp % 4) != 0)) return false;
    if (emode == EPI_RELUMASK_STATS && E.dmask && !aligned16(E.dmask)) return false;
  }
  return true;
}

// ------------------------------------------------------------------ warp-specialised row GEMM (chunked K, P-row layers) ---
// The kernel above runs "synthesise A -> MMA -> epilogue" one after the other inside a CTA.  For the per-point layers
// (K = 192 ... 512, 256-column tiles) that leaves the tensor pipe at 12-21 % (profiles/r1_gemm_summary.md).  Here one
// persistent CTA per SM splits the roles (17 warps):
//   * warps 0-7  (producers): every thread streams ITS OWN 2 rows x 8 channels of each 32-channel chunk into a private
//     shared-memory slot with cp.async (48-64 KB in flight per SM, no registers held, no cross-thread hand-over), then
//     applies the operand map (BN+ReLU / BN-backward affine), splits into bf16 hi + lo and writes the UMMA image of a
//     2-stage ring; one release-arrival per warp;
//   * warp 16 (one thread): waits for operand image + weight chunk, runs the generic->async proxy fence, issues the chunk's
//     6 tcgen05.mma (hi*hi + lo*hi + hi*lo) and refills the 2-3 stage ring of pre-split weight chunks (wprep_kernel image)
//     with one bulk copy per chunk.  Every CTA walks the K chunks in its own rotation so that the 148 SMs do not all pull
//     the same 33 KB chunk from the same L2 slices at the same time;
//   * warps 8-15 (epilogue): drain the finished 128 x <=256 accumulator (two TMEM accumulators alternate, so the next tile's
//     MMAs run under this tile's epilogue) in warp-private 32 x 32 blocks -- no block-wide barrier -- with the fused
//     epilogue (bias / per-cloud row bias / BN sums / ReLU mask + BN-backward sums / pooled arg-max keys).
// Measured at cfg-3 (profiles/r2_rowgemm_summary.md): tensor pipe 30 -> 45 %; the limiter left is the producers' instruction
// stream (~17 instructions per operand value at 2 rows per thread and chunk).
constexpr int WS_THREADS = 544;                 // 8 producer warps + 8 epilogue warps + 1 MMA / weight-copy warp
constexpr int WS_PROD = 256;
constexpr int WS_EPI = 256;
constexpr int WS_KC = 32;                       // channels per ring stage
constexpr int WS_NA = 2;                        // operand-image ring
constexpr int WS_NA_IMG = 4;                    // ... when the images arrive ready-made (OP_IMG): raw area + ring = 4 stages
constexpr int WS_NB_1 = 3, WS_NB_2 = 2;          // weight-chunk ring: one-operand maps / OP_DY (two operand streams to land)
constexpr int WS_AGRP = TILE_M * 16 + 32;       // image group stride: 4 groups x 2 rows per quarter-warp hit 8 distinct 16-B banks
constexpr int WS_RAW_BYTES_1 = 48 * 1024;       // cp.async landing slots (thread-private): 3 chunks of one operand in flight
constexpr int WS_RAW_BYTES_2 = 64 * 1024;       // ... or 2 chunks of OP_DY's two operands
constexpr int WS_STAGE_LD = 36;                 // epilogue staging pitch (floats): float4 rows land on distinct banks

struct WsSmem {
  int b_group_bytes;
  uint32_t bstage;
  size_t off_raw, off_a, off_b, off_stage, off_bar, off_pool, total;
};
__host__ __device__ inline WsSmem ws_smem_plan(int NtMax, bool two_ops) {
  WsSmem s;
  const int Npad = (NtMax + 15) / 16 * 16;
  s.b_group_bytes = Npad * 16 + 16;
  s.bstage = 2u * (WS_KC / 8) * (uint32_t)s.b_group_bytes;      // [hi | lo] of one chunk, as wprep_kernel writes it
  size_t o = 0;
  s.off_raw = o; o += two_ops ? WS_RAW_BYTES_2 : WS_RAW_BYTES_1;
  s.off_a = o; o += (size_t)WS_NA * 2 * (WS_KC / 8) * WS_AGRP;
  s.off_b = o; o += ((size_t)(two_ops ? WS_NB_2 : WS_NB_1) * s.bstage + 127) / 128 * 128;
  s.off_stage = o; o += (size_t)8 * 32 * WS_STAGE_LD * 4;      // one 32 x 32 staging tile per epilogue warp
  s.off_bar = o; o += 256;
  s.off_pool = o; o += 8 * 128 * 8;                 // EPI_STATS_POOL: running (value, row) keys, 128 columns per epilogue warp
  s.total = o;
  return s;
}

// order-preserving map float -> uint32 (and back) for the packed (value, row) keys of EPI_STATS_POOL
__host__ __device__ __forceinline__ uint32_t f32_ordered(float f) {
#ifdef __CUDA_ARCH__
  const uint32_t u = __float_as_uint(f);
#else
  uint32_t u; memcpy(&u, &f, 4);
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_) : "memory"); }

template <int AMODE, int EMODE, int MAXPASS>
__global__ void __launch_bounds__(WS_THREADS, 1)
rowgemm_ws_kernel(const Operand A, long long M, int N, int K, const Epilogue E, int num_tiles, int NtMax, int ntn,
                  const unsigned char* __restrict__ wimg) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int NOPS = (AMODE == OP_DY) ? 2 : 1;
  constexpr int WS_NB = (NOPS == 2) ? WS_NB_2 : WS_NB_1;
  // OP_IMG: the operand arrives as ready-made chunk images (wspc_rows_image) by bulk copy -- the producer warps have nothing
  // to do, and the cp.async landing area joins the image ring (4 stages instead of 2)
  constexpr bool IMG = (AMODE == OP_IMG);
  constexpr int NA_ = IMG ? WS_NA_IMG : WS_NA;
  const WsSmem sp = ws_smem_plan(NtMax, NOPS == 2);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + sp.off_bar);
  uint64_t *fullA = bars, *emptyA = bars + 4, *fullB = bars + 8, *emptyB = bars + 12, *accFull = bars + 16, *accEmpty = bars + 18;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
  unsigned char* const a_ring = smem + (IMG ? sp.off_raw : sp.off_a);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nkc = K / WS_KC;
  // CTA -> (column tile ny, every tstride-th row tile): neighbouring CTAs work on the same rows at the same time (the A rows
  // of the second column tile come from L2) and a CTA's BN statistics stay with fixed columns
  // the pooled epilogue walks CONTIGUOUS row tiles (tstride = 1) so that a CTA sees a cloud's tiles one after the other and
  // keeps the running per-column extreme on chip; the other epilogues take every (grid / ntn)-th tile
  constexpr bool kContig = (EMODE == EPI_STATS_POOL);
  const int ngroups = gridDim.x / ntn;
  const int per_cta = (num_tiles + ngroups - 1) / ngroups;
  const int ny = blockIdx.x % ntn, tile0 = kContig ? (blockIdx.x / ntn) * per_cta : blockIdx.x / ntn;
  const int tstride = kContig ? 1 : ngroups;
  const int n_items = kContig ? (num_tiles > tile0 ? (num_tiles - tile0 < per_cta ? num_tiles - tile0 : per_cta) : 0)
                              : ((num_tiles > tile0) ? (num_tiles - tile0 + tstride - 1) / tstride : 0);
  const int n0 = ny * NtMax;
  const int Nt = (N - n0 < NtMax) ? (N - n0) : NtMax;
  const int Ntp = (Nt + 15) / 16 * 16;
  // every CTA walks the K chunks of a tile in its own rotation: all SMs fetching the SAME 33 KB weight chunk at the same time
  // concentrates 148 readers on the few L2 slices that hold it (measured: the MMA thread waited on the weight ring 2/3 of the time)
  const int rot = (int)(blockIdx.x / ntn) % nkc;

  if (warp == 8) tc_alloc(tmem_slot, 512u);
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) {
      mbar_init(fullB + i, 1);
      mbar_init(emptyB + i, 1);
      mbar_init(fullA + i, IMG ? 1 : WS_PROD / 32);  // one arrival per producer warp / the bulk copy's expect_tx
      mbar_init(emptyA + i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(accFull + i, 1);
      mbar_init(accEmpty + i, WS_EPI / 32);
    }
    mbar_fence_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (tid < WS_PROD) {
    // ======================================================================== producers =====
    if constexpr (!IMG) {
    constexpr int DEPTH = (NOPS == 2 ? WS_RAW_BYTES_2 : WS_RAW_BYTES_1) / (NOPS * 4 * WS_PROD * 16);   // chunks in flight: 3 / 2
    // thread = (16-byte piece pc of the chunk's 128-byte rows, rows rl + 4j): every cp.async instruction of a warp covers four
    // whole 128-byte lines (the group-per-thread mapping touched eight half-used lines: twice the shared-memory write passes)
    const int pc = lane & 7, rl = (tid >> 5) * 16 + (lane >> 3);
    unsigned char* raw = smem + sp.off_raw;
    auto slot = [&](int d, int piece) { return raw + ((size_t)(d * (NOPS * 4) + piece) * WS_PROD + tid) * 16; };
    const bool two_ops = (AMODE == OP_DY) && A.c1 != nullptr;
    const int Q = n_items * nkc;
    // fetch cursor: chunk index inside the tile (already rotated) and the thread's first row of the tile
    int f_kc = rot, f_left = nkc;
    long long f_row = (long long)tile0 * TILE_M + rl;
    const long long row_step = (long long)tstride * TILE_M;
    auto issue = [&](int q) {
      if (q < Q) {
        const int c0 = f_kc * WS_KC + pc * 4;
        const int d = q % DEPTH;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const long long row = f_row + 4 * j;
          if (row < M) {
            cp_async16(slot(d, j), A.p + row * A.ld + c0);
            if (NOPS == 2 && two_ops) cp_async16(slot(d, 4 + j), A.y + row * A.ldy + c0);
          }
        }
        if (++f_kc == nkc) f_kc = 0;
        if (--f_left == 0) { f_left = nkc; f_row += row_step; }
      }
      cp_async_commit();
    };
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) issue(d);
    const uint32_t a_bytes_half = (WS_KC / 8) * WS_AGRP;
    unsigned char* const a_img = a_ring + (size_t)(pc >> 1) * WS_AGRP + (pc & 1) * 8;
    // consume cursor
    int b_kc = rot, b_left = nkc;
    long long b_row = (long long)tile0 * TILE_M + rl;
    // one chunk: raw slots -> operand map -> bf16 hi / lo image of ring stage s
    auto build = [&](int q, int s, int kcr, long long row_a) {
      const int c0 = kcr * WS_KC + pc * 4;
      float4 k0 = make_float4(0.f, 0.f, 0.f, 0.f), k1 = k0, k2 = k0;
      if (AMODE == OP_BNRELU) {
        k0 = *reinterpret_cast<const float4*>(A.sc + c0);
        k1 = *reinterpret_cast<const float4*>(A.sh + c0);
      }
      if (AMODE == OP_DY && two_ops) {
        k0 = *reinterpret_cast<const float4*>(A.c1 + c0);
        k1 = *reinterpret_cast<const float4*>(A.c2 + c0);
        k2 = *reinterpret_cast<const float4*>(A.c3 + c0);
      }
      const int d = q % DEPTH;
      float4 a[4], b[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        a[j] = *reinterpret_cast<const float4*>(slot(d, j));
        if (NOPS == 2) b[j] = *reinterpret_cast<const float4*>(slot(d, 4 + j));
      }
      unsigned char* a_hi = a_img + (size_t)s * 2 * a_bytes_half;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v[4] = {a[j].x, a[j].y, a[j].z, a[j].w};
        if (AMODE == OP_BNRELU) {
          v[0] = fmaxf(fmaf(v[0], k0.x, k1.x), 0.f); v[1] = fmaxf(fmaf(v[1], k0.y, k1.y), 0.f);
          v[2] = fmaxf(fmaf(v[2], k0.z, k1.z), 0.f); v[3] = fmaxf(fmaf(v[3], k0.w, k1.w), 0.f);
        }
        if (AMODE == OP_DY && two_ops) {
          v[0] = fmaf(k0.x, v[0], fmaf(k2.x, b[j].x, k1.x)); v[1] = fmaf(k0.y, v[1], fmaf(k2.y, b[j].y, k1.y));
          v[2] = fmaf(k0.z, v[2], fmaf(k2.z, b[j].z, k1.z)); v[3] = fmaf(k0.w, v[3], fmaf(k2.w, b[j].w, k1.w));
        }
        if (row_a + 4 * j >= M) { v[0] = 0.f; v[1] = 0.f; v[2] = 0.f; v[3] = 0.f; }
        uint2 hi, lo;
        split4(v, hi, lo);
        *reinterpret_cast<uint2*>(a_hi + (rl + 4 * j) * 16) = hi;
        *reinterpret_cast<uint2*>(a_hi + a_bytes_half + (rl + 4 * j) * 16) = lo;
      }
    };
    auto advance = [&]() {
      if (++b_kc == nkc) b_kc = 0;
      if (--b_left == 0) { b_left = nkc; b_row += row_step; }
    };
#pragma unroll 1
    for (int q = 0; q < Q; ++q) {
      const int s = q & 1;
      cp_async_wait<DEPTH - 1>();
      if (lane == 0) mbar_wait(emptyA + s, ((uint32_t)(q >> 1) & 1u) ^ 1u);   // the MMAs that read this stage two chunks ago retired
      __syncwarp();
      build(q, s, b_kc, b_row);
      advance();
      // one release-arrival per warp; the generic -> async proxy fence is executed by the MMA thread after its acquire (a
      // fence.proxy.async here would also wait for this thread's cp.async prefetches)
      __syncwarp();
      if (lane == 0) mbar_arrive(fullA + s);
      issue(q + DEPTH);                                                 // refill the slot just consumed
    }
    cp_async_wait<0>();
    }   // !IMG
  } else if (tid >= WS_PROD + WS_EPI) {
    // ================================================================ MMA issue + weight ring =====
    if (lane == 0) {
      const uint32_t bstage = sp.bstage;
      unsigned char* sB = smem + sp.off_b;
      const unsigned char* wsrc = wimg + (size_t)ny * nkc * bstage;     // this CTA's column tile: chunk kc at wsrc + kc*bstage
      const int Q = n_items * nkc;
      for (int c = 0; c < WS_NB && c < Q; ++c) {
        mbar_expect_tx(fullB + c, bstage);
        bulk_g2s(sB + (size_t)c * bstage, wsrc + (size_t)((c + rot) % nkc) * bstage, bstage, fullB + c);
      }
      const uint32_t idesc = umma_idesc(Ntp);
      const uint32_t a_bytes_half = (WS_KC / 8) * WS_AGRP;
      // OP_IMG: operand chunk c of this CTA = image block (row tile, rotated chunk), one bulk copy into ring stage c % NA_
      const unsigned char* aimg = reinterpret_cast<const unsigned char*>(A.p);
      const uint32_t astage = 2u * a_bytes_half;
      int ia_item = 0, ia_kc = 0;                                        // cursor of the next operand chunk to fetch
      auto fetch_a = [&](int c) {
        const int cs = c % NA_;
        const size_t blk = (size_t)(tile0 + ia_item * tstride) * nkc + (size_t)((ia_kc + rot) % nkc);
        mbar_expect_tx(fullA + cs, astage);
        bulk_g2s(a_ring + (size_t)cs * astage, aimg + blk * astage, astage, fullA + cs);
        if (++ia_kc == nkc) { ia_kc = 0; ++ia_item; }
      };
      if (IMG) {
        for (int c = 0; c < NA_ && c < Q; ++c) fetch_a(c);
      }
      int item = 0, kc = 0;
#pragma unroll 1
      for (int q = 0; q < Q; ++q) {
        const int s = q % NA_, acc = item & 1, sb = q % WS_NB;
        if (kc == 0) mbar_wait(accEmpty + acc, ((uint32_t)(item >> 1) & 1u) ^ 1u);   // the epilogue drained this accumulator
        mbar_wait(fullB + sb, (uint32_t)(q / WS_NB) & 1u);
        mbar_wait(fullA + s, (uint32_t)(q / NA_) & 1u);
        if (!IMG) fence_proxy_async_smem();
        tc_fence_after();
        const uint32_t ah = smem_u32(a_ring + (size_t)s * 2 * a_bytes_half), al = ah + a_bytes_half;
        const uint32_t bh = smem_u32(sB + (size_t)sb * bstage), bl = bh + (WS_KC / 8) * (uint32_t)sp.b_group_bytes;
        const uint32_t dcol = tmem_base + (uint32_t)acc * 256u;
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {
          const uint32_t ab = (pass == 1) ? al : ah;
          const uint32_t bb = (pass == 2) ? bl : bh;
#pragma unroll
          for (int kk = 0; kk < WS_KC / 16; ++kk) {
            const uint64_t ad = umma_desc(ab + (uint32_t)(2 * kk) * WS_AGRP, WS_AGRP, 128);
            const uint64_t bd = umma_desc(bb + (uint32_t)(2 * kk) * sp.b_group_bytes, sp.b_group_bytes, 128);
            tc_mma_bf16(dcol, ad, bd, idesc, (kc | pass | kk) ? 1u : 0u);
          }
        }
        tc_commit(emptyA + s);
        tc_commit(emptyB + sb);
        if (kc == nkc - 1) tc_commit(accFull + acc);
        const int c = q + WS_NB - 1;                                    // weight chunk q+NB-1 goes where chunk q-1 was
        if (c >= WS_NB && c < Q) {
          const int cs = c % WS_NB;
          mbar_wait(emptyB + cs, ((uint32_t)(c / WS_NB) & 1u) ^ 1u);
          mbar_expect_tx(fullB + cs, bstage);
          bulk_g2s(sB + (size_t)cs * bstage, wsrc + (size_t)((c + rot) % nkc) * bstage, bstage, fullB + cs);
        }
        if (IMG) {
          const int ca = q + NA_ - 1;                                   // operand chunk q+NA-1 goes where chunk q-1 was
          if (ca >= NA_ && ca < Q) {
            mbar_wait(emptyA + ca % NA_, ((uint32_t)(ca / NA_) & 1u) ^ 1u);
            fetch_a(ca);
          }
        }
        if (++kc == nkc) { kc = 0; ++item; }
      }
    }
  } else {
    // ========================================================================= epilogue =====
    // Every warp drains its own 32-row x 128-column part of the accumulator in 32 x 32 blocks through a warp-private
    // staging tile (TMEM lane = row -> shared -> 4 rows x 128 B per store instruction): no block-wide barrier anywhere, so
    // the eight warps drift apart and hide each other's TMEM / store latencies.
    const int et = tid - WS_PROD, ew = et >> 5;
    const int lq = ew & 3, chh = ew >> 2;         // TMEM lane quadrant (hardware warp id % 4) / 128-column half of the tile
    float* stage_w = reinterpret_cast<float*>(smem + sp.off_stage) + ew * (32 * WS_STAGE_LD);
    const int sr = lane >> 3, cq = lane & 7;      // store phase: sub-row within a group of 4 rows, 4-column group
    constexpr bool kStats = (EMODE == EPI_STORE_STATS || EMODE == EPI_RELUMASK_STATS || EMODE == EPI_STATS_POOL);
    constexpr int NBLK = 4;
    double st0[kStats ? NBLK : 1][4], st1[kStats ? NBLK : 1][4];
    if (kStats) {
#pragma unroll
      for (int p = 0; p < NBLK; ++p)
#pragma unroll
        for (int j = 0; j < 4; ++j) { st0[p][j] = 0.0; st1[p][j] = 0.0; }
    }
    // EPI_STATS_POOL: lanes with sr == 0 own 16 columns (4 blocks x 4) of this warp's running keys for the current cloud
    unsigned long long* pool_run = reinterpret_cast<unsigned long long*>(smem + sp.off_pool) + ew * 128 + cq * 4;
    long long pool_cloud = -1;
    auto pool_flush = [&]() {
      if (EMODE == EPI_STATS_POOL && sr == 0 && pool_cloud >= 0) {
        unsigned long long* keys = reinterpret_cast<unsigned long long*>(E.dx) + pool_cloud * N + n0 + chh * 128 + cq * 4;
#pragma unroll
        for (int nb = 0; nb < NBLK; ++nb)
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (chh * 128 + nb * 32 + cq * 4 + j < Nt && pool_run[nb * 32 + j] != 0ull) atomicMax(keys + nb * 32 + j, pool_run[nb * 32 + j]);
      }
    };
    if (EMODE == EPI_STATS_POOL && sr == 0) {
#pragma unroll
      for (int nb = 0; nb < NBLK; ++nb)
#pragma unroll
        for (int j = 0; j < 4; ++j) pool_run[nb * 32 + j] = 0ull;
    }
    int last_blk = -1;                            // last 32-column block of this warp's half that holds columns
#pragma unroll
    for (int nb = 0; nb < NBLK; ++nb)
      if (chh * 128 + nb * 32 < Nt) last_blk = nb;
#pragma unroll 1
    for (int item = 0; item < n_items; ++item) {
      const long long row0 = (long long)(tile0 + item * tstride) * TILE_M;
      const int acc = item & 1;
      if (lane == 0) mbar_wait(accFull + acc, (uint32_t)(item >> 1) & 1u);
      __syncwarp();
      tc_fence_after();
      const uint32_t tacc = tmem_base + (uint32_t)acc * 256u + ((uint32_t)(lq * 32) << 16);
      if (last_blk < 0) {                         // this warp's column half is empty: nothing to read, hand the accumulator back
        if (lane == 0) mbar_arrive(accEmpty + acc);
        continue;
      }
      if (EMODE == EPI_STATS_POOL) {              // a new cloud starts: publish the finished one, restart the running keys
        const long long cloud = row0 / E.npts;
        if (cloud != pool_cloud) {
          pool_flush();
          if (sr == 0) {
#pragma unroll
            for (int nb = 0; nb < NBLK; ++nb)
#pragma unroll
              for (int j = 0; j < 4; ++j) pool_run[nb * 32 + j] = 0ull;
          }
          pool_cloud = cloud;
        }
      }
      long long rb_cloud0 = 0;
      uint32_t rb_rem0 = 0;
      if ((EMODE == EPI_STORE || EMODE == EPI_STORE_STATS) && E.rowbias) {
        rb_cloud0 = div_pos(row0, E.rb_rows);
        rb_rem0 = (uint32_t)(row0 - rb_cloud0 * E.rb_rows);
      }
#pragma unroll
      for (int nb = 0; nb < NBLK; ++nb) {
        const int c0 = chh * 128 + nb * 32;       // first column of the block inside this N tile
        if (c0 >= Nt) break;
        {
          float v[32];
          tc_ld32(tacc + (uint32_t)c0, v);
          float* srow = stage_w + lane * WS_STAGE_LD;
#pragma unroll
          for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(srow + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
        if (nb == last_blk) {                     // every tcgen05.ld of this warp on the accumulator has completed: hand it back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(accEmpty + acc);
        }
        __syncwarp();
        const int cl = c0 + cq * 4;               // column within this N tile
        const int cbase = n0 + cl;                // global column
        if (cl < Nt) {
          float bias[4] = {0.f, 0.f, 0.f, 0.f}, scp[4] = {0.f, 0.f, 0.f, 0.f}, shp[4] = {-1.f, -1.f, -1.f, -1.f};
          if ((EMODE == EPI_STORE || EMODE == EPI_STORE_STATS || EMODE == EPI_STATS_POOL) && E.bias) {
#pragma unroll
            for (int j = 0; j < 4; ++j) bias[j] = E.bias[cbase + j];
          }
          if (EMODE == EPI_RELUMASK_STATS) {
#pragma unroll
            for (int j = 0; j < 4; ++j) { scp[j] = E.scp[cbase + j]; shp[j] = E.shp[cbase + j]; }
          }
          float f0[4] = {0.f, 0.f, 0.f, 0.f}, f1[4] = {0.f, 0.f, 0.f, 0.f};
          const int rbase = lq * 32 + sr;         // tile-local row of iteration 0; + 4 per iteration
          if (EMODE == EPI_STATS_POOL) {
            // BN (gamma * invstd > 0 or < 0) and ReLU are monotone per column: the pooled maximum over the cloud's points sits at
            // the row with the largest (gamma >= 0) or smallest (gamma < 0) pre-BN value; first row on ties (max_pool2d's arg-max)
            float sgn[4], best[4];
            int brow[4] = {0, 0, 0, 0};
#pragma unroll
            for (int j = 0; j < 4; ++j) { sgn[j] = (E.scp[cbase + j] < 0.f) ? -1.f : 1.f; best[j] = -INFINITY; }
            const uint32_t rin0 = (uint32_t)(row0 % E.npts);              // tiles never straddle clouds (npts % 128 == 0)
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              const int r = rbase + 4 * it;
              if (row0 + r >= M) break;
              const float4 s4 = *reinterpret_cast<const float4*>(stage_w + (it * 4 + sr) * WS_STAGE_LD + cq * 4);
              const float o[4] = {s4.x + bias[0], s4.y + bias[1], s4.z + bias[2], s4.w + bias[3]};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                f0[j] += o[j];
                f1[j] = fmaf(o[j], o[j], f1[j]);
                const float ys = o[j] * sgn[j];
                if (ys > best[j]) { best[j] = ys; brow[j] = r; }
              }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              unsigned long long key = (row0 + rbase < M)
                  ? ((unsigned long long)f32_ordered(best[j]) << 32) | (unsigned long long)(0xFFFFFFFFu - (rin0 + (uint32_t)brow[j]))
                  : 0ull;
              unsigned long long other = __shfl_xor_sync(0xffffffffu, key, 8);
              if (other > key) key = other;
              other = __shfl_xor_sync(0xffffffffu, key, 16);
              if (other > key) key = other;
              if (sr == 0 && key > pool_run[nb * 32 + j]) pool_run[nb * 32 + j] = key;
            }
          } else {
            float4 pre[8];
            if (EMODE == EPI_RELUMASK_STATS) {
#pragma unroll
              for (int it = 0; it < 8; ++it) {
                const long long row = row0 + rbase + 4 * it;
                pre[it] = (row < M) ? *reinterpret_cast<const float4*>(E.yprev + row * E.ldyp + cbase) : make_float4(0.f, 0.f, 0.f, 0.f);
              }
            }
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              const int r = rbase + 4 * it;
              const long long row = row0 + r;
              if (row >= M) break;
              const float4 s4 = *reinterpret_cast<const float4*>(stage_w + (it * 4 + sr) * WS_STAGE_LD + cq * 4);
              float o[4] = {s4.x, s4.y, s4.z, s4.w};
              if (EMODE == EPI_STORE || EMODE == EPI_STORE_STATS) {
                const float* rb = nullptr;
                if (E.rowbias) {
                  long long cl0 = rb_cloud0;
                  const uint32_t x = rb_rem0 + (uint32_t)r;
                  if (x >= (uint32_t)E.rb_rows) { cl0 += x / (uint32_t)E.rb_rows; }
                  rb = E.rowbias + cl0 * E.ldrb + cbase;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  o[j] += bias[j];
                  if (rb) o[j] += rb[j];
                }
                if (EMODE == EPI_STORE_STATS) {
#pragma unroll
                  for (int j = 0; j < 4; ++j) { f0[j] += o[j]; f1[j] = fmaf(o[j], o[j], f1[j]); }
                }
              } else if (EMODE == EPI_RELUMASK_STATS) {
                const float4 y4 = pre[it];
                const float yv[4] = {y4.x, y4.y, y4.z, y4.w};
                float dm[4] = {1.f, 1.f, 1.f, 1.f};
                if (E.dmask) {
                  const float4 m4 = *reinterpret_cast<const float4*>(E.dmask + row * N + cbase);
                  dm[0] = m4.x * E.dscale; dm[1] = m4.y * E.dscale; dm[2] = m4.z * E.dscale; dm[3] = m4.w * E.dscale;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const bool on = fmaf(yv[j], scp[j], shp[j]) > 0.f;
                  o[j] = on ? o[j] * dm[j] : 0.f;
                  f0[j] += o[j];
                  f1[j] = fmaf(o[j], yv[j], f1[j]);
                }
              }
              *reinterpret_cast<float4*>(E.out + row * E.ldo + cbase) = make_float4(o[0], o[1], o[2], o[3]);
            }
          }
          if (kStats) {
#pragma unroll
            for (int j = 0; j < 4; ++j) { st0[nb][j] += (double)f0[j]; st1[nb][j] += (double)f1[j]; }
          }
        }
        __syncwarp();        // the staging tile may be overwritten by the next block
      }
    }
    pool_flush();
    if (kStats) {
#pragma unroll
      for (int nb = 0; nb < NBLK; ++nb) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          double a = st0[nb][j], b = st1[nb][j];
          a += __shfl_xor_sync(0xffffffffu, a, 8);    // the four sub-row lanes of a column group
          b += __shfl_xor_sync(0xffffffffu, b, 8);
          a += __shfl_xor_sync(0xffffffffu, a, 16);
          b += __shfl_xor_sync(0xffffffffu, b, 16);
          const int cl = chh * 128 + nb * 32 + cq * 4 + j;
          if (sr == 0 && cl < Nt) {
            atomicAdd(E.stats + n0 + cl, a);
            atomicAdd(E.stats + N + n0 + cl, b);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tc_dealloc(tmem_base, 512u);
}

// the serial kernel stays the path for single-chunk layers, edge operands and short inputs; WSPC_ROWGEMM_KERNEL=serial forces it
bool ws_rowgemm_enabled() {
  static const bool on = [] {
    const char* e = getenv("WSPC_ROWGEMM_KERNEL");
    return !(e && strcmp(e, "serial") == 0);
  }();
  return on;
}

// ---- OP_IMG: pre-split operand image (wspc_rows_image) -------------------------------------------------------------
bool ws_img_supported(long long M, int N, int K) {
  if (!ws_rowgemm_enabled()) return false;
  const TcPlan pl = tc_plan(N, K);
  const long long num_tiles = (M + TILE_M - 1) / TILE_M;
  return K > 128 && K % WS_KC == 0 && N % 16 == 0 && pl.ntiles_n <= kNumSM / 2 && num_tiles >= 4 * kNumSM && tc_wimg_bytes(N, K) > 0;
}
size_t ws_img_bytes(long long M, int K) {
  return (size_t)((M + TILE_M - 1) / TILE_M) * (size_t)(K / WS_KC) * 2 * (WS_KC / 8) * WS_AGRP;
}

// grid = row tiles, block 256: thread = (channel group of the chunk, rows rr and rr + 64), as the producers of rowgemm_ws_kernel
__global__ void __launch_bounds__(256)
rows_image_kernel(const float* __restrict__ x, long long ldx, long long M, int K, unsigned char* __restrict__ img) {
  const int tid = threadIdx.x, kg = tid & 3, rr = tid >> 2;
  const int nkc = K / WS_KC;
  const long long row0 = (long long)blockIdx.x * TILE_M;
  constexpr uint32_t half = (WS_KC / 8) * WS_AGRP;
  for (int kc = 0; kc < nkc; ++kc) {
    unsigned char* blk = img + ((size_t)blockIdx.x * nkc + kc) * (2 * half);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = rr + 64 * i;
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = 0.f;
      if (row0 + r < M) ld8(x + (row0 + r) * ldx + kc * WS_KC + kg * 8, v);
      uint4 hi, lo;
      split8(v, hi, lo);
      *reinterpret_cast<uint4*>(blk + (size_t)kg * WS_AGRP + r * 16) = hi;
      *reinterpret_cast<uint4*>(blk + half + (size_t)kg * WS_AGRP + r * 16) = lo;
    }
    if (tid < 16)    // the 32 pad bytes of every group travel with the bulk copy: keep them defined
      *reinterpret_cast<uint4*>(blk + (size_t)(tid & 7) * WS_AGRP + TILE_M * 16 + (tid >> 3) * 16) = make_uint4(0, 0, 0, 0);
  }
}

template <int EMODE>
int launch_ws_img(const Operand& A, const float* Bm, long long ldb, int bT, long long M, int N, int K, const Epilogue& E, void* ws,
                  cudaStream_t st) {
  const TcPlan pl = tc_plan(N, K);
  const int num_tiles = (int)((M + TILE_M - 1) / TILE_M);
  wprep_kernel<<<dim3(K / WS_KC, pl.ntiles_n), 256, 0, st>>>(Bm, ldb, bT, N, K, WS_KC, pl.NtMax, static_cast<unsigned char*>(ws));
  count_launch();
  const WsSmem wp = ws_smem_plan(pl.NtMax, false);
  const int grid = kNumSM / pl.ntiles_n * pl.ntiles_n;
  auto kern = rowgemm_ws_kernel<OP_IMG, EMODE, 4>;
  WSPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wp.total));
  kern<<<grid, WS_THREADS, wp.total, st>>>(A, M, N, K, E, num_tiles, pl.NtMax, pl.ntiles_n, static_cast<const unsigned char*>(ws));
  count_launch();
  WSPC_LAUNCH_CHECK("rowgemm_ws_kernel(image operand)");
  return WSPC_OK;
}

// conv2d -> (BN -> ReLU) -> max over the cloud's points without writing the conv output: eligibility and launch
bool ws_pool_supported(long long M, int N, int K, int npts) {
  if (!ws_rowgemm_enabled()) return false;
  const TcPlan pl = tc_plan(N, K);
  const long long num_tiles = (M + TILE_M - 1) / TILE_M;
  return K > 128 && K % WS_KC == 0 && N % 16 == 0 && pl.ntiles_n <= kNumSM / 2 && num_tiles >= 4 * kNumSM && npts >= TILE_M &&
         npts % TILE_M == 0 && M % npts == 0 && tc_wimg_bytes(N, K) > 0;
}

__global__ void __launch_bounds__(256)
pool_keys_finish_kernel(const unsigned long long* __restrict__ keys, const float* __restrict__ gamma, const float* __restrict__ sc,
                        const float* __restrict__ sh, long long total, int C, float* __restrict__ g, int32_t* __restrict__ amax,
                        float* __restrict__ ymax) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  const unsigned long long key = keys[i];
  const uint32_t u = (uint32_t)(key >> 32);
  const float ys = __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
  const float y = (gamma[c] < 0.f) ? -ys : ys;
  g[i] = fmaxf(fmaf(y, sc[c], sh[c]), 0.f);
  amax[i] = (int32_t)(0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull));
  ymax[i] = y;
}

template <int AMODE>
int launch_ws_pool(const Operand& A, const float* Bm, long long ldb, long long M, int N, int K, const Epilogue& E, void* ws,
                   cudaStream_t st) {
  const TcPlan pl = tc_plan(N, K);
  const int num_tiles = (int)((M + TILE_M - 1) / TILE_M);
  wprep_kernel<<<dim3(K / WS_KC, pl.ntiles_n), 256, 0, st>>>(Bm, ldb, 0, N, K, WS_KC, pl.NtMax, static_cast<unsigned char*>(ws));
  count_launch();
  const WsSmem wp = ws_smem_plan(pl.NtMax, AMODE == OP_DY);
  const int grid = kNumSM / pl.ntiles_n * pl.ntiles_n;
  auto kern = rowgemm_ws_kernel<AMODE, EPI_STATS_POOL, 4>;
  WSPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wp.total));
  kern<<<grid, WS_THREADS, wp.total, st>>>(A, M, N, K, E, num_tiles, pl.NtMax, pl.ntiles_n, static_cast<const unsigned char*>(ws));
  count_launch();
  WSPC_LAUNCH_CHECK("rowgemm_ws_kernel(pool)");
  return WSPC_OK;
}

template <int AMODE, int EMODE>
int launch_tc(const Operand& A, const float* Bm, long long ldb, int bT, long long M, int N, int K, const Epilogue& E,
              void* ws, size_t ws_bytes, cudaStream_t st) {
  const int generic = tc_operand_fast(A, AMODE, K) ? 0 : 1;
  const TcPlan pl = tc_plan(N, K);
  const unsigned char* wimg = nullptr;
  const size_t need = tc_wimg_bytes(N, K);
  const int num_tiles = (int)((M + TILE_M - 1) / TILE_M);
  constexpr bool kWsModes = (AMODE == OP_PLAIN || AMODE == OP_BNRELU || AMODE == OP_DY) &&
                            (EMODE == EPI_STORE || EMODE == EPI_STORE_STATS || EMODE == EPI_RELUMASK_STATS);
  if constexpr (kWsModes) {
    if (ws_rowgemm_enabled() && need && ws && ws_bytes >= need && aligned16(ws) && !generic && K % WS_KC == 0 && pl.ntiles_n <= kNumSM / 2 &&
        num_tiles >= 4 * kNumSM && !(AMODE == OP_BNRELU && A.dmask)) {
      // pre-split weights in 32-channel chunks (same bytes as the 64-channel image when K % 64 == 0, fewer otherwise)
      wprep_kernel<<<dim3(K / WS_KC, pl.ntiles_n), 256, 0, st>>>(Bm, ldb, bT, N, K, WS_KC, pl.NtMax, static_cast<unsigned char*>(ws));
      count_launch();
      const WsSmem wp = ws_smem_plan(pl.NtMax, AMODE == OP_DY);
      const int grid = kNumSM / pl.ntiles_n * pl.ntiles_n;
      auto kern = rowgemm_ws_kernel<AMODE, EMODE, 4>;
      WSPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wp.total));
      kern<<<grid, WS_THREADS, wp.total, st>>>(A, M, N, K, E, num_tiles, pl.NtMax, pl.ntiles_n, static_cast<const unsigned char*>(ws));
      count_launch();
      WSPC_LAUNCH_CHECK("rowgemm_ws_kernel");
      return WSPC_OK;
    }
  }
  if (need && ws && ws_bytes >= need && aligned16(ws)) {     // chunked K: stage pre-split weights with bulk copies
    const int nkc = (K + pl.KC - 1) / pl.KC;
    wprep_kernel<<<dim3(nkc, pl.ntiles_n), 256, 0, st>>>(Bm, ldb, bT, N, K, pl.KC, pl.NtMax, static_cast<unsigned char*>(ws));
    count_launch();
    wimg = static_cast<const unsigned char*>(ws);
  }
  const int ctas = pl.minb * kNumSM;
  int gx = (ctas + pl.ntiles_n - 1) / pl.ntiles_n;
  if (gx > num_tiles) gx = num_tiles;
  if (gx < 1) gx = 1;
  dim3 grid(gx, pl.ntiles_n);
#define WSPC_TC_LAUNCH(MINB_, NP_)                                                                             \
  {                                                                                                            \
    auto kern = rowgemm_tc_kernel<AMODE, EMODE, MINB_, NP_>;                                                   \
    WSPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));          \
    kern<<<grid, TC_THREADS, pl.smem, st>>>(A, Bm, ldb, bT, M, N, K, E, num_tiles, pl.tmem_cols, pl.KC, pl.NtMax, generic, wimg); \
  }
  if (pl.npass <= 1) { if (pl.minb == 3) WSPC_TC_LAUNCH(3, 1) else if (pl.minb == 2) WSPC_TC_LAUNCH(2, 1) else WSPC_TC_LAUNCH(1, 1) }
  else if (pl.npass <= 2) { if (pl.minb >= 2) WSPC_TC_LAUNCH(2, 2) else WSPC_TC_LAUNCH(1, 2) }
  else { if (pl.minb >= 2) WSPC_TC_LAUNCH(2, 4) else WSPC_TC_LAUNCH(1, 4) }
#undef WSPC_TC_LAUNCH
  count_launch();
  WSPC_LAUNCH_CHECK("rowgemm_tc_kernel");
  return WSPC_OK;
}

// ------------------------------------------------------------------ weight gradient on tensor cores ---
// dW(K1,K2) = sum_rows A(row,:)^T dY(row,:).  The reduction runs over rows, so both operands are "MN-major"
// for the MMA (channels contiguous, rows = K).  The shared-memory image is the same [channel group][row][8]
// layout as above — read by the tensor core with the roles of LBO/SBO swapped — so the loaders are shared.
// Each CTA owns a slab of rows and one (128 x <=256) tile of dW, accumulates it in TMEM over the whole slab
// and writes a partial; slab_reduce_kernel adds the partials in fp64 in a fixed order (deterministic).
__device__ __forceinline__ uint32_t umma_idesc_mn(int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(TILE_M >> 4) << 24);
}

// FUSE: the same pass over the rows also produces the layer's DATA gradient dA(row, 0:Nf) = dY(row,:) * Wf^T (Wf = the
// layer's (Nf, K2) weight matrix) with the ReLU-mask / BN-sum epilogue of the producing layer (EPI_RELUMASK_STATS) -- the
// dY tile staged for the weight gradient is read a second time K-major by the tensor core, so G, y and the previous
// activation are read from HBM once for both gradients (Conv2DBackpropFilter + Conv2DBackpropInput of one conv2d).
struct FuseArgs {
  const float* Wf;      // (Nf, K2) row-major, leading dimension ldw
  long long ldw;
  int Nf;               // columns of dA (= the layer's input channels), Nf % 16 == 0, Nf <= 128
  Epilogue E;           // out, ldo, stats, yprev, ldyp, scp, shp (, dmask, dscale)
};

template <int AMODE, int GMODE, int MINB, bool FUSE>
__global__ void __launch_bounds__(TC_THREADS, MINB)
colgemm_tc_kernel(const Operand A, const Operand G, long long M, long long rows_per_slab, int K1p, int K2p, int K2t,
                  float* __restrict__ partial, float* __restrict__ partial_b, int tmem_cols, int genA, int genG,
                  int aGroups, int RT, const FuseArgs F) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int gGroups = K2t / 8;                         // power of two, <= 32
  // RT = rows per tile (the K extent of one accumulation step): 128, or 64 when that lets two CTAs share an SM
  const int agb = RT * 16 + 16;                        // bytes of one 8-channel group of a row-tile image
  // A images: aGroups = 8 when the operand has <= 64 channels (the M=128 MMA then also reads the 8 groups that follow --
  // the lo image / the dY image -- into accumulator lanes 64..127, which are never stored), else 16
  unsigned char* sAhi = smem;
  unsigned char* sAlo = sAhi + (size_t)aGroups * agb;
  unsigned char* sGhi = sAlo + (size_t)aGroups * agb;
  unsigned char* sGlo = sGhi + (size_t)gGroups * agb;
  const int w_group_bytes = FUSE ? F.Nf * 16 + 16 : 0;                  // fused data gradient: resident weight image
  unsigned char* sWhi = sGlo + (size_t)gGroups * agb;
  unsigned char* sWlo = sWhi + (size_t)gGroups * w_group_bytes;
  unsigned char* misc = sWlo + (size_t)gGroups * w_group_bytes;
  float* stage = reinterpret_cast<float*>(sAhi);                        // epilogue staging aliases the A images
  uint64_t* mma_bar = reinterpret_cast<uint64_t*>(misc);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 16);
  float* bred = reinterpret_cast<float*>(misc + 32);   // [K2t] bias-gradient reduction

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int k1_0 = blockIdx.y * 128 /*channels per dW tile = MMA M*/, k2_0 = blockIdx.z * 256;
  const long long r_begin = (long long)blockIdx.x * rows_per_slab;
  const long long r_end = (r_begin + rows_per_slab < M) ? r_begin + rows_per_slab : M;

  if (warp == 0) tc_alloc(tmem_slot, (uint32_t)tmem_cols);
  if (tid == 32) { mbar_init(mma_bar, 1); mbar_fence_init(); }
  for (int i = tid; i < K2t; i += TC_THREADS) bred[i] = 0.f;
  if (FUSE) {   // Wf^T as the K-major B operand: element (n, c) -> group c/8, row n, slot c%8 (as load_w of the row GEMM)
    for (int e = tid; e < F.Nf * gGroups; e += TC_THREADS) {
      const int n = e % F.Nf, g = e / F.Nf;
      float w[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = g * 8 + i;
        w[i] = (c < G.C) ? F.Wf[(long long)n * F.ldw + c] : 0.f;
      }
      uint4 hi, lo;
      split8(w, hi, lo);
      *reinterpret_cast<uint4*>(sWhi + (size_t)g * w_group_bytes + n * 16) = hi;
      *reinterpret_cast<uint4*>(sWlo + (size_t)g * w_group_bytes + n * 16) = lo;
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t idesc = umma_idesc_mn(K2t);
  const uint32_t idesc_f = FUSE ? umma_idesc(F.Nf) : 0u;
  const int e_c4 = tid & 15, e_r0 = tid >> 4;          // fused epilogue: 4 fixed columns, 16 rows per sweep
  double fst0[4] = {0.0, 0.0, 0.0, 0.0}, fst1[4] = {0.0, 0.0, 0.0, 0.0};

  // fixed per-thread channel groups
  const int gA = tid & (aGroups - 1), rA0 = (aGroups == 8) ? tid >> 3 : tid >> 4;   // A: aGroups (8 or 16) channel groups
  const int gG = tid % gGroups, rG0 = tid / gGroups;       // dY: gGroups groups
  const int rGstep = TC_THREADS / gGroups;
  const int cA = k1_0 / 8 + gA, cG = k2_0 / 8 + gG;        // absolute channel groups
  const bool vA = cA * 8 < A.C, vG = cG * 8 < G.C;
  float a0[8], a1[8], a2[8], g0[8], g1[8], g2[8];
  load_consts<AMODE>(A, cA, A.C, a0, a1, a2);
  load_consts<GMODE>(G, cG, G.C, g0, g1, g2);
  float bs[8];     // bias-gradient partial: <= rows_per_slab / rGstep values per thread, fp32 is ample (slabs add in fp64)
#pragma unroll
  for (int i = 0; i < 8; ++i) bs[i] = 0.f;

  uint32_t phase = 0, accum = 0;
  bool pending = false;
  const uint64_t a_inv = (AMODE == OP_EDGE) ? rowmap_inv(A.k) : 0;
  const uint64_t g_inv = (GMODE == OP_DY_MAXK) ? rowmap_inv(G.k) : 0;
  for (long long rb = r_begin; rb < r_end; rb += RT) {
    if (pending) { mbar_wait(mma_bar, phase); phase ^= 1; pending = false; }   // MMAs done reading smem
    RowMap arm, grm;
    CloudMap gcm;
    if (AMODE == OP_EDGE) arm = rowmap_tile(rb, A.k, A.npts, a_inv);
    if (GMODE == OP_DY_SPARSE) gcm = cloudmap_tile(rb, G.npts);
    if (GMODE == OP_DY_MAXK) grm = rowmap_tile(rb, G.k, G.npts, g_inv);
    // A operand: RT / rAstep chunks per thread in batches of UB, all loads of a batch in flight before the first use.
    // The group count is a compile-time constant of each copy of the loop (16: 16 rows per sweep, the batch of 4 never
    // leaves the tile; 8: 32 rows per sweep, every thread owns a group that exists).
    constexpr int UB = 4;
    auto stage_A = [&](auto ag_tag) {
      constexpr int AG = decltype(ag_tag)::value;
      constexpr int rAstep = TC_THREADS / AG;
      constexpr bool guard = rAstep * UB > 64;               // a batch may reach past a 64-row tile
      if (!genA) {
#pragma unroll 1
        for (int rbase = rA0; rbase < RT; rbase += rAstep * UB) {
          long long pt[UB], cb[UB], nb[UB];
          bool ok[UB];
#pragma unroll
          for (int u = 0; u < UB; ++u) {
            const int r = rbase + rAstep * u;
            const bool in = !guard || r < RT;
            ok[u] = vA && in && rb + r < r_end;
            pt[u] = 0; cb[u] = 0; nb[u] = 0;
            if (AMODE == OP_EDGE && in) {
              rowmap_point(arm, r, pt[u], cb[u]);
              if (ok[u] && cA * 8 >= (A.C >> 1)) nb[u] = cb[u] + A.idx[rb + r];
            }
          }
          RawChunk w[UB];
#pragma unroll
          for (int u = 0; u < UB; ++u) fetch_chunk<AMODE>(A, rb + rbase + rAstep * u, cA, ok[u], pt[u], nb[u], w[u]);
#pragma unroll
          for (int u = 0; u < UB; ++u) {
            const int r = rbase + rAstep * u;
            if (!guard || r < RT) {
              float v[8];
              finish_chunk<AMODE>(A, rb + r, cA, ok[u], pt[u], cb[u], a0, a1, a2, w[u], v);
              uint4 hi, lo;
              split8(v, hi, lo);
              *reinterpret_cast<uint4*>(sAhi + (size_t)gA * agb + r * 16) = hi;
              *reinterpret_cast<uint4*>(sAlo + (size_t)gA * agb + r * 16) = lo;
            }
          }
        }
      } else {
#pragma unroll 2
        for (int r = rA0; r < RT; r += rAstep) {
          const long long row = rb + r;
          float v[8];
          load_chunk<AMODE>(A, row, cA, vA && row < r_end, a0, a1, a2, v, true);
          uint4 hi, lo;
          split8(v, hi, lo);
          *reinterpret_cast<uint4*>(sAhi + (size_t)gA * agb + r * 16) = hi;
          *reinterpret_cast<uint4*>(sAlo + (size_t)gA * agb + r * 16) = lo;
        }
      }
    };
    if (aGroups == 8) stage_A(std::integral_constant<int, 8>{});
    else stage_A(std::integral_constant<int, 16>{});
    if (!genG) {
      // dY operand: RT / rGstep chunks per thread (4, 8 or 16; 2 when K2t = 16), same batching
#pragma unroll 1
      for (int rbase = rG0; rbase < RT; rbase += rGstep * UB) {
        long long pt[UB], cb[UB];
        bool ok[UB];
#pragma unroll
        for (int u = 0; u < UB; ++u) {
          const int r = rbase + rGstep * u;
          ok[u] = vG && r < RT && rb + r < r_end;
          pt[u] = 0; cb[u] = 0;
          if (GMODE == OP_DY_SPARSE) cloudmap_point(gcm, r, pt[u], cb[u]);
          if (GMODE == OP_DY_MAXK && r < RT) rowmap_point(grm, r, pt[u], cb[u]);
        }
        RawChunk w[UB];
#pragma unroll
        for (int u = 0; u < UB; ++u) fetch_chunk<GMODE>(G, rb + rbase + rGstep * u, cG, ok[u], pt[u], 0, w[u]);
#pragma unroll
        for (int u = 0; u < UB; ++u) {
          const int r = rbase + rGstep * u;
          if (r < RT) {
            float v[8];
            finish_chunk<GMODE>(G, rb + r, cG, ok[u], pt[u], cb[u], g0, g1, g2, w[u], v);
#pragma unroll
            for (int i = 0; i < 8; ++i) bs[i] += v[i];
            uint4 hi, lo;
            split8(v, hi, lo);
            *reinterpret_cast<uint4*>(sGhi + (size_t)gG * agb + r * 16) = hi;
            *reinterpret_cast<uint4*>(sGlo + (size_t)gG * agb + r * 16) = lo;
          }
        }
      }
    } else {
#pragma unroll 2
      for (int r = rG0; r < RT; r += rGstep) {
        const long long row = rb + r;
        float v[8];
        load_chunk<GMODE>(G, row, cG, vG && row < r_end, g0, g1, g2, v, true);
#pragma unroll
        for (int i = 0; i < 8; ++i) bs[i] += v[i];
        uint4 hi, lo;
        split8(v, hi, lo);
        *reinterpret_cast<uint4*>(sGhi + (size_t)gG * agb + r * 16) = hi;
        *reinterpret_cast<uint4*>(sGlo + (size_t)gG * agb + r * 16) = lo;
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t ah = smem_u32(sAhi), al = smem_u32(sAlo), gh = smem_u32(sGhi), gl = smem_u32(sGlo);
#pragma unroll 1
      for (int pass = 0; pass < 3; ++pass) {
        const uint32_t ab = (pass == 1) ? al : ah;
        const uint32_t gb = (pass == 2) ? gl : gh;
        for (int j = 0; j < RT / 16; ++j) {
          // rows 16j..16j+15 = one K=16 slice: k-groups of 8 rows are 128 B apart, channel groups agb apart
          const uint64_t ad = umma_desc(ab + (uint32_t)j * 256, 128, agb);
          const uint64_t gd = umma_desc(gb + (uint32_t)j * 256, 128, agb);
          tc_mma_bf16(tmem_base, ad, gd, idesc, accum);
          accum = 1;
        }
      }
      if (FUSE) {   // dA tile = dY * Wf^T : the dY image read K-major, accumulator at TMEM column K2t
        const uint32_t wh = smem_u32(sWhi), wl = smem_u32(sWlo);
        uint32_t acc2 = 0;
#pragma unroll 1
        for (int pass = 0; pass < 3; ++pass) {
          const uint32_t gb = (pass == 1) ? gl : gh;
          const uint32_t wb = (pass == 2) ? wl : wh;
          for (int kk = 0; kk < gGroups / 2; ++kk) {
            const uint64_t ad = umma_desc(gb + (uint32_t)(2 * kk) * agb, agb, 128);
            const uint64_t bd = umma_desc(wb + (uint32_t)(2 * kk) * w_group_bytes, w_group_bytes, 128);
            tc_mma_bf16(tmem_base + (uint32_t)K2t, ad, bd, idesc_f, acc2);
            acc2 = 1;
          }
        }
      }
      tc_commit(mma_bar);
    }
    pending = true;
    if (FUSE) {
      mbar_wait(mma_bar, phase);
      phase ^= 1;
      pending = false;
      tc_fence_after();
      // ---- fused epilogue (EPI_RELUMASK_STATS of the row GEMM): TMEM -> staging -> masked rows of dA + BN-backward sums
      const Epilogue& E = F.E;
      const int lq = warp & 3, ch = warp >> 2;
      const int trow = lq * 32 + lane;
      const int npass = (F.Nf + 63) / 64;
      for (int p = 0; p < npass; ++p) {
        float v[32];
        tc_ld32(tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(K2t + p * 64 + ch * 32), v);
        float* srow = stage + trow * STAGE_LD + ch * 32;
#pragma unroll
        for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(srow + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        __syncthreads();
        const int cl = p * 64 + e_c4 * 4;
        if (cl < F.Nf) {
          float scp[4], shp[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) { scp[j] = E.scp[cl + j]; shp[j] = E.shp[cl + j]; }
          float f0[4] = {0.f, 0.f, 0.f, 0.f}, f1[4] = {0.f, 0.f, 0.f, 0.f};
          float4 yq[8];     // the previous layer's pre-activations of the thread's 8 rows: all loads in flight at once
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const long long row = rb + e_r0 + 16 * it;
            yq[it] = (e_r0 + 16 * it < RT && row < r_end) ? *reinterpret_cast<const float4*>(E.yprev + row * E.ldyp + cl)
                                                          : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int r = e_r0 + 16 * it;
            const long long row = rb + r;
            if (r >= RT || row >= r_end) break;
            const float4 s4 = *reinterpret_cast<const float4*>(stage + r * STAGE_LD + e_c4 * 4);
            float o[4] = {s4.x, s4.y, s4.z, s4.w};
            const float4 y4 = yq[it];
            const float yv[4] = {y4.x, y4.y, y4.z, y4.w};
            float dm[4] = {1.f, 1.f, 1.f, 1.f};
            if (E.dmask) {
              const float4 m4 = *reinterpret_cast<const float4*>(E.dmask + row * F.Nf + cl);
              dm[0] = m4.x * E.dscale; dm[1] = m4.y * E.dscale; dm[2] = m4.z * E.dscale; dm[3] = m4.w * E.dscale;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const bool on = fmaf(yv[j], scp[j], shp[j]) > 0.f;
              o[j] = on ? o[j] * dm[j] : 0.f;
              f0[j] += o[j];
              f1[j] = fmaf(o[j], yv[j], f1[j]);
            }
            *reinterpret_cast<float4*>(E.out + row * E.ldo + cl) = make_float4(o[0], o[1], o[2], o[3]);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) { fst0[j] += (double)f0[j]; fst1[j] += (double)f1[j]; }
        }
        __syncthreads();   // staging (= the A images) free for the next pass / tile
      }
      tc_fence_before();
    }
  }
  if (FUSE) {   // flush the BN-backward sums of the data gradient (Nf <= 64 per pass; lanes l and l+16 own the same columns)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double a = fst0[j], b = fst1[j];
      a += __shfl_xor_sync(0xffffffffu, a, 16);
      b += __shfl_xor_sync(0xffffffffu, b, 16);
      const int cl = e_c4 * 4 + j;
      if (lane < 16 && cl < F.Nf) {
        atomicAdd(F.E.stats + cl, a);
        atomicAdd(F.E.stats + F.Nf + cl, b);
      }
    }
  }
  if (pending) { mbar_wait(mma_bar, phase); phase ^= 1; }
  tc_fence_after();

  // epilogue: TMEM (lane = k1 within tile, column = k2) -> partial
  {
    const int lq = warp & 3, ch = warp >> 2;
    const int k1 = k1_0 + lq * 32 + lane;
    float* prow = partial + ((size_t)blockIdx.x * K1p + k1) * K2p + k2_0;
    for (int c0 = ch * 32; c0 < K2t; c0 += 64) {
      float v[32];
      if (r_begin < r_end) tc_ld32(tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)c0, v);
      else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0.f;
      }
      if (k1 < K1p) {
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          if (k2_0 + c0 + i < K2p) *reinterpret_cast<float4*>(prow + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      }
    }
  }
  if (blockIdx.y == 0 && partial_b) {
#pragma unroll
    for (int i = 0; i < 8; ++i) atomicAdd(&bred[gG * 8 + i], bs[i]);
    __syncthreads();
    for (int i = tid; i < K2t; i += TC_THREADS)
      if (k2_0 + i < K2p) partial_b[(size_t)blockIdx.x * K2p + k2_0 + i] = bred[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tc_dealloc(tmem_base, (uint32_t)tmem_cols);
}

bool wgrad_tc_supported(const Operand& A, int amode, const Operand& G, int gmode) {
  if (amode != OP_PLAIN && amode != OP_BNRELU && amode != OP_EDGE) return false;
  if (gmode != OP_DY && gmode != OP_DY_SPARSE && gmode != OP_DY_MAXK) return false;
  if (A.C < 12 || G.C < 12) return false;
  if (amode == OP_EDGE && (A.k < 1 || A.k > 32768 || A.npts < 1)) return false;          // RowMap range
  if (gmode == OP_DY_MAXK && (G.k < 1 || G.k > 32768 || G.npts < 1)) return false;
  return true;
}

template <int AMODE, int GMODE, bool FUSE>
int launch_wgrad_tc(const Operand& A, const Operand& G, long long M, int S, int K1p, int K2p, float* partial,
                    float* partial_b, const FuseArgs& F, cudaStream_t st) {
  const int genA = tc_operand_fast(A, AMODE, A.C) ? 0 : 1, genG = tc_operand_fast(G, GMODE, G.C) ? 0 : 1;
  const int t1 = (A.C + TILE_M - 1) / TILE_M, t2 = (G.C + 255) / 256;
  int K2t = 16;
  const int k2max = G.C < 256 ? G.C : 256;
  while (K2t < k2max) K2t <<= 1;
  int tmem_cols = 64;
  while (tmem_cols < K2t + (FUSE ? F.Nf : 0)) tmem_cols <<= 1;
  const size_t wimg = FUSE ? (size_t)2 * (K2t / 8) * (F.Nf * 16 + 16) : 0;
  // <= 64 operand channels: 8-group A images, provided the 16 groups an M=128 MMA reads from each image base stay inside
  // the A + dY region (2*8 + 2*(K2t/8) >= 24 groups) and the fused epilogue's staging tile fits in it
  const int aGroups = (A.C <= 64 && K2t >= 32) ? 8 : 16;
  auto smem_for = [&](int rt) { return (size_t)(2 * aGroups + 2 * (K2t / 8)) * (rt * 16 + 16) + wimg + 32 + (size_t)K2t * 4 + 64; };
  // 64-row tiles when the 128-row images would leave one CTA (8 warps) per SM: the kernel is latency-bound on its loads
  static const bool rt64 = []() { const char* e = getenv("WSPC_WGRAD_RT"); return !(e && strcmp(e, "128") == 0); }();
  const int RT = (rt64 && !FUSE && smem_for(128) > 110 * 1024 && smem_for(64) <= 110 * 1024) ? 64 : 128;
  const size_t smem = smem_for(RT);
  long long rps = (M + S - 1) / S;
  rps = (rps + TILE_M - 1) / TILE_M * TILE_M;
  dim3 grid(S, t1, t2);
  if (smem <= 110 * 1024) {
    auto kern = colgemm_tc_kernel<AMODE, GMODE, 2, FUSE>;
    WSPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, TC_THREADS, smem, st>>>(A, G, M, rps, K1p, K2p, K2t, partial, partial_b, tmem_cols, genA, genG, aGroups, RT, F);
  } else {
    auto kern = colgemm_tc_kernel<AMODE, GMODE, 1, FUSE>;
    WSPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, TC_THREADS, smem, st>>>(A, G, M, rps, K1p, K2p, K2t, partial, partial_b, tmem_cols, genA, genG, aGroups, RT, F);
  }
  count_launch();
  WSPC_LAUNCH_CHECK("colgemm_tc_kernel");
  return WSPC_OK;
}

}  // namespace

int wgrad_tc_slabs(int K1, int K2) {
  const int tiles = ((K1 + TILE_M - 1) / TILE_M) * ((K2 + 255) / 256);
  int S = (2 * kNumSM) / tiles;
  return S < 1 ? 1 : S;
}

// returns 1 if handled (partials written for S slabs), 0 if not eligible, <0 on error
int wgrad_tc_dispatch(const Operand& A, int amode, const Operand& G, int gmode, long long M, int S, int K1p, int K2p,
                      float* partial, float* partial_b, cudaStream_t st) {
  if (!wgrad_tc_supported(A, amode, G, gmode)) return 0;
  int rc = -100;
  const FuseArgs none{};
#define WSPC_WG(AM, GM) \
  if (amode == AM && gmode == GM) rc = launch_wgrad_tc<AM, GM, false>(A, G, M, S, K1p, K2p, partial, partial_b, none, st);
  WSPC_WG(OP_PLAIN, OP_DY) WSPC_WG(OP_BNRELU, OP_DY) WSPC_WG(OP_EDGE, OP_DY) WSPC_WG(OP_PLAIN, OP_DY_SPARSE)
  WSPC_WG(OP_BNRELU, OP_DY_MAXK)
#undef WSPC_WG
  if (rc == -100) return 0;
  return rc == WSPC_OK ? 1 : rc;
}

// fused weight + data gradient of one conv2d (see FuseArgs): eligible when dW is a single (<=128 x <=256) tile, so every
// row tile is visited exactly once, and dA has at most 64 columns.  returns 1 if handled, 0 if not eligible, <0 on error
int bwd_fused_tc_dispatch(const Operand& A, int amode, const Operand& G, int gmode, long long M, int S, int K1p, int K2p,
                          float* partial, float* partial_b, const float* Wf, long long ldw, int Nf, const Epilogue& E,
                          cudaStream_t st) {
  if (!wgrad_tc_supported(A, amode, G, gmode)) return 0;
  if (amode != OP_BNRELU || (gmode != OP_DY && gmode != OP_DY_MAXK)) return 0;
  if (A.C > TILE_M || G.C > 256 || G.C % 16 != 0 || Nf % 16 != 0 || Nf < 16 || Nf > 64) return 0;
  if (!aligned16(E.out) || (E.ldo % 4) != 0 || !aligned16(E.yprev) || (E.ldyp % 4) != 0 || !E.stats || !E.scp || !E.shp) return 0;
  if (E.dmask && !aligned16(E.dmask)) return 0;
  FuseArgs F;
  F.Wf = Wf; F.ldw = ldw; F.Nf = Nf; F.E = E;
  int rc = -100;
  if (gmode == OP_DY) rc = launch_wgrad_tc<OP_BNRELU, OP_DY, true>(A, G, M, S, K1p, K2p, partial, partial_b, F, st);
  else rc = launch_wgrad_tc<OP_BNRELU, OP_DY_MAXK, true>(A, G, M, S, K1p, K2p, partial, partial_b, F, st);
  return rc == WSPC_OK ? 1 : rc;
}

namespace {
}  // namespace

// returns 1 if the tensor-core path handled the call, 0 if the shape is not eligible, <0 on error
size_t rowgemm_tc_workspace_bytes(int N, int K) { return tc_wimg_bytes(N, K); }

int rowgemm_tc_dispatch(const Operand& A, int amode, const float* Bm, long long ldb, int bT, long long M, int N, int K,
                        const Epilogue& E, int emode, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (amode == OP_IMG) {       // pre-split image operand: the warp-specialised kernel or nothing
    const bool ok = ws_img_supported(M, N, K) && ws && ws_bytes >= tc_wimg_bytes(N, K) && aligned16(ws) && aligned16(A.p) &&
                    (emode == EPI_STORE || emode == EPI_STORE_STATS) && aligned16(E.out) && (E.ldo % 4) == 0;
    if (!ok) {
      set_error("conv1x1_rows: WSPC_OP_IMG needs an eligible shape (wspc_rows_image_supported), the weight-image workspace and a "
                "STORE / STORE_STATS epilogue");
      return WSPC_ERR_INVALID;
    }
    const int rc = emode == EPI_STORE ? launch_ws_img<EPI_STORE>(A, Bm, ldb, bT, M, N, K, E, ws, st)
                                      : launch_ws_img<EPI_STORE_STATS>(A, Bm, ldb, bT, M, N, K, E, ws, st);
    return rc == WSPC_OK ? 1 : rc;
  }
  if (!tc_supported(A, amode, Bm, M, N, K, E, emode)) return 0;
  int rc = -100;
#define WSPC_TC(AM, EM) \
  if (amode == AM && emode == EM) rc = launch_tc<AM, EM>(A, Bm, ldb, bT, M, N, K, E, ws, ws_bytes, st);
  WSPC_TC(OP_PLAIN, EPI_STORE) WSPC_TC(OP_PLAIN, EPI_STORE_STATS) WSPC_TC(OP_PLAIN, EPI_ACCUM)
  WSPC_TC(OP_BNRELU, EPI_STORE) WSPC_TC(OP_BNRELU, EPI_STORE_STATS)
  WSPC_TC(OP_EDGE, EPI_STORE) WSPC_TC(OP_EDGE, EPI_STORE_STATS)
  WSPC_TC(OP_DY, EPI_STORE) WSPC_TC(OP_DY, EPI_RELUMASK_STATS) WSPC_TC(OP_DY, EPI_EDGE_SCATTER) WSPC_TC(OP_DY, EPI_ACCUM)
  WSPC_TC(OP_DY_SPARSE, EPI_STORE) WSPC_TC(OP_DY_SPARSE, EPI_ACCUM)
  WSPC_TC(OP_DY_MAXK, EPI_RELUMASK_STATS)
#undef WSPC_TC
  if (rc == -100) return 0;
  return rc == WSPC_OK ? 1 : rc;
}

}  // namespace wspc

// ---------------------------------------------------------------------------------------------------------------------
// adj_conv7 + max_pool2d([N,1]) (DGCNN_S3DIS.py:80-85, DGCNN_ShapeNet.py:80-85) in one pass; see include/wspc.h
extern "C" int wspc_conv1x1_pool_supported(long long M, int N, int K, int npts) {
  return wspc::ws_pool_supported(M, N, K, npts) ? 1 : 0;
}

extern "C" int wspc_conv1x1_pool_fwd(const wspc_operand_t* A, int a_mode, const float* W, long long ldw, long long M, int N, int K,
                                     int npts, const float* bias, const float* gamma, double* stats, unsigned long long* keys,
                                     void* workspace, size_t workspace_bytes, wspc_stream_t stream) {
  using namespace wspc;
  if (int rc = check_arch()) return rc;
  WSPC_REQUIRE(A && W && gamma && stats && keys && workspace, "conv1x1_pool_fwd: null argument");
  WSPC_REQUIRE(a_mode == OP_PLAIN || a_mode == OP_BNRELU || a_mode == OP_IMG, "conv1x1_pool_fwd: operand mode %d (PLAIN, BNRELU or IMG)",
               a_mode);
  WSPC_REQUIRE(A->p && A->C == K, "conv1x1_pool_fwd: operand channels %d != K %d", A->C, K);
  WSPC_REQUIRE(ws_pool_supported(M, N, K, npts), "conv1x1_pool_fwd: shape M=%lld N=%d K=%d npts=%d is not eligible "
               "(wspc_conv1x1_pool_supported)", M, N, K, npts);
  WSPC_REQUIRE(a_mode == OP_IMG ? aligned16(A->p) : (tc_operand_fast(*A, a_mode, K) && !(a_mode == OP_BNRELU && A->dmask)),
               "conv1x1_pool_fwd: operand must be 16-byte aligned");
  if (workspace_bytes < tc_wimg_bytes(N, K) || !aligned16(workspace)) {
    set_error("conv1x1_pool_fwd: workspace too small (wspc_conv1x1_rows_workspace_bytes)");
    return WSPC_ERR_WORKSPACE;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  WSPC_CUDA(cudaMemsetAsync(keys, 0, (size_t)(M / npts) * N * sizeof(unsigned long long), st));
  Epilogue E{};
  E.bias = bias;
  E.stats = stats;
  E.scp = gamma;
  E.dx = reinterpret_cast<float*>(keys);
  E.npts = npts;
  if (a_mode == OP_IMG) return launch_ws_pool<OP_IMG>(*A, W, ldw, M, N, K, E, workspace, st);
  if (a_mode == OP_PLAIN) return launch_ws_pool<OP_PLAIN>(*A, W, ldw, M, N, K, E, workspace, st);
  return launch_ws_pool<OP_BNRELU>(*A, W, ldw, M, N, K, E, workspace, st);
}

extern "C" int wspc_maxn_from_keys(const unsigned long long* keys, const float* gamma, const float* sc, const float* sh, int B,
                                   int C, float* g, int32_t* amax, float* ymax, wspc_stream_t stream) {
  using namespace wspc;
  if (int rc = check_arch()) return rc;
  WSPC_REQUIRE(keys && gamma && sc && sh && g && amax && ymax && B >= 1 && C >= 1, "maxn_from_keys: bad argument");
  const long long total = (long long)B * C;
  pool_keys_finish_kernel<<<(unsigned)((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(keys, gamma, sc, sh,
                                                                                                          total, C, g, amax, ymax);
  count_launch();
  WSPC_LAUNCH_CHECK("pool_keys_finish_kernel");
  return WSPC_OK;
}

// pre-split operand image (see include/wspc.h)
extern "C" size_t wspc_rows_image_bytes(long long M, int K) {
  return (M >= 1 && K >= 32 && K % 32 == 0) ? wspc::ws_img_bytes(M, K) : 0;
}
extern "C" int wspc_rows_image_supported(long long M, int N, int K) { return wspc::ws_img_supported(M, N, K) ? 1 : 0; }
extern "C" int wspc_rows_image(const float* x, long long ldx, long long M, int K, void* image, wspc_stream_t stream) {
  using namespace wspc;
  if (int rc = check_arch()) return rc;
  WSPC_REQUIRE(x && image && M >= 1, "rows_image: null argument");
  WSPC_REQUIRE(K >= 32 && K % 32 == 0 && (ldx % 4) == 0 && ldx >= K && aligned16(x) && aligned16(image),
               "rows_image: K=%d must be a multiple of 32 and the rows 16-byte aligned", K);
  const long long tiles = (M + TILE_M - 1) / TILE_M;
  rows_image_kernel<<<(unsigned)tiles, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, ldx, M, K, static_cast<unsigned char*>(image));
  count_launch();
  WSPC_LAUNCH_CHECK("rows_image_kernel");
  return WSPC_OK;
}
