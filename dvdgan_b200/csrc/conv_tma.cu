// TMA + tcgen05 implicit-GEMM convolution (forward / dgrad and weight gradient) for sm_100a.
//
// Operand preparation (HBM-bound, once per conv call): the fp32 NC(D)HW activation (or dY) is split into two bf16
// planes in channels-last layout [N][D][H][W][Cp] (ReLU / nearest-x2 upsample of the reference's F.relu /
// F.interpolate fused in); the fp32 packed weights [tap][Cin][Cout] become [tap][CoutP][CinP] planes.
// Two plane formats (A and B of one tcgen05.mma must share one):
//   fp16:  hi = fp16(x), lo = fp16((x - hi) * 2^11)   22 significant bits, |x| <= 65504.  Forward activations and
//          weights (x_kind = 1).  Needed for the 1e-3 contract: with 16-bit-precision operands the Generator's output is
//          1.2e-3 off the fp32 reference after 48 recurrent frames, with these 2.3e-4 (profiles/r2).  Elements beyond
//          fp16's range are clamped AND counted (dvd_saturation_count); option "fwd_bf16" switches every launch to
//   bf16:  hi = bf16(x), lo = bf16(x - hi)             16-17 significant bits, fp32's exponent range.  Everything that
//          holds gradients (dgrad / wgrad operands), and the forward too under "fwd_bf16" or "oneacc".
// The GEMM issues three MMAs per algorithmic MAC,  A_hi*B_hi,  A_lo*B_hi,  A_hi*B_lo,  into fp32 TMEM accumulators:
//   two-accumulator kernels:   D_main += A_hi*B_hi ;  D_lo += A_lo*B_hi + A_hi*B_lo ;  result = D_main + D_lo / s
//                              (the many small cross terms stay out of the large accumulator, whose adds truncate)
//   ONEACC kernels (bf16 planes only: the low plane is unscaled):   D += all three.  Halves the TMEM footprint, which
//                              lets the persistent 256-wide CTA-pair kernel keep TWO accumulator sets: the epilogue of
//                              tile i drains set i&1 while the MMAs of tile i+1 fill the other one.  Measured +7..16 %
//                              on isolated GEMMs, +2 % on the step, but tied to the bf16 planes' precision: off by default.
//
// forward CTA (64 + 32*EW threads): warp 0 = TMA producer (one lane), warp 1 = TMEM alloc + tcgen05.mma issuer (one
// lane), EW = 4 or 8 epilogue warps.  A tile = 128 output pixels x 64 channels, fetched per tap as ONE 5-D TMA box
// (c, w, h, d, n) whose start coordinate carries the tap's shift; out-of-bounds = zero fill = the conv padding.
// B tile = BN couts x 64 channels, a 2-D box of the prepared weights.  K-major, 128B-swizzled, 2-5 stage ring.
// Variants (template parameters, chosen per shape in tma_fwd_launch_ex): CG = 2 CTA pairs (cta_group::2, 256-row tiles,
// each CTA stages half of B), OCC = 2 co-resident CTAs for short reductions, PERSIST = strided tile walk per SM pair,
// plain / ConvGRU epilogues; the measured bound that drives these choices is the ~64 B/clk an SM can ingest
// (profiles/r1/README.md).
//
// wgrad CTA: D[ci][co] += sum_pixels X[pix + tap][ci] * dY[pix][co]: both operands are MN-major views of the same
// channels-last planes (k = 64 consecutive pixels = one TMA box), one tap and one pixel range per CTA (pair: 256 ci),
// taps fastest in the grid so that a wave of CTAs re-uses one pixel range from L2.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "conv_params.cuh"

namespace dvd {
namespace tma {

constexpr int BM = 128;
constexpr int BKC = 64;
constexpr int NT = 192;

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctaid_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctaid.x;" : "=r"(r));
  return r;
}

// ---- CTA pair (cta_group::2): the two CTAs of a (2,1,1) cluster run one M = 256 MMA; rank 0 issues it
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {      // same offset in CTA `rank`
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// loads into THIS CTA's shared memory, complete_tx on a barrier that may live in the peer CTA
__device__ __forceinline__ void tma_load_5d_pair(uint32_t dst, const CUtensorMap* tm, uint32_t bar_cluster, int c0,
                                                 int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* tm, uint32_t bar_cluster, int c0,
                                                 int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}

// shared-memory matrix descriptor, SWIZZLE_128B.  K-major: rows at 128 B, SBO = 1024 (8-row group).
// MN-major: 64 MN elements per 128-B row, 8 k-rows per 1024-B atom (SBO), next 64-wide MN block LBO bytes away.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;      // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;      // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: D = f32, A/B = bf16 (fmt 1) or fp16 (fmt 0), M = m, N = n; mn_major: both operands
// MN-major
__device__ __forceinline__ uint32_t make_idesc(int n, int mn_major, int m, uint32_t fmt = 1) {
  uint32_t d = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
  if (mn_major) d |= (1u << 15) | (1u << 16);
  return d;
}
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// pair commit: arrives on the barrier at this offset in both CTAs of the pair
__device__ __forceinline__ void mma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 8 consecutive fp32 -> 8 bf16 hi + 8 bf16 lo (x = hi + lo)
__device__ __forceinline__ void bf16_split8(const float* v, uint4* hi, uint4* lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 hp = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    const float2 hf = __bfloat1622float2(hp);
    const __nv_bfloat162 lp = __floats2bfloat162_rn(v[2 * i] - hf.x, v[2 * i + 1] - hf.y);
    h[i] = *reinterpret_cast<const uint32_t*>(&hp);
    l[i] = *reinterpret_cast<const uint32_t*>(&lp);
  }
  *hi = make_uint4(h[0], h[1], h[2], h[3]);
  *lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// fp16 planes: hi = fp16(x), lo = fp16((x - hi) * 2^11); values beyond fp16's range are clamped and counted
constexpr float kLoScaleFp16 = 2048.f;
__device__ unsigned int g_fp16_saturated = 0;
__device__ __forceinline__ void fp16_split8(const float* v, uint4* hi, uint4* lo) {
  const float lim = 65504.f;
  uint32_t h[4], l[4];
  bool sat = false;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float a = v[2 * i], b = v[2 * i + 1];
    sat = sat || fabsf(a) > lim || fabsf(b) > lim;
    const __half2 hp = __floats2half2_rn(fminf(fmaxf(a, -lim), lim), fminf(fmaxf(b, -lim), lim));
    const float2 hf = __half22float2(hp);
    const float ra = (a - hf.x) * kLoScaleFp16, rb = (b - hf.y) * kLoScaleFp16;
    const __half2 lp = __floats2half2_rn(fminf(fmaxf(ra, -lim), lim), fminf(fmaxf(rb, -lim), lim));
    h[i] = *reinterpret_cast<const uint32_t*>(&hp);
    l[i] = *reinterpret_cast<const uint32_t*>(&lp);
  }
  if (sat) atomicAdd(&g_fp16_saturated, 1u);          // rare: one count per 8-element group
  *hi = make_uint4(h[0], h[1], h[2], h[3]);
  *lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// ------------------------------------------------------------------------------------------------ operand prep
// src: fp32, element (n, c, pix) at n1*s1 + n2*s2 + c*cs + srcpix(pix)   ->  dst planes [n][pix][Cp] bf16 hi / lo.
// Block = 32 output pixels x 64 channels, transposed through shared memory (coalesced on both sides).
struct PrepP {
  const float* src;
  __nv_bfloat16* hi;
  __nv_bfloat16* lo;
  int N2, C, Cp;
  int64_t s1, s2, cs;
  int in_pix;        // valid output pixels per image (rows >= in_pix are zero-filled)
  int out_pix;       // rows per image in dst
  int W, HW, up;     // output W, H*W (for the upsample source map); up = 1: source is (H/2, W/2)
  int relu;
  int fp16;          // plane format (see the file header)
};

__global__ void __launch_bounds__(256) prep_planes_kernel(const PrepP p) {
  __shared__ float tile[64][33];
  const int n = blockIdx.z;
  const int pix0 = blockIdx.x * 32;
  const int c0 = blockIdx.y * 64;
  const int n1 = n / p.N2, n2 = n - n1 * p.N2;
  const float* src = p.src + (int64_t)n1 * p.s1 + (int64_t)n2 * p.s2;
  const int tid = threadIdx.x;
  {
    const int px = tid & 31;
    const int pix = pix0 + px;
    int64_t so = -1;
    if (pix < p.in_pix) {
      if (p.up) {
        const int z = pix / p.HW;
        const int r = pix - z * p.HW;
        const int y = r / p.W, x = r - y * p.W;
        so = ((int64_t)z * (p.HW >> 2)) + (int64_t)(y >> 1) * (p.W >> 1) + (x >> 1);
      } else {
        so = pix;
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = (tid >> 5) + 8 * j;
      float v = 0.f;
      if (so >= 0 && c0 + c < p.C) {
        v = __ldg(src + (int64_t)(c0 + c) * p.cs + so);
        if (p.relu) v = fmaxf(v, 0.f);
      }
      tile[c][px] = v;
    }
  }
  __syncthreads();
  {
    const int px = tid >> 3, q = tid & 7;
    const int pix = pix0 + px;
    if (pix < p.out_pix) {
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = tile[q * 8 + i][px];
      uint4 h, l;
      if (p.fp16) fp16_split8(v, &h, &l);
      else bf16_split8(v, &h, &l);
      const int64_t o = ((int64_t)n * p.out_pix + pix) * p.Cp + c0 + q * 8;
      *reinterpret_cast<uint4*>(p.hi + o) = h;
      *reinterpret_cast<uint4*>(p.lo + o) = l;
    }
  }
}

// Vectorised variant for the common case (no upsample, pixel counts and strides that are multiples of 4, 16-byte aligned
// source, no padded rows): block = 64 pixels x 64 channels, 128-bit loads along the pixel axis (256 contiguous bytes
// per channel row and half-warp), 128-bit plane stores.
__global__ void __launch_bounds__(256) prep_planes_vec_kernel(const PrepP p) {
  __shared__ float tile[64][65];
  const int n = blockIdx.z;
  const int pix0 = blockIdx.x * 64;
  const int c0 = blockIdx.y * 64;
  const int n1 = n / p.N2, n2 = n - n1 * p.N2;
  const float* src = p.src + (int64_t)n1 * p.s1 + (int64_t)n2 * p.s2;
  const int tid = threadIdx.x;
  {
    const int g = tid & 15, pix = pix0 + 4 * g;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = (tid >> 4) + 16 * j;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (pix < p.in_pix && c0 + c < p.C) {
        v = __ldg(reinterpret_cast<const float4*>(src + (int64_t)(c0 + c) * p.cs + pix));
        if (p.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
      }
      tile[c][4 * g] = v.x; tile[c][4 * g + 1] = v.y; tile[c][4 * g + 2] = v.z; tile[c][4 * g + 3] = v.w;
    }
  }
  __syncthreads();
#pragma unroll
  for (int rep = 0; rep < 2; ++rep) {
    const int item = tid + 256 * rep;          // 64 pixels x 8 chunks of 8 channels
    const int px = item >> 3, q = item & 7;
    const int pix = pix0 + px;
    if (pix < p.out_pix) {
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = tile[q * 8 + i][px];
      uint4 h, l;
      if (p.fp16) fp16_split8(v, &h, &l);
      else bf16_split8(v, &h, &l);
      const int64_t o = ((int64_t)n * p.out_pix + pix) * p.Cp + c0 + q * 8;
      *reinterpret_cast<uint4*>(p.hi + o) = h;
      *reinterpret_cast<uint4*>(p.lo + o) = l;
    }
  }
}

// ------------------------------------------------------------------------------------------------ shared pieces
struct TileGeom {
  int bw, bh, bd, bn;      // box extents (w, h, d, images); bw*bh*bd*bn = rows per box
};

constexpr int LATE_ITERS = 24; // the epilogue's L2 prefetch starts this many k-blocks before the end of the main loop
                               // (earlier, the lines are evicted again by the operand stream of a long main loop)

constexpr int tmem_cols_for(int need) { return need <= 32 ? 32 : need <= 64 ? 64 : need <= 128 ? 128 : need <= 256 ? 256 : 512; }

// barrier block: full[S] | empty[S] | accum[2] | late[2] | tfree[2] | tmem slot      (index 1 of the pairs is used by
// the double-buffered persistent kernel only)
template <int STAGES>
struct Bars {
  uint64_t* base;
  __device__ uint64_t* full(int s) const { return base + s; }
  __device__ uint64_t* empty(int s) const { return base + STAGES + s; }
  __device__ uint64_t* accum(int b) const { return base + 2 * STAGES + b; }           // "accumulator set b is complete"
  __device__ uint64_t* late(int b) const { return base + 2 * STAGES + 2 + b; }        // "the main loop is about to finish"
  __device__ uint64_t* tfree(int b) const { return base + 2 * STAGES + 4 + b; }       // "set b has been read out"
  __device__ uint32_t* slot() const { return reinterpret_cast<uint32_t*>(base + 2 * STAGES + 6); }
  __device__ void init(uint32_t tfree_count) const {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(full(s)), 1);
      mbar_init(smem_u32(empty(s)), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(accum(b)), 1);
      mbar_init(smem_u32(late(b)), 1);
      mbar_init(smem_u32(tfree(b)), tfree_count);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
};

// The single-thread MMA issue loop shared by the forward and wgrad kernels: n_iters k-blocks into the accumulator(s)
// at TMEM columns main_col / lo_col (ONEACC: lo_col == main_col).  lbo/sbo: descriptor strides; kadv: start-address
// advance (16-byte units) per UMMA_K = 16 step.  `buf` selects the accum / late barriers that are signalled.
template <int BN, bool ONEACC, int STAGES, int STAGE_BYTES, int A_BYTES, int B_BYTES, int CG>
__device__ __forceinline__ void mma_issue_loop(uint8_t* smem, const Bars<STAGES>& bars, uint32_t main_col,
                                               uint32_t lo_col, int buf, int n_iters, int mn_major, uint32_t lbo,
                                               uint32_t sbo, uint32_t kadv, uint32_t fmt /* 0: fp16, 1: bf16 planes */,
                                               int& stage, uint32_t& phase /* ring position, carried across tiles */) {
  const uint32_t idesc = make_idesc(BN, mn_major, BM * CG, fmt);
  const int late_it = n_iters > LATE_ITERS ? n_iters - LATE_ITERS : 0;
  for (int it = 0; it < n_iters; ++it) {
    if (it == late_it) {          // tell the epilogue warps (of both CTAs of a pair) to start prefetching their operands
      mbar_arrive(smem_u32(bars.late(buf)));
      if constexpr (CG == 2) mbar_arrive_cluster(mapa_rank(smem_u32(bars.late(buf)), 1));
    }
    mbar_wait(smem_u32(bars.full(stage)), phase);
    tc_fence_after();
    const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
    const uint64_t a_hi = make_desc(sa, lbo, sbo), a_lo = make_desc(sa + A_BYTES, lbo, sbo);
    const uint64_t b_hi = make_desc(sa + 2 * A_BYTES, lbo, sbo), b_lo = make_desc(sa + 2 * A_BYTES + B_BYTES, lbo, sbo);
#pragma unroll
    for (int kk = 0; kk < BKC / 16; ++kk) {
      const uint64_t adv = (uint64_t)(kk * kadv);
      const uint32_t not_first = (it | kk) != 0;
      if constexpr (CG == 2) {
        mma_f16_pair(main_col, a_hi + adv, b_hi + adv, idesc, not_first);
        mma_f16_pair(lo_col, a_lo + adv, b_hi + adv, idesc, ONEACC ? 1u : not_first);
        mma_f16_pair(lo_col, a_hi + adv, b_lo + adv, idesc, 1);
      } else {
        mma_f16(main_col, a_hi + adv, b_hi + adv, idesc, not_first);
        mma_f16(lo_col, a_lo + adv, b_hi + adv, idesc, ONEACC ? 1u : not_first);
        mma_f16(lo_col, a_hi + adv, b_lo + adv, idesc, 1);
      }
    }
    if constexpr (CG == 2) mma_commit_pair(smem_u32(bars.empty(stage)));
    else mma_commit(smem_u32(bars.empty(stage)));
    if (++stage == STAGES) { stage = 0; phase ^= 1; }
  }
  if constexpr (CG == 2) mma_commit_pair(smem_u32(bars.accum(buf)));
  else mma_commit(smem_u32(bars.accum(buf)));
}

// CG = 1: one CTA per 128 x BN tile.  CG = 2: a CTA pair (cluster (2,1,1)) per 256 x BN tile: each CTA stages its own
// 128 rows of A and HALF of the BN weight rows, rank 0 issues tcgen05.mma.cta_group::2 (M = 256) and each CTA's TMEM
// receives its 128 accumulator rows.  The pair exists because the main loop is bound by the bytes an SM can pull in
// per cycle (64 KB of planes per 128x128x64 k-block = 3 MMAs per 4 bytes): halving the B bytes per SM is the lever.
// OCC = 2: two CTAs per SM (two-stage rings, <= 168 registers, 2 * BN <= 256 TMEM columns each) for short-K
// convolutions, where one CTA's prologue / epilogue would otherwise leave the tensor pipe idle: the co-resident CTA's
// main loop runs underneath it.
template <int BN, int CG, int OCC, bool PERSIST, bool ONEACC>
struct Cfg {
  static constexpr int A_BYTES = BM * 128;
  static constexpr int B_BYTES = (BN / CG) * 128;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int STAGES_FIT = (OCC == 2 ? 108 * 1024 : 200 * 1024) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_FIT > 5 ? 5 : STAGES_FIT;
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 + 256;
  static constexpr int NBUF = (PERSIST && ONEACC) ? 2 : 1;          // accumulator sets
  static constexpr int SET_COLS = ONEACC ? BN : 2 * BN;             // TMEM columns of one set
  static constexpr int TMEM_COLS = tmem_cols_for(NBUF * SET_COLS);
  static_assert(NBUF * SET_COLS <= 512, "TMEM has 512 columns");
  static_assert(OCC == 1 || TMEM_COLS <= 256, "two CTAs per SM share 512 TMEM columns");
  static_assert(STAGES >= 2, "pipeline needs two stages");
};

struct FwdP {
  ConvP c;
  TileGeom g;
  int CoutP;
  int fp16;             // operand planes are fp16 (forward values) rather than bf16
  float lo_inv;         // 1 / scale of the low-order planes
  GruEpi gru;           // ConvGRU gate / state epilogue (mode 0: plain conv epilogue)
  int prefetch;         // epilogue operands are prefetched into L2 late in the main loop
  int nt, tiles;        // persistent kernels: n-tiles and total (m-unit, n-tile) tiles
  int a_c_off;          // first channel of the A operand inside (shared) planes
};

// 32 consecutive channels of one pixel -> hi / lo planes (same split as prep_planes_kernel)
__device__ __forceinline__ void store_planes32(__nv_bfloat16* hi, __nv_bfloat16* lo, const float* v, int fp16) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 h, l;
    if (fp16) fp16_split8(v + q * 8, &h, &l);
    else bf16_split8(v + q * 8, &h, &l);
    reinterpret_cast<uint4*>(hi)[q] = h;
    reinterpret_cast<uint4*>(lo)[q] = l;
  }
}

// EW = epilogue warps: 4 (one per TMEM lane quarter) or 8 (two per quarter, alternating 32-column chunks).
// PERSIST: one CTA (pair) per SM (pair) walks tiles t = cluster, cluster + #clusters, ... ; the TMA ring keeps running
// into the next tile while the epilogue drains the accumulators, and barrier setup / TMEM allocation / tensor-map fetch
// are paid once per SM instead of once per tile.  With two accumulator sets (ONEACC) the MMAs of tile i+1 start as soon
// as the set of tile i-1 has been read out; with one set they wait for the epilogue of tile i.
template <int BN, int CG, int OCC, int EW, bool PERSIST, bool ONEACC>
__global__ void __launch_bounds__(64 + 32 * EW, OCC)
conv_tma_fwd_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                    const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                    const FwdP fp) {
  using C = Cfg<BN, CG, OCC, PERSIST, ONEACC>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const Bars<C::STAGES> bars{reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES)};
  uint64_t* full_bar = bars.full(0);
  uint64_t* empty_bar = bars.empty(0);
  uint32_t* tmem_slot = bars.slot();

  const ConvP& p = fp.c;
  const dvd_conv_desc& d = p.d;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t cx = CG == 2 ? cluster_ctaid_x() : 0u;
  const bool leader = CG == 1 || cx == 0;       // pair: rank 0 owns the full barriers and issues the MMAs

  static_assert(EW == 4 || EW == 8, "one or two epilogue warps per TMEM lane quarter");
  if (tid == 0) bars.init((uint32_t)(32 * EW * CG));
  if (warp == 1) {
    if constexpr (CG == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"((uint32_t)C::TMEM_COLS));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"((uint32_t)C::TMEM_COLS));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();        // the peer's barriers are initialised before anything is signalled to them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int it_begin = blockIdx.z * p.iters_per_split;
  const int it_end = min(it_begin + p.iters_per_split, p.iters_total);
  const int n_iters = it_end - it_begin;
  // tiles of this CTA: persistent = a strided walk over (m-unit, n-tile) with n fastest; otherwise the one tile of
  // blockIdx.  An m-unit is CG consecutive 128-row tiles (one per CTA of the pair).
  const int tile_first = PERSIST ? (int)(blockIdx.x / CG) : 0;
  const int tile_stride = PERSIST ? (int)(gridDim.x / CG) : 1;
  const int tile_end = PERSIST ? fp.tiles : 1;
  auto tile_origin = [&](int tile, int* m0, int* n0) {
    if (PERSIST) {
      const int nt = tile % fp.nt, mu = tile / fp.nt;
      *m0 = (mu * CG + (int)cx) * BM;
      *n0 = nt * BN;
    } else {
      *m0 = blockIdx.x * BM;
      *n0 = blockIdx.y * BN;
    }
  };

  if (warp == 0) {
    if (lane == 0) {
     int stage = 0;
     uint32_t phase = 0;
     for (int tile = tile_first; tile < tile_end; tile += tile_stride) {
      int m0, n0;
      tile_origin(tile, &m0, &n0);
      // box origin of this CTA's 128 rows of the A tile -> (image, z, y, x); B slice: this CTA's half of the n-tile
      constexpr int b_rows = BN / CG;
      const int img = m0 / p.DHW;
      int rem = m0 - img * p.DHW;
      const int z0 = rem / p.HW;
      rem -= z0 * p.HW;
      const int y0 = rem / d.W;
      const int x0 = rem - y0 * d.W;
      int tap = it_begin / p.ck;
      int cchunk = it_begin - tap * p.ck;
      for (int it = 0; it < n_iters; ++it) {
        const int kw = tap % d.kW;
        const int t2 = tap / d.kW;
        const int kh = t2 % d.kH;
        const int kd = t2 / d.kH;
        mbar_wait(smem_u32(empty_bar + stage), phase ^ 1);
        const uint32_t fb = smem_u32(full_bar + stage);
        const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
        const int c0 = cchunk * BKC;                 // channel block of the weights; the activations may be offset
        const int ca = c0 + fp.a_c_off;
        const int px = x0 + kw - d.kW / 2, py = y0 + kh - d.kH / 2, pz = z0 + kd - d.kD / 2;
        const int brow = tap * fp.CoutP + n0 + (int)cx * b_rows;
        if constexpr (CG == 2) {
          // both CTAs' bytes are counted on rank 0's barrier
          if (leader) mbar_expect_tx(fb, 2 * C::STAGE_BYTES);
          const uint32_t fbl = mapa_rank(fb, 0);
          tma_load_5d_pair(sa, &tmA_hi, fbl, ca, px, py, pz, img);
          tma_load_5d_pair(sa + C::A_BYTES, &tmA_lo, fbl, ca, px, py, pz, img);
          tma_load_2d_pair(sa + 2 * C::A_BYTES, &tmB_hi, fbl, c0, brow);
          tma_load_2d_pair(sa + 2 * C::A_BYTES + C::B_BYTES, &tmB_lo, fbl, c0, brow);
        } else {
          mbar_expect_tx(fb, C::STAGE_BYTES);
          tma_load_5d(sa, &tmA_hi, fb, ca, px, py, pz, img);
          tma_load_5d(sa + C::A_BYTES, &tmA_lo, fb, ca, px, py, pz, img);
          tma_load_2d(sa + 2 * C::A_BYTES, &tmB_hi, fb, c0, brow);
          tma_load_2d(sa + 2 * C::A_BYTES + C::B_BYTES, &tmB_lo, fb, c0, brow);
        }
        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        if (++cchunk == p.ck) { cchunk = 0; ++tap; }
      }
     }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      int stage = 0, ti = 0;
      uint32_t phase = 0;
      for (int tile = tile_first; tile < tile_end; tile += tile_stride, ++ti) {
        const int buf = C::NBUF == 2 ? (ti & 1) : 0;
        const int use = C::NBUF == 2 ? (ti >> 1) : ti;        // how many times this accumulator set has been filled
        if (PERSIST && use > 0) {       // the epilogue warps (of both CTAs of a pair) have read the set's last tile out
          mbar_wait(smem_u32(bars.tfree(buf)), (uint32_t)((use - 1) & 1));
          tc_fence_after();
        }
        const uint32_t main_col = tmem_base + (uint32_t)(buf * C::SET_COLS);
        // K-major operands: LBO unused (16), SBO = 1024 (8 rows of 128 B), 32 bytes (2 units) per UMMA_K step
        mma_issue_loop<BN, ONEACC, C::STAGES, C::STAGE_BYTES, C::A_BYTES, C::B_BYTES, CG>(
            smem, bars, main_col, ONEACC ? main_col : main_col + BN, buf, n_iters, 0, 16, 1024, 2, fp.fp16 ? 0u : 1u,
            stage, phase);
      }
    }
    __syncwarp();
  } else {
   // ---------------- epilogue: warps 2..(EW+1); a warp may only touch TMEM lanes [32*(warp%4), +32); with EW = 8 the
   // two warps of a quarter take alternate 32-column chunks
   const int q = warp & 3;
   const int half = (warp - 2) >> 2;
   int ti = 0;
   for (int tile = tile_first; tile < tile_end; tile += tile_stride, ++ti) {
    int m0, n0;
    tile_origin(tile, &m0, &n0);
    const int buf = C::NBUF == 2 ? (ti & 1) : 0;
    const uint32_t tpar = (uint32_t)((C::NBUF == 2 ? (ti >> 1) : ti) & 1);     // phase of the set's barriers
    const int row = q * 32 + lane;
    const int m = m0 + row;
    const bool ok = m < p.M;
    int64_t yo = 0, ro = 0;
    if (ok) {
      const int n = m / p.DHW;
      const int rem = m - n * p.DHW;
      const int n1 = n / d.N2, n2 = n - n1 * d.N2;
      yo = (int64_t)n1 * d.y_s1 + (int64_t)n2 * d.y_s2 + rem;
      if (p.res) {
        int rr = rem;
        if (d.res_up) {
          const int z = rem / p.HW;
          const int r2 = rem - z * p.HW;
          const int hh = r2 / d.W, ww = r2 - hh * d.W;
          rr = (z * (d.H >> 1) + (hh >> 1)) * (d.W >> 1) + (ww >> 1);
        }
        ro = (int64_t)n1 * d.r_s1 + (int64_t)n2 * d.r_s2 + rr;
      }
    }
    const bool lead = blockIdx.z == 0;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * C::SET_COLS);
    // 32 output channels of this thread's pixel.  Everything that has to be READ (previous value for accumulate,
    // residual, bias) is fetched for all 32 channels before the first store: a load-add-store chain per channel
    // would serialise 256 DRAM round trips per tile (measured: +60 % on the per-timestep h-half GEMMs).
    const GruEpi& ge = fp.gru;
    // ConvGRU epilogue for 32 channels (Ch % 32 == 0, so a chunk never straddles the update | reset boundary)
    auto emit_gru32 = [&](int cb, const float* v) {
      const int co0 = n0 + cb;
      if (co0 >= d.Cout) return;
      const int nb = m / p.DHW;                      // image (= batch row) and pixel of this thread
      const int pix = m - nb * p.DHW;
      float pre[32], hp[32] = {}, ug[32] = {};
      const bool state = ge.mode == 2 || co0 >= ge.Ch;        // chunk that produces out2 (r*h or the new h)
      const int c0 = ge.mode == 2 ? co0 : co0 - ge.Ch;        // channel inside the hidden state
#pragma unroll
      for (int j = 0; j < 32; ++j) pre[j] = __ldcg(p.y + yo + (int64_t)(co0 + j) * d.y_cs);
      if (state) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          hp[j] = ge.hprev ? __ldg(ge.hprev + (int64_t)nb * ge.hp_s1 + (int64_t)(c0 + j) * p.DHW + pix) : 0.f;
      }
      if (ge.mode == 2) {
#pragma unroll
        for (int j = 0; j < 32; ++j) ug[j] = __ldcg(ge.ugate + (int64_t)nb * ge.u_s1 + (int64_t)(c0 + j) * p.DHW + pix);
      }
      float o2[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float a = v[j] + pre[j];
        const float g = ge.mode == 2 ? tanhf(a) : sigmoidf_(a);
        p.y[yo + (int64_t)(co0 + j) * d.y_cs] = g;
        if (ge.mode == 2) o2[j] = hp[j] * (1.f - ug[j]) + g * ug[j];
        else o2[j] = g * hp[j];
      }
      if (state) {
#pragma unroll
        for (int j = 0; j < 32; ++j) ge.out2[(int64_t)nb * ge.o2_s1 + (int64_t)(c0 + j) * p.DHW + pix] = o2[j];
        if (ge.pl_hi)
          store_planes32(reinterpret_cast<__nv_bfloat16*>(ge.pl_hi) + (int64_t)m * ge.pl_Cp + c0,
                         reinterpret_cast<__nv_bfloat16*>(ge.pl_lo) + (int64_t)m * ge.pl_Cp + c0, o2, fp.fp16);
      }
    };
    // BPTT epilogues (modes 3 / 4, see GruEpi), eight channels at a time to bound the live registers
    auto emit_bptt32 = [&](int cb, const float* v) {
      const int co0 = n0 + cb;
      if (co0 >= d.Cout) return;
      const int nb = m / p.DHW;
      const int pix = m - nb * p.DHW;
      float* gate = const_cast<float*>(ge.ugate) + (int64_t)nb * ge.u_s1 + pix;        // + channel * DHW
      __nv_bfloat16* ph = reinterpret_cast<__nv_bfloat16*>(ge.pl_hi) + (int64_t)nb * ge.pl_img + (int64_t)pix * ge.pl_Cp;
      __nv_bfloat16* pl = reinterpret_cast<__nv_bfloat16*>(ge.pl_lo) + (int64_t)nb * ge.pl_img + (int64_t)pix * ge.pl_Cp;
      float* cw = ge.out2 + (int64_t)nb * ge.o2_s1 + pix;
#pragma unroll
      for (int s8 = 0; s8 < 4; ++s8) {
        const int c0 = co0 + 8 * s8;
        if (ge.mode == 3) {
          float r[8], h[8], cr[8], dar[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            r[j] = __ldcg(gate + (int64_t)(ge.Ch + c0 + j) * p.DHW);
            h[j] = __ldg(ge.hprev + (int64_t)nb * ge.hp_s1 + (int64_t)(c0 + j) * p.DHW + pix);
            cr[j] = __ldcg(cw + (int64_t)(c0 + j) * p.DHW);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float drh = v[8 * s8 + j];
            cw[(int64_t)(c0 + j) * p.DHW] = fmaf(drh, r[j], cr[j]);
            dar[j] = drh * h[j] * r[j] * (1.f - r[j]);
            gate[(int64_t)(ge.Ch + c0 + j) * p.DHW] = dar[j];
          }
          uint4 hi, lo;
          bf16_split8(dar, &hi, &lo);
          *reinterpret_cast<uint4*>(ph + ge.Ch + c0) = hi;
          *reinterpret_cast<uint4*>(pl + ge.Ch + c0) = lo;
        } else {
          float u[8], o[8], h2[8], dhn[8], dau[8], dao[8];
          const float* cin = ge.carry_in + (int64_t)nb * ge.o2_s1 + pix;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            u[j] = __ldcg(gate + (int64_t)(c0 + j) * p.DHW);
            o[j] = __ldcg(gate + (int64_t)(2 * ge.Ch + c0 + j) * p.DHW);
            h2[j] = ge.hprev ? __ldg(ge.hprev + (int64_t)nb * ge.hp_s1 + (int64_t)(c0 + j) * p.DHW + pix) : 0.f;
            dhn[j] = __ldg(ge.dh_prev + (int64_t)nb * ge.dh_s1 + (int64_t)(c0 + j) * p.DHW + pix) +
                     __ldcg(cin + (int64_t)(c0 + j) * p.DHW) + v[8 * s8 + j];
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            dao[j] = dhn[j] * u[j] * (1.f - o[j] * o[j]);
            dau[j] = dhn[j] * (o[j] - h2[j]) * u[j] * (1.f - u[j]);
            gate[(int64_t)(c0 + j) * p.DHW] = dau[j];
            gate[(int64_t)(2 * ge.Ch + c0 + j) * p.DHW] = dao[j];
            cw[(int64_t)(c0 + j) * p.DHW] = dhn[j] * (1.f - u[j]);
          }
          uint4 hi, lo;
          bf16_split8(dau, &hi, &lo);
          *reinterpret_cast<uint4*>(ph + c0) = hi;
          *reinterpret_cast<uint4*>(pl + c0) = lo;
          bf16_split8(dao, &hi, &lo);
          *reinterpret_cast<uint4*>(ph + 2 * ge.Ch + c0) = hi;
          *reinterpret_cast<uint4*>(pl + 2 * ge.Ch + c0) = lo;
        }
      }
    };
    auto emit32 = [&](int cb, const float* v) {
      if (ge.mode >= 3) { emit_bptt32(cb, v); return; }
      if (ge.mode) { emit_gru32(cb, v); return; }
      const int co0 = n0 + cb;
      if (co0 >= d.Cout) return;
      float add[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) add[j] = 0.f;
      if (!p.atomic_out && d.accumulate) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (co0 + j < d.Cout) add[j] = __ldcg(p.y + yo + (int64_t)(co0 + j) * d.y_cs);
      }
      if (lead && p.res) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (co0 + j < d.Cout) add[j] += __ldg(p.res + ro + (int64_t)(co0 + j) * d.r_cs);
      }
      if (lead && p.bias) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (co0 + j < d.Cout) add[j] += __ldg(p.bias + co0 + j);
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (co0 + j >= d.Cout) break;
        float o = v[j] + add[j];
        float* dst = p.y + yo + (int64_t)(co0 + j) * d.y_cs;
        if (p.atomic_out) {
          atomicAdd(dst, o);
        } else {
          if (d.out_act == 1) o = fmaxf(o, 0.f);
          else if (d.out_act == 2) o = tanhf(o);
          *dst = o;
        }
      }
    };
    // While the main loop runs these warps are idle: pull everything the epilogue will read into L2, so its loads pay
    // an L2 hit instead of a DRAM round trip per 32-channel chunk.
    if (fp.prefetch && !p.atomic_out && (d.accumulate || ge.mode)) mbar_wait(smem_u32(bars.late(buf)), tpar);
    if (fp.prefetch && ok && ge.mode >= 3) {
      const int nb = m / p.DHW, pix = m - nb * p.DHW;
      const int cend = min(BN, d.Cout - n0);
      const float* gate = ge.ugate + (int64_t)nb * ge.u_s1 + pix;
      for (int j = 0; j < cend; ++j) {
        if (EW == 8 && ((j >> 5) & 1) != half) continue;
        const int c = n0 + j;
        if (ge.hprev) prefetch_l2(ge.hprev + (int64_t)nb * ge.hp_s1 + (int64_t)c * p.DHW + pix);
        if (ge.mode == 3) {
          prefetch_l2(gate + (int64_t)(ge.Ch + c) * p.DHW);
          prefetch_l2(ge.out2 + (int64_t)nb * ge.o2_s1 + (int64_t)c * p.DHW + pix);
        } else {
          prefetch_l2(gate + (int64_t)c * p.DHW);
          prefetch_l2(gate + (int64_t)(2 * ge.Ch + c) * p.DHW);
          prefetch_l2(ge.dh_prev + (int64_t)nb * ge.dh_s1 + (int64_t)c * p.DHW + pix);
          prefetch_l2(ge.carry_in + (int64_t)nb * ge.o2_s1 + (int64_t)c * p.DHW + pix);
        }
      }
    } else if (fp.prefetch && ok && !p.atomic_out && (d.accumulate || ge.mode)) {
      const int nb = m / p.DHW, pix = m - nb * p.DHW;
      const int cend = min(BN, d.Cout - n0);
      for (int j = 0; j < cend; ++j) {
        if (EW == 8 && ((j >> 5) & 1) != half) continue;
        prefetch_l2(p.y + yo + (int64_t)(n0 + j) * d.y_cs);
      }
      if (ge.mode) {
        for (int j = 0; j < cend; ++j) {
          if (EW == 8 && ((j >> 5) & 1) != half) continue;
          const int co = n0 + j;
          if (ge.mode == 1 && co < ge.Ch) continue;
          const int c = ge.mode == 2 ? co : co - ge.Ch;
          if (ge.hprev) prefetch_l2(ge.hprev + (int64_t)nb * ge.hp_s1 + (int64_t)c * p.DHW + pix);
          if (ge.mode == 2) prefetch_l2(ge.ugate + (int64_t)nb * ge.u_s1 + (int64_t)c * p.DHW + pix);
        }
      }
    }
    mbar_wait(smem_u32(bars.accum(buf)), tpar);
    tc_fence_after();
    for (int cb = 0; cb < BN; cb += 32) {
      if (n0 + cb >= d.Cout) break;
      if (EW == 8 && ((cb >> 5) & 1) != half) continue;
      float v[32];
      if constexpr (ONEACC) {
        uint32_t r0[32];
        tmem_ld32(taddr + cb, r0);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r0[j]);
      } else {
        uint32_t r0[32], r1[32];
        tmem_ld32(taddr + cb, r0);
        tmem_ld32(taddr + BN + cb, r1);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaf(__uint_as_float(r1[j]), fp.lo_inv, __uint_as_float(r0[j]));
      }
      if (!ok) continue;
      emit32(cb, v);
    }
    tc_fence_before();
    if (PERSIST) {        // this thread's TMEM reads of the set are complete: a later tile's MMAs may overwrite it
      if constexpr (CG == 2) mbar_arrive_cluster(mapa_rank(smem_u32(bars.tfree(buf)), 0));
      else mbar_arrive(smem_u32(bars.tfree(buf)));
    }
   }
  }

  __syncthreads();
  if (CG == 2) cluster_sync_all();        // no CTA leaves while the peer may still signal its barriers / read its smem
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CG == 2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS));
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS));
  }
}

// ------------------------------------------------------------------------------------------------ wgrad kernel
// A = X^T (rows ci, MN-major), B = dY^T (rows co, MN-major), k = 64 pixels per stage.
struct WgP {
  ConvP c;
  TileGeom g;           // box of 64 pixels
  int nsplit, per_split;   // pixel range per CTA (multiple of 64)
  int y_c_off;             // shared dY planes: channel offset, frames per clip in the planes, first frame, frames used
  int y_T, y_t_off, y_n2;  // (y_T = 0: the planes hold exactly this conv's images)
};

template <int BN, int CG>
struct WCfg {
  static constexpr int A_BYTES = 2 * 64 * 128;           // 128 ci = two 64-wide MN blocks of [64 k][128 B]
  static constexpr int B_BLOCKS = BN / CG / 64;          // 64-wide co blocks staged by this CTA (pair: half of BN)
  static constexpr int B_BYTES = B_BLOCKS * 64 * 128;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int STAGES_FIT = (200 * 1024) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_FIT > 5 ? 5 : STAGES_FIT;
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 + 256;
  static constexpr int TMEM_COLS = tmem_cols_for(2 * BN);
  static_assert(BN % (64 * CG) == 0, "whole 64-wide co blocks per CTA");
};

// CG = 2: CTA pair = 256 ci x BN co; each CTA stages its own 128 ci of X and half of the dY blocks.
template <int BN, int CG>
__global__ void __launch_bounds__(NT, 1)
conv_tma_wgrad_kernel(const __grid_constant__ CUtensorMap tmX_hi, const __grid_constant__ CUtensorMap tmX_lo,
                      const __grid_constant__ CUtensorMap tmY_hi, const __grid_constant__ CUtensorMap tmY_lo,
                      const WgP wp, float* __restrict__ dwp) {
  using C = WCfg<BN, CG>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const Bars<C::STAGES> bars{reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES)};
  uint64_t* full_bar = bars.full(0);
  uint64_t* empty_bar = bars.empty(0);
  uint32_t* tmem_slot = bars.slot();

  const ConvP& p = wp.c;
  const dvd_conv_desc& d = p.d;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t cx = CG == 2 ? cluster_ctaid_x() : 0u;
  const bool leader = CG == 1 || cx == 0;

  if (tid == 0) bars.init(128u * CG);
  if (warp == 1) {
    if constexpr (CG == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"((uint32_t)C::TMEM_COLS));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"((uint32_t)C::TMEM_COLS));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int ci0 = blockIdx.x * BM;
  const int co0 = blockIdx.y * BN;
  // tap fastest: the CTAs of one wave sweep the SAME pixel range for all taps / channel blocks, so the X and dY planes
  // come out of L2 for all but the first of them (tap-major order re-read both planes from DRAM once per tap: 24x the
  // algorithmic bytes in the ncu capture of profiles/)
  const int split = blockIdx.z / p.taps;
  const int tap = blockIdx.z - split * p.taps;
  const int m_lo = split * wp.per_split;
  const int m_hi = min(m_lo + wp.per_split, p.M);
  const int n_iters = (m_hi - m_lo + BKC - 1) / BKC;

  if (warp == 0) {
    if (lane == 0) {
      const int kw = tap % d.kW;
      const int t2 = tap / d.kW;
      const int kh = t2 % d.kH;
      const int kd = t2 / d.kH;
      const int ox = kw - d.kW / 2, oy = kh - d.kH / 2, oz = kd - d.kD / 2;
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < n_iters; ++it) {
        const int mk = m_lo + it * BKC;
        const int img = mk / p.DHW;
        int rem = mk - img * p.DHW;
        const int z0 = rem / p.HW;
        rem -= z0 * p.HW;
        const int y0 = rem / d.W;
        const int x0 = rem - y0 * d.W;
        // dY may live in planes shared with other convs of the layer: frame t of clip b is image b * y_T + t + y_t_off
        const int yimg = wp.y_T ? (img / wp.y_n2) * wp.y_T + (img % wp.y_n2) + wp.y_t_off : img;
        mbar_wait(smem_u32(empty_bar + stage), phase ^ 1);
        const uint32_t fb = smem_u32(full_bar + stage);
        const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
        if constexpr (CG == 2) {
          if (leader) mbar_expect_tx(fb, 2 * C::STAGE_BYTES);
          const uint32_t fbl = mapa_rank(fb, 0);
          const int cob = wp.y_c_off + co0 + (int)cx * (BN / 2);    // this CTA's half of the pair's co range
#pragma unroll
          for (int b = 0; b < 2; ++b) {
            tma_load_5d_pair(sa + b * 8192, &tmX_hi, fbl, ci0 + b * 64, x0 + ox, y0 + oy, z0 + oz, img);
            tma_load_5d_pair(sa + C::A_BYTES + b * 8192, &tmX_lo, fbl, ci0 + b * 64, x0 + ox, y0 + oy, z0 + oz, img);
          }
#pragma unroll
          for (int b = 0; b < C::B_BLOCKS; ++b) {
            tma_load_5d_pair(sa + 2 * C::A_BYTES + b * 8192, &tmY_hi, fbl, cob + b * 64, x0, y0, z0, yimg);
            tma_load_5d_pair(sa + 2 * C::A_BYTES + C::B_BYTES + b * 8192, &tmY_lo, fbl, cob + b * 64, x0, y0, z0, yimg);
          }
        } else {
          mbar_expect_tx(fb, C::STAGE_BYTES);
#pragma unroll
          for (int b = 0; b < 2; ++b) {
            tma_load_5d(sa + b * 8192, &tmX_hi, fb, ci0 + b * 64, x0 + ox, y0 + oy, z0 + oz, img);
            tma_load_5d(sa + C::A_BYTES + b * 8192, &tmX_lo, fb, ci0 + b * 64, x0 + ox, y0 + oy, z0 + oz, img);
          }
#pragma unroll
          for (int b = 0; b < C::B_BLOCKS; ++b) {
            const int yc = wp.y_c_off + co0 + b * 64;
            tma_load_5d(sa + 2 * C::A_BYTES + b * 8192, &tmY_hi, fb, yc, x0, y0, z0, yimg);
            tma_load_5d(sa + 2 * C::A_BYTES + C::B_BYTES + b * 8192, &tmY_lo, fb, yc, x0, y0, z0, yimg);
          }
        }
        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      // MN-major operands: LBO = 8192 (next 64-wide MN block), SBO = 1024 (next 8 k-rows); one UMMA_K step =
      // 16 k-rows of 128 B = 2048 B = 128 units.
      int stage = 0;
      uint32_t phase = 0;
      mma_issue_loop<BN, false, C::STAGES, C::STAGE_BYTES, C::A_BYTES, C::B_BYTES, CG>(
          smem, bars, tmem_base, tmem_base + BN, 0, n_iters, 1, 8192, 1024, 128, 1u, stage, phase);
    }
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int ci = ci0 + q * 32 + lane;
    const bool ok = ci < d.Cin;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    float* dst = dwp + ((int64_t)tap * d.Cin + ci) * d.Cout + co0;
    mbar_wait(smem_u32(bars.accum(0)), 0);
    tc_fence_after();
    for (int cb = 0; cb < BN; cb += 32) {
      if (co0 + cb >= d.Cout) break;
      uint32_t r0[32], r1[32];
      tmem_ld32(taddr + cb, r0);
      tmem_ld32(taddr + BN + cb, r1);
      tmem_ld_wait();
      if (!ok) continue;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (co0 + cb + j >= d.Cout) break;
        const float v = __uint_as_float(r0[j]) + __uint_as_float(r1[j]);
        if (p.atomic_out) atomicAdd(dst + cb + j, v);
        else dst[cb + j] = v;
      }
    }
    tc_fence_before();
  }

  __syncthreads();
  if (CG == 2) cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CG == 2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS));
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS));
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeFn get_encode() {
  static EncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {          // a driver entry point: process-wide, not per device
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(ptr);
  });
  return fn;
}

// channels-last planes [N][D][H][W][Cp] -> 5-D map (c, w, h, d, n), box (64, bw, bh, bd, bn), 128B swizzle
static int make_act_map(CUtensorMap* tm, const void* base, int N, int D, int H, int W, int Cp, const TileGeom& g,
                        int64_t img_stride_elems = 0) {
  EncodeFn enc = get_encode();
  if (!enc) return fail("cuTensorMapEncodeTiled unavailable%s (%s:%d)", "", __FILE__, __LINE__);
  cuuint64_t gdim[5] = {(cuuint64_t)Cp, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
  cuuint64_t gstr[4] = {(cuuint64_t)Cp * 2, (cuuint64_t)W * Cp * 2, (cuuint64_t)H * W * Cp * 2,
                        img_stride_elems ? (cuuint64_t)img_stride_elems * 2 : (cuuint64_t)D * H * W * Cp * 2};
  cuuint32_t box[5] = {64, (cuuint32_t)g.bw, (cuuint32_t)g.bh, (cuuint32_t)g.bd, (cuuint32_t)g.bn};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled (activation) failed%s (%s:%d)", "", __FILE__, __LINE__);
  return 0;
}

static int make_w_map(CUtensorMap* tm, const void* base, int rows, int Cp, int box_rows) {
  EncodeFn enc = get_encode();
  if (!enc) return fail("cuTensorMapEncodeTiled unavailable%s (%s:%d)", "", __FILE__, __LINE__);
  cuuint64_t gdim[2] = {(cuuint64_t)Cp, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)Cp * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled (weights) failed%s (%s:%d)", "", __FILE__, __LINE__);
  return 0;
}

// rows consecutive flat pixels (rows = 128 or 64) as a box (bn, bd, bh, bw); false if the geometry does not tile
static bool tile_geom(int rows, int N, int D, int H, int W, TileGeom* g) {
  if (W > rows) {                      // a fraction of one image row
    if (W % rows) return false;
    g->bw = rows; g->bh = g->bd = g->bn = 1;
    return true;
  }
  if (rows % W) return false;
  g->bw = W;
  int left = rows / W;
  g->bh = H < left ? H : left;
  if (H % g->bh || left % g->bh) return false;
  left /= g->bh;
  g->bd = D < left ? D : left;
  if (D % g->bd || left % g->bd) return false;
  left /= g->bd;
  g->bn = left;
  if (g->bn > 256) return false;
  // a box must not straddle: if it spans several rows it must span full rows etc. (guaranteed by construction:
  // bh < H only when bd = bn = 1; bd < D only when bn = 1)
  if (g->bh < H && (g->bd != 1 || g->bn != 1)) return false;
  if (g->bd < D && g->bn != 1) return false;
  (void)N;
  return true;
}

static inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

// Stream-ordered scratch for the operand planes of one call: cudaMallocFromPoolAsync from a memory pool the library
// keeps PER (device, stream).  One shared pool served a single stream well, but with the ConvGRU layers on one stream
// each (and the batch chains on helper streams) blocks freed on one stream are not reusable on another until that free
// has been reached, so a shared pool kept growing by fresh driver allocations at timing-dependent moments in the middle
// of a step (measured: the same configuration 964 or 1394 ms per step).  A pool per stream re-uses its own blocks in
// stream order: after the first step nothing is allocated from the driver.  The pools keep what they were given
// (release threshold = max); dvd_scratch_bytes() reports the sum of their high-water marks.
struct PoolEntry { int dev; cudaStream_t st; cudaMemPool_t pool; };
static std::mutex g_pool_mu;
static std::vector<PoolEntry> g_pools;

static int scratch_pool(cudaStream_t s, cudaMemPool_t* out) {
  int dev = 0;
  DVD_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_pool_mu);
  for (const PoolEntry& e : g_pools)
    if (e.dev == dev && e.st == s) { *out = e.pool; return 0; }
  cudaMemPoolProps props;
  memset(&props, 0, sizeof(props));
  props.allocType = cudaMemAllocationTypePinned;
  props.handleTypes = cudaMemHandleTypeNone;
  props.location.type = cudaMemLocationTypeDevice;
  props.location.id = dev;
  cudaMemPool_t pool;
  DVD_CUDA(cudaMemPoolCreate(&pool, &props));
  uint64_t thr = UINT64_MAX;
  DVD_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
  g_pools.push_back({dev, s, pool});
  *out = pool;
  return 0;
}

struct Scratch {
  void* ptr = nullptr;
  cudaStream_t st;
  int alloc(size_t bytes, cudaStream_t s) {
    st = s;
    cudaMemPool_t pool;
    DVD_TRY(scratch_pool(s, &pool));
    DVD_CUDA(cudaMallocFromPoolAsync(&ptr, bytes, pool, s));
    return 0;
  }
  ~Scratch() {
    if (ptr) cudaFreeAsync(ptr, st);
  }
};

static int prep_planes(const float* src, int N1, int N2, int C, int Cp, int64_t s1, int64_t s2, int64_t cs, int in_pix,
                       int out_pix, int W, int HW, int up, int relu, int fp16, __nv_bfloat16* hi, __nv_bfloat16* lo,
                       cudaStream_t st) {
  PrepP p;
  p.src = src; p.hi = hi; p.lo = lo; p.N2 = N2; p.C = C; p.Cp = Cp; p.s1 = s1; p.s2 = s2; p.cs = cs;
  p.in_pix = in_pix; p.out_pix = out_pix; p.W = W; p.HW = HW; p.up = up; p.relu = relu; p.fp16 = fp16;
  const int N = N1 * N2;
  DVD_CHECK_ARG(N <= 65535);          // gridDim.z
  const bool vec = !up && in_pix == out_pix && in_pix % 4 == 0 && cs % 4 == 0 && s1 % 4 == 0 && s2 % 4 == 0 &&
                   (reinterpret_cast<uintptr_t>(src) & 15) == 0 && out_pix >= 64;
  prof_tag("prep N%d pix%d C%d", N, out_pix, Cp);
  prof_begin(2, (double)N * out_pix * Cp * 8.0, st);          // "flops" = bytes moved (4 in + 4 out per element)
  if (vec) prep_planes_vec_kernel<<<dim3(ceil_div(out_pix, 64), Cp / 64, N), 256, 0, st>>>(p);
  else prep_planes_kernel<<<dim3(ceil_div(out_pix, 32), Cp / 64, N), 256, 0, st>>>(p);
  prof_end(2, st);
  DVD_LAUNCH_CHECK();
  return 0;
}

template <int BN, int CG, int OCC = 1, int EW = 4, bool PERSIST = false, bool ONEACC = false>
static int launch_fwd(const CUtensorMap* m, const FwdP& fp, dim3 grid, cudaStream_t st) {
  using C = Cfg<BN, CG, OCC, PERSIST, ONEACC>;
  constexpr int NT = 64 + 32 * EW;
  static std::atomic<uint64_t> configured{0};         // one bit per device: the attribute is per (function, device)
  if (!device_bit_test_and_set(configured))
    DVD_CUDA(cudaFuncSetAttribute(conv_tma_fwd_kernel<BN, CG, OCC, EW, PERSIST, ONEACC>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
  if (PERSIST) {          // one CTA (pair) per SM (pair), or fewer when there are fewer tiles
    const int slots = num_sms() / CG;
    grid = dim3((unsigned)(CG * std::min(fp.tiles, slots)), 1, 1);
  }
  if (CG == 2) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = C::SMEM; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    DVD_CUDA(cudaLaunchKernelEx(&cfg, conv_tma_fwd_kernel<BN, CG, OCC, EW, PERSIST, ONEACC>, m[0], m[1], m[2], m[3], fp));
  } else {
    conv_tma_fwd_kernel<BN, CG, OCC, EW, PERSIST, ONEACC><<<grid, NT, C::SMEM, st>>>(m[0], m[1], m[2], m[3], fp);
  }
  return 0;
}
template <int BN, int CG>
static int launch_wgrad(const CUtensorMap* m, const WgP& wp, float* dwp, dim3 grid, cudaStream_t st) {
  using C = WCfg<BN, CG>;
  static std::atomic<uint64_t> configured{0};
  if (!device_bit_test_and_set(configured))
    DVD_CUDA(cudaFuncSetAttribute(conv_tma_wgrad_kernel<BN, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
  if (CG == 2) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = C::SMEM; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    DVD_CUDA(cudaLaunchKernelEx(&cfg, conv_tma_wgrad_kernel<BN, CG>, m[0], m[1], m[2], m[3], wp, dwp));
  } else {
    conv_tma_wgrad_kernel<BN, CG><<<grid, NT, C::SMEM, st>>>(m[0], m[1], m[2], m[3], wp, dwp);
  }
  return 0;
}

// widest tile with the least padded columns (ties -> wider)
static int pick_bn(int Cout, bool allow_192) {
  if (Cout <= 64) return 64;
  const int cand[3] = {256, 192, 128};
  int best = 128, best_waste = 1 << 30;
  for (int i = 0; i < 3; ++i) {
    const int bn = cand[i];
    if (bn == 192 && !allow_192) continue;
    const int waste = (Cout + bn - 1) / bn * bn - Cout;
    if (waste < best_waste) { best = bn; best_waste = waste; }
  }
  return best;
}

}  // namespace tma

bool tma_fwd_eligible(const ConvP& p) {
  if (get_option(OPT_SIMT_ONLY)) return false;
  const dvd_conv_desc& d = p.d;
  // narrow layers (3-channel image convs) ride the tensor path, zero-padded to one 64-wide block, once there are
  // enough pixels for the padding not to matter: the SIMT engine runs them at < 1 TFLOP/s
  if (p.M < 128 || ((d.Cin < 32 || d.Cout < 64) && p.M < (1 << 18))) return false;
  if ((int64_t)d.N1 * d.N2 > 65535) return false;
  tma::TileGeom g;
  return tma::tile_geom(128, d.N1 * d.N2, d.D, d.H, d.W, &g) && tma::get_encode() != nullptr;
}

int tma_fwd_launch(ConvP& p, cudaStream_t st) { return tma_fwd_launch_ex(p, nullptr, nullptr, st); }
bool tma_fwd_launch_ex_eligible(const ConvP& p) { return tma_fwd_eligible(p); }

// forward operands (x_kind = 1) are split into fp16 planes unless "fwd_bf16" / "oneacc" ask for bf16 everywhere
bool tma_forward_planes_fp16() { return !get_option(OPT_FWD_BF16) && !get_option(OPT_ONEACC); }

int tma_split_weights(const float* w_packed, int taps, int Cin, int Cout, int CoutP, int fp16, void* hi, void* lo,
                      cudaStream_t st) {
  // [tap][Cin][Cout] fp32 -> [tap][CoutP][CinP]   (image = tap, channel = cin, pixel = cout)
  return tma::prep_planes(w_packed, taps, 1, Cin, tma_round64(Cin), (int64_t)Cin * Cout, 0, Cout, Cout, CoutP, 1, 1, 0,
                          0, fp16, reinterpret_cast<__nv_bfloat16*>(hi), reinterpret_cast<__nv_bfloat16*>(lo), st);
}
int tma_split_activations(const float* x, int N, int C, int64_t n_stride, int64_t c_stride, int pix, int fp16, void* hi,
                          void* lo, cudaStream_t st) {
  return tma::prep_planes(x, N, 1, C, tma_round64(C), n_stride, 0, c_stride, pix, pix, 1, 1, 0, 0, fp16,
                          reinterpret_cast<__nv_bfloat16*>(hi), reinterpret_cast<__nv_bfloat16*>(lo), st);
}

int tma_fwd_launch_ex(ConvP& p, const TmaOperands* ops, const GruEpi* epi, cudaStream_t st) {
  using namespace tma;
  const dvd_conv_desc& d = p.d;
  const int nsm = num_sms();
  const int N = d.N1 * d.N2;
  FwdP fp;
  if (!tile_geom(128, N, d.D, d.H, d.W, &fp.g)) return fail("internal: geometry%s (%s:%d)", "", __FILE__, __LINE__);
  const int CinP = round_up(d.Cin, 64);
  p.ck = CinP / 64;
  p.iters_total = p.taps * p.ck;
  int bn = pick_bn(d.Cout, true);
  const int mt = ceil_div(p.M, BM);
  if (bn > 128 && (int64_t)mt * ceil_div(d.Cout, bn) < nsm) bn = 128;        // small grids: more, narrower tiles
  // short reductions prefer 128-wide tiles with two CTAs per SM over one 256-wide tile (+7 % on the 3x3 convs at
  // 256 channels)
  if (bn > 128 && p.iters_total <= 40 && d.Cout % 128 == 0 && !(epi && epi->mode)) bn = 128;
  const bool ext_w = ops && ops->w_hi, ext_a = ops && ops->a_hi;
  const int CoutP = ext_w ? ops->CoutP : round_up(d.Cout, bn);
  fp.CoutP = CoutP;
  if (epi) fp.gru = *epi;
  fp.prefetch = get_option(OPT_EPI_PREFETCH);
  if (fp.gru.mode) DVD_CHECK_ARG(fp.gru.Ch % 32 == 0 && !d.out_act && !p.res && !p.bias);
  if (fp.gru.mode == 1 || fp.gru.mode == 2) DVD_CHECK_ARG(d.accumulate);
  if (fp.gru.mode >= 3) DVD_CHECK_ARG(!d.accumulate && d.Cout == fp.gru.Ch && fp.gru.pl_hi && fp.gru.pl_lo && fp.gru.out2);
  fp.fp16 = (d.x_kind == 1 && tma_forward_planes_fp16()) ? 1 : 0;
  fp.lo_inv = fp.fp16 ? 1.f / kLoScaleFp16 : 1.f;
  const int64_t ctas = (int64_t)mt * ceil_div(d.Cout, bn);
  int nsplit = 1;
  if (ctas < nsm && d.out_act == 0 && p.iters_total >= 8 && !fp.gru.mode) {
    nsplit = (int)ceil_div<int64_t>(nsm, ctas);
    const int maxs = p.iters_total / 4;
    if (nsplit > maxs) nsplit = maxs;
    if (nsplit > 16) nsplit = 16;
    if (nsplit < 1) nsplit = 1;
  }
  const int per = ceil_div(p.iters_total, nsplit);
  nsplit = ceil_div(p.iters_total, per);
  p.nsplit = nsplit;
  p.iters_per_split = per;
  p.atomic_out = nsplit > 1;
  if (p.atomic_out && !d.accumulate) DVD_TRY(zero_output_view(p, st));

  // ---- operand preparation in stream-ordered scratch
  const int64_t pix = (int64_t)p.DHW;
  const size_t a_elems = (size_t)N * pix * CinP;
  const size_t w_elems = (size_t)p.taps * CoutP * CinP;
  Scratch sc;
  const size_t need = (ext_a ? 0 : 2 * a_elems) + (ext_w ? 0 : 2 * w_elems);
  if (need) DVD_TRY(sc.alloc(need * sizeof(__nv_bfloat16) + 1024, st));
  __nv_bfloat16* cur = reinterpret_cast<__nv_bfloat16*>(sc.ptr);
  const __nv_bfloat16 *a_hi, *a_lo, *w_hi, *w_lo;
  if (ext_a) {
    a_hi = reinterpret_cast<const __nv_bfloat16*>(ops->a_hi);
    a_lo = reinterpret_cast<const __nv_bfloat16*>(ops->a_lo);
  } else {
    __nv_bfloat16* h = cur; __nv_bfloat16* l = cur + a_elems; cur += 2 * a_elems;
    DVD_TRY(prep_planes(p.x, d.N1, d.N2, d.Cin, CinP, d.x_s1, d.x_s2, d.x_cs, p.DHW, p.DHW, d.W, p.HW, d.in_up,
                        d.in_relu, fp.fp16, h, l, st));
    a_hi = h; a_lo = l;
  }
  if (ext_w) {
    w_hi = reinterpret_cast<const __nv_bfloat16*>(ops->w_hi);
    w_lo = reinterpret_cast<const __nv_bfloat16*>(ops->w_lo);
  } else {
    __nv_bfloat16* h = cur; __nv_bfloat16* l = cur + w_elems;
    // weights: [tap][Cin][Cout] fp32 -> [tap][CoutP][CinP]   (image = tap, channel = cin, pixel = cout)
    DVD_TRY(prep_planes(p.w, p.taps, 1, d.Cin, CinP, (int64_t)d.Cin * d.Cout, 0, d.Cout, d.Cout, CoutP, 1, 1, 0, 0,
                        fp.fp16, h, l, st));
    w_hi = h; w_lo = l;
  }
  CUtensorMap maps[4];
  // CTA pairs (256 x bn tiles) whenever there is more than one wave of tiles
  const bool pair = get_option(OPT_PAIR) && mt % 2 == 0 && ctas >= nsm && bn >= 64;
  const int a_Cp = (ext_a && ops->a_Cp) ? ops->a_Cp : CinP;
  const int64_t a_stride = ext_a ? ops->a_img_stride : 0;
  fp.a_c_off = ext_a ? ops->a_c_off : 0;
  if (ext_a) DVD_CHECK_ARG(a_Cp % 64 == 0 && fp.a_c_off % 64 == 0 && fp.a_c_off + CinP <= a_Cp);
  DVD_TRY(make_act_map(&maps[0], a_hi, N, d.D, d.H, d.W, a_Cp, fp.g, a_stride));
  DVD_TRY(make_act_map(&maps[1], a_lo, N, d.D, d.H, d.W, a_Cp, fp.g, a_stride));
  DVD_TRY(make_w_map(&maps[2], w_hi, p.taps * CoutP, CinP, pair ? bn / 2 : bn));
  DVD_TRY(make_w_map(&maps[3], w_lo, p.taps * CoutP, CinP, pair ? bn / 2 : bn));
  fp.c = p;
  dim3 grid(mt, ceil_div(d.Cout, bn), nsplit);
  // persistent CTA pairs when there is more than one wave of tiles
  fp.nt = ceil_div(d.Cout, bn);
  fp.tiles = (mt / (pair ? 2 : 1)) * fp.nt;
  const bool persist = get_option(OPT_PERSIST) && pair && nsplit == 1 && bn >= 128 && fp.tiles > nsm / 2;
  const bool oneacc = persist && get_option(OPT_ONEACC) && !fp.fp16;
  // short reductions on narrow tiles: two CTAs per SM
  const bool occ2 = get_option(OPT_OCC2) && !fp.gru.mode && bn <= 128 && p.iters_total <= 40 &&
                    ctas >= 2 * (int64_t)nsm && (pair || bn == 64);     // (one CTA, 128 wide) stages are 64 KB: no room
  prof_tag(pair ? "fwd M%d Ci%d Co%d t%d bn%d pair acc%d" : "fwd M%d Ci%d Co%d t%d bn%d acc%d", p.M, d.Cin, d.Cout,
           p.taps, bn, d.accumulate + 2 * (nsplit > 1) + 4 * (fp.gru.mode != 0));
  prof_begin(0, 2.0 * p.M * (double)d.Cout * d.Cin * p.taps, st);
  int rc;
  if (occ2) {
    if (pair) rc = bn == 128 ? launch_fwd<128, 2, 2>(maps, fp, grid, st) : launch_fwd<64, 2, 2>(maps, fp, grid, st);
    else rc = launch_fwd<64, 1, 2>(maps, fp, grid, st);
  } else if (persist && oneacc) {           // two accumulator sets: the epilogue is hidden under the next tile
    if (bn == 256) rc = launch_fwd<256, 2, 1, 8, true, true>(maps, fp, grid, st);
    else if (bn == 192) rc = launch_fwd<192, 2, 1, 8, true, true>(maps, fp, grid, st);
    else rc = launch_fwd<128, 2, 1, 8, true, true>(maps, fp, grid, st);
  } else if (persist) {                     // persistent CTA pairs, 8 epilogue warps
    if (bn == 256) rc = launch_fwd<256, 2, 1, 8, true>(maps, fp, grid, st);
    else if (bn == 192) rc = launch_fwd<192, 2, 1, 8, true>(maps, fp, grid, st);
    else rc = launch_fwd<128, 2, 1, 8, true>(maps, fp, grid, st);
  } else if (pair) {
    if (bn == 256) rc = launch_fwd<256, 2, 1, 8>(maps, fp, grid, st);
    else if (bn == 192) rc = launch_fwd<192, 2, 1, 8>(maps, fp, grid, st);
    else if (bn == 128) rc = launch_fwd<128, 2, 1, 8>(maps, fp, grid, st);
    else rc = launch_fwd<64, 2>(maps, fp, grid, st);
  } else {
    if (bn == 256) rc = launch_fwd<256, 1>(maps, fp, grid, st);
    else if (bn == 192) rc = launch_fwd<192, 1>(maps, fp, grid, st);
    else if (bn == 128) rc = launch_fwd<128, 1, 1, 8>(maps, fp, grid, st);
    else rc = launch_fwd<64, 1>(maps, fp, grid, st);
  }
  prof_end(0, st);
  if (rc) return rc;
  DVD_LAUNCH_CHECK();
  return 0;
}

bool tma_wgrad_eligible(const ConvP& p) {
  if (get_option(OPT_SIMT_ONLY)) return false;
  const dvd_conv_desc& d = p.d;
  if (p.M < 4096 || d.in_up || ((d.Cin < 32 || d.Cout < 64) && p.M < (1 << 18))) return false;
  if ((int64_t)d.N1 * d.N2 > 65535) return false;
  if (p.DHW % 64 != 0 && 64 % p.DHW != 0) return false;
  tma::TileGeom g;
  return tma::tile_geom(64, d.N1 * d.N2, d.D, d.H, d.W, &g) && tma::get_encode() != nullptr;
}

int tma_scratch_alloc(void** p, size_t bytes, cudaStream_t st) {
  tma::Scratch sc;
  DVD_TRY(sc.alloc(bytes, st));
  *p = sc.ptr;
  sc.ptr = nullptr;          // ownership passes to the caller
  return 0;
}
void tma_scratch_free(void* p, cudaStream_t st) {
  if (p) cudaFreeAsync(p, st);
}
// sums over the scratch pools of the current device: high-water marks of used memory, memory currently reserved
int tma_scratch_stats(long long* high_water, long long* reserved) {
  int dev = 0;
  DVD_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(tma::g_pool_mu);
  long long hw = 0, res = 0;
  for (const tma::PoolEntry& e : tma::g_pools) {
    if (e.dev != dev) continue;
    uint64_t a = 0, b = 0;
    DVD_CUDA(cudaMemPoolGetAttribute(e.pool, cudaMemPoolAttrUsedMemHigh, &a));
    DVD_CUDA(cudaMemPoolGetAttribute(e.pool, cudaMemPoolAttrReservedMemCurrent, &b));
    hw += (long long)a;
    res += (long long)b;
  }
  *high_water = hw;
  *reserved = res;
  return 0;
}
int tma_split_gradients(const float* g, int N, int C, int64_t n_stride, int64_t c_stride, int pix, void* hi, void* lo,
                        cudaStream_t st) {
  return tma_split_activations(g, N, C, n_stride, c_stride, pix, 0, hi, lo, st);
}

// groups of 8 operand elements that were clamped to fp16's range since the last reset (current device); synchronises
int tma_saturation_count(unsigned int* count, int reset, cudaStream_t st) {
  DVD_CUDA(cudaMemcpyFromSymbolAsync(count, tma::g_fp16_saturated, sizeof(unsigned int), 0, cudaMemcpyDeviceToHost, st));
  if (reset) {
    const unsigned int zero = 0;
    DVD_CUDA(cudaMemcpyToSymbolAsync(tma::g_fp16_saturated, &zero, sizeof(unsigned int), 0, cudaMemcpyHostToDevice, st));
  }
  DVD_CUDA(cudaStreamSynchronize(st));
  return 0;
}
int tma_wgrad_launch(ConvP& p, float* dwp, cudaStream_t st) { return tma_wgrad_launch_ex(p, dwp, nullptr, st); }
// shared dY planes need whole images per 64-pixel box unless they map one to one
bool tma_wgrad_ex_ok(const ConvP& p, const TmaWgOperands* ops) {
  if (!ops || !ops->y_hi) return true;
  if (ops->y_Cp % 64 || ops->y_c_off % 64 || ops->y_c_off + p.d.Cout > ops->y_Cp) return false;
  const bool identity = ops->y_T == p.d.N2 && ops->y_t_off == 0;
  return identity || p.DHW % 64 == 0;
}

int tma_wgrad_launch_ex(ConvP& p, float* dwp, const TmaWgOperands* ops, cudaStream_t st) {
  using namespace tma;
  const dvd_conv_desc& d = p.d;
  const int nsm = num_sms();
  const int N = d.N1 * d.N2;
  WgP wp;
  if (!tile_geom(64, N, d.D, d.H, d.W, &wp.g)) return fail("internal: geometry%s (%s:%d)", "", __FILE__, __LINE__);
  const int bn = pick_bn(d.Cout, false);
  const bool pair = get_option(OPT_PAIR) && d.Cin % 256 == 0 && bn >= 128;
  const int64_t base = (int64_t)ceil_div(d.Cin, BM) * ceil_div(d.Cout, bn) * p.taps;
  // split the pixel range: cost model = waves of CTAs x (k-blocks per CTA + a fixed per-CTA cost of ~40 k-blocks for
  // the prologue and the atomic epilogue); whole waves matter for the long reductions (M = millions of pixels), the
  // fixed cost for the short ones
  int nsplit = 1;
  {
    const int maxs = std::max(1, std::min(32, p.M / 2048));
    double best = 1e30;
    for (int ns = 1; ns <= maxs; ++ns) {
      const int64_t waves = ceil_div<int64_t>(base * ns, nsm);
      const double cost = (double)waves * (ceil_div(ceil_div(p.M, ns), BKC) + 40.0);
      if (cost < best * 0.99) { best = cost; nsplit = ns; }
    }
  }
  int per = ceil_div(p.M, nsplit);
  per = round_up(per, BKC);
  nsplit = ceil_div(p.M, per);
  p.atomic_out = (nsplit > 1) || d.accumulate;
  if (nsplit > 1 && !d.accumulate)
    DVD_CUDA(cudaMemsetAsync(dwp, 0, sizeof(float) * (size_t)p.taps * d.Cin * d.Cout, st));
  wp.nsplit = nsplit;
  wp.per_split = per;

  const bool ext_y = ops && ops->y_hi;
  const int CinP = round_up(d.Cin, 64), CoutP = ext_y ? ops->y_Cp : round_up(d.Cout, 64);
  const size_t x_elems = (size_t)N * p.DHW * CinP, y_elems = ext_y ? 0 : (size_t)N * p.DHW * CoutP;
  Scratch sc;
  DVD_TRY(sc.alloc((2 * x_elems + 2 * y_elems) * sizeof(__nv_bfloat16) + 1024, st));
  __nv_bfloat16* x_hi = reinterpret_cast<__nv_bfloat16*>(sc.ptr);
  __nv_bfloat16* x_lo = x_hi + x_elems;
  const __nv_bfloat16 *y_hi, *y_lo;
  DVD_TRY(prep_planes(p.x, d.N1, d.N2, d.Cin, CinP, d.x_s1, d.x_s2, d.x_cs, p.DHW, p.DHW, d.W, p.HW, 0, d.in_relu,
                      0, x_hi, x_lo, st));
  int y_images = N;
  wp.y_c_off = 0; wp.y_T = 0; wp.y_t_off = 0; wp.y_n2 = d.N2;
  if (ext_y) {
    y_hi = reinterpret_cast<const __nv_bfloat16*>(ops->y_hi);
    y_lo = reinterpret_cast<const __nv_bfloat16*>(ops->y_lo);
    y_images = d.N1 * ops->y_T;
    wp.y_c_off = ops->y_c_off;
    if (!(ops->y_T == d.N2 && ops->y_t_off == 0)) { wp.y_T = ops->y_T; wp.y_t_off = ops->y_t_off; }
  } else {
    __nv_bfloat16* h = x_lo + x_elems; __nv_bfloat16* l = h + y_elems;
    DVD_TRY(prep_planes(p.y, d.N1, d.N2, d.Cout, CoutP, d.y_s1, d.y_s2, d.y_cs, p.DHW, p.DHW, d.W, p.HW, 0, 0, 0, h, l, st));
    y_hi = h; y_lo = l;
  }
  CUtensorMap maps[4];
  DVD_TRY(make_act_map(&maps[0], x_hi, N, d.D, d.H, d.W, CinP, wp.g));
  DVD_TRY(make_act_map(&maps[1], x_lo, N, d.D, d.H, d.W, CinP, wp.g));
  DVD_TRY(make_act_map(&maps[2], y_hi, y_images, d.D, d.H, d.W, CoutP, wp.g));
  DVD_TRY(make_act_map(&maps[3], y_lo, y_images, d.D, d.H, d.W, CoutP, wp.g));
  wp.c = p;
  dim3 grid(ceil_div(d.Cin, BM), ceil_div(d.Cout, bn), p.taps * nsplit);
  prof_tag(pair ? "wgrad M%d Ci%d Co%d t%d bn%d pair ns%d" : "wgrad M%d Ci%d Co%d t%d bn%d ns%d", p.M, d.Cin, d.Cout,
           p.taps, bn, nsplit);
  prof_begin(1, 2.0 * p.M * (double)d.Cout * d.Cin * p.taps, st);
  int rc;
  if (pair) {
    rc = bn == 256 ? launch_wgrad<256, 2>(maps, wp, dwp, grid, st) : launch_wgrad<128, 2>(maps, wp, dwp, grid, st);
  } else {
    if (bn == 256) rc = launch_wgrad<256, 1>(maps, wp, dwp, grid, st);
    else if (bn == 128) rc = launch_wgrad<128, 1>(maps, wp, dwp, grid, st);
    else rc = launch_wgrad<64, 1>(maps, wp, dwp, grid, st);
  }
  prof_end(1, st);
  if (rc) return rc;
  DVD_LAUNCH_CHECK();
  return 0;
}

}  // namespace dvd
