// Attention on the 5th-generation tensor cores without ever materialising the N x N map
// (Discriminators.py:108-114, Attention.py:165-176):   S = Q^T K,  A = softmax_rows(S),  O = V A^T   and its backward.
//
// One kernel template does every contraction of the forward and the backward pass.  A CTA owns 128 "rows" (tokens of
// one side) of one batch item and streams the other side's tokens ("cols") in tiles of 64:
//   stage 1   S[128 x 64]  = X1 . Y1^T          (K = the q/k channel count, padded to 64)
//            [dP[128 x 64] = X2 . Y2^T]          (NS = 2: K = the v channel count)        -> fp32 in TMEM
//   elementwise (one thread per row, TMEM -> registers):
//             e = exp(s - L)  [* (dp - D)]       L = log-sum-exp, D = rowsum(dO * O) of the QUERY of the pair, indexed
//                                                by the row (COL = false) or by the column (COL = true)
//             -> bf16 (hi, lo) planes of E in shared memory, written in the 128B-swizzled K-major layout UMMA expects
//   stage 2   OUT[128 x nz] += E . Z^T            (K = the 64 streamed tokens)              -> fp32 in TMEM
//   epilogue  OUT -> global (channel-major, coalesced over the 128 rows)
//
//   pass            rows     cols     X1 | X2        Y1 | Y2        L, D by   Z            OUT
//   LSE (forward)   queries  keys     q              k              -         -            L[query] = log sum_j exp(S_ij)
//   O   (forward)   queries  keys     q              k              row       v            O (dv x Nq)
//   dV              keys     queries  k              q              col       dO           dV (dv x Nk)   = P^T dO
//   dQ              queries  keys     q | dO         k | v          row       k            dQ (dq x Nq)   = dS K
//   dK              keys     queries  k | v          q | dO         col       q            dK (dq x Nk)   = dS^T Q
//
// Precision: operands are fp32 tensors split into bf16 planes.  The logits feed an exponential, so Q and K are split
// into THREE planes (x = h + m + l, 24 bits) and S sums the six products down to 2^-24; everything else uses the
// engine's usual two planes and three products (2^-16).  exp() is expf on fp32 registers.  The softmax is two-pass
// (L first, then exp(s - L) directly), which is what the reference's nn.Softmax computes and needs no rescaling of the
// accumulator; the backward recomputes P from L the same way, so "attn" and "dattn" never exist in memory.
#include <cuda.h>
#include <cuda_bf16.h>

#include <mutex>

#include "common.cuh"

namespace dvd {
namespace attn {

constexpr int BR = 128;      // rows per CTA
constexpr int BC = 64;       // streamed tokens per tile
constexpr int KP = 64;       // q/k channels, padded

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy writes (st.shared) -> visible to the async proxy (tcgen05.mma reads its operands through it)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// K-major SWIZZLE_128B operand: rows of 128 B, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(16 >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16: D = f32, A = B = bf16, K-major, M = 128, N = n
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BR >> 4) << 24);
}
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
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

// ------------------------------------------------------------------------------------------------ operand planes
// token-major planes: src (B, C, N) fp32 -> NP bf16 planes [B][N][CP] (x = sum of the planes), channels >= C zero.
// Block = 32 tokens x 64 channels, transposed through shared memory.
template <int NP>
__global__ void __launch_bounds__(256) split_tm_kernel(const float* __restrict__ src, int64_t bs, int C, int N, int CP,
                                                       __nv_bfloat16* __restrict__ dst, int64_t plane_stride) {
  __shared__ float tile[64][33];
  const int b = blockIdx.z, n0 = blockIdx.x * 32, c0 = blockIdx.y * 64, tid = threadIdx.x;
  {
    const int tn = tid & 31, n = n0 + tn;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = (tid >> 5) + 8 * j;
      tile[c][tn] = (n < N && c0 + c < C) ? __ldg(src + b * bs + (int64_t)(c0 + c) * N + n) : 0.f;
    }
  }
  __syncthreads();
  const int tn = tid >> 3, q = tid & 7, n = n0 + tn;
  if (n >= N) return;
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = tile[q * 8 + i][tn];
  const int64_t o = ((int64_t)b * N + n) * CP + c0 + q * 8;
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      const float2 hf = __bfloat1622float2(h);
      v[2 * i] -= hf.x;              // the remainder goes to the next plane
      v[2 * i + 1] -= hf.y;
      w[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(dst + p * plane_stride + o) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}
// channel-major planes: src (B, C, N) fp32 -> 2 bf16 planes [B][RP][N], rows >= C zero
__global__ void split_cm_kernel(const float* __restrict__ src, int64_t bs, int C, int N, int RP,
                                __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i % N);
    const int64_t t = i / N;
    const int r = (int)(t % RP), b = (int)(t / RP);
    const float x = r < C ? __ldg(src + b * bs + (int64_t)r * N + n) : 0.f;
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    hi[i] = h;
    lo[i] = __float2bfloat16_rn(x - __bfloat162float(h));
  }
}
// D[b][i] = sum_c dO[b][c][i] * O[b][c][i]
__global__ void rowdot_kernel(const float* __restrict__ a, int64_t a_bs, const float* __restrict__ o, int64_t o_bs,
                              int C, int N, float* __restrict__ d) {
  const int b = blockIdx.y, n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float s = 0.f;
  for (int c = 0; c < C; ++c) s = fmaf(__ldg(a + b * a_bs + (int64_t)c * N + n), __ldg(o + b * o_bs + (int64_t)c * N + n), s);
  d[(int64_t)b * N + n] = s;
}

// ------------------------------------------------------------------------------------------------ the kernel
struct AttnP {
  int rows, cols;          // tokens on the CTA side / streamed side
  int nz;                  // width of OUT (multiple of 16, <= 256)
  int k2;                  // 64-wide K blocks of the second stage-1 GEMM (NS = 2)
  int kq;                  // UMMA_K = 16 steps of the S GEMM that hold data: ceil(dq / 16) (the planes are zero beyond dq)
  int out_z;               // valid columns of OUT
  int ld;                  // queries per batch item (stride of L / D)
  const float* L;
  const float* D;
  float* Lout;             // LSE pass: [B][rows]
  float* out;              // element (b, z, row) at b*out_bs + z*out_zs + row
  int64_t out_bs, out_zs;
  // shared-memory carve-up (bytes from the 1024-aligned base)
  uint32_t off_x2, off_y1, off_y2, off_z, off_e, off_bar;
};

struct Maps {
  CUtensorMap x1[3], y1[3], x2[2], y2[2], z[2];
};

// NS: stage-1 GEMMs (1: S, 2: S and dP).  COL: L / D are indexed by the column.  LSE: only the log-sum-exp of each row.
template <int NS, bool COL, bool LSE>
__global__ void __launch_bounds__(192, 1) attn_tc_kernel(const __grid_constant__ Maps mp, const AttnP ap) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ap.off_bar);
  // x_full | y_full | y_empty | s_full | s_free | e_full | out_full | tmem slot
  const uint32_t b_xfull = smem_u32(bars + 0), b_yfull = smem_u32(bars + 1), b_yempty = smem_u32(bars + 2),
                 b_sfull = smem_u32(bars + 3), b_sfree = smem_u32(bars + 4), b_efull = smem_u32(bars + 5),
                 b_ofull = smem_u32(bars + 6);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * BR, b = blockIdx.y;
  const int n_tiles = (ap.cols + BC - 1) / BC;
  constexpr uint32_t TMEM_COLS = 512;
  constexpr uint32_t COL_S = 0, COL_DP = 64, COL_OUT = 128;

  if (tid == 0) {
    mbar_init(b_xfull, 1); mbar_init(b_yfull, 1); mbar_init(b_yempty, 1); mbar_init(b_sfull, 1);
    mbar_init(b_sfree, 128); mbar_init(b_efull, 128); mbar_init(b_ofull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const uint32_t s_x1 = smem_u32(smem), s_x2 = smem_u32(smem + ap.off_x2), s_y1 = smem_u32(smem + ap.off_y1),
                 s_y2 = smem_u32(smem + ap.off_y2), s_z = smem_u32(smem + ap.off_z), s_e = smem_u32(smem + ap.off_e);
  constexpr uint32_t XB = BR * 128, YB = BC * 128;       // bytes of one 64-channel block of the row / column operands
  const uint32_t zb = (uint32_t)ap.nz * 128;             // bytes of one Z plane

  if (warp == 0) {
    if (lane == 0) {
      // the CTA's own rows, once
      mbar_expect_tx(b_xfull, 3 * XB + (NS == 2 ? 2 * ap.k2 * XB : 0));
#pragma unroll
      for (int p = 0; p < 3; ++p) tma_load_3d(s_x1 + p * XB, &mp.x1[p], b_xfull, 0, row0, b);
      if (NS == 2)
        for (int p = 0; p < 2; ++p)
          for (int kb = 0; kb < ap.k2; ++kb)
            tma_load_3d(s_x2 + (p * ap.k2 + kb) * XB, &mp.x2[p], b_xfull, kb * 64, row0, b);
      for (int t = 0; t < n_tiles; ++t) {
        const int col0 = t * BC;
        mbar_wait(b_yempty, (uint32_t)((t & 1) ^ 1));
        mbar_expect_tx(b_yfull, 3 * YB + (NS == 2 ? 2 * ap.k2 * YB : 0) + (LSE ? 0 : 2 * zb));
#pragma unroll
        for (int p = 0; p < 3; ++p) tma_load_3d(s_y1 + p * YB, &mp.y1[p], b_yfull, 0, col0, b);
        if (NS == 2)
          for (int p = 0; p < 2; ++p)
            for (int kb = 0; kb < ap.k2; ++kb)
              tma_load_3d(s_y2 + (p * ap.k2 + kb) * YB, &mp.y2[p], b_yfull, kb * 64, col0, b);
        if (!LSE)
          for (int p = 0; p < 2; ++p) tma_load_3d(s_z + p * zb, &mp.z[p], b_yfull, col0, 0, b);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t id_s = make_idesc(BC), id_o = make_idesc(ap.nz);
      mbar_wait(b_xfull, 0);
      for (int t = 0; t < n_tiles; ++t) {
        mbar_wait(b_yfull, (uint32_t)(t & 1));
        if (t > 0) mbar_wait(b_sfree, (uint32_t)((t - 1) & 1));
        tc_fence_after();
        // S = X1 . Y1^T with three planes each: the six products above 2^-24 (smallest first)
        {
          const int pa[6] = {1, 2, 0, 1, 0, 0}, pb[6] = {1, 0, 2, 0, 1, 0};     // (m,m) (l,h) (h,l) (m,h) (h,m) (h,h)
#pragma unroll
          for (int i = 0; i < 6; ++i) {
            const uint64_t da = make_desc(s_x1 + pa[i] * XB), db = make_desc(s_y1 + pb[i] * YB);
            for (int kk = 0; kk < ap.kq; ++kk)
              mma_f16(tmem_base + COL_S, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), id_s, (i | kk) != 0);
          }
        }
        if (NS == 2) {      // dP = X2 . Y2^T, two planes: (l,h) (h,l) (h,h)
          const int pa[3] = {1, 0, 0}, pb[3] = {0, 1, 0};
          for (int i = 0; i < 3; ++i)
            for (int kb = 0; kb < ap.k2; ++kb) {
              const uint64_t da = make_desc(s_x2 + (pa[i] * ap.k2 + kb) * XB), db = make_desc(s_y2 + (pb[i] * ap.k2 + kb) * YB);
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                mma_f16(tmem_base + COL_DP, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), id_s, (i | kb | kk) != 0);
            }
        }
        mma_commit(b_sfull);
        if (LSE) {
          mma_commit(b_yempty);
        } else {
          mbar_wait(b_efull, (uint32_t)(t & 1));
          tc_fence_after();
          // OUT += E . Z^T over the 64 streamed tokens: (l,h) (h,l) (h,h)
          const int pa[3] = {1, 0, 0}, pb[3] = {0, 1, 0};
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const uint64_t da = make_desc(s_e + pa[i] * XB), db = make_desc(s_z + pb[i] * zb);
#pragma unroll
            for (int kk = 0; kk < BC / 16; ++kk)
              mma_f16(tmem_base + COL_OUT, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), id_o, (t | i | kk) != 0);
          }
          mma_commit(b_yempty);
        }
      }
      mma_commit(b_ofull);
    }
    __syncwarp();
  } else {
    // ---------------- elementwise + epilogue: one thread per row; warp w may touch TMEM lanes [32*(w%4), +32)
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int row = row0 + r;
    const bool ok = row < ap.rows;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    const float* Lb = ap.L ? ap.L + (int64_t)b * ap.ld : nullptr;
    const float* Db = ap.D ? ap.D + (int64_t)b * ap.ld : nullptr;
    float Lr = 0.f, Dr = 0.f;
    if (!LSE && !COL && ok) { Lr = __ldg(Lb + row); if (NS == 2) Dr = __ldg(Db + row); }
    float run_m = -INFINITY, run_l = 0.f;
    for (int t = 0; t < n_tiles; ++t) {
      const int col0 = t * BC;
      mbar_wait(b_sfull, (uint32_t)(t & 1));
      tc_fence_after();
      float e[BC];
      {
        uint32_t r0[32], r1[32];
        tmem_ld32(taddr + COL_S, r0);
        tmem_ld32(taddr + COL_S + 32, r1);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) { e[j] = __uint_as_float(r0[j]); e[32 + j] = __uint_as_float(r1[j]); }
      }
      if (LSE) {
        tc_fence_before();
        mbar_arrive(b_sfree);
        float tm = -INFINITY;
#pragma unroll
        for (int j = 0; j < BC; ++j) if (col0 + j < ap.cols) tm = fmaxf(tm, e[j]);
        const float m_new = fmaxf(run_m, tm);
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < BC; ++j) if (col0 + j < ap.cols) s += expf(e[j] - m_new);
        run_l = run_l * expf(run_m - m_new) + s;
        run_m = m_new;
        continue;
      }
      if (NS == 2) {
        uint32_t r0[32], r1[32];
        tmem_ld32(taddr + COL_DP, r0);
        tmem_ld32(taddr + COL_DP + 32, r1);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(b_sfree);
#pragma unroll
        for (int j = 0; j < BC; ++j) {
          const int c = col0 + j;
          const float dp = __uint_as_float(j < 32 ? r0[j] : r1[j - 32]);
          float v = 0.f;
          if (c < ap.cols) {
            const float l = COL ? __ldg(Lb + c) : Lr, dd = COL ? __ldg(Db + c) : Dr;
            v = expf(e[j] - l) * (dp - dd);
          }
          e[j] = v;
        }
      } else {
        tc_fence_before();
        mbar_arrive(b_sfree);
#pragma unroll
        for (int j = 0; j < BC; ++j) {
          const int c = col0 + j;
          e[j] = c < ap.cols ? expf(e[j] - (COL ? __ldg(Lb + c) : Lr)) : 0.f;
        }
      }
      // E -> bf16 (hi, lo) planes, K-major rows of 128 B, 16-byte chunk j of row r at position j ^ (r & 7)
      uint8_t* e_hi = smem + ap.off_e + r * 128;
      uint8_t* e_lo = e_hi + XB;
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        uint32_t h[4], l[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float a = e[ch * 8 + 2 * i], c = e[ch * 8 + 2 * i + 1];
          const __nv_bfloat162 hp = __floats2bfloat162_rn(a, c);
          const float2 hf = __bfloat1622float2(hp);
          const __nv_bfloat162 lp = __floats2bfloat162_rn(a - hf.x, c - hf.y);
          h[i] = *reinterpret_cast<const uint32_t*>(&hp);
          l[i] = *reinterpret_cast<const uint32_t*>(&lp);
        }
        const int pos = (ch ^ (r & 7)) << 4;
        *reinterpret_cast<uint4*>(e_hi + pos) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(e_lo + pos) = make_uint4(l[0], l[1], l[2], l[3]);
      }
      fence_proxy_async();
      mbar_arrive(b_efull);
    }
    if (LSE) {
      if (ok) ap.Lout[(int64_t)b * ap.rows + row] = run_m + logf(run_l);
    } else {
      mbar_wait(b_ofull, 0);
      tc_fence_after();
      float* dst = ap.out + (int64_t)b * ap.out_bs + row;
      for (int cb = 0; cb < ap.nz; cb += 32) {
        if (cb >= ap.out_z) break;
        uint32_t r0[32];
        if (ap.nz - cb >= 32) {
          tmem_ld32(taddr + COL_OUT + cb, r0);
        } else {        // nz is a multiple of 16: a 16-column tail
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
              "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
              : "=r"(r0[0]), "=r"(r0[1]), "=r"(r0[2]), "=r"(r0[3]), "=r"(r0[4]), "=r"(r0[5]), "=r"(r0[6]), "=r"(r0[7]),
                "=r"(r0[8]), "=r"(r0[9]), "=r"(r0[10]), "=r"(r0[11]), "=r"(r0[12]), "=r"(r0[13]), "=r"(r0[14]), "=r"(r0[15])
              : "r"(taddr + COL_OUT + cb));
        }
        tmem_ld_wait();
        if (!ok) continue;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (cb + j < ap.out_z) dst[(int64_t)(cb + j) * ap.out_zs] = __uint_as_float(r0[j]);
      }
      tc_fence_before();
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeFn get_encode() {
  static EncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(ptr);
  });
  return fn;
}
// planes [B][R][Cc] (Cc contiguous) -> 3-D map (c, r, b), box (64, box_r, 1), 128B swizzle, zero fill out of bounds
static int make_map(CUtensorMap* tm, const void* base, int B, int R, int Cc, int box_r) {
  EncodeFn enc = get_encode();
  if (!enc) return fail("cuTensorMapEncodeTiled unavailable%s (%s:%d)", "", __FILE__, __LINE__);
  cuuint64_t gdim[3] = {(cuuint64_t)Cc, (cuuint64_t)R, (cuuint64_t)B};
  cuuint64_t gstr[2] = {(cuuint64_t)Cc * 2, (cuuint64_t)R * Cc * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)box_r, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled (attention) failed%s (%s:%d)", "", __FILE__, __LINE__);
  return 0;
}

static inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

struct Planes {          // one split tensor: np planes of `elems` bf16 each
  __nv_bfloat16* p = nullptr;
  size_t elems = 0;
  __nv_bfloat16* plane(int i) const { return p + i * elems; }
};

struct Carver {
  char* base;
  size_t off = 0;
  void* take(size_t bytes) {
    void* r = base ? base + off : nullptr;
    off += (bytes + 255) / 256 * 256;
    return r;
  }
  Planes planes(int np, size_t elems) {
    Planes pl;
    pl.elems = (elems + 127) / 128 * 128;
    pl.p = reinterpret_cast<__nv_bfloat16*>(take(sizeof(__nv_bfloat16) * np * pl.elems));
    return pl;
  }
};

struct Ws {
  Planes q_tm, k_tm, v_cm;                       // forward
  Planes v_tm, do_tm, do_cm, k_cm, q_cm;         // backward
  float* D;
  size_t bytes;
};
static Ws carve(void* base, int B, int dq, int dv, int Nq, int Nk, bool bwd) {
  Carver c{reinterpret_cast<char*>(base)};
  Ws w;
  const int dvP = round_up(dv, 64), zv = round_up(dv, 16), zq = round_up(dq, 16);
  w.q_tm = c.planes(3, (size_t)B * Nq * KP);
  w.k_tm = c.planes(3, (size_t)B * Nk * KP);
  if (!bwd) {
    w.v_cm = c.planes(2, (size_t)B * zv * Nk);
    w.D = nullptr;
  } else {
    w.v_tm = c.planes(2, (size_t)B * Nk * dvP);
    w.do_tm = c.planes(2, (size_t)B * Nq * dvP);
    w.do_cm = c.planes(2, (size_t)B * zv * Nq);
    w.k_cm = c.planes(2, (size_t)B * zq * Nk);
    w.q_cm = c.planes(2, (size_t)B * zq * Nq);
    w.D = reinterpret_cast<float*>(c.take(sizeof(float) * (size_t)B * Nq));
  }
  w.bytes = c.off;
  return w;
}

static int split_tm(const float* src, int64_t bs, int B, int C, int N, int CP, int np, const Planes& pl, cudaStream_t st) {
  dim3 grid(ceil_div(N, 32), CP / 64, B);
  DVD_CHECK_ARG(B <= 65535);
  if (np == 3) split_tm_kernel<3><<<grid, 256, 0, st>>>(src, bs, C, N, CP, pl.p, (int64_t)pl.elems);
  else split_tm_kernel<2><<<grid, 256, 0, st>>>(src, bs, C, N, CP, pl.p, (int64_t)pl.elems);
  DVD_LAUNCH_CHECK();
  return 0;
}
static int split_cm(const float* src, int64_t bs, int B, int C, int N, int RP, const Planes& pl, cudaStream_t st) {
  const int64_t total = (int64_t)B * RP * N;
  split_cm_kernel<<<ew_blocks(total, 4), 256, 0, st>>>(src, bs, C, N, RP, pl.plane(0), pl.plane(1), total);
  DVD_LAUNCH_CHECK();
  return 0;
}

struct Pass {
  int B, rows, cols;
  const Planes* x1; const Planes* y1;             // 3 planes, [B][rows|cols][64]
  int dq = KP;                                    // channels of x1 / y1 that hold data
  const Planes* x2 = nullptr; const Planes* y2 = nullptr; int k2 = 0;      // 2 planes, [B][rows|cols][64*k2]
  const Planes* z = nullptr; int nz = 0;          // 2 planes, [B][nz][cols]
  const float* L = nullptr; const float* D = nullptr; int ld = 0;
  float* Lout = nullptr;
  float* out = nullptr; int64_t out_bs = 0, out_zs = 0; int out_z = 0;
};

template <int NS, bool COL, bool LSE>
static int launch(const Pass& ps, cudaStream_t st) {
  Maps mp;
  memset(&mp, 0, sizeof(mp));
  for (int p = 0; p < 3; ++p) {
    DVD_TRY(make_map(&mp.x1[p], ps.x1->plane(p), ps.B, ps.rows, KP, BR));
    DVD_TRY(make_map(&mp.y1[p], ps.y1->plane(p), ps.B, ps.cols, KP, BC));
  }
  if (NS == 2)
    for (int p = 0; p < 2; ++p) {
      DVD_TRY(make_map(&mp.x2[p], ps.x2->plane(p), ps.B, ps.rows, 64 * ps.k2, BR));
      DVD_TRY(make_map(&mp.y2[p], ps.y2->plane(p), ps.B, ps.cols, 64 * ps.k2, BC));
    }
  if (!LSE)
    for (int p = 0; p < 2; ++p) DVD_TRY(make_map(&mp.z[p], ps.z->plane(p), ps.B, ps.nz, ps.cols, ps.nz));
  AttnP ap;
  memset(&ap, 0, sizeof(ap));
  ap.rows = ps.rows; ap.cols = ps.cols; ap.nz = LSE ? 16 : ps.nz; ap.k2 = ps.k2; ap.out_z = ps.out_z; ap.ld = ps.ld;
  ap.kq = ceil_div(ps.dq, 16);
  ap.L = ps.L; ap.D = ps.D; ap.Lout = ps.Lout; ap.out = ps.out; ap.out_bs = ps.out_bs; ap.out_zs = ps.out_zs;
  uint32_t off = 3 * BR * 128;
  ap.off_x2 = off; off += NS == 2 ? 2 * ps.k2 * BR * 128 : 0;
  ap.off_y1 = off; off += 3 * BC * 128;
  ap.off_y2 = off; off += NS == 2 ? 2 * ps.k2 * BC * 128 : 0;
  ap.off_z = off; off += LSE ? 0 : (uint32_t)(2 * ps.nz * 128 + 1023) / 1024 * 1024;
  ap.off_e = off; off += LSE ? 0 : 2 * BR * 128;
  ap.off_bar = off; off += 128;
  const int smem = (int)off + 1024;
  DVD_CHECK_ARG(smem <= 227 * 1024);
  // dynamic shared memory differs per shape: raise the kernel's limit to the device maximum once per device
  static std::atomic<uint64_t> configured{0};
  if (!device_bit_test_and_set(configured))
    DVD_CUDA(cudaFuncSetAttribute(attn_tc_kernel<NS, COL, LSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  dim3 grid(ceil_div(ps.rows, BR), ps.B);
  attn_tc_kernel<NS, COL, LSE><<<grid, 192, smem, st>>>(mp, ap);
  DVD_LAUNCH_CHECK();
  return 0;
}

}  // namespace attn
}  // namespace dvd

using namespace dvd;

// Shapes the tensor-core path covers (everything else: dvd_attn_fwd / dvd_attn_bwd, the materialised SIMT path):
// channel-major q / k / v, dq <= 64, dv <= 256 forward / 128 backward, token counts that are multiples of 8 (TMA
// strides of the channel-major planes), batch <= 65535.
extern "C" int dvd_attn_flash_supported(int batch, int dq, int dv, int Nq, int Nk, int backward) {
  if (!get_option(OPT_FLASH_ATTN) || attn::get_encode() == nullptr) return 0;
  if (batch <= 0 || batch > 65535 || dq <= 0 || dq > 64 || dv <= 0) return 0;
  if (dv > (backward ? 128 : 256)) return 0;
  if (Nq % 8 || Nk % 8 || Nq < 8 || Nk < 8) return 0;
  return 1;
}
extern "C" size_t dvd_attn_flash_workspace_bytes(int batch, int dq, int dv, int Nq, int Nk, int backward) {
  return attn::carve(nullptr, batch, dq, dv, Nq, Nk, backward != 0).bytes + 1024;
}

// q (B,dq,Nq), k (B,dq,Nk), v (B,dv,Nk) -> out (B,dv,Nq), lse (B,Nq) (kept for the backward)
extern "C" int dvd_attn_flash_fwd(const float* q, int64_t q_bs, const float* k, int64_t k_bs, const float* v,
                                  int64_t v_bs, float* out, int64_t o_bs, float* lse, int batch, int dq, int dv, int Nq,
                                  int Nk, void* workspace, size_t ws_bytes, void* stream) {
  dvd::ProfScope _ps(3, "attn_flash_fwd", dvd::as_stream(stream));
  DVD_CHECK_ARG(q && k && v && out && lse && workspace);
  DVD_CHECK_ARG(dvd_attn_flash_supported(batch, dq, dv, Nq, Nk, 0));
  cudaStream_t st = as_stream(stream);
  void* base = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
  attn::Ws ws = attn::carve(base, batch, dq, dv, Nq, Nk, false);
  DVD_CHECK_ARG(ws.bytes + 1024 <= ws_bytes);
  const int zv = attn::round_up(dv, 16);
  DVD_TRY(attn::split_tm(q, q_bs, batch, dq, Nq, attn::KP, 3, ws.q_tm, st));
  DVD_TRY(attn::split_tm(k, k_bs, batch, dq, Nk, attn::KP, 3, ws.k_tm, st));
  DVD_TRY(attn::split_cm(v, v_bs, batch, dv, Nk, zv, ws.v_cm, st));
  attn::Pass ps;
  ps.B = batch; ps.rows = Nq; ps.cols = Nk; ps.x1 = &ws.q_tm; ps.y1 = &ws.k_tm; ps.Lout = lse; ps.dq = dq;
  DVD_TRY((attn::launch<1, false, true>(ps, st)));
  ps.Lout = nullptr; ps.L = lse; ps.ld = Nq; ps.z = &ws.v_cm; ps.nz = zv; ps.out = out; ps.out_bs = o_bs; ps.out_zs = Nq;
  ps.out_z = dv;
  DVD_TRY((attn::launch<1, false, false>(ps, st)));
  return 0;
}

// gradients of the above: dq_ (B,dq,Nq), dk_ (B,dq,Nk), dv_ (B,dv,Nk) from dout (B,dv,Nq), the saved out and lse
extern "C" int dvd_attn_flash_bwd(const float* q, int64_t q_bs, const float* k, int64_t k_bs, const float* v,
                                  int64_t v_bs, const float* out, int64_t o_bs, const float* dout, int64_t do_bs,
                                  const float* lse, float* dq_, int64_t dq_bs, float* dk_, int64_t dk_bs, float* dv_,
                                  int64_t dv_bs, int batch, int dq, int dv, int Nq, int Nk, void* workspace,
                                  size_t ws_bytes, void* stream) {
  dvd::ProfScope _ps(3, "attn_flash_bwd", dvd::as_stream(stream));
  DVD_CHECK_ARG(q && k && v && out && dout && lse && dq_ && dk_ && dv_ && workspace);
  DVD_CHECK_ARG(dvd_attn_flash_supported(batch, dq, dv, Nq, Nk, 1));
  cudaStream_t st = as_stream(stream);
  void* base = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
  attn::Ws ws = attn::carve(base, batch, dq, dv, Nq, Nk, true);
  DVD_CHECK_ARG(ws.bytes + 1024 <= ws_bytes);
  const int dvP = attn::round_up(dv, 64), zv = attn::round_up(dv, 16), zq = attn::round_up(dq, 16);
  DVD_TRY(attn::split_tm(q, q_bs, batch, dq, Nq, attn::KP, 3, ws.q_tm, st));
  DVD_TRY(attn::split_tm(k, k_bs, batch, dq, Nk, attn::KP, 3, ws.k_tm, st));
  DVD_TRY(attn::split_tm(v, v_bs, batch, dv, Nk, dvP, 2, ws.v_tm, st));
  DVD_TRY(attn::split_tm(dout, do_bs, batch, dv, Nq, dvP, 2, ws.do_tm, st));
  DVD_TRY(attn::split_cm(dout, do_bs, batch, dv, Nq, zv, ws.do_cm, st));
  DVD_TRY(attn::split_cm(k, k_bs, batch, dq, Nk, zq, ws.k_cm, st));
  DVD_TRY(attn::split_cm(q, q_bs, batch, dq, Nq, zq, ws.q_cm, st));
  {
    dim3 grid(ceil_div(Nq, 256), batch);
    attn::rowdot_kernel<<<grid, 256, 0, st>>>(dout, do_bs, out, o_bs, dv, Nq, ws.D);
    DVD_LAUNCH_CHECK();
  }
  // dV = P^T dO: rows = keys, columns = queries
  {
    attn::Pass ps;
    ps.B = batch; ps.rows = Nk; ps.cols = Nq; ps.x1 = &ws.k_tm; ps.y1 = &ws.q_tm; ps.L = lse; ps.ld = Nq; ps.dq = dq;
    ps.z = &ws.do_cm; ps.nz = zv; ps.out = dv_; ps.out_bs = dv_bs; ps.out_zs = Nk; ps.out_z = dv;
    DVD_TRY((attn::launch<1, true, false>(ps, st)));
  }
  // dQ = dS K: rows = queries, columns = keys
  {
    attn::Pass ps;
    ps.B = batch; ps.rows = Nq; ps.cols = Nk; ps.x1 = &ws.q_tm; ps.y1 = &ws.k_tm; ps.x2 = &ws.do_tm; ps.y2 = &ws.v_tm;
    ps.k2 = dvP / 64; ps.L = lse; ps.D = ws.D; ps.ld = Nq; ps.dq = dq;
    ps.z = &ws.k_cm; ps.nz = zq; ps.out = dq_; ps.out_bs = dq_bs; ps.out_zs = Nq; ps.out_z = dq;
    DVD_TRY((attn::launch<2, false, false>(ps, st)));
  }
  // dK = dS^T Q: rows = keys, columns = queries
  {
    attn::Pass ps;
    ps.B = batch; ps.rows = Nk; ps.cols = Nq; ps.x1 = &ws.k_tm; ps.y1 = &ws.q_tm; ps.x2 = &ws.v_tm; ps.y2 = &ws.do_tm;
    ps.k2 = dvP / 64; ps.L = lse; ps.D = ws.D; ps.ld = Nq; ps.dq = dq;
    ps.z = &ws.q_cm; ps.nz = zq; ps.out = dk_; ps.out_bs = dk_bs; ps.out_zs = Nk; ps.out_z = dq;
    DVD_TRY((attn::launch<2, true, false>(ps, st)));
  }
  return 0;
}
