// tcgen05 (5th-gen tensor core) implicit-GEMM convolution for sm_100a: forward / dgrad and weight gradient.
//
// Precision: fp32 operands are split on the fly into bf16 hi + lo (x = hi + lo to ~2^-17) and every k-block issues
// three MMAs into one fp32 TMEM accumulator:  D += A_hi*B_hi + A_lo*B_hi + A_hi*B_lo  ("BF16x3").  Single-pass
// TF32/BF16 miss the 1e-3 output bound of the 48-frame recurrence; the 3-pass split is ~1e-5.
//
// Structure of one CTA (288 threads, one 128 x BN output tile, one accumulator in TMEM):
//   warps 0-3  A producers: thread t owns tile row t.  They gather the row's 64 k-values of the current k-block
//              straight from the fp32 NC(D)HW tensors (coalesced across the warp: lanes = consecutive rows),
//              split to bf16 hi/lo and write K-major, 128B-swizzled operand tiles into shared memory.
//              After the main loop the same warps are the epilogue (row t == TMEM lane t).
//   warps 4-7  B producers, same scheme for the BN rows of the B operand.
//   warp  8    TMEM allocator + single-thread tcgen05.mma issuer; tcgen05.commit frees smem stages / signals
//              the epilogue through mbarriers.
// forward:  rows = output pixels, k = (tap, ci), B rows = cout          (A gathered with the tap's shift)
// wgrad:    rows = ci,            k = pixels,    B rows = cout (from dY) (one tap and one pixel range per CTA)
#include <cuda_bf16.h>

#include "common.cuh"
#include "conv_params.cuh"

namespace dvd {

namespace tc {

constexpr int BM = 128;        // tile rows (UMMA M)
constexpr int BKC = 64;        // k elements per stage: 64 bf16 = one 128-byte swizzle row
constexpr int NT = 288;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
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
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (rows at 128 B pitch, 8-row groups 1024 B apart)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);      // start address, 16-byte units
  d |= (uint64_t)1 << 16;                       // leading byte offset (ignored for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;             // stride byte offset: 8 rows * 128 B
  d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=n
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
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
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// split 8 fp32 values into bf16 hi / lo and store the two 16-byte chunks of tile row `row`, k-chunk `q`
__device__ __forceinline__ void store_chunk(uint8_t* hi_tile, uint8_t* lo_tile, int row, int q, const float* v) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * i]), h1 = __float2bfloat16_rn(v[2 * i + 1]);
    const __nv_bfloat16 l0 = __float2bfloat16_rn(v[2 * i] - __bfloat162float(h0));
    const __nv_bfloat16 l1 = __float2bfloat16_rn(v[2 * i + 1] - __bfloat162float(h1));
    h[i] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    l[i] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
  }
  const int off = row * 128 + ((q ^ (row & 7)) << 4);
  *reinterpret_cast<uint4*>(hi_tile + off) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(lo_tile + off) = make_uint4(l[0], l[1], l[2], l[3]);
}

template <int BN>
struct Cfg {
  static constexpr int A_BYTES = BM * 128;            // one bf16 plane of the A tile
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 2 : (BN == 128 ? 3 : 4);
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
};

// ------------------------------------------------------------------------------------------------------------
// MODE 0: forward / dgrad.   MODE 1: weight gradient.
// ------------------------------------------------------------------------------------------------------------
template <int BN, int MODE>
__global__ void __launch_bounds__(NT, 1) conv_tc_kernel(const ConvP p, float* __restrict__ dwp, int nsplit,
                                                        int per_split) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* full_bar = bars;                        // [STAGES] producers -> MMA
  uint64_t* empty_bar = bars + C::STAGES;           // [STAGES] MMA -> producers
  uint64_t* accum_bar = bars + 2 * C::STAGES;       // MMA -> epilogue
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 1);

  const dvd_conv_desc& d = p.d;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(smem_u32(full_bar + s), 256);
      mbar_init(smem_u32(empty_bar + s), 1);
    }
    mbar_init(smem_u32(accum_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)C::TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ---- iteration space of this CTA
  int it_begin, it_end;            // k-block range
  int row0, n0;                    // first tile row / first B row
  int tap_w = 0, m_lo = 0, m_hi = 0;
  if (MODE == 0) {
    row0 = blockIdx.x * BM;
    n0 = blockIdx.y * BN;
    it_begin = blockIdx.z * per_split;
    it_end = min(it_begin + per_split, p.iters_total);
  } else {
    row0 = blockIdx.x * BM;        // ci
    n0 = blockIdx.y * BN;          // co
    tap_w = blockIdx.z / nsplit;
    const int split = blockIdx.z - tap_w * nsplit;
    m_lo = split * per_split;      // pixel range (multiple of 64)
    m_hi = min(m_lo + per_split, p.M);
    it_begin = 0;
    it_end = (m_hi - m_lo + BKC - 1) / BKC;
  }
  const int n_iters = it_end - it_begin;

  if (warp < 8) {
    // =============================================================== producers
    const bool is_a = warp < 4;
    const int t = tid & 127;                       // row within the A tile / first row within the B tile
    constexpr int ROWS_PER_THREAD = BN > 128 ? BN / 128 : 1;
    int stage = 0;
    uint32_t phase = 0;

    if (MODE == 0) {
      // ---------------- forward: A rows = pixels (gather with tap shift), B rows = cout (packed fp32 weights)
      const int m = row0 + t;
      const bool m_valid = is_a && m < p.M;
      int pz = 0, py = 0, px = 0;
      int64_t x_base = 0;
      if (m_valid) {
        const int n = m / p.DHW;
        int rem = m - n * p.DHW;
        pz = rem / p.HW;
        rem -= pz * p.HW;
        py = rem / d.W;
        px = rem - py * d.W;
        const int n1 = n / d.N2, n2 = n - n1 * d.N2;
        x_base = (int64_t)n1 * d.x_s1 + (int64_t)n2 * d.x_s2;
      }
      int tap = it_begin / p.ck;
      int cchunk = it_begin - tap * p.ck;
      bool a_valid = false;
      int64_t a_off = 0;
      auto set_tap = [&](int tp) {
        const int kw = tp % d.kW;
        const int t2 = tp / d.kW;
        const int kh = t2 % d.kH;
        const int kd = t2 / d.kH;
        const int iz = pz + kd - d.kD / 2, iy = py + kh - d.kH / 2, ix = px + kw - d.kW / 2;
        a_valid = m_valid && iz >= 0 && iz < d.D && iy >= 0 && iy < d.H && ix >= 0 && ix < d.W;
        a_off = x_base + (int64_t)iz * p.Hs * p.Ws + (int64_t)(iy >> d.in_up) * p.Ws + (ix >> d.in_up);
      };
      if (is_a) set_tap(tap);
      for (int it = 0; it < n_iters; ++it) {
        const int c0 = cchunk * BKC;
        mbar_wait(smem_u32(empty_bar + stage), phase ^ 1);
        uint8_t* st = smem + stage * C::STAGE_BYTES;
        if (is_a) {
          const float* src = p.x + a_off + (int64_t)c0 * d.x_cs;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int c = c0 + half * 32 + j;
              float val = 0.f;
              if (a_valid && c < d.Cin) {
                val = __ldg(src + (int64_t)(half * 32 + j) * d.x_cs);
                if (d.in_relu) val = fmaxf(val, 0.f);
              }
              v[j] = val;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) store_chunk(st, st + C::A_BYTES, t, half * 4 + q, v + 8 * q);
          }
        } else {
          const float* wt = p.w + ((int64_t)tap * d.Cin + c0) * d.Cout;
#pragma unroll
          for (int rr = 0; rr < ROWS_PER_THREAD; ++rr) {
            const int row = t + rr * 128;
            if (row < BN) {
              const int co = n0 + row;
              const bool rv = co < d.Cout;
#pragma unroll
              for (int half = 0; half < 2; ++half) {
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  const int c = c0 + half * 32 + j;
                  v[j] = (rv && c < d.Cin) ? __ldg(wt + (int64_t)(half * 32 + j) * d.Cout + co) : 0.f;
                }
#pragma unroll
                for (int q = 0; q < 4; ++q)
                  store_chunk(st + 2 * C::A_BYTES, st + 2 * C::A_BYTES + C::B_BYTES, row, half * 4 + q, v + 8 * q);
              }
            }
          }
        }
        fence_async_smem();
        mbar_arrive(smem_u32(full_bar + stage));
        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        if (++cchunk == p.ck) {
          cchunk = 0;
          ++tap;
          if (is_a && it + 1 < n_iters) set_tap(tap);
        }
      }
    } else {
      // ---------------- wgrad: A rows = ci (x shifted by the tap), B rows = co (dY); k = 64 consecutive pixels.
      // Lanes walk along pixels (contiguous in NC(D)HW); each warp covers 32 rows per pass.
      const int kw = tap_w % d.kW;
      const int t2 = tap_w / d.kW;
      const int kh = t2 % d.kH;
      const int kd = t2 / d.kH;
      const int oz = kd - d.kD / 2, oy = kh - d.kH / 2, ox = kw - d.kW / 2;
      const int wq = warp & 3;                      // which quarter of the rows
      constexpr int NROWS = 128;                    // rows handled per producer group pass
      for (int it = 0; it < n_iters; ++it) {
        const int mk = m_lo + it * BKC;
        mbar_wait(smem_u32(empty_bar + stage), phase ^ 1);
        uint8_t* st = smem + stage * C::STAGE_BYTES;
        // decode this lane's two pixels (k positions lane and lane + 32)
        int64_t off[2];
        bool ok[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int m = mk + h * 32 + lane;
          ok[h] = false;
          off[h] = 0;
          if (m < m_hi) {
            const int n = m / p.DHW;
            const int rem = m - n * p.DHW;
            const int n1 = n / d.N2, n2 = n - n1 * d.N2;
            if (is_a) {
              const int pz = rem / p.HW;
              const int r2 = rem - pz * p.HW;
              const int py = r2 / d.W;
              const int px = r2 - py * d.W;
              const int iz = pz + oz, iy = py + oy, ix = px + ox;
              ok[h] = iz >= 0 && iz < d.D && iy >= 0 && iy < d.H && ix >= 0 && ix < d.W;
              off[h] = (int64_t)n1 * d.x_s1 + (int64_t)n2 * d.x_s2 + (int64_t)iz * p.Hs * p.Ws +
                       (int64_t)(iy >> d.in_up) * p.Ws + (ix >> d.in_up);
            } else {
              ok[h] = true;
              off[h] = (int64_t)n1 * d.y_s1 + (int64_t)n2 * d.y_s2 + rem;
            }
          }
        }
        const float* base = is_a ? p.x : p.y;
        const int64_t cs = is_a ? d.x_cs : d.y_cs;
        const int rlimit = is_a ? d.Cin : d.Cout;
        const int rbase = is_a ? row0 : n0;
        const int nrows = is_a ? BM : BN;
        uint8_t* hi_tile = st + (is_a ? 0 : 2 * C::A_BYTES);
        uint8_t* lo_tile = hi_tile + (is_a ? C::A_BYTES : C::B_BYTES);
        for (int r0 = wq * 32; r0 < nrows; r0 += NROWS) {
#pragma unroll 4
          for (int rr = 0; rr < 32; ++rr) {
            const int row = r0 + rr;
            const int ch = rbase + row;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              float val = 0.f;
              if (ok[h] && ch < rlimit) {
                val = __ldg(base + off[h] + (int64_t)ch * cs);
                if (is_a && d.in_relu) val = fmaxf(val, 0.f);
              }
              const __nv_bfloat16 hv = __float2bfloat16_rn(val);
              const __nv_bfloat16 lv = __float2bfloat16_rn(val - __bfloat162float(hv));
              const int kpos = h * 32 + lane;                      // k index within the stage (0..63)
              const int o = row * 128 + (((kpos >> 3) ^ (row & 7)) << 4) + (kpos & 7) * 2;
              *reinterpret_cast<__nv_bfloat16*>(hi_tile + o) = hv;
              *reinterpret_cast<__nv_bfloat16*>(lo_tile + o) = lv;
            }
          }
        }
        fence_async_smem();
        mbar_arrive(smem_u32(full_bar + stage));
        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
      }
    }

    // =============================================================== epilogue (warps 0-3: row t == TMEM lane t)
    if (is_a) {
      mbar_wait(smem_u32(accum_bar), 0);
      tc_fence_after();
      const uint32_t taddr_row = tmem_base + ((uint32_t)(warp * 32) << 16);
      if (MODE == 0) {
        const int m = row0 + t;
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
        for (int cb = 0; cb < BN; cb += 32) {
          if (n0 + cb >= d.Cout) break;                              // warp-uniform
          uint32_t r[32];
          tmem_ld32(taddr_row + cb, r);
          if (!ok) continue;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int co = n0 + cb + j;
            if (co >= d.Cout) break;
            float v = __uint_as_float(r[j]);
            if (lead) {
              if (p.bias) v += __ldg(p.bias + co);
              if (p.res) v += __ldg(p.res + ro + (int64_t)co * d.r_cs);
            }
            float* dst = p.y + yo + (int64_t)co * d.y_cs;
            if (p.atomic_out) {
              atomicAdd(dst, v);
            } else {
              if (d.accumulate) v += *dst;
              if (d.out_act == 1) v = fmaxf(v, 0.f);
              else if (d.out_act == 2) v = tanhf(v);
              *dst = v;
            }
          }
        }
      } else {
        const int ci = row0 + t;
        const bool ok = ci < d.Cin;
        for (int cb = 0; cb < BN; cb += 32) {
          if (n0 + cb >= d.Cout) break;
          uint32_t r[32];
          tmem_ld32(taddr_row + cb, r);
          if (!ok) continue;
          float* dst = dwp + ((int64_t)tap_w * d.Cin + ci) * d.Cout + n0 + cb;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (n0 + cb + j >= d.Cout) break;
            const float v = __uint_as_float(r[j]);
            if (p.atomic_out) atomicAdd(dst + j, v);
            else dst[j] = v;
          }
        }
      }
      tc_fence_before();
    }
  } else {
    // =============================================================== MMA issuer (warp 8, one elected lane)
    if (lane == 0) {
      const uint32_t idesc = make_idesc(BN);
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < n_iters; ++it) {
        mbar_wait(smem_u32(full_bar + stage), phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
        const uint64_t a_hi = make_desc(sa), a_lo = make_desc(sa + C::A_BYTES);
        const uint64_t b_hi = make_desc(sa + 2 * C::A_BYTES), b_lo = make_desc(sa + 2 * C::A_BYTES + C::B_BYTES);
#pragma unroll
        for (int kk = 0; kk < BKC / 16; ++kk) {
          const uint64_t adv = (uint64_t)(kk * 32 >> 4);     // 16 bf16 = 32 bytes along K inside the swizzle row
          mma_f16(tmem_base, a_hi + adv, b_hi + adv, idesc, (it | kk) != 0);
          mma_f16(tmem_base, a_lo + adv, b_hi + adv, idesc, 1);
          mma_f16(tmem_base, a_hi + adv, b_lo + adv, idesc, 1);
        }
        mma_commit(smem_u32(empty_bar + stage));             // frees the stage when the MMAs have read it
        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
      }
      mma_commit(smem_u32(accum_bar));
    }
    __syncwarp();
  }

  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS));
  }
}

template <int BN, int MODE>
static int launch(const ConvP& p, float* dwp, dim3 grid, int nsplit, int per_split, cudaStream_t st) {
  using C = Cfg<BN>;
  static bool configured = false;     // per process & template instance; attribute is per-device but cheap to re-set
  if (!configured) {
    DVD_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    configured = true;
  }
  conv_tc_kernel<BN, MODE><<<grid, NT, C::SMEM, st>>>(p, dwp, nsplit, per_split);
  return 0;
}

}  // namespace tc

// which implementation: env DVD_CONV_IMPL = "simt" | "tc" (default tc when eligible)
static int impl_pref() {
  static int pref = -1;
  if (pref < 0) {
    const char* e = getenv("DVD_CONV_IMPL");
    pref = (e && strcmp(e, "simt") == 0) ? 0 : 1;
  }
  return pref;
}

bool tc_fwd_eligible(const ConvP& p) {
  if (!impl_pref()) return false;
  const dvd_conv_desc& d = p.d;
  return d.Cin >= 32 && d.Cout >= 64 && p.M >= 128;
}

int tc_fwd_launch(ConvP& p, cudaStream_t st) {
  const dvd_conv_desc& d = p.d;
  const int nsm = num_sms();
  int bn = d.Cout >= 256 ? 256 : (d.Cout >= 128 ? 128 : 64);
  const int mt = ceil_div(p.M, tc::BM);
  if (bn == 256 && (int64_t)mt * ceil_div(d.Cout, 256) < nsm) bn = 128;     // fill the machine first
  p.ck = ceil_div(d.Cin, tc::BKC);
  p.iters_total = p.taps * p.ck;
  const int64_t ctas = (int64_t)mt * ceil_div(d.Cout, bn);
  int nsplit = 1;
  if (ctas < nsm && d.out_act == 0 && p.iters_total >= 8) {
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
  dim3 grid(mt, ceil_div(d.Cout, bn), nsplit);
  prof_tag("tc1 fwd M%d Ci%d Co%d t%d", p.M, d.Cin, d.Cout, p.taps);
  prof_begin(0, 2.0 * p.M * (double)d.Cout * d.Cin * p.taps, st);
  int rc;
  if (bn == 256) rc = tc::launch<256, 0>(p, nullptr, grid, nsplit, per, st);
  else if (bn == 128) rc = tc::launch<128, 0>(p, nullptr, grid, nsplit, per, st);
  else rc = tc::launch<64, 0>(p, nullptr, grid, nsplit, per, st);
  prof_end(0, st);
  if (rc) return rc;
  DVD_LAUNCH_CHECK();
  return 0;
}

bool tc_wgrad_eligible(const ConvP& p) {
  if (!impl_pref()) return false;
  const dvd_conv_desc& d = p.d;
  return d.Cin >= 32 && d.Cout >= 64 && p.M >= 4096;
}

int tc_wgrad_launch(ConvP& p, float* dwp, cudaStream_t st) {
  const dvd_conv_desc& d = p.d;
  const int nsm = num_sms();
  const int bn = d.Cout >= 256 ? 256 : (d.Cout >= 128 ? 128 : 64);
  const int64_t base = (int64_t)ceil_div(d.Cin, tc::BM) * ceil_div(d.Cout, bn) * p.taps;
  int nsplit = 1;
  if (base < nsm) {
    nsplit = (int)ceil_div<int64_t>(nsm, base);
    const int maxs = p.M / 1024 > 0 ? p.M / 1024 : 1;
    if (nsplit > maxs) nsplit = maxs;
  }
  int per = ceil_div(p.M, nsplit);
  per = ceil_div(per, tc::BKC) * tc::BKC;
  nsplit = ceil_div(p.M, per);
  p.atomic_out = (nsplit > 1) || d.accumulate;
  if (nsplit > 1 && !d.accumulate)
    DVD_CUDA(cudaMemsetAsync(dwp, 0, sizeof(float) * (size_t)p.taps * d.Cin * d.Cout, st));
  dim3 grid(ceil_div(d.Cin, tc::BM), ceil_div(d.Cout, bn), p.taps * nsplit);
  prof_tag("tc1 wgrad M%d Ci%d Co%d t%d", p.M, d.Cin, d.Cout, p.taps);
  prof_begin(1, 2.0 * p.M * (double)d.Cout * d.Cin * p.taps, st);
  int rc;
  if (bn == 256) rc = tc::launch<256, 1>(p, dwp, grid, nsplit, per, st);
  else if (bn == 128) rc = tc::launch<128, 1>(p, dwp, grid, nsplit, per, st);
  else rc = tc::launch<64, 1>(p, dwp, grid, nsplit, per, st);
  prof_end(1, st);
  if (rc) return rc;
  DVD_LAUNCH_CHECK();
  return 0;
}

}  // namespace dvd
