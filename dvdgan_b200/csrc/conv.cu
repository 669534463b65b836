// Implicit-GEMM convolution engine (fp32 SIMT path): forward / dgrad, weight gradient, weight packing,
// strided-batched SGEMM.  No im2col buffer is ever materialised: the A operand of the GEMM is gathered
// tap by tap (shifted 1x1 projections) straight into shared memory.
//
// GEMM view of the forward:  Y[m][co] = sum_{tap, ci} X[m shifted by tap][ci] * Wp[tap][ci][co]
//   m = output pixel (image, d, h, w)   -- contiguous in NC(D)HW => coalesced gathers and stores
//   K order = (tap outer, ci inner)     -- the tap's shift / zero-padding predicate is hoisted
#include <string>
#include <vector>

#include "common.cuh"
#include "conv_params.cuh"

namespace dvd {

thread_local char g_last_error[512] = "";

constexpr int BK = 8;

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(256, 2) conv_fwd_kernel(const ConvP p) {
  constexpr int TXN = BM / TM;  // threads along M (pixels)
  constexpr int TYN = BN / TN;  // threads along N (cout)
  static_assert(TXN * TYN == 256, "256 threads");
  static_assert(TM % 4 == 0 && TN % 4 == 0, "float4 micro tiles");
  constexpr int AL = BM * BK / 256;
  constexpr int KSTEP = 256 / BM;
  constexpr int MCH = TM / 4, NCH = TN / 4;
  constexpr int MCS = BM / MCH, NCS = BN / NCH;  // chunk strides

  __shared__ __align__(16) float As[2][BK][BM];
  __shared__ __align__(16) float Bs[2][BK][BN];

  const dvd_conv_desc& d = p.d;
  const int tid = threadIdx.x;
  const int tx = tid % TXN, ty = tid / TXN;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

  // ---- per-thread gather state: one pixel, AL channels per k-tile
  const int a_m = tid % BM;
  const int a_k0 = tid / BM;
  const int m = m0 + a_m;
  const bool m_valid = m < p.M;
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
  bool a_valid = false;
  int64_t a_off = 0;
  auto set_tap = [&](int tap) {
    const int kw = tap % d.kW;
    const int t2 = tap / d.kW;
    const int kh = t2 % d.kH;
    const int kd = t2 / d.kH;
    const int iz = pz + kd - d.kD / 2, iy = py + kh - d.kH / 2, ix = px + kw - d.kW / 2;
    a_valid = m_valid && iz >= 0 && iz < d.D && iy >= 0 && iy < d.H && ix >= 0 && ix < d.W;
    a_off = x_base + (int64_t)iz * p.Hs * p.Ws + (int64_t)(iy >> d.in_up) * p.Ws + (ix >> d.in_up);
  };

  int it = blockIdx.z * p.iters_per_split;
  int it_end = it + p.iters_per_split;
  if (it_end > p.iters_total) it_end = p.iters_total;
  int tap = it / p.ck;
  int cchunk = it - tap * p.ck;
  if (it < it_end) set_tap(tap);

  float a_reg[AL];
  float b_reg[4];

  auto load_tile = [&]() {
    const int c0 = cchunk * BK;
#pragma unroll
    for (int j = 0; j < AL; ++j) {
      const int c = c0 + a_k0 + j * KSTEP;
      float v = 0.f;
      if (a_valid && c < d.Cin) {
        v = __ldg(p.x + a_off + (int64_t)c * d.x_cs);
        if (d.in_relu) v = fmaxf(v, 0.f);
      }
      a_reg[j] = v;
    }
    const float* wt = p.w + ((int64_t)tap * d.Cin + c0) * d.Cout;
    if (p.vecB) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (tid < BN * BK / 4) {
        const int k = tid / (BN / 4), n4 = tid % (BN / 4);
        const int co = n0 + n4 * 4;
        if (c0 + k < d.Cin && co < d.Cout) v = __ldg(reinterpret_cast<const float4*>(wt + (int64_t)k * d.Cout + co));
      }
      b_reg[0] = v.x; b_reg[1] = v.y; b_reg[2] = v.z; b_reg[3] = v.w;
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int idx = tid + 256 * e;
        float v = 0.f;
        if (idx < BN * BK) {
          const int k = idx / BN, nn = idx % BN;
          if (c0 + k < d.Cin && n0 + nn < d.Cout) v = __ldg(wt + (int64_t)k * d.Cout + n0 + nn);
        }
        b_reg[e] = v;
      }
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int j = 0; j < AL; ++j) As[buf][a_k0 + j * KSTEP][a_m] = a_reg[j];
    if (p.vecB) {
      if (tid < BN * BK / 4) {
        const int k = tid / (BN / 4), n4 = tid % (BN / 4);
        *reinterpret_cast<float4*>(&Bs[buf][k][n4 * 4]) = make_float4(b_reg[0], b_reg[1], b_reg[2], b_reg[3]);
      }
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int idx = tid + 256 * e;
        if (idx < BN * BK) Bs[buf][idx / BN][idx % BN] = b_reg[e];
      }
    }
  };
  auto advance = [&]() {
    ++it;
    if (++cchunk == p.ck) {
      cchunk = 0;
      ++tap;
      if (it < it_end) set_tap(tap);
    }
  };

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  int cur = 0;
  if (it < it_end) {
    load_tile();
    store_tile(0);
    advance();
  }
  __syncthreads();
  const int n_iters = it_end - (int)(blockIdx.z * p.iters_per_split);
  for (int s = 0; s < n_iters; ++s) {
    const bool more = (s + 1) < n_iters;
    if (more) load_tile();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int c = 0; c < MCH; ++c)
        *reinterpret_cast<float4*>(&a[c * 4]) = *reinterpret_cast<const float4*>(&As[cur][k][c * MCS + tx * 4]);
#pragma unroll
      for (int c = 0; c < NCH; ++c)
        *reinterpret_cast<float4*>(&b[c * 4]) = *reinterpret_cast<const float4*>(&Bs[cur][k][c * NCS + ty * 4]);
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) {
      store_tile(cur ^ 1);
      advance();
    }
    __syncthreads();
    cur ^= 1;
  }

  // ---- epilogue: bias, residual, activation, accumulate / split-K atomics
  const bool lead = (blockIdx.z == 0);
#pragma unroll
  for (int c = 0; c < MCH; ++c) {
    const int mb = m0 + c * MCS + tx * 4;
    if (mb >= p.M) continue;
    // decode the 4 pixels of this chunk
    int64_t yo[4], ro[4];
    bool ok[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int mm = mb + i;
      ok[i] = mm < p.M;
      const int n = ok[i] ? mm / p.DHW : 0;
      const int rem = mm - n * p.DHW;
      const int n1 = n / d.N2, n2 = n - n1 * d.N2;
      yo[i] = (int64_t)n1 * d.y_s1 + (int64_t)n2 * d.y_s2 + rem;
      if (p.res) {
        int rr = rem;
        if (d.res_up) {
          const int z = rem / p.HW;
          const int r2 = rem - z * p.HW;
          const int hh = r2 / d.W, ww = r2 - hh * d.W;
          rr = (z * (d.H >> 1) + (hh >> 1)) * (d.W >> 1) + (ww >> 1);
        }
        ro[i] = (int64_t)n1 * d.r_s1 + (int64_t)n2 * d.r_s2 + rr;
      } else {
        ro[i] = 0;
      }
    }
#pragma unroll
    for (int cn = 0; cn < NCH; ++cn) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int co = n0 + cn * NCS + ty * 4 + j;
        if (co >= d.Cout) continue;
        float v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = acc[c * 4 + i][cn * 4 + j];
        if (lead) {
          if (p.bias) {
            const float bb = __ldg(p.bias + co);
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] += bb;
          }
          if (p.res) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (ok[i]) v[i] += __ldg(p.res + ro[i] + (int64_t)co * d.r_cs);
          }
        }
        if (p.atomic_out) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (ok[i]) atomicAdd(p.y + yo[i] + (int64_t)co * d.y_cs, v[i]);
        } else if (p.vecY) {
          float4* dst = reinterpret_cast<float4*>(p.y + yo[0] + (int64_t)co * d.y_cs);
          float4 o = make_float4(v[0], v[1], v[2], v[3]);
          if (d.accumulate) {
            const float4 old = *dst;
            o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
          }
          if (d.out_act == 1) {
            o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
          } else if (d.out_act == 2) {
            o.x = tanhf(o.x); o.y = tanhf(o.y); o.z = tanhf(o.z); o.w = tanhf(o.w);
          }
          *dst = o;
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (!ok[i]) continue;
            float* dst = p.y + yo[i] + (int64_t)co * d.y_cs;
            float o = v[i];
            if (d.accumulate) o += *dst;
            if (d.out_act == 1) o = fmaxf(o, 0.f);
            else if (d.out_act == 2) o = tanhf(o);
            *dst = o;
          }
        }
      }
    }
  }
}

// zero-fill a strided NC(D)HW view (split-K without accumulate on a non-dense output)
__global__ void zero_view_kernel(float* y, int N2, int Cout, int DHW, int64_t s1, int64_t s2, int64_t cs,
                                 int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int pix = (int)(i % DHW);
    int64_t t = i / DHW;
    const int co = (int)(t % Cout);
    const int n = (int)(t / Cout);
    const int n1 = n / N2, n2 = n - n1 * N2;
    y[(int64_t)n1 * s1 + (int64_t)n2 * s2 + (int64_t)co * cs + pix] = 0.f;
  }
}

int zero_output_view(const ConvP& p, cudaStream_t st) {
  const dvd_conv_desc* d = &p.d;
  const int64_t img = (int64_t)d->Cout * p.DHW;
  const bool dense_y = (d->y_cs == p.DHW) && (d->N2 == 1 || d->y_s2 == img) &&
                       (d->N1 == 1 || d->y_s1 == (int64_t)d->N2 * img);
  const int64_t total = (int64_t)d->N1 * d->N2 * img;
  if (dense_y) {
    DVD_CUDA(cudaMemsetAsync(p.y, 0, sizeof(float) * (size_t)total, st));
  } else {
    zero_view_kernel<<<ew_blocks(total, 1), 256, 0, st>>>(p.y, d->N2, d->Cout, p.DHW, d->y_s1, d->y_s2, d->y_cs, total);
    DVD_LAUNCH_CHECK();
  }
  return 0;
}

static int check_desc(const dvd_conv_desc* d) {
  DVD_CHECK_ARG(d != nullptr);
  DVD_CHECK_ARG(d->N1 > 0 && d->N2 > 0 && d->Cin > 0 && d->Cout > 0);
  DVD_CHECK_ARG(d->D > 0 && d->H > 0 && d->W > 0);
  DVD_CHECK_ARG((d->kD & 1) && (d->kH & 1) && (d->kW & 1));
  DVD_CHECK_ARG(d->in_up == 0 || d->in_up == 1);
  DVD_CHECK_ARG(d->in_up == 0 || ((d->H % 2 == 0) && (d->W % 2 == 0)));
  DVD_CHECK_ARG(d->res_up == 0 || ((d->H % 2 == 0) && (d->W % 2 == 0)));
  DVD_CHECK_ARG((int64_t)d->N1 * d->N2 * d->D * d->H * d->W < (int64_t)1 << 31);
  return 0;
}

static void fill_common(ConvP& p, const dvd_conv_desc* d) {
  p.d = *d;
  p.HW = d->H * d->W;
  p.DHW = d->D * p.HW;
  p.M = d->N1 * d->N2 * p.DHW;
  p.taps = d->kD * d->kH * d->kW;
  p.Hs = d->H >> d->in_up;
  p.Ws = d->W >> d->in_up;
  p.ck = ceil_div(d->Cin, BK);
  p.iters_total = p.taps * p.ck;
}

template <int BM, int BN, int TM, int TN>
static int launch_fwd(ConvP& p, cudaStream_t st) {
  dim3 grid(ceil_div(p.M, BM), ceil_div(p.d.Cout, BN), p.nsplit);
  prof_tag("simt fwd M%d Ci%d Co%d t%d", p.M, p.d.Cin, p.d.Cout, p.taps);
  prof_begin(0, 2.0 * p.M * (double)p.d.Cout * p.d.Cin * p.taps, st);
  conv_fwd_kernel<BM, BN, TM, TN><<<grid, 256, 0, st>>>(p);
  prof_end(0, st);
  DVD_LAUNCH_CHECK();
  return 0;
}

}  // namespace dvd

namespace dvd {
// library-internal: forward conv on the TMA/tcgen05 engine with pre-split operands and/or a ConvGRU epilogue
bool conv_fwd_ex_eligible(const dvd_conv_desc* d) {
  if (check_desc(d)) return false;
  ConvP p;
  fill_common(p, d);
  return tma_fwd_launch_ex_eligible(p);
}
int conv_fwd_ex(const dvd_conv_desc* d, const float* x, const float* w_packed, float* y, const TmaOperands* ops,
                const GruEpi* epi, cudaStream_t st) {
  DVD_TRY(check_desc(d));
  DVD_CHECK_ARG(y && (x || (ops && ops->a_hi)) && (w_packed || (ops && ops->w_hi)));
  ConvP p;
  fill_common(p, d);
  p.x = x; p.w = w_packed; p.bias = nullptr; p.res = nullptr; p.y = y;
  p.vecB = p.vecY = 0;
  if (!tma_fwd_launch_ex_eligible(p)) return fail("conv_fwd_ex: shape not eligible%s (%s:%d)", "", __FILE__, __LINE__);
  return tma_fwd_launch_ex(p, ops, epi, st);
}
bool conv_wgrad_ex_eligible(const dvd_conv_desc* d, const TmaWgOperands* ops) {
  if (check_desc(d)) return false;
  ConvP p;
  fill_common(p, d);
  return tma_wgrad_eligible(p) && tma_wgrad_ex_ok(p, ops);
}
int conv_wgrad_ex(const dvd_conv_desc* d, const float* x, float* dwp, const TmaWgOperands* ops, cudaStream_t st) {
  DVD_TRY(check_desc(d));
  DVD_CHECK_ARG(x && dwp && ops && ops->y_hi && ops->y_lo);
  ConvP p;
  fill_common(p, d);
  p.x = x; p.y = nullptr; p.w = nullptr; p.bias = nullptr; p.res = nullptr;
  if (!(tma_wgrad_eligible(p) && tma_wgrad_ex_ok(p, ops)))
    return fail("conv_wgrad_ex: shape not eligible%s (%s:%d)", "", __FILE__, __LINE__);
  return tma_wgrad_launch_ex(p, dwp, ops, st);
}
}  // namespace dvd

using namespace dvd;

extern "C" int dvd_conv_fwd(const dvd_conv_desc* d, const float* x, const float* w_packed, const float* bias,
                            const float* res, float* y, void* stream) {
  DVD_TRY(check_desc(d));
  DVD_CHECK_ARG(x && w_packed && y);
  cudaStream_t st = as_stream(stream);
  ConvP p;
  fill_common(p, d);
  p.x = x; p.w = w_packed; p.bias = bias; p.res = res; p.y = y;
  p.vecB = (d->Cout % 4 == 0) && ((reinterpret_cast<uintptr_t>(w_packed) & 15) == 0);
  p.vecY = (p.DHW % 4 == 0) && ((reinterpret_cast<uintptr_t>(y) & 15) == 0) && (d->y_s1 % 4 == 0) &&
           (d->y_s2 % 4 == 0) && (d->y_cs % 4 == 0);
  if (tma_fwd_eligible(p)) return tma_fwd_launch(p, st);
  const int nsm = num_sms();
  // tile selection
  int tile;  // 0: 128x128  1: 128x64  2: 64x64  3: 256x16
  if (d->Cout <= 16) tile = 3;
  else if (d->Cout <= 64) tile = 1;
  else tile = 0;
  auto ctas_of = [&](int t) {
    const int bm = t == 3 ? 256 : (t == 2 ? 64 : 128), bn = t == 0 ? 128 : (t == 3 ? 16 : 64);
    return (int64_t)ceil_div(p.M, bm) * ceil_div(d->Cout, bn);
  };
  if (tile <= 1 && ctas_of(tile) < 2 * nsm) tile = 2;
  int64_t ctas = ctas_of(tile);
  // split-K for under-filled grids (small-spatial ConvGRU stages)
  p.nsplit = 1;
  if (ctas < nsm && d->out_act == 0 && p.iters_total >= 32) {
    int want = (int)ceil_div<int64_t>(2 * nsm, ctas);
    int maxs = p.iters_total / 16;
    if (want > maxs) want = maxs;
    if (want > 32) want = 32;
    if (want > 1) p.nsplit = want;
  }
  p.iters_per_split = ceil_div(p.iters_total, p.nsplit);
  p.nsplit = ceil_div(p.iters_total, p.iters_per_split);
  p.atomic_out = p.nsplit > 1;
  if (p.atomic_out && !d->accumulate) DVD_TRY(zero_output_view(p, st));
  switch (tile) {
    case 0: return launch_fwd<128, 128, 8, 8>(p, st);
    case 1: return launch_fwd<128, 64, 8, 4>(p, st);
    case 2: return launch_fwd<64, 64, 4, 4>(p, st);
    default: return launch_fwd<256, 16, 4, 4>(p, st);
  }
}

// ---------------------------------------------------------------------------------------------
// weight gradient:  dWp[tap][ci][co] += sum_m X[m + tap][ci] * dY[m][co]
// GEMM with M' = ci, N' = co, K' = pixels; one tap (and one pixel range) per CTA.
// ---------------------------------------------------------------------------------------------
namespace dvd {

template <int BC, int BO>
__global__ void __launch_bounds__(256, 2) conv_wgrad_kernel(const ConvP p, float* __restrict__ dwp, int nsplit,
                                                         int pix_per_split, int atomic_out) {
  constexpr int KP = 8;          // pixels per k-tile
  constexpr int PAD = 4;
  constexpr int TC = BC / 16, TO = BO / 16;  // micro tile (ci x co)
  static_assert(TC % 4 == 0 && TO % 4 == 0, "float4 micro tiles");
  constexpr int CCH = TC / 4, OCH = TO / 4;
  constexpr int CCS = BC / CCH, OCS = BO / OCH;
  constexpr int ALD = BC * KP / 256, BLD = BO * KP / 256;

  __shared__ __align__(16) float As[2][KP][BC + PAD];
  __shared__ __align__(16) float Bs[2][KP][BO + PAD];

  const dvd_conv_desc& d = p.d;
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;      // tx -> co (contiguous in dwp), ty -> ci
  const int c0 = blockIdx.x * BC, o0 = blockIdx.y * BO;
  const int tap = blockIdx.z / nsplit;
  const int split = blockIdx.z - tap * nsplit;
  const int kw = tap % d.kW;
  const int t2 = tap / d.kW;
  const int kh = t2 % d.kH;
  const int kd = t2 / d.kH;
  const int oz = kd - d.kD / 2, oy = kh - d.kH / 2, ox = kw - d.kW / 2;

  int mk = split * pix_per_split;
  int m_end = mk + pix_per_split;
  if (m_end > p.M) m_end = p.M;

  const int l_pix = tid % KP, l_row = tid / KP;  // 32 rows per pass
  float a_reg[ALD], b_reg[BLD];

  auto load_tile = [&](int mbase) {
    const int m = mbase + l_pix;
    bool mv = m < m_end;
    bool av = false;
    int64_t xo = 0, yo = 0;
    if (mv) {
      const int n = m / p.DHW;
      int rem = m - n * p.DHW;
      const int pz = rem / p.HW;
      const int r2 = rem - pz * p.HW;
      const int py = r2 / d.W;
      const int px = r2 - py * d.W;
      const int n1 = n / d.N2, n2 = n - n1 * d.N2;
      yo = (int64_t)n1 * d.y_s1 + (int64_t)n2 * d.y_s2 + rem;
      const int iz = pz + oz, iy = py + oy, ix = px + ox;
      av = iz >= 0 && iz < d.D && iy >= 0 && iy < d.H && ix >= 0 && ix < d.W;
      xo = (int64_t)n1 * d.x_s1 + (int64_t)n2 * d.x_s2 + (int64_t)iz * p.Hs * p.Ws +
           (int64_t)(iy >> d.in_up) * p.Ws + (ix >> d.in_up);
    }
#pragma unroll
    for (int j = 0; j < ALD; ++j) {
      const int ci = c0 + l_row + 32 * j;
      float v = 0.f;
      if (av && ci < d.Cin) {
        v = __ldg(p.x + xo + (int64_t)ci * d.x_cs);
        if (d.in_relu) v = fmaxf(v, 0.f);
      }
      a_reg[j] = v;
    }
#pragma unroll
    for (int j = 0; j < BLD; ++j) {
      const int co = o0 + l_row + 32 * j;
      float v = 0.f;
      if (mv && co < d.Cout) v = __ldg(p.y + yo + (int64_t)co * d.y_cs);
      b_reg[j] = v;
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int j = 0; j < ALD; ++j) As[buf][l_pix][l_row + 32 * j] = a_reg[j];
#pragma unroll
    for (int j = 0; j < BLD; ++j) Bs[buf][l_pix][l_row + 32 * j] = b_reg[j];
  };

  float acc[TC][TO];
#pragma unroll
  for (int i = 0; i < TC; ++i)
#pragma unroll
    for (int j = 0; j < TO; ++j) acc[i][j] = 0.f;

  int cur = 0;
  if (mk < m_end) {
    load_tile(mk);
    store_tile(0);
  }
  __syncthreads();
  for (; mk < m_end; mk += KP) {
    const bool more = (mk + KP) < m_end;
    if (more) load_tile(mk + KP);
#pragma unroll
    for (int k = 0; k < KP; ++k) {
      float a[TC], b[TO];
#pragma unroll
      for (int c = 0; c < CCH; ++c)
        *reinterpret_cast<float4*>(&a[c * 4]) = *reinterpret_cast<const float4*>(&As[cur][k][c * CCS + ty * 4]);
#pragma unroll
      for (int c = 0; c < OCH; ++c)
        *reinterpret_cast<float4*>(&b[c * 4]) = *reinterpret_cast<const float4*>(&Bs[cur][k][c * OCS + tx * 4]);
#pragma unroll
      for (int i = 0; i < TC; ++i)
#pragma unroll
        for (int j = 0; j < TO; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) store_tile(cur ^ 1);
    __syncthreads();
    cur ^= 1;
  }

#pragma unroll
  for (int c = 0; c < CCH; ++c)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int ci = c0 + c * CCS + ty * 4 + i;
      if (ci >= d.Cin) continue;
#pragma unroll
      for (int cn = 0; cn < OCH; ++cn)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int co = o0 + cn * OCS + tx * 4 + j;
          if (co >= d.Cout) continue;
          float* dst = dwp + ((int64_t)tap * d.Cin + ci) * d.Cout + co;
          const float v = acc[c * 4 + i][cn * 4 + j];
          if (atomic_out) atomicAdd(dst, v);
          else *dst = v;
        }
    }
}

}  // namespace dvd

extern "C" int dvd_conv_wgrad(const dvd_conv_desc* d, const float* x, const float* dy, float* dwp, void* stream) {
  DVD_TRY(check_desc(d));
  DVD_CHECK_ARG(x && dy && dwp);
  cudaStream_t st = as_stream(stream);
  ConvP p;
  fill_common(p, d);
  p.x = x; p.y = const_cast<float*>(dy); p.w = nullptr; p.bias = nullptr; p.res = nullptr;
  if (tma_wgrad_eligible(p)) return tma_wgrad_launch(p, dwp, st);
  const int nsm = num_sms();
  const bool small = (d->Cin <= 64 && d->Cout <= 64);
  const int bc = small ? 64 : 128, bo = small ? 64 : 128;
  const int64_t base = (int64_t)ceil_div(d->Cin, bc) * ceil_div(d->Cout, bo) * p.taps;
  int nsplit = 1;
  if (base < 2 * nsm) {
    nsplit = (int)ceil_div<int64_t>(2 * nsm, base);
    const int maxs = p.M / 512 > 0 ? p.M / 512 : 1;
    if (nsplit > maxs) nsplit = maxs;
  }
  int pps = ceil_div(p.M, nsplit);
  pps = ceil_div(pps, 8) * 8;
  nsplit = ceil_div(p.M, pps);
  const int atomic_out = (nsplit > 1) || d->accumulate;
  if (nsplit > 1 && !d->accumulate)
    DVD_CUDA(cudaMemsetAsync(dwp, 0, sizeof(float) * (size_t)p.taps * d->Cin * d->Cout, st));
  dim3 grid(ceil_div(d->Cin, bc), ceil_div(d->Cout, bo), p.taps * nsplit);
  prof_tag("simt wgrad M%d Ci%d Co%d t%d", p.M, d->Cin, d->Cout, p.taps);
  prof_begin(1, 2.0 * p.M * (double)d->Cout * d->Cin * p.taps, st);
  if (small) conv_wgrad_kernel<64, 64><<<grid, 256, 0, st>>>(p, dwp, nsplit, pps, atomic_out);
  else conv_wgrad_kernel<128, 128><<<grid, 256, 0, st>>>(p, dwp, nsplit, pps, atomic_out);
  prof_end(1, st);
  DVD_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// weight pack / unpack
// ---------------------------------------------------------------------------------------------
namespace dvd {

__global__ void weight_pack_kernel(const float* __restrict__ w, int Ci_total, int taps, int co0, int Cout, int ci0,
                                   int Cin, const float* __restrict__ sigma, int transpose, float* __restrict__ dst,
                                   int dst_rows, int row_off, int dst_ld, int col_off) {
  const float scale = sigma ? 1.f / __ldg(sigma) : 1.f;
  const int64_t total = (int64_t)taps * Cin * Cout;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    // destination-major enumeration => coalesced stores
    int tap, ci, co;
    int64_t o;
    if (!transpose) {
      co = (int)(i % Cout);
      int64_t t = i / Cout;
      ci = (int)(t % Cin);
      tap = (int)(t / Cin);
      o = ((int64_t)tap * dst_rows + row_off + ci) * dst_ld + col_off + co;
    } else {
      ci = (int)(i % Cin);
      int64_t t = i / Cin;
      co = (int)(t % Cout);
      const int tapf = (int)(t / Cout);
      tap = taps - 1 - tapf;
      o = ((int64_t)tapf * dst_rows + row_off + co) * dst_ld + col_off + ci;
    }
    dst[o] = __ldg(w + ((int64_t)(co0 + co) * Ci_total + ci0 + ci) * taps + tap) * scale;
  }
}

__global__ void weight_unpack_kernel(const float* __restrict__ src, int src_ld, int src_off, int Ci_total, int taps,
                                     int co0, int Cout, int ci0, int Cin, int accumulate, float* __restrict__ wg) {
  const int64_t total = (int64_t)taps * Cin * Cout;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    // destination-major: (co, ci, tap)
    const int tap = (int)(i % taps);
    int64_t t = i / taps;
    const int ci = (int)(t % Cin);
    const int co = (int)(t / Cin);
    const float v = __ldg(src + ((int64_t)tap * Cin + ci) * src_ld + src_off + co);
    float* dst = wg + ((int64_t)(co0 + co) * Ci_total + ci0 + ci) * taps + tap;
    *dst = accumulate ? *dst + v : v;
  }
}

}  // namespace dvd

extern "C" int dvd_weight_pack(const float* w, int Ci_total, int taps, int co0, int Cout, int ci0, int Cin,
                               const float* sigma, int transpose, float* dst, int dst_rows, int dst_row_off,
                               int dst_ld, int dst_col_off, void* stream) {
  dvd::ProfScope _ps(3, "weight_pack", dvd::as_stream(stream));
  DVD_CHECK_ARG(w && dst && Ci_total > 0 && taps > 0 && Cout > 0 && Cin > 0 && ci0 >= 0 && co0 >= 0);
  DVD_CHECK_ARG(ci0 + Cin <= Ci_total);
  const int64_t total = (int64_t)taps * Cin * Cout;
  weight_pack_kernel<<<ew_blocks(total, 2), 256, 0, as_stream(stream)>>>(w, Ci_total, taps, co0, Cout, ci0, Cin, sigma,
                                                                        transpose, dst, dst_rows, dst_row_off, dst_ld,
                                                                        dst_col_off);
  DVD_LAUNCH_CHECK();
  return 0;
}

extern "C" int dvd_weight_unpack(const float* src, int src_ld, int src_off, int Ci_total, int taps, int co0, int Cout,
                                 int ci0, int Cin, int accumulate, float* w_grad, void* stream) {
  dvd::ProfScope _ps(3, "weight_unpack", dvd::as_stream(stream));
  DVD_CHECK_ARG(src && w_grad && Ci_total > 0 && taps > 0 && Cout > 0 && Cin > 0);
  DVD_CHECK_ARG(ci0 + Cin <= Ci_total);
  const int64_t total = (int64_t)taps * Cin * Cout;
  weight_unpack_kernel<<<ew_blocks(total, 2), 256, 0, as_stream(stream)>>>(src, src_ld, src_off, Ci_total, taps, co0,
                                                                          Cout, ci0, Cin, accumulate, w_grad);
  DVD_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// strided-batched SGEMM (row-major), 64x64x16 tiles, 4x4 micro tiles
// ---------------------------------------------------------------------------------------------
namespace dvd {

__global__ void __launch_bounds__(256) bgemm_kernel(int transA, int transB, int M, int N, int K, float alpha,
                                                    const float* __restrict__ A, int lda, int64_t sA,
                                                    const float* __restrict__ B, int ldb, int64_t sB, float beta,
                                                    float* __restrict__ C, int ldc, int64_t sC,
                                                    const float* __restrict__ bias) {
  constexpr int T = 64, KT = 16;
  __shared__ __align__(16) float As[KT][T + 4];
  __shared__ __align__(16) float Bs[KT][T + 4];
  const int b = blockIdx.z;
  A += (int64_t)b * sA;
  B += (int64_t)b * sB;
  C += (int64_t)b * sC;
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  const int m0 = blockIdx.y * T, n0 = blockIdx.x * T;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += KT) {
    // A tile: element (m, k).  Pick the thread->element map that is contiguous in memory.
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = tid + 256 * e;
      int mm, kk;
      if (transA) { mm = idx % T; kk = idx / T; }   // A[k][m]: m contiguous
      else { kk = idx % KT; mm = idx / KT; }        // A[m][k]: k contiguous
      const int gm = m0 + mm, gk = k0 + kk;
      float v = 0.f;
      if (gm < M && gk < K) v = transA ? __ldg(A + (int64_t)gk * lda + gm) : __ldg(A + (int64_t)gm * lda + gk);
      As[kk][mm] = v;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = tid + 256 * e;
      int nn, kk;
      if (transB) { kk = idx % KT; nn = idx / KT; }  // B[n][k]: k contiguous
      else { nn = idx % T; kk = idx / T; }           // B[k][n]: n contiguous
      const int gn = n0 + nn, gk = k0 + kk;
      float v = 0.f;
      if (gn < N && gk < K) v = transB ? __ldg(B + (int64_t)gn * ldb + gk) : __ldg(B + (int64_t)gk * ldb + gn);
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < KT; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 bb = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float* dst = C + (int64_t)gm * ldc + gn;
      float v = alpha * acc[i][j];
      if (bias) v += __ldg(bias + gn);
      if (beta != 0.f) v += beta * *dst;
      *dst = v;
    }
  }
}

}  // namespace dvd

extern "C" int dvd_bgemm(int transA, int transB, int M, int N, int K, float alpha, const float* A, int lda,
                         int64_t strideA, const float* B, int ldb, int64_t strideB, float beta, float* C, int ldc,
                         int64_t strideC, int batch, const float* bias, void* stream) {
  dvd::ProfScope _ps(3, "bgemm", dvd::as_stream(stream));
  DVD_CHECK_ARG(A && B && C && M > 0 && N > 0 && K >= 0 && batch > 0 && batch <= 65535);
  dim3 grid(ceil_div(N, 64), ceil_div(M, 64), batch);
  DVD_CHECK_ARG(grid.y <= 65535);
  bgemm_kernel<<<grid, 256, 0, as_stream(stream)>>>(transA, transB, M, N, K, alpha, A, lda, strideA, B, ldb, strideB,
                                                    beta, C, ldc, strideC, bias);
  DVD_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// launch counter + CUDA-event profiler for the dense engines
// ---------------------------------------------------------------------------------------------
#include <mutex>
#include <vector>
namespace dvd {
std::atomic<long long> g_launches{0};
namespace {
struct ProfRec { cudaEvent_t a, b; double flops; char tag[56]; };
thread_local char g_prof_tag[56] = "";
std::mutex g_prof_mu;
unsigned g_prof_on = 0;            // bit c set: category c is being recorded
constexpr int kProfCats = 4;       // 0: conv fwd/dgrad GEMM, 1: wgrad GEMM, 2: operand-plane preparation, 3: helpers
std::vector<ProfRec> g_prof[kProfCats];
std::vector<ProfRec> g_pool;       // recycled event pairs
long long g_prof_dropped[kProfCats] = {0, 0, 0, 0};
constexpr size_t kProfMax = 1 << 17;
}  // namespace

thread_local int g_prof_open[kProfCats] = {-1, -1, -1, -1};   // index of the record the depth-0 prof_end closes
thread_local int g_prof_depth[kProfCats] = {0, 0, 0, 0};      // brackets of one category may nest (entry points that
                                                               // call other entry points): only the outermost records

void prof_begin(int cat, double flops, cudaStream_t st) {
  if (g_prof_depth[cat]++ > 0) return;          // nested bracket of the same category: the outer one covers it
  if (!(g_prof_on >> cat & 1u)) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (g_prof[cat].size() >= kProfMax) { ++g_prof_dropped[cat]; return; }
  ProfRec r;
  if (!g_pool.empty()) { r = g_pool.back(); g_pool.pop_back(); }
  else if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
  r.flops = flops;
  memcpy(r.tag, g_prof_tag, sizeof(r.tag));
  cudaEventRecord(r.a, st);
  g_prof[cat].push_back(r);
  g_prof_open[cat] = (int)g_prof[cat].size() - 1;
}
void prof_tag(const char* fmt, int a, int b, int c, int d, int e, int f) {
  if (!g_prof_on) return;
  snprintf(g_prof_tag, sizeof(g_prof_tag), fmt, a, b, c, d, e, f);
}
void prof_end(int cat, cudaStream_t st) {
  if (g_prof_depth[cat] > 0 && --g_prof_depth[cat] > 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  const int i = g_prof_open[cat];
  if (i < 0) return;
  if (i < (int)g_prof[cat].size()) cudaEventRecord(g_prof[cat][i].b, st);
  g_prof_open[cat] = -1;
}
}  // namespace dvd

extern "C" long long dvd_launch_count(void) { return dvd::g_launches.load(); }

extern "C" int dvd_prof_enable(int on) {
  std::lock_guard<std::mutex> lk(dvd::g_prof_mu);
  // on = 1: the dense engines only (categories 0, 1); otherwise a bit mask of categories (0xF: everything)
  dvd::g_prof_on = on == 1 ? 0x3u : (unsigned)on;
  return 0;
}

extern "C" int dvd_prof_read(int category, double* ms, double* flops, long long* launches) {
  DVD_CHECK_ARG(category >= 0 && category < dvd::kProfCats);
  DVD_CHECK_ARG(ms && flops && launches);
  std::lock_guard<std::mutex> lk(dvd::g_prof_mu);
  double t = 0.0, f = 0.0;
  for (auto& r : dvd::g_prof[category]) {
    float e = 0.f;
    if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&e, r.a, r.b) == cudaSuccess) {
      t += e;
      f += r.flops;
    } else {
      cudaGetLastError();      // a bracket that never closed (error path): drop it, do not poison later launches
    }
    dvd::g_pool.push_back(r);
  }
  *ms = t;
  *flops = f;
  *launches = (long long)dvd::g_prof[category].size() + dvd::g_prof_dropped[category];
  dvd::g_prof[category].clear();
  dvd::g_prof_dropped[category] = 0;
  return 0;
}

// per-shape table of the recorded launches (does not clear them): "cat \t tag \t launches \t ms \t flops"
extern "C" int dvd_prof_dump(const char* path) {
  DVD_CHECK_ARG(path);
  std::lock_guard<std::mutex> lk(dvd::g_prof_mu);
  FILE* f = fopen(path, "w");
  if (!f) return dvd::fail("cannot open %s (%s:%d)", path, __FILE__, __LINE__);
  for (int cat = 0; cat < dvd::kProfCats; ++cat) {
    std::vector<std::string> keys;
    std::vector<double> ms, fl;
    std::vector<long long> cnt;
    for (auto& r : dvd::g_prof[cat]) {
      if (cudaEventSynchronize(r.b) != cudaSuccess) { cudaGetLastError(); continue; }
      float e = 0.f;
      if (cudaEventElapsedTime(&e, r.a, r.b) != cudaSuccess) { cudaGetLastError(); continue; }
      size_t i = 0;
      for (; i < keys.size(); ++i) if (keys[i] == r.tag) break;
      if (i == keys.size()) { keys.push_back(r.tag); ms.push_back(0); fl.push_back(0); cnt.push_back(0); }
      ms[i] += e; fl[i] += r.flops; cnt[i] += 1;
    }
    for (size_t i = 0; i < keys.size(); ++i)
      fprintf(f, "%d\t%s\t%lld\t%.4f\t%.6g\n", cat, keys[i].c_str(), cnt[i], ms[i], fl[i]);
  }
  fclose(f);
  return 0;
}

extern "C" const char* dvd_last_error(void) { return dvd::g_last_error; }
extern "C" int dvd_abi_version(void) { return 2; }

// ---------------------------------------------------------------------------------------------
// process-wide options
// ---------------------------------------------------------------------------------------------
namespace dvd {
namespace {
struct OptDef { const char* name; int def; };
const OptDef kOptDefs[OPT_COUNT] = {
    {"simt_only", 0}, {"pair", 1}, {"persist", 1}, {"oneacc", 0}, {"occ2", 1}, {"epi_prefetch", 1},
    {"gru_fused", 1}, {"gru_share_planes", 1}, {"gru_bwd_planes", 1}, {"gru_bwd_fused", 0}, {"flash_attn", 1},
    {"fwd_bf16", 0}, {"gru_streams", 2},
};
std::atomic<int> g_opts[OPT_COUNT];
std::atomic<bool> g_opts_init{false};
void opts_init() {
  if (g_opts_init.load(std::memory_order_acquire)) return;
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  if (g_opts_init.load(std::memory_order_relaxed)) return;
  for (int i = 0; i < OPT_COUNT; ++i) g_opts[i].store(kOptDefs[i].def, std::memory_order_relaxed);
  g_opts_init.store(true, std::memory_order_release);
}
int opt_index(const char* name) {
  for (int i = 0; i < OPT_COUNT; ++i)
    if (strcmp(kOptDefs[i].name, name) == 0) return i;
  return -1;
}
}  // namespace
int get_option(int opt) {
  opts_init();
  return g_opts[opt].load(std::memory_order_relaxed);
}
}  // namespace dvd

extern "C" int dvd_set_option(const char* name, int value) {
  DVD_CHECK_ARG(name != nullptr);
  dvd::opts_init();
  const int i = dvd::opt_index(name);
  if (i < 0) return dvd::fail("unknown option '%s' (%s:%d)", name, __FILE__, __LINE__);
  dvd::g_opts[i].store(value, std::memory_order_relaxed);
  return 0;
}
extern "C" int dvd_get_option(const char* name, int* value) {
  DVD_CHECK_ARG(name != nullptr && value != nullptr);
  dvd::opts_init();
  const int i = dvd::opt_index(name);
  if (i < 0) return dvd::fail("unknown option '%s' (%s:%d)", name, __FILE__, __LINE__);
  *value = dvd::g_opts[i].load(std::memory_order_relaxed);
  return 0;
}

// Forward operands are split into fp16 planes (22 bits); elements beyond +-65504 are clamped there.  Reads -- and with
// reset != 0 clears -- the number of 8-element groups that were clamped on the current device.  Synchronises `stream`.
extern "C" int dvd_saturation_count(unsigned int* count, int reset, void* stream) {
  DVD_CHECK_ARG(count != nullptr);
  return dvd::tma_saturation_count(count, reset, dvd::as_stream(stream));
}

// Bytes of the library's stream-ordered scratch pools on the current device (one pool per stream that has run a
// tensor-core GEMM): the operand planes live there (cudaMallocFromPoolAsync), next to whatever allocator the caller uses
// for its tensors.  high_water = sum of the pools' used-memory high-water marks, reserved = what they hold right now.
extern "C" int dvd_scratch_bytes(long long* high_water, long long* reserved) {
  DVD_CHECK_ARG(high_water != nullptr && reserved != nullptr);
  return dvd::tma_scratch_stats(high_water, reserved);
}
