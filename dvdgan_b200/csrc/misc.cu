// Memory-bound helpers of the step: pooling, phi, frame gather, permutes, activations, reductions,
// embedding, discriminator head, GAN losses, fused Adam.  Grid-stride loops over a grid that is a
// multiple of the SM count; coalesced along the innermost dimension.
#include "common.cuh"

namespace dvd {

#define GRID_STRIDE(i, n) \
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (n); i += (int64_t)gridDim.x * blockDim.x)

// ---------------------------------------------------------------------------------------------- pools
__global__ void avgpool_fwd_kernel(const float* __restrict__ x, int64_t NC, int D, int H, int W, int pd, int ph, int pw,
                                   float scale, int accumulate, float* __restrict__ y) {
  const int Do = D / pd, Ho = H / ph, Wo = W / pw;
  const float inv = scale / (pd * ph * pw);
  const int64_t total = NC * Do * Ho * Wo;
  GRID_STRIDE(i, total) {
    const int xo = (int)(i % Wo);
    int64_t t = i / Wo;
    const int yo = (int)(t % Ho);
    t /= Ho;
    const int zo = (int)(t % Do);
    const int64_t nc = t / Do;
    const float* src = x + ((nc * D + (int64_t)zo * pd) * H + (int64_t)yo * ph) * W + (int64_t)xo * pw;
    float s = 0.f;
    for (int a = 0; a < pd; ++a)
      for (int b = 0; b < ph; ++b)
        for (int c = 0; c < pw; ++c) s += __ldg(src + ((int64_t)a * H + b) * W + c);
    s *= inv;
    y[i] = accumulate ? y[i] + s : s;
  }
}

__global__ void avgpool_bwd_kernel(const float* __restrict__ dy, int64_t NC, int D, int H, int W, int pd, int ph,
                                   int pw, int accumulate, float* __restrict__ dx) {
  const int Do = D / pd, Ho = H / ph, Wo = W / pw;
  const float inv = 1.f / (pd * ph * pw);
  const int64_t total = NC * D * H * W;
  GRID_STRIDE(i, total) {
    const int xx = (int)(i % W);
    int64_t t = i / W;
    const int yy = (int)(t % H);
    t /= H;
    const int zz = (int)(t % D);
    const int64_t nc = t / D;
    const int zo = zz / pd, yo = yy / ph, xo = xx / pw;
    float v = 0.f;
    if (zo < Do && yo < Ho && xo < Wo) v = __ldg(dy + ((nc * Do + zo) * Ho + yo) * Wo + xo) * inv;
    dx[i] = accumulate ? dx[i] + v : v;
  }
}

// 2x2 spatial pooling (pd = 1) with W a multiple of 4 and W/4 a power of two: rows of the flattened (NC*D*H, W) input
// map to output row = input row / 2; one thread = two outputs (fwd) / a 2x4 input patch (bwd); 128-bit accesses.
__global__ void avgpool2x2_fwd_vec_kernel(const float* __restrict__ x, int64_t rows_out, int W, int wq_shift,
                                          float inv, int accumulate, float* __restrict__ y) {
  const int Wo = W >> 1;
  const int64_t total = rows_out << wq_shift;            // (W / 4) vectors per output row
  GRID_STRIDE(i, total) {
    const int64_t ro = i >> wq_shift;
    const int q = (int)(i - (ro << wq_shift));
    const float4 a = __ldg(reinterpret_cast<const float4*>(x + (2 * ro) * W) + q);
    const float4 b = __ldg(reinterpret_cast<const float4*>(x + (2 * ro + 1) * W) + q);
    float2 o = make_float2(((a.x + a.y) + (b.x + b.y)) * inv, ((a.z + a.w) + (b.z + b.w)) * inv);
    float2* dst = reinterpret_cast<float2*>(y + ro * Wo) + q;
    if (accumulate) { const float2 p = *dst; o.x += p.x; o.y += p.y; }
    *dst = o;
  }
}
__global__ void avgpool2x2_bwd_vec_kernel(const float* __restrict__ dy, int64_t rows_out, int W, int wq_shift,
                                          int accumulate, float* __restrict__ dx) {
  const int Wo = W >> 1;
  const int64_t total = rows_out << wq_shift;
  GRID_STRIDE(i, total) {
    const int64_t ro = i >> wq_shift;
    const int q = (int)(i - (ro << wq_shift));
    const float2 g = __ldg(reinterpret_cast<const float2*>(dy + ro * Wo) + q);
    float4 o = make_float4(g.x * 0.25f, g.x * 0.25f, g.y * 0.25f, g.y * 0.25f);
    float4* d0 = reinterpret_cast<float4*>(dx + (2 * ro) * W) + q;
    float4* d1 = reinterpret_cast<float4*>(dx + (2 * ro + 1) * W) + q;
    if (accumulate) {
      const float4 p0 = *d0, p1 = *d1;
      *d0 = make_float4(p0.x + o.x, p0.y + o.y, p0.z + o.z, p0.w + o.w);
      *d1 = make_float4(p1.x + o.x, p1.y + o.y, p1.z + o.z, p1.w + o.w);
    } else {
      *d0 = o;
      *d1 = o;
    }
  }
}
static int pool2x2_shift(const void* a, const void* b, int D, int H, int W, int pd, int ph, int pw) {
  if (pd != 1 || ph != 2 || pw != 2 || (W & 3) || (H & 1)) return -1;
  if ((reinterpret_cast<uintptr_t>(a) & 15) || (reinterpret_cast<uintptr_t>(b) & 15)) return -1;
  const int wq = W >> 2;
  if (wq & (wq - 1)) return -1;
  int sh = 0;
  while ((1 << sh) < wq) ++sh;
  (void)D;
  return sh;
}

__global__ void maxpool_fwd_kernel(const float* __restrict__ x, int64_t NC, int D, int H, int W, int pd, int ph, int pw,
                                   float* __restrict__ y) {
  const int Do = D / pd, Ho = H / ph, Wo = W / pw;
  const int64_t total = NC * Do * Ho * Wo;
  GRID_STRIDE(i, total) {
    const int xo = (int)(i % Wo);
    int64_t t = i / Wo;
    const int yo = (int)(t % Ho);
    t /= Ho;
    const int zo = (int)(t % Do);
    const int64_t nc = t / Do;
    const float* src = x + ((nc * D + (int64_t)zo * pd) * H + (int64_t)yo * ph) * W + (int64_t)xo * pw;
    float m = -INFINITY;
    for (int a = 0; a < pd; ++a)
      for (int b = 0; b < ph; ++b)
        for (int c = 0; c < pw; ++c) m = fmaxf(m, __ldg(src + ((int64_t)a * H + b) * W + c));
    y[i] = m;
  }
}

// gradient goes to the first maximum in scan order (what ATen's max_pool3d backward does)
__global__ void maxpool_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, int64_t NC, int D, int H,
                                   int W, int pd, int ph, int pw, float* __restrict__ dx) {
  const int Do = D / pd, Ho = H / ph, Wo = W / pw;
  const int64_t total = NC * Do * Ho * Wo;
  GRID_STRIDE(i, total) {
    const int xo = (int)(i % Wo);
    int64_t t = i / Wo;
    const int yo = (int)(t % Ho);
    t /= Ho;
    const int zo = (int)(t % Do);
    const int64_t nc = t / Do;
    const int64_t base = ((nc * D + (int64_t)zo * pd) * H + (int64_t)yo * ph) * W + (int64_t)xo * pw;
    float m = -INFINITY;
    int64_t arg = base;
    for (int a = 0; a < pd; ++a)
      for (int b = 0; b < ph; ++b)
        for (int c = 0; c < pw; ++c) {
          const int64_t o = base + ((int64_t)a * H + b) * W + c;
          const float v = __ldg(x + o);
          if (v > m) { m = v; arg = o; }
        }
    const float g = __ldg(dy + i);
    for (int a = 0; a < pd; ++a)
      for (int b = 0; b < ph; ++b)
        for (int c = 0; c < pw; ++c) {
          const int64_t o = base + ((int64_t)a * H + b) * W + c;
          dx[o] = (o == arg) ? g : 0.f;
        }
  }
}

// ---------------------------------------------------------------------------------------------- phi
__global__ void phi_fwd_kernel(const float* __restrict__ x, int B, int T, int C, int H, int W, float* __restrict__ y) {
  const int Ho = H / 2, Wo = W / 2;
  const int64_t total = (int64_t)B * C * T * Ho * Wo;
  GRID_STRIDE(i, total) {
    const int xo = (int)(i % Wo);
    int64_t t = i / Wo;
    const int yo = (int)(t % Ho);
    t /= Ho;
    const int tt = (int)(t % T);
    t /= T;
    const int c = (int)(t % C);
    const int b = (int)(t / C);
    const float* s = x + ((((int64_t)b * T + tt) * C + c) * H + 2 * yo) * W + 2 * xo;
    y[i] = (__ldg(s) + __ldg(s + 1) + __ldg(s + W) + __ldg(s + W + 1)) * 0.25f;
  }
}

__global__ void phi_bwd_kernel(const float* __restrict__ dy, int B, int T, int C, int H, int W, int accumulate,
                               float* __restrict__ dx) {
  const int Ho = H / 2, Wo = W / 2;
  const int64_t total = (int64_t)B * T * C * H * W;
  GRID_STRIDE(i, total) {
    const int xx = (int)(i % W);
    int64_t t = i / W;
    const int yy = (int)(t % H);
    t /= H;
    const int c = (int)(t % C);
    t /= C;
    const int tt = (int)(t % T);
    const int b = (int)(t / T);
    float v = 0.f;
    if ((yy >> 1) < Ho && (xx >> 1) < Wo)
      v = 0.25f * __ldg(dy + ((((int64_t)b * C + c) * T + tt) * Ho + (yy >> 1)) * Wo + (xx >> 1));
    dx[i] = accumulate ? dx[i] + v : v;
  }
}

// ---------------------------------------------------------------------------------------------- frames
__global__ void gather_frames_kernel(const float* __restrict__ x, const int64_t* __restrict__ idx, int B, int T, int k,
                                     int64_t fe, float* __restrict__ y) {
  const int64_t total = (int64_t)B * k * fe;
  GRID_STRIDE(i, total) {
    const int64_t e = i % fe;
    const int64_t t = i / fe;
    const int j = (int)(t % k);
    const int b = (int)(t / k);
    y[i] = __ldg(x + ((int64_t)b * T + idx[j]) * fe + e);
  }
}

__global__ void scatter_frames_kernel(const float* __restrict__ dy, const int64_t* __restrict__ idx, int B, int T,
                                      int k, int64_t fe, float* __restrict__ dx) {
  const int64_t total = (int64_t)B * k * fe;
  GRID_STRIDE(i, total) {
    const int64_t e = i % fe;
    const int64_t t = i / fe;
    const int j = (int)(t % k);
    const int b = (int)(t / k);
    dx[((int64_t)b * T + idx[j]) * fe + e] += __ldg(dy + i);   // indices are distinct (randperm)
  }
}

__global__ void permute_bctp_kernel(const float* __restrict__ x, int B, int C, int T, int64_t P,
                                    float* __restrict__ y) {
  const int64_t total = (int64_t)B * C * T * P;
  GRID_STRIDE(i, total) {   // i enumerates the destination (B,T,C,P)
    const int64_t p = i % P;
    int64_t t = i / P;
    const int c = (int)(t % C);
    t /= C;
    const int tt = (int)(t % T);
    const int b = (int)(t / T);
    y[i] = __ldg(x + (((int64_t)b * C + c) * T + tt) * P + p);
  }
}

struct Perm5 { int od[5]; int64_t istr[5]; };
__global__ void permute5_kernel(const float* __restrict__ x, Perm5 p, int64_t total, float* __restrict__ y) {
  GRID_STRIDE(i, total) {
    int64_t t = i, off = 0;
#pragma unroll
    for (int a = 4; a >= 0; --a) {
      const int c = (int)(t % p.od[a]);
      t /= p.od[a];
      off += (int64_t)c * p.istr[a];
    }
    y[i] = __ldg(x + off);
  }
}

// ---------------------------------------------------------------------------------------------- elementwise
__global__ void act_fwd_kernel(const float* __restrict__ x, int64_t n, int act, float* __restrict__ y) {
  GRID_STRIDE(i, n) {
    const float v = __ldg(x + i);
    y[i] = act == 1 ? fmaxf(v, 0.f) : (act == 2 ? tanhf(v) : v);
  }
}
__global__ void act_bwd_kernel(const float* __restrict__ ref, const float* __restrict__ dy, int64_t n, int act,
                               float* __restrict__ dx) {
  GRID_STRIDE(i, n) {
    const float r = __ldg(ref + i), g = __ldg(dy + i);
    dx[i] = act == 1 ? (r > 0.f ? g : 0.f) : (act == 2 ? g * (1.f - r * r) : g);
  }
}
__global__ void scale_residual_fwd_kernel(const float* __restrict__ o, const float* __restrict__ x,
                                          const float* __restrict__ gamma, int64_t n, float* __restrict__ y) {
  const float g = __ldg(gamma);
  GRID_STRIDE(i, n) y[i] = fmaf(g, __ldg(o + i), __ldg(x + i));
}
__global__ void scale_residual_bwd_kernel(const float* __restrict__ o, const float* __restrict__ dy,
                                          const float* __restrict__ gamma, int64_t n, float* __restrict__ do_,
                                          double* __restrict__ acc) {
  __shared__ double red[32];
  const float g = __ldg(gamma);
  double s = 0.0;
  GRID_STRIDE(i, n) {
    const float d = __ldg(dy + i);
    s += (double)d * (double)__ldg(o + i);
    do_[i] = g * d;
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(acc, s);
}
__global__ void double_to_float_kernel(const double* __restrict__ a, int n, int accumulate, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = accumulate ? out[i] + (float)a[i] : (float)a[i];
}
// one (channel, image range) per block: rows of P contiguous floats, 128-bit loads when P % 4 == 0, no per-element
// division; per-thread fp32 partials over one row, fp64 across rows and blocks
__global__ void channel_sum_kernel(const float* __restrict__ x, int N, int C, int64_t P, int64_t n_stride,
                                   int n_per_block, double* __restrict__ acc) {
  __shared__ double red[32];
  const int c = blockIdx.x;
  const int nb = blockIdx.y * n_per_block;
  int ne = nb + n_per_block;
  if (ne > N) ne = N;
  double s = 0.0;
  const bool vec = (P % 4 == 0) && (n_stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  if (vec && (P >> 2) <= blockDim.x && P >= blockDim.x) {
    // at most one 128-bit vector per thread and row: fix the column, walk the rows four at a time so that four
    // independent loads are in flight per thread (one load per loop iteration left this kernel latency-bound)
    const int P4 = (int)(P >> 2);
    const int rpp = blockDim.x / P4;
    const int r = threadIdx.x / P4, q = threadIdx.x - r * P4;
    if (r < rpp) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      const float* base = x + (int64_t)c * P;
      int n = nb + r;
      for (; n + 3 * rpp < ne; n += 4 * rpp) {
        const float4 v0 = __ldg(reinterpret_cast<const float4*>(base + (int64_t)n * n_stride) + q);
        const float4 v1 = __ldg(reinterpret_cast<const float4*>(base + (int64_t)(n + rpp) * n_stride) + q);
        const float4 v2 = __ldg(reinterpret_cast<const float4*>(base + (int64_t)(n + 2 * rpp) * n_stride) + q);
        const float4 v3 = __ldg(reinterpret_cast<const float4*>(base + (int64_t)(n + 3 * rpp) * n_stride) + q);
        a0 += (v0.x + v0.y) + (v0.z + v0.w);
        a1 += (v1.x + v1.y) + (v1.z + v1.w);
        a2 += (v2.x + v2.y) + (v2.z + v2.w);
        a3 += (v3.x + v3.y) + (v3.z + v3.w);
      }
      for (; n < ne; n += rpp) {
        const float4 v0 = __ldg(reinterpret_cast<const float4*>(base + (int64_t)n * n_stride) + q);
        a0 += (v0.x + v0.y) + (v0.z + v0.w);
      }
      s = ((double)a0 + (double)a1) + ((double)a2 + (double)a3);
    }
  } else if (P >= blockDim.x) {
    for (int n = nb; n < ne; ++n) {
      const float* row = x + (int64_t)n * n_stride + (int64_t)c * P;
      float part = 0.f;
      if (vec) {
        const float4* r4 = reinterpret_cast<const float4*>(row);
        const int64_t P4 = P >> 2;
        for (int64_t i = threadIdx.x; i < P4; i += blockDim.x) {
          const float4 v = __ldg(r4 + i);
          part += (v.x + v.y) + (v.z + v.w);
        }
      } else {
        for (int64_t i = threadIdx.x; i < P; i += blockDim.x) part += __ldg(row + i);
      }
      s += (double)part;
    }
  } else {
    // short rows (P < blockDim): lay the threads out as (rows per pass) x P
    const int Pi = (int)P;
    const int rpp = blockDim.x / Pi;
    const int r = threadIdx.x / Pi, pp = threadIdx.x - r * Pi;
    if (r < rpp) {
      float part = 0.f;
      for (int n = nb + r; n < ne; n += rpp) part += __ldg(x + (int64_t)n * n_stride + (int64_t)c * P + pp);
      s = (double)part;
    }
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(acc + c, s);
}
// copy (or zero-fill, src == nullptr) up to kMax tensors into their slices of a flat arena in ONE launch
struct GatherP {
  static constexpr int kMax = 96;
  const float* src[kMax];
  int64_t off[kMax];
  int64_t cnt[kMax];
  int n;
};
__global__ void gather_flat_kernel(const GatherP gp, float* __restrict__ flat) {
  const int t = blockIdx.y;
  const float* __restrict__ src = gp.src[t];
  float* dst = flat + gp.off[t];
  const int64_t n = gp.cnt[t];
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = src ? __ldg(src + i) : 0.f;
}
__global__ void axpby_kernel(const float* __restrict__ x, float a, float b, int64_t n, float* __restrict__ y) {
  GRID_STRIDE(i, n) y[i] = b == 0.f ? a * __ldg(x + i) : fmaf(a, __ldg(x + i), b * y[i]);
}
// Class ids index embedding tables (nn.Embedding raises on an id outside [0, rows); a raw pointer cannot): an
// out-of-range id is counted in g_index_errors (dvd_index_errors reads it) and clamped so no access leaves the table.
__device__ unsigned int g_index_errors = 0;
__device__ __forceinline__ int64_t checked_index(int64_t i, int rows, bool count) {
  if (i < 0 || i >= rows) {
    if (count) atomicAdd(&g_index_errors, 1u);
    return i < 0 ? 0 : rows - 1;
  }
  return i;
}

__global__ void embedding_fwd_kernel(const float* __restrict__ w, const int64_t* __restrict__ idx, int n, int dim,
                                     int rows, float* __restrict__ y) {
  GRID_STRIDE(i, (int64_t)n * dim) {
    const int r = (int)(i / dim), c = (int)(i % dim);
    y[i] = __ldg(w + checked_index(idx[r], rows, c == 0) * dim + c);       // one count per looked-up row
  }
}
__global__ void embedding_bwd_kernel(const float* __restrict__ dy, const int64_t* __restrict__ idx, int n, int dim,
                                     int rows, float* __restrict__ dw) {
  GRID_STRIDE(i, (int64_t)n * dim) {
    const int r = (int)(i / dim), c = (int)(i % dim);
    atomicAdd(dw + checked_index(idx[r], rows, false) * dim + c, __ldg(dy + i));       // counted by the forward
  }
}

// ---------------------------------------------------------------------------------------------- D head
__global__ void dhead_fwd_kernel(const float* __restrict__ x, int C, int HW, int T, const float* __restrict__ wl,
                                 const float* __restrict__ sl, const float* __restrict__ bl,
                                 const float* __restrict__ emb, const float* __restrict__ se,
                                 const int64_t* __restrict__ cls, int n_class, float* __restrict__ feat,
                                 float* __restrict__ out) {
  __shared__ float red[32];
  const int n = blockIdx.x;
  const float isl = 1.f / __ldg(sl), ise = 1.f / __ldg(se);
  int64_t cl = cls[n / T];
  if (cl < 0 || cl >= n_class) {          // one count per frame, not per thread
    if (threadIdx.x == 0) atomicAdd(&g_index_errors, 1u);
    cl = cl < 0 ? 0 : n_class - 1;
  }
  const float* e = emb + cl * C;
  float acc = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float* p = x + ((int64_t)n * C + c) * HW;
    float f = 0.f;
    for (int q = 0; q < HW; ++q) f += fmaxf(__ldg(p + q), 0.f);
    feat[(int64_t)n * C + c] = f;
    acc += f * (__ldg(wl + c) * isl) + f * (__ldg(e + c) * ise);
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) out[n] = acc + __ldg(bl);
}

__global__ void dhead_bwd_kernel(const float* __restrict__ x, const float* __restrict__ feat,
                                 const float* __restrict__ dout, int C, int HW, int T, const float* __restrict__ wl,
                                 const float* __restrict__ sl, const float* __restrict__ emb,
                                 const float* __restrict__ se, const int64_t* __restrict__ cls, int n_class,
                                 float* __restrict__ dx, float* __restrict__ dwl, float* __restrict__ db,
                                 float* __restrict__ demb) {
  const int n = blockIdx.x;
  const float isl = 1.f / __ldg(sl), ise = 1.f / __ldg(se);
  int64_t cl = cls[n / T];
  if (cl < 0 || cl >= n_class) cl = cl < 0 ? 0 : n_class - 1;      // counted by the forward
  const float* e = emb + cl * C;
  const float g = __ldg(dout + n);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float dfeat = g * (__ldg(wl + c) * isl + __ldg(e + c) * ise);
    const float* p = x + ((int64_t)n * C + c) * HW;
    float* q = dx + ((int64_t)n * C + c) * HW;
    for (int r = 0; r < HW; ++r) q[r] = __ldg(p + r) > 0.f ? dfeat : 0.f;
    const float f = __ldg(feat + (int64_t)n * C + c);
    atomicAdd(dwl + c, g * f);
    atomicAdd(demb + cl * C + c, g * f);
  }
  if (threadIdx.x == 0) atomicAdd(db, g);
}

// ---------------------------------------------------------------------------------------------- losses
__global__ void gan_loss_fwd_kernel(const float* __restrict__ x, int n, float sign, int hinge, int accumulate,
                                    float* __restrict__ loss) {
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = sign * __ldg(x + i);
    s += hinge ? fmaxf(1.f + v, 0.f) : v;
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) loss[0] = (accumulate ? loss[0] : 0.f) + s / n;
}
__global__ void gan_loss_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gout, int n, float sign,
                                    int hinge, float* __restrict__ dx) {
  const float g = __ldg(gout) / n;
  GRID_STRIDE(i, n) {
    const float v = sign * __ldg(x + i);
    dx[i] = (!hinge || 1.f + v > 0.f) ? sign * g : 0.f;
  }
}

// ---------------------------------------------------------------------------------------------- Adam
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t n, float beta1, float beta2, float eps, float step_size,
                            float inv_bc2_sqrt, float grad_scale) {
  GRID_STRIDE(i, n) {
    const float gr = __ldg(g + i) * grad_scale;
    // torch.optim.Adam: exp_avg.lerp_(grad, 1-beta1); exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1-beta2)
    const float w = 1.f - beta1;
    const float m0 = m[i];
    const float mn = w < 0.5f ? m0 + w * (gr - m0) : gr - (gr - m0) * (1.f - w);
    const float vn = v[i] * beta2 + (1.f - beta2) * gr * gr;
    m[i] = mn;
    v[i] = vn;
    const float denom = sqrtf(vn) * inv_bc2_sqrt + eps;
    p[i] = p[i] - step_size * (mn / denom);
  }
}

}  // namespace dvd

using namespace dvd;

#define POOL_ARGS_OK (NC > 0 && D > 0 && H > 0 && W > 0 && pd >= 1 && ph >= 1 && pw >= 1)

extern "C" int dvd_avgpool_fwd(const float* x, int64_t NC, int D, int H, int W, int pd, int ph, int pw, float scale,
                               int accumulate, float* y, void* stream) {
  dvd::ProfScope _ps(3, "avgpool_fwd", dvd::as_stream(stream));
  DVD_CHECK_ARG(x && y && POOL_ARGS_OK);
  const int64_t total = NC * (D / pd) * (H / ph) * (W / pw);
  DVD_CHECK_ARG(total > 0);
  const int sh = pool2x2_shift(x, y, D, H, W, pd, ph, pw);
  if (sh >= 0) {
    const int64_t rows_out = NC * D * (H / 2);
    avgpool2x2_fwd_vec_kernel<<<ew_blocks(rows_out << sh, 2), 256, 0, as_stream(stream)>>>(x, rows_out, W, sh, scale * 0.25f,
                                                                                          accumulate, y);
  } else {
    avgpool_fwd_kernel<<<ew_blocks(total, 2), 256, 0, as_stream(stream)>>>(x, NC, D, H, W, pd, ph, pw, scale, accumulate, y);
  }
  DVD_LAUNCH_CHECK();
  return 0;
}
extern "C" int dvd_avgpool_bwd(const float* dy, int64_t NC, int D, int H, int W, int pd, int ph, int pw,
                               int accumulate, float* dx, void* stream) {
  dvd::ProfScope _ps(3, "avgpool_bwd", dvd::as_stream(stream));
  DVD_CHECK_ARG(dy && dx && POOL_ARGS_OK);
  const int sh = pool2x2_shift(dy, dx, D, H, W, pd, ph, pw);
  if (sh >= 0) {
    const int64_t rows_out = NC * D * (H / 2);
    avgpool2x2_bwd_vec_kernel<<<ew_blocks(rows_out << sh, 2), 256, 0, as_stream(stream)>>>(dy, rows_out, W, sh, accumulate, dx);
  } else {
    avgpool_bwd_kernel<<<ew_blocks(NC * D * H * W, 4), 256, 0, as_stream(stream)>>>(dy, NC, D, H, W, pd, ph, pw,
                                                                                    accumulate, dx);
  }
  DVD_LAUNCH_CHECK();
  return 0;
}
extern "C" int dvd_maxpool_fwd(const float* x, int64_t NC, int D, int H, int W, int pd, int ph, int pw, float* y,
                               void* stream) {
  dvd::ProfScope _ps(3, "maxpool_fwd", dvd::as_stream(stream));
  DVD_CHECK_ARG(x && y && POOL_ARGS_OK);
  DVD_CHECK_ARG(D % pd == 0 && H % ph == 0 && W % pw == 0);
  const int64_t total = NC * (D / pd) * (H / ph) * (W / pw);
  maxpool_fwd_kernel<<<ew_blocks(total, 2), 256, 0, as_stream(stream)>>>(x, NC, D, H, W, pd, ph, pw, y);
  DVD_LAUNCH_CHECK();
  return 0;
}
extern "C" int dvd_maxpool_bwd(const float* x, const float* dy, int64_t NC, int D, int H, int W, int pd, int ph, int pw,
                               float* dx, void* stream) {
  dvd::ProfScope _ps(3, "maxpool_bwd", dvd::as_stream(stream));
  DVD_CHECK_ARG(x && dy && dx && POOL_ARGS_OK);
  DVD_CHECK_ARG(D % pd == 0 && H % ph == 0 && W % pw == 0);
  const int64_t total = NC * (D / pd) * (H / ph) * (W / pw);
  maxpool_bwd_kernel<<<ew_blocks(total, 2), 256, 0, as_stream(stream)>>>(x, dy, NC, D, H, W, pd, ph, pw, dx);
  DVD_LAUNCH_CHECK();
  return 0;
}
extern "C" int dvd_phi_fwd(const float* x, int B, int T, int C, int H, int W, float* y, void* stream) {
  dvd::ProfScope _ps(3, "phi_fwd", dvd::as_stream(stream));
  DVD_CHECK_ARG(x && y && B > 0 && T > 0 && C > 0 && H > 1 && W > 1);
  const int64_t total = (int64_t)B * C * T * (H / 2) * (W / 2);
  phi_fwd_kernel<<<ew_blocks(total, 2), 256, 0, as_stream(stream)>>>(x, B, T, C, H, W, y);
  DVD_LAUNCH_CHECK();
  return 0;
}
extern "C" int dvd_phi_bwd(const float* dy, int B, int T, int C, int H, int W, int accumulate, float* dx,
                           void* stream) {
  dvd::ProfScope _ps(3, "phi_bwd", dvd::as_stream(stream));
  DVD_CHECK_ARG(dy && dx && B > 0 && T > 0 && C > 0 && H > 1 && W > 1);
  phi_bwd_kernel<<<ew_blocks((int64_t)B * T * C * H * W, 4), 256, 0, as_stream(stream)>>>(dy, B, T, C, H, W, accumulate,
                                                                                          dx);
  DVD_LAUNCH_CHECK();
  return 0;
}
extern "C" int dvd_gather_frames_fwd(const float* x, const int64_t* idx, int B, int T, int k, int64_t frame_elems,
                                     float* y, void* stream) {
  dvd::ProfScope _ps(3, "gather_frames_fwd", dvd::as_stream(stream));
  DVD_CHECK_ARG(x && idx && y && B > 0 && T > 0 && k > 0 && frame_elems > 0);
  gather_frames_kernel<<<ew_blocks((int64_t)B * k * frame_elems, 4), 256, 0, as_stream(stream)>>>(x, idx, B, T, k,
                                                                                                  frame_elems, y);
  DVD_LAUNCH_CHECK();
  return 0;
}
extern "C" int dvd_gather_frames_bwd(const float* dy, const int64_t* idx, int B, int T, int k, int64_t frame_elems,
                                     int accumulate, float* dx, void* stream) {
  dvd::ProfScope _ps(3, "gather_frames_bwd", dvd::as_stream(stream));
  DVD_CHECK_ARG(dy && idx && dx && B > 0 && T > 0 && k > 0 && frame_elems > 0);
  cudaStream_t st = as_stream(stream);
  if (!accumulate) DVD_CUDA(cudaMemsetAsync(dx, 0, sizeof(float) * (size_t)B * T * frame_elems, st));
  scatter_frames_kernel<<<ew_blocks((int64_t)B * k * frame_elems, 4), 256, 0, st>>>(dy, idx, B, T, k, frame_elems, dx);
  DVD_LAUNCH_CHECK();
  return 0;
}
extern "C" int dvd_permute_bctp(const float* x, int B, int C, int T, int64_t P, float* y, void* stream) {
  dvd::ProfScope _ps(3, "permute_bctp", dvd::as_stream(stream));
  DVD_CHECK_ARG(x && y && B > 0 && C > 0 && T > 0 && P > 0);
  permute_bctp_kernel<<<ew_blocks((int64_t)B * C * T * P, 4), 256, 0, as_stream(stream)>>>(x, B, C, T, P, y);
  DVD_LAUNCH_CHECK();
  return 0;
}
extern "C" int dvd_permute5(const float* x, const int* dims, const int* perm, float* y, void* stream) {
  dvd::ProfScope _ps(3, "permute5", dvd::as_stream(stream));
  DVD_CHECK_ARG(x && dims && perm && y);
  int64_t str[5];
  int64_t total = 1;
  for (int a = 4; a >= 0; --a) {
    DVD_CHECK_ARG(dims[a] > 0 && perm[a] >= 0 && perm[a] < 5);
    str[a] = total;
    total *= dims[a];
  }
  Perm5 p;
  for (int a = 0; a < 5; ++a) {
    p.od[a] = dims[perm[a]];
    p.istr[a] = str[perm[a]];
  }
  permute5_kernel<<<ew_blocks(total, 4), 256, 0, as_stream(stream)>>>(x, p, total, y);
  DVD_LAUNCH_CHECK();
  return 0;
}
extern "C" int dvd_act_fwd(const float* x, int64_t n, int act, float* y, void* stream) {
  dvd::ProfScope _ps(3, "act_fwd", dvd::as_stream(stream));
  DVD_CHECK_ARG(x && y && n > 0);
  act_fwd_kernel<<<ew_blocks(n), 256, 0, as_stream(stream)>>>(x, n, act, y);
  DVD_LAUNCH_CHECK();
  return 0;
}
extern "C" int dvd_act_bwd(const float* ref, const float* dy, int64_t n, int act, float* dx, void* stream) {
  dvd::ProfScope _ps(3, "act_bwd", dvd::as_stream(stream));
  DVD_CHECK_ARG(ref && dy && dx && n > 0);
  act_bwd_kernel<<<ew_blocks(n), 256, 0, as_stream(stream)>>>(ref, dy, n, act, dx);
  DVD_LAUNCH_CHECK();
  return 0;
}
extern "C" int dvd_scale_residual_fwd(const float* o, const float* x, const float* gamma, int64_t n, float* y,
                                      void* stream) {
  dvd::ProfScope _ps(3, "scale_residual_fwd", dvd::as_stream(stream));
  DVD_CHECK_ARG(o && x && gamma && y && n > 0);
  scale_residual_fwd_kernel<<<ew_blocks(n), 256, 0, as_stream(stream)>>>(o, x, gamma, n, y);
  DVD_LAUNCH_CHECK();
  return 0;
}
extern "C" int dvd_scale_residual_bwd(const float* o, const float* dy, const float* gamma, int64_t n, float* do_,
                                      float* dgamma, void* scratch, void* stream) {
  dvd::ProfScope _ps(3, "scale_residual_bwd", dvd::as_stream(stream));
  DVD_CHECK_ARG(o && dy && gamma && do_ && dgamma && scratch && n > 0);
  cudaStream_t st = as_stream(stream);
  double* acc = reinterpret_cast<double*>(scratch);
  DVD_CUDA(cudaMemsetAsync(acc, 0, sizeof(double), st));
  scale_residual_bwd_kernel<<<ew_blocks(n, 8), 256, 0, st>>>(o, dy, gamma, n, do_, acc);
  DVD_LAUNCH_CHECK();
  double_to_float_kernel<<<1, 32, 0, st>>>(acc, 1, 0, dgamma);
  DVD_LAUNCH_CHECK();
  return 0;
}
extern "C" int dvd_channel_sum(const float* x, int N, int C, int64_t P, int64_t n_stride, int accumulate, float* out,
                               void* scratch, void* stream) {
  dvd::ProfScope _ps(3, "channel_sum", dvd::as_stream(stream));
  DVD_CHECK_ARG(x && out && scratch && N > 0 && C > 0 && P > 0);
  cudaStream_t st = as_stream(stream);
  double* acc = reinterpret_cast<double*>(scratch);
  DVD_CUDA(cudaMemsetAsync(acc, 0, sizeof(double) * C, st));
  int splits = ceil_div(4 * num_sms(), C);
  if (splits > N) splits = N;
  if (splits < 1) splits = 1;
  const int npb = ceil_div(N, splits);
  splits = ceil_div(N, npb);
  channel_sum_kernel<<<dim3(C, splits), 256, 0, st>>>(x, N, C, P, n_stride, npb, acc);
  DVD_LAUNCH_CHECK();
  double_to_float_kernel<<<ceil_div(C, 128), 128, 0, st>>>(acc, C, accumulate, out);
  DVD_LAUNCH_CHECK();
  return 0;
}
extern "C" int dvd_gather_flat(const float* const* srcs, const int64_t* offs, const int64_t* counts, int n, float* flat,
                               void* stream) {
  dvd::ProfScope _ps(3, "gather_flat", dvd::as_stream(stream));
  DVD_CHECK_ARG(srcs && offs && counts && flat && n >= 0);
  cudaStream_t st = as_stream(stream);
  for (int i0 = 0; i0 < n; i0 += GatherP::kMax) {
    GatherP gp;
    gp.n = n - i0 < GatherP::kMax ? n - i0 : GatherP::kMax;
    int64_t longest = 1;
    for (int i = 0; i < gp.n; ++i) {
      DVD_CHECK_ARG(counts[i0 + i] >= 0 && offs[i0 + i] >= 0);
      gp.src[i] = srcs[i0 + i];
      gp.off[i] = offs[i0 + i];
      gp.cnt[i] = counts[i0 + i];
      if (gp.cnt[i] > longest) longest = gp.cnt[i];
    }
    int bx = (int)ceil_div<int64_t>(longest, 256 * 8);
    if (bx > 2 * num_sms()) bx = 2 * num_sms();
    gather_flat_kernel<<<dim3(bx, gp.n), 256, 0, st>>>(gp, flat);
    DVD_LAUNCH_CHECK();
  }
  return 0;
}
extern "C" int dvd_axpby(const float* x, float a, float b, int64_t n, float* y, void* stream) {
  dvd::ProfScope _ps(3, "axpby", dvd::as_stream(stream));
  DVD_CHECK_ARG(x && y && n > 0);
  axpby_kernel<<<ew_blocks(n), 256, 0, as_stream(stream)>>>(x, a, b, n, y);
  DVD_LAUNCH_CHECK();
  return 0;
}
extern "C" int dvd_embedding_fwd(const float* w, const int64_t* idx, int n, int dim, int rows, float* y,
                                 void* stream) {
  dvd::ProfScope _ps(3, "embedding_fwd", dvd::as_stream(stream));
  DVD_CHECK_ARG(w && idx && y && n > 0 && dim > 0 && rows > 0);
  embedding_fwd_kernel<<<ew_blocks((int64_t)n * dim, 1), 256, 0, as_stream(stream)>>>(w, idx, n, dim, rows, y);
  DVD_LAUNCH_CHECK();
  return 0;
}
extern "C" int dvd_embedding_bwd(const float* dy, const int64_t* idx, int n, int dim, int rows, float* dw,
                                 void* stream) {
  dvd::ProfScope _ps(3, "embedding_bwd", dvd::as_stream(stream));
  DVD_CHECK_ARG(dy && idx && dw && n > 0 && dim > 0 && rows > 0);
  embedding_bwd_kernel<<<ew_blocks((int64_t)n * dim, 1), 256, 0, as_stream(stream)>>>(dy, idx, n, dim, rows, dw);
  DVD_LAUNCH_CHECK();
  return 0;
}
extern "C" int dvd_dhead_fwd(const float* x, int N, int C, int HW, int T, int n_class, const float* w_lin,
                             const float* sigma_l, const float* b_lin, const float* emb, const float* sigma_e,
                             const int64_t* class_id, float* feat, float* out, void* stream) {
  dvd::ProfScope _ps(3, "dhead_fwd", dvd::as_stream(stream));
  DVD_CHECK_ARG(x && w_lin && sigma_l && b_lin && emb && sigma_e && class_id && feat && out);
  DVD_CHECK_ARG(N > 0 && C > 0 && HW > 0 && T > 0 && n_class > 0);
  dhead_fwd_kernel<<<N, 256, 0, as_stream(stream)>>>(x, C, HW, T, w_lin, sigma_l, b_lin, emb, sigma_e, class_id,
                                                     n_class, feat, out);
  DVD_LAUNCH_CHECK();
  return 0;
}
extern "C" int dvd_dhead_bwd(const float* x, const float* feat, const float* dout, int N, int C, int HW, int T,
                             int n_class, const float* w_lin, const float* sigma_l, const float* emb,
                             const float* sigma_e, const int64_t* class_id, float* dx, float* dwl, float* db,
                             float* demb, void* stream) {
  dvd::ProfScope _ps(3, "dhead_bwd", dvd::as_stream(stream));
  DVD_CHECK_ARG(x && feat && dout && w_lin && sigma_l && emb && sigma_e && class_id && dx && dwl && db && demb);
  DVD_CHECK_ARG(N > 0 && C > 0 && HW > 0 && T > 0 && n_class > 0);
  cudaStream_t st = as_stream(stream);
  DVD_CUDA(cudaMemsetAsync(dwl, 0, sizeof(float) * C, st));
  DVD_CUDA(cudaMemsetAsync(db, 0, sizeof(float), st));
  DVD_CUDA(cudaMemsetAsync(demb, 0, sizeof(float) * (size_t)n_class * C, st));
  dhead_bwd_kernel<<<N, 256, 0, st>>>(x, feat, dout, C, HW, T, w_lin, sigma_l, emb, sigma_e, class_id, n_class, dx,
                                      dwl, db, demb);
  DVD_LAUNCH_CHECK();
  return 0;
}
// Reads (and optionally clears) the count of out-of-range class ids seen by the embedding / head kernels of the
// current device since the last clear.  The ONE entry point that synchronises: it waits for `stream`.
extern "C" int dvd_index_errors(unsigned int* count, int reset, void* stream) {
  DVD_CHECK_ARG(count != nullptr);
  cudaStream_t st = as_stream(stream);
  DVD_CUDA(cudaMemcpyFromSymbolAsync(count, g_index_errors, sizeof(unsigned int), 0, cudaMemcpyDeviceToHost, st));
  if (reset) {
    const unsigned int zero = 0;
    DVD_CUDA(cudaMemcpyToSymbolAsync(g_index_errors, &zero, sizeof(unsigned int), 0, cudaMemcpyHostToDevice, st));
  }
  DVD_CUDA(cudaStreamSynchronize(st));
  return 0;
}
extern "C" int dvd_gan_loss_fwd(const float* x, int n, float sign, int hinge, int accumulate, float* loss,
                                void* stream) {
  dvd::ProfScope _ps(3, "gan_loss_fwd", dvd::as_stream(stream));
  DVD_CHECK_ARG(x && loss && n > 0);
  gan_loss_fwd_kernel<<<1, 1024, 0, as_stream(stream)>>>(x, n, sign, hinge, accumulate, loss);
  DVD_LAUNCH_CHECK();
  return 0;
}
extern "C" int dvd_gan_loss_bwd(const float* x, const float* gout, int n, float sign, int hinge, float* dx,
                                void* stream) {
  dvd::ProfScope _ps(3, "gan_loss_bwd", dvd::as_stream(stream));
  DVD_CHECK_ARG(x && gout && dx && n > 0);
  gan_loss_bwd_kernel<<<ew_blocks(n, 1), 256, 0, as_stream(stream)>>>(x, gout, n, sign, hinge, dx);
  DVD_LAUNCH_CHECK();
  return 0;
}
extern "C" int dvd_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                             float beta2, float eps, int step, float grad_scale, void* stream) {
  dvd::ProfScope _ps(3, "adam_step", dvd::as_stream(stream));
  DVD_CHECK_ARG(p && g && m && v && n > 0 && step >= 1);
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1);
  const float inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  adam_kernel<<<ew_blocks(n, 4), 256, 0, as_stream(stream)>>>(p, g, m, v, n, beta1, beta2, eps, step_size, inv_bc2_sqrt,
                                                              grad_scale);
  DVD_LAUNCH_CHECK();
  return 0;
}
