// Spectral norm (power iteration + backward) and class-conditional BatchNorm (statistics, fused
// affine + ReLU + nearest-upsample apply, backward).  All HBM-bound: coalesced along the innermost
// dimension, warp-shuffle reductions, fp64 accumulation of the statistics.
#include "common.cuh"

namespace dvd {

// ------------------------------------------------------------------------------------------------
// spectral norm
// ------------------------------------------------------------------------------------------------

// vraw[j] = sum_i W[i][j] * u[i]; one block = 32 columns x 8 row groups
__global__ void sn_colsum_kernel(const float* __restrict__ w, const float* __restrict__ u, int rows, int cols,
                                 float* __restrict__ vraw) {
  __shared__ float part[8][33];
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + lane;
  float acc = 0.f;
  if (j < cols)
    for (int i = grp; i < rows; i += 8) acc = fmaf(__ldg(w + (int64_t)i * cols + j), __ldg(u + i), acc);
  part[grp][lane] = acc;
  __syncthreads();
  if (grp == 0 && j < cols) {
    float s = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) s += part[g][lane];
    vraw[j] = s;
  }
}

// s[i] = sum_j W[i][j] * vraw[j]; one warp per row
__global__ void sn_rowdot_kernel(const float* __restrict__ w, const float* __restrict__ vraw, int rows, int cols,
                                 float* __restrict__ s) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= rows) return;
  float acc = 0.f;
  for (int j = lane; j < cols; j += 32) acc = fmaf(__ldg(w + (int64_t)i * cols + j), __ldg(vraw + j), acc);
  acc = warp_sum(acc);
  if (lane == 0) s[i] = acc;
}

// v = vraw/(|vraw|+eps); uraw = s/(|vraw|+eps) (= W v); u = uraw/(|uraw|+eps); sigma = u . uraw
__global__ void sn_finalize_kernel(const float* __restrict__ vraw, const float* __restrict__ s, int rows, int cols,
                                   float* __restrict__ u, float* __restrict__ v, float* __restrict__ sigma) {
  __shared__ float red[32];
  const float eps = 1e-12f;
  float a = 0.f;
  for (int j = threadIdx.x; j < cols; j += blockDim.x) a = fmaf(vraw[j], vraw[j], a);
  const float nv = sqrtf(block_sum(a, red));
  const float inv_v = 1.f / (nv + eps);
  for (int j = threadIdx.x; j < cols; j += blockDim.x) v[j] = vraw[j] * inv_v;
  float b = 0.f;
  for (int i = threadIdx.x; i < rows; i += blockDim.x) {
    const float ur = s[i] * inv_v;
    b = fmaf(ur, ur, b);
  }
  const float nu = sqrtf(block_sum(b, red));
  const float inv_u = 1.f / (nu + eps);
  float c = 0.f;
  for (int i = threadIdx.x; i < rows; i += blockDim.x) {
    const float ur = s[i] * inv_v;
    const float un = ur * inv_u;
    u[i] = un;
    c = fmaf(un, ur, c);
  }
  const float sg = block_sum(c, red);
  if (threadIdx.x == 0) sigma[0] = sg;
}

__global__ void sn_dot_kernel(const float* __restrict__ g, const float* __restrict__ w, int64_t n, double* out) {
  __shared__ double red[32];
  double acc = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    acc += (double)__ldg(g + i) * (double)__ldg(w + i);
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(out, acc);
}

__global__ void sn_bwd_kernel(const float* __restrict__ g, const float* __restrict__ u, const float* __restrict__ v,
                              const float* __restrict__ sigma, const double* __restrict__ dot, int rows, int cols,
                              float* __restrict__ dw, int accumulate) {
  const float sg = __ldg(sigma);
  const float inv = 1.f / sg;
  const float coef = (float)(dot[0] / ((double)sg * (double)sg));   // <G, W_bar> / sigma^2
  const int64_t n = (int64_t)rows * cols;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i - (int64_t)r * cols);
    const float val = __ldg(g + i) * inv - coef * __ldg(u + r) * __ldg(v + c);
    dw[i] = accumulate ? dw[i] + val : val;
  }
}

// ------------------------------------------------------------------------------------------------
// batch statistics
// ------------------------------------------------------------------------------------------------
__global__ void bn_partial_kernel(const float* __restrict__ x, int N, int C, int HW, int n_per_block,
                                  double* __restrict__ acc /* [2C] */) {
  __shared__ double red[32];
  const int c = blockIdx.x;
  const int nb = blockIdx.y * n_per_block;
  int ne = nb + n_per_block;
  if (ne > N) ne = N;
  double s = 0.0, ss = 0.0;
  const int64_t cnt = (int64_t)(ne - nb) * HW;
  for (int64_t i = threadIdx.x; i < cnt; i += blockDim.x) {
    const int n = nb + (int)(i / HW);
    const int p = (int)(i % HW);
    const float v = __ldg(x + ((int64_t)n * C + c) * HW + p);
    s += v;
    ss += (double)v * v;
  }
  s = block_sum(s, red);
  ss = block_sum(ss, red);
  if (threadIdx.x == 0) {
    atomicAdd(acc + c, s);
    atomicAdd(acc + C + c, ss);
  }
}

__global__ void bn_finalize_kernel(const double* __restrict__ acc, int C, double cnt, int training, float momentum,
                                   float eps, float* __restrict__ rm, float* __restrict__ rv,
                                   long long* __restrict__ nbt, float* __restrict__ mean, float* __restrict__ rstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && training && nbt) nbt[0] += 1;
  if (c >= C) return;
  if (training) {
    const double m = acc[c] / cnt;
    double var = acc[C + c] / cnt - m * m;
    if (var < 0.0) var = 0.0;
    mean[c] = (float)m;
    rstd[c] = (float)(1.0 / sqrt(var + (double)eps));
    if (rm) rm[c] = (1.f - momentum) * rm[c] + momentum * (float)m;
    if (rv) {
      const double unb = cnt > 1.0 ? var * cnt / (cnt - 1.0) : var;
      rv[c] = (1.f - momentum) * rv[c] + momentum * (float)unb;
    }
  } else {
    mean[c] = rm[c];
    rstd[c] = 1.f / sqrtf(rv[c] + eps);
  }
}

// y[n][c][yo][xo] = act(g * (x - mean) * rstd + b), nearest-upsampled by 2^up
__global__ void cbn_apply_kernel(const float* __restrict__ x, const float* __restrict__ gb, int R,
                                 const float* __restrict__ mean, const float* __restrict__ rstd, int N, int C, int H,
                                 int W, int relu, int up, float* __restrict__ y) {
  const int Ho = H << up, Wo = W << up;
  const int64_t total = (int64_t)N * C * Ho * Wo;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int xo = (int)(i % Wo);
    int64_t t = i / Wo;
    const int yo = (int)(t % Ho);
    t /= Ho;
    const int c = (int)(t % C);
    const int n = (int)(t / C);
    const int r = n % R;
    const float g = __ldg(gb + (int64_t)r * 2 * C + c), b = __ldg(gb + (int64_t)r * 2 * C + C + c);
    const float xv = __ldg(x + (((int64_t)n * C + c) * H + (yo >> up)) * W + (xo >> up));
    float v = g * ((xv - __ldg(mean + c)) * __ldg(rstd + c)) + b;
    if (relu) v = fmaxf(v, 0.f);
    y[i] = v;
  }
}

// per-plane sums: dgamma[n][c] = sum g * xhat, dbeta[n][c] = sum g, with g = act'(pre) * sum_{window} dy
__global__ void cbn_bwd_plane_kernel(const float* __restrict__ x, const float* __restrict__ gb, int R,
                                     const float* __restrict__ mean, const float* __restrict__ rstd,
                                     const float* __restrict__ dy, int N, int C, int H, int W, int relu, int up,
                                     float* __restrict__ dgb) {
  const int lane = threadIdx.x & 31;
  const int64_t plane = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (plane >= (int64_t)N * C) return;
  const int c = (int)(plane % C);
  const int n = (int)(plane / C);
  const int r = n % R;
  const float g = __ldg(gb + (int64_t)r * 2 * C + c), b = __ldg(gb + (int64_t)r * 2 * C + C + c);
  const float mu = __ldg(mean + c), rs = __ldg(rstd + c);
  const int HW = H * W, Wo = W << up;
  const float* xp = x + plane * HW;
  const float* dp = dy + plane * ((int64_t)HW << (2 * up));
  float sg = 0.f, sgx = 0.f;
  for (int p = lane; p < HW; p += 32) {
    const float xh = (__ldg(xp + p) - mu) * rs;
    float d;
    if (up) {
      const int yy = p / W, xx = p - yy * W;
      const float* q = dp + (int64_t)(2 * yy) * Wo + 2 * xx;
      d = __ldg(q) + __ldg(q + 1) + __ldg(q + Wo) + __ldg(q + Wo + 1);
    } else {
      d = __ldg(dp + p);
    }
    if (relu && !(g * xh + b > 0.f)) d = 0.f;
    sg += d;
    sgx = fmaf(d, xh, sgx);
  }
  sg = warp_sum(sg);
  sgx = warp_sum(sgx);
  if (lane == 0) {
    if (R == N) {
      dgb[(int64_t)r * 2 * C + c] = sgx;
      dgb[(int64_t)r * 2 * C + C + c] = sg;
    } else {
      atomicAdd(dgb + (int64_t)r * 2 * C + c, sgx);
      atomicAdd(dgb + (int64_t)r * 2 * C + C + c, sg);
    }
  }
}

// per channel: m1 = mean(dxhat) = sum_n gamma*dbeta / cnt, m2 = mean(dxhat*xhat) = sum_n gamma*dgamma / cnt
__global__ void cbn_bwd_chan_kernel(const float* __restrict__ gb, const float* __restrict__ dgb, int N, int C,
                                    float inv_cnt, float* __restrict__ m /* [2C] */) {
  __shared__ float red[32];
  const int c = blockIdx.x;
  float a = 0.f, b = 0.f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const float g = __ldg(gb + (int64_t)n * 2 * C + c);
    a = fmaf(g, __ldg(dgb + (int64_t)n * 2 * C + C + c), a);
    b = fmaf(g, __ldg(dgb + (int64_t)n * 2 * C + c), b);
  }
  a = block_sum(a, red);
  b = block_sum(b, red);
  if (threadIdx.x == 0) {
    m[c] = a * inv_cnt;
    m[C + c] = b * inv_cnt;
  }
}

__global__ void cbn_bwd_dx_kernel(const float* __restrict__ x, const float* __restrict__ gb, int R,
                                  const float* __restrict__ mean, const float* __restrict__ rstd,
                                  const float* __restrict__ dy, const float* __restrict__ m, int N, int C, int H,
                                  int W, int relu, int up, int training, float* __restrict__ dx) {
  const int HW = H * W, Wo = W << up;
  const int64_t total = (int64_t)N * C * HW;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int p = (int)(i % HW);
    const int64_t plane = i / HW;
    const int c = (int)(plane % C);
    const int n = (int)(plane / C);
    const int r = n % R;
    const float g = __ldg(gb + (int64_t)r * 2 * C + c), b = __ldg(gb + (int64_t)r * 2 * C + C + c);
    const float rs = __ldg(rstd + c);
    const float xh = (__ldg(x + i) - __ldg(mean + c)) * rs;
    float d;
    if (up) {
      const int yy = p / W, xx = p - yy * W;
      const float* q = dy + plane * ((int64_t)HW << 2) + (int64_t)(2 * yy) * Wo + 2 * xx;
      d = __ldg(q) + __ldg(q + 1) + __ldg(q + Wo) + __ldg(q + Wo + 1);
    } else {
      d = __ldg(dy + i);
    }
    if (relu && !(g * xh + b > 0.f)) d = 0.f;
    float v = d * g;
    if (training) v = v - __ldg(m + c) - xh * __ldg(m + C + c);
    dx[i] = v * rs;
  }
}


// ------------------------------------------------------------------------------------------------
// vectorised plane walkers (the layouts the model uses: H*W % 4 == 0, W even).  A plane = one (n, c) image of H*W
// floats; `lpp` lanes (a power of two <= 32) share a plane, 32/lpp planes per warp pass, so 4x4 planes still fill a
// warp.  Per-plane parameters are read once per plane; no per-element integer division; 128-bit loads / stores.
// ------------------------------------------------------------------------------------------------
struct PlaneWalk {
  int lane, lpp, ppw, sub, l;
  int64_t first, stride;
  __device__ PlaneWalk(int lanes_per_plane) {
    lane = threadIdx.x & 31;
    lpp = lanes_per_plane;
    ppw = 32 / lpp;
    sub = lane / lpp;
    l = lane - sub * lpp;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    first = warp * ppw + sub;
    stride = (int64_t)gridDim.x * (blockDim.x >> 5) * ppw;
  }
};

__global__ void bn_partial_vec_kernel(const float* __restrict__ x, int N, int C, int HW, int n_per_block,
                                      double* __restrict__ acc /* [2C] */) {
  __shared__ double red[32];
  const int c = blockIdx.x;
  const int nb = blockIdx.y * n_per_block;
  int ne = nb + n_per_block;
  if (ne > N) ne = N;
  double s = 0.0, ss = 0.0;
  const int HW4 = HW >> 2;
  if (HW4 >= (int)blockDim.x) {
    for (int n = nb; n < ne; ++n) {
      const float4* r4 = reinterpret_cast<const float4*>(x + ((int64_t)n * C + c) * HW);
      float a = 0.f, b = 0.f;
      for (int i = threadIdx.x; i < HW4; i += blockDim.x) {
        const float4 v = __ldg(r4 + i);
        a += (v.x + v.y) + (v.z + v.w);
        b += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
      }
      s += (double)a;
      ss += (double)b;
    }
  } else {
    // short planes: threads laid out as (planes per pass) x HW4
    const int rpp = blockDim.x / HW4;
    const int r = threadIdx.x / HW4, q = threadIdx.x - r * HW4;
    if (r < rpp) {
      for (int n = nb + r; n < ne; n += rpp) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(x + ((int64_t)n * C + c) * HW) + q);
        s += (double)((v.x + v.y) + (v.z + v.w));
        ss += (double)((v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w));
      }
    }
  }
  s = block_sum(s, red);
  ss = block_sum(ss, red);
  if (threadIdx.x == 0) {
    atomicAdd(acc + c, s);
    atomicAdd(acc + C + c, ss);
  }
}

template <bool UP>
__global__ void cbn_apply_vec_kernel(const float* __restrict__ x, const float* __restrict__ gb, int R,
                                     const float* __restrict__ mean, const float* __restrict__ rstd, int64_t planes,
                                     int C, int H, int W, int relu, int lpp, float* __restrict__ y) {
  const PlaneWalk pw(lpp);
  const int HW = H * W;
  const int Wh = W >> 1, Wo = W << 1;
  for (int64_t plane = pw.first; plane < planes; plane += pw.stride) {
    const int c = (int)(plane % C);
    const int n = (int)(plane / C);
    const int r = n % R;
    const float g = __ldg(gb + (int64_t)r * 2 * C + c), b = __ldg(gb + (int64_t)r * 2 * C + C + c);
    const float mu = __ldg(mean + c), rs = __ldg(rstd + c);
    auto f = [&](float v) {
      v = g * ((v - mu) * rs) + b;
      return relu ? fmaxf(v, 0.f) : v;
    };
    const float* xp = x + plane * HW;
    if (!UP) {
      const float4* x4 = reinterpret_cast<const float4*>(xp);
      float4* y4 = reinterpret_cast<float4*>(y + plane * HW);
      for (int i = pw.l; i < (HW >> 2); i += pw.lpp) {
        const float4 v = __ldg(x4 + i);
        y4[i] = make_float4(f(v.x), f(v.y), f(v.z), f(v.w));
      }
    } else {
      const float2* x2 = reinterpret_cast<const float2*>(xp);
      float* yp = y + plane * ((int64_t)HW << 2);
      for (int i = pw.l; i < (HW >> 1); i += pw.lpp) {
        const int yy = i / Wh, xq = i - yy * Wh;
        const float2 v = __ldg(x2 + i);
        const float a = f(v.x), bb = f(v.y);
        const float4 o = make_float4(a, a, bb, bb);
        float* row = yp + (int64_t)(2 * yy) * Wo + 4 * xq;
        *reinterpret_cast<float4*>(row) = o;
        *reinterpret_cast<float4*>(row + Wo) = o;
      }
    }
  }
}

// sum over the lanes that share a plane
__device__ __forceinline__ float group_sum(float v, int lpp) {
  for (int o = lpp >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <bool UP>
__global__ void cbn_bwd_plane_vec_kernel(const float* __restrict__ x, const float* __restrict__ gb, int R,
                                         const float* __restrict__ mean, const float* __restrict__ rstd,
                                         const float* __restrict__ dy, int64_t planes, int N, int C, int H, int W,
                                         int relu, int lpp, float* __restrict__ dgb) {
  const PlaneWalk pw(lpp);
  const int HW = H * W;
  const int Wh = W >> 1, Wo = W << 1;
  // every lane runs the same number of passes (the group reduction is warp-wide)
  const int64_t base0 = pw.first - pw.sub;
  for (int64_t base = base0; base < planes; base += pw.stride) {
    const int64_t plane = base + pw.sub;
    const bool live = plane < planes;
    float sg = 0.f, sgx = 0.f;
    int c = 0, r = 0;
    if (live) {
      c = (int)(plane % C);
      const int n = (int)(plane / C);
      r = n % R;
      const float g = __ldg(gb + (int64_t)r * 2 * C + c), b = __ldg(gb + (int64_t)r * 2 * C + C + c);
      const float mu = __ldg(mean + c), rs = __ldg(rstd + c);
      auto acc = [&](float xv, float d) {
        const float xh = (xv - mu) * rs;
        if (relu && !(g * xh + b > 0.f)) d = 0.f;
        sg += d;
        sgx = fmaf(d, xh, sgx);
      };
      const float* xp = x + plane * HW;
      if (!UP) {
        const float4* x4 = reinterpret_cast<const float4*>(xp);
        const float4* d4 = reinterpret_cast<const float4*>(dy + plane * HW);
        for (int i = pw.l; i < (HW >> 2); i += pw.lpp) {
          const float4 v = __ldg(x4 + i), d = __ldg(d4 + i);
          acc(v.x, d.x); acc(v.y, d.y); acc(v.z, d.z); acc(v.w, d.w);
        }
      } else {
        const float2* x2 = reinterpret_cast<const float2*>(xp);
        const float* dp = dy + plane * ((int64_t)HW << 2);
        for (int i = pw.l; i < (HW >> 1); i += pw.lpp) {
          const int yy = i / Wh, xq = i - yy * Wh;
          const float2 v = __ldg(x2 + i);
          const float* row = dp + (int64_t)(2 * yy) * Wo + 4 * xq;
          const float4 q0 = __ldg(reinterpret_cast<const float4*>(row));
          const float4 q1 = __ldg(reinterpret_cast<const float4*>(row + Wo));
          acc(v.x, (q0.x + q0.y) + (q1.x + q1.y));
          acc(v.y, (q0.z + q0.w) + (q1.z + q1.w));
        }
      }
    }
    sg = group_sum(sg, pw.lpp);
    sgx = group_sum(sgx, pw.lpp);
    if (live && pw.l == 0) {
      if (R == N) {
        dgb[(int64_t)r * 2 * C + c] = sgx;
        dgb[(int64_t)r * 2 * C + C + c] = sg;
      } else {
        atomicAdd(dgb + (int64_t)r * 2 * C + c, sgx);
        atomicAdd(dgb + (int64_t)r * 2 * C + C + c, sg);
      }
    }
  }
}

template <bool UP>
__global__ void cbn_bwd_dx_vec_kernel(const float* __restrict__ x, const float* __restrict__ gb, int R,
                                      const float* __restrict__ mean, const float* __restrict__ rstd,
                                      const float* __restrict__ dy, const float* __restrict__ m, int64_t planes, int C,
                                      int H, int W, int relu, int training, int lpp, float* __restrict__ dx) {
  const PlaneWalk pw(lpp);
  const int HW = H * W;
  const int Wh = W >> 1, Wo = W << 1;
  for (int64_t plane = pw.first; plane < planes; plane += pw.stride) {
    const int c = (int)(plane % C);
    const int n = (int)(plane / C);
    const int r = n % R;
    const float g = __ldg(gb + (int64_t)r * 2 * C + c), b = __ldg(gb + (int64_t)r * 2 * C + C + c);
    const float mu = __ldg(mean + c), rs = __ldg(rstd + c);
    const float m1 = training ? __ldg(m + c) : 0.f, m2 = training ? __ldg(m + C + c) : 0.f;
    auto f = [&](float xv, float d) {
      const float xh = (xv - mu) * rs;
      if (relu && !(g * xh + b > 0.f)) d = 0.f;
      float v = d * g;
      if (training) v = v - m1 - xh * m2;
      return v * rs;
    };
    const float* xp = x + plane * HW;
    if (!UP) {
      const float4* x4 = reinterpret_cast<const float4*>(xp);
      const float4* d4 = reinterpret_cast<const float4*>(dy + plane * HW);
      float4* o4 = reinterpret_cast<float4*>(dx + plane * HW);
      for (int i = pw.l; i < (HW >> 2); i += pw.lpp) {
        const float4 v = __ldg(x4 + i), d = __ldg(d4 + i);
        o4[i] = make_float4(f(v.x, d.x), f(v.y, d.y), f(v.z, d.z), f(v.w, d.w));
      }
    } else {
      const float2* x2 = reinterpret_cast<const float2*>(xp);
      const float* dp = dy + plane * ((int64_t)HW << 2);
      float2* o2 = reinterpret_cast<float2*>(dx + plane * HW);
      for (int i = pw.l; i < (HW >> 1); i += pw.lpp) {
        const int yy = i / Wh, xq = i - yy * Wh;
        const float2 v = __ldg(x2 + i);
        const float* row = dp + (int64_t)(2 * yy) * Wo + 4 * xq;
        const float4 q0 = __ldg(reinterpret_cast<const float4*>(row));
        const float4 q1 = __ldg(reinterpret_cast<const float4*>(row + Wo));
        o2[i] = make_float2(f(v.x, (q0.x + q0.y) + (q1.x + q1.y)), f(v.y, (q0.z + q0.w) + (q1.z + q1.w)));
      }
    }
  }
}

// lanes per plane for `vecs` vectors per plane: the smallest power of two >= vecs, capped at 32
static int lanes_per_plane(int vecs) {
  int l = 1;
  while (l < vecs && l < 32) l <<= 1;
  return l;
}
static bool vec_ok(const void* a, const void* b, const void* c, int H, int W, int up) {
  const int HW = H * W;
  const auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!al(a) || !al(b) || (c && !al(c))) return false;
  return up ? (W % 2 == 0 && HW % 2 == 0) : (HW % 4 == 0);
}
static int walker_blocks(int64_t planes, int lpp) {
  const int64_t warps = ceil_div<int64_t>(planes, 32 / lpp);
  const int64_t want = ceil_div<int64_t>(warps, 8);
  const int64_t cap = (int64_t)num_sms() * 16;
  return (int)(want < cap ? (want < 1 ? 1 : want) : cap);
}

}  // namespace dvd

using namespace dvd;

extern "C" int dvd_specnorm_fwd(const float* w_bar, int rows, int cols, float* u, float* v, float* sigma,
                                float* scratch, void* stream) {
  dvd::ProfScope _ps(3, "specnorm_fwd", dvd::as_stream(stream));
  DVD_CHECK_ARG(w_bar && u && v && sigma && scratch && rows > 0 && cols > 0);
  cudaStream_t st = as_stream(stream);
  float* vraw = scratch;
  float* s = scratch + cols;
  sn_colsum_kernel<<<ceil_div(cols, 32), 256, 0, st>>>(w_bar, u, rows, cols, vraw);
  DVD_LAUNCH_CHECK();
  sn_rowdot_kernel<<<ceil_div(rows, 8), 256, 0, st>>>(w_bar, vraw, rows, cols, s);
  DVD_LAUNCH_CHECK();
  sn_finalize_kernel<<<1, 256, 0, st>>>(vraw, s, rows, cols, u, v, sigma);
  DVD_LAUNCH_CHECK();
  return 0;
}

extern "C" int dvd_specnorm_bwd(const float* g, const float* w_bar, const float* u, const float* v,
                                const float* sigma, int rows, int cols, float* dw_bar, int accumulate, void* scratch,
                                void* stream) {
  dvd::ProfScope _ps(3, "specnorm_bwd", dvd::as_stream(stream));
  DVD_CHECK_ARG(g && w_bar && u && v && sigma && dw_bar && scratch && rows > 0 && cols > 0);
  cudaStream_t st = as_stream(stream);
  double* dot = reinterpret_cast<double*>(scratch);
  DVD_CUDA(cudaMemsetAsync(dot, 0, sizeof(double), st));
  const int64_t n = (int64_t)rows * cols;
  sn_dot_kernel<<<ew_blocks(n, 8), 256, 0, st>>>(g, w_bar, n, dot);
  DVD_LAUNCH_CHECK();
  sn_bwd_kernel<<<ew_blocks(n, 4), 256, 0, st>>>(g, u, v, sigma, dot, rows, cols, dw_bar, accumulate);
  DVD_LAUNCH_CHECK();
  return 0;
}

// phase 0: everything.  Cross-replica statistics (Generator.py:57-58's TODO) split the call around a collective:
// phase 1 leaves the per-channel (sum, sum of squares) of this rank's N*HW elements in `scratch` (2C doubles), the caller
// sums that over the ranks, phase 2 finalises with `count_scale` x (N*HW) elements.
extern "C" int dvd_bn_stats_ex(const float* x, int N, int C, int HW, int training, float momentum, float eps,
                               float* running_mean, float* running_var, int64_t* num_batches_tracked, float* mean,
                               float* rstd, void* scratch, int phase, int count_scale, void* stream) {
  dvd::ProfScope _ps(3, "bn_stats", dvd::as_stream(stream));
  DVD_CHECK_ARG(x && mean && rstd && scratch && N > 0 && C > 0 && HW > 0 && phase >= 0 && phase <= 2 && count_scale >= 1);
  DVD_CHECK_ARG(training || (running_mean && running_var));
  DVD_CHECK_ARG(phase == 0 || training);
  cudaStream_t st = as_stream(stream);
  double* acc = reinterpret_cast<double*>(scratch);
  if (training && phase != 2) {
    DVD_CUDA(cudaMemsetAsync(acc, 0, sizeof(double) * 2 * C, st));
    // enough blocks to cover the machine: C x splits over the batch
    int splits = ceil_div(4 * num_sms(), C);
    if (splits > N) splits = N;
    if (splits < 1) splits = 1;
    const int npb = ceil_div(N, splits);
    splits = ceil_div(N, npb);
    if (HW % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (HW >> 2) <= 256 * 1024)
      bn_partial_vec_kernel<<<dim3(C, splits), 256, 0, st>>>(x, N, C, HW, npb, acc);
    else
      bn_partial_kernel<<<dim3(C, splits), 256, 0, st>>>(x, N, C, HW, npb, acc);
    DVD_LAUNCH_CHECK();
  }
  if (phase == 1) return 0;
  bn_finalize_kernel<<<ceil_div(C, 128), 128, 0, st>>>(acc, C, (double)N * HW * count_scale, training, momentum, eps,
                                                       running_mean, running_var,
                                                       reinterpret_cast<long long*>(num_batches_tracked), mean, rstd);
  DVD_LAUNCH_CHECK();
  return 0;
}
extern "C" int dvd_bn_stats(const float* x, int N, int C, int HW, int training, float momentum, float eps,
                            float* running_mean, float* running_var, int64_t* num_batches_tracked, float* mean,
                            float* rstd, void* scratch, void* stream) {
  return dvd_bn_stats_ex(x, N, C, HW, training, momentum, eps, running_mean, running_var, num_batches_tracked, mean, rstd,
                         scratch, 0, 1, stream);
}

extern "C" int dvd_cbn_apply(const float* x, const float* gb, int gb_rows, const float* mean, const float* rstd, int N,
                             int C, int H, int W, int relu, int up, float* y, void* stream) {
  dvd::ProfScope _ps(3, "cbn_apply", dvd::as_stream(stream));
  DVD_CHECK_ARG(x && gb && mean && rstd && y && N > 0 && C > 0 && H > 0 && W > 0 && (up == 0 || up == 1));
  DVD_CHECK_ARG(gb_rows > 0 && gb_rows <= N);
  const int64_t total = ((int64_t)N * C * H * W) << (2 * up);
  if (vec_ok(x, y, nullptr, H, W, up)) {
    const int64_t planes = (int64_t)N * C;
    const int lpp = lanes_per_plane(up ? (H * W) >> 1 : (H * W) >> 2);
    const int nb = walker_blocks(planes, lpp);
    if (up) cbn_apply_vec_kernel<true><<<nb, 256, 0, as_stream(stream)>>>(x, gb, gb_rows, mean, rstd, planes, C, H, W, relu, lpp, y);
    else cbn_apply_vec_kernel<false><<<nb, 256, 0, as_stream(stream)>>>(x, gb, gb_rows, mean, rstd, planes, C, H, W, relu, lpp, y);
  } else {
    cbn_apply_kernel<<<ew_blocks(total, 4), 256, 0, as_stream(stream)>>>(x, gb, gb_rows, mean, rstd, N, C, H, W, relu, up, y);
  }
  DVD_LAUNCH_CHECK();
  return 0;
}

// phase 0: everything.  Cross-replica statistics: phase 1 computes dgb and leaves this rank's per-channel
// (mean(dxhat), mean(dxhat * xhat)) in `scratch` (2C floats), the caller AVERAGES that over the (equal-sized) ranks,
// phase 2 computes dx from the averaged means.
extern "C" int dvd_cbn_bwd_ex(const float* x, const float* gb, int gb_rows, const float* mean, const float* rstd,
                              const float* dy, int N, int C, int H, int W, int relu, int up, int training, float* dx,
                              float* dgb, float* scratch, int phase, void* stream) {
  dvd::ProfScope _ps(3, "cbn_bwd", dvd::as_stream(stream));
  DVD_CHECK_ARG(x && gb && mean && rstd && dy && dx && dgb && scratch && N > 0 && C > 0 && (up == 0 || up == 1));
  DVD_CHECK_ARG(gb_rows > 0 && gb_rows <= N && phase >= 0 && phase <= 2 && (phase == 0 || training));
  cudaStream_t st = as_stream(stream);
  const int64_t planes = (int64_t)N * C;
  const bool vec = vec_ok(x, dy, dx, H, W, up);
  const int lpp = lanes_per_plane(up ? (H * W) >> 1 : (H * W) >> 2);
  const int nb = walker_blocks(planes, lpp);
  if (phase != 2) {
  if (gb_rows != N) DVD_CUDA(cudaMemsetAsync(dgb, 0, sizeof(float) * (size_t)gb_rows * 2 * C, st));
  if (vec) {
    if (up) cbn_bwd_plane_vec_kernel<true><<<nb, 256, 0, st>>>(x, gb, gb_rows, mean, rstd, dy, planes, N, C, H, W, relu, lpp, dgb);
    else cbn_bwd_plane_vec_kernel<false><<<nb, 256, 0, st>>>(x, gb, gb_rows, mean, rstd, dy, planes, N, C, H, W, relu, lpp, dgb);
  } else {
    cbn_bwd_plane_kernel<<<(unsigned)ceil_div<int64_t>(planes, 8), 256, 0, st>>>(x, gb, gb_rows, mean, rstd, dy, N, C, H,
                                                                               W, relu, up, dgb);
  }
  DVD_LAUNCH_CHECK();
  if (training) {
    cbn_bwd_chan_kernel<<<C, 256, 0, st>>>(gb, dgb, gb_rows, C, 1.f / ((float)N * H * W), scratch);
    DVD_LAUNCH_CHECK();
  }
  }
  if (phase == 1) return 0;
  if (vec) {
    if (up) cbn_bwd_dx_vec_kernel<true><<<nb, 256, 0, st>>>(x, gb, gb_rows, mean, rstd, dy, scratch, planes, C, H, W, relu, training, lpp, dx);
    else cbn_bwd_dx_vec_kernel<false><<<nb, 256, 0, st>>>(x, gb, gb_rows, mean, rstd, dy, scratch, planes, C, H, W, relu, training, lpp, dx);
  } else {
    cbn_bwd_dx_kernel<<<ew_blocks(planes * H * W, 4), 256, 0, st>>>(x, gb, gb_rows, mean, rstd, dy, scratch, N, C, H, W,
                                                                    relu, up, training, dx);
  }
  DVD_LAUNCH_CHECK();
  return 0;
}
extern "C" int dvd_cbn_bwd(const float* x, const float* gb, int gb_rows, const float* mean, const float* rstd,
                           const float* dy, int N, int C, int H, int W, int relu, int up, int training, float* dx,
                           float* dgb, float* scratch, void* stream) {
  return dvd_cbn_bwd_ex(x, gb, gb_rows, mean, rstd, dy, N, C, H, W, relu, up, training, dx, dgb, scratch, 0, stream);
}
