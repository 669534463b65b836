// ConvGRU layer over a whole clip (ConvGRU.py:29-54 driven by Generator.py:87-97), forward and BPTT.
//
// conv(cat[x,h]) = conv_x(x) + conv_h(h): the x-halves of all three gates are one batched implicit GEMM over
// all T frames (Cout = 3*Ch, bias folded in); only the h-halves are sequential in time.  Gates live in one
// (B,T,3Ch,H,W) buffer = (update, reset, out) that is activated in place and, in the backward sweep,
// overwritten in place with the pre-activation gradients, which then feed ONE batched x-dgrad and the batched
// weight gradients.
//
// Two ways of overlapping the sequential part: the whole-clip entry points split the batch into
// independent chains on helper streams (SliceFork below); the frame-range entry points run frames [t0, t1) only, with
// the sweep's state kept in the caller's workspace between calls, so that a caller can run the layers of a stack as a
// wavefront, one stream per layer (ops.GRUStackFn).
#include <cuda_bf16.h>

#include "common.cuh"
#include "conv_params.cuh"

namespace dvd {

// u = sigmoid(gu), r = sigmoid(gr) in place; rh = r * h_prev
__global__ void gru_gate_ur_kernel(float* __restrict__ g_t, int64_t g_bs, const float* __restrict__ hp, int64_t hp_bs,
                                   float* __restrict__ rh_t, int64_t rh_bs, int B, int64_t chw) {
  const int64_t total = (int64_t)B * chw;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / chw);
    const int64_t r = i - (int64_t)b * chw;
    float* g = g_t + b * g_bs;
    const float u = sigmoidf_(g[r]);
    const float rr = sigmoidf_(g[chw + r]);
    g[r] = u;
    g[chw + r] = rr;
    const float h = hp ? hp[b * hp_bs + r] : 0.f;
    rh_t[b * rh_bs + r] = rr * h;
  }
}

// o = tanh(go) in place; h = h_prev * (1 - u) + o * u
__global__ void gru_out_kernel(float* __restrict__ g_t, int64_t g_bs, const float* __restrict__ hp, int64_t hp_bs,
                               float* __restrict__ h_t, int64_t h_bs, int B, int64_t chw) {
  const int64_t total = (int64_t)B * chw;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / chw);
    const int64_t r = i - (int64_t)b * chw;
    float* g = g_t + b * g_bs;
    const float u = g[r];
    const float o = tanhf(g[2 * chw + r]);
    g[2 * chw + r] = o;
    const float h = hp ? hp[b * hp_bs + r] : 0.f;
    h_t[b * h_bs + r] = h * (1.f - u) + o * u;
  }
}

// backward, part 1: dhn = dh_ext + carry_in;  da_u -> u slot, da_o -> o slot, carry_out = dhn * (1 - u)
__global__ void gru_bwd1_kernel(float* __restrict__ g_t, int64_t g_bs, const float* __restrict__ hp, int64_t hp_bs,
                                const float* __restrict__ dh_t, int64_t dh_bs, const float* __restrict__ carry_in,
                                float* __restrict__ carry_out, int B, int64_t chw) {
  const int64_t total = (int64_t)B * chw;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / chw);
    const int64_t r = i - (int64_t)b * chw;
    float* g = g_t + b * g_bs;
    const float u = g[r], o = g[2 * chw + r];
    const float h = hp ? hp[b * hp_bs + r] : 0.f;
    float dhn = dh_t[b * dh_bs + r];
    if (carry_in) dhn += carry_in[i];
    const float du = dhn * (o - h);
    const float d_o = dhn * u;
    g[2 * chw + r] = d_o * (1.f - o * o);
    g[r] = du * u * (1.f - u);
    carry_out[i] = dhn * (1.f - u);
  }
}

// backward, part 2: dr = d_rh * h_prev; carry_out += d_rh * r; da_r -> r slot
__global__ void gru_bwd2_kernel(float* __restrict__ g_t, int64_t g_bs, const float* __restrict__ hp, int64_t hp_bs,
                                const float* __restrict__ d_rh, float* __restrict__ carry_out, int B, int64_t chw) {
  const int64_t total = (int64_t)B * chw;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / chw);
    const int64_t r = i - (int64_t)b * chw;
    float* g = g_t + b * g_bs;
    if (!hp) {          // zero previous state: no gradient reaches the reset gate
      g[chw + r] = 0.f;
      continue;
    }
    const float rr = g[chw + r];
    const float h = hp[b * hp_bs + r];
    const float drh = d_rh[i];
    carry_out[i] += drh * rr;
    g[chw + r] = drh * h * rr * (1.f - rr);
  }
}

// ---- backward kernels that ALSO write the bf16 (hi, lo) operand planes of the pre-activation gradients.
// Planes: [B*T][HW][G3P] channels-last, channel order (da_u | da_r | da_o) = the layer's gate order; they are read
// by the two per-step dgrad GEMMs, the batched x-dgrad and the three weight-gradient GEMMs, so the gradients are
// split exactly once, by the kernel that produces them.  Block = 32 pixels x 64 channels of one image, transposed
// through shared memory (NCHW-coalesced loads and fp32 stores, channels-last 128-bit plane stores).
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

// part 1 (see gru_bwd1_kernel) for image b = blockIdx.z, channels [64*blockIdx.y, +64), pixels [32*blockIdx.x, +32)
__global__ void __launch_bounds__(256) gru_bwd1_planes_kernel(
    float* __restrict__ g_t, int64_t g_bs, const float* __restrict__ hp, int64_t hp_bs, const float* __restrict__ dh_t,
    int64_t dh_bs, const float* __restrict__ carry_in, float* __restrict__ carry_out, int Ch, int HW,
    __nv_bfloat16* __restrict__ pl_hi, __nv_bfloat16* __restrict__ pl_lo, int64_t pl_img_stride, int G3P) {
  __shared__ float t_o[64][33], t_u[64][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 64, pix0 = blockIdx.x * 32, tid = threadIdx.x;
  const int64_t chw = (int64_t)Ch * HW;
  {
    const int px = tid & 31, pix = pix0 + px;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int cl = (tid >> 5) + 8 * j, c = c0 + cl;
      float da_o = 0.f, da_u = 0.f;
      if (pix < HW) {
        const int64_t r = (int64_t)c * HW + pix;
        float* g = g_t + b * g_bs;
        const float u = g[r], o = g[2 * chw + r];
        const float h = hp ? hp[b * hp_bs + r] : 0.f;
        float dhn = dh_t[b * dh_bs + r];
        if (carry_in) dhn += carry_in[b * chw + r];
        da_o = dhn * u * (1.f - o * o);
        da_u = dhn * (o - h) * u * (1.f - u);
        g[2 * chw + r] = da_o;
        g[r] = da_u;
        carry_out[b * chw + r] = dhn * (1.f - u);
      }
      t_o[cl][px] = da_o;
      t_u[cl][px] = da_u;
    }
  }
  __syncthreads();
  {
    const int px = tid >> 3, q = tid & 7, pix = pix0 + px;
    if (pix < HW) {
      float vo[8], vu[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) { vo[i] = t_o[q * 8 + i][px]; vu[i] = t_u[q * 8 + i][px]; }
      const int64_t row = (int64_t)b * pl_img_stride + (int64_t)pix * G3P + c0 + q * 8;
      uint4 hi, lo;
      bf16_split8(vu, &hi, &lo);
      *reinterpret_cast<uint4*>(pl_hi + row) = hi;
      *reinterpret_cast<uint4*>(pl_lo + row) = lo;
      bf16_split8(vo, &hi, &lo);
      *reinterpret_cast<uint4*>(pl_hi + row + 2 * Ch) = hi;
      *reinterpret_cast<uint4*>(pl_lo + row + 2 * Ch) = lo;
    }
  }
}

// part 2 (see gru_bwd2_kernel): da_r -> reset slot and the planes at channel offset Ch
__global__ void __launch_bounds__(256) gru_bwd2_planes_kernel(
    float* __restrict__ g_t, int64_t g_bs, const float* __restrict__ hp, int64_t hp_bs, const float* __restrict__ d_rh,
    float* __restrict__ carry_out, int Ch, int HW, __nv_bfloat16* __restrict__ pl_hi, __nv_bfloat16* __restrict__ pl_lo,
    int64_t pl_img_stride, int G3P) {
  __shared__ float t_r[64][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 64, pix0 = blockIdx.x * 32, tid = threadIdx.x;
  const int64_t chw = (int64_t)Ch * HW;
  {
    const int px = tid & 31, pix = pix0 + px;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int cl = (tid >> 5) + 8 * j, c = c0 + cl;
      float da_r = 0.f;
      if (pix < HW) {
        const int64_t r = (int64_t)c * HW + pix;
        float* g = g_t + b * g_bs;
        if (hp) {
          const float rr = g[chw + r];
          const float h = hp[b * hp_bs + r];
          const float drh = d_rh[b * chw + r];
          carry_out[b * chw + r] += drh * rr;
          da_r = drh * h * rr * (1.f - rr);
        }
        g[chw + r] = da_r;
      }
      t_r[cl][px] = da_r;
    }
  }
  __syncthreads();
  {
    const int px = tid >> 3, q = tid & 7, pix = pix0 + px;
    if (pix < HW) {
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = t_r[q * 8 + i][px];
      const int64_t row = (int64_t)b * pl_img_stride + (int64_t)pix * G3P + Ch + c0 + q * 8;
      uint4 hi, lo;
      bf16_split8(v, &hi, &lo);
      *reinterpret_cast<uint4*>(pl_hi + row) = hi;
      *reinterpret_cast<uint4*>(pl_lo + row) = lo;
    }
  }
}

struct GruWs {
  float *wx, *whur, *who, *bias;        // forward operands
  float *wxT, *whurT, *whoT;            // dgrad operands
  float *dwx, *dwhur, *dwho;            // packed weight grads
  float *d_rh, *carry0, *carry1, *dbias;
  double* dscratch;
  // fused forward path: hi/lo planes (forward format) of the h-half weights and of the per-step GEMM inputs
  uint16_t *whur_hi, *whur_lo, *who_hi, *who_lo;      // [taps][CoutP][ChP]
  uint16_t *hpl_hi[2], *hpl_lo[2];                    // planes of h_{t-1} / h_t (ping-pong), [B*HW][ChP]
  uint16_t *rhpl_hi, *rhpl_lo;                        // planes of r * h_{t-1}
  uint16_t *whoT_hi, *whoT_lo, *whurT_hi, *whurT_lo;  // backward: [taps][ChP][ChP] / [taps][ChP][round64(2Ch)]
  uint16_t* gplanes;                                  // frame-range API only: (hi | lo) planes of the gate gradients
  size_t bytes;
};

// T > 0: also room for the bf16 planes [B*T][HW][round64(3Ch)] x (hi, lo) of the pre-activation gradients, which the
// whole-clip entry point takes from the stream-ordered scratch pool but a sweep spread over several calls keeps here
static GruWs carve(void* base, int B, int Cx, int Ch, int HW, int taps, bool bwd, int T = 0) {
  GruWs w;
  char* p = reinterpret_cast<char*>(base);
  size_t off = 0;
  auto take = [&](size_t n_floats) {
    float* r = reinterpret_cast<float*>(p + off);
    off += ((n_floats * sizeof(float) + 255) / 256) * 256;
    return r;
  };
  const size_t nx = (size_t)taps * Cx * 3 * Ch, nur = (size_t)taps * Ch * 2 * Ch, no = (size_t)taps * Ch * Ch;
  w.wx = take(nx); w.whur = take(nur); w.who = take(no); w.bias = take(3 * Ch);
  if (bwd) {
    w.wxT = take(nx); w.whurT = take(nur); w.whoT = take(no);
    w.dwx = take(nx); w.dwhur = take(nur); w.dwho = take(no);
    const size_t st = (size_t)B * Ch * HW;
    w.d_rh = take(st); w.carry0 = take(st); w.carry1 = take(st); w.dbias = take(3 * Ch);
    w.dscratch = reinterpret_cast<double*>(take(2 * 3 * Ch));
  } else {
    w.wxT = w.whurT = w.whoT = w.dwx = w.dwhur = w.dwho = w.d_rh = w.carry0 = w.carry1 = w.dbias = nullptr;
    w.dscratch = nullptr;
  }
  {
    const size_t ChP = (size_t)tma_round64(Ch), Co2P = (size_t)tma_round64(2 * Ch);
    auto take16 = [&](size_t n) { return reinterpret_cast<uint16_t*>(take((n + 1) / 2)); };
    w.whur_hi = take16((size_t)taps * Co2P * ChP); w.whur_lo = take16((size_t)taps * Co2P * ChP);
    w.who_hi = take16((size_t)taps * ChP * ChP); w.who_lo = take16((size_t)taps * ChP * ChP);
    const size_t pl = (size_t)B * HW * ChP;
    for (int i = 0; i < 2; ++i) { w.hpl_hi[i] = take16(pl); w.hpl_lo[i] = take16(pl); }
    w.rhpl_hi = take16(pl); w.rhpl_lo = take16(pl);
    if (bwd) {     // bf16 planes of the transposed h-half weights (dgrad operands), split once per layer
      w.whoT_hi = take16((size_t)taps * ChP * ChP); w.whoT_lo = take16((size_t)taps * ChP * ChP);
      w.whurT_hi = take16((size_t)taps * ChP * Co2P); w.whurT_lo = take16((size_t)taps * ChP * Co2P);
    } else {
      w.whoT_hi = w.whoT_lo = w.whurT_hi = w.whurT_lo = nullptr;
    }
    w.gplanes = T > 0 ? take16(2 * (size_t)B * T * HW * tma_round64(3 * Ch) + 128) : nullptr;
  }
  w.bytes = off;
  return w;
}

// option "gru_fused" = 0: keep the gate math in separate elementwise kernels
static bool gru_fused_enabled() { return get_option(OPT_GRU_FUSED) != 0; }

static dvd_conv_desc base_desc(int B, int T, int Cin, int Cout, int H, int W, int k) {
  dvd_conv_desc d;
  memset(&d, 0, sizeof(d));
  d.N1 = B; d.N2 = T; d.Cin = Cin; d.Cout = Cout; d.D = 1; d.H = H; d.W = W; d.kD = 1; d.kH = k; d.kW = k;
  d.x_cs = d.y_cs = d.r_cs = (int64_t)H * W;
  return d;
}

// ---- batch slices of the time loop as independent chains on helper streams (option "gru_streams").
// The two GEMMs of a time step depend on each other, so one layer's loop is a chain of launches that each leave the
// tail of their last wave of tiles empty (256 pair-tiles on 74 SM pairs = 3.46 waves) and expose their epilogues.  The
// clips of a batch do not interact inside a ConvGRU, so the loop runs as `ns` chains over B / ns clips each: while one
// chain's kernel drains, the CTAs of the other chain's kernel take the free SMs, and the elementwise kernels of the BPTT
// sweep overlap the other chain's GEMMs.  Chain 0 stays on the caller's stream; chains 1.. run on library-owned
// non-blocking streams (per host thread and device) that are forked from / joined to the caller's stream with events,
// so the call keeps its stream-ordered semantics.
constexpr int kMaxSlices = 4;
struct HelperStreams {
  cudaStream_t s[kMaxSlices - 1] = {nullptr, nullptr, nullptr};
  cudaEvent_t fork = nullptr;
  cudaEvent_t join[kMaxSlices - 1] = {nullptr, nullptr, nullptr};
  bool ready = false;
};
static HelperStreams* helper_streams() {
  static thread_local HelperStreams per_dev[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  HelperStreams& h = per_dev[dev];
  if (!h.ready) {
    if (cudaEventCreateWithFlags(&h.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    for (int i = 0; i < kMaxSlices - 1; ++i) {
      if (cudaStreamCreateWithFlags(&h.s[i], cudaStreamNonBlocking) != cudaSuccess) return nullptr;
      if (cudaEventCreateWithFlags(&h.join[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
    }
    h.ready = true;
  }
  return &h;
}

class SliceFork {
 public:
  SliceFork(cudaStream_t main_stream, int n) : main_(main_stream), n_(n) {}
  ~SliceFork() { if (open_) end(); }
  int begin() {           // everything queued on the caller's stream so far happens before the helper chains
    if (n_ <= 1) return 0;
    hs_ = helper_streams();
    if (!hs_) return fail("cannot create the helper streams%s (%s:%d)", "", __FILE__, __LINE__);
    DVD_CUDA(cudaEventRecord(hs_->fork, main_));
    for (int i = 0; i < n_ - 1; ++i) DVD_CUDA(cudaStreamWaitEvent(hs_->s[i], hs_->fork, 0));
    open_ = true;
    return 0;
  }
  cudaStream_t stream(int i) const { return i == 0 ? main_ : hs_->s[i - 1]; }
  int end() {             // the caller's stream continues after every chain
    if (!open_) return 0;
    open_ = false;
    for (int i = 0; i < n_ - 1; ++i) {
      DVD_CUDA(cudaEventRecord(hs_->join[i], hs_->s[i]));
      DVD_CUDA(cudaStreamWaitEvent(main_, hs_->join[i], 0));
    }
    return 0;
  }

 private:
  cudaStream_t main_;
  int n_;
  HelperStreams* hs_ = nullptr;
  bool open_ = false;
};

// number of chains: the option, lowered until it divides the batch and the sliced GEMMs stay on the tensor path
template <typename Ok>
static int pick_slices(int B, int T, Ok ok) {
  int ns = get_option(OPT_GRU_STREAMS);
  if (ns > kMaxSlices) ns = kMaxSlices;
  if (T <= 1 || ns < 1) ns = 1;
  while (ns > 1 && (B % ns != 0 || !ok(B / ns))) --ns;
  return ns;
}

}  // namespace dvd

using namespace dvd;

extern "C" size_t dvd_convgru_layer_workspace_bytes(int B, int T, int Cx, int Ch, int H, int W, int k) {
  (void)T;
  return carve(nullptr, B, Cx, Ch, H * W, k * k, true).bytes + 256;
}
extern "C" size_t dvd_convgru_layer_range_workspace_bytes(int B, int T, int Cx, int Ch, int H, int W, int k) {
  return carve(nullptr, B, Cx, Ch, H * W, k * k, true, T).bytes + 256;
}

// Frames [t0, t1) of the forward sweep.  range = false: the whole clip in one call (t0 = 0, t1 = T).  range = true: one
// of several calls that share `workspace` and visit the frames in order -- the call with t0 = 0 packs / splits the
// weights into the workspace, later calls find them (and the operand planes of h_{t0-1}) there.
static int gru_fwd_impl(const float* x, int64_t x_bs, int64_t x_ts, const float* h0, const float* wu, const float* wr,
                        const float* wo, const float* bu, const float* br, const float* bo, float* gates, float* h,
                        float* rh, int B, int T, int Cx, int Ch, int H, int W, int k, int t0, int t1, bool range,
                        void* workspace, size_t ws_bytes, void* stream) {
  DVD_CHECK_ARG(x && wu && wr && wo && bu && br && bo && gates && h && rh && workspace);
  DVD_CHECK_ARG(B > 0 && T > 0 && Cx > 0 && Ch > 0 && H > 0 && W > 0 && (k & 1));
  DVD_CHECK_ARG(0 <= t0 && t0 < t1 && t1 <= T);
  const int HW = H * W, taps = k * k, Ct = Cx + Ch;
  GruWs ws = carve(workspace, B, Cx, Ch, HW, taps, false);
  DVD_CHECK_ARG(ws.bytes <= ws_bytes);
  cudaStream_t st = as_stream(stream);
  const float* wsrc[3] = {wu, wr, wo};
  const float* bsrc[3] = {bu, br, bo};
  for (int g = 0; g < 3 && t0 == 0; ++g) {
    DVD_TRY(dvd_weight_pack(wsrc[g], Ct, taps, 0, Ch, 0, Cx, nullptr, 0, ws.wx, Cx, 0, 3 * Ch, g * Ch, stream));
    if (g < 2) DVD_TRY(dvd_weight_pack(wsrc[g], Ct, taps, 0, Ch, Cx, Ch, nullptr, 0, ws.whur, Ch, 0, 2 * Ch, g * Ch, stream));
    else DVD_TRY(dvd_weight_pack(wsrc[g], Ct, taps, 0, Ch, Cx, Ch, nullptr, 0, ws.who, Ch, 0, Ch, 0, stream));
    DVD_CUDA(cudaMemcpyAsync(ws.bias + g * Ch, bsrc[g], sizeof(float) * Ch, cudaMemcpyDeviceToDevice, st));
  }
  const int64_t chw = (int64_t)Ch * HW;
  const int64_t g_ts = 3 * chw, g_bs = (int64_t)T * g_ts, h_ts = chw, h_bs = (int64_t)T * chw;
  // x-halves of all gates, all frames of the range: one implicit GEMM
  {
    dvd_conv_desc d = base_desc(B, t1 - t0, Cx, 3 * Ch, H, W, k); d.x_kind = 1;
    d.x_s1 = x_bs; d.x_s2 = x_ts; d.y_s1 = g_bs; d.y_s2 = g_ts;
    DVD_TRY(dvd_conv_fwd(&d, x + (int64_t)t0 * x_ts, ws.wx, ws.bias, nullptr, gates + (int64_t)t0 * g_ts, stream));
  }
  // Fused path: the two h-half GEMMs of a step run on the TMA/tcgen05 engine with the gate math in their epilogues
  // (update|reset: sigmoid, r*h;  out: tanh, state update) which also emit the operand planes the next GEMM reads, and the
  // h-half weight planes are split once per layer instead of once per step.
  auto step_descs = [&](int Bn, dvd_conv_desc* ur, dvd_conv_desc* o) {
    *ur = base_desc(Bn, 1, Ch, 2 * Ch, H, W, k); *o = base_desc(Bn, 1, Ch, Ch, H, W, k);
    ur->x_kind = o->x_kind = 1; ur->accumulate = o->accumulate = 1;
    ur->y_s1 = o->y_s1 = g_bs; ur->x_s1 = o->x_s1 = chw;
  };
  dvd_conv_desc d_ur, d_o;
  const bool can_fuse = gru_fused_enabled() && T > 1 && Ch % 32 == 0;
  // chains over batch slices (see SliceFork): only where every slice stays on the fused tensor path; a caller of the
  // frame-range API overlaps whole layers instead
  const int ns = range ? 1 : pick_slices(B, T, [&](int Bn) {
    step_descs(Bn, &d_ur, &d_o);
    return can_fuse && conv_fwd_ex_eligible(&d_ur) && conv_fwd_ex_eligible(&d_o);
  });
  const int Bn = B / ns;
  step_descs(Bn, &d_ur, &d_o);
  const int eb = ew_blocks((int64_t)Bn * chw);
  const int ChP = tma_round64(Ch), Co2P = tma_round64(2 * Ch);
  const bool fused = can_fuse && conv_fwd_ex_eligible(&d_ur) && conv_fwd_ex_eligible(&d_o);
  const int f16 = tma_forward_planes_fp16() ? 1 : 0;       // the epilogues write the step-to-step planes in this format
  if (fused && t0 == 0) {
    DVD_TRY(tma_split_weights(ws.whur, taps, Ch, 2 * Ch, Co2P, f16, ws.whur_hi, ws.whur_lo, st));
    DVD_TRY(tma_split_weights(ws.who, taps, Ch, Ch, ChP, f16, ws.who_hi, ws.who_lo, st));
    if (ChP != Ch) {      // padded channels of the planes the epilogues write stay zero
      const size_t pl = (size_t)B * HW * ChP * sizeof(uint16_t);
      for (int i = 0; i < 2; ++i) {
        DVD_CUDA(cudaMemsetAsync(ws.hpl_hi[i], 0, pl, st));
        DVD_CUDA(cudaMemsetAsync(ws.hpl_lo[i], 0, pl, st));
      }
      DVD_CUDA(cudaMemsetAsync(ws.rhpl_hi, 0, pl, st));
      DVD_CUDA(cudaMemsetAsync(ws.rhpl_lo, 0, pl, st));
    }
  }
  SliceFork fk(st, ns);
  DVD_TRY(fk.begin());
  // planes of h_{t-1} already sit in slot (t-1) & 1: written by the epilogue of step t-1 (a fused step, i.e. one with a
  // previous state), possibly in the previous call
  bool have_planes = fused && t0 > 0 && (t0 > 1 || h0);
  for (int t = t0; t < t1; ++t) {
    const bool first = t == 0 && !h0;      // zero previous state
    const int64_t hp_bs = t > 0 ? h_bs : chw;
    for (int sl = 0; sl < ns; ++sl) {      // step t of every chain, interleaved so that all chains have work queued
      cudaStream_t ss = fk.stream(sl);
      const int b0 = sl * Bn;
      const float* hp = first ? nullptr : (t > 0 ? h + (int64_t)(t - 1) * h_ts : h0) + (int64_t)b0 * hp_bs;
      float* g_t = gates + (int64_t)t * g_ts + (int64_t)b0 * g_bs;
      float* rh_t = rh + (int64_t)t * h_ts + (int64_t)b0 * h_bs;
      float* h_t = h + (int64_t)t * h_ts + (int64_t)b0 * h_bs;
      if (fused && hp) {
        const int sp = (t + 1) & 1, sn = t & 1;          // slots of h_{t-1} and h_t
        const size_t po = (size_t)b0 * HW * ChP;         // this slice's rows of the [B*HW][ChP] planes
        if (!have_planes)
          DVD_TRY(tma_split_activations(hp, Bn, Ch, hp_bs, HW, HW, f16, ws.hpl_hi[sp] + po, ws.hpl_lo[sp] + po, ss));
        TmaOperands op;
        GruEpi ge;
        op.a_hi = ws.hpl_hi[sp] + po; op.a_lo = ws.hpl_lo[sp] + po; op.w_hi = ws.whur_hi; op.w_lo = ws.whur_lo;
        op.CoutP = Co2P;
        ge.mode = 1; ge.Ch = Ch; ge.hprev = hp; ge.hp_s1 = hp_bs; ge.out2 = rh_t; ge.o2_s1 = h_bs;
        ge.pl_hi = ws.rhpl_hi + po; ge.pl_lo = ws.rhpl_lo + po; ge.pl_Cp = ChP;
        DVD_TRY(conv_fwd_ex(&d_ur, nullptr, nullptr, g_t, &op, &ge, ss));
        op.a_hi = ws.rhpl_hi + po; op.a_lo = ws.rhpl_lo + po; op.w_hi = ws.who_hi; op.w_lo = ws.who_lo; op.CoutP = ChP;
        ge.mode = 2; ge.ugate = g_t; ge.u_s1 = g_bs; ge.out2 = h_t; ge.o2_s1 = h_bs;
        ge.pl_hi = ws.hpl_hi[sn] + po; ge.pl_lo = ws.hpl_lo[sn] + po;
        DVD_TRY(conv_fwd_ex(&d_o, nullptr, nullptr, g_t + 2 * chw, &op, &ge, ss));
        continue;
      }
      if (hp) {
        dvd_conv_desc d = base_desc(Bn, 1, Ch, 2 * Ch, H, W, k); d.x_kind = 1;
        d.x_s1 = hp_bs; d.y_s1 = g_bs; d.accumulate = 1;
        DVD_TRY(dvd_conv_fwd(&d, hp, ws.whur, nullptr, nullptr, g_t, ss));
      }
      { ProfScope ps(3, "gru_gate_ur", ss); gru_gate_ur_kernel<<<eb, 256, 0, ss>>>(g_t, g_bs, hp, hp_bs, rh_t, h_bs, Bn, chw); }
      DVD_LAUNCH_CHECK();
      if (hp) {
        dvd_conv_desc d = base_desc(Bn, 1, Ch, Ch, H, W, k); d.x_kind = 1;
        d.x_s1 = h_bs; d.y_s1 = g_bs; d.accumulate = 1;
        DVD_TRY(dvd_conv_fwd(&d, rh_t, ws.who, nullptr, nullptr, g_t + 2 * chw, ss));
      }
      { ProfScope ps(3, "gru_out", ss); gru_out_kernel<<<eb, 256, 0, ss>>>(g_t, g_bs, hp, hp_bs, h_t, h_bs, Bn, chw); }
      DVD_LAUNCH_CHECK();
    }
    have_planes = fused && !first;
  }
  DVD_TRY(fk.end());
  return 0;
}

extern "C" int dvd_convgru_layer_fwd(const float* x, int64_t x_bs, int64_t x_ts, const float* h0, const float* wu,
                                     const float* wr, const float* wo, const float* bu, const float* br,
                                     const float* bo, float* gates, float* h, float* rh, int B, int T, int Cx, int Ch,
                                     int H, int W, int k, void* workspace, size_t ws_bytes, void* stream) {
  return gru_fwd_impl(x, x_bs, x_ts, h0, wu, wr, wo, bu, br, bo, gates, h, rh, B, T, Cx, Ch, H, W, k, 0, T, false,
                      workspace, ws_bytes, stream);
}

extern "C" int dvd_convgru_layer_fwd_range(const float* x, int64_t x_bs, int64_t x_ts, const float* h0, const float* wu,
                                           const float* wr, const float* wo, const float* bu, const float* br,
                                           const float* bo, float* gates, float* h, float* rh, int B, int T, int Cx,
                                           int Ch, int H, int W, int k, int t0, int t1, void* workspace, size_t ws_bytes,
                                           void* stream) {
  return gru_fwd_impl(x, x_bs, x_ts, h0, wu, wr, wo, bu, br, bo, gates, h, rh, B, T, Cx, Ch, H, W, k, t0, t1, true,
                      workspace, ws_bytes, stream);
}

// Frames [t0, t1) of the BPTT sweep, latest frame first.  range = false: the whole clip in one call.  range = true: one
// of several calls that share `workspace` (dvd_convgru_layer_range_workspace_bytes) and visit the frames in descending
// order: the call with t1 = T prepares the weights, every call leaves dx of its frames, the call with t0 = 0 finishes
// with dh0, the weight and bias gradients.
static int gru_bwd_impl(const float* x, int64_t x_bs, int64_t x_ts, const float* h0, const float* wu, const float* wr,
                        const float* wo, float* gates, const float* h, const float* rh, const float* dh, float* dx,
                        float* dh0, float* dwu, float* dwr, float* dwo, float* dbu, float* dbr, float* dbo, int B, int T,
                        int Cx, int Ch, int H, int W, int k, int t0, int t1, bool range, void* workspace,
                        size_t ws_bytes, void* stream) {
  DVD_CHECK_ARG(x && wu && wr && wo && gates && h && rh && dh && dx && dwu && dwr && dwo && dbu && dbr && dbo);
  DVD_CHECK_ARG(workspace && B > 0 && T > 0 && Cx > 0 && Ch > 0 && H > 0 && W > 0 && (k & 1));
  DVD_CHECK_ARG(dh0 == nullptr || h0 != nullptr);
  DVD_CHECK_ARG(0 <= t0 && t0 < t1 && t1 <= T);
  const int HW = H * W, taps = k * k, Ct = Cx + Ch;
  GruWs ws = carve(workspace, B, Cx, Ch, HW, taps, true, range ? T : 0);
  DVD_CHECK_ARG(ws.bytes <= ws_bytes);
  cudaStream_t st = as_stream(stream);
  const bool begin = t1 == T, finish = t0 == 0;
  const float* wsrc[3] = {wu, wr, wo};
  float* dwdst[3] = {dwu, dwr, dwo};
  float* dbdst[3] = {dbu, dbr, dbo};
  // dgrad operands (transposed + flipped): [tap'][gate rows (u|r|o)][ci]
  for (int g = 0; g < 3 && begin; ++g) {
    DVD_TRY(dvd_weight_pack(wsrc[g], Ct, taps, 0, Ch, 0, Cx, nullptr, 1, ws.wxT, 3 * Ch, g * Ch, Cx, 0, stream));
    if (g < 2) DVD_TRY(dvd_weight_pack(wsrc[g], Ct, taps, 0, Ch, Cx, Ch, nullptr, 1, ws.whurT, 2 * Ch, g * Ch, Ch, 0, stream));
    else DVD_TRY(dvd_weight_pack(wsrc[g], Ct, taps, 0, Ch, Cx, Ch, nullptr, 1, ws.whoT, Ch, 0, Ch, 0, stream));
  }
  const int64_t chw = (int64_t)Ch * HW;
  const int64_t g_ts = 3 * chw, g_bs = (int64_t)T * g_ts, h_ts = chw, h_bs = (int64_t)T * chw;
  // the two dgrad GEMMs of a step read the same transposed weights every step: split them into planes once
  auto step_descs = [&](int Bn, dvd_conv_desc* rhd, dvd_conv_desc* hpd) {
    *rhd = base_desc(Bn, 1, Ch, Ch, H, W, k); *hpd = base_desc(Bn, 1, 2 * Ch, Ch, H, W, k);
    rhd->x_s1 = hpd->x_s1 = g_bs; rhd->y_s1 = hpd->y_s1 = chw; hpd->accumulate = 1;
  };
  dvd_conv_desc d_rh, d_hp;
  const bool can_planes = gru_fused_enabled() && T > 1;
  // chains over batch slices (see SliceFork); a caller of the frame-range API overlaps whole layers instead
  const int ns = range ? 1 : pick_slices(B, T, [&](int Bn) {
    step_descs(Bn, &d_rh, &d_hp);
    return can_planes && conv_fwd_ex_eligible(&d_rh) && conv_fwd_ex_eligible(&d_hp);
  });
  const int Bn = B / ns;
  step_descs(Bn, &d_rh, &d_hp);
  const int eb = ew_blocks((int64_t)Bn * chw);
  const int ChP = tma_round64(Ch), Co2P = tma_round64(2 * Ch);
  const bool wplanes = can_planes && conv_fwd_ex_eligible(&d_rh) && conv_fwd_ex_eligible(&d_hp);
  if (wplanes && begin) {
    DVD_TRY(tma_split_weights(ws.whoT, taps, Ch, Ch, ChP, 0, ws.whoT_hi, ws.whoT_lo, st));
    DVD_TRY(tma_split_weights(ws.whurT, taps, 2 * Ch, Ch, ChP, 0, ws.whurT_hi, ws.whurT_lo, st));
  }
  // The pre-activation gradients (da_u | da_r | da_o) of all frames feed four GEMMs (x-dgrad and the three weight
  // gradients): split them into bf16 planes ONCE and hand the planes to all four (channel / frame offsets in the
  // TMA coordinates) instead of re-splitting the largest tensor of the layer per GEMM.
  struct PlanesGuard {
    void* p = nullptr; cudaStream_t st;
    ~PlanesGuard() { tma_scratch_free(p, st); }
  } gp;
  gp.st = st;
  TmaWgOperands yo;
  const int G3P = tma_round64(3 * Ch);
  dvd_conv_desc d_dx = base_desc(B, T, 3 * Ch, Cx, H, W, k);
  d_dx.x_s1 = g_bs; d_dx.x_s2 = g_ts; d_dx.y_s1 = (int64_t)T * Cx * HW; d_dx.y_s2 = (int64_t)Cx * HW;
  dvd_conv_desc d_wx = base_desc(B, T, Cx, 3 * Ch, H, W, k); d_wx.x_kind = 1;
  d_wx.x_s1 = x_bs; d_wx.x_s2 = x_ts; d_wx.y_s1 = g_bs; d_wx.y_s2 = g_ts;
  const bool share_on = get_option(OPT_GRU_SHARE_PLANES) != 0;
  yo.y_Cp = G3P; yo.y_T = T;
  bool share = share_on && Ch % 32 == 0 && conv_fwd_ex_eligible(&d_dx);
  if (share) {
    yo.y_hi = reinterpret_cast<void*>(1);                     // (eligibility only looks at the geometry)
    share = conv_wgrad_ex_eligible(&d_wx, &yo);
    yo.y_hi = nullptr;
  }
  // gplanes: the BPTT kernels write the planes themselves (frame by frame) and the per-step dgrad GEMMs read their
  // A operand from them too; otherwise the planes are split from the fp32 buffer after the sweep
  const bool gplanes_on = get_option(OPT_GRU_BWD_PLANES) != 0;
  const bool gplanes = gplanes_on && share && wplanes && Ch % 64 == 0;
  __nv_bfloat16 *pl_hi = nullptr, *pl_lo = nullptr;
  const int64_t pl_frame = (int64_t)HW * G3P, pl_img = (int64_t)T * pl_frame;
  if (share) {
    const size_t elems = (size_t)B * T * HW * G3P;
    void* base = ws.gplanes;          // frame-range calls: the planes live in the workspace from call to call
    if (!range) {
      DVD_TRY(tma_scratch_alloc(&gp.p, 2 * elems * sizeof(uint16_t) + 256, st));
      base = gp.p;
    }
    yo.y_hi = base;
    yo.y_lo = reinterpret_cast<uint16_t*>(base) + elems;
    pl_hi = reinterpret_cast<__nv_bfloat16*>(base);
    pl_lo = pl_hi + elems;
  }
  // Fused BPTT (option "gru_bwd_fused", off by default): the gate-gradient math between the two dgrad GEMMs of a step,
  // and between one step and the next, runs in their epilogues (GruEpi modes 3 / 4): two launches per step instead of
  // four.  Only where the GEMMs fill the machine without split-K (the small early stages keep the separate kernels and
  // their split-K dgrads).  Measured on config 2 (profiles/r2): helpers 135 -> 102 ms, fwd/dgrad GEMMs 1026 -> 1066 ms,
  // step unchanged -- with one accumulator set per CTA pair the tensor pipe waits for the longer epilogue.
  const bool bptt_fused = gplanes && get_option(OPT_GRU_BWD_FUSED) &&
                          (int64_t)ceil_div(Bn * HW, 128) * ceil_div(Ch, 128) >= num_sms();
  SliceFork fk(st, ns);
  DVD_TRY(fk.begin());
  // carry buffers ping-pong with the step: step t writes carry[(T-1-t) & 1] and reads the other one
  float* const carry_buf[2] = {ws.carry0, ws.carry1};
  // the elementwise part 1 of a step has not been done by the previous step's epilogue (mode 4).  The last step of a
  // call never runs mode 4: it would read dh of frame t0 - 1, which a frame-range caller has not produced yet.
  bool need_k1 = true;
  for (int t = t1 - 1; t >= t0; --t) {
    const bool first = t == 0 && !h0;
    const int64_t hp_bs = t > 0 ? h_bs : chw;
    const dim3 pgrid(ceil_div(HW, 32), Ch / 64, Bn);
    bool did_k1_next = false;
    for (int sl = 0; sl < ns; ++sl) {      // step t of every chain, interleaved
      cudaStream_t ss = fk.stream(sl);
      const int b0 = sl * Bn;
      const int64_t co = (int64_t)b0 * chw;                       // this slice's rows of the (B, Ch*HW) buffers
      const float* hp = first ? nullptr : (t > 0 ? h + (int64_t)(t - 1) * h_ts : h0) + (int64_t)b0 * hp_bs;
      float* g_t = gates + (int64_t)t * g_ts + (int64_t)b0 * g_bs;
      const float* dh_t = dh + (int64_t)t * h_ts + (int64_t)b0 * h_bs;
      float* carry_out = carry_buf[(T - 1 - t) & 1] + co;
      float* carry_next = carry_buf[(T - t) & 1] + co;
      const float* carry_in = t == T - 1 ? nullptr : carry_next;
      float* d_rh_s = ws.d_rh + co;
      __nv_bfloat16* p_hi = pl_hi ? pl_hi + t * pl_frame + (int64_t)b0 * pl_img : nullptr;
      __nv_bfloat16* p_lo = pl_lo ? pl_lo + t * pl_frame + (int64_t)b0 * pl_img : nullptr;
      if (need_k1) {
        if (gplanes) {
          ProfScope ps(3, "gru_bwd1", ss);
          gru_bwd1_planes_kernel<<<pgrid, 256, 0, ss>>>(g_t, g_bs, hp, hp_bs, dh_t, h_bs, carry_in, carry_out, Ch, HW,
                                                         p_hi, p_lo, pl_img, G3P);
        } else {
          ProfScope ps(3, "gru_bwd1", ss);
          gru_bwd1_kernel<<<eb, 256, 0, ss>>>(g_t, g_bs, hp, hp_bs, dh_t, h_bs, carry_in, carry_out, Bn, chw);
        }
        DVD_LAUNCH_CHECK();
      }
      if (hp && bptt_fused) {
        // d(rh) = conv_o^T(da_o) with the reset-gate gradient in the epilogue (mode 3)
        TmaOperands op;
        op.w_hi = ws.whoT_hi; op.w_lo = ws.whoT_lo; op.CoutP = ChP;
        op.a_hi = p_hi; op.a_lo = p_lo;
        op.a_Cp = G3P; op.a_c_off = 2 * Ch; op.a_img_stride = pl_img;
        GruEpi ge;
        ge.mode = 3; ge.Ch = Ch; ge.hprev = hp; ge.hp_s1 = hp_bs; ge.ugate = g_t; ge.u_s1 = g_bs;
        ge.out2 = carry_out; ge.o2_s1 = chw;
        ge.pl_hi = p_hi; ge.pl_lo = p_lo; ge.pl_Cp = G3P; ge.pl_img = pl_img;
        DVD_TRY(conv_fwd_ex(&d_rh, g_t + 2 * chw, ws.whoT, carry_out, &op, &ge, ss));
        // dh_{t-1} += conv_ur^T(da_u | da_r); up to the last step of the call the epilogue (mode 4) goes straight on with
        // step t-1's part 1
        op.w_hi = ws.whurT_hi; op.w_lo = ws.whurT_lo; op.a_c_off = 0;
        if (t > t0) {
          dvd_conv_desc d4 = d_hp;
          d4.accumulate = 0;
          GruEpi g4;
          g4.mode = 4; g4.Ch = Ch; g4.ugate = g_t - g_ts; g4.u_s1 = g_bs;
          g4.hprev = t > 1 ? h + (int64_t)(t - 2) * h_ts + (int64_t)b0 * h_bs : (h0 ? h0 + (int64_t)b0 * chw : nullptr);
          g4.hp_s1 = t > 1 ? h_bs : chw;
          g4.carry_in = carry_out; g4.out2 = carry_next; g4.o2_s1 = chw;
          g4.dh_prev = dh_t - h_ts; g4.dh_s1 = h_bs;
          g4.pl_hi = p_hi - pl_frame; g4.pl_lo = p_lo - pl_frame; g4.pl_Cp = G3P; g4.pl_img = pl_img;
          DVD_TRY(conv_fwd_ex(&d4, g_t, ws.whurT, carry_next, &op, &g4, ss));
          did_k1_next = true;
        } else {
          DVD_TRY(conv_fwd_ex(&d_hp, g_t, ws.whurT, carry_out, &op, nullptr, ss));
        }
      } else {
        if (hp) {                                                 // d(rh) = conv_o^T(da_o), h-half
          if (wplanes) {
            TmaOperands op;
            op.w_hi = ws.whoT_hi; op.w_lo = ws.whoT_lo; op.CoutP = ChP;
            if (gplanes) {
              op.a_hi = p_hi; op.a_lo = p_lo;
              op.a_Cp = G3P; op.a_c_off = 2 * Ch; op.a_img_stride = pl_img;
            }
            DVD_TRY(conv_fwd_ex(&d_rh, g_t + 2 * chw, ws.whoT, d_rh_s, &op, nullptr, ss));
          } else {
            DVD_TRY(dvd_conv_fwd(&d_rh, g_t + 2 * chw, ws.whoT, nullptr, nullptr, d_rh_s, ss));
          }
        }
        if (gplanes) {
          ProfScope ps(3, "gru_bwd2", ss);
          gru_bwd2_planes_kernel<<<pgrid, 256, 0, ss>>>(g_t, g_bs, hp, hp_bs, d_rh_s, carry_out, Ch, HW, p_hi, p_lo,
                                                         pl_img, G3P);
        } else {
          ProfScope ps(3, "gru_bwd2", ss);
          gru_bwd2_kernel<<<eb, 256, 0, ss>>>(g_t, g_bs, hp, hp_bs, d_rh_s, carry_out, Bn, chw);
        }
        DVD_LAUNCH_CHECK();
        if (hp) {                                                 // dh_prev += conv_u^T(da_u) + conv_r^T(da_r)
          if (wplanes) {
            TmaOperands op;
            op.w_hi = ws.whurT_hi; op.w_lo = ws.whurT_lo; op.CoutP = ChP;
            if (gplanes) {
              op.a_hi = p_hi; op.a_lo = p_lo;
              op.a_Cp = G3P; op.a_c_off = 0; op.a_img_stride = pl_img;
            }
            DVD_TRY(conv_fwd_ex(&d_hp, g_t, ws.whurT, carry_out, &op, nullptr, ss));
          } else {
            DVD_TRY(dvd_conv_fwd(&d_hp, g_t, ws.whurT, nullptr, nullptr, carry_out, ss));
          }
        }
      }
    }
    need_k1 = !did_k1_next;
  }
  DVD_TRY(fk.end());
  const float* carry_in = carry_buf[(T - 1) & 1];          // what step 0 wrote: dL/dh_{-1}
  if (dh0 && finish) DVD_CUDA(cudaMemcpyAsync(dh0, carry_in, sizeof(float) * (size_t)B * chw, cudaMemcpyDeviceToDevice, st));
  if (share && !gplanes && finish)
    DVD_TRY(tma_split_gradients(gates, B * T, 3 * Ch, g_ts, HW, HW, const_cast<void*>(yo.y_hi), const_cast<void*>(yo.y_lo), st));
  // dx of the frames of this call: one implicit GEMM over the (da_u | da_r | da_o) buffer
  if (share && begin && finish) {
    TmaOperands op;
    op.a_hi = yo.y_hi; op.a_lo = yo.y_lo;
    DVD_TRY(conv_fwd_ex(&d_dx, gates, ws.wxT, dx, &op, nullptr, st));
  } else {
    // (a frame range of the [B][T] planes is not one strided run of images: this GEMM splits its operand itself)
    dvd_conv_desc d = d_dx;
    d.N2 = t1 - t0;
    DVD_TRY(dvd_conv_fwd(&d, gates + (int64_t)t0 * g_ts, ws.wxT, nullptr, nullptr, dx + (int64_t)t0 * d_dx.y_s2, stream));
  }
  if (!finish) return 0;
  // weight gradients, batched over time
  if (share) {
    yo.y_c_off = 0; yo.y_t_off = 0;
    DVD_TRY(conv_wgrad_ex(&d_wx, x, ws.dwx, &yo, st));
  } else {
    DVD_TRY(dvd_conv_wgrad(&d_wx, x, gates, ws.dwx, stream));
  }
  {
    bool have = false;
    if (T > 1) {
      dvd_conv_desc d = base_desc(B, T - 1, Ch, 2 * Ch, H, W, k);   // pairs (h_{t-1}, da_t), t = 1..T-1
      d.x_kind = 1;
      d.x_s1 = h_bs; d.x_s2 = h_ts; d.y_s1 = g_bs; d.y_s2 = g_ts;
      yo.y_c_off = 0; yo.y_t_off = 1;
      if (share && conv_wgrad_ex_eligible(&d, &yo)) DVD_TRY(conv_wgrad_ex(&d, h, ws.dwhur, &yo, st));
      else DVD_TRY(dvd_conv_wgrad(&d, h, gates + g_ts, ws.dwhur, stream));
      have = true;
    }
    if (h0) {
      dvd_conv_desc d = base_desc(B, 1, Ch, 2 * Ch, H, W, k); d.x_kind = 1;
      d.x_s1 = chw; d.y_s1 = g_bs; d.accumulate = have ? 1 : 0;
      DVD_TRY(dvd_conv_wgrad(&d, h0, gates, ws.dwhur, stream));
      have = true;
    }
    if (!have) DVD_CUDA(cudaMemsetAsync(ws.dwhur, 0, sizeof(float) * (size_t)taps * Ch * 2 * Ch, st));
  }
  {
    dvd_conv_desc d = base_desc(B, T, Ch, Ch, H, W, k);             // pairs (rh_t, da_o,t)
    d.x_kind = 1;
    d.x_s1 = h_bs; d.x_s2 = h_ts; d.y_s1 = g_bs; d.y_s2 = g_ts;
    yo.y_c_off = 2 * Ch; yo.y_t_off = 0;
    if (share && conv_wgrad_ex_eligible(&d, &yo)) DVD_TRY(conv_wgrad_ex(&d, rh, ws.dwho, &yo, st));
    else DVD_TRY(dvd_conv_wgrad(&d, rh, gates + 2 * chw, ws.dwho, stream));
  }
  for (int g = 0; g < 3; ++g) {
    DVD_TRY(dvd_weight_unpack(ws.dwx, 3 * Ch, g * Ch, Ct, taps, 0, Ch, 0, Cx, 0, dwdst[g], stream));
    if (g < 2) DVD_TRY(dvd_weight_unpack(ws.dwhur, 2 * Ch, g * Ch, Ct, taps, 0, Ch, Cx, Ch, 0, dwdst[g], stream));
    else DVD_TRY(dvd_weight_unpack(ws.dwho, Ch, 0, Ct, taps, 0, Ch, Cx, Ch, 0, dwdst[g], stream));
  }
  DVD_TRY(dvd_channel_sum(gates, B * T, 3 * Ch, HW, g_ts, 0, ws.dbias, ws.dscratch, stream));
  for (int g = 0; g < 3; ++g)
    DVD_CUDA(cudaMemcpyAsync(dbdst[g], ws.dbias + g * Ch, sizeof(float) * Ch, cudaMemcpyDeviceToDevice, st));
  return 0;
}

extern "C" int dvd_convgru_layer_bwd(const float* x, int64_t x_bs, int64_t x_ts, const float* h0, const float* wu,
                                     const float* wr, const float* wo, float* gates, const float* h, const float* rh,
                                     const float* dh, float* dx, float* dh0, float* dwu, float* dwr, float* dwo,
                                     float* dbu, float* dbr, float* dbo, int B, int T, int Cx, int Ch, int H, int W,
                                     int k, void* workspace, size_t ws_bytes, void* stream) {
  return gru_bwd_impl(x, x_bs, x_ts, h0, wu, wr, wo, gates, h, rh, dh, dx, dh0, dwu, dwr, dwo, dbu, dbr, dbo, B, T, Cx,
                      Ch, H, W, k, 0, T, false, workspace, ws_bytes, stream);
}

extern "C" int dvd_convgru_layer_bwd_range(const float* x, int64_t x_bs, int64_t x_ts, const float* h0, const float* wu,
                                           const float* wr, const float* wo, float* gates, const float* h,
                                           const float* rh, const float* dh, float* dx, float* dh0, float* dwu,
                                           float* dwr, float* dwo, float* dbu, float* dbr, float* dbo, int B, int T,
                                           int Cx, int Ch, int H, int W, int k, int t0, int t1, void* workspace,
                                           size_t ws_bytes, void* stream) {
  return gru_bwd_impl(x, x_bs, x_ts, h0, wu, wr, wo, gates, h, rh, dh, dx, dh0, dwu, dwr, dwo, dbu, dbr, dbo, B, T, Cx,
                      Ch, H, W, k, t0, t1, true, workspace, ws_bytes, stream);
}
