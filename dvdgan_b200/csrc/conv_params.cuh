// Shared between the SIMT (conv.cu) and TMA + tcgen05 (conv_tma.cu) convolution engines.
#pragma once
#include "common.cuh"

namespace dvd {

struct ConvP {
  dvd_conv_desc d;
  const float* x;
  const float* w;
  const float* bias;
  const float* res;
  float* y;            // forward: output; wgrad: dY (read only)
  int M, DHW, HW, taps, Hs, Ws, ck, iters_total, iters_per_split, nsplit, vecB, vecY, atomic_out;
};

// zero-fill the (possibly strided) output view described by p (split-K without accumulate)
int zero_output_view(const ConvP& p, cudaStream_t st);

// Optional extras of the TMA + tcgen05 forward engine (library-internal, used by the ConvGRU time loop).
// Pre-split operand planes: activations [N][D*H*W][CinP] and weights [taps][CoutP][CinP], each as a hi and a lo
// plane of 16-bit floats (fp16 with lo scaled by 2^11 for forward operands, bf16 otherwise; see conv_tma.cu).
struct TmaOperands {
  const void* a_hi = nullptr;     // nullptr: split p.x inside the call
  const void* a_lo = nullptr;
  const void* w_hi = nullptr;     // nullptr: split p.w inside the call
  const void* w_lo = nullptr;
  int CoutP = 0;                  // rows per tap of the weight planes
  // a_hi / a_lo may be a window of larger shared planes: channels [a_c_off, a_c_off + Cin) of pixels that hold a_Cp
  // channels each, image n at element offset n * a_img_stride from a_hi (0 / 0 / 0: dense [N][D*H*W][round64(Cin)])
  int a_Cp = 0;
  int a_c_off = 0;
  int64_t a_img_stride = 0;
};
// ConvGRU epilogues (ConvGRU.py:47-52) applied to v = accumulator + y (y holds the x-half pre-activations):
//  mode 1 (h-half of update|reset, Cout = 2*Ch): co <  Ch: y = u = sigmoid(v)
//                                                co >= Ch: y = r = sigmoid(v); out2 = r * hprev  (+ planes of it)
//  mode 2 (h-half of out, Cout = Ch):            y = o = tanh(v); out2 = hprev * (1 - u) + o * u  (+ planes of it)
// hprev / ugate / out2 are (B, Ch, H, W) views with batch strides hp_s1 / u_s1 / o2_s1 and channel stride H*W;
// the planes are dense [B*H*W][pl_Cp] in the forward plane format, the A operand of the next h-half GEMM.
// BPTT epilogues of the two per-step dgrad GEMMs; v = accumulator, gates hold (u | r | o) activated:
//  mode 3 (d(rh) = conv_o^T(da_o), Cout = Ch):  carry += v * r;  da_r = v * hprev * r * (1 - r)  -> reset slot of the
//         frame's gates (fp32) and channels [Ch, 2Ch) of its bf16 gradient planes.   gate = ugate, carry = out2.
//  mode 4 (conv_ur^T(da_u | da_r), Cout = Ch):  dhn = dh_prev + carry_in + v  is the whole gradient of h_{t-1}; the
//         elementwise part of step t-1 follows at once:  da_o = dhn * u * (1 - o^2),  da_u = dhn * (o - h2) * u * (1 - u)
//         -> update / out slots of frame t-1's gates and channels [0, Ch) / [2Ch, 3Ch) of its planes;
//         carry_next = dhn * (1 - u).   gate(t-1) = ugate, h2 = hprev (h_{t-2}, may be null), carry_next = out2.
// planes: frame image b at pl + b * pl_img + pix * pl_Cp (channels-last, pl_Cp = round64(3Ch)).
struct GruEpi {
  int mode = 0;
  int Ch = 0;
  const float* hprev = nullptr;
  int64_t hp_s1 = 0;
  const float* ugate = nullptr;
  int64_t u_s1 = 0;
  float* out2 = nullptr;
  int64_t o2_s1 = 0;
  void* pl_hi = nullptr;
  void* pl_lo = nullptr;
  int pl_Cp = 0;
  // modes 3 / 4 only
  int64_t pl_img = 0;
  const float* carry_in = nullptr;     // mode 4: carry after mode 3 of the same step (batch stride o2_s1)
  const float* dh_prev = nullptr;      // mode 4: external gradient of h_{t-1}
  int64_t dh_s1 = 0;
};
bool tma_fwd_launch_ex_eligible(const ConvP& p);
int tma_fwd_launch_ex(ConvP& p, const TmaOperands* ops, const GruEpi* epi, cudaStream_t st);
// split fp32 packed weights [taps][Cin][Cout] / activations into (hi, lo) planes; fp16 = 1: the forward format
int tma_split_weights(const float* w_packed, int taps, int Cin, int Cout, int CoutP, int fp16, void* hi, void* lo,
                      cudaStream_t st);
int tma_split_activations(const float* x, int N, int C, int64_t n_stride, int64_t c_stride, int pix, int fp16, void* hi,
                          void* lo, cudaStream_t st);
bool tma_forward_planes_fp16();   // forward operands use fp16 planes (default; options "fwd_bf16" / "oneacc" turn it off)
int tma_saturation_count(unsigned int* count, int reset, cudaStream_t st);
inline int tma_round64(int c) { return (c + 63) / 64 * 64; }
// stream-ordered scratch (the engine's per-stream cudaMallocAsync pools); free with tma_scratch_free on the same stream
int tma_scratch_alloc(void** p, size_t bytes, cudaStream_t st);
void tma_scratch_free(void* p, cudaStream_t st);
int tma_scratch_stats(long long* high_water, long long* reserved);

// Pre-split dY planes for the weight-gradient engine: bf16 (hi, lo) planes [clips*T][D*H*W][Cp] of a (clips, T, C)
// tensor; the conv's dY is channels [c_off, c_off + Cout) of frames [t_off, t_off + d.N2) of every clip (d.N1 clips).
struct TmaWgOperands {
  const void* y_hi = nullptr;
  const void* y_lo = nullptr;
  int y_Cp = 0;        // channels per pixel in the planes (multiple of 64)
  int y_c_off = 0;     // first channel of this conv's dY (multiple of 64)
  int y_T = 0;         // frames per clip in the planes
  int y_t_off = 0;     // first frame used
};
bool conv_wgrad_ex_eligible(const dvd_conv_desc* d, const TmaWgOperands* ops);
int conv_wgrad_ex(const dvd_conv_desc* d, const float* x, float* dwp, const TmaWgOperands* ops, cudaStream_t st);
int tma_split_gradients(const float* g, int N, int C, int64_t n_stride, int64_t c_stride, int pix, void* hi, void* lo,
                        cudaStream_t st);     // bf16 planes [N][pix][round64(C)]
bool conv_fwd_ex_eligible(const dvd_conv_desc* d);
int conv_fwd_ex(const dvd_conv_desc* d, const float* x, const float* w_packed, float* y, const TmaOperands* ops,
                const GruEpi* epi, cudaStream_t st);

// TMA + tcgen05 path (conv_tma.cu)
bool tma_fwd_eligible(const ConvP& p);
int tma_fwd_launch(ConvP& p, cudaStream_t st);
bool tma_wgrad_eligible(const ConvP& p);
int tma_wgrad_launch(ConvP& p, float* dwp, cudaStream_t st);
bool tma_wgrad_ex_ok(const ConvP& p, const TmaWgOperands* ops);
int tma_wgrad_launch_ex(ConvP& p, float* dwp, const TmaWgOperands* ops, cudaStream_t st);

}  // namespace dvd
