// Shared between the SIMT (conv.cu) and tcgen05 (conv_tc.cu) convolution engines.
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

// tcgen05 path (conv_tc.cu)
bool tc_fwd_eligible(const ConvP& p);
int tc_fwd_launch(ConvP& p, cudaStream_t st);
bool tc_wgrad_eligible(const ConvP& p);
int tc_wgrad_launch(ConvP& p, float* dwp, cudaStream_t st);

// TMA + tcgen05 path (conv_tma.cu)
bool tma_fwd_eligible(const ConvP& p);
int tma_fwd_launch(ConvP& p, cudaStream_t st);
bool tma_wgrad_eligible(const ConvP& p);
int tma_wgrad_launch(ConvP& p, float* dwp, cudaStream_t st);

}  // namespace dvd
