// GPU side of the input pipeline (reference: Dataloader/datasets/ucf101.py:177-199 driven by the transforms main.py:33-56
// composes): decoded uint8 frames -> crop -> PIL-exact bilinear resize -> horizontal flip -> ToTensor / Normalize ->
// the (B, C, T, H, W) float clip the trainer consumes.  One launch per batch instead of B*T PIL calls on the host.
//
// The resize reproduces Pillow's ImagingResample for 8-bit images bit for bit: two passes (horizontal, then vertical),
// each a weighted sum with 22-bit fixed-point coefficients, the intermediate image rounded back to uint8.  The
// coefficient tables depend only on (crop size, output size); the host computes them in double precision exactly as
// Pillow's precompute_coeffs / normalize_coeffs_8bpc do (dvdgan_b200/data.py) and the kernel only does integer math.
#include "common.cuh"

namespace dvd {

constexpr int kPrecisionBits = 32 - 8 - 2;      // Pillow: PRECISION_BITS

__device__ __forceinline__ int clip8(int v) {
  v >>= kPrecisionBits;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// grid (OH, T, B); block: 64 x 3 threads per pass chunk.  Shared: the horizontally resampled rows this output row needs.
__global__ void __launch_bounds__(192) clip_transform_kernel(
    const uint8_t* __restrict__ frames, int T, int Hs, int Ws, const int* __restrict__ box, const int* __restrict__ flip,
    const int* __restrict__ xb, const int* __restrict__ xk, int xks, const int* __restrict__ yb,
    const int* __restrict__ yk, int yks, int OH, int OW, float norm_value, float m0, float m1, float m2, float s0,
    float s1, float s2, float* __restrict__ out) {
  extern __shared__ uint8_t rows[];               // [yks][OW][3]
  const int oy = blockIdx.x, t = blockIdx.y, b = blockIdx.z;
  const int x0 = box[4 * b], y0 = box[4 * b + 1];
  const uint8_t* src = frames + ((int64_t)b * T + t) * Hs * Ws * 3;
  const int ymin = yb[((int64_t)b * OH + oy) * 2], ycnt = yb[((int64_t)b * OH + oy) * 2 + 1];
  const int* kyrow = yk + ((int64_t)b * OH + oy) * yks;
  // horizontal pass for the ycnt source rows of this output row (Pillow: ImagingResampleHorizontal_8bpc)
  for (int i = threadIdx.x; i < ycnt * OW * 3; i += blockDim.x) {
    const int c = i % 3, ox = (i / 3) % OW, r = i / (3 * OW);
    const int xmin = xb[((int64_t)b * OW + ox) * 2], xcnt = xb[((int64_t)b * OW + ox) * 2 + 1];
    const int* k = xk + ((int64_t)b * OW + ox) * xks;
    const uint8_t* p = src + ((int64_t)(y0 + ymin + r) * Ws + (x0 + xmin)) * 3 + c;
    int ss = 1 << (kPrecisionBits - 1);
    for (int j = 0; j < xcnt; ++j) ss += (int)p[3 * j] * k[j];
    rows[(r * OW + ox) * 3 + c] = (uint8_t)clip8(ss);
  }
  __syncthreads();
  // vertical pass (ImagingResampleVertical_8bpc), flip, ToTensor(norm_value), Normalize(mean, std)
  const int fl = flip[b];
  for (int i = threadIdx.x; i < OW * 3; i += blockDim.x) {
    const int ox = i % OW, c = i / OW;
    int ss = 1 << (kPrecisionBits - 1);
    for (int r = 0; r < ycnt; ++r) ss += (int)rows[(r * OW + ox) * 3 + c] * kyrow[r];
    const float px = (float)clip8(ss);
    const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2), sd = c == 0 ? s0 : (c == 1 ? s1 : s2);
    // img.float().div(norm_value).sub_(mean).div_(std): three correctly rounded fp32 operations, no contraction
    const float v = __fdiv_rn(__fsub_rn(__fdiv_rn(px, norm_value), mean), sd);
    const int dx = fl ? OW - 1 - ox : ox;
    out[((((int64_t)b * 3 + c) * T + t) * OH + oy) * OW + dx] = v;
  }
}

}  // namespace dvd

using namespace dvd;

extern "C" int dvd_clip_transform(const uint8_t* frames, int B, int T, int Hs, int Ws, const int* box, const int* flip,
                                  const int* xb, const int* xk, int xks, const int* yb, const int* yk, int yks, int OH,
                                  int OW, float norm_value, const float* mean3, const float* std3, float* out,
                                  void* stream) {
  dvd::ProfScope _ps(3, "clip_transform", dvd::as_stream(stream));
  DVD_CHECK_ARG(frames && box && flip && xb && xk && yb && yk && mean3 && std3 && out);
  DVD_CHECK_ARG(B > 0 && B <= 65535 && T > 0 && T <= 65535 && Hs > 0 && Ws > 0 && OH > 0 && OW > 0 && xks > 0 && yks > 0);
  DVD_CHECK_ARG(norm_value > 0.f);
  const size_t smem = (size_t)yks * OW * 3;
  DVD_CHECK_ARG(smem <= 48 * 1024);
  dim3 grid(OH, T, B);
  clip_transform_kernel<<<grid, 192, smem, as_stream(stream)>>>(frames, T, Hs, Ws, box, flip, xb, xk, xks, yb, yk, yks, OH,
                                                               OW, norm_value, mean3[0], mean3[1], mean3[2], std3[0],
                                                               std3[1], std3[2], out);
  DVD_LAUNCH_CHECK();
  return 0;
}
