// Attention core: S = Q^T K, row softmax, O = V A^T and the matching backward, built on the
// strided-batched SGEMM plus warp-per-row softmax kernels.  Layouts follow the reference's views
// (channel-major q/k/v/out); SeparableAttnCell's raw `.view` gives a token-major q.
#include "common.cuh"

namespace dvd {

// in-place softmax over the last dim of a (rows, n) matrix; one warp per row
__global__ void softmax_rows_kernel(float* __restrict__ a, int64_t rows, int n) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float* r = a + row * n;
  float mx = -INFINITY;
  for (int j = lane; j < n; j += 32) mx = fmaxf(mx, r[j]);
  mx = warp_max(mx);
  float s = 0.f;
  for (int j = lane; j < n; j += 32) {
    const float e = expf(r[j] - mx);
    r[j] = e;
    s += e;
  }
  s = warp_sum(s);
  const float inv = 1.f / s;
  for (int j = lane; j < n; j += 32) r[j] *= inv;
}

// dS = A * (dA - rowsum(dA * A)), written over dA
__global__ void softmax_bwd_rows_kernel(const float* __restrict__ a, float* __restrict__ da, int64_t rows, int n) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* ar = a + row * n;
  float* dr = da + row * n;
  float s = 0.f;
  for (int j = lane; j < n; j += 32) s = fmaf(ar[j], dr[j], s);
  s = warp_sum(s);
  for (int j = lane; j < n; j += 32) dr[j] = ar[j] * (dr[j] - s);
}

}  // namespace dvd

using namespace dvd;

extern "C" int dvd_attn_fwd(const float* q, int64_t q_bs, const float* k, int64_t k_bs, const float* v, int64_t v_bs,
                            float* attn, float* out, int64_t o_bs, int batch, int dq, int dv, int Nq, int Nk,
                            int q_token_major, void* stream) {
  dvd::ProfScope _ps(3, "attn_fwd", dvd::as_stream(stream));
  DVD_CHECK_ARG(q && k && v && attn && out && batch > 0 && dq > 0 && dv > 0 && Nq > 0 && Nk > 0);
  const int64_t a_bs = (int64_t)Nq * Nk;
  for (int b0 = 0; b0 < batch; b0 += 65535) {
    const int nb = batch - b0 < 65535 ? batch - b0 : 65535;
    // S[i][j] = sum_c q[c][i] k[c][j]
    DVD_TRY(dvd_bgemm(q_token_major ? 0 : 1, 0, Nq, Nk, dq, 1.f, q + b0 * q_bs, q_token_major ? dq : Nq, q_bs,
                      k + b0 * k_bs, Nk, k_bs, 0.f, attn + b0 * a_bs, Nk, a_bs, nb, nullptr, stream));
  }
  const int64_t rows = (int64_t)batch * Nq;
  softmax_rows_kernel<<<(unsigned)ceil_div<int64_t>(rows, 8), 256, 0, as_stream(stream)>>>(attn, rows, Nk);
  DVD_LAUNCH_CHECK();
  for (int b0 = 0; b0 < batch; b0 += 65535) {
    const int nb = batch - b0 < 65535 ? batch - b0 : 65535;
    // out[c][i] = sum_j v[c][j] attn[i][j]
    DVD_TRY(dvd_bgemm(0, 1, dv, Nq, Nk, 1.f, v + b0 * v_bs, Nk, v_bs, attn + b0 * a_bs, Nk, a_bs, 0.f, out + b0 * o_bs,
                      Nq, o_bs, nb, nullptr, stream));
  }
  return 0;
}

extern "C" int dvd_attn_bwd(const float* q, int64_t q_bs, const float* k, int64_t k_bs, const float* v, int64_t v_bs,
                            const float* attn, float* dattn, const float* dout, int64_t do_bs, float* dq_,
                            int64_t dq_bs, float* dk_, int64_t dk_bs, float* dv_, int64_t dv_bs, int batch, int dq,
                            int dv, int Nq, int Nk, int q_token_major, void* stream) {
  dvd::ProfScope _ps(3, "attn_bwd", dvd::as_stream(stream));
  DVD_CHECK_ARG(q && k && v && attn && dattn && dout && dq_ && dk_ && dv_);
  DVD_CHECK_ARG(batch > 0 && dq > 0 && dv > 0 && Nq > 0 && Nk > 0);
  const int64_t a_bs = (int64_t)Nq * Nk;
  for (int b0 = 0; b0 < batch; b0 += 65535) {
    const int nb = batch - b0 < 65535 ? batch - b0 : 65535;
    // dV[c][j] = sum_i dout[c][i] attn[i][j]
    DVD_TRY(dvd_bgemm(0, 0, dv, Nk, Nq, 1.f, dout + b0 * do_bs, Nq, do_bs, attn + b0 * a_bs, Nk, a_bs, 0.f,
                      dv_ + b0 * dv_bs, Nk, dv_bs, nb, nullptr, stream));
    // dA[i][j] = sum_c dout[c][i] v[c][j]
    DVD_TRY(dvd_bgemm(1, 0, Nq, Nk, dv, 1.f, dout + b0 * do_bs, Nq, do_bs, v + b0 * v_bs, Nk, v_bs, 0.f,
                      dattn + b0 * a_bs, Nk, a_bs, nb, nullptr, stream));
  }
  const int64_t rows = (int64_t)batch * Nq;
  softmax_bwd_rows_kernel<<<(unsigned)ceil_div<int64_t>(rows, 8), 256, 0, as_stream(stream)>>>(attn, dattn, rows, Nk);
  DVD_LAUNCH_CHECK();
  for (int b0 = 0; b0 < batch; b0 += 65535) {
    const int nb = batch - b0 < 65535 ? batch - b0 : 65535;
    const float* dS = dattn + b0 * a_bs;
    if (q_token_major) {
      // dq[i][c] = sum_j dS[i][j] k[c][j] ;  dk[c][j] = sum_i q[i][c] dS[i][j]
      DVD_TRY(dvd_bgemm(0, 1, Nq, dq, Nk, 1.f, dS, Nk, a_bs, k + b0 * k_bs, Nk, k_bs, 0.f, dq_ + b0 * dq_bs, dq, dq_bs,
                        nb, nullptr, stream));
      DVD_TRY(dvd_bgemm(1, 0, dq, Nk, Nq, 1.f, q + b0 * q_bs, dq, q_bs, dS, Nk, a_bs, 0.f, dk_ + b0 * dk_bs, Nk, dk_bs,
                        nb, nullptr, stream));
    } else {
      // dq[c][i] = sum_j k[c][j] dS[i][j] ;  dk[c][j] = sum_i q[c][i] dS[i][j]
      DVD_TRY(dvd_bgemm(0, 1, dq, Nq, Nk, 1.f, k + b0 * k_bs, Nk, k_bs, dS, Nk, a_bs, 0.f, dq_ + b0 * dq_bs, Nq, dq_bs,
                        nb, nullptr, stream));
      DVD_TRY(dvd_bgemm(0, 0, dq, Nk, Nq, 1.f, q + b0 * q_bs, Nq, q_bs, dS, Nk, a_bs, 0.f, dk_ + b0 * dk_bs, Nk, dk_bs,
                        nb, nullptr, stream));
    }
  }
  return 0;
}
