/*
 * dvdgan_b200 -- C ABI of the B200 (sm_100a) DVD-GAN training hot path.
 *
 * The reference (Harrypotterrrr/DVD-GAN) has no native/FFI layer: its hot path is the Python
 * nn.Module API.  This header is the boundary a binding sits on: plain
 * pointers and sizes, no torch types.  Every entry point cites the reference code it replaces
 * (file:line into the reference tree).
 *
 * Conventions
 *  - all tensors are fp32, device memory, caller-owned; class ids / frame indices are int64.
 *  - every call only enqueues work on `stream` (a cudaStream_t passed as void*) of the CURRENT device; no host sync
 *    (except dvd_index_errors and the dvd_prof_* measurement aids).  Persistent buffers are the caller's; the
 *    tensor-core engine additionally takes stream-ordered scratch for its operand planes from the device's default
 *    cudaMallocAsync pool (dvd_scratch_bytes).  Re-entrant per (device, stream): one-time kernel attributes and pool
 *    settings are tracked per device; the only process-wide state is the option table (dvd_set_option), the launch
 *    counter and the profiler.
 *  - return value: 0 = ok, non-zero = error; dvd_last_error() gives a thread-local message.
 *  - "packed" conv weights: [tap][Cin][Cout] (Cout contiguous), produced by dvd_weight_pack.
 */
#ifndef DVDGAN_B200_H
#define DVDGAN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* dvd_last_error(void);
int dvd_abi_version(void);
/* Number of CUDA kernels this library has launched in this process (all threads). */
long long dvd_launch_count(void);
/* Measurement aid: while enabled, every launch of the dense engines (category 0: dvd_conv_fwd incl. dgrads,
 * 1: dvd_conv_wgrad) is bracketed by CUDA events on its stream.  dvd_prof_read waits for the recorded events,
 * returns the summed device time (ms), algorithmic FLOPs and launch count, and resets the category. */
int dvd_prof_enable(int on);   /* 0: off; 1: categories 0 and 1; else a bit mask of categories (0xF = all four) */
int dvd_prof_read(int category, double* ms, double* flops, long long* launches);
/* per-shape table of the launches recorded so far (not cleared): "category \t tag \t launches \t ms \t flops-or-bytes";
 * category 2 = operand-plane preparation of the tcgen05 engine (value column = bytes moved) */
int dvd_prof_dump(const char* path);

/* Process-wide tuning switches (atomic; read at every call, so they can be flipped between calls).  Defaults are the
 * validated configuration; nothing is read from the environment inside the library.
 *   "simt_only" 0   fp32 FFMA conv engine for every convolution (the tests' cross-check of the tensor-core engine)
 *   "pair" 1, "persist" 1, "occ2" 1, "epi_prefetch" 1     tile / scheduling variants of the tcgen05 engine
 *   "oneacc" 0      persistent tiles use ONE fp32 accumulator for all three products and double-buffer it (implies
 *                   "fwd_bf16": measured +2 % on the step, Generator output 1.2e-3 from the fp32 reference)
 *   "fwd_bf16" 0    bf16 operand planes in the forward too: fp32's exponent range at 16-bit operand precision; the
 *                   default fp16 planes (22 bits) clamp |x| > 65504 and count it (dvd_saturation_count)
 *   "gru_fused" 1, "gru_share_planes" 1, "gru_bwd_planes" 1, "gru_bwd_fused" 0   ConvGRU fusion levels (the last one
 *                   moves the BPTT gate-gradient math into the dgrad GEMM epilogues: measured neutral, off)
 *   "flash_attn" 1  attention on the tensor cores without the N x N map (0: materialised SIMT path)
 * Unknown names are an error. */
int dvd_set_option(const char* name, int value);
int dvd_get_option(const char* name, int* value);
/* Forward operands of the tensor-core engine are fp16 (hi, lo) planes; |x| > 65504 is clamped and counted.  Reads -- and
 * with reset != 0 clears -- the count of clamped 8-element groups on the current device (0 in a healthy run; the fp32
 * reference would carry such values on).  Synchronises `stream`.  Remedy: dvd_set_option("fwd_bf16", 1). */
int dvd_saturation_count(unsigned int* count, int reset, void* stream);
/* High-water mark / currently reserved bytes of the current device's default stream-ordered memory pool: the bf16
 * operand planes of the tensor-core engine are cudaMallocAsync'ed there (freed in stream order after each call). */
int dvd_scratch_bytes(long long* high_water, long long* reserved);

/* ------------------------------------------------------------------------------------------
 * Dense engines
 * ---------------------------------------------------------------------------------------- */

/* Stride-1 "same" convolution, 1-D..3-D, odd kernel, NC(D)HW, implicit GEMM (no im2col buffer).
 * Image index n = n1*N2 + n2; offsets n1*s1 + n2*s2 (two-level batch so (B,T,...) slices and
 * time-shifted pairs need no copies).  Replaces every nn.Conv2d/Conv3d call site on the path:
 * ConvGRU.py:47-52, GResBlock.py:57,64,73, Generator.py:114, Discriminators.py:89-91,190-206,
 * 222-226,314-324,376-382 -- and their autograd dgrads (run with weights packed transposed). */
typedef struct {
  int N1, N2;                 /* images = N1*N2 */
  int Cin, Cout;
  int D, H, W;                /* output extent (D = 1 for 2-D) */
  int kD, kH, kW;             /* odd; padding = k/2 */
  int64_t x_s1, x_s2, x_cs;   /* input strides: batch level 1, level 2, channel (elements) */
  int64_t y_s1, y_s2, y_cs;   /* output strides */
  int in_relu;                /* apply ReLU to the input while loading (F.relu -> conv) */
  int in_up;                  /* input is nearest-upsampled x2 in H,W while loading (F.interpolate) */
  int accumulate;             /* y += conv(x) instead of y = conv(x) */
  int out_act;                /* 0 none, 1 relu, 2 tanh (applied after bias / residual) */
  int res_up;                 /* residual is read at (h>>res_up, w>>res_up) */
  int64_t r_s1, r_s2, r_cs;   /* residual strides (if res != NULL) */
  int x_kind;                 /* tensor-core path: 1 = x holds forward activations / the conv is a forward conv: fp16
                                 operand planes (22 bits, |x| <= 65504, see dvd_saturation_count); 0 = x may hold
                                 gradients: bf16 planes (fp32's exponent range) */
} dvd_conv_desc;

int dvd_conv_fwd(const dvd_conv_desc* d, const float* x, const float* w_packed, const float* bias,
                 const float* res, float* y, void* stream);

/* Weight gradient: dwp[tap][Cin][Cout] (+)= sum_pixels x[pixel+tap, ci] * dy[pixel, co].
 * x_* strides describe x, y_* strides describe dy; in_relu / in_up as in the forward.
 * `dwp` must be zero-filled by the caller when accumulate == 0 (split-K uses atomics). */
int dvd_conv_wgrad(const dvd_conv_desc* d, const float* x, const float* dy, float* dwp, void* stream);

/* Pack rows [co0,co0+Cout) x input channels [ci0,ci0+Cin) of a reference-layout weight
 * w[Co_total][Ci_total][taps] into the GEMM operand layout, optionally scaled by 1/(*sigma)
 * (spectral norm, Normalization.py:31).  The destination is [tap][dst_rows][dst_ld]:
 *   transpose == 0 (forward operand):  dst[tap][row_off + ci][col_off + co]
 *   transpose != 0 (dgrad operand):    dst[taps-1-tap][row_off + co][col_off + ci]
 * row/col offsets let several weights be concatenated along either channel axis. */
int dvd_weight_pack(const float* w, int Ci_total, int taps, int co0, int Cout, int ci0, int Cin,
                    const float* sigma, int transpose, float* dst, int dst_rows, int dst_row_off, int dst_ld,
                    int dst_col_off, void* stream);
/* Inverse of dvd_weight_pack(transpose=0) for gradients: w_grad[co0+co][ci0+ci][tap] (+)= src[tap][ci][src_off+co]. */
int dvd_weight_unpack(const float* src, int src_ld, int src_off, int Ci_total, int taps, int co0, int Cout,
                      int ci0, int Cin, int accumulate, float* w_grad, void* stream);

/* Strided-batched SGEMM, row-major: C[b] = alpha * op(A[b]) * op(B[b]) + beta * C[b] (+ bias[n] per column).
 * Replaces F.linear (Generator.py:75, Normalization.py:80) and torch.bmm (Discriminators.py:110,114;
 * Attention.py:94,101,170,176) and their gradients. */
int dvd_bgemm(int transA, int transB, int M, int N, int K, float alpha, const float* A, int lda, int64_t strideA,
              const float* B, int ldb, int64_t strideB, float beta, float* C, int ldc, int64_t strideC, int batch,
              const float* bias, void* stream);

/* ------------------------------------------------------------------------------------------
 * ConvGRU (ConvGRU.py:29-54,104-133 driven by Generator.py:87-97): one layer over all T frames.
 * Buffers are b-major: x (B,T,Cx,H,W) [x_bs/x_ts strides allow a frame broadcast, Q13],
 * gates (B,T,3Ch,H,W) = (update, reset, out), h (B,T,Ch,H,W), rh (B,T,Ch,H,W) = reset * h_prev.
 * wu/wr/wo are the reference-layout gate weights [Ch][Cx+Ch][k][k]; h0 may be NULL (zero state).
 * ---------------------------------------------------------------------------------------- */
size_t dvd_convgru_layer_workspace_bytes(int B, int T, int Cx, int Ch, int H, int W, int k);
int dvd_convgru_layer_fwd(const float* x, int64_t x_bs, int64_t x_ts, const float* h0,
                          const float* wu, const float* wr, const float* wo,
                          const float* bu, const float* br, const float* bo,
                          float* gates, float* h, float* rh,
                          int B, int T, int Cx, int Ch, int H, int W, int k,
                          void* workspace, size_t ws_bytes, void* stream);
/* Backward through time.  dh (B,T,Ch,H,W) = grad wrt every h_t (read only); gates are overwritten
 * with the pre-activation grads.  Outputs: dx (B,T,Cx,H,W), dh0 (may be NULL), dwu/dwr/dwo/dbu/dbr/dbo
 * (overwritten, reference layout). */
int dvd_convgru_layer_bwd(const float* x, int64_t x_bs, int64_t x_ts, const float* h0,
                          const float* wu, const float* wr, const float* wo,
                          float* gates, const float* h, const float* rh, const float* dh,
                          float* dx, float* dh0, float* dwu, float* dwr, float* dwo,
                          float* dbu, float* dbr, float* dbo,
                          int B, int T, int Cx, int Ch, int H, int W, int k,
                          void* workspace, size_t ws_bytes, void* stream);
/* Both calls split the batch into `gru_streams` (option, default 2) independent chains over B / n clips: chain 0 runs
 * on `stream`, the others on library-owned non-blocking helper streams (one set per host thread and device) that are
 * forked from and joined to `stream` with events inside the call, so the call stays stream-ordered for its caller.
 *
 * The same sweeps over a RANGE of frames [t0, t1), for callers that overlap the layers of a ConvGRU stack
 * (Generator.py:87-97 runs layer l+1 of frame t right after layer l: nothing forces a layer to finish its clip before
 * the next one starts).  All tensor arguments are the whole-clip buffers of the calls above; a call touches frames
 * [t0, t1) only.  Forward: calls that share `workspace` visit the frames in ascending order, the first one (t0 = 0)
 * prepares the weights there.  Backward: calls that share `workspace` (dvd_convgru_layer_range_workspace_bytes: it
 * also holds the gate-gradient operand planes of the whole clip) visit the frames in descending order starting at
 * t1 = T; every call needs dh of its frames and leaves dx of its frames; the last one (t0 = 0) also writes dh0 and the
 * weight / bias gradients.  No helper streams here: the caller supplies the concurrency (one stream per layer). */
size_t dvd_convgru_layer_range_workspace_bytes(int B, int T, int Cx, int Ch, int H, int W, int k);
int dvd_convgru_layer_fwd_range(const float* x, int64_t x_bs, int64_t x_ts, const float* h0,
                                const float* wu, const float* wr, const float* wo,
                                const float* bu, const float* br, const float* bo,
                                float* gates, float* h, float* rh,
                                int B, int T, int Cx, int Ch, int H, int W, int k, int t0, int t1,
                                void* workspace, size_t ws_bytes, void* stream);
int dvd_convgru_layer_bwd_range(const float* x, int64_t x_bs, int64_t x_ts, const float* h0,
                                const float* wu, const float* wr, const float* wo,
                                float* gates, const float* h, const float* rh, const float* dh,
                                float* dx, float* dh0, float* dwu, float* dwr, float* dwo,
                                float* dbu, float* dbr, float* dbo,
                                int B, int T, int Cx, int Ch, int H, int W, int k, int t0, int t1,
                                void* workspace, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Normalisation
 * ---------------------------------------------------------------------------------------- */

/* Spectral norm, one power iteration (Normalization.py:19-31): updates u[rows], v[cols] in place,
 * writes sigma[0].  scratch: rows + cols + 4 floats. */
int dvd_specnorm_fwd(const float* w_bar, int rows, int cols, float* u, float* v, float* sigma, float* scratch,
                     void* stream);
/* dW_bar = (G - <G, W_bar/sigma> u v^T) / sigma with the CURRENT u, v (SURVEY Q3 / DESIGN Q17).
 * scratch: 2 doubles. */
int dvd_specnorm_bwd(const float* g, const float* w_bar, const float* u, const float* v, const float* sigma,
                     int rows, int cols, float* dw_bar, int accumulate, void* scratch, void* stream);

/* BatchNorm2d(affine=False) statistics (Normalization.py:79): mean/rstd over (N,HW) per channel;
 * training: batch stats, running stats updated (momentum, unbiased var), eval: running stats.
 * scratch: 2*C doubles. */
int dvd_bn_stats(const float* x, int N, int C, int HW, int training, float momentum, float eps,
                 float* running_mean, float* running_var, int64_t* num_batches_tracked, float* mean, float* rstd,
                 void* scratch, void* stream);
/* The same in phases, for statistics over the replicas of a data-parallel job (the reference's TODO at
 * Generator.py:57-58): phase 0 = everything (dvd_bn_stats); phase 1 leaves this rank's per-channel (sum, sum of squares)
 * in `scratch` (2*C doubles) for the caller to sum over the ranks; phase 2 finalises mean / rstd / running statistics
 * from the summed scratch, counting count_scale * N * HW elements (count_scale = number of equal-sized ranks). */
int dvd_bn_stats_ex(const float* x, int N, int C, int HW, int training, float momentum, float eps,
                    float* running_mean, float* running_var, int64_t* num_batches_tracked, float* mean, float* rstd,
                    void* scratch, int phase, int count_scale, void* stream);
/* y = relu?(gb[r][c] * xhat + gb[r][C+c]) nearest-upsampled by 2^up (Normalization.py:82-86 +
 * GResBlock.py:52-55): the conditional affine, activation and F.interpolate in one pass.
 * gb has gb_rows rows and row r = n % gb_rows (gb_rows == N: one row per image; gb_rows == B reproduces
 * Generator.py:109-110's condition.repeat(T,1) against b-major frames, SURVEY Q1, without materialising it). */
int dvd_cbn_apply(const float* x, const float* gb, int gb_rows, const float* mean, const float* rstd, int N, int C,
                  int H, int W, int relu, int up, float* y, void* stream);
/* Backward of dvd_cbn_apply + batch statistics.  dgb[gb_rows][2C] receives (dgamma | dbeta), summed over
 * the images that share a row.  scratch: 2*C floats. */
int dvd_cbn_bwd(const float* x, const float* gb, int gb_rows, const float* mean, const float* rstd, const float* dy,
                int N, int C, int H, int W, int relu, int up, int training, float* dx, float* dgb, float* scratch,
                void* stream);
/* In phases (cross-replica statistics): phase 1 computes dgb and leaves this rank's per-channel (mean(dxhat),
 * mean(dxhat * xhat)) in `scratch` (2*C floats) for the caller to AVERAGE over the equal-sized ranks; phase 2 computes dx
 * from the averaged scratch; phase 0 = everything (dvd_cbn_bwd). */
int dvd_cbn_bwd_ex(const float* x, const float* gb, int gb_rows, const float* mean, const float* rstd, const float* dy,
                   int N, int C, int H, int W, int relu, int up, int training, float* dx, float* dgb, float* scratch,
                   int phase, void* stream);

/* ------------------------------------------------------------------------------------------
 * Attention core (Discriminators.py:108-114, Attention.py:92-101,165-176):
 * out[b][:, i] = sum_j softmax_j(q_i . k_j) v[b][:, j]; channel-major q [dq][Nq], k [dq][Nk],
 * v [dv][Nk], out [dv][Nq].  q_token_major: q is [Nq][dq] (SeparableAttnCell's raw view).
 * attn (batch, Nq, Nk) is written by fwd and consumed by bwd; dattn is a same-sized scratch.
 * ---------------------------------------------------------------------------------------- */
int dvd_attn_fwd(const float* q, int64_t q_bs, const float* k, int64_t k_bs, const float* v, int64_t v_bs,
                 float* attn, float* out, int64_t o_bs, int batch, int dq, int dv, int Nq, int Nk,
                 int q_token_major, void* stream);
int dvd_attn_bwd(const float* q, int64_t q_bs, const float* k, int64_t k_bs, const float* v, int64_t v_bs,
                 const float* attn, float* dattn, const float* dout, int64_t do_bs, float* dq_, int64_t dq_bs, float* dk_,
                 int64_t dk_bs, float* dv_, int64_t dv_bs, int batch, int dq, int dv, int Nq, int Nk,
                 int q_token_major, void* stream);

/* The same contraction on the tensor cores (tcgen05, bf16-split operands, fp32 TMEM accumulators) WITHOUT the N x N map:
 * a two-pass softmax (log-sum-exp per query first, then exp(s - lse) straight into the P.V product), the backward
 * recomputes P from lse.  Covers channel-major q/k/v with dq <= 64, dv <= 256 (backward: 128), token counts that are
 * multiples of 8 (dvd_attn_flash_supported); the caller falls back to dvd_attn_fwd / dvd_attn_bwd otherwise.
 * `workspace`: dvd_attn_flash_workspace_bytes bytes of device memory (operand planes), free to reuse after the call. */
int dvd_attn_flash_supported(int batch, int dq, int dv, int Nq, int Nk, int backward);
size_t dvd_attn_flash_workspace_bytes(int batch, int dq, int dv, int Nq, int Nk, int backward);
int dvd_attn_flash_fwd(const float* q, int64_t q_bs, const float* k, int64_t k_bs, const float* v, int64_t v_bs,
                       float* out, int64_t o_bs, float* lse, int batch, int dq, int dv, int Nq, int Nk,
                       void* workspace, size_t ws_bytes, void* stream);
int dvd_attn_flash_bwd(const float* q, int64_t q_bs, const float* k, int64_t k_bs, const float* v, int64_t v_bs,
                       const float* out, int64_t o_bs, const float* dout, int64_t do_bs, const float* lse,
                       float* dq_, int64_t dq_bs, float* dk_, int64_t dk_bs, float* dv_, int64_t dv_bs, int batch,
                       int dq, int dv, int Nq, int Nk, void* workspace, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Pooling, resampling, helpers
 * ---------------------------------------------------------------------------------------- */
/* Average pool, window = stride = (pd,ph,pw) in {1,2} (F.avg_pool2d/3d, Discriminators.py:197,206,352,361).
 * fwd: y (+)= scale * pool(x) ; bwd: dx (+)= broadcast(dy)/window.  NC = product of the leading dims. */
int dvd_avgpool_fwd(const float* x, int64_t NC, int D, int H, int W, int pd, int ph, int pw, float scale,
                    int accumulate, float* y, void* stream);
int dvd_avgpool_bwd(const float* dy, int64_t NC, int D, int H, int W, int pd, int ph, int pw, int accumulate, float* dx,
                    void* stream);
/* Max pool, window = stride = (pd,ph,pw) (nn.MaxPool3d, Attention.py:48,138). */
int dvd_maxpool_fwd(const float* x, int64_t NC, int D, int H, int W, int pd, int ph, int pw, float* y, void* stream);
int dvd_maxpool_bwd(const float* x, const float* dy, int64_t NC, int D, int H, int W, int pd, int ph, int pw, float* dx,
                    void* stream);
/* phi (utils.py:77-83): (B,T,C,H,W) -> 2x2 avg pool -> (B,C,T,H/2,W/2). */
int dvd_phi_fwd(const float* x, int B, int T, int C, int H, int W, float* y, void* stream);
int dvd_phi_bwd(const float* dy, int B, int T, int C, int H, int W, int accumulate, float* dx, void* stream);
/* sample_k_frames gather (utils.py:60-63): y[b][j] = x[b][idx[j]]; bwd scatters (dx zero-filled by callee
 * unless accumulate). */
int dvd_gather_frames_fwd(const float* x, const int64_t* idx, int B, int T, int k, int64_t frame_elems, float* y,
                          void* stream);
int dvd_gather_frames_bwd(const float* dy, const int64_t* idx, int B, int T, int k, int64_t frame_elems,
                          int accumulate, float* dx, void* stream);
/* General 5-D permute: y = x.permute(perm).contiguous(); dims = sizes of x (Attention.py:81-84,103-108). */
int dvd_permute5(const float* x, const int* dims, const int* perm, float* y, void* stream);
/* permute (B,C,T,HW) <-> (B,T,C,HW) (Discriminators.py:413-415). */
int dvd_permute_bctp(const float* x, int B, int C, int T, int64_t P, float* y, void* stream);
/* Elementwise helpers: y = act(x) [0 copy,1 relu,2 tanh]; bwd from the OUTPUT for tanh, from x for relu. */
int dvd_act_fwd(const float* x, int64_t n, int act, float* y, void* stream);
int dvd_act_bwd(const float* ref, const float* dy, int64_t n, int act, float* dx, void* stream);
/* y = gamma[0]*o + x (Discriminators.py:118); bwd: dgamma = <dy,o> (scratch: 1 double), do = gamma*dy. */
int dvd_scale_residual_fwd(const float* o, const float* x, const float* gamma, int64_t n, float* y, void* stream);
int dvd_scale_residual_bwd(const float* o, const float* dy, const float* gamma, int64_t n, float* do_,
                           float* dgamma, void* scratch, void* stream);
/* per-channel sum over (N, P): out[c] (+)= sum x[n][c][p]  (bias gradients). scratch: C doubles. */
int dvd_channel_sum(const float* x, int N, int C, int64_t P, int64_t n_stride, int accumulate, float* out,
                    void* scratch, void* stream);
/* axpby: y = a*x + b*y */
int dvd_axpby(const float* x, float a, float b, int64_t n, float* y, void* stream);
/* Multi-tensor gather of the flat-arena optimizer (trainer.py:136-141 keeps one gradient tensor per parameter):
 * flat[offs[i] .. offs[i]+counts[i]) = srcs[i][0 .. counts[i])  (zeros when srcs[i] == NULL), i < n, in ceil(n/96)
 * launches.  srcs / offs / counts are HOST arrays of device pointers / element offsets / element counts. */
int dvd_gather_flat(const float* const* srcs, const int64_t* offs, const int64_t* counts, int n, float* flat, void* stream);
/* Embedding rows (Generator.py:70): y[i] = w[idx[i]]; bwd: dw[idx[i]] += dy[i] (dw pre-zeroed by caller). */
int dvd_embedding_fwd(const float* w, const int64_t* idx, int n, int dim, int rows, float* y, void* stream);
int dvd_embedding_bwd(const float* dy, const int64_t* idx, int n, int dim, int rows, float* dw, void* stream);
/* nn.Embedding raises on an index outside [0, rows); the kernels above and the head below count such ids (and clamp
 * them so no access leaves the table).  Reads -- and with reset != 0 clears -- the current device's count.  This is the
 * one entry point that synchronises (it waits for `stream`). */
int dvd_index_errors(unsigned int* count, int reset, void* stream);

/* Discriminator head (Discriminators.py:264-291, 421-447): feat[n][c] = sum_hw relu(x);
 * out[n] = feat . w_lin/sigma_l + b + feat . emb[class[n / T]]/sigma_e. */
int dvd_dhead_fwd(const float* x, int N, int C, int HW, int T, int n_class, const float* w_lin,
                  const float* sigma_l, const float* b_lin, const float* emb, const float* sigma_e,
                  const int64_t* class_id, float* feat, float* out, void* stream);
/* Grads wrt x, W_lin_sn (dwl[C]), bias (db[1]) and Emb_sn (demb[n_class][C]); outputs are overwritten. */
int dvd_dhead_bwd(const float* x, const float* feat, const float* dout, int N, int C, int HW, int T, int n_class,
                  const float* w_lin, const float* sigma_l, const float* emb, const float* sigma_e,
                  const int64_t* class_id, float* dx, float* dwl, float* db, float* demb, void* stream);

/* Input pipeline on the GPU (Dataloader/datasets/ucf101.py:177-199 with the transforms of main.py:33-56):
 * frames uint8 [B][T][Hs][Ws][3] (decoded RGB) -> per-clip crop `box` = (x0, y0, w, h), already rounded the way
 * PIL.Image.crop rounds -> bilinear resize to OH x OW exactly as Pillow's ImagingResample does for 8-bit images (two
 * passes, 22-bit fixed-point coefficients, uint8 intermediate) -> horizontal flip where flip[b] != 0 ->
 * (pixel / norm_value - mean[c]) / std[c] in fp32 -> out (B, 3, T, OH, OW).
 * xb / yb: [B][OW|OH][2] = (first source index inside the crop, tap count); xk / yk: [B][OW|OH][xks|yks] fixed-point
 * coefficients; both come from the host restatement of Pillow's precompute_coeffs (dvdgan_b200/data.py).  mean3 / std3
 * are HOST pointers to three floats. */
int dvd_clip_transform(const uint8_t* frames, int B, int T, int Hs, int Ws, const int* box, const int* flip,
                       const int* xb, const int* xk, int xks, const int* yb, const int* yk, int yks, int OH, int OW,
                       float norm_value, const float* mean3, const float* std3, float* out, void* stream);

/* Losses (trainer.py:114-121): sign = -1 for real_flag; hinge: mean(relu(1 + sign*x)); wgan: mean(sign*x).
 * loss[0] (+)= value;  bwd: dx = gout[0] * dloss/dx. */
int dvd_gan_loss_fwd(const float* x, int n, float sign, int hinge, int accumulate, float* loss, void* stream);
int dvd_gan_loss_bwd(const float* x, const float* gout, int n, float sign, int hinge, float* dx, void* stream);

/* Fused Adam over a flat fp32 arena (trainer.py:136-141 torch.optim.Adam, eps 1e-8, no weight decay),
 * grads optionally pre-scaled (1/world_size after a sum all-reduce). */
int dvd_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                  float eps, int step, float grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DVDGAN_B200_H */
