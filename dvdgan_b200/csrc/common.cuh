// Shared helpers for the dvdgan_b200 CUDA library (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/dvdgan_b200.h"

namespace dvd {

extern thread_local char g_last_error[512];
extern std::atomic<long long> g_launches;   // kernels launched by this library (bench.py reports it)

// Process-wide tuning switches (dvd_set_option / dvd_get_option in the C ABI).  The defaults are the validated
// configuration; the switches exist for A/B measurements and for the tests that cross-check one engine against another.
enum Option {
  OPT_SIMT_ONLY = 0,     // "simt_only":   fp32 FFMA conv engine for everything (default 0)
  OPT_PAIR,              // "pair":        cta_group::2 CTA-pair tiles (1)
  OPT_PERSIST,           // "persist":     persistent CTA pairs when there is more than one wave of tiles (1)
  OPT_ONEACC,            // "oneacc":      persistent tiles accumulate all three MMA products into ONE fp32 TMEM
                         //                accumulator and double-buffer it (epilogue hidden under the next tile);
                         //                implies bf16 planes in the forward (0)
  OPT_OCC2,              // "occ2":        two CTAs per SM for short reductions on narrow tiles (1)
  OPT_EPI_PREFETCH,      // "epi_prefetch": L2 prefetch of the epilogue operands late in the main loop (1)
  OPT_GRU_FUSED,         // "gru_fused":   ConvGRU gate math in the h-half GEMM epilogues (1)
  OPT_GRU_SHARE_PLANES,  // "gru_share_planes": BPTT gate-gradient planes shared by x-dgrad and the weight gradients (1)
  OPT_GRU_BWD_PLANES,    // "gru_bwd_planes":  the BPTT elementwise kernels write those planes themselves (1)
  OPT_GRU_BWD_FUSED,     // "gru_bwd_fused":   BPTT gate-gradient math in the per-step dgrad GEMM epilogues (0: measured
                         //                    neutral -- 1000 fewer launches per step, but the single accumulator set
                         //                    cannot hide the longer epilogues, so the time only moves into the GEMMs)
  OPT_FLASH_ATTN,        // "flash_attn":  tcgen05 attention that never materialises the N x N map (1)
  OPT_FWD_BF16,          // "fwd_bf16":    bf16 operand planes in the forward too (fp32 range, 16-bit operand precision) (0)
  OPT_GRU_STREAMS,       // "gru_streams": the ConvGRU time loops run as this many independent chains over batch slices,
                         //                chain 0 on the caller's stream, the others on library-owned helper streams
                         //                forked from / joined to it (1 = one chain, no helper streams) (2)
  OPT_COUNT
};
int get_option(int opt);

// Optional CUDA-event profiling of the dense engines (category 0: conv fwd/dgrad, 1: wgrad).
void prof_begin(int category, double flops, cudaStream_t st);
void prof_end(int category, cudaStream_t st);
void prof_tag(const char* fmt, int a = 0, int b = 0, int c = 0, int d = 0, int e = 0, int f = 0);   // label of the next record

// RAII bracket for a whole C-ABI call (category 3 = memory-bound helpers): tag = the entry point's name
struct ProfScope {
  int cat;
  cudaStream_t st;
  ProfScope(int c, const char* tag, cudaStream_t s) : cat(c), st(s) { prof_tag(tag); prof_begin(c, 0.0, s); }
  ~ProfScope() { prof_end(cat, st); }
};

inline int fail(const char* fmt, const char* a = "", const char* file = "", int line = 0) {
  snprintf(g_last_error, sizeof(g_last_error), fmt, a, file, line);
  return 1;
}

#define DVD_CHECK_ARG(cond)                                                          \
  do {                                                                               \
    if (!(cond)) return ::dvd::fail("invalid argument: %s (%s:%d)", #cond, __FILE__, __LINE__); \
  } while (0)

#define DVD_CUDA(expr)                                                               \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess)                                                           \
      return ::dvd::fail("CUDA error: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define DVD_LAUNCH_CHECK()                                                           \
  do {                                                                               \
    ::dvd::g_launches.fetch_add(1, std::memory_order_relaxed);                       \
    cudaError_t _e = cudaGetLastError();                                             \
    if (_e != cudaSuccess)                                                           \
      return ::dvd::fail("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define DVD_TRY(expr)            \
  do {                           \
    int _r = (expr);             \
    if (_r != 0) return _r;      \
  } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

inline int num_sms() {
  static thread_local int cached_dev = -1;
  static thread_local int cached = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) cached = n;
    cached_dev = dev;
  }
  return cached;
}

// One-time per-device setup (cudaFuncSetAttribute, mem-pool thresholds): returns whether the current device's bit was
// already set, and sets it.  Devices >= 64 are configured on every call (harmless).
inline bool device_bit_test_and_set(std::atomic<uint64_t>& bits) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return false;
  const uint64_t bit = 1ull << dev;
  return (bits.fetch_or(bit, std::memory_order_acq_rel) & bit) != 0;
}

template <typename T>
__host__ __device__ inline T ceil_div(T a, T b) { return (a + b - 1) / b; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum (blockDim.x multiple of 32, <= 1024). Result valid in every thread.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* smem /* >= 32 */) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) smem[wid] = v;
  __syncthreads();
  T r = (lane < nw) ? smem[lane] : T(0);
  r = warp_sum(r);
  return r;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// Elementwise launch geometry: a multiple of the SM count, grid-stride loops inside.
inline int ew_blocks(int64_t n, int per_thread = 4, int threads = 256) {
  int64_t want = ceil_div<int64_t>(n, (int64_t)per_thread * threads);
  int64_t cap = (int64_t)num_sms() * 8;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

}  // namespace dvd
