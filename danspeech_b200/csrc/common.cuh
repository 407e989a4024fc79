// Shared helpers for the danspeech_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>
#include "../../include/danspeech_b200.h"

namespace dsb {

extern thread_local char g_err[512];
extern std::atomic<uint64_t> g_launches;

int set_error(int code, const char* fmt, ...);

inline void count_launch(int n = 1) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

#define DSB_CUDA(expr)                                                                         \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      return dsb::set_error(DSB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                            __FILE__, __LINE__);                                               \
  } while (0)

#define DSB_CHECK_LAUNCH()                                                                     \
  do {                                                                                         \
    dsb::count_launch();                                                                       \
    cudaError_t _e = cudaGetLastError();                                                       \
    if (_e != cudaSuccess)                                                                     \
      return dsb::set_error(DSB_ERR_CUDA, "kernel launch failed: %s (%s:%d)",                  \
                            cudaGetErrorString(_e), __FILE__, __LINE__);                       \
  } while (0)

#define DSB_REQUIRE(cond, ...)                                        \
  do {                                                                \
    if (!(cond)) return dsb::set_error(DSB_ERR_INVALID, __VA_ARGS__); \
  } while (0)

// Stage timers (CUDA events on the launching stream), enabled by dsb_profile_enable().
enum Stage { ST_SPECT = 0, ST_CONV, ST_PROJ, ST_RNN, ST_TAIL, ST_GREEDY, ST_BEAM, ST_COMBINE, ST_COUNT };
void prof_begin(int stage, cudaStream_t st);
void prof_end(int stage, cudaStream_t st);
struct ProfScope {
  int stage;
  cudaStream_t st;
  ProfScope(int s, cudaStream_t t) : stage(s), st(t) { prof_begin(stage, st); }
  ~ProfScope() { prof_end(stage, st); }
};

// Tuning knobs (dsb_tune_set / dsb_tune_get; defaults are the production settings).
struct Tune {
  std::atomic<int> rnn_in_flight{3};
  std::atomic<int> rnn_max_slots{0};
  std::atomic<int> rnn_ksplit{0};
  std::atomic<int> rnn_ring_gsz{0};
  std::atomic<int> rnn_producers{1};
  std::atomic<int> rnn_pair{1};             // CTA-pair recurrence (rnn_pair.cu) for batches of two or more groups
  std::atomic<int> rnn_batch_minor{1};      // batch-minor pre-activations / outputs around the CTA-pair recurrence
  std::atomic<int> rnn_pair_min_rows{1};    // smallest batch the CTA-pair kernel takes (a single group leaves half of every pair idle and is still ~20 % faster than the one-CTA kernel: shorter publish chain)
  std::atomic<int> rnn_pair_in_flight{2};   // pair items (two groups of 64 sequences each) in flight per CTA pair
};
extern Tune g_tune;

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

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

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

}  // namespace dsb
