// Pieces shared by the persistent tensor-core recurrence kernels (rnn_tc.cu: one CTA per W_hh slice; rnn_pair.cu:
// CTA pairs, cta_group::2): shared-memory plan, kernel parameters, bounded waits, gate non-linearities.
#pragma once
#include "tc_common.cuh"
#include "model_types.cuh"

namespace dsb {
namespace tc {

constexpr int RT_BK = 64;
// W_hh rows per CTA = 2 halves x HS accumulator columns.  Default: HS = 32 (64 rows, 8 KB per K chunk); "narrow" slices
// for wide layers whose 64-row slice does not fit shared memory (H > ~1500): HS = 24 (48 rows, 6 KB per K chunk).
__host__ __device__ constexpr int rt_hs(bool narrow) { return narrow ? 24 : 32; }
__host__ __device__ constexpr int rt_rows(bool narrow) { return 2 * rt_hs(narrow); }
__host__ __device__ constexpr int rt_units(int gates, bool narrow) { return 2 * (rt_hs(narrow) / gates); }
constexpr int RT_GROUP = 4;        // most K chunks per ring slot / elected issue region
constexpr int RT_MAX_GROUPS = 4;   // barrier slots; the ring holds 2 groups of 4 chunks or up to 4 groups of 2
constexpr int RT_MAX_NIF = 3;      // batch groups in flight per CTA (each with its own TMEM accumulator)
// (Consecutive tcgen05.mma into the same accumulator do not stall each other -- tested with 4 independent
// accumulators: no change -- so a single 64-column TMEM accumulator per batch group is used.)
constexpr int RT_THREADS = 64 + 256 + 32;   // producer, MMA issuer, 8 epilogue warps, second producer
constexpr long long RT_TIMEOUT_CYCLES = 4000000000LL;
constexpr int RT_SMEM_LIMIT = 227 * 1024;

// Shared-memory plan: [W slice: nkc x 8 KB][h ring: slots x gsz stages x BP*128 B][h store staging][row times][barriers].
// The MMA is M = BP (64 or 128 batch rows): with M = 64 only the 64 valid rows are read from shared memory.
// Every elected issue region (elect.sync + single-lane branch + reconvergence) costs ~200 cycles on top
// of ~35 cycles per tcgen05.mma / TMA instruction (scripts/mma_microbench.py), so the producer and the MMA
// warp work in groups of `gsz` chunks: one region issues one TMA box of gsz chunks, one region issues 4*gsz MMAs.
// The ring slot is the unit of flow control: a slot is refilled when its MMAs have completed.  The h stream is bound
// by the SM's TMA intake (~35 B/clk: 152 KB per step at H = 1200), so the ring should keep the TMA unit busy ACROSS
// steps: with two slots the unit idles while the last two slots of a step drain; three slots of three chunks let it
// run ahead into the next group's step.  The kernel has no static shared memory, so the dynamic window starts on a
// 1 KB boundary and the plan may use all of the 227 KB.
struct RtPlan {
  int groups, gsz, stage_bytes, stage_off, stg_off, st_off, bar_off, total;
};
__host__ __device__ inline RtPlan rt_plan(int nkc, int BP, int U, int want_gsz = 0, bool narrow = false) {
  RtPlan pl;
  pl.stage_bytes = BP * RT_BK * 2;
  const int w_bytes = nkc * rt_rows(narrow) * RT_BK * 2;
  const int stg = (BP * U * 2 + 127) / 128 * 128;   // h (bf16) staging for coalesced stores
  const int st = BP * 4;                            // time index of every row of the group (-1 = inactive)
  const int room = RT_SMEM_LIMIT - 256 - w_bytes - stg - st;
  const int stages = room > 0 ? room / pl.stage_bytes : 0;
  int gsz, groups;
  if (want_gsz > 0) {
    gsz = want_gsz;
    groups = stages / gsz;
  } else if (stages >= 8) {      // (three slots of three chunks fit too and were measured slower: the cost is per box)
    gsz = 4; groups = 2;
  } else {
    gsz = 2; groups = stages / 2;
  }
  if (gsz > nkc) { gsz = nkc; groups = gsz ? stages / gsz : 0; }
  if (groups > RT_MAX_GROUPS) groups = RT_MAX_GROUPS;
  pl.groups = groups;
  pl.gsz = gsz;
  pl.stage_off = w_bytes;
  pl.stg_off = w_bytes + groups * gsz * pl.stage_bytes;
  pl.st_off = pl.stg_off + stg;
  pl.bar_off = pl.st_off + st;
  pl.total = pl.bar_off + 256;
  return pl;
}

struct RnnTcParams {
  const float* gx;          // [T*B][dirs*G*H]
  const float* b_hn;        // [dirs][H] GRU n-gate hidden bias (else nullptr)
  float* y;                 // [dirs][T][B][H]
  __nv_bfloat16* hbuf;      // [n_bgroups][2][dirs][BP][HP]
  const int32_t* lens;      // [B] sorted descending, or nullptr (every sequence runs Tmax steps)
  unsigned int* counters;   // [dirs][slots][NIF] step counters, kRnnCounterStride words apart
  int* abort_flag;
  const float* h0;          // [dirs][B][H] initial hidden state or nullptr (zeros)
  const float* c0;          // LSTM cell state, likewise
  float* hT;                // [dirs][B][H] final hidden state or nullptr
  float* cT;
  int B, H, HP, BP, T, Tmax;
  int dirs;   // directions in gx / y / hbuf / counters
  int dir0;   // first direction handled by this launch
  int cpd;    // CTAs per (direction, slot)
  int n_bgroups;   // the batch is processed in groups of BP rows ...
  int slots;       // ... by `slots` independent CTA sets per direction (set k takes groups k, k+slots, ...),
                   // NIF groups of a set in flight at a time
  int U;      // hidden units per CTA (2 * units per half)
  int ring_gsz;       // K chunks per ring slot, 0 = default (rt_plan)
  int n_producers;    // TMA producer warps (1 or 2)
  int nkc;    // K chunks of 64 (HP / 64)
  int skip;           // diagnostic (DSB_RNN_SKIP bit mask, results are then WRONG): 1 no gx loads, 2 no y stores, 4 no gate math
  long long ldt;      // > 0 (rnn_pair.cu only): gx is [dirs*G*H][ldt] and y is [dirs][H][ldt], column t*B + b (batch-minor)
  unsigned long long* dbg;   // optional [grid][128] cycle counters (DSB_RNN_DEBUG=1)
};

__device__ __forceinline__ bool wait_abortable(uint64_t* bar, uint32_t parity, int* abort_flag) {
  long long t0 = 0;
  unsigned n = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++n & 0xFF) == 0) {
      if (*(volatile int*)abort_flag) return false;
      long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > RT_TIMEOUT_CYCLES) {
        atomicExch(abort_flag, 1);
        return false;
      }
    }
  }
  return true;
}

__device__ __forceinline__ bool bar_red_and(bool pred, int id, int nthreads) {
  uint32_t r;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.u32 q, %1, 0;\n\t"
      "barrier.cta.red.and.pred p, %2, %3, q;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(r)
      : "r"((uint32_t)pred), "r"(id), "r"(nthreads)
      : "memory");
  return r != 0;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Gate non-linearities on the MUFU pipe: ex2.approx / rcp.approx with flush-to-zero, two MUFU operations each.
// (__expf / __fdividef expand to the same MUFU instructions plus range handling for denormal results -- a compare, two
// predicated multiplies and a squaring per call -- which a sigmoid / tanh does not need: 1 + denormal == 1.)
// tanh.approx.f32 is only good to ~5e-4 absolute, which is visible after 9 recurrent layers; the exp-based form is
// accurate to ~1e-6.
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_ftz(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_sigmoid(float x) { return rcp_ftz(1.0f + ex2_ftz(-1.4426950408889634f * x)); }
__device__ __forceinline__ float fast_tanh(float x) {
  return fmaf(-2.0f, rcp_ftz(ex2_ftz(2.8853900817779268f * x) + 1.0f), 1.0f);
}

// Steps of batch group bg: the lengths are sorted descending (forward() rejects anything else, like
// pack_padded_sequence behind model.py:117), so the group's first row is its longest sequence.
__device__ __forceinline__ int rt_group_steps(const RnnTcParams& p, int bg) {
  if (bg >= p.n_bgroups) return 0;
  if (!p.lens) return p.Tmax;
  const int l = p.lens[bg * p.BP];
  return l < p.Tmax ? l : p.Tmax;
}

}  // namespace tc
}  // namespace dsb
