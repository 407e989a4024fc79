// K-split persistent tensor-core recurrence (GRU / LSTM / tanh-RNN) on sm_100a: a CTA PAIR owns 128 rows of W_hh.
//
// Same contract as rnn_tc.cu (the sequential half of torch.nn.GRU/LSTM/RNN inside BatchRNN.forward,
// danspeech/deepspeech/model.py:114-122, packed-sequence semantics, both directions in one cooperative launch,
// batch groups of 64 sequences in flight per CTA).  What changes is the decomposition.  Measured on the one-CTA
// kernel (profiles/r02_recurrence.md): a step of one group is bound by the h all-gather -- every CTA pulls the whole
// h_{t-1} (64 x H bf16 = 152 KB at H = 1200) through its TMA port at ~35 B/clk, 4.7 k cycles -- and by 76 dependent
// M = 64 MMAs at the small-MMA floor.  Here the two CTAs of a cluster split K:
//   * the pair owns 2*UR hidden units = up to 128 rows of W_hh (all gates); CTA r keeps W[128 rows, K half r] resident
//     in shared memory (16 KB per K chunk of 64) and streams only ITS half of h_{t-1}: half the bytes per SM;
//   * the MMA is D[128 W rows, 64 batch] += W_chunk[128 x 16] * h_chunk[64 x 16]^T: M = 128 fills the tensor datapath,
//     half as many instructions per step (38 instead of 76 at H = 1200);
//   * each CTA then holds partial sums of all 128 rows over its K half.  Rows 64*rho..64*rho+63 belong to CTA rho's
//     units: the epilogue warps that hold the peer's rows push them into the peer's shared memory
//     (st.shared::cluster, 16 KB per step and direction), the others add the received partials to their own, and all
//     256 epilogue threads do the gate math on (unit, batch) pairs taken from shared memory -- consecutive threads take
//     consecutive units, so the pre-activation loads, the h stores and the y stores are (nearly) coalesced, which the
//     row-per-thread epilogue of the one-CTA kernel is not (16 cache lines per load instruction).
// Shared memory per CTA at H = 1200: W 160 KB + h ring 48 KB (2 x 3 chunks) + exchange 16 KB + barriers.
#include "tc_common.cuh"
#include "model_types.cuh"
#include <cstdlib>
#include <vector>

namespace dsb {
namespace tc {

constexpr int KS_BK = 64;
constexpr int KS_M = 128;                        // W_hh rows per CTA pair
constexpr int KS_N = 64;                         // sequences per batch group
constexpr int KS_W_BYTES = KS_M * KS_BK * 2;     // one resident W chunk (16 KB)
constexpr int KS_STAGE = KS_N * KS_BK * 2;       // one h chunk (8 KB)
constexpr int KS_XS = 68;                        // floats per row of the exchange buffer: 64 batch + 4 (bank spread)
constexpr int KS_X_BYTES = 64 * KS_XS * 4;       // partial-sum exchange [64 rows][64 batch (+4)] fp32
constexpr int KS_MAX_SLOTS = 4;
constexpr int KS_MAX_NIF = 3;
constexpr int KS_THREADS = 64 + 256 + 32;   // producer, MMA issuer, 8 epilogue warps, publisher
constexpr long long KS_TIMEOUT_CYCLES = 4000000000LL;
constexpr int KS_SMEM_LIMIT = 227 * 1024;

constexpr int KS_LEN_BYTES = KS_MAX_NIF * KS_N * 4;   // sequence lengths of the groups in flight
struct KsPlan {
  int slots, gsz, ring_off, x_off, len_off, bar_off, total;
};
// half = K chunks of 64 per CTA; want_gsz = chunks per ring slot (0 = default).  A ring slot is one TMA box and one
// elected MMA region; the ring is as deep as shared memory allows: with few large slots the MMAs of slot g wait for
// (MMA completion of slot g-2 + TMA latency), with more small ones they run back to back.
__host__ __device__ inline KsPlan ks_plan(int half, int want_gsz = 0) {
  KsPlan pl;
  const int w_bytes = half * KS_W_BYTES;
  const int room = KS_SMEM_LIMIT - 1024 - 256 - KS_LEN_BYTES - w_bytes - KS_X_BYTES;   // 1 KB alignment slack, barriers
  const int stages = room > 0 ? room / KS_STAGE : 0;
  int gsz = want_gsz > 0 ? want_gsz : 2;
  if (gsz > 4) gsz = 4;
  while (gsz > 1 && stages / gsz < 2) --gsz;
  if (gsz > half) gsz = half;
  int slots = gsz > 0 ? stages / gsz : 0;
  if (slots > KS_MAX_SLOTS) slots = KS_MAX_SLOTS;
  pl.slots = slots;
  pl.gsz = gsz;
  pl.ring_off = w_bytes;
  pl.x_off = w_bytes + slots * gsz * KS_STAGE;
  pl.len_off = pl.x_off + KS_X_BYTES;
  pl.bar_off = pl.len_off + KS_LEN_BYTES;
  pl.total = pl.bar_off + 256 + 1024;
  return pl;
}

struct KsParams {
  const float* gx;          // [T*B][dirs*G*H]
  const float* b_hn;        // [dirs][H] GRU n-gate hidden bias (else nullptr)
  float* y;                 // [dirs][T][B][H]
  __nv_bfloat16* hbuf;      // [n_bgroups][2][dirs][64][HP]
  const int32_t* lens;      // [B] sorted descending, or nullptr
  unsigned int* counters;   // [dirs][slots][NIF], kRnnCounterStride words apart
  int* abort_flag;
  const float* h0;
  const float* c0;
  float* hT;
  float* cT;
  int B, H, HP, T, Tmax;
  int dirs, dir0;
  int pairs;       // CTA pairs per (direction, slot)
  int n_bgroups, slots;
  int nkc, half;   // K chunks of 64 in total / per CTA
  int ring_gsz;    // chunks per ring slot (ks_plan)
  unsigned long long* dbg;
};

__device__ __forceinline__ bool ks_wait(uint64_t* bar, uint32_t parity, int* abort_flag) {
  long long t0 = 0;
  unsigned n = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++n & 0xFF) == 0) {
      if (*(volatile int*)abort_flag) return false;
      long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > KS_TIMEOUT_CYCLES) {
        atomicExch(abort_flag, 1);
        return false;
      }
    }
  }
  return true;
}
// same at cluster scope, relaxed: the barrier is arrived on by the peer CTA and only orders the REUSE of a buffer whose
// reads completed before the arrive (no data travels with it)
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.relaxed.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool ks_wait_cluster(uint64_t* bar, uint32_t parity, int* abort_flag) {
  long long t0 = 0;
  unsigned n = 0;
  while (!mbar_try_wait_cluster(bar, parity)) {
    if ((++n & 0xFF) == 0) {
      if (*(volatile int*)abort_flag) return false;
      long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > KS_TIMEOUT_CYCLES) {
        atomicExch(abort_flag, 1);
        return false;
      }
    }
  }
  return true;
}
// Asynchronous store into the peer CTA's shared memory that counts its bytes on the PEER's mbarrier when it lands
// (SASS: STAS): producer/consumer hand-over through distributed shared memory without any fence -- a cluster-scope
// release / acquire pair costs a MEMBAR + CCTL.IVALL (L1 invalidate) in every participating thread per item.
__device__ __forceinline__ void st_async_v4(uint32_t addr, uint32_t mbar, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(addr),
               "r"(a), "r"(b), "r"(c), "r"(d), "r"(mbar)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t bar_cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
__device__ __forceinline__ float4 lds_v4(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts_v4(uint32_t a, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// explicit shared-space accesses through 32-bit addresses: no 64-bit generic pointers to keep alive (the epilogue runs
// at the register cap, and a spilled pointer costs an L2 round trip per reload -- every cluster-scope acquire /
// release of the exchange invalidates L1, local memory included)
__device__ __forceinline__ float lds_f32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts_f32(uint32_t a, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}
__device__ __forceinline__ int lds_s32(uint32_t a) {
  int v;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ bool ks_bar_red_and(bool pred, int id, int nthreads) {
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
__device__ __forceinline__ void ks_named_bar(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ unsigned ks_ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void ks_red_release(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float ks_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float ks_tanh(float x) { return 1.0f - __fdividef(2.0f, __expf(2.0f * x) + 1.0f); }

__device__ __forceinline__ int ks_group_steps(const KsParams& p, int bg) {
  if (bg >= p.n_bgroups) return 0;
  if (!p.lens) return p.Tmax;
  const int l = p.lens[bg * KS_N];
  return l < p.Tmax ? l : p.Tmax;
}

// units per CTA: as many as fit 64 rows, a multiple of 4
__host__ __device__ constexpr int ks_units(int gates) { return (64 / gates) / 4 * 4; }

template <int GATES, int NIF>
__global__ void __launch_bounds__(KS_THREADS, 1)
rnn_ks_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_h,
              const KsParams p) {
  constexpr int UR = ks_units(GATES);            // units per CTA
  static_assert(UR % 4 == 0, "four epilogue threads share a batch row");
  constexpr int TMEM_COLS = NIF == 1 ? 64 : (NIF == 2 ? 128 : 256);
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  const KsPlan pl = ks_plan(p.half, p.ring_gsz);
  unsigned char* sW = smem;
  unsigned char* sA = smem + pl.ring_off;
  float* sX = reinterpret_cast<float*>(smem + pl.x_off);
  int* sLen = reinterpret_cast<int*>(smem + pl.len_off);             // [KS_MAX_NIF][64] lengths of the groups in flight
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + pl.bar_off);   // [KS_MAX_SLOTS] ring slot landed
  uint64_t* gempty = full + KS_MAX_SLOTS;                            // [KS_MAX_SLOTS] ring slot consumed
  uint64_t* wbar = gempty + KS_MAX_SLOTS;
  uint64_t* dfull = wbar + 1;                                        // [KS_MAX_NIF] accumulator complete
  uint64_t* xfull = dfull + KS_MAX_NIF;     // the peer's partial sums for my rows have landed in sX (16 KB of st.async)
  uint64_t* xfree = xfull + 1;              // the peer is done with ITS sX: I may overwrite it (1 arrival)
  uint64_t* hdone = xfree + 1;              // [4] all 256 epilogue threads have issued the h stores of item ic (slot ic & 3)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(hdone + 4);
  const int n_slots = pl.slots, gsz = pl.gsz;
  const int gps = (p.half + gsz - 1) / gsz;   // ring-slot uses per item

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per_set = 2 * p.pairs;
  const int set = blockIdx.x / per_set;            // (direction, slot)
  const int dir = p.dir0 + set / p.slots;
  const int slot = set % p.slots;
  const int pair = (blockIdx.x % per_set) >> 1;
  const int rank = (int)cluster_ctarank();
  const int g_rot = (int)(((long long)pair * gps) / p.pairs);
  unsigned* const ctr0 = p.counters + (size_t)((dir * p.slots + slot) * NIF) * kRnnCounterStride;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_w);
    prefetch_tmap(&tmap_h);
    for (int i = 0; i < KS_MAX_SLOTS; ++i) mbar_init(&full[i], 1);
    for (int i = 0; i < KS_MAX_SLOTS; ++i) mbar_init(&gempty[i], 1);
    mbar_init(wbar, 1);
    for (int i = 0; i < KS_MAX_NIF; ++i) mbar_init(&dfull[i], 1);
    mbar_init(xfull, 1);       // one local arrive.expect_tx per item + the bytes of the peer's st.async
    mbar_init(xfree, 1);
    for (int i = 0; i < 4; ++i) mbar_init(&hdone[i], 256);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();      // the peer's barriers exist before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    // ---- TMA producer ----
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(wbar, (uint32_t)p.half * KS_W_BYTES);
      const int row0 = ((dir * p.pairs + pair) * 2 + rank) * KS_M;   // [dirs][pairs][2 ranks][128 rows]
      for (int kc = 0; kc < p.half; ++kc) tma_load_2d(sW + (size_t)kc * KS_W_BYTES, &tmap_w, wbar, kc * KS_BK, row0);
    }
    __syncwarp();
    bool ok = true;
    unsigned long long d_spin = 0, d_issue = 0, d_empty = 0;
    int arm_slot = 0, load_slot = 0;
    uint32_t arm_phase = 0;
    auto arm_group = [&]() -> bool {
      const int grp = arm_slot;
      if (!__all_sync(0xffffffffu, ks_wait(&gempty[grp], arm_phase ^ 1, p.abort_flag))) return false;
      if (elect_one_sync()) mbar_arrive_expect_tx(&full[grp], (uint32_t)(gsz * KS_STAGE));
      __syncwarp();
      if (++arm_slot == n_slots) { arm_slot = 0; arm_phase ^= 1; }
      return true;
    };
    auto item = [&](int i, int s, int bg, unsigned steps_before) -> bool {
      long long c0 = clock64();
      if (steps_before + (unsigned)s > 0) {
        const unsigned target = (unsigned)per_set * (steps_before + (unsigned)s);
        const unsigned* ctr = ctr0 + i * kRnnCounterStride;
        long long t0 = 0;
        unsigned n = 0;
        bool good = true;
        while (ks_ld_acquire(ctr) < target) {
          if ((++n & 0x3F) == 0) {
            if (*(volatile int*)p.abort_flag) { good = false; break; }
            long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > KS_TIMEOUT_CYCLES) { atomicExch(p.abort_flag, 1); good = false; break; }
          }
        }
        if (!__all_sync(0xffffffffu, good)) return false;
        asm volatile("fence.proxy.async.global;" ::: "memory");   // generic-proxy writes -> async-proxy (TMA) reads
      }
      long long c1 = clock64();
      d_spin += c1 - c0;
      const int row0 = ((bg * 2 + (s & 1)) * p.dirs + dir) * KS_N;
      for (int g = 0; g < gps; ++g) {
        long long w0 = clock64();
        if (!arm_group()) return false;
        d_empty += clock64() - w0;
        const int grp = load_slot;
        if (++load_slot == n_slots) load_slot = 0;
        int gg = g + g_rot;
        if (gg >= gps) gg -= gps;
        if (elect_one_sync())
          tma_load_3d(sA + grp * gsz * KS_STAGE, &tmap_h, &full[grp], 0, row0, rank * p.half + gg * gsz);
        __syncwarp();
      }
      d_issue += clock64() - c1;
      return true;
    };
    unsigned before[NIF];
#pragma unroll
    for (int i = 0; i < NIF; ++i) before[i] = 0;
    for (int k0 = 0; ok && slot + k0 * p.slots < p.n_bgroups; k0 += NIF) {
      int Tg[NIF], Tw = 0;
#pragma unroll
      for (int i = 0; i < NIF; ++i) {
        Tg[i] = ks_group_steps(p, slot + (k0 + i) * p.slots);
        Tw = max(Tw, Tg[i]);
      }
      for (int s = 0; s < Tw && ok; ++s) {
#pragma unroll
        for (int i = 0; i < NIF; ++i)
          if (ok && s < Tg[i]) ok = item(i, s, slot + (k0 + i) * p.slots, before[i]);
      }
#pragma unroll
      for (int i = 0; i < NIF; ++i) before[i] += (unsigned)Tg[i];
    }
    if (p.dbg && lane == 0) {
      p.dbg[blockIdx.x * 16 + 0] = d_spin;
      p.dbg[blockIdx.x * 16 + 1] = d_issue;
      p.dbg[blockIdx.x * 16 + 2] = d_empty;
    }
  } else if (warp == 1) {
    // ---- MMA issuer: D[128 W rows, 64 batch] += W_chunk * h_chunk^T ----
    const uint32_t idesc = make_idesc_bf16(KS_M, KS_N);
    const uint64_t desc0 = make_smem_desc(0, 16, 1024, 2);
    const uint32_t a_lo = smem_u32(sA) >> 4, w_lo = smem_u32(sW) >> 4;
    bool ok = __all_sync(0xffffffffu, ks_wait(wbar, 0, p.abort_flag));
    unsigned long long d_wait0 = 0, d_rest = 0;
    int grp = 0;
    uint32_t fphase = 0;
    auto item = [&](int i) -> bool {
      long long m0 = clock64();
      const uint32_t d_tmem = tmem_base + (uint32_t)(i * 64);
      for (int g = 0; g < gps; ++g) {
        int gg = g + g_rot;
        if (gg >= gps) gg -= gps;
        const int i0 = gg * gsz, nch = min(gsz, p.half - i0);
        if (!__all_sync(0xffffffffu, ks_wait(&full[grp], fphase, p.abort_flag))) return false;
        if (g == 0) { long long m1 = clock64(); d_wait0 += m1 - m0; m0 = m1; }
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t h0 = a_lo + (uint32_t)(grp * gsz) * (KS_STAGE >> 4);
          const uint32_t w0 = w_lo + (uint32_t)i0 * (KS_W_BYTES >> 4);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (j < nch) {
              const uint64_t wdesc = desc0 + (uint64_t)(w0 + (uint32_t)j * (KS_W_BYTES >> 4));
              const uint64_t hdesc = desc0 + (uint64_t)(h0 + (uint32_t)j * (KS_STAGE >> 4));
#pragma unroll
              for (int k = 0; k < KS_BK / 16; ++k)
                umma_bf16(d_tmem, wdesc + (uint64_t)(k * 2), hdesc + (uint64_t)(k * 2), idesc,
                          (j | k) ? 1u : (uint32_t)(g != 0));
            }
          }
          umma_commit(&gempty[grp]);
          if (g == gps - 1) umma_commit(&dfull[i]);
        }
        __syncwarp();
        if (++grp == n_slots) { grp = 0; fphase ^= 1; }
      }
      d_rest += clock64() - m0;
      return true;
    };
    for (int k0 = 0; ok && slot + k0 * p.slots < p.n_bgroups; k0 += NIF) {
      int Tg[NIF], Tw = 0;
#pragma unroll
      for (int i = 0; i < NIF; ++i) {
        Tg[i] = ks_group_steps(p, slot + (k0 + i) * p.slots);
        Tw = max(Tw, Tg[i]);
      }
      for (int s = 0; s < Tw && ok; ++s) {
#pragma unroll
        for (int i = 0; i < NIF; ++i)
          if (ok && s < Tg[i]) ok = item(i);
      }
    }
    if (p.dbg && lane == 0) {
      p.dbg[blockIdx.x * 16 + 3] = d_wait0;
      p.dbg[blockIdx.x * 16 + 4] = d_rest;
    }
  } else if (warp == 10) {
    // ---- publisher: once all epilogue threads have issued the h stores of an item, publish the step (release at GPU
    // scope: waits for the acknowledgements of those stores, ~2 k cycles) and hand the exchange buffer back to the
    // peer.  On its own warp this round trip is off the epilogue warps' path: they go straight on to the next item.
    bool ok = true;
    unsigned ic = 0;
    unsigned long long d_pub = 0;
    const uint32_t peer_xfree = mapa_u32(smem_u32(xfree), (uint32_t)(rank ^ 1));
    for (int k0 = 0; ok && slot + k0 * p.slots < p.n_bgroups; k0 += NIF) {
      int Tg[NIF], Tw = 0;
#pragma unroll
      for (int i = 0; i < NIF; ++i) {
        Tg[i] = ks_group_steps(p, slot + (k0 + i) * p.slots);
        Tw = max(Tw, Tg[i]);
      }
      for (int s = 0; s < Tw && ok; ++s) {
#pragma unroll
        for (int i = 0; i < NIF; ++i)
          if (ok && s < Tg[i]) {
            ok = __all_sync(0xffffffffu, ks_wait(&hdone[ic & 3], (ic >> 2) & 1u, p.abort_flag));
            if (ok && lane == 0) {
              long long c0 = clock64();
              mbar_arrive_cluster_relaxed(peer_xfree);            // the peer may overwrite my sX with its next item
              ks_red_release(ctr0 + i * kRnnCounterStride, 1u);   // publish h_t (release: cumulative over the arrivals)
              d_pub += clock64() - c0;
            }
            ++ic;
          }
      }
    }
    if (p.dbg && lane == 0) p.dbg[blockIdx.x * 16 + 10] = d_pub;
  } else {
    // ---- epilogue: 8 warps.  Accumulator row m = 32*q + lane (q = warp % 4) is W row m of the pair; rows
    //      [64*rho, 64*rho + 64) hold the gates of CTA rho's units; the two warps of a quarter split the 64 batch columns.
    const int et = threadIdx.x - 64;     // 0..255
    const int q = warp & 3;
    const int ch = (warp - 2) >> 2;
    const int j = ((q & 1) << 5) + lane;           // row inside its owner's 64
    const bool is_own = (q >> 1) == rank;
    const bool x_leader = is_own && ch == 0 && (q & 1) == 0 && lane == 0;
    const int cb = ch * 32;                        // first batch column of this thread
    const int ncol = p.dirs * GATES * p.H;
    const int unit0 = (pair * 2 + rank) * UR;
    // gate math: thread (bq = et / 4, uq = et % 4) owns UPT consecutive units of batch row bq -- one length, one time
    // index and one set of row base pointers per thread and item (the instruction count of the epilogue is what the
    // eight warps compete for: every instruction here is issued 8 x per item)
    constexpr int UPT = UR / 4;
    const int bq = et >> 2, uq = et & 3;
    const int ubase = unit0 + uq * UPT;
    const int nvalid = min(UPT, max(0, p.H - ubase));          // units of this thread inside H
    float bhn[UPT];
#pragma unroll
    for (int u = 0; u < UPT; ++u)
      bhn[u] = (GATES == 3 && p.b_hn && u < nvalid) ? p.b_hn[(size_t)dir * p.H + ubase + u] : 0.f;
    float hprev[NIF][UPT], cst[NIF][UPT];
    const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)cb;
    const uint32_t sx = smem_u32(sX), slen = smem_u32(sLen);
    const uint32_t my_x = sx + (uint32_t)((j * KS_XS + cb) * 4);                    // (row j, batch cb) of my sX
    const uint32_t peer_x = mapa_u32(sx, (uint32_t)(rank ^ 1)) + (uint32_t)((j * KS_XS + cb) * 4);
    const uint32_t peer_xfull = mapa_u32(smem_u32(xfull), (uint32_t)(rank ^ 1));
    const uint32_t gate_x = sx + (uint32_t)((uq * UPT * GATES * KS_XS + bq) * 4);   // (unit uq*UPT, gate 0, batch bq)
    unsigned ic = 0;   // items processed (both CTAs of the pair walk the same item sequence)
    unsigned long long e_load = 0, e_wait = 0, e_xchg = 0, e_math = 0, e_pub = 0, e_len = 0;

    auto item = [&](int i, int s, int bg, unsigned steps_before, float (&hp)[UPT], float (&cs)[UPT]) -> bool {
      long long e0 = clock64();
      // 1. input-projection pre-activations of this step
      float gxv[UPT][GATES];
      const int len = nvalid > 0 ? lds_s32(slen + (uint32_t)((i * KS_N + bq) * 4)) : 0;   // 0 for rows beyond the batch
      const bool act = s < len;
      const int t = dir == 0 ? s : len - 1 - s;
      const size_t brow = (size_t)(bg * KS_N + bq);
      long long e0b = clock64();
      if (act) {
        const float* gp = p.gx + ((size_t)t * p.B + brow) * ncol + (size_t)dir * GATES * p.H + ubase;
#pragma unroll
        for (int g = 0; g < GATES; ++g)
#pragma unroll
          for (int u = 0; u < UPT; ++u) gxv[u][g] = u < nvalid ? __ldg(gp + (size_t)g * p.H + u) : 0.f;
        if (s + 1 < len) {   // the group's next step: pulled into L2 one round of items ahead
          const float* gn = gp + (dir == 0 ? (ptrdiff_t)p.B * ncol : -(ptrdiff_t)p.B * ncol);
#pragma unroll
          for (int g = 0; g < GATES; ++g) asm volatile("prefetch.global.L2 [%0];" ::"l"(gn + (size_t)g * p.H));
        }
      }
      long long e1 = clock64();
      // 2. accumulator of this item
      bool ok = ks_wait(&dfull[i], (uint32_t)((steps_before + (unsigned)s) & 1u), p.abort_flag);
      long long e2 = clock64();
      tc_fence_after();
      // 3. partial-sum exchange: my K half of the peer's rows goes to the peer, the peer's half of my rows is added
      if (!is_own) {
        if (ic > 0) ok = ks_wait_cluster(xfree, (ic - 1) & 1u, p.abort_flag) && ok;   // the peer has consumed item ic-1
      } else {
        if (x_leader) mbar_arrive_expect_tx(xfull, 64u * 64u * 4u);                    // this item's 16 KB from the peer
        ok = ks_wait(xfull, ic & 1u, p.abort_flag) && ok;
      }
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t r[16];
        tmem_ld16(t_addr + (uint32_t)(i * 64 + hh * 16), r);
        tmem_ld_wait();
        if (ok) {
          // 16-byte accesses: a quarter warp touches 8 rows 272 bytes apart = all 32 banks once
          if (!is_own) {
#pragma unroll
            for (int c = 0; c < 16; c += 4)
              st_async_v4(peer_x + (uint32_t)((hh * 16 + c) * 4), peer_xfull, r[c], r[c + 1], r[c + 2], r[c + 3]);
          } else {
#pragma unroll
            for (int c = 0; c < 16; c += 4) {
              const uint32_t a = my_x + (uint32_t)((hh * 16 + c) * 4);
              float4 v = lds_v4(a);
              v.x += __uint_as_float(r[c]); v.y += __uint_as_float(r[c + 1]);
              v.z += __uint_as_float(r[c + 2]); v.w += __uint_as_float(r[c + 3]);
              sts_v4(a, v);
            }
          }
        }
      }
      tc_fence_before();
      long long e3 = clock64();
      if (!ks_bar_red_and(ok, 1, 256)) return false;     // sX complete (and uniform abort decision)
      // 4. gates
      if (act) {
        __nv_bfloat16* hb = p.hbuf + ((size_t)((bg * 2 + ((s + 1) & 1)) * p.dirs + dir) * KS_N + bq) * p.HP + ubase;
#pragma unroll
        for (int u = 0; u < UPT; ++u) {
          if (u < nvalid) {
            const uint32_t a = gate_x + (uint32_t)(u * GATES * KS_XS * 4);
            float hn;
            if (GATES == 3) {
              const float rg = ks_sigmoid(gxv[u][0] + lds_f32(a));
              const float zg = ks_sigmoid(gxv[u][1 % GATES] + lds_f32(a + 4 * KS_XS * (1 % GATES)));
              const float ng = ks_tanh(gxv[u][2 % GATES] + rg * (lds_f32(a + 4 * KS_XS * (2 % GATES)) + bhn[u]));
              hn = (1.0f - zg) * ng + zg * hp[u];
            } else if (GATES == 4) {
              const float ig = ks_sigmoid(gxv[u][0] + lds_f32(a));
              const float fg = ks_sigmoid(gxv[u][1 % GATES] + lds_f32(a + 4 * KS_XS * (1 % GATES)));
              const float gg = ks_tanh(gxv[u][2 % GATES] + lds_f32(a + 4 * KS_XS * (2 % GATES)));
              const float og = ks_sigmoid(gxv[u][3 % GATES] + lds_f32(a + 4 * KS_XS * (3 % GATES)));
              cs[u] = fg * cs[u] + ig * gg;
              hn = og * ks_tanh(cs[u]);
            } else {
              hn = ks_tanh(gxv[u][0] + lds_f32(a));
            }
            hp[u] = hn;
            hb[u] = __float2bfloat16_rn(hn);     // h_t -> exchange buffer of the next step
          }
        }
      }
      long long e4 = clock64();
      // all reads of sX done and all h stores issued: the publisher warp takes it from here (release = this arrive)
      mbar_arrive(&hdone[ic & 3]);
      // y_t -> global (fp32): nobody waits on these stores
      if (act) {
        float* yo = p.y + (((size_t)dir * p.T + t) * p.B + brow) * p.H + ubase;
#pragma unroll
        for (int u = 0; u < UPT; ++u)
          if (u < nvalid) yo[u] = hp[u];
      }
      ++ic;
      e_len += e0b - e0;
      e_load += e1 - e0; e_wait += e2 - e1; e_xchg += e3 - e2; e_math += e4 - e3; e_pub += clock64() - e4;
      return true;
    };

    unsigned before[NIF];
#pragma unroll
    for (int i = 0; i < NIF; ++i) before[i] = 0;
    bool alive = true;
    for (int k0 = 0; alive && slot + k0 * p.slots < p.n_bgroups; k0 += NIF) {
      int Tg[NIF], Tw = 0;
      ks_named_bar(2, 256);      // every thread is done with the previous wave's lengths
#pragma unroll
      for (int i = 0; i < NIF; ++i) {
        const int bg = slot + (k0 + i) * p.slots;
        Tg[i] = ks_group_steps(p, bg);
        Tw = max(Tw, Tg[i]);
        if (et < KS_N) {
          const int b = bg * KS_N + et;
          sLen[i * KS_N + et] = (bg < p.n_bgroups && b < p.B) ? (p.lens ? p.lens[b] : p.Tmax) : 0;   // read back with lds_s32
        }
#pragma unroll
        for (int u = 0; u < UPT; ++u) {
          const int b = bg * KS_N + bq;
          const bool in = bg < p.n_bgroups && b < p.B && u < nvalid;
          hprev[i][u] = (p.h0 && in) ? p.h0[((size_t)dir * p.B + b) * p.H + ubase + u] : 0.f;
          cst[i][u] = (GATES == 4 && p.c0 && in) ? p.c0[((size_t)dir * p.B + b) * p.H + ubase + u] : 0.f;
        }
      }
      ks_named_bar(2, 256);
      for (int s = 0; s < Tw && alive; ++s) {
#pragma unroll
        for (int i = 0; i < NIF; ++i)
          if (alive && s < Tg[i]) alive = item(i, s, slot + (k0 + i) * p.slots, before[i], hprev[i], cst[i]);
      }
#pragma unroll
      for (int i = 0; i < NIF; ++i) {
        before[i] += (unsigned)Tg[i];
        const int bg = slot + (k0 + i) * p.slots;
        if (alive && bg < p.n_bgroups && (p.hT || p.cT)) {   // carry the state out (streaming)
          const int b = bg * KS_N + bq;
#pragma unroll
          for (int u = 0; u < UPT; ++u) {
            if (b < p.B && u < nvalid) {
              if (p.hT) p.hT[((size_t)dir * p.B + b) * p.H + ubase + u] = hprev[i][u];
              if (GATES == 4 && p.cT) p.cT[((size_t)dir * p.B + b) * p.H + ubase + u] = cst[i][u];
            }
          }
        }
      }
    }
    if (p.dbg && et == 64) {
      p.dbg[blockIdx.x * 16 + 5] = e_load;
      p.dbg[blockIdx.x * 16 + 6] = e_wait;
      p.dbg[blockIdx.x * 16 + 7] = e_xchg;
      p.dbg[blockIdx.x * 16 + 8] = e_math;
      p.dbg[blockIdx.x * 16 + 9] = e_pub;
      p.dbg[blockIdx.x * 16 + 11] = e_len;
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();     // nobody leaves while the peer may still store into / arrive on this CTA
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// W_hh [dirs][G*H][H] fp32 -> per-CTA slices [dirs][pairs][2 ranks][128 rows][half*64] bf16.  Row m of a pair:
// rho = m/64 (owning CTA), jj = m%64, u = jj/G, g = jj%G  <->  W_hh[g*H + (pair*2+rho)*UR + u][rank*half*64 + k].
__global__ void pack_whh_ks_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int dirs, int pairs,
                                   int G, int H, int UR, int half) {
  const int KH = half * KS_BK;
  const int64_t total = (int64_t)dirs * pairs * 2 * KS_M * KH;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % KH);
    int64_t rr = i / KH;
    const int m = (int)(rr % KS_M);
    rr /= KS_M;
    const int rank = (int)(rr % 2);
    rr /= 2;
    const int pr = (int)(rr % pairs);
    const int d = (int)(rr / pairs);
    const int rho = m / 64, jj = m % 64, u = jj / G, g = jj % G;
    const int unit = (pr * 2 + rho) * UR + u;
    const int kk = rank * KH + k;
    float v = 0.f;
    if (u < UR && unit < H && kk < H) v = w[((int64_t)d * G * H + (int64_t)g * H + unit) * H + kk];
    out[i] = __float2bfloat16_rn(v);
  }
}

}  // namespace tc

static int ks_half(int H) { return cdiv(cdiv(H, 64), 2); }

bool rnn_ks_supported(const RnnLayer& L, int sms, int* pairs_out, int* launches_out) {
  const int UR = tc::ks_units(L.gates);
  const int pairs = cdiv(L.H, 2 * UR);
  const tc::KsPlan pl = tc::ks_plan(ks_half(L.H), g_tune.rnn_ring_gsz.load());
  if (pl.slots < 2 || pl.gsz < 1 || pl.total > tc::KS_SMEM_LIMIT || 2 * pairs > sms) return false;
  if (pairs_out) *pairs_out = pairs;
  if (launches_out) *launches_out = (L.dirs * 2 * pairs <= sms) ? 1 : L.dirs;
  return true;
}

size_t rnn_ks_pack_elems(const RnnLayer& L) {
  const int UR = tc::ks_units(L.gates);
  return (size_t)L.dirs * cdiv(L.H, 2 * UR) * 2 * tc::KS_M * ks_half(L.H) * tc::KS_BK;
}

int pack_whh_ks(const RnnLayer& L, __nv_bfloat16* out, cudaStream_t st) {
  const int UR = tc::ks_units(L.gates), pairs = cdiv(L.H, 2 * UR), half = ks_half(L.H);
  const int64_t total = (int64_t)rnn_ks_pack_elems(L);
  tc::pack_whh_ks_kernel<<<(int)(cdiv64(total, 256) < 2048 ? cdiv64(total, 256) : 2048), 256, 0, st>>>(
      L.w_hh, out, L.dirs, pairs, L.gates, L.H, UR, half);
  DSB_CHECK_LAUNCH();
  return 0;
}

// One BatchRNN layer on CTA pairs; arguments as rnn_layer_tc (hbuf: rnn_tc_hbuf_elems() with groups of 64 rows).
int rnn_layer_ks(const RnnLayer& L, const float* gx, const int32_t* d_len, int B, int T, int Tmax, float* y,
                 __nv_bfloat16* hbuf, unsigned int* sync_words, int* abort_flag, cudaStream_t st, const float* h0,
                 const float* c0, float* hT, float* cT) {
  using namespace tc;
  int dev = 0, sms = 148, pairs = 0, launches = 1;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (!rnn_ks_supported(L, sms, &pairs, &launches))
    return set_error(DSB_ERR_UNSUPPORTED, "rnn_layer_ks: shape H=%d not supported", L.H);
  const int HP = (L.H + 63) / 64 * 64, nkc = HP / 64, half = ks_half(L.H);
  const int n_bgroups = cdiv(B, KS_N);
  const int dirs_per_launch = L.dirs / launches;
  int slots = sms / (dirs_per_launch * 2 * pairs);
  if (slots > n_bgroups) slots = n_bgroups;
  if (g_tune.rnn_max_slots.load() > 0 && slots > g_tune.rnn_max_slots.load()) slots = g_tune.rnn_max_slots.load();
  if (slots < 1) slots = 1;
  int nif = cdiv(n_bgroups, slots);
  if (nif > rnn_tc_max_in_flight()) nif = rnn_tc_max_in_flight();
  while (slots > 1 && L.dirs * slots * nif > kRnnMaxCounters) --slots;
  DSB_CUDA(cudaMemsetAsync(hbuf, 0, sizeof(__nv_bfloat16) * (size_t)n_bgroups * 2 * L.dirs * KS_N * HP, st));
  DSB_CUDA(cudaMemsetAsync(sync_words, 0, sizeof(unsigned int) * kRnnSyncCounters, st));
  if (h0)
    if (int e = rnn_tc_init_hbuf(h0, hbuf, L.dirs, B, L.H, HP, KS_N, n_bgroups, st)) return e;

  const int ring_gsz = g_tune.rnn_ring_gsz.load();
  const KsPlan pl = ks_plan(half, ring_gsz);
  CUtensorMap tw, th;
  uint64_t dw[2] = {(uint64_t)half * KS_BK, (uint64_t)L.dirs * pairs * 2 * KS_M}, sw[2] = {2, (uint64_t)half * KS_BK * 2};
  uint32_t bw[2] = {KS_BK, KS_M};
  if (int e = make_tmap_bf16(&tw, L.w_hh_pack_ks, 2, dw, sw, bw, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
  uint64_t dh[3] = {(uint64_t)KS_BK, (uint64_t)n_bgroups * 2 * L.dirs * KS_N, (uint64_t)nkc};
  uint64_t sh[3] = {2, (uint64_t)HP * 2, (uint64_t)KS_BK * 2};
  uint32_t bh[3] = {KS_BK, (uint32_t)KS_N, (uint32_t)pl.gsz};
  if (int e = make_tmap_bf16(&th, hbuf, 3, dh, sh, bh, CU_TENSOR_MAP_SWIZZLE_128B)) return e;

  KsParams p{};
  p.gx = gx; p.b_hn = L.b_hn; p.y = y; p.hbuf = hbuf; p.lens = d_len;
  p.counters = sync_words; p.abort_flag = abort_flag;
  p.h0 = h0; p.c0 = c0; p.hT = hT; p.cT = cT;
  p.B = B; p.H = L.H; p.HP = HP; p.T = T; p.Tmax = Tmax;
  p.dirs = L.dirs; p.pairs = pairs; p.n_bgroups = n_bgroups; p.slots = slots; p.nkc = nkc; p.half = half; p.ring_gsz = ring_gsz;
  const void* fn = nullptr;
#define KS_PICK(G) \
  fn = nif == 1 ? (const void*)rnn_ks_kernel<G, 1> : nif == 2 ? (const void*)rnn_ks_kernel<G, 2> : (const void*)rnn_ks_kernel<G, 3>
  if (L.gates == 3) KS_PICK(3);
  else if (L.gates == 4) KS_PICK(4);
  else KS_PICK(1);
#undef KS_PICK
  DSB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, pl.total));
  static const bool debug = getenv("DSB_RNN_DEBUG") != nullptr;
  const int grid = dirs_per_launch * slots * 2 * pairs;
  unsigned long long* dbg = nullptr;
  if (debug) {
    DSB_CUDA(cudaMalloc(&dbg, sizeof(unsigned long long) * 16 * grid));
    DSB_CUDA(cudaMemsetAsync(dbg, 0, sizeof(unsigned long long) * 16 * grid, st));
  }
  p.dbg = dbg;
  for (int l = 0; l < launches; ++l) {
    p.dir0 = l * dirs_per_launch;
    void* args[] = {(void*)&tw, (void*)&th, (void*)&p};
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(KS_THREADS);
    cfg.dynamicSmemBytes = pl.total;
    cfg.stream = st;
    cudaLaunchAttribute attrs[2];
    attrs[0].id = cudaLaunchAttributeCooperative;          // all CTAs co-resident (they spin on each other)
    attrs[0].val.cooperative = 1;
    attrs[1].id = cudaLaunchAttributeClusterDimension;
    attrs[1].val.clusterDim.x = 2;
    attrs[1].val.clusterDim.y = 1;
    attrs[1].val.clusterDim.z = 1;
    cfg.attrs = attrs;
    cfg.numAttrs = 2;
    const cudaError_t le = cudaLaunchKernelExC(&cfg, fn, args);
    if (le != cudaSuccess)
      return set_error(DSB_ERR_CUDA, "rnn_layer_ks: launch failed: %s", cudaGetErrorString(le));
    count_launch();
  }
  if (debug) {
    std::vector<unsigned long long> h(16 * grid);
    DSB_CUDA(cudaStreamSynchronize(st));
    DSB_CUDA(cudaMemcpy(h.data(), dbg, sizeof(unsigned long long) * 16 * grid, cudaMemcpyDeviceToHost));
    cudaFree(dbg);
    const char* names[12] = {"prod.spin", "prod.issue", "prod.wait_empty", "mma.wait_first", "mma.rest", "epi.gload",
                             "epi.wait_mma", "epi.exchange", "epi.math_store", "epi.y_store", "publisher.release", "epi.gload.lens"};
    const int items = cdiv(n_bgroups, slots) * Tmax;
    fprintf(stderr, "[rnn_ks debug] H=%d B=%d Tmax=%d grid=%d groups=%d slots=%d in flight=%d ring=%dx%d  cycles/item "
                    "(avg over CTAs | max CTA)\n", L.H, B, Tmax, grid, n_bgroups, slots, nif, pl.slots, pl.gsz);
    for (int k = 0; k < 12; ++k) {
      double sum = 0, mx = 0;
      for (int c = 0; c < grid; ++c) { double v = (double)h[c * 16 + k] / items; sum += v; mx = v > mx ? v : mx; }
      fprintf(stderr, "   %-16s %9.0f | %9.0f\n", names[k], sum / grid, mx);
    }
  }
  return 0;
}

}  // namespace dsb
