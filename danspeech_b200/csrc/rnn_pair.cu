// CTA-pair persistent tensor-core recurrence (GRU / LSTM / tanh-RNN) on sm_100a: tcgen05.mma.cta_group::2.
//
// Same contract as rnn_tc.cu (the sequential half of torch.nn.GRU/LSTM/RNN inside BatchRNN.forward,
// danspeech/deepspeech/model.py:114-122, packed-sequence semantics, both directions in one cooperative launch, batch
// groups of 64 sequences, several of them in flight per CTA) and the same per-CTA W_hh slices (pack_whh_tc).  What
// changes is who multiplies what.  Measured on the one-CTA kernel (profiles/r02_recurrence.md): a step of one group
// costs every CTA the whole h_{t-1} of the group through its TMA port (152 KB at H = 1200, ~4.7 k cycles) and 76
// dependent M = 64 MMAs (~3.5 k), for 64 x 64 outputs.  Here two neighbouring CTAs (a cluster) work as one MMA unit:
//   * the pair processes TWO batch groups per item: CTA r streams h_{t-1} of group 2k + r only (A operand, its 64 rows
//     of M = 128);
//   * the B operand is the pair's 128 rows of W_hh: each CTA keeps ITS resident 64-row slice (N half r) and the tensor
//     cores read the peer's half across the pair -- no copy, no exchange of partial sums;
//   * one tcgen05.mma.cta_group::2 (M = 128, N = 128, K = 16), issued by the leader CTA, gives CTA r the accumulator
//     D[64 sequences of group 2k + r][128 gate rows of the pair]: per group and step half the TMA bytes per SM and half
//     the MMA instructions of the one-CTA kernel;
//   * accumulator layout (scripts/tmem_layout_probe.py, cta_group::2 M = 128 N = 128): TMEM lanes 0-63 hold rows 0-63 x
//     columns 0-63 (the leader's W rows), lanes 64-127 rows 0-63 x columns 64-127 (the peer's W rows), so all 32 lanes
//     of all eight epilogue warps own a (sequence, 32 gate columns) strip: no idle lanes, no shuffles.
// The two groups of an item are independent recurrences that only share the MMA issue: each CTA polls / bumps the step
// counter of ITS group's CTA set (the rank-r CTAs of all pairs of the direction).  An odd group count leaves the last
// item's second half empty (rows inactive, its exchange buffer stays zero).
//
// Roles per CTA (352 threads): warp 0 TMA producer (h_{t-1} boxes of ITS group into the shared ring; both CTAs'
// completions are counted on the leader's barriers); warp 1 MMA issuer (leader CTA only; tcgen05.commit multicast
// frees the ring slot / signals the accumulator in both CTAs); warps 2-9 epilogue -- tcgen05.ld, gate math (ex2 / rcp
// on the MUFU pipe, fp32 state in registers), h_t staged as bf16 pairs, bar.arrive, then the fp32 y stores and the
// pre-activations of the group's next step into registers; no epilogue warp ever waits for another one -- and warp 10,
// the publisher: completes the item's named barrier, stores the staged [64][2U] box into the exchange buffer with ONE
// TMA store, waits for its completion and bumps the group's counter with red.release.gpu.
// Layouts (batch_minor, the default inside forward_tc): pre-activations [dirs*G*H][T*B] and outputs [dirs][H][T*B], so
// that a warp (32 consecutive sequences) touches one 128-byte line per column; the backward direction walks
// t = steps-1 .. 0 for ALL sequences of an item (a shorter one joins at t = len-1 with its initial state), which keeps
// that true for ragged batches.  Numbers and the measurements behind each choice: profiles/r02_ncu_pair_recurrence.md.
//
// Every wait is bounded: a stuck barrier sets the abort flag instead of hanging the GPU.
#include "rnn_tc.cuh"
#include <algorithm>
#include <cstdlib>
#include <string>
#include <vector>

namespace dsb {
namespace tc {

constexpr int RP_THREADS = 64 + 256 + 32;   // producer, MMA issuer (leader only), 8 epilogue warps, publisher
constexpr int RP_PUB_WARP = 10;
constexpr int RP_HS = 32;              // accumulator columns per epilogue thread
constexpr int RP_N = 64;               // W_hh rows per CTA = half of the pair's N
constexpr int RP_W_BYTES = RP_N * RT_BK * 2;
constexpr int RP_DBG = 256;            // debug words per CTA: 128 counters + an event trace of steps 100 / 101
// event slot of (item in flight i, step s in {100, 101}, event e < 20)
#define RP_EV(i, s, e) (blockIdx.x * RP_DBG + 128 + ((i) * 2 + ((s) - 100)) * 20 + (e))
#define RP_TRACE(i, s) (p.dbg && ((s) == 100 || (s) == 101) && (i) < 3)

template <int GATES, int NIF>
__global__ void __launch_bounds__(RP_THREADS, 1)
rnn_pair_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_h,
                const __grid_constant__ CUtensorMap tmap_hs, const RnnTcParams p) {
  constexpr int UH = RP_HS / GATES;        // units per epilogue thread
  constexpr int U = 2 * UH;                // units per CTA slice
  constexpr int TMEM_COLS = NIF == 1 ? 64 : (NIF == 2 ? 128 : 256);
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = smem_dyn;
  constexpr int U2 = 2 * U;                // units per pair
  constexpr int STG_BYTES = 64 * U2 * 2;   // h_t of one item: [64 sequences][U2 units] bf16, the box of one TMA store
  const RtPlan pl = rt_plan(p.nkc, 64, NIF * U2, p.ring_gsz, false);
  unsigned char* sW = smem;
  unsigned char* sA = smem + pl.stage_off;
  unsigned char* sStg = smem + pl.stg_off;   // [NIF][64][U2] bf16
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + pl.bar_off);   // [RT_MAX_GROUPS] ring slot landed in BOTH CTAs (leader's copy is used)
  uint64_t* gempty = full + RT_MAX_GROUPS;                           // [RT_MAX_GROUPS] ring slot consumed (multicast commit)
  uint64_t* wbar = gempty + RT_MAX_GROUPS;                           // both W slices resident (leader's copy is used)
  uint64_t* dfull = wbar + 1;                                        // [RT_MAX_NIF] accumulator of item-in-flight i complete (multicast)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dfull + RT_MAX_NIF);
  const int n_groups = pl.groups;
  const int gsz = pl.gsz;
  const int gps = (p.nkc + gsz - 1) / gsz;   // ring slot uses per item

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int set = blockIdx.x / p.cpd;          // (direction, slot)
  const int dir = p.dir0 + set / p.slots;
  const int slot = set % p.slots;
  const int c = blockIdx.x % p.cpd;            // p.cpd is even: c & 1 == rank
  const int half_cpd = p.cpd >> 1;
  const int n_items = (p.n_bgroups + 1) >> 1;  // pair items: groups (2k, 2k + 1)
  const int g_rot = (int)(((long long)(c >> 1) * gps) / half_cpd);   // K-chunk rotation: the same in both CTAs of a pair and as in rnn_tc.cu
  // step counter of (item in flight i, rank): bumped by the half_cpd rank-r CTAs of the set
  unsigned* const ctr0 = p.counters + (size_t)((dir * p.slots + slot) * NIF * 2 + rank) * kRnnCounterStride;
  // the 128B-swizzled tiles need 1 KB alignment (no static shared memory here); both CTAs see the same offset
  const bool misaligned = (smem_u32(smem_dyn) & 1023u) != 0 || (c & 1) != rank;
  if (misaligned && threadIdx.x == 0) atomicExch(p.abort_flag, 1);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_w);
    prefetch_tmap(&tmap_h);
    prefetch_tmap(&tmap_hs);
    for (int i = 0; i < RT_MAX_GROUPS; ++i) mbar_init(&full[i], 1);
    for (int i = 0; i < RT_MAX_GROUPS; ++i) mbar_init(&gempty[i], 1);
    mbar_init(wbar, 1);
    for (int i = 0; i < RT_MAX_NIF; ++i) mbar_init(&dfull[i], 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2cta<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer's barriers exist before anything is signalled on them
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  if (misaligned) goto done;

  if (warp == 0) {
    // ---- TMA producer (both CTAs): whole warp in warp-uniform control flow, one elected lane issues.  Completion of
    //      both CTAs' boxes is counted on the LEADER's barriers (its MMA warp is the only consumer).
    const uint32_t lead_full = mapa_u32(smem_u32(&full[0]), 0);
    const uint32_t lead_wbar = mapa_u32(smem_u32(wbar), 0);
    if (elect_one_sync()) {
      if (rank == 0) mbar_arrive_expect_tx(wbar, 2u * (uint32_t)p.nkc * RP_W_BYTES);
      for (int kc = 0; kc < p.nkc; ++kc)
        tma_load_2d_2cta(sW + (size_t)kc * RP_W_BYTES, &tmap_w, lead_wbar, kc * RT_BK, (dir * p.cpd + c) * RP_N);
    }
    __syncwarp();
    bool ok = true;
    unsigned long long d_spin = 0, d_fence = 0, d_issue = 0, d_empty = 0;
    int cur_slot = 0;
    uint32_t cur_phase = 0;
    auto item = [&](int i, int s, int k, unsigned steps_before) -> bool {
      long long c0 = clock64();
      const int bg = 2 * k + rank;
      if (steps_before + (unsigned)s > 0) {
        // every rank-r CTA of the set has published h_{s-1} of this group (at s = 0 of a later wave: has finished the
        // item that used this accumulator / counter before)
        const unsigned target = (unsigned)half_cpd * (steps_before + (unsigned)s);
        const unsigned* ctr = ctr0 + (size_t)i * 2 * kRnnCounterStride;
        long long t0 = 0;
        unsigned n = 0;
        bool good = true;
        while (ld_acquire_gpu(ctr) < target) {
          if ((++n & 0x3F) == 0) {
            if (*(volatile int*)p.abort_flag) { good = false; break; }
            long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > RT_TIMEOUT_CYCLES) { atomicExch(p.abort_flag, 1); good = false; break; }
          }
        }
        if (!__all_sync(0xffffffffu, good)) return false;
        long long c1 = clock64();
        asm volatile("fence.proxy.async.global;" ::: "memory");   // generic-proxy writes -> async-proxy (TMA) reads
        d_spin += c1 - c0;
        d_fence += clock64() - c1;
      }
      long long c2 = clock64();
      if (p.dbg && s == 100 && i == 0 && lane == 0) p.dbg[blockIdx.x * RP_DBG + 15] = c2;
      if (RP_TRACE(i, s) && lane == 0) { p.dbg[RP_EV(i, s, 0)] = c0; p.dbg[RP_EV(i, s, 1)] = c2; }
      const int row0 = ((bg * 2 + (s & 1)) * p.dirs + dir) * 64;
      for (int g = 0; g < gps; ++g) {
        long long w0 = clock64();
        if (!__all_sync(0xffffffffu, wait_abortable(&gempty[cur_slot], cur_phase ^ 1, p.abort_flag))) return false;
        d_empty += clock64() - w0;
        int gg = g + g_rot;
        if (gg >= gps) gg -= gps;
        if (elect_one_sync()) {
          if (rank == 0) mbar_arrive_expect_tx(&full[cur_slot], 2u * (uint32_t)(gsz * pl.stage_bytes));
          tma_load_3d_2cta(sA + cur_slot * gsz * pl.stage_bytes, &tmap_h, lead_full + (uint32_t)cur_slot * 8u, 0, row0,
                           gg * gsz);
        }
        __syncwarp();
        if (p.dbg && s == 100 && i == 0 && g < 8 && lane == 0) p.dbg[blockIdx.x * RP_DBG + 16 + g] = clock64();
        if (RP_TRACE(i, s) && g < 5 && lane == 0) p.dbg[RP_EV(i, s, 2 + g)] = clock64();
        if (++cur_slot == n_groups) { cur_slot = 0; cur_phase ^= 1; }
      }
      d_issue += clock64() - c2;
      return true;
    };
    unsigned before[NIF];
#pragma unroll
    for (int i = 0; i < NIF; ++i) before[i] = 0;
    for (int k0 = 0; ok && slot + k0 * p.slots < n_items; k0 += NIF) {
      int Tg[NIF], Tw = 0;
#pragma unroll
      for (int i = 0; i < NIF; ++i) {
        const int k = slot + (k0 + i) * p.slots;
        Tg[i] = k < n_items ? rt_group_steps(p, 2 * k) : 0;
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
      p.dbg[blockIdx.x * RP_DBG + 0] = d_spin;
      p.dbg[blockIdx.x * RP_DBG + 1] = d_fence;
      p.dbg[blockIdx.x * RP_DBG + 2] = d_issue;
      p.dbg[blockIdx.x * RP_DBG + 11] = d_empty;
    }
  } else if (warp == 1) {
    // ---- MMA issuer (leader CTA only): one tcgen05.mma.cta_group::2 per K = 16 slice drives both SMs ----
    if (rank == 0) {
      const uint32_t idesc = make_idesc_bf16(128, 128);
      const uint64_t desc0 = make_smem_desc(0, 16, 1024, 2);
      const uint32_t a_lo = smem_u32(sA) >> 4, w_lo = smem_u32(sW) >> 4, stage16 = (uint32_t)pl.stage_bytes >> 4;
      bool ok = __all_sync(0xffffffffu, wait_abortable(wbar, 0, p.abort_flag));
      unsigned long long d_wait0 = 0, d_rest = 0, d_waitn = 0;
      int grp = 0;
      uint32_t fphase = 0;
      auto item = [&](int i, int s) -> bool {
        long long m0 = clock64();
        const uint32_t d_tmem = tmem_base + (uint32_t)(i * 64);
        for (int g = 0; g < gps; ++g) {
          int gg = g + g_rot;
          if (gg >= gps) gg -= gps;
          const int i0 = gg * gsz, i1 = min(p.nkc, i0 + gsz);
          long long w0 = clock64();
          if (!__all_sync(0xffffffffu, wait_abortable(&full[grp], fphase, p.abort_flag))) return false;
          if (g == 0) { long long m1 = clock64(); d_wait0 += m1 - m0; m0 = m1; }
          else d_waitn += clock64() - w0;
          tc_fence_after();
          if (elect_one_sync()) {
            if (p.dbg && s == 100 && i == 0 && g < 8) p.dbg[blockIdx.x * RP_DBG + 40 + g] = clock64();
            if (RP_TRACE(i, s) && g < 5) p.dbg[RP_EV(i, s, 7 + g)] = clock64();
            const uint32_t a0 = a_lo + (uint32_t)(grp * gsz) * stage16;
            const uint32_t b0 = w_lo + (uint32_t)i0 * (RP_W_BYTES >> 4);
            const int nch = i1 - i0;
#pragma unroll
            for (int j = 0; j < RT_GROUP; ++j) {
              if (j < nch) {
                const uint64_t adesc = desc0 + (uint64_t)(a0 + (uint32_t)j * stage16);
                const uint64_t bdesc = desc0 + (uint64_t)(b0 + (uint32_t)j * (RP_W_BYTES >> 4));
#pragma unroll
                for (int kk = 0; kk < RT_BK / 16; ++kk)
                  umma_bf16_2cta(d_tmem, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc,
                                 (j | kk) ? 1u : (uint32_t)(g != 0));
              }
            }
            umma_commit_2cta(&gempty[grp], 3);
            if (g == gps - 1) umma_commit_2cta(&dfull[i], 3);
          }
          __syncwarp();
          if (++grp == n_groups) { grp = 0; fphase ^= 1; }
        }
        if (p.dbg && s == 100 && i == 0 && lane == 0) p.dbg[blockIdx.x * RP_DBG + 64] = clock64();
        if (RP_TRACE(i, s) && lane == 0) p.dbg[RP_EV(i, s, 12)] = clock64();
        d_rest += clock64() - m0;
        return true;
      };
      for (int k0 = 0; ok && slot + k0 * p.slots < n_items; k0 += NIF) {
        int Tg[NIF], Tw = 0;
#pragma unroll
        for (int i = 0; i < NIF; ++i) {
          const int k = slot + (k0 + i) * p.slots;
          Tg[i] = k < n_items ? rt_group_steps(p, 2 * k) : 0;
          Tw = max(Tw, Tg[i]);
        }
        for (int s = 0; s < Tw && ok; ++s) {
#pragma unroll
          for (int i = 0; i < NIF; ++i)
            if (ok && s < Tg[i]) ok = item(i, s);
        }
      }
      if (p.dbg && lane == 0) {
        p.dbg[blockIdx.x * RP_DBG + 3] = d_wait0;
        p.dbg[blockIdx.x * RP_DBG + 4] = d_rest;
        p.dbg[blockIdx.x * RP_DBG + 10] = d_waitn;
      }
    }
  } else if (warp == RP_PUB_WARP) {
    // ---- publisher: completes the item's named barrier (all 256 epilogue threads have issued their h stores) and
    //      bumps the group's step counter with release semantics -- the ~1.4 k cycles a red.release.gpu waits for the
    //      CTA's outstanding stores are spent here, not in an epilogue warp
    unsigned long long d_pub = 0;
    bool go = true;
    for (int k0 = 0; go && slot + k0 * p.slots < n_items; k0 += NIF) {
      int Tg[NIF], Tw = 0;
#pragma unroll
      for (int i = 0; i < NIF; ++i) {
        const int k = slot + (k0 + i) * p.slots;
        Tg[i] = k < n_items ? rt_group_steps(p, 2 * k) : 0;
        Tw = max(Tw, Tg[i]);
      }
      for (int s = 0; s < Tw && go; ++s) {
#pragma unroll
        for (int i = 0; i < NIF; ++i) {
          if (go && s < Tg[i]) {
            named_bar_sync(2 + i, 256 + 32);
            long long q0 = clock64();
            if (elect_one_sync()) {
              // h_t of the item: ONE TMA store of the staged [64][U2] box into the group's exchange buffer (rows the
              // sequences of which have ended carry their last state: every row only feeds its own accumulator row)
              const int bg = 2 * (slot + (k0 + i) * p.slots) + rank;
              tma_store_2d(&tmap_hs, sStg + i * STG_BYTES, (c & ~1) * U, ((bg * 2 + ((s + 1) & 1)) * p.dirs + dir) * 64);
              bulk_commit_group();
              const long long q2 = clock64();
              bulk_wait_group0();
              const long long q3 = clock64();
              // The bulk store has COMPLETED here (its writes are visible to this thread); the releasing red makes them
              // visible to whoever acquires the counter.  A relaxed red is ~7 % faster per step -- no MEMBAR.GPU waiting
              // for the SM's other outstanding accesses -- and WRONG: measured, readers on other SMs then now and again
              // stream a stale row of h_t (results differ from run to run).
              red_release_gpu_add(ctr0 + (size_t)i * 2 * kRnnCounterStride, 1u);   // publish h_t of this group
              if (p.dbg && i == 0 && s == 100) {
                p.dbg[blockIdx.x * RP_DBG + 120] = q0; p.dbg[blockIdx.x * RP_DBG + 121] = q2; p.dbg[blockIdx.x * RP_DBG + 122] = q3;
              }
            }
            __syncwarp();
            if (*(volatile int*)p.abort_flag) { go = false; break; }   // (after the publish: off the step's critical path)
            const long long q1 = clock64();
            d_pub += q1 - q0;
            if (RP_TRACE(i, s) && lane == 0) { p.dbg[RP_EV(i, s, 15)] = q0; p.dbg[RP_EV(i, s, 16)] = q1; }
            if (p.dbg && i == 0 && s == 99 && lane == 0) p.dbg[blockIdx.x * RP_DBG + 66] = q1;
            if (p.dbg && i == 0 && s == 100 && lane == 0) p.dbg[blockIdx.x * RP_DBG + 67] = q1;
          }
        }
      }
    }
    if (p.dbg && lane == 0) p.dbg[blockIdx.x * RP_DBG + 12] = d_pub;
  } else {
    // ---- epilogue: 8 warps.  TMEM lane quarter q = warp % 4: sequences (q & 1) * 32 + lane of this CTA's group, gate
    //      columns of CTA (q >> 1) of the pair; the two warps of a quarter split those 64 columns.
    const int et = threadIdx.x - 64;     // 0..255
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int lrow = (q & 1) * 32 + lane;        // row inside the batch group (= hbuf row)
    const int csel = q >> 1;
    const int cpair = c & ~1;                    // first CTA of the pair inside the set
    const int j0 = (cpair + csel) * U + half * UH;
    const int ncol = p.dirs * GATES * p.H;
    float hprev[NIF][UH], cst[NIF][UH], gxr[NIF][GATES][UH], bhn[UH];
#pragma unroll
    for (int u = 0; u < UH; ++u)
      bhn[u] = (GATES == 3 && p.b_hn && j0 + u < p.H) ? p.b_hn[(size_t)dir * p.H + j0 + u] : 0.f;
    const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * RP_HS);
    const bool vec2 = ((p.H & 1) == 0) && ((UH & 1) == 0) && j0 + UH <= p.H;
    const bool full_units = j0 + UH <= p.H;
    const size_t gate_stride = (size_t)p.H * (size_t)p.ldt;
    unsigned long long e_load = 0, e_wait = 0, e_math = 0, e_bar = 0, e_pub = 0;

    // pre-activations of step s of sequence b (row t of gx) into registers; they are asked for one step of the group
    // ahead, right after the group's previous step was handed to the publisher, so their latency hides behind the
    // other items
    auto load_gx = [&](float (&gxv)[GATES][UH], int t, int b) {
      if (p.skip & 1) return;
      if (p.ldt > 0) {
        // batch-minor pre-activations (gemm_bias_rows_tc): the 32 lanes of a warp read 32 consecutive floats per column
        const float* gp = p.gx + ((size_t)dir * GATES * p.H + j0) * p.ldt + (size_t)t * p.B + b;
        if (full_units) {   // one pointer per gate, stepped by a column
#pragma unroll
          for (int g = 0; g < GATES; ++g) {
            const float* q = gp + (size_t)g * gate_stride;
#pragma unroll
            for (int u = 0; u < UH; ++u) { gxv[g][u] = __ldg(q); q += p.ldt; }
          }
        } else {
#pragma unroll
          for (int g = 0; g < GATES; ++g)
#pragma unroll
            for (int u = 0; u < UH; ++u)
              gxv[g][u] = (j0 + u < p.H) ? __ldg(gp + ((size_t)g * p.H + u) * p.ldt) : 0.f;
        }
        return;
      }
      const float* gp = p.gx + ((size_t)t * p.B + b) * ncol + (size_t)dir * GATES * p.H + j0;
      if (vec2) {
#pragma unroll
        for (int g = 0; g < GATES; ++g)
#pragma unroll
          for (int u = 0; u < UH; u += 2) {
            const float2 v = __ldg(reinterpret_cast<const float2*>(gp + (size_t)g * p.H + u));
            gxv[g][u] = v.x;
            gxv[g][u + 1 < UH ? u + 1 : u] = v.y;
          }
      } else {
#pragma unroll
        for (int g = 0; g < GATES; ++g)
#pragma unroll
          for (int u = 0; u < UH; ++u) gxv[g][u] = (j0 + u < p.H) ? __ldg(gp + (size_t)g * p.H + u) : 0.f;
      }
    };

    // one step of one item.  hp / cs / gxv: this group's recurrent state and pre-activations (registers).
    // No thread of the epilogue waits for another one: h_t goes from registers to the exchange buffer (bf16 pairs), the
    // thread ARRIVES on the item's named barrier and carries on with its y stores and the next pre-activations; the
    // publisher warp completes the barrier and pays for the release.
    auto item = [&](int i, int s, int steps, int bg, int b, bool row_ok, int len, unsigned steps_before,
                    float (&hp)[UH], float (&cs)[UH], float (&gxv)[GATES][UH]) -> bool {
      long long e1 = clock64();
      // Every sequence of the item is at the SAME time index in a step (the batch-minor loads / stores of a warp are then
      // one line each however ragged the batch is): forwards t = s; backwards the item walks t = steps-1 .. 0 and a
      // shorter sequence JOINS at t = len-1 with its initial state -- the same arithmetic per sequence as starting
      // them together (pack_padded_sequence semantics, model.py:117-120), just at a later step.
      const int t = dir == 0 ? s : steps - 1 - s;
      const bool active = row_ok && t < len;
      const bool ok = wait_abortable(&dfull[i], (uint32_t)((steps_before + (unsigned)s) & 1u), p.abort_flag);
      long long e2 = clock64();
      if (p.dbg && s == 100 && i == 0 && et == 64) p.dbg[blockIdx.x * RP_DBG + 65] = e2;
      if (RP_TRACE(i, s) && et == 64) { p.dbg[RP_EV(i, s, 13)] = e2; p.dbg[RP_EV(i, s, 17)] = e1; }
      tc_fence_after();
      uint32_t r[32];
      tmem_ld32(t_addr + (uint32_t)(i * 64), r);
      tmem_ld_wait();
      tc_fence_before();
      if (ok && active && !(p.skip & 4)) {
#pragma unroll
        for (int u = 0; u < UH; ++u) {
          float hn;
          if (GATES == 3) {
            const float rg = fast_sigmoid(gxv[0][u] + __uint_as_float(r[u * GATES + 0]));
            const float zg = fast_sigmoid(gxv[1 % GATES][u] + __uint_as_float(r[u * GATES + (1 % GATES)]));
            const float ng = fast_tanh(gxv[2 % GATES][u] + rg * (__uint_as_float(r[u * GATES + (2 % GATES)]) + bhn[u]));
            hn = (1.0f - zg) * ng + zg * hp[u];
          } else if (GATES == 4) {
            const float ig = fast_sigmoid(gxv[0][u] + __uint_as_float(r[u * GATES + 0]));
            const float fg = fast_sigmoid(gxv[1 % GATES][u] + __uint_as_float(r[u * GATES + (1 % GATES)]));
            const float gg = fast_tanh(gxv[2 % GATES][u] + __uint_as_float(r[u * GATES + (2 % GATES)]));
            const float og = fast_sigmoid(gxv[3 % GATES][u] + __uint_as_float(r[u * GATES + (3 % GATES)]));
            cs[u] = fg * cs[u] + ig * gg;
            hn = og * fast_tanh(cs[u]);
          } else {
            hn = fast_tanh(gxv[0][u] + __uint_as_float(r[u]));
          }
          hp[u] = hn;
        }
      }
      long long e3 = clock64();
      {
        // h_t (bf16 pairs) -> the item's staging box; unconditionally: ended / missing sequences restage their last state
        uint32_t* sh = reinterpret_cast<uint32_t*>(sStg + i * STG_BYTES) + ((lrow * U2 + csel * U + half * UH) >> 1);
#pragma unroll
        for (int u = 0; u < UH; u += 2) {
          const __nv_bfloat162 v = __floats2bfloat162_rn(hp[u], hp[u + 1 < UH ? u + 1 : u]);
          sh[u >> 1] = *reinterpret_cast<const uint32_t*>(&v);
        }
        fence_proxy_async();   // generic-proxy writes -> the publisher's TMA store (async proxy)
      }
      long long e4 = clock64();
      asm volatile("bar.arrive %0, %1;" ::"r"(2 + i), "r"(256 + 32) : "memory");   // -> publisher (also on abort)
      if (RP_TRACE(i, s) && et == 64) p.dbg[RP_EV(i, s, 14)] = e4;
      if (p.dbg && s == 100 && i == 0 && lane == 0) {
        unsigned long long* d = p.dbg + blockIdx.x * RP_DBG + 72 + (warp - 2) * 6;
        d[0] = e1; d[1] = e2; d[2] = e3; d[3] = e4; d[4] = 0; d[5] = 0;
      }
      if (!ok) return false;
      // pre-activations of the group's next step: nobody waits on them before the group's next item
      {
        const int tn = dir == 0 ? s + 1 : steps - 2 - s;
        if (row_ok && s + 1 < steps && tn < len) load_gx(gxv, tn, b);
      }
      long long e5 = clock64();
      // y_t (fp32) straight from registers: nobody waits on these stores
      if (ok && active && !(p.skip & 2)) {
        if (p.ldt > 0) {
          float* yo = p.y + ((size_t)dir * p.H + j0) * p.ldt + (size_t)t * p.B + b;   // batch-minor: coalesced per unit
          if (full_units) {
#pragma unroll
            for (int u = 0; u < UH; ++u) { *yo = hp[u]; yo += p.ldt; }
          } else {
#pragma unroll
            for (int u = 0; u < UH; ++u)
              if (j0 + u < p.H) yo[(size_t)u * p.ldt] = hp[u];
          }
        } else {
          float* yo = p.y + (((size_t)dir * p.T + t) * p.B + b) * p.H + j0;
          if (vec2) {
#pragma unroll
            for (int u = 0; u < UH; u += 2)
              *reinterpret_cast<float2*>(yo + u) = make_float2(hp[u], hp[u + 1 < UH ? u + 1 : u]);
          } else {
#pragma unroll
            for (int u = 0; u < UH; ++u)
              if (j0 + u < p.H) yo[u] = hp[u];
          }
        }
      }
      e_pub += clock64() - e5;
      e_load += e5 - e4; e_wait += e2 - e1; e_math += e3 - e2; e_bar += e4 - e3;
      return true;
    };

    unsigned before[NIF];
#pragma unroll
    for (int i = 0; i < NIF; ++i) before[i] = 0;
    bool alive = true;
    for (int k0 = 0; alive && slot + k0 * p.slots < n_items; k0 += NIF) {
      int Tg[NIF], Tw = 0, bb[NIF], ln[NIF], bgs[NIF];
      bool rok[NIF];
#pragma unroll
      for (int i = 0; i < NIF; ++i) {
        const int k = slot + (k0 + i) * p.slots;
        const int bg = 2 * k + rank;
        bgs[i] = bg;
        Tg[i] = k < n_items ? rt_group_steps(p, 2 * k) : 0;
        Tw = max(Tw, Tg[i]);
        bb[i] = bg * 64 + lrow;
        rok[i] = k < n_items && bg < p.n_bgroups && bb[i] < p.B;
        ln[i] = rok[i] ? (p.lens ? p.lens[bb[i]] : p.Tmax) : 0;
#pragma unroll
        for (int u = 0; u < UH; ++u) {
          const bool in = rok[i] && j0 + u < p.H;
          hprev[i][u] = (p.h0 && in) ? p.h0[((size_t)dir * p.B + bb[i]) * p.H + j0 + u] : 0.f;
          cst[i][u] = (GATES == 4 && p.c0 && in) ? p.c0[((size_t)dir * p.B + bb[i]) * p.H + j0 + u] : 0.f;
#pragma unroll
          for (int g = 0; g < GATES; ++g) gxr[i][g][u] = 0.f;
        }
        if (rok[i] && Tg[i] > 0 && (dir == 0 ? 0 : Tg[i] - 1) < ln[i]) load_gx(gxr[i], dir == 0 ? 0 : Tg[i] - 1, bb[i]);
      }
      for (int s = 0; s < Tw && alive; ++s) {
#pragma unroll
        for (int i = 0; i < NIF; ++i)
          if (alive && s < Tg[i])
            alive = item(i, s, Tg[i], bgs[i], bb[i], rok[i], ln[i], before[i], hprev[i], cst[i], gxr[i]);
      }
#pragma unroll
      for (int i = 0; i < NIF; ++i) {
        before[i] += (unsigned)Tg[i];
        if (alive && rok[i]) {   // carry the state out (streaming): the last active step's h (and c)
#pragma unroll
          for (int u = 0; u < UH; ++u)
            if (j0 + u < p.H) {
              if (p.hT) p.hT[((size_t)dir * p.B + bb[i]) * p.H + j0 + u] = hprev[i][u];
              if (GATES == 4 && p.cT) p.cT[((size_t)dir * p.B + bb[i]) * p.H + j0 + u] = cst[i][u];
            }
        }
      }
    }
    if (p.dbg && et == 64) {   // warp 4: quarter 0, an active row
      p.dbg[blockIdx.x * RP_DBG + 5] = e_load;
      p.dbg[blockIdx.x * RP_DBG + 6] = e_wait;
      p.dbg[blockIdx.x * RP_DBG + 7] = e_math;
      p.dbg[blockIdx.x * RP_DBG + 8] = e_bar;
      p.dbg[blockIdx.x * RP_DBG + 9] = e_pub;
    }
  }
done:
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer's shared memory / TMEM stay valid until both CTAs are done
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta<TMEM_COLS>(tmem_base);
  }
}

}  // namespace tc

// The pair kernel takes the layers the one-CTA kernel takes with 64-row slices, when a direction's CTA count is even
// and there are at least two batch groups to pair.
bool rnn_pair_supported(const RnnLayer& L, int B, int sms) {
  if (!g_tune.rnn_pair.load() || rnn_tc_max_in_flight() < 2 || B < g_tune.rnn_pair_min_rows.load()) return false;
  if (rnn_tc_narrow(L, B)) return false;
  const int U = tc::rt_units(L.gates, false), cpd = cdiv(L.H, U), HP = (L.H + 63) / 64 * 64;
  if (cpd & 1) return false;
  const int nif_max = g_tune.rnn_pair_in_flight.load();
  const tc::RtPlan pl = tc::rt_plan(HP / 64, 64, nif_max * 2 * U, g_tune.rnn_ring_gsz.load(), false);
  return pl.groups >= 1 && pl.gsz >= 1 && pl.total <= tc::RT_SMEM_LIMIT && cpd <= sms;
}

// One BatchRNN layer on CTA pairs; arguments as rnn_layer_tc (hbuf: rnn_tc_hbuf_elems(), an even number of groups).
int rnn_layer_pair(const RnnLayer& L, const float* gx, const int32_t* d_len, int B, int T, int Tmax, float* y,
                   __nv_bfloat16* hbuf, unsigned int* sync_words, int* abort_flag, cudaStream_t st, const float* h0,
                   const float* c0, float* hT, float* cT, bool batch_minor) {
  using namespace tc;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (!rnn_pair_supported(L, B, sms))
    return set_error(DSB_ERR_UNSUPPORTED, "rnn_layer_pair: shape H=%d B=%d not supported", L.H, B);
  const int U = rt_units(L.gates, false), cpd = cdiv(L.H, U);
  const int HP = (L.H + 63) / 64 * 64, nkc = HP / 64;
  const int n_bgroups = cdiv(B, 64), n_items = (n_bgroups + 1) / 2, n_alloc = 2 * n_items;
  const int launches = (L.dirs * cpd <= sms) ? 1 : L.dirs;
  const int dirs_per_launch = L.dirs / launches;
  int slots = sms / (dirs_per_launch * cpd);
  if (slots > n_items) slots = n_items;
  if (g_tune.rnn_max_slots.load() > 0 && slots > g_tune.rnn_max_slots.load()) slots = g_tune.rnn_max_slots.load();
  if (slots < 1) slots = 1;
  int nif = cdiv(n_items, slots);            // items per CTA set; up to RT_MAX_NIF of them in flight
  const int nif_max = g_tune.rnn_pair_in_flight.load();
  if (nif > nif_max) nif = nif_max;
  while (slots > 1 && L.dirs * slots * nif * 2 > kRnnMaxCounters) --slots;
  DSB_CUDA(cudaMemsetAsync(hbuf, 0, sizeof(__nv_bfloat16) * rnn_tc_hbuf_elems(L, B), st));
  DSB_CUDA(cudaMemsetAsync(sync_words, 0, sizeof(unsigned int) * kRnnSyncCounters, st));
  if (h0)
    if (int e = rnn_tc_init_hbuf(h0, hbuf, L.dirs, B, L.H, HP, 64, n_bgroups, st)) return e;

  const int ring_gsz = g_tune.rnn_ring_gsz.load();
  const RtPlan pl = rt_plan(nkc, 64, nif * 2 * U, ring_gsz, false);
  CUtensorMap tw, th, ths;
  uint64_t dw[2] = {(uint64_t)HP, (uint64_t)L.dirs * cpd * RP_N}, sw[2] = {2, (uint64_t)HP * 2};
  uint32_t bw[2] = {RT_BK, (uint32_t)RP_N};
  if (int e = make_tmap_bf16(&tw, L.w_hh_pack, 2, dw, sw, bw, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
  uint64_t dh[3] = {(uint64_t)RT_BK, (uint64_t)n_alloc * 2 * L.dirs * 64, (uint64_t)nkc};
  uint64_t sh[3] = {2, (uint64_t)HP * 2, (uint64_t)RT_BK * 2};
  uint32_t bh[3] = {RT_BK, 64u, (uint32_t)pl.gsz};
  if (int e = make_tmap_bf16(&th, hbuf, 3, dh, sh, bh, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
  // the same buffer as a plain [rows][HP] matrix for the publisher's store of one item's h_t: box {2U units, 64 rows}
  uint64_t ds[2] = {(uint64_t)HP, (uint64_t)n_alloc * 2 * L.dirs * 64}, ss[2] = {2, (uint64_t)HP * 2};
  uint32_t bs[2] = {(uint32_t)(2 * U), 64u};
  if (int e = make_tmap_bf16(&ths, hbuf, 2, ds, ss, bs, CU_TENSOR_MAP_SWIZZLE_NONE)) return e;

  RnnTcParams p{};
  p.gx = gx; p.b_hn = L.b_hn; p.y = y; p.hbuf = hbuf; p.lens = d_len;
  p.counters = sync_words; p.abort_flag = abort_flag;
  p.h0 = h0; p.c0 = c0; p.hT = hT; p.cT = cT;
  p.n_bgroups = n_bgroups; p.slots = slots;
  p.B = B; p.H = L.H; p.HP = HP; p.BP = 64; p.T = T; p.Tmax = Tmax;
  p.dirs = L.dirs; p.cpd = cpd; p.U = U; p.nkc = nkc;
  p.ring_gsz = ring_gsz;
  p.n_producers = 1;
  static const int skip_env = getenv("DSB_RNN_SKIP") ? atoi(getenv("DSB_RNN_SKIP")) : 0;
  p.skip = skip_env;
  p.ldt = batch_minor ? (long long)T * B : 0;
  const void* fn = nullptr;
#define RP_PICK(G) \
  fn = nif == 1 ? (const void*)rnn_pair_kernel<G, 1> : nif == 2 ? (const void*)rnn_pair_kernel<G, 2> : (const void*)rnn_pair_kernel<G, 3>
  if (L.gates == 3) RP_PICK(3);
  else if (L.gates == 4) RP_PICK(4);
  else RP_PICK(1);
#undef RP_PICK
  DSB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, pl.total));
  static const bool debug = getenv("DSB_RNN_DEBUG") != nullptr;
  const int grid = dirs_per_launch * slots * cpd;
  unsigned long long* dbg = nullptr;
  if (debug) {
    DSB_CUDA(cudaMalloc(&dbg, sizeof(unsigned long long) * RP_DBG * grid));
    DSB_CUDA(cudaMemsetAsync(dbg, 0, sizeof(unsigned long long) * RP_DBG * grid, st));
  }
  p.dbg = dbg;
  for (int l = 0; l < launches; ++l) {
    p.dir0 = l * dirs_per_launch;
    void* args[] = {(void*)&tw, (void*)&th, (void*)&ths, (void*)&p};
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(RP_THREADS);
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
    // Profilers cannot replay a cooperative cluster launch (ncu: LaunchFailed).  DSB_RNN_NONCOOP=1 drops the cooperative
    // attribute for a profiling run on an otherwise idle GPU, where the <= 148 single-CTA-per-SM blocks are co-resident
    // anyway; production keeps it (it is what makes spinning on the other CTAs safe next to other work).
    static const bool noncoop = getenv("DSB_RNN_NONCOOP") != nullptr;
    if (noncoop) {
      attrs[0] = attrs[1];
      cfg.numAttrs = 1;
    }
    const cudaError_t le = cudaLaunchKernelExC(&cfg, fn, args);
    if (le != cudaSuccess)
      return set_error(DSB_ERR_CUDA, "rnn_layer_pair: launch failed: %s", cudaGetErrorString(le));
    count_launch();
  }
  if (debug) {
    std::vector<unsigned long long> h((size_t)RP_DBG * grid);
    DSB_CUDA(cudaStreamSynchronize(st));
    DSB_CUDA(cudaMemcpy(h.data(), dbg, sizeof(unsigned long long) * RP_DBG * grid, cudaMemcpyDeviceToHost));
    cudaFree(dbg);
    const char* names[13] = {"prod.spin", "prod.fence", "prod.issue", "mma.wait_first", "mma.rest", "epi.gx_loads",
                             "epi.wait_mma", "epi.tmem+math", "epi.stage", "epi.y_store", "mma.wait_rest",
                             "prod.wait_empty", "publisher.release"};
    const int items = cdiv(n_items, slots) * Tmax;   // (step, pair item) items per CTA (upper bound for ragged groups)
    fprintf(stderr, "[rnn_pair debug] H=%d B=%d Tmax=%d grid=%d groups=%d pair items=%d slots=%d in flight=%d ring=%dx%d  "
                    "cycles/item (avg over CTAs | max CTA)\n", L.H, B, Tmax, grid, n_bgroups, n_items, slots, nif, pl.groups, pl.gsz);
    for (int k = 0; k < 13; ++k) {
      double sum = 0, mx = 0;
      for (int cc = 0; cc < grid; ++cc) { double v = (double)h[(size_t)cc * RP_DBG + k] / items; sum += v; mx = v > mx ? v : mx; }
      fprintf(stderr, "   %-16s %9.0f | %9.0f\n", names[k], sum / grid, mx);
    }
    for (int cc = 0; cc < grid; cc += grid / 2 + 1) {   // step-100 timeline of two CTAs (cycles since this CTA published step 99)
      const unsigned long long* d = &h[(size_t)cc * RP_DBG];
      const long long t0 = (long long)d[66];
      fprintf(stderr, "   [cta %d] barrier passed %+lld | tma group issued:", cc, (long long)d[15] - t0);
      for (int i = 0; i < (nkc + pl.gsz - 1) / pl.gsz && i < 8; ++i) fprintf(stderr, " %lld", (long long)d[16 + i] - t0);
      fprintf(stderr, "\n   [cta %d] group ready:", cc);
      for (int i = 0; i < (nkc + pl.gsz - 1) / pl.gsz && i < 8; ++i) fprintf(stderr, " %lld", (long long)d[40 + i] - t0);
      fprintf(stderr, "\n   [cta %d] mma issued %+lld | epilogue saw dfull %+lld | published %+lld\n", cc,
              (long long)d[64] - t0, (long long)d[65] - t0, (long long)d[67] - t0);
      fprintf(stderr, "   [cta %d] publisher: barrier complete %+lld | store issued %+lld | store complete %+lld\n", cc,
              (long long)d[120] - t0, (long long)d[121] - t0, (long long)d[122] - t0);
      {   // event trace of steps 100 and 101, all items in flight, in time order
        static const char* ev[18] = {"producer: at the barrier", "producer: barrier passed", "box 1 issued", "box 2 issued", "box 3 issued",
                                     "box 4 issued", "box 5 issued", "mma: box 1 landed", "mma: box 2 landed", "mma: box 3 landed",
                                     "mma: box 4 landed", "mma: box 5 landed", "mma: all issued", "epilogue: accumulator complete",
                                     "epilogue: staged, arrived", "publisher: barrier complete (issue time)", "publisher: published",
                                     "epilogue: starts waiting"};
        std::vector<std::pair<long long, std::string>> tr;
        for (int i = 0; i < nif; ++i)
          for (int ss = 0; ss < 2; ++ss)
            for (int e = 0; e < 18; ++e) {
              const unsigned long long v = d[128 + (i * 2 + ss) * 20 + e];
              if (v) tr.emplace_back((long long)v - t0, "item " + std::to_string(i) + " step " + std::to_string(100 + ss) + "  " + ev[e]);
            }
        std::sort(tr.begin(), tr.end());
        for (auto& x : tr) fprintf(stderr, "   [cta %d] %+8lld  %s\n", cc, x.first, x.second.c_str());
      }
      for (int w = 0; w < 8; w += 3) {
        const unsigned long long* e = d + 72 + w * 6;
        fprintf(stderr, "   [cta %d] epilogue warp %d: wait from %+lld | dfull %+lld | math done %+lld | h stored, arrived %+lld\n",
                cc, w + 2, (long long)e[0] - t0, (long long)e[1] - t0, (long long)e[2] - t0, (long long)e[3] - t0);
      }
    }
  }
  return 0;
}

}  // namespace dsb
