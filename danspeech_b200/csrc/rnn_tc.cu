// Persistent tensor-core recurrence for BatchRNN (GRU / LSTM / tanh-RNN) on sm_100a.
//
// Replaces the sequential half of torch.nn.GRU/LSTM/RNN inside BatchRNN.forward
// (danspeech/deepspeech/model.py:114-122) with packed-sequence semantics: sequence b runs
// t = 0..len_b-1 forwards and len_b-1..0 backwards from its own end; both directions run
// concurrently in one cooperative launch.
//
// Decomposition: a CTA owns U hidden units of ONE direction (all gates: N = 64 rows of W_hh) for
// the whole layer.  Its W_hh slice (bf16, K padded to a multiple of 64) is loaded ONCE by TMA into
// 128B-swizzled shared memory and stays resident for all T steps.  Per step:
//   producer warp : arms the ring, polls the step barrier of its CTA set (ld.acquire on a counter every CTA of
//                   the set bumps with red.release), then TMA-streams h_{t-1} (bf16, [rows, H]) into the ring in
//                   3-D boxes of four K-chunks of 64 (two such groups in flight),
//   MMA warp      : tcgen05.mma  D[batch rows (M = 64 or 128), 64 gate columns] += h_chunk * W_chunk^T, fp32 in
//                   TMEM, 16 MMAs per elected issue region,
//   8 epilogue warps: prefetch the input-projection pre-activations of the step while the MMAs run,
//                   tcgen05.ld the accumulators, apply the gate non-linearities with the fp32 hidden
//                   state kept in registers, stage h_t (bf16) in shared memory, store it coalesced for the next
//                   step's TMA, publish (one red.release per CTA per step), then store y_t (fp32).
// Batches above 64 rows run as groups of 64 inside the same launch (W_hh stays resident), up to three groups IN
// FLIGHT per CTA (see rnn_tc_kernel); groups also run side by side on independent CTA sets ("slots") when a
// direction's CTAs leave SMs free, and an initial / final hidden state can be carried (streaming).
// The step is latency-bound (grid barrier + L2 round trips), not tensor-bound: algorithmic work is
// 2*B*3H*H flop per step per direction and is reported against the tensor roof.
//
// Every wait is bounded: a stuck barrier sets an abort flag instead of hanging the GPU.
#include "rnn_tc.cuh"
#include <cstdlib>

namespace dsb {
namespace tc {


// NIF batch groups of a CTA set are in flight at a time.  One step of one group is a serial chain -- all-gather of
// h (publish -> barrier -> TMA: ~3.4 k cycles of latency) -> 76 MMAs (~4.5 k) -> gate math and publish (~2.7 k) --
// in which the tensor pipe is busy less than half of the time and nothing can overlap, because every chunk of h
// becomes available at the same moment.  Independent groups can: all three roles walk the same item sequence
// (wave, step, group-in-flight); while group A sits in its epilogue and communication chain, the MMA warp runs
// group B against the same resident W_hh slice.  Per group in flight: its own TMEM accumulator, accumulator-full
// barrier, step counter and exchange buffers; the TMA ring is shared (the MMA phases are serialised on the tensor
// pipe anyway) and the eight epilogue warps take the groups in turn.
//
// SPLITM (GRU, batch groups of 64 rows): an M = 64 accumulator quarter only fills TMEM lanes 0-15, so lanes 16-31 of
// every epilogue warp would idle through the MUFU-bound gate math.  Lane l+16 computes the second half of lane l's
// hidden units on values handed over by shuffle and hands h back; loads, stores and the recurrent state stay with
// lanes 0-15 (splitting those as well doubled the number of memory requests and was measured slower).
template <int GATES, bool SPLITM, int NIF, bool NARROW>
__global__ void __launch_bounds__(RT_THREADS, 1)
rnn_tc_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_h,
              const RnnTcParams p) {
  constexpr int HS = rt_hs(NARROW);        // accumulator columns per half
  constexpr int RT_N = 2 * HS;             // W_hh rows of this CTA
  constexpr int RT_W_BYTES = RT_N * RT_BK * 2;
  constexpr int UH = HS / GATES;           // units per half
  constexpr int U = 2 * UH;
  constexpr int TMEM_COLS = NIF == 1 ? 64 : (NIF == 2 ? 128 : 256);
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = smem_dyn;
  if ((smem_u32(smem_dyn) & 1023u) != 0) {   // the 128B-swizzled tiles need 1 KB alignment (no static shared memory here)
    if (threadIdx.x == 0) atomicExch(p.abort_flag, 1);
    return;
  }
  const RtPlan pl = rt_plan(p.nkc, p.BP, U, p.ring_gsz, NARROW);
  unsigned char* sW = smem;
  unsigned char* sA = smem + pl.stage_off;
  unsigned char* sStg = smem + pl.stg_off;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + pl.bar_off);   // [RT_MAX_GROUPS] group landed (TMA tx)
  uint64_t* gempty = full + RT_MAX_GROUPS;                           // [RT_MAX_GROUPS] group consumed by the MMAs
  uint64_t* wbar = gempty + RT_MAX_GROUPS;
  uint64_t* dfull = wbar + 1;                                        // [RT_MAX_NIF] accumulator of group-in-flight i complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dfull + RT_MAX_NIF);
  const int n_groups = pl.groups;
  const int gsz = pl.gsz;
  const int gps = (p.nkc + gsz - 1) / gsz;   // ring-group uses per step

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int set = blockIdx.x / p.cpd;          // (direction, slot)
  const int dir = p.dir0 + set / p.slots;
  const int slot = set % p.slots;
  const int c = blockIdx.x % p.cpd;
  // every CTA walks the K chunks in its own rotation so that the readers do not hit the same L2 lines together
  // (neighbouring CTAs share a rotation: the CTA-pair kernel, rnn_pair.cu, then accumulates in the same order)
  const int g_rot = (int)(((long long)(c >> 1) * gps) / ((p.cpd + 1) >> 1));   // rotation in whole ring groups
  unsigned* const ctr0 = p.counters + (size_t)((dir * p.slots + slot) * NIF) * kRnnCounterStride;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_w);
    prefetch_tmap(&tmap_h);
    for (int i = 0; i < RT_MAX_GROUPS; ++i) mbar_init(&full[i], 1);
    for (int i = 0; i < RT_MAX_GROUPS; ++i) mbar_init(&gempty[i], 1);
    mbar_init(wbar, 1);
    for (int i = 0; i < RT_MAX_NIF; ++i) mbar_init(&dfull[i], 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0 || warp == 10) {
    // ---- TMA producers: whole warp in warp-uniform control flow, one elected lane issues (elect_one_sync).
    // Measured (scripts/tma_microbench.py, one CTA): ONE warp gets a TMA box into flight every ~700 cycles whatever
    // its size (8 KB or 64 KB, tensor box or bulk copy, 1 or 8 boxes in flight), two warps together every ~500.  A
    // second producer warp on alternate ring slots (dsb_tune_set("rnn_producers", 2)) did NOT shorten the h stream
    // of a step inside this kernel, though: with all CTAs of a direction pulling the same h the boxes land ~940
    // cycles apart whoever issues them.  Default: one producer (warp 0).
    const int pi = warp == 0 ? 0 : 1;
    const int NP = p.n_producers;
    if (pi >= NP) goto done;
    if (pi == 0 && elect_one_sync()) {
      // resident W_hh slice, loaded once
      mbar_arrive_expect_tx(wbar, (uint32_t)p.nkc * RT_W_BYTES);
      for (int kc = 0; kc < p.nkc; ++kc)
        tma_load_2d(sW + (size_t)kc * RT_W_BYTES, &tmap_w, wbar, kc * RT_BK, (dir * p.cpd + c) * RT_N);
    }
    __syncwarp();
    bool ok = true;
    unsigned long long d_spin = 0, d_fence = 0, d_issue = 0, d_empty = 0;
    // ring position of the next slot use (slot index + phase kept incrementally: a 64-bit modulo per group costs the
    // single issuing warp hundreds of cycles); `turn` says whose slot use it is
    int cur_slot = 0, turn = 0;
    uint32_t cur_phase = 0;
    // One TMA per ring slot: the h buffer is mapped as a 3-D tensor {64 k, BP rows, nkc chunks}, so a box
    // {64, BP, gsz} lands as gsz consecutive 128B-swizzled K-chunk tiles (chunks beyond nkc are zero-filled and
    // never used by the MMAs).
    auto item = [&](int i, int s, int bg, unsigned steps_before) -> bool {
      long long c0 = clock64();
      if (steps_before + (unsigned)s > 0) {
        // set-wide barrier: every CTA of this set has published h_{s-1} of this group (at s = 0 of a later wave:
        // has finished the group that used this accumulator / counter before).  With several groups in flight it is
        // usually open already (the group published while the others ran).
        const unsigned target = (unsigned)p.cpd * (steps_before + (unsigned)s);
        const unsigned* ctr = ctr0 + i * kRnnCounterStride;
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
      if (p.dbg && s == 100 && i == 0 && lane == 0 && pi == 0) p.dbg[blockIdx.x * 128 + 15] = c2;
      const int row0 = ((bg * 2 + (s & 1)) * p.dirs + dir) * p.BP;
      for (int g = 0; g < gps; ++g) {
        if (turn == pi) {
          long long w0 = clock64();
          if (!__all_sync(0xffffffffu, wait_abortable(&gempty[cur_slot], cur_phase ^ 1, p.abort_flag))) return false;
          d_empty += clock64() - w0;
          int gg = g + g_rot;
          if (gg >= gps) gg -= gps;
          if (elect_one_sync()) {
            mbar_arrive_expect_tx(&full[cur_slot], (uint32_t)(gsz * pl.stage_bytes));
            tma_load_3d(sA + cur_slot * gsz * pl.stage_bytes, &tmap_h, &full[cur_slot], 0, row0, gg * gsz);
          }
          __syncwarp();
          if (p.dbg && s == 100 && i == 0 && g < 8 && lane == 0) p.dbg[blockIdx.x * 128 + 16 + g] = clock64();
        }
        if (++turn == NP) turn = 0;
        if (++cur_slot == n_groups) { cur_slot = 0; cur_phase ^= 1; }
      }
      d_issue += clock64() - c2;
      return true;
    };
    unsigned before[NIF];
#pragma unroll
    for (int i = 0; i < NIF; ++i) before[i] = 0;
    for (int k0 = 0; ok && slot + k0 * p.slots < p.n_bgroups; k0 += NIF) {
      int Tg[NIF], Tw = 0;
#pragma unroll
      for (int i = 0; i < NIF; ++i) {
        Tg[i] = rt_group_steps(p, slot + (k0 + i) * p.slots);
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
    if (p.dbg && lane == 0 && pi == 0) {
      p.dbg[blockIdx.x * 128 + 0] = d_spin;
      p.dbg[blockIdx.x * 128 + 1] = d_fence;
      p.dbg[blockIdx.x * 128 + 2] = d_issue;
      p.dbg[blockIdx.x * 128 + 11] = d_empty;
    }
  } else if (warp == 1) {
    // ---- MMA issuer: whole warp waits for the chunks of a group, one elected lane issues its 16 MMAs ----
    const uint32_t idesc = make_idesc_bf16(p.BP, RT_N);
    // Descriptors differ only in their 14-bit address field (shared-memory offset >> 4), so the four chunks of a
    // group get theirs by one add each and the 16 MMAs go out back to back (building every descriptor from
    // scratch inside the loop costs the single issuing thread more than the MMAs take).
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
          if (p.dbg && s == 100 && i == 0 && g < 8) p.dbg[blockIdx.x * 128 + 40 + g] = clock64();
          const uint32_t a0 = a_lo + (uint32_t)(grp * gsz) * stage16;
          const uint32_t b0 = w_lo + (uint32_t)i0 * (RT_W_BYTES >> 4);
          const int nch = i1 - i0;
#pragma unroll
          for (int j = 0; j < RT_GROUP; ++j) {
            if (j < nch) {
              const uint64_t adesc = desc0 + (uint64_t)(a0 + (uint32_t)j * stage16);
              const uint64_t bdesc = desc0 + (uint64_t)(b0 + (uint32_t)j * (RT_W_BYTES >> 4));
#pragma unroll
              for (int k = 0; k < RT_BK / 16; ++k)
                umma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc,
                          (j | k) ? 1u : (uint32_t)(g != 0));
            }
          }
          umma_commit(&gempty[grp]);
          if (g == gps - 1) umma_commit(&dfull[i]);
        }
        __syncwarp();
        if (++grp == n_groups) { grp = 0; fphase ^= 1; }
      }
      if (p.dbg && s == 100 && i == 0 && lane == 0) p.dbg[blockIdx.x * 128 + 64] = clock64();
      d_rest += clock64() - m0;
      return true;
    };
    for (int k0 = 0; ok && slot + k0 * p.slots < p.n_bgroups; k0 += NIF) {
      int Tg[NIF], Tw = 0;
#pragma unroll
      for (int i = 0; i < NIF; ++i) {
        Tg[i] = rt_group_steps(p, slot + (k0 + i) * p.slots);
        Tw = max(Tw, Tg[i]);
      }
      for (int s = 0; s < Tw && ok; ++s) {
#pragma unroll
        for (int i = 0; i < NIF; ++i)
          if (ok && s < Tg[i]) ok = item(i, s);
      }
    }
    if (p.dbg && lane == 0) {
      p.dbg[blockIdx.x * 128 + 3] = d_wait0;
      p.dbg[blockIdx.x * 128 + 4] = d_rest;
      p.dbg[blockIdx.x * 128 + 10] = d_waitn;
    }
  } else {
    // ---- epilogue: 8 warps.  TMEM lane quarter q = warp % 4 holds batch rows [q*rpq, (q+1)*rpq)
    //      (rpq = 16 for the M = 64 MMA, 32 for M = 128); the two warps of a quarter split the columns.
    const int et = threadIdx.x - 64;     // 0..255
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int rpq = p.BP >> 2;
    const int lrow = q * rpq + lane;     // row inside the batch group (= TMEM lane / hbuf row)
    const int j0 = c * U + half * UH;
    const int ncol = p.dirs * GATES * p.H;
    float hprev[NIF][UH], cst[NIF][UH], bhn[UH];
#pragma unroll
    for (int u = 0; u < UH; ++u)
      bhn[u] = (GATES == 3 && p.b_hn && j0 + u < p.H) ? p.b_hn[(size_t)dir * p.H + j0 + u] : 0.f;
    const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * HS);
    // h store staging sH [BP][U] bf16 + row validity sT [BP].  Single-buffered: the next write happens after
    // this CTA's publish of the item (behind the second named barrier), i.e. after every read of it.
    __nv_bfloat16* sH = reinterpret_cast<__nv_bfloat16*>(sStg);
    int* sT = reinterpret_cast<int*>(smem + pl.st_off);
    const bool vec2 = ((p.H & 1) == 0) && ((UH & 1) == 0);
    unsigned long long e_load = 0, e_wait = 0, e_math = 0, e_bar = 0, e_pub = 0;

    // one step of one group.  hp / cs: this group's recurrent state (registers); returns false on abort.
    auto item = [&](int i, int s, int bg, int b, bool row_ok, int len, unsigned steps_before, float (&hp)[UH],
                    float (&cs)[UH]) -> bool {
      long long e0 = clock64();
      const bool active = row_ok && s < len;
      const int t = dir == 0 ? s : len - 1 - s;
      float gxv[GATES][UH];
      if (active) {
        const float* gp = p.gx + ((size_t)t * p.B + b) * ncol + (size_t)dir * GATES * p.H + j0;
        if (vec2 && j0 + UH <= p.H) {
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
      }
      if (half == 0 && lane < rpq) sT[lrow] = active ? t : -1;
      if (NIF > 1 && row_ok && s + 1 < len) {
        // With several groups in flight the loads above are no longer hidden behind this group's MMA wait; the
        // pre-activations of the group's NEXT step are pulled into L2 now, one round of items ahead.
        const int tn = dir == 0 ? s + 1 : len - 2 - s;
        const float* gn = p.gx + ((size_t)tn * p.B + b) * ncol + (size_t)dir * GATES * p.H + j0;
#pragma unroll
        for (int g = 0; g < GATES; ++g) {
          asm volatile("prefetch.global.L2 [%0];" ::"l"(gn + (size_t)g * p.H));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(gn + (size_t)g * p.H + (UH - 1)));
        }
      }
      long long e1 = clock64();
      const bool ok = wait_abortable(&dfull[i], (uint32_t)((steps_before + (unsigned)s) & 1u), p.abort_flag);
      long long e2 = clock64();
      if (p.dbg && s == 100 && i == 0 && et == 64) p.dbg[blockIdx.x * 128 + 65] = e2;
      tc_fence_after();
      uint32_t r[32];
      tmem_ld32(t_addr + (uint32_t)(i * 64), r);
      tmem_ld_wait();
      tc_fence_before();
      if constexpr (SPLITM && GATES == 3) {
        constexpr int UL = UH / 2;
        const int src = lane & 15;
        const bool hi = lane >= 16;
        const bool act2 = __shfl_sync(0xffffffffu, (int)(ok && active), src) != 0;
        float hn[UL];
#pragma unroll
        for (int u = 0; u < UL; ++u) {
          // unit u of the lower half stays, unit UL+u travels to lane+16
          float av[3], gv[3];
#pragma unroll
          for (int g = 0; g < 3; ++g) {
            const float a_hi = __shfl_sync(0xffffffffu, __uint_as_float(r[(UL + u) * 3 + g]), src);
            const float g_hi = __shfl_sync(0xffffffffu, gxv[g][UL + u], src);
            av[g] = hi ? a_hi : __uint_as_float(r[u * 3 + g]);
            gv[g] = hi ? g_hi : gxv[g][u];
          }
          const float h_hi = __shfl_sync(0xffffffffu, hp[UL + u], src);
          const float hpv = hi ? h_hi : hp[u];
          const float bh = hi ? bhn[UL + u] : bhn[u];
          const float rg = fast_sigmoid(gv[0] + av[0]);
          const float zg = fast_sigmoid(gv[1] + av[1]);
          const float ng = fast_tanh(gv[2] + rg * (av[2] + bh));
          hn[u] = (1.0f - zg) * ng + zg * hpv;
        }
        __nv_bfloat16* sh = sH + (size_t)(q * rpq + src) * U + half * UH + (hi ? UL : 0);
#pragma unroll
        for (int u = 0; u < UL; ++u) {
          const float back = __shfl_sync(0xffffffffu, hn[u], src + 16);
          if (act2) sh[u] = __float2bfloat16_rn(hn[u]);
          if (ok && active) {   // lanes 0-15: the state of all UH units
            hp[u] = hn[u];
            hp[UL + u] = back;
          }
        }
      } else
      if (ok && active) {
        __nv_bfloat16* sh = sH + (size_t)lrow * U + half * UH;
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
          sh[u] = __float2bfloat16_rn(hn);
        }
      }
      long long e3 = clock64();
      const bool all_ok = bar_red_and(ok, 1, 256);     // staging complete (and uniform abort decision)
      if (!all_ok) return false;
      // h_t -> global (bf16), coalesced: each row contributes U contiguous values
      {
        const int n_valid = min(U, p.H - c * U);       // units of this CTA inside H
        __nv_bfloat16* hrow0 = p.hbuf + ((size_t)((bg * 2 + ((s + 1) & 1)) * p.dirs + dir) * p.BP) * p.HP + c * U;
        if ((U & 3) == 0 && n_valid == U && (p.HP & 3) == 0) {
          const int per_row = U / 4;                   // 8-byte pieces
          for (int k = et; k < p.BP * per_row; k += 256) {
            const int row = k / per_row, part = k - row * per_row;
            if (sT[row] >= 0)
              *reinterpret_cast<uint2*>(hrow0 + (size_t)row * p.HP + part * 4) =
                  *reinterpret_cast<const uint2*>(sH + (size_t)row * U + part * 4);
          }
        } else {
          for (int k = et; k < p.BP * U; k += 256) {
            const int row = k / U, u = k - row * U;
            if (sT[row] >= 0 && u < n_valid) hrow0[(size_t)row * p.HP + u] = sH[(size_t)row * U + u];
          }
        }
      }
      named_bar_sync(2, 256);                          // all h stores issued
      long long e4 = clock64();
      if (et == 0) red_release_gpu_add(ctr0 + i * kRnnCounterStride, 1u);   // publish h_t (release: cumulative over the CTA)
      if (p.dbg && i == 0 && s == 99 && et == 0) p.dbg[blockIdx.x * 128 + 66] = clock64();
      if (p.dbg && i == 0 && s == 100 && et == 0) p.dbg[blockIdx.x * 128 + 67] = clock64();
      // y_t -> global (fp32) straight from registers, after the publish: nobody waits on these stores
      if (ok && active) {
        float* yo = p.y + (((size_t)dir * p.T + t) * p.B + b) * p.H + j0;
        if ((UH & 1) == 0 && (p.H & 1) == 0 && j0 + UH <= p.H) {
#pragma unroll
          for (int u = 0; u < UH; u += 2)
            *reinterpret_cast<float2*>(yo + u) = make_float2(hp[u], hp[u + 1 < UH ? u + 1 : u]);
        } else {
#pragma unroll
          for (int u = 0; u < UH; ++u)
            if (j0 + u < p.H) yo[u] = hp[u];
        }
      }
      e_load += e1 - e0; e_wait += e2 - e1; e_math += e3 - e2; e_bar += e4 - e3; e_pub += clock64() - e4;
      return true;
    };

    unsigned before[NIF];
#pragma unroll
    for (int i = 0; i < NIF; ++i) before[i] = 0;
    bool alive = true;
    for (int k0 = 0; alive && slot + k0 * p.slots < p.n_bgroups; k0 += NIF) {
      int Tg[NIF], Tw = 0, bb[NIF], ln[NIF];
      bool rok[NIF];
#pragma unroll
      for (int i = 0; i < NIF; ++i) {
        const int bg = slot + (k0 + i) * p.slots;
        Tg[i] = rt_group_steps(p, bg);
        Tw = max(Tw, Tg[i]);
        bb[i] = bg * p.BP + lrow;
        rok[i] = bg < p.n_bgroups && lane < rpq && bb[i] < p.B;
        ln[i] = rok[i] ? (p.lens ? p.lens[bb[i]] : p.Tmax) : 0;
#pragma unroll
        for (int u = 0; u < UH; ++u) {
          const bool in = rok[i] && j0 + u < p.H;
          hprev[i][u] = (p.h0 && in) ? p.h0[((size_t)dir * p.B + bb[i]) * p.H + j0 + u] : 0.f;
          cst[i][u] = (GATES == 4 && p.c0 && in) ? p.c0[((size_t)dir * p.B + bb[i]) * p.H + j0 + u] : 0.f;
        }
      }
      for (int s = 0; s < Tw && alive; ++s) {
#pragma unroll
        for (int i = 0; i < NIF; ++i)
          if (alive && s < Tg[i])
            alive = item(i, s, slot + (k0 + i) * p.slots, bb[i], rok[i], ln[i], before[i], hprev[i], cst[i]);
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
      p.dbg[blockIdx.x * 128 + 5] = e_load;
      p.dbg[blockIdx.x * 128 + 6] = e_wait;
      p.dbg[blockIdx.x * 128 + 7] = e_math;
      p.dbg[blockIdx.x * 128 + 8] = e_bar;
      p.dbg[blockIdx.x * 128 + 9] = e_pub;
    }
  }
done:
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// W_hh [dirs][G*H][H] fp32 -> per-CTA slices [dirs][cpd][64 rows][HP] bf16.  Row n of a slice:
// half = n/HS, r = n%HS, u = r/G, g = r%G  <->  W_hh[g*H + c*U + half*UH + u][:]  (zero rows/cols beyond H).
__global__ void pack_whh_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int dirs, int cpd,
                                int G, int H, int HP, int U, int HS) {
  const int UH = HS / G, RT_N = 2 * HS;
  const int64_t total = (int64_t)dirs * cpd * RT_N * HP;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % HP);
    int64_t rr = i / HP;
    const int n = (int)(rr % RT_N);
    rr /= RT_N;
    const int c = (int)(rr % cpd);
    const int d = (int)(rr / cpd);
    const int half = n / HS, r = n % HS, u = r / G, g = r % G;
    const int j = c * U + half * UH + u;
    float v = 0.f;
    if (u < UH && j < H && k < H) v = w[((int64_t)d * G * H + (int64_t)g * H + j) * H + k];
    out[i] = __float2bfloat16_rn(v);
  }
}

// next-layer operand: x[t*B+b][j] = bf16(y_fwd + y_bwd) for t < len_b, else 0; optional fp32 copy
__global__ void combine_dirs_kernel(const float* __restrict__ y, int dirs, int T, int B, int H,
                                    const int32_t* __restrict__ lens, __nv_bfloat16* __restrict__ xb, int ldx,
                                    float* __restrict__ xf) {
  const int64_t total = (int64_t)T * B * H;
  const int64_t dstride = total;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % H);
    const int64_t tb = i / H;
    const int b = (int)(tb % B), t = (int)(tb / B);
    float v = 0.f;
    if (t < lens[b]) {
      v = y[i];
      if (dirs == 2) v += y[i + dstride];
    }
    if (xb) xb[tb * ldx + j] = __float2bfloat16_rn(v);
    if (xf) xf[i] = v;
  }
}

}  // namespace tc

// Batches above 64 rows run as groups of 64 rows inside the same launch (W_hh stays resident), up to RT_MAX_NIF of
// them in flight per CTA; when the CTAs of one direction leave SMs free, groups also run side by side on
// independent CTA sets ("slots").  dsb_tune_set("rnn_in_flight", 1) restores the round-1 scheme (groups of 128 rows,
// one at a time).
int rnn_tc_max_in_flight() {
  const int v = g_tune.rnn_in_flight.load();
  return v < 1 ? 1 : (v > tc::RT_MAX_NIF ? tc::RT_MAX_NIF : v);
}
static int rt_bp(int B) { return (B <= 64 || rnn_tc_max_in_flight() > 1) ? 64 : 128; }

// Picks the W_hh slicing of a layer: 64-row slices when they fit next to a ring of at least one slot, else 48-row
// ("narrow") slices -- more CTAs per direction, which for wide bidirectional layers means one launch per direction.
bool rnn_tc_narrow(const RnnLayer& L, int B) {
  const int HP = (L.H + 63) / 64 * 64;
  const tc::RtPlan pl = tc::rt_plan(HP / 64, rt_bp(B), tc::rt_units(L.gates, false), g_tune.rnn_ring_gsz.load(), false);
  return pl.groups < 1 || pl.gsz < 1 || pl.total > tc::RT_SMEM_LIMIT;
}

// True when this layer / batch takes the CTA-pair kernel and no other kernel has been asked for: the caller then hands
// over batch-minor pre-activations and gets batch-minor outputs (see rnn_layer_pair).
bool rnn_batch_minor(const RnnLayer& L, int B) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const bool ks = L.ks_recurrence && g_tune.rnn_ksplit.load() && rnn_tc_max_in_flight() > 1;
  return !ks && g_tune.rnn_batch_minor.load() && rnn_pair_supported(L, B, sms);
}

bool rnn_tc_supported(const RnnLayer& L, int B, int sms, int* cpd_out, int* launches_out) {
  const bool narrow = rnn_tc_narrow(L, B);
  const int U = tc::rt_units(L.gates, narrow);
  const int cpd = cdiv(L.H, U);
  const int HP = (L.H + 63) / 64 * 64;
  const tc::RtPlan pl = tc::rt_plan(HP / 64, rt_bp(B), U, g_tune.rnn_ring_gsz.load(), narrow);
  if (B < 1 || pl.groups < 1 || pl.gsz < 1 || pl.total > tc::RT_SMEM_LIMIT || cpd > sms) return false;
  if (cpd_out) *cpd_out = cpd;
  if (launches_out) *launches_out = (L.dirs * cpd <= sms) ? 1 : L.dirs;
  return true;
}

size_t rnn_tc_hbuf_elems(const RnnLayer& L, int B) {
  const int HP = (L.H + 63) / 64 * 64, BP = rt_bp(B);
  // an even number of groups: the CTA-pair kernel (rnn_pair.cu) streams a (zero) buffer for the missing partner of the last group
  return (size_t)((cdiv(B, BP) + 1) / 2 * 2) * 2 * L.dirs * BP * HP;
}

size_t rnn_tc_pack_elems(const RnnLayer& L) {
  const bool narrow = rnn_tc_narrow(L, 64);
  return (size_t)L.dirs * cdiv(L.H, tc::rt_units(L.gates, narrow)) * tc::rt_rows(narrow) * ((L.H + 63) / 64 * 64);
}

int pack_whh_tc(const RnnLayer& L, __nv_bfloat16* out, cudaStream_t st) {
  const bool narrow = rnn_tc_narrow(L, 64);
  const int U = tc::rt_units(L.gates, narrow), cpd = cdiv(L.H, U), HP = (L.H + 63) / 64 * 64;
  const int64_t total = (int64_t)rnn_tc_pack_elems(L);
  tc::pack_whh_kernel<<<(int)(cdiv64(total, 256) < 2048 ? cdiv64(total, 256) : 2048), 256, 0, st>>>(
      L.w_hh, out, L.dirs, cpd, L.gates, L.H, HP, U, tc::rt_hs(narrow));
  DSB_CHECK_LAUNCH();
  return 0;
}

namespace tc {
// 4 hidden units per thread (H % 4 == 0): float4 loads of both directions, 8-byte bf16 stores
__global__ void combine_dirs_vec4_kernel(const float* __restrict__ y, int dirs, int T, int B, int H,
                                         const int32_t* __restrict__ lens, __nv_bfloat16* __restrict__ xb, int ldx,
                                         float* __restrict__ xf) {
  const int H4 = H >> 2;
  const int64_t total4 = (int64_t)T * B * H4;
  const int64_t dstride4 = total4;
  const float4* y4 = reinterpret_cast<const float4*>(y);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
    const int j4 = (int)(i % H4);
    const int64_t tb = i / H4;
    const int b = (int)(tb % B), t = (int)(tb / B);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t < lens[b]) {
      v = y4[i];
      if (dirs == 2) {
        const float4 w = y4[i + dstride4];
        v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
      }
    }
    if (xb) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo);
      pk.y = *reinterpret_cast<uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(xb + tb * ldx + j4 * 4) = pk;
    }
    if (xf) reinterpret_cast<float4*>(xf)[i] = v;
  }
}
}  // namespace tc

namespace tc {
// Batch-minor per-direction outputs yt [dirs][H][T*B] (rnn_pair.cu) -> next-layer operand x[t*B+b][j] = bf16(fwd + bwd)
// for t < len_b, else 0 (and / or the fp32 copy): a 64 x 64 tile transpose through shared memory, coalesced both ways.
__global__ void __launch_bounds__(256) combine_dirs_t_kernel(const float* __restrict__ yt, int dirs, int T, int B, int H,
                                                             const int32_t* __restrict__ lens,
                                                             __nv_bfloat16* __restrict__ xb, int ldx,
                                                             float* __restrict__ xf) {
  __shared__ float tile[64][65];
  __shared__ int live[64];
  const int64_t M = (int64_t)T * B;
  const int64_t m0 = (int64_t)blockIdx.x * 64;
  const int j0 = blockIdx.y * 64;
  const int tid = threadIdx.x;
  if (tid < 64) {
    const int64_t m = m0 + tid;
    int lv = 0;
    if (m < M) {
      const int b = (int)(m % B), t = (int)(m / B);
      lv = t < lens[b];
    }
    live[tid] = lv;
  }
  __syncthreads();
  const int g = tid & 15, r = tid >> 4;   // 16 groups of four columns x 16 rows per pass
  if ((M & 3) == 0) {
    // 16-byte loads along the (t, b) axis: a quarter warp reads 256 contiguous bytes of one unit
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {
      const int jj = r + 16 * pass, j = j0 + jj;
      const int64_t m = m0 + 4 * g;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j < H && m < M) {
        v = *reinterpret_cast<const float4*>(yt + (int64_t)j * M + m);
        if (dirs == 2) {
          const float4 w = *reinterpret_cast<const float4*>(yt + ((int64_t)H + j) * M + m);
          v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
        }
      }
      tile[jj][4 * g + 0] = live[4 * g + 0] ? v.x : 0.f;
      tile[jj][4 * g + 1] = live[4 * g + 1] ? v.y : 0.f;
      tile[jj][4 * g + 2] = live[4 * g + 2] ? v.z : 0.f;
      tile[jj][4 * g + 3] = live[4 * g + 3] ? v.w : 0.f;
    }
  } else {
    const int tx = tid & 63, ty = tid >> 6;
    for (int jj = ty; jj < 64; jj += 4) {
      const int j = j0 + jj;
      const int64_t m = m0 + tx;
      float v = 0.f;
      if (live[tx] && j < H) {
        v = yt[(int64_t)j * M + m];
        if (dirs == 2) v += yt[((int64_t)H + j) * M + m];
      }
      tile[jj][tx] = v;
    }
  }
  __syncthreads();
  // rows of the output: four consecutive units per thread (8 bytes of bf16, 16 bytes of fp32)
  const bool vec = (H & 3) == 0 && (ldx & 3) == 0;
#pragma unroll
  for (int pass = 0; pass < 4; ++pass) {
    const int mm = r + 16 * pass;
    const int64_t m = m0 + mm;
    const int j = j0 + 4 * g;
    if (m >= M || j >= H) continue;
    const float v0 = tile[4 * g + 0][mm], v1 = tile[4 * g + 1][mm], v2 = tile[4 * g + 2][mm], v3 = tile[4 * g + 3][mm];
    if (vec) {     // H % 4 == 0: the four units are all inside H
      if (xb) {
        const __nv_bfloat162 lo = __floats2bfloat162_rn(v0, v1), hi = __floats2bfloat162_rn(v2, v3);
        uint2 pk;
        pk.x = *reinterpret_cast<const uint32_t*>(&lo);
        pk.y = *reinterpret_cast<const uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(xb + m * ldx + j) = pk;
      }
      if (xf) *reinterpret_cast<float4*>(xf + m * H + j) = make_float4(v0, v1, v2, v3);
    } else {
      const float vv[4] = {v0, v1, v2, v3};
      for (int e = 0; e < 4; ++e)
        if (j + e < H) {
          if (xb) xb[m * ldx + j + e] = __float2bfloat16_rn(vv[e]);
          if (xf) xf[m * H + j + e] = vv[e];
        }
    }
  }
}
}  // namespace tc

int combine_dirs_t_tc(const float* yt, int dirs, int T, int B, int H, const int32_t* d_len, __nv_bfloat16* xb, int ldx,
                      float* xf, cudaStream_t st) {
  const int64_t M = (int64_t)T * B;
  dim3 grid((unsigned)cdiv64(M, 64), (unsigned)cdiv(H, 64));
  tc::combine_dirs_t_kernel<<<grid, 256, 0, st>>>(yt, dirs, T, B, H, d_len, xb, ldx, xf);
  DSB_CHECK_LAUNCH();
  return 0;
}

int combine_dirs_tc(const float* y, int dirs, int T, int B, int H, const int32_t* d_len, __nv_bfloat16* xb, int ldx,
                    float* xf, cudaStream_t st) {
  if ((H & 3) == 0 && (ldx & 3) == 0) {
    const int64_t total4 = (int64_t)T * B * (H >> 2);
    tc::combine_dirs_vec4_kernel<<<(int)(cdiv64(total4, 256) < 148 * 16 ? cdiv64(total4, 256) : 148 * 16), 256, 0, st>>>(
        y, dirs, T, B, H, d_len, xb, ldx, xf);
    DSB_CHECK_LAUNCH();
    return 0;
  }
  const int64_t total = (int64_t)T * B * H;
  tc::combine_dirs_kernel<<<(int)(cdiv64(total, 256) < 148 * 16 ? cdiv64(total, 256) : 148 * 16), 256, 0, st>>>(
      y, dirs, T, B, H, d_len, xb, ldx, xf);
  DSB_CHECK_LAUNCH();
  return 0;
}

namespace tc {
// slot 0 of every batch group's exchange buffer <- bf16(h0); rows/columns beyond B/H stay zero
__global__ void init_hbuf_kernel(const float* __restrict__ h0, __nv_bfloat16* __restrict__ hbuf, int dirs, int B, int H,
                                 int HP, int BP, int n_bgroups) {
  const int64_t total = (int64_t)n_bgroups * dirs * BP * H;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % H);
    int64_t r = i / H;
    const int row = (int)(r % BP);
    r /= BP;
    const int d = (int)(r % dirs), bg = (int)(r / dirs);
    const int b = bg * BP + row;
    if (b < B) hbuf[(((int64_t)(bg * 2) * dirs + d) * BP + row) * HP + k] = __float2bfloat16_rn(h0[((int64_t)d * B + b) * H + k]);
  }
}
}  // namespace tc

int rnn_tc_init_hbuf(const float* h0, __nv_bfloat16* hbuf, int dirs, int B, int H, int HP, int BP, int n_bgroups,
                     cudaStream_t st) {
  const int64_t total = (int64_t)n_bgroups * dirs * BP * H;
  tc::init_hbuf_kernel<<<(int)(cdiv64(total, 256) < 1184 ? cdiv64(total, 256) : 1184), 256, 0, st>>>(h0, hbuf, dirs, B, H, HP, BP,
                                                                                                     n_bgroups);
  DSB_CHECK_LAUNCH();
  return 0;
}

// One BatchRNN layer.  gx [T*B][dirs*G*H] fp32, y [dirs][T][B][H] fp32 (rows t >= len_b are NOT written),
// hbuf rnn_tc_hbuf_elems() bf16, sync_words = kRnnSyncCounters u32 step counters, abort_flag = one sticky i32 (device).
// d_len == nullptr: every sequence runs Tmax steps.  h0/c0 (nullptr = zeros) and hT/cT (nullptr = dropped)
// are [dirs][B][H] fp32 and may alias (streaming state carried across chunks, model.py:219-237).
int rnn_layer_tc(const RnnLayer& L, const float* gx, const int32_t* d_len, int B, int T, int Tmax, float* y,
                 __nv_bfloat16* hbuf, unsigned int* sync_words, int* abort_flag, cudaStream_t st, const float* h0,
                 const float* c0, float* hT, float* cT, bool batch_minor) {
  using namespace tc;
  if (batch_minor) {   // the caller asked rnn_batch_minor() first
    if (!rnn_batch_minor(L, B)) return set_error(DSB_ERR_INVALID, "rnn_layer_tc: batch-minor layout needs the CTA-pair kernel");
    return rnn_layer_pair(L, gx, d_len, B, T, Tmax, y, hbuf, sync_words, abort_flag, st, h0, c0, hT, cT, true);
  }
  if (L.ks_recurrence && g_tune.rnn_ksplit.load() && rnn_tc_max_in_flight() > 1)
    return rnn_layer_ks(L, gx, d_len, B, T, Tmax, y, hbuf, sync_words, abort_flag, st, h0, c0, hT, cT);
  int dev = 0, sms = 148, cpd = 0, launches = 1;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (rnn_pair_supported(L, B, sms))
    return rnn_layer_pair(L, gx, d_len, B, T, Tmax, y, hbuf, sync_words, abort_flag, st, h0, c0, hT, cT);
  if (!rnn_tc_supported(L, B, sms, &cpd, &launches))
    return set_error(DSB_ERR_UNSUPPORTED, "rnn_layer_tc: shape H=%d B=%d not supported", L.H, B);
  const int HP = (L.H + 63) / 64 * 64, BP = rt_bp(B), nkc = HP / 64;
  const int n_bgroups = cdiv(B, BP);
  const int dirs_per_launch = L.dirs / launches;
  int slots = sms / (dirs_per_launch * cpd);
  if (slots > n_bgroups) slots = n_bgroups;
  if (g_tune.rnn_max_slots.load() > 0 && slots > g_tune.rnn_max_slots.load()) slots = g_tune.rnn_max_slots.load();
  if (slots < 1) slots = 1;
  int nif = cdiv(n_bgroups, slots);          // groups per CTA set; up to RT_MAX_NIF of them in flight
  if (nif > rnn_tc_max_in_flight()) nif = rnn_tc_max_in_flight();
  while (slots > 1 && L.dirs * slots * nif > kRnnMaxCounters) --slots;
  DSB_CUDA(cudaMemsetAsync(hbuf, 0, sizeof(__nv_bfloat16) * rnn_tc_hbuf_elems(L, B), st));
  DSB_CUDA(cudaMemsetAsync(sync_words, 0, sizeof(unsigned int) * kRnnSyncCounters, st));   // step counters only; abort flag is sticky
  if (h0)
    if (int e = rnn_tc_init_hbuf(h0, hbuf, L.dirs, B, L.H, HP, BP, n_bgroups, st)) return e;

  CUtensorMap tw, th;
  const bool narrow = rnn_tc_narrow(L, B);
  const int U = rt_units(L.gates, narrow);
  uint64_t dw[2] = {(uint64_t)HP, (uint64_t)L.dirs * cpd * rt_rows(narrow)}, sw[2] = {2, (uint64_t)HP * 2};
  uint32_t bw[2] = {RT_BK, (uint32_t)rt_rows(narrow)};
  if (int e = make_tmap_bf16(&tw, L.w_hh_pack, 2, dw, sw, bw, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
  // h exchange buffer as {64 k, rows, K chunks}: one box {64, BP, RT_GROUP} = RT_GROUP consecutive chunk tiles
  uint64_t dh[3] = {(uint64_t)RT_BK, (uint64_t)n_bgroups * 2 * L.dirs * BP, (uint64_t)nkc};
  uint64_t sh[3] = {2, (uint64_t)HP * 2, (uint64_t)RT_BK * 2};
  const int ring_gsz = g_tune.rnn_ring_gsz.load();
  const int gsz = rt_plan(nkc, BP, U, ring_gsz, narrow).gsz;
  uint32_t bh[3] = {RT_BK, (uint32_t)BP, (uint32_t)gsz};
  if (int e = make_tmap_bf16(&th, hbuf, 3, dh, sh, bh, CU_TENSOR_MAP_SWIZZLE_128B)) return e;

  RnnTcParams p{};
  p.gx = gx;
  p.b_hn = L.b_hn;
  p.y = y;
  p.hbuf = hbuf;
  p.lens = d_len;
  p.counters = sync_words;
  p.abort_flag = abort_flag;
  p.h0 = h0; p.c0 = c0; p.hT = hT; p.cT = cT;
  p.n_bgroups = n_bgroups; p.slots = slots;
  p.B = B; p.H = L.H; p.HP = HP; p.BP = BP; p.T = T; p.Tmax = Tmax;
  p.dirs = L.dirs; p.cpd = cpd; p.U = U; p.nkc = nkc;
  p.ring_gsz = ring_gsz;
  // two producers own alternate ring slots, which only works when the ring has an even number of them (a producer
  // must never be two uses of the same slot ahead of the MMAs: the slot barriers carry one parity bit)
  p.n_producers = (g_tune.rnn_producers.load() == 2 && rt_plan(nkc, BP, U, ring_gsz, narrow).groups % 2 == 0) ? 2 : 1;
  const size_t smem = (size_t)rt_plan(nkc, BP, U, ring_gsz, narrow).total;
  static const int split_env = getenv("DSB_RNN_SPLIT") ? atoi(getenv("DSB_RNN_SPLIT")) : 1;
  const bool split = L.gates == 3 && BP == 64 && split_env;
  const void* fn = nullptr;
#define RT_PICK2(G, S, W)                                                                             \
  fn = nif == 1 ? (const void*)rnn_tc_kernel<G, S, 1, W> : nif == 2 ? (const void*)rnn_tc_kernel<G, S, 2, W> \
                                                                   : (const void*)rnn_tc_kernel<G, S, 3, W>
#define RT_PICK(G, S) do { if (narrow) { RT_PICK2(G, S, true); } else { RT_PICK2(G, S, false); } } while (0)
  if (L.gates == 3) { if (split) RT_PICK(3, true); else RT_PICK(3, false); }
  else if (L.gates == 4) RT_PICK(4, false);
  else RT_PICK(1, false);
#undef RT_PICK2
#undef RT_PICK
  DSB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  static const bool debug = getenv("DSB_RNN_DEBUG") != nullptr;
  unsigned long long* dbg = nullptr;
  const int grid = dirs_per_launch * slots * cpd;
  if (debug) {
    DSB_CUDA(cudaMalloc(&dbg, sizeof(unsigned long long) * 128 * grid));
    DSB_CUDA(cudaMemsetAsync(dbg, 0, sizeof(unsigned long long) * 128 * grid, st));
  }
  p.dbg = dbg;
  for (int l = 0; l < launches; ++l) {
    p.dir0 = l * dirs_per_launch;
    void* args[] = {(void*)&tw, (void*)&th, (void*)&p};
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(RT_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attrs[1];
    attrs[0].id = cudaLaunchAttributeCooperative;          // all CTAs co-resident (they spin on each other)
    attrs[0].val.cooperative = 1;
    cfg.attrs = attrs;
    cfg.numAttrs = 1;
    const cudaError_t le = cudaLaunchKernelExC(&cfg, fn, args);
    if (le != cudaSuccess)
      return set_error(DSB_ERR_CUDA, "rnn_layer_tc: launch failed: %s", cudaGetErrorString(le));
    count_launch();
  }
  if (debug) {
    std::vector<unsigned long long> h(128 * grid);
    DSB_CUDA(cudaStreamSynchronize(st));
    DSB_CUDA(cudaMemcpy(h.data(), dbg, sizeof(unsigned long long) * 128 * grid, cudaMemcpyDeviceToHost));
    cudaFree(dbg);
    const char* names[12] = {"prod.spin", "prod.fence", "prod.issue", "mma.wait_first", "mma.rest", "epi.gload",
                             "epi.wait_mma", "epi.math_store", "epi.bar", "epi.publish", "mma.wait_rest",
                             "prod.wait_empty"};
    const int items = cdiv(n_bgroups, slots) * Tmax;   // (step, group) items per CTA (upper bound for ragged groups)
    const RtPlan dpl = rt_plan(nkc, BP, U, ring_gsz, narrow);
    fprintf(stderr, "[rnn_tc debug] H=%d B=%d Tmax=%d grid=%d groups=%d slots=%d in flight=%d ring=%dx%d  cycles/item (avg over CTAs | max CTA)\n", L.H, B, Tmax, grid, n_bgroups, slots, nif, dpl.groups, dpl.gsz);
    for (int k = 0; k < 12; ++k) {
      double sum = 0, mx = 0;
      for (int c = 0; c < grid; ++c) { double v = (double)h[c * 128 + k] / items; sum += v; mx = v > mx ? v : mx; }
      fprintf(stderr, "   %-16s %9.0f | %9.0f\n", names[k], sum / grid, mx);
    }
    for (int c = 0; c < grid; c += grid / 2 + 1) {   // step-100 timeline of two CTAs (cycles since this CTA published step 99)
      const unsigned long long* d = &h[c * 128];
      const long long t0 = (long long)d[66];
      fprintf(stderr, "   [cta %d] barrier passed %+lld | tma group issued:", c, (long long)d[15] - t0);
      for (int i = 0; i < (nkc + gsz - 1) / gsz && i < 8; ++i) fprintf(stderr, " %lld", (long long)d[16 + i] - t0);
      fprintf(stderr, "\n   [cta %d] group ready:", c);
      for (int i = 0; i < (nkc + gsz - 1) / gsz && i < 8; ++i) fprintf(stderr, " %lld", (long long)d[40 + i] - t0);
      fprintf(stderr, "\n   [cta %d] mma issued %+lld | epilogue saw dfull %+lld | published %+lld\n", c,
              (long long)d[64] - t0, (long long)d[65] - t0, (long long)d[67] - t0);
    }
  }
  return 0;
}

}  // namespace dsb
