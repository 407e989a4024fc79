// fp32 (CUDA-core) recurrence for BatchRNN: GRU / LSTM / tanh-RNN, uni- or bi-directional,
// packed-sequence semantics.
//
// Replaces the recurrent half of torch.nn.GRU/LSTM/RNN as used by BatchRNN.forward
// (danspeech/deepspeech/model.py:114-122): pack_padded_sequence -> rnn -> pad_packed_sequence ->
// sum of the two directions.  Sequence b runs t = 0..len_b-1 forwards and t = len_b-1..0 backwards
// from its own end, outputs at t >= len_b stay exactly 0, and the directions are summed.
//
// One launch per time step (exact-fp32 verification path; the persistent tensor-core recurrence is
// rnn_tc.cu).  A CTA owns 16 hidden units of one direction for up to 64 sequences and streams
// W_hh[:, k-chunk] and h_{t-1}[:, k-chunk] through shared memory: rnn_step_f32_v2_kernel (cp.async ring,
// 24-accumulator register tiles) when H % 4 == 0, rnn_step_f32_kernel otherwise (DSB_RNN_F32_V1=1 forces it).
#include "model_types.cuh"

namespace dsb {

constexpr int RJ = 16;   // hidden units per CTA
constexpr int RB = 64;   // sequences per CTA
constexpr int RK = 32;   // k-chunk

__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

template <int GATES>
__global__ void __launch_bounds__(256)
rnn_step_f32_kernel(const float* __restrict__ gx,      // [T*B][dirs*GATES*H]
                    const float* __restrict__ w_hh,    // [dirs][GATES*H][H]
                    const float* __restrict__ b_hh,    // [dirs][GATES*H]
                    const float* __restrict__ h_prev,  // [dirs][B][H]
                    float* __restrict__ h_next,        // [dirs][B][H]
                    float* __restrict__ c_state,       // [dirs][B][H] (LSTM)
                    float* __restrict__ y,             // [T][B][H]
                    const int32_t* __restrict__ lens, int step, int B, int H, int dirs) {
  __shared__ __align__(16) float hs[RK][RB + 4];
  __shared__ float wsm[RK][GATES * RJ + 1];
  const int dir = blockIdx.y;
  const int j0 = blockIdx.x * RJ;
  const int b0 = blockIdx.z * RB;
  const int tid = threadIdx.x;
  const int jj = tid & 15, bg = tid >> 4;

  // whole tile inactive?  (lens sorted descending: row b0 is the longest of the tile)
  if (step >= lens[b0]) return;

  const float* hp = h_prev + (int64_t)dir * B * H;
  const float* wd = w_hh + (int64_t)dir * GATES * H * H;

  float acc[GATES][4];
#pragma unroll
  for (int g = 0; g < GATES; ++g)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[g][i] = 0.0f;

  for (int k0 = 0; k0 < H; k0 += RK) {
    __syncthreads();
    for (int i = tid; i < RB * RK; i += 256) {
      int k = i % RK, bb = i / RK;
      float v = 0.0f;
      if (b0 + bb < B && k0 + k < H) v = hp[(int64_t)(b0 + bb) * H + k0 + k];
      hs[k][bb] = v;
    }
    for (int i = tid; i < GATES * RJ * RK; i += 256) {
      int k = i % RK, r = i / RK;
      int g = r / RJ, j = j0 + (r % RJ);
      float v = 0.0f;
      if (j < H && k0 + k < H) v = __ldg(wd + ((int64_t)g * H + j) * H + k0 + k);
      wsm[k][r] = v;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < RK; ++k) {
      const float4 hv4 = *reinterpret_cast<const float4*>(&hs[k][bg * 4]);
      const float hv[4] = {hv4.x, hv4.y, hv4.z, hv4.w};
#pragma unroll
      for (int g = 0; g < GATES; ++g) {
        const float wv = wsm[k][g * RJ + jj];
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[g][i] = fmaf(wv, hv[i], acc[g][i]);
      }
    }
  }

  const int j = j0 + jj;
  if (j >= H) return;
  const float* bh = b_hh + (int64_t)dir * GATES * H;
  const int ncol = dirs * GATES * H;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int b = b0 + bg * 4 + i;
    if (b >= B) continue;
    const int len = lens[b];
    if (step >= len) continue;
    const int t = dir == 0 ? step : len - 1 - step;
    const float* gxr = gx + ((int64_t)t * B + b) * ncol + (int64_t)dir * GATES * H;
    const int64_t sidx = ((int64_t)dir * B + b) * H + j;
    float hnew;
    if (GATES == 3) {   // GRU, torch gate order r, z, n
      const float r = sigmoid_acc(gxr[j] + acc[0][i] + bh[j]);
      const float z = sigmoid_acc(gxr[H + j] + acc[1 % GATES][i] + bh[H + j]);
      const float n = tanhf(gxr[2 * H + j] + r * (acc[2 % GATES][i] + bh[2 * H + j]));
      const float hprev = hp[(int64_t)b * H + j];
      hnew = (1.0f - z) * n + z * hprev;
    } else if (GATES == 4) {   // LSTM, torch gate order i, f, g, o
      const float ig = sigmoid_acc(gxr[j] + acc[0][i] + bh[j]);
      const float fg = sigmoid_acc(gxr[H + j] + acc[1 % GATES][i] + bh[H + j]);
      const float gg = tanhf(gxr[2 * H + j] + acc[2 % GATES][i] + bh[2 * H + j]);
      const float og = sigmoid_acc(gxr[3 * H + j] + acc[3 % GATES][i] + bh[3 * H + j]);
      const float c = fg * c_state[sidx] + ig * gg;
      c_state[sidx] = c;
      hnew = og * tanhf(c);
    } else {   // nn.RNN (tanh)
      hnew = tanhf(gxr[j] + acc[0][i] + bh[j]);
    }
    h_next[sidx] = hnew;
    float* yo = y + ((int64_t)t * B + b) * H + j;
    if (dirs == 2) atomicAdd(yo, hnew);   // two addends onto 0: order-independent, exact
    else *yo = hnew;
  }
}

// ---- v2 step kernel: same arithmetic (one fmaf chain over k per output, ascending k: bit-identical to the kernel
// above), restructured for the FMA pipe.  A CTA of 4 warps owns 64 sequences x 16 hidden units of one direction; a
// thread 4 sequences x 2 units x GATES rows (24 accumulators for a GRU); W_hh and h_{t-1} arrive in k-chunks of 32
// through a 3-stage cp.async ring (zero fill outside B / H), rows padded to 36 floats so that the 128-bit
// shared-memory reads of a warp (8 distinct h rows, 4 distinct W rows) are conflict-free.  Needs H % 4 == 0 and
// 16-byte aligned operands; everything else takes the kernel above.
constexpr int R2_B = 64, R2_U = 16, R2_K = 32, R2_LD = 36, R2_STAGES = 3, R2_THREADS = 128;

__device__ __forceinline__ void cp_async16_zfill(float* dst, const float* src, bool ok) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  const int n = ok ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}

template <int GATES>
__global__ void __launch_bounds__(R2_THREADS)
rnn_step_f32_v2_kernel(const float* __restrict__ gx, const float* __restrict__ w_hh, const float* __restrict__ b_hh,
                       const float* __restrict__ h_prev, float* __restrict__ h_next, float* __restrict__ c_state,
                       float* __restrict__ y, const int32_t* __restrict__ lens, int step, int B, int H, int dirs) {
  extern __shared__ __align__(16) float r2_smem[];
  constexpr int ROWS = GATES * R2_U;
  constexpr int STAGE = (R2_B + ROWS) * R2_LD;
  const int dir = blockIdx.y;
  const int j0 = blockIdx.x * R2_U;
  const int b0 = blockIdx.z * R2_B;
  if (step >= lens[b0]) return;   // lens sorted descending: row b0 is the longest of the tile
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wb = warp & 1, wu = warp >> 1, bl = lane & 7, ul = lane >> 3;
  const float* hp = h_prev + (int64_t)dir * B * H;
  const float* wd = w_hh + (int64_t)dir * GATES * H * H;
  const float* bh = b_hh + (int64_t)dir * GATES * H;
  const int ncol = dirs * GATES * H;
  const int nk = (H + R2_K - 1) / R2_K;

  auto issue = [&](int kc) {
    float* hs = r2_smem + (kc % R2_STAGES) * STAGE;
    float* ws = hs + R2_B * R2_LD;
    const int k0 = kc * R2_K;
    for (int v = tid; v < R2_B * (R2_K / 4); v += R2_THREADS) {
      const int row = v >> 3, kv = (v & 7) * 4;
      const bool ok = b0 + row < B && k0 + kv < H;
      cp_async16_zfill(hs + row * R2_LD + kv, ok ? hp + (int64_t)(b0 + row) * H + k0 + kv : hp, ok);
    }
    for (int v = tid; v < ROWS * (R2_K / 4); v += R2_THREADS) {
      const int r = v >> 3, kv = (v & 7) * 4;
      const int g = r / R2_U, j = j0 + (r % R2_U);
      const bool ok = j < H && k0 + kv < H;
      cp_async16_zfill(ws + r * R2_LD + kv, ok ? wd + ((int64_t)g * H + j) * H + k0 + kv : wd, ok);
    }
  };
#pragma unroll
  for (int kc = 0; kc < R2_STAGES - 1; ++kc) {
    if (kc < nk) issue(kc);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }

  // this thread's outputs: sequences b0 + wb*32 + i*8 + bl (i < 4), units j0 + wu*8 + s*4 + ul (s < 2).
  // Their gate pre-activations, biases and previous state are fetched now, under the k loop.
  float gxv[GATES][2][4], bhv[GATES][2], hpv[2][4], cpv[2][4];
  int tt[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int b = b0 + wb * 32 + i * 8 + bl;
    const int len = b < B ? lens[b] : 0;
    tt[i] = step < len ? (dir == 0 ? step : len - 1 - step) : -1;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int j = j0 + wu * 8 + s * 4 + ul;
      const bool ok = tt[i] >= 0 && j < H;
      const float* gxr = gx + ((int64_t)(ok ? tt[i] : 0) * B + (ok ? b : 0)) * ncol + (int64_t)dir * GATES * H;
#pragma unroll
      for (int g = 0; g < GATES; ++g) gxv[g][s][i] = ok ? __ldg(gxr + (int64_t)g * H + j) : 0.0f;
      hpv[s][i] = ok ? hp[(int64_t)b * H + j] : 0.0f;
      cpv[s][i] = (ok && GATES == 4) ? c_state[((int64_t)dir * B + b) * H + j] : 0.0f;
    }
  }
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const int j = j0 + wu * 8 + s * 4 + ul;
#pragma unroll
    for (int g = 0; g < GATES; ++g) bhv[g][s] = j < H ? __ldg(bh + (int64_t)g * H + j) : 0.0f;
  }

  float acc[GATES][2][4];
#pragma unroll
  for (int g = 0; g < GATES; ++g)
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[g][s][i] = 0.0f;

  for (int kc = 0; kc < nk; ++kc) {
    asm volatile("cp.async.wait_group %0;" ::"n"(R2_STAGES - 2) : "memory");
    __syncthreads();   // chunk kc has landed for every thread; everyone is done with chunk kc-1 (its stage is refilled next)
    if (kc + R2_STAGES - 1 < nk) issue(kc + R2_STAGES - 1);
    asm volatile("cp.async.commit_group;" ::: "memory");
    const float* hs = r2_smem + (kc % R2_STAGES) * STAGE + (wb * 32 + bl) * R2_LD;
    const float* ws = r2_smem + (kc % R2_STAGES) * STAGE + R2_B * R2_LD + (wu * 8 + ul) * R2_LD;
#pragma unroll 2
    for (int kv = 0; kv < R2_K; kv += 4) {
      float4 hv[4], wv[GATES][2];
#pragma unroll
      for (int i = 0; i < 4; ++i) hv[i] = *reinterpret_cast<const float4*>(hs + i * 8 * R2_LD + kv);
#pragma unroll
      for (int g = 0; g < GATES; ++g)
#pragma unroll
        for (int s = 0; s < 2; ++s)
          wv[g][s] = *reinterpret_cast<const float4*>(ws + (g * R2_U + s * 4) * R2_LD + kv);
#pragma unroll
      for (int g = 0; g < GATES; ++g)
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float a = acc[g][s][i];
            a = fmaf(wv[g][s].x, hv[i].x, a);
            a = fmaf(wv[g][s].y, hv[i].y, a);
            a = fmaf(wv[g][s].z, hv[i].z, a);
            a = fmaf(wv[g][s].w, hv[i].w, a);
            acc[g][s][i] = a;
          }
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (tt[i] < 0) continue;
    const int b = b0 + wb * 32 + i * 8 + bl;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int j = j0 + wu * 8 + s * 4 + ul;
      if (j >= H) continue;
      const int64_t sidx = ((int64_t)dir * B + b) * H + j;
      float hnew;
      if (GATES == 3) {
        const float r = sigmoid_acc(gxv[0][s][i] + acc[0][s][i] + bhv[0][s]);
        const float z = sigmoid_acc(gxv[1 % GATES][s][i] + acc[1 % GATES][s][i] + bhv[1 % GATES][s]);
        const float n = tanhf(gxv[2 % GATES][s][i] + r * (acc[2 % GATES][s][i] + bhv[2 % GATES][s]));
        hnew = (1.0f - z) * n + z * hpv[s][i];
      } else if (GATES == 4) {
        const float ig = sigmoid_acc(gxv[0][s][i] + acc[0][s][i] + bhv[0][s]);
        const float fg = sigmoid_acc(gxv[1 % GATES][s][i] + acc[1 % GATES][s][i] + bhv[1 % GATES][s]);
        const float gg = tanhf(gxv[2 % GATES][s][i] + acc[2 % GATES][s][i] + bhv[2 % GATES][s]);
        const float og = sigmoid_acc(gxv[3 % GATES][s][i] + acc[3 % GATES][s][i] + bhv[3 % GATES][s]);
        const float c = fg * cpv[s][i] + ig * gg;
        c_state[sidx] = c;
        hnew = og * tanhf(c);
      } else {
        hnew = tanhf(gxv[0][s][i] + acc[0][s][i] + bhv[0][s]);
      }
      h_next[sidx] = hnew;
      float* yo = y + ((int64_t)tt[i] * B + b) * H + j;
      if (dirs == 2) atomicAdd(yo, hnew);   // two addends onto 0: order-independent, exact
      else *yo = hnew;
    }
  }
}

// ---- v3 step kernel: the v2 tile with more warps per SM and no second wave.  ncu on v2 (H = 1200, 150 CTAs of 4 warps
// on 148 SMs): one warp per scheduler issues 0.42 instructions per cycle (fixed-latency and shared-memory waits are
// exposed) and the two SMs that hold two CTAs set the step time.  Here a CTA owns 64 sequences x 20 units (120 CTAs
// for H = 1200: one per SM) and runs 12 warps: two k-groups of six warps take the even / odd k-chunks through their
// own cp.async rings (named barriers), warps 0-7 are the 2 x 2 full tiles of v2 (units 0-15), warps 8-11 the half
// tiles of units 16-19, so that every scheduler gets two full and one half warp; the k-groups' partial sums meet in
// shared memory before the gate math.
constexpr int R3_U = 20, R3_THREADS = 384, R3_GT = 192;

template <int GATES, int S>
__device__ __forceinline__ void r3_chunk(const float* hs, const float* ws, float (&acc)[GATES][2][4]) {
#pragma unroll 2
  for (int kv = 0; kv < R2_K; kv += 4) {
    float4 hv[4], wv[GATES][S];
#pragma unroll
    for (int i = 0; i < 4; ++i) hv[i] = *reinterpret_cast<const float4*>(hs + i * 8 * R2_LD + kv);
#pragma unroll
    for (int g = 0; g < GATES; ++g)
#pragma unroll
      for (int s = 0; s < S; ++s) wv[g][s] = *reinterpret_cast<const float4*>(ws + (g * R3_U + s * 4) * R2_LD + kv);
#pragma unroll
    for (int g = 0; g < GATES; ++g)
#pragma unroll
      for (int s = 0; s < S; ++s)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float a = acc[g][s][i];
          a = fmaf(wv[g][s].x, hv[i].x, a);
          a = fmaf(wv[g][s].y, hv[i].y, a);
          a = fmaf(wv[g][s].z, hv[i].z, a);
          a = fmaf(wv[g][s].w, hv[i].w, a);
          acc[g][s][i] = a;
        }
  }
}

template <int GATES>
__global__ void __launch_bounds__(R3_THREADS, 1)
rnn_step_f32_v3_kernel(const float* __restrict__ gx, const float* __restrict__ w_hh, const float* __restrict__ b_hh,
                       const float* __restrict__ h_prev, float* __restrict__ h_next, float* __restrict__ c_state,
                       float* __restrict__ y, const int32_t* __restrict__ lens, int step, int B, int H, int dirs) {
  extern __shared__ __align__(16) float r3_smem[];
  constexpr int ROWS = GATES * R3_U;
  constexpr int STAGE = (R2_B + ROWS) * R2_LD;
  const int dir = blockIdx.y;
  const int j0 = blockIdx.x * R3_U;
  const int b0 = blockIdx.z * R2_B;
  if (step >= lens[b0]) return;   // lens sorted descending: row b0 is the longest of the tile
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool half_tile = warp >= 8;                       // units 16..19, one unit slot
  const int kg = half_tile ? ((warp - 8) >> 1) : (warp >> 2);
  const int wb = warp & 1;
  const int wu = half_tile ? 2 : ((warp & 3) >> 1);
  const int gt = (half_tile ? 4 + (warp & 1) : (warp & 3)) * 32 + lane;   // thread index inside the k-group
  const int bl = lane & 7, ul = lane >> 3;
  const int ubase = wu * 8 + ul;                          // unit of slot 0; slot 1 is ubase + 4
  const int nslots = half_tile ? 1 : 2;
  const float* hp = h_prev + (int64_t)dir * B * H;
  const float* wd = w_hh + (int64_t)dir * GATES * H * H;
  const float* bh = b_hh + (int64_t)dir * GATES * H;
  const int ncol = dirs * GATES * H;
  const int nk = (H + R2_K - 1) / R2_K;
  const int nkg = (nk - kg + 1) / 2;                       // chunks kc = 2 i + kg of this k-group
  float* ring = r3_smem + kg * (R2_STAGES * STAGE);

  auto issue = [&](int i) {
    float* hs = ring + (i % R2_STAGES) * STAGE;
    float* ws = hs + R2_B * R2_LD;
    const int k0 = (2 * i + kg) * R2_K;
    for (int v = gt; v < R2_B * (R2_K / 4); v += R3_GT) {
      const int row = v >> 3, kv = (v & 7) * 4;
      const bool ok = b0 + row < B && k0 + kv < H;
      cp_async16_zfill(hs + row * R2_LD + kv, ok ? hp + (int64_t)(b0 + row) * H + k0 + kv : hp, ok);
    }
    for (int v = gt; v < ROWS * (R2_K / 4); v += R3_GT) {
      const int r = v >> 3, kv = (v & 7) * 4;
      const int g = r / R3_U, j = j0 + (r % R3_U);
      const bool ok = j < H && k0 + kv < H;
      cp_async16_zfill(ws + r * R2_LD + kv, ok ? wd + ((int64_t)g * H + j) * H + k0 + kv : wd, ok);
    }
  };
#pragma unroll
  for (int i = 0; i < R2_STAGES - 1; ++i) {
    if (i < nkg) issue(i);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }

  // k-group 0 finishes the step: its gate pre-activations, biases and previous state are fetched under the k loop
  float gxv[GATES][2][4], bhv[GATES][2], hpv[2][4], cpv[2][4];
  int tt[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int b = b0 + wb * 32 + i * 8 + bl;
    const int len = (kg == 0 && b < B) ? lens[b] : 0;
    tt[i] = step < len ? (dir == 0 ? step : len - 1 - step) : -1;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int j = j0 + ubase + s * 4;
      const bool ok = tt[i] >= 0 && s < nslots && j < H;
      const float* gxr = gx + ((int64_t)(ok ? tt[i] : 0) * B + (ok ? b : 0)) * ncol + (int64_t)dir * GATES * H;
#pragma unroll
      for (int g = 0; g < GATES; ++g) gxv[g][s][i] = ok ? __ldg(gxr + (int64_t)g * H + j) : 0.0f;
      hpv[s][i] = ok ? hp[(int64_t)b * H + j] : 0.0f;
      cpv[s][i] = (ok && GATES == 4) ? c_state[((int64_t)dir * B + b) * H + j] : 0.0f;
    }
  }
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const int j = j0 + ubase + s * 4;
#pragma unroll
    for (int g = 0; g < GATES; ++g) bhv[g][s] = (kg == 0 && s < nslots && j < H) ? __ldg(bh + (int64_t)g * H + j) : 0.0f;
  }

  float acc[GATES][2][4];
#pragma unroll
  for (int g = 0; g < GATES; ++g)
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[g][s][i] = 0.0f;

  for (int i = 0; i < nkg; ++i) {
    asm volatile("cp.async.wait_group %0;" ::"n"(R2_STAGES - 2) : "memory");
    asm volatile("bar.sync %0, %1;" ::"r"(1 + kg), "n"(R3_GT) : "memory");   // chunk i landed for the whole k-group
    if (i + R2_STAGES - 1 < nkg) issue(i + R2_STAGES - 1);
    asm volatile("cp.async.commit_group;" ::: "memory");
    const float* hs = ring + (i % R2_STAGES) * STAGE + (wb * 32 + bl) * R2_LD;
    const float* ws = ring + (i % R2_STAGES) * STAGE + R2_B * R2_LD + ubase * R2_LD;
    if (half_tile) r3_chunk<GATES, 1>(hs, ws, acc);
    else r3_chunk<GATES, 2>(hs, ws, acc);
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();   // both rings are idle: k-group 1 hands its partial sums over through its own ring
  float* red = r3_smem + R2_STAGES * STAGE;
  if (kg == 1) {
#pragma unroll
    for (int g = 0; g < GATES; ++g)
#pragma unroll
      for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int i = 0; i < 4; ++i) red[((g * 2 + s) * 4 + i) * R3_GT + gt] = acc[g][s][i];
  }
  __syncthreads();
  if (kg == 1) return;
#pragma unroll
  for (int g = 0; g < GATES; ++g)
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[g][s][i] += red[((g * 2 + s) * 4 + i) * R3_GT + gt];

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (tt[i] < 0) continue;
    const int b = b0 + wb * 32 + i * 8 + bl;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int j = j0 + ubase + s * 4;
      if (s >= nslots || j >= H) continue;
      const int64_t sidx = ((int64_t)dir * B + b) * H + j;
      float hnew;
      if (GATES == 3) {
        const float r = sigmoid_acc(gxv[0][s][i] + acc[0][s][i] + bhv[0][s]);
        const float z = sigmoid_acc(gxv[1 % GATES][s][i] + acc[1 % GATES][s][i] + bhv[1 % GATES][s]);
        const float n = tanhf(gxv[2 % GATES][s][i] + r * (acc[2 % GATES][s][i] + bhv[2 % GATES][s]));
        hnew = (1.0f - z) * n + z * hpv[s][i];
      } else if (GATES == 4) {
        const float ig = sigmoid_acc(gxv[0][s][i] + acc[0][s][i] + bhv[0][s]);
        const float fg = sigmoid_acc(gxv[1 % GATES][s][i] + acc[1 % GATES][s][i] + bhv[1 % GATES][s]);
        const float gg = tanhf(gxv[2 % GATES][s][i] + acc[2 % GATES][s][i] + bhv[2 % GATES][s]);
        const float og = sigmoid_acc(gxv[3 % GATES][s][i] + acc[3 % GATES][s][i] + bhv[3 % GATES][s]);
        const float c = fg * cpv[s][i] + ig * gg;
        c_state[sidx] = c;
        hnew = og * tanhf(c);
      } else {
        hnew = tanhf(gxv[0][s][i] + acc[0][s][i] + bhv[0][s]);
      }
      h_next[sidx] = hnew;
      float* yo = y + ((int64_t)tt[i] * B + b) * H + j;
      if (dirs == 2) atomicAdd(yo, hnew);   // two addends onto 0: order-independent, exact
      else *yo = hnew;
    }
  }
}

template <int GATES>
static bool launch_step_v3(dim3 grid, cudaStream_t st, const float* gx, const float* w_hh, const float* b_hh,
                           const float* hp, float* hn, float* c_state, float* y, const int32_t* lens, int step, int B,
                           int H, int dirs) {
  constexpr int BYTES = 2 * R2_STAGES * (R2_B + GATES * R3_U) * R2_LD * (int)sizeof(float);
  static_assert(GATES * 2 * 4 * R3_GT <= R2_STAGES * (R2_B + GATES * R3_U) * R2_LD, "partial sums must fit one ring");
  static bool configured = false;
  static int configured_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (!configured || configured_dev != dev) {
    if (cudaFuncSetAttribute(rnn_step_f32_v3_kernel<GATES>, cudaFuncAttributeMaxDynamicSharedMemorySize, BYTES) != cudaSuccess)
      return false;
    configured = true;
    configured_dev = dev;
  }
  rnn_step_f32_v3_kernel<GATES><<<grid, R3_THREADS, BYTES, st>>>(gx, w_hh, b_hh, hp, hn, c_state, y, lens, step, B, H, dirs);
  return true;
}

template <int GATES>
static bool launch_step_v2(dim3 grid, cudaStream_t st, const float* gx, const float* w_hh, const float* b_hh,
                           const float* hp, float* hn, float* c_state, float* y, const int32_t* lens, int step, int B,
                           int H, int dirs) {
  constexpr int BYTES = R2_STAGES * (R2_B + GATES * R2_U) * R2_LD * (int)sizeof(float);
  static bool configured = false;   // one attribute call per instantiation (per process; the attribute is per device
  static int configured_dev = -1;   // context, so re-issue it when the current device changes)
  int dev = 0;
  cudaGetDevice(&dev);
  if (!configured || configured_dev != dev) {
    if (cudaFuncSetAttribute(rnn_step_f32_v2_kernel<GATES>, cudaFuncAttributeMaxDynamicSharedMemorySize, BYTES) != cudaSuccess)
      return false;
    configured = true;
    configured_dev = dev;
  }
  rnn_step_f32_v2_kernel<GATES><<<grid, R2_THREADS, BYTES, st>>>(gx, w_hh, b_hh, hp, hn, c_state, y, lens, step, B, H, dirs);
  return true;
}

static bool step_v2_ok(const RnnLayer& L, const float* h_state) {
  static const bool off = [] { const char* e = getenv("DSB_RNN_F32_V1"); return e && e[0] == '1'; }();
  return !off && (L.H % 4) == 0 && ((reinterpret_cast<uintptr_t>(L.w_hh) | reinterpret_cast<uintptr_t>(h_state)) & 15) == 0;
}

// one recurrence step of a layer: the v2 kernel when the layer qualifies, the general kernel otherwise
static void launch_step(const RnnLayer& L, bool v2, cudaStream_t st, const float* gx, const float* hp, float* hn,
                        float* c_state, float* y, const int32_t* lens, int step, int B, int dirs) {
  const int H = L.H;
  static const bool v3 = [] { const char* e = getenv("DSB_RNN_F32_V3"); return !(e && e[0] == '0'); }();
  if (v2 && v3) {
    dim3 grid(cdiv(H, R3_U), dirs, cdiv(B, R2_B));
    bool ok = L.gates == 3   ? launch_step_v3<3>(grid, st, gx, L.w_hh, L.b_hh, hp, hn, c_state, y, lens, step, B, H, dirs)
              : L.gates == 4 ? launch_step_v3<4>(grid, st, gx, L.w_hh, L.b_hh, hp, hn, c_state, y, lens, step, B, H, dirs)
                             : launch_step_v3<1>(grid, st, gx, L.w_hh, L.b_hh, hp, hn, c_state, y, lens, step, B, H, dirs);
    if (ok) return;
  }
  if (v2) {
    dim3 grid(cdiv(H, R2_U), dirs, cdiv(B, R2_B));
    bool ok = L.gates == 3   ? launch_step_v2<3>(grid, st, gx, L.w_hh, L.b_hh, hp, hn, c_state, y, lens, step, B, H, dirs)
              : L.gates == 4 ? launch_step_v2<4>(grid, st, gx, L.w_hh, L.b_hh, hp, hn, c_state, y, lens, step, B, H, dirs)
                             : launch_step_v2<1>(grid, st, gx, L.w_hh, L.b_hh, hp, hn, c_state, y, lens, step, B, H, dirs);
    if (ok) return;
  }
  dim3 grid(cdiv(H, RJ), dirs, cdiv(B, RB));
  if (L.gates == 3)
    rnn_step_f32_kernel<3><<<grid, 256, 0, st>>>(gx, L.w_hh, L.b_hh, hp, hn, c_state, y, lens, step, B, H, dirs);
  else if (L.gates == 4)
    rnn_step_f32_kernel<4><<<grid, 256, 0, st>>>(gx, L.w_hh, L.b_hh, hp, hn, c_state, y, lens, step, B, H, dirs);
  else
    rnn_step_f32_kernel<1><<<grid, 256, 0, st>>>(gx, L.w_hh, L.b_hh, hp, hn, c_state, y, lens, step, B, H, dirs);
}

int rnn_layer_f32(const dsb_model* m, const RnnLayer& L, const float* gates_x, const int32_t* d_len, int B, int Tmax,
                  int Trows, float* y, float* h_state, float* c_state, cudaStream_t st) {
  const int H = L.H, dirs = L.dirs;
  const size_t hbytes = sizeof(float) * (size_t)dirs * B * H;
  DSB_CUDA(cudaMemsetAsync(h_state, 0, 2 * hbytes, st));
  if (L.gates == 4) DSB_CUDA(cudaMemsetAsync(c_state, 0, hbytes, st));
  DSB_CUDA(cudaMemsetAsync(y, 0, sizeof(float) * (size_t)Trows * B * H, st));
  float* hbuf[2] = {h_state, h_state + (size_t)dirs * B * H};
  const bool v2 = step_v2_ok(L, h_state) && ((hbytes & 15) == 0);
  for (int s = 0; s < Tmax; ++s) {
    launch_step(L, v2, st, gates_x, hbuf[s & 1], hbuf[(s + 1) & 1], c_state, y, d_len, s, B, dirs);
    count_launch();
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(DSB_ERR_CUDA, "rnn step launch failed: %s", cudaGetErrorString(e));
  (void)m;
  return 0;
}


// Uni-directional layer with carried state (BatchRNNStream, model.py:219-237): h_io/c_io [B][H] hold the
// state between chunks; has_state = false starts from zeros (first chunk of an utterance).
int rnn_layer_f32_state(const RnnLayer& L, const float* gates_x, const int32_t* d_len, int B, int T, float* y,
                        float* h_scratch, float* c_scratch, float* h_io, float* c_io, bool has_state, cudaStream_t st) {
  (void)c_scratch;
  const int H = L.H;
  const size_t hbytes = sizeof(float) * (size_t)B * H;
  if (L.dirs != 1) return set_error(DSB_ERR_UNSUPPORTED, "rnn_layer_f32_state: uni-directional layers only");
  if (has_state) {
    DSB_CUDA(cudaMemcpyAsync(h_scratch, h_io, hbytes, cudaMemcpyDeviceToDevice, st));
  } else {
    DSB_CUDA(cudaMemsetAsync(h_scratch, 0, hbytes, st));
    if (L.gates == 4) DSB_CUDA(cudaMemsetAsync(c_io, 0, hbytes, st));
  }
  float* hbuf[2] = {h_scratch, h_scratch + (size_t)B * H};
  const bool v2 = step_v2_ok(L, h_scratch) && ((hbytes & 15) == 0);
  for (int s = 0; s < T; ++s) {
    launch_step(L, v2, st, gates_x, hbuf[s & 1], hbuf[(s + 1) & 1], c_io, y, d_len, s, B, 1);
    count_launch();
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(DSB_ERR_CUDA, "rnn step launch failed: %s", cudaGetErrorString(e));
  DSB_CUDA(cudaMemcpyAsync(h_io, hbuf[T & 1], hbytes, cudaMemcpyDeviceToDevice, st));
  return 0;
}

}  // namespace dsb
