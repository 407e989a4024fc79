// Model tail + greedy CTC kernels.
//
//  * lookahead_htanh_f32  -- Lookahead.forward + Hardtanh(0,20)  (model.py:143-148, :407-411)
//  * softmax_argmax_f32   -- InferenceBatchSoftmax (model.py:89-93) fused with the [T,B,C]->[B,T,C]
//                            transpose of model.py:512 and with torch.max(probs, 2) (decoder.py:195)
//  * dsb_greedy_decode    -- GreedyDecoder.decode/process_string (decoder.py:166-198): warp-level
//                            collapse-repeats / drop-blank compaction with ballot + popc.
#include "model_types.cuh"

namespace dsb {

__global__ void lookahead_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ y,
                                 int T, int BH, int H, int context) {
  const int64_t total = (int64_t)T * BH;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int t = (int)(i / BH);
    const int bc = (int)(i - (int64_t)t * BH);
    const int c = bc % H;
    float acc = 0.0f;
    for (int j = 0; j < context && t + j < T; ++j) acc = fmaf(__ldg(w + c * context + j), x[i + (int64_t)j * BH], acc);
    y[i] = fminf(fmaxf(acc, 0.0f), 20.0f);
  }
}

// Sliding-window variant for the usual context of 20: one thread per (batch, channel) column walks the time axis
// with the window in registers, so every input element is read once (the generic kernel reads it `context` times).
template <int CTX>
__global__ void lookahead_window_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ y,
                                        int T, int BH, int H, int t_chunk) {
  const int bc = blockIdx.x * blockDim.x + threadIdx.x;
  if (bc >= BH) return;
  const int t0 = blockIdx.y * t_chunk, t1 = min(T, t0 + t_chunk);
  const int c = bc % H;
  float wt[CTX], win[CTX];
#pragma unroll
  for (int j = 0; j < CTX; ++j) {
    wt[j] = __ldg(w + c * CTX + j);
    win[j] = (j > 0 && t0 + j - 1 < T) ? x[(int64_t)(t0 + j - 1) * BH + bc] : 0.f;   // win[j] = x[t + j - 1] before the shift
  }
  for (int t = t0; t < t1; ++t) {
#pragma unroll
    for (int j = 0; j < CTX - 1; ++j) win[j] = win[j + 1];
    win[CTX - 1] = t + CTX - 1 < T ? x[(int64_t)(t + CTX - 1) * BH + bc] : 0.f;
    float acc = 0.0f;
#pragma unroll
    for (int j = 0; j < CTX; ++j) acc = fmaf(wt[j], win[j], acc);
    y[(int64_t)t * BH + bc] = fminf(fmaxf(acc, 0.0f), 20.0f);
  }
}

int lookahead_htanh_f32(const float* x, const float* w, float* y, int T, int B, int H, int context, cudaStream_t st) {
  const int BH = B * H;
  if (context == 20) {
    // enough (column, time-chunk) threads to fill the GPU; a chunk re-reads context-1 frames at its start
    int chunks = 1;
    while ((int64_t)BH * chunks < 148 * 2048 && chunks * 64 < T) chunks *= 2;
    const int t_chunk = cdiv(T, chunks);
    dim3 grid(cdiv(BH, 128), cdiv(T, t_chunk));
    lookahead_window_kernel<20><<<grid, 128, 0, st>>>(x, w, y, T, BH, H, t_chunk);
    DSB_CHECK_LAUNCH();
    return 0;
  }
  const int64_t total = (int64_t)T * B * H;
  int blocks = (int)(cdiv64(total, 256) < 148 * 8 ? cdiv64(total, 256) : 148 * 8);
  lookahead_kernel<<<blocks, 256, 0, st>>>(x, w, y, T, BH, H, context);
  DSB_CHECK_LAUNCH();
  return 0;
}

// One warp per (t, b) row.  probs[b][t][:] = softmax(logits[t*B+b][:]); argmax over the *rounded*
// probabilities, first index wins ties (torch.max semantics on the tensor the reference sees).
__global__ void softmax_argmax_kernel(const float* __restrict__ logits, float* __restrict__ probs,
                                      int32_t* __restrict__ argmax, int T, int B, int C) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= T * B) return;
  const int t = row / B, b = row - t * B;
  const float* in = logits + (int64_t)row * C;
  float mx = -INFINITY;
  for (int c = lane; c < C; c += 32) mx = fmaxf(mx, in[c]);
  mx = warp_max(mx);
  float sum = 0.0f;
  for (int c = lane; c < C; c += 32) sum += expf(in[c] - mx);
  sum = warp_sum(sum);
  float* out = probs + ((int64_t)b * T + t) * C;
  float best = -1.0f;
  int besti = 0x7fffffff;
  for (int c = lane; c < C; c += 32) {
    const float p = expf(in[c] - mx) / sum;
    out[c] = p;
    if (p > best) { best = p; besti = c; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
    if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
  }
  if (argmax && lane == 0) argmax[(int64_t)b * T + t] = besti;
}

int softmax_argmax_f32(const float* logits, float* probs, int32_t* argmax, int T, int B, int C, cudaStream_t st) {
  const int rows = T * B;
  softmax_argmax_kernel<<<cdiv(rows, 8), 256, 0, st>>>(logits, probs, argmax, T, B, C);
  DSB_CHECK_LAUNCH();
  return 0;
}

// SequenceWise fc (BatchNorm1d folded into W, b; model.py:414-420) fused with InferenceBatchSoftmax, the transpose
// to [B,T,C] and the argmax: the C x H weight matrix lives in shared memory for the whole (persistent) CTA, a warp
// takes R rows at a time, every lane accumulates 4 consecutive k per 128 for all classes (one 128-bit shared load
// feeds 4R FMAs; 16 warps of ~110 registers hide the load latencies better than 8 warps with R = 4 did), then the R*C sums are reduced across the warp and the softmax is done from a small staging
// area.  HBM-bound in principle (H*4 bytes per row); replaces a narrow-N GEMM plus a second pass over the logits.
constexpr int FC_WARPS = 16;
template <int NC, int R>
__global__ void __launch_bounds__(FC_WARPS * 32, 1)
fc_softmax_argmax_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ bias,
                         float* __restrict__ probs, int32_t* __restrict__ argmax, int T, int B, int C, int H, int Hp) {
  extern __shared__ float fc_smem[];
  float* sW = fc_smem;                          // [NC][Hp], zero padded
  float* sL = fc_smem + (size_t)NC * Hp;        // [FC_WARPS][R][NC]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < NC * Hp; i += blockDim.x) {
    const int c = i / Hp, k = i - c * Hp;
    sW[i] = (c < C && k < H) ? W[(size_t)c * H + k] : 0.f;
  }
  __syncthreads();
  const int64_t rows = (int64_t)T * B;
  const int64_t n_groups = (rows + R - 1) / R;
  float* myL = sL + warp * R * NC;
  for (int64_t grp = (int64_t)blockIdx.x * FC_WARPS + warp; grp < n_groups; grp += (int64_t)gridDim.x * FC_WARPS) {
    const int64_t row0 = grp * R;
    float acc[R][NC];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int c = 0; c < NC; ++c) acc[r][c] = 0.f;
    // the x loads of the next 128-wide slice are issued before the FMAs of the current one (two warps per
    // scheduler cannot hide a DRAM round trip otherwise)
    auto load_x = [&](int k0, float4 (&xv)[R]) {
#pragma unroll
      for (int r = 0; r < R; ++r)
        xv[r] = (k0 < H && row0 + r < rows) ? __ldg(reinterpret_cast<const float4*>(x + (row0 + r) * H + k0))
                                            : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    float4 xn[R];
    load_x(lane * 4, xn);
    for (int k0 = lane * 4; k0 < H; k0 += 128) {
      float4 xv[R];
#pragma unroll
      for (int r = 0; r < R; ++r) xv[r] = xn[r];
      load_x(k0 + 128, xn);
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const float4 w = *reinterpret_cast<const float4*>(sW + (size_t)c * Hp + k0);
#pragma unroll
        for (int r = 0; r < R; ++r)
          acc[r][c] = fmaf(xv[r].x, w.x, fmaf(xv[r].y, w.y, fmaf(xv[r].z, w.z, fmaf(xv[r].w, w.w, acc[r][c]))));
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const float v = warp_sum(acc[r][c]);
        if (lane == 0) myL[r * NC + c] = v + ((bias && c < C) ? __ldg(bias + c) : 0.f);
      }
    __syncwarp();
    for (int r = 0; r < R && row0 + r < rows; ++r) {
      const int64_t row = row0 + r;
      const int t = (int)(row / B), b = (int)(row - (int64_t)t * B);
      const float* in = myL + r * NC;
      float mx = -INFINITY;
      for (int c = lane; c < C; c += 32) mx = fmaxf(mx, in[c]);
      mx = warp_max(mx);
      float sum = 0.0f;
      for (int c = lane; c < C; c += 32) sum += expf(in[c] - mx);
      sum = warp_sum(sum);
      float* out = probs + ((int64_t)b * T + t) * C;
      float best = -1.0f;
      int besti = 0x7fffffff;
      for (int c = lane; c < C; c += 32) {
        const float pv = expf(in[c] - mx) / sum;
        out[c] = pv;
        if (pv > best) { best = pv; besti = c; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
        if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
      }
      if (argmax && lane == 0) argmax[(int64_t)b * T + t] = besti;
    }
    __syncwarp();
  }
}

// x [T*B][H] fp32 -> probs [B][T][C] (+ argmax).  Uses the fused kernel when C <= 33, H % 4 == 0 and the weights fit
// in shared memory; otherwise the fp32 GEMM into `logits_scratch` [T*B][C] followed by softmax_argmax_f32.
int fc_softmax_argmax_f32(const float* x, const float* W, const float* bias, float* probs, int32_t* argmax,
                          float* logits_scratch, int T, int B, int C, int H, cudaStream_t st) {
  constexpr int NC = 33, R = 2;
  const int Hp = (H + 127) / 128 * 128;
  const size_t smem = ((size_t)NC * Hp + FC_WARPS * R * NC) * sizeof(float);
  if (C <= NC && (H & 3) == 0 && smem <= 227 * 1024 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    // per device (and a cheap host-side call): set on every launch, not once per process
    DSB_CUDA(cudaFuncSetAttribute(fc_softmax_argmax_kernel<NC, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    const int64_t groups = ((int64_t)T * B + R - 1) / R;
    const int grid = (int)(cdiv64(groups, FC_WARPS) < 148 ? cdiv64(groups, FC_WARPS) : 148);
    fc_softmax_argmax_kernel<NC, R><<<grid, FC_WARPS * 32, smem, st>>>(x, W, bias, probs, argmax, T, B, C, H, Hp);
    DSB_CHECK_LAUNCH();
    return 0;
  }
  if (int e = gemm_bias_f32(x, W, bias, logits_scratch, (int64_t)T * B, C, H, st)) return e;
  return softmax_argmax_f32(logits_scratch, probs, argmax, T, B, C, st);
}

// One warp per utterance.
__global__ void greedy_kernel(const float* __restrict__ probs, const int32_t* __restrict__ argmax,
                              const int32_t* __restrict__ sizes, int B, int T, int C, int blank,
                              int32_t* __restrict__ tokens, int32_t* __restrict__ offsets,
                              int32_t* __restrict__ out_len) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  int size = sizes ? sizes[b] : T;
  size = min(max(size, 0), T);
  int count = 0;
  int carry = -1;   // symbol of frame base-1
  for (int base = 0; base < size; base += 32) {
    const int t = base + lane;
    int s = blank;
    if (t < size) {
      if (argmax) {
        s = argmax[(int64_t)b * T + t];
      } else {
        const float* p = probs + ((int64_t)b * T + t) * C;
        float best = p[0];
        s = 0;
        for (int c = 1; c < C; ++c) {
          const float v = p[c];
          if (v > best) { best = v; s = c; }
        }
      }
    }
    int prev = __shfl_up_sync(0xffffffffu, s, 1);
    if (lane == 0) prev = carry;
    const bool keep = (t < size) && (s != blank) && !(t > 0 && s == prev);
    const unsigned mask = __ballot_sync(0xffffffffu, keep);
    if (keep) {
      const int pos = count + __popc(mask & ((1u << lane) - 1u));
      tokens[(int64_t)b * T + pos] = s;
      offsets[(int64_t)b * T + pos] = t;
    }
    count += __popc(mask);
    carry = __shfl_sync(0xffffffffu, s, 31);
  }
  if (lane == 0) out_len[b] = count;
}

}  // namespace dsb

using namespace dsb;

extern "C" int dsb_greedy_decode(const float* probs, const int32_t* argmax, const int32_t* sizes, int B, int T, int C,
                                 int blank, int32_t* tokens, int32_t* offsets, int32_t* out_len, void* stream) {
  DSB_REQUIRE((probs || argmax) && tokens && offsets && out_len, "dsb_greedy_decode: null argument");
  DSB_REQUIRE(B > 0 && T >= 0 && C > 0, "dsb_greedy_decode: bad shape B=%d T=%d C=%d", B, T, C);
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope scope(ST_GREEDY, st);
  greedy_kernel<<<cdiv(B, 4), 128, 0, st>>>(probs, argmax, sizes, B, T, C, blank, tokens, offsets, out_len);
  DSB_CHECK_LAUNCH();
  return 0;
}
