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

int lookahead_htanh_f32(const float* x, const float* w, float* y, int T, int B, int H, int context, cudaStream_t st) {
  const int64_t total = (int64_t)T * B * H;
  int blocks = (int)(cdiv64(total, 256) < 148 * 8 ? cdiv64(total, 256) : 148 * 8);
  lookahead_kernel<<<blocks, 256, 0, st>>>(x, w, y, T, B * H, H, context);
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
