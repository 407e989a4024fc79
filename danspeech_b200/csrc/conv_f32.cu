// fp32 (CUDA-core) MaskConv block: Conv2d + folded eval-BatchNorm2d + Hardtanh(0,20) + length mask.
//
// Replaces one (Conv2d, BatchNorm2d, Hardtanh) triple of danspeech/deepspeech/model.py:357-392 together
// with the masking loop of MaskConv.forward (model.py:65-81).  This is the exact-fp32 path used by
// DSB_PREC_FP32; the bf16 tensor-core path lives in conv_tc.cu.
//
// Direct convolution: a CTA produces 32 output channels x 128 output frames of one (utterance,
// output frequency row).  For every kernel row it stages an 8-channel slab of the input row and the
// matching [8][11][32] weights in shared memory; each thread accumulates 8 channels x 4 frames.
#include "model_types.cuh"

namespace dsb {

constexpr int CV_TT = 128;   // output frames per CTA
constexpr int CV_CO = 32;    // output channels per CTA
constexpr int CV_CI = 8;     // input channels per staged slab

template <int ST>
__global__ void __launch_bounds__(128)
conv_f32_kernel(const float* __restrict__ x, int cin, int din, int tin, const float* __restrict__ w /*[kh][cin][11][cout]*/,
                const float* __restrict__ bias, int cout, int kh_n, int sd, int pd, int dout, int tout,
                const int32_t* __restrict__ lens, float* __restrict__ y, int rnn_layout, int B) {
  constexpr int XW = ST * (CV_TT - 1) + kConvKW;
  __shared__ float xs[CV_CI][XW + 1];
  __shared__ __align__(16) float ws[CV_CI][kConvKW][CV_CO];

  const int b = blockIdx.z;
  const int co_tiles = cout / CV_CO;
  const int d = blockIdx.y / co_tiles;
  const int co0 = (blockIdx.y % co_tiles) * CV_CO;
  const int t0 = blockIdx.x * CV_TT;
  const int tid = threadIdx.x, lane = tid & 31, cg = tid >> 5;

  float acc[8][4];
#pragma unroll
  for (int c = 0; c < 8; ++c)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[c][i] = 0.0f;

  const int tin0 = ST * t0 - kConvPT;
  for (int kh = 0; kh < kh_n; ++kh) {
    const int row = sd * d + kh - pd;
    if (row < 0 || row >= din) continue;     // zero padding in frequency (uniform over the CTA)
    for (int ci0 = 0; ci0 < cin; ci0 += CV_CI) {
      const int nci = min(CV_CI, cin - ci0);
      __syncthreads();
      for (int i = tid; i < nci * XW; i += 128) {
        int ci = i / XW, p = i - ci * XW;
        int t = tin0 + p;
        float v = 0.0f;
        if (t >= 0 && t < tin) v = __ldg(x + (((int64_t)b * cin + ci0 + ci) * din + row) * tin + t);
        xs[ci][p] = v;
      }
      for (int i = tid; i < nci * kConvKW * CV_CO; i += 128) {
        int co = i % CV_CO;
        int r = i / CV_CO;
        int kw = r % kConvKW, ci = r / kConvKW;
        ws[ci][kw][co] = __ldg(w + (((int64_t)kh * cin + ci0 + ci) * kConvKW + kw) * cout + co0 + co);
      }
      __syncthreads();
      for (int ci = 0; ci < nci; ++ci) {
#pragma unroll
        for (int kw = 0; kw < kConvKW; ++kw) {
          float xv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) xv[i] = xs[ci][ST * (lane + 32 * i) + kw];
          const float4 w0 = *reinterpret_cast<const float4*>(&ws[ci][kw][cg * 8]);
          const float4 w1 = *reinterpret_cast<const float4*>(&ws[ci][kw][cg * 8 + 4]);
          const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
          for (int c = 0; c < 8; ++c)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[c][i] = fmaf(wv[c], xv[i], acc[c][i]);
        }
      }
    }
  }

  const int len = lens[b];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int co = co0 + cg * 8 + c;
    const float bv = bias[co];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int t = t0 + lane + 32 * i;
      if (t >= tout) continue;
      float v = fminf(fmaxf(acc[c][i] + bv, 0.0f), 20.0f);
      if (t >= len) v = 0.0f;
      if (rnn_layout)
        y[((int64_t)t * B + b) * ((int64_t)cout * dout) + (int64_t)co * dout + d] = v;
      else
        y[(((int64_t)b * cout + co) * dout + d) * tout + t] = v;
    }
  }
}

int conv2d_bn_htanh_f32(const float* x, int B, int cin, int din, int tin, const ConvLayer& L, const int32_t* d_len,
                        float* y, int tout, bool rnn_layout, cudaStream_t st) {
  if (L.cout % CV_CO != 0) return set_error(DSB_ERR_UNSUPPORTED, "conv: cout %d not a multiple of %d", L.cout, CV_CO);
  dim3 grid(cdiv(tout, CV_TT), L.dout * (L.cout / CV_CO), B);
  if (L.st == 2)
    conv_f32_kernel<2><<<grid, 128, 0, st>>>(x, cin, din, tin, L.w, L.bias, L.cout, L.kh, L.sd, L.pd, L.dout, tout,
                                            d_len, y, rnn_layout ? 1 : 0, B);
  else
    conv_f32_kernel<1><<<grid, 128, 0, st>>>(x, cin, din, tin, L.w, L.bias, L.cout, L.kh, L.sd, L.pd, L.dout, tout,
                                            d_len, y, rnn_layout ? 1 : 0, B);
  DSB_CHECK_LAUNCH();
  return 0;
}

}  // namespace dsb
