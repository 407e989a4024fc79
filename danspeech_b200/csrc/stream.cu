// Chunked streaming forward for S lock-step streams.
//
// Replaces DeepSpeech.streaming_forward (danspeech/deepspeech/model.py:517-537) and the stateful
// modules MaskConvStream (:169-201), BatchRNNStream (:219-237) and LookaheadStream (:255-279).  The
// per-module-instance state of the reference (left conv contexts, GRU hidden state, lookahead tail)
// becomes an explicit device-resident dsb_stream_state for S streams, so one model serves many
// concurrent streams.  Quirks kept on purpose: the convs still apply their own symmetric
// time padding on top of the carried context (overlapping frames are re-emitted), the first call only
// fills the lookahead buffer and emits nothing, the hidden state is reset on is_last only.
// DSB_PREC_FP32 runs the exact-fp32 kernels; DSB_PREC_BF16 runs the tcgen05 kernels: the S streams are laid
// side by side on the time axis (conv 1: [161][S*t1] rows of the time-expanded input; conv 2: one segment
// [5 zeros | carried context + chunk | 5 zeros] per stream, so that 128-frame MMA tiles stay full although a
// chunk is only ~40 frames), the projection is one GEMM over all S*T2 rows and the recurrence runs the
// streams in groups of 128 on independent CTA sets with the hidden state carried in fp32.
#include "model_types.cuh"

struct dsb_stream_state {
  int S = 0, max_k = 0;
  int t1_max = 0, t2_max = 0, look_cap = 0;
  float* left1 = nullptr;   // [S][161][10]      last 10 frames of the conv-1 input
  float* left2 = nullptr;   // [S][32][81][10]   last 10 frames of the conv-2 input
  float* h = nullptr;       // [layers][S][H]
  float* c = nullptr;       // [layers][S][H] (LSTM)
  float* look = nullptr;    // [look_cap][S][H]  frames waiting for their right context
  int look_len = 0;
  bool look_init = false, h_init = false;
  // scratch
  float *in1 = nullptr, *c1 = nullptr, *in2 = nullptr, *x0 = nullptr, *gates = nullptr, *ya = nullptr, *yb = nullptr,
        *cat = nullptr, *lo = nullptr, *logits = nullptr, *hs = nullptr, *cs = nullptr;
  int32_t* lens = nullptr;
  // bf16 tensor-core path
  bool tc = false;
  __nv_bfloat16 *left2b = nullptr;   // [81][S][10][32] last 10 frames of the conv-2 input (channels-last)
  __nv_bfloat16 *x1 = nullptr, *c1b = nullptr, *in2b = nullptr, *xb = nullptr, *hbuf = nullptr;
  unsigned int* sync_words = nullptr;
  std::vector<void*> owned;
};

namespace dsb {

// dst[s][c][d][t] over t in [0, Tout): [pad_l zeros | left (n_left frames, if any) | x (Tx frames) | zeros]
__global__ void assemble_kernel(const float* __restrict__ x, int Tx, const float* __restrict__ left, int n_left,
                                int pad_l, float* __restrict__ dst, int Tout, int64_t rows) {
  const int64_t total = rows * Tout;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int t = (int)(i % Tout);
    const int64_t r = i / Tout;
    float v = 0.f;
    int u = t - pad_l;
    if (u >= 0) {
      if (u < n_left) v = left[r * n_left + u];
      else if (u - n_left < Tx) v = x[r * Tx + (u - n_left)];
    }
    dst[i] = v;
  }
}
// left[r][0..n) = src[r][T-n .. T)
__global__ void save_tail_kernel(const float* __restrict__ src, int T, float* __restrict__ left, int n, int64_t rows) {
  const int64_t total = rows * n;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int u = (int)(i % n);
    const int64_t r = i / n;
    left[i] = src[r * T + (T - n + u)];
  }
}
// ---- bf16 path helpers (channels-last, streams side by side on the time axis) ----
// Block-1 operand tiles for the streams laid side by side in time (frame u = s*t1 + t of one long "utterance"):
// x1[u][j] = in1[s][d][2t + j - 5] (j < 11; zero outside [0, tin1) and for j >= 11), stored as
// [u / 128][d][128 frames x 32 B, 32B-swizzled] like conv_tc.cu's im2col_time_kernel; frames >= S*t1 are 0.
__global__ void im2col_time_stream_kernel(const float* __restrict__ in1, __nv_bfloat16* __restrict__ x1, int S, int D,
                                          int tin1, int t1, int t_blocks) {
  const int Upad = t_blocks * 128;
  const int64_t total = (int64_t)D * Upad;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int u = (int)(i % Upad);
    const int d = (int)(i / Upad);
    const int s = u / t1, t = u - s * t1;
    const bool in = s < S;
    const float* src = in1 + ((int64_t)(in ? s : 0) * D + d) * tin1;
    uint32_t w[8];
#pragma unroll
    for (int h = 0; h < 8; ++h) {
      const int j0 = 2 * h, j1 = 2 * h + 1;
      const int s0 = 2 * t + j0 - 5, s1 = 2 * t + j1 - 5;
      const float v0 = (in && j0 < kConvKW && s0 >= 0 && s0 < tin1) ? src[s0] : 0.f;
      const float v1 = (in && j1 < kConvKW && s1 >= 0 && s1 < tin1) ? src[s1] : 0.f;
      __nv_bfloat162 pk = __floats2bfloat162_rn(v0, v1);
      w[h] = *reinterpret_cast<uint32_t*>(&pk);
    }
    const int tb = u >> 7, tl = u & 127, sw = (tl >> 2) & 1;
    uint4* dst = reinterpret_cast<uint4*>(x1 + ((((int64_t)tb * D + d) << 7) + tl) * 16);
    dst[sw] = make_uint4(w[0], w[1], w[2], w[3]);
    dst[sw ^ 1] = make_uint4(w[4], w[5], w[6], w[7]);
  }
}
// in2[d][s*seg + u][c], u in [0, seg = tin2 + 10): 5 zeros, then the tin2 logical frames
// [pad_l zeros | left (n_left) | c1 (t1) | zeros], then 5 zeros.  One thread per 8 channels (16 bytes).
__global__ void assemble_cl_kernel(const __nv_bfloat16* __restrict__ c1, int t1, const __nv_bfloat16* __restrict__ left,
                                   int n_left, int pad_l, __nv_bfloat16* __restrict__ in2, int tin2, int S, int D) {
  const int seg = tin2 + 10;
  const int64_t total = (int64_t)D * S * seg * 4;
  const uint4* c1v = reinterpret_cast<const uint4*>(c1);
  const uint4* lv = reinterpret_cast<const uint4*>(left);
  uint4* out = reinterpret_cast<uint4*>(in2);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int part = (int)(i & 3);
    int64_t r = i >> 2;
    const int u = (int)(r % seg);
    r /= seg;   // d*S + s
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    const int w = u - 5 - pad_l;
    if (u - 5 >= 0 && u - 5 < tin2 && w >= 0) {
      if (w < n_left) v = lv[(r * 10 + w) * 4 + part];
      else if (w - n_left < t1) v = c1v[(r * t1 + (w - n_left)) * 4 + part];
    }
    out[i] = v;
  }
}
// left[d][s][u][c] = last 10 logical frames of in2's segment (u < 10)
__global__ void save_tail_cl_kernel(const __nv_bfloat16* __restrict__ in2, int tin2, __nv_bfloat16* __restrict__ left,
                                    int64_t rows) {
  const int seg = tin2 + 10;
  const int64_t total = rows * 10 * 4;
  const uint4* src = reinterpret_cast<const uint4*>(in2);
  uint4* dst = reinterpret_cast<uint4*>(left);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int part = (int)(i & 3);
    const int64_t r = (i >> 2) / 10;
    const int u = (int)((i >> 2) % 10);
    dst[i] = src[(r * seg + 5 + tin2 - 10 + u) * 4 + part];
  }
}

__global__ void fill_i32_kernel(int32_t* p, int n, int v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

static int sgrid(int64_t n) { return (int)(cdiv64(n, 256) < 148 * 16 ? cdiv64(n, 256) : 148 * 16); }

template <typename T>
static int salloc(dsb_stream_state* s, T** p, int64_t n) {
  void* q = nullptr;
  DSB_CUDA(cudaMalloc(&q, sizeof(T) * (size_t)(n > 0 ? n : 1)));
  DSB_CUDA(cudaMemset(q, 0, sizeof(T) * (size_t)(n > 0 ? n : 1)));
  s->owned.push_back(q);
  *p = reinterpret_cast<T*>(q);
  return 0;
}

int rnn_layer_f32_state(const RnnLayer& L, const float* gates_x, const int32_t* d_len, int B, int T, float* y,
                        float* h_scratch, float* c_scratch, float* h_io, float* c_io, bool has_state, cudaStream_t st);

}  // namespace dsb

using namespace dsb;

static void stream_frames(int k, int is_first, int is_last, int& tin1, int& t1, int& tin2) {
  tin1 = k + ((is_first || is_last) ? 5 : 0) + (is_first ? 0 : 10);
  t1 = (tin1 - 1) / 2 + 1;
  tin2 = t1 + ((is_first || is_last) ? 5 : 0) + (is_first ? 0 : 10);
}

extern "C" int dsb_stream_state_create(dsb_model* m, int n_streams, int max_chunk_frames, dsb_stream_state** out) {
  DSB_REQUIRE(m && out && n_streams > 0 && max_chunk_frames > 0, "dsb_stream_state_create: bad argument");
  if (!m->finalized) return set_error(DSB_ERR_STATE, "dsb_stream_state_create: model not finalized");
  if (!m->desc.streaming) return set_error(DSB_ERR_STATE, "dsb_stream_state_create: not a streaming model");
  if (m->convs.size() != 2)
    return set_error(DSB_ERR_UNSUPPORTED, "streaming models need exactly 2 conv blocks (the reference's streaming_init "
                                          "sizes the RNN for the 2-conv stack, model.py:477-484)");
  dsb_stream_state* s = new dsb_stream_state();
  s->S = n_streams;
  s->max_k = max_chunk_frames;
  const int S = n_streams, H = m->desc.rnn_hidden_size, C = m->desc.num_classes, ctx = m->desc.context;
  const int layers = (int)m->rnns.size(), G = m->rnns[0].gates;
  const int tin1 = max_chunk_frames + 15, t1 = (tin1 - 1) / 2 + 1, tin2 = t1 + 15;
  s->t1_max = t1;
  s->t2_max = tin2;
  s->look_cap = tin2 > ctx - 1 ? tin2 : ctx - 1;
  const int cat_max = s->look_cap + tin2;
  auto fail = [&](int e) {
    for (void* p : s->owned) cudaFree(p);
    delete s;
    return e;
  };
#define SA(ptr, n) if (int e = salloc(s, &s->ptr, (int64_t)(n))) return fail(e)
  s->tc = m->precision == DSB_PREC_BF16;
  SA(left1, (int64_t)S * kFreqBins * 10);
  if (s->tc) {
    const RnnLayer& R0 = m->rnns[0];
    const int ld = R0.in_ld > (H + 7) / 8 * 8 ? R0.in_ld : (H + 7) / 8 * 8;
    SA(left2b, (int64_t)81 * S * 10 * 32);
    SA(x1, (int64_t)conv1_tiles_elems(1, S * t1));
    SA(c1b, (int64_t)81 * S * t1 * 32);
    SA(in2b, (int64_t)81 * S * (tin2 + 10) * 32);
    SA(xb, (int64_t)tin2 * S * ld);
    SA(hbuf, (int64_t)rnn_tc_hbuf_elems(R0, S));
    SA(sync_words, kRnnSyncCounters + 1);
  } else {
    SA(left2, (int64_t)S * 32 * 81 * 10);
    SA(c1, (int64_t)S * 32 * 81 * t1);
    SA(in2, (int64_t)S * 32 * 81 * tin2);
    SA(x0, (int64_t)tin2 * S * m->rnn_input);
  }
  SA(h, (int64_t)layers * S * H);
  SA(c, (int64_t)layers * S * H);
  SA(look, (int64_t)s->look_cap * S * H);
  SA(in1, (int64_t)S * kFreqBins * tin1);
  SA(gates, (int64_t)tin2 * S * G * H);
  SA(ya, (int64_t)tin2 * S * H);
  SA(yb, (int64_t)tin2 * S * H);
  SA(cat, (int64_t)cat_max * S * H);
  SA(lo, (int64_t)cat_max * S * H);
  SA(logits, (int64_t)cat_max * S * C);
  SA(hs, (int64_t)2 * S * H);
  SA(cs, (int64_t)S * H);
  SA(lens, S);
#undef SA
  *out = s;
  return 0;
}

extern "C" void dsb_stream_state_destroy(dsb_stream_state* s) {
  if (!s) return;
  for (void* p : s->owned) cudaFree(p);
  delete s;
}

extern "C" int dsb_stream_max_out_frames(const dsb_stream_state* s, int k) {
  if (!s || k <= 0) return 0;
  int tin1, t1, tin2;
  stream_frames(k, 0, 1, tin1, t1, tin2);
  return s->look_cap + tin2;
}

extern "C" int dsb_streaming_forward(dsb_model* m, dsb_stream_state* s, const float* chunk, int k, int is_first,
                                     int is_last, float* probs, int32_t* k_out, void* stream) {
  DSB_REQUIRE(m && s && chunk && probs && k_out, "dsb_streaming_forward: null argument");
  DSB_REQUIRE(k >= 1 && k <= s->max_k, "dsb_streaming_forward: chunk of %d frames outside [1,%d]", k, s->max_k);
  if (!m->desc.streaming) return set_error(DSB_ERR_STATE, "dsb_streaming_forward: not a streaming model");
  cudaStream_t st = (cudaStream_t)stream;
  const int S = s->S, H = m->desc.rnn_hidden_size, C = m->desc.num_classes, ctx = m->desc.context;
  int tin1, t1, tin2;
  stream_frames(k, is_first, is_last, tin1, t1, tin2);
  const int T2 = tin2;   // conv-2 has unit time stride
  *k_out = 0;

  DSB_REQUIRE(tin1 >= 10 && tin2 >= 10, "dsb_streaming_forward: chunk of %d frames is too short to carry the 10-frame "
              "conv contexts", k);
  const int64_t rows1 = (int64_t)S * kFreqBins;
  const float* x = nullptr;   // output of the last recurrent layer, fp32 [T2][S][H]

  // ---- MaskConvStream (model.py:169-201): conv-1 input = [5 zeros if first | 10 carried frames | chunk | 5 zeros if last] ----
  prof_begin(ST_CONV, st);
  assemble_kernel<<<sgrid(rows1 * tin1), 256, 0, st>>>(chunk, k, s->left1, is_first ? 0 : 10, is_first ? 5 : 0, s->in1,
                                                     tin1, rows1);
  DSB_CHECK_LAUNCH();
  if (!is_last) {
    save_tail_kernel<<<sgrid(rows1 * 10), 256, 0, st>>>(s->in1, tin1, s->left1, 10, rows1);
    DSB_CHECK_LAUNCH();
  }
  if (s->tc) {
    const int64_t rows2 = (int64_t)81 * S;
    const int seg = tin2 + 10;
    const int t_blocks1 = cdiv(S * t1, 128);
    im2col_time_stream_kernel<<<sgrid((int64_t)kFreqBins * t_blocks1 * 128), 256, 0, st>>>(s->in1, s->x1, S, kFreqBins, tin1, t1,
                                                                                       t_blocks1);
    DSB_CHECK_LAUNCH();
    if (int e = conv_block_tc(s->x1, m->convs[0], true, nullptr, 1, S * t1, s->c1b, false, 0, st)) return e;
    assemble_cl_kernel<<<sgrid(rows2 * seg * 4), 256, 0, st>>>(s->c1b, t1, s->left2b, is_first ? 0 : 10, is_first ? 5 : 0,
                                                              s->in2b, tin2, S, 81);
    DSB_CHECK_LAUNCH();
    if (!is_last) {
      save_tail_cl_kernel<<<sgrid(rows2 * 40), 256, 0, st>>>(s->in2b, tin2, s->left2b, rows2);
      DSB_CHECK_LAUNCH();
    }
    const int segp[4] = {seg, 5, tin2, S};
    if (int e = conv_block_tc(s->in2b, m->convs[1], false, nullptr, 1, S * seg, s->xb, true, m->rnns[0].in_ld, st, segp))
      return e;
    prof_end(ST_CONV, st);

    // ---- BatchRNNStream stack (model.py:219-237) on the tensor cores ----
    fill_i32_kernel<<<cdiv(S, 256), 256, 0, st>>>(s->lens, S, T2);
    DSB_CHECK_LAUNCH();
    DSB_CUDA(cudaMemsetAsync(s->sync_words, 0, sizeof(unsigned int) * (kRnnSyncCounters + 1), st));
    const int next_ld = (H + 7) / 8 * 8;
    bool used_tc_rnn = false;
    int dev_id = 0, dev_sms = 148;
    cudaGetDevice(&dev_id);
    cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, dev_id);
    for (size_t l = 0; l < m->rnns.size(); ++l) {
      const RnnLayer& R = m->rnns[l];
      const bool last = l + 1 == m->rnns.size();
      const int N = R.gates * R.H;
      const bool tc_rnn = R.tc_recurrence && rnn_tc_supported(R, S, dev_sms, nullptr, nullptr);   // plan depends on the group size
      float* h_io = s->h + (int64_t)l * S * H;
      float* c_io = R.gates == 4 ? s->c + (int64_t)l * S * H : nullptr;
      prof_begin(ST_PROJ, st);
      if (int e = gemm_bias_tc(s->xb, R.in_ld, R.w_ih_tc, R.in_ld, tc_rnn ? R.b_ih_tc : R.b_ih, s->gates, N,
                               T2 * S, N, R.in_size, st))
        return e;
      prof_end(ST_PROJ, st);
      prof_begin(ST_RNN, st);
      if (tc_rnn) {
        used_tc_rnn = true;
        if (int e = rnn_layer_tc(R, s->gates, nullptr, S, T2, T2, s->ya, s->hbuf, s->sync_words,
                                 reinterpret_cast<int*>(s->sync_words + kRnnSyncCounters), st,
                                 s->h_init ? h_io : nullptr, (s->h_init && c_io) ? c_io : nullptr, h_io, c_io))
          return e;
        if (int e = combine_dirs_tc(s->ya, 1, T2, S, H, s->lens, last ? nullptr : s->xb, next_ld, last ? s->yb : nullptr, st))
          return e;
      } else {
        if (int e = rnn_layer_f32_state(R, s->gates, s->lens, S, T2, s->yb, s->hs, s->cs, h_io, s->c + (int64_t)l * S * H,
                                        s->h_init, st))
          return e;
        if (!last)
          if (int e = f32_to_bf16_ld(s->yb, s->xb, (int64_t)T2 * S, H, next_ld, st)) return e;
      }
      prof_end(ST_RNN, st);
    }
    x = s->yb;
    if (used_tc_rnn) {
      int abort_flag = 0;
      DSB_CUDA(cudaMemcpyAsync(&abort_flag, s->sync_words + kRnnSyncCounters, sizeof(int), cudaMemcpyDeviceToHost, st));
      DSB_CUDA(cudaStreamSynchronize(st));
      if (abort_flag) return set_error(DSB_ERR_CUDA, "dsb_streaming_forward: persistent recurrence step barrier timed out");
    }
  } else {
    const int64_t rows2 = (int64_t)S * 32 * 81;
    fill_i32_kernel<<<cdiv(S, 256), 256, 0, st>>>(s->lens, S, 1 << 30);   // no masking in the streaming convs
    DSB_CHECK_LAUNCH();
    if (int e = conv2d_bn_htanh_f32(s->in1, S, 1, kFreqBins, tin1, m->convs[0], s->lens, s->c1, t1, false, st)) return e;
    assemble_kernel<<<sgrid(rows2 * tin2), 256, 0, st>>>(s->c1, t1, s->left2, is_first ? 0 : 10, is_first ? 5 : 0, s->in2,
                                                       tin2, rows2);
    DSB_CHECK_LAUNCH();
    if (!is_last) {
      save_tail_kernel<<<sgrid(rows2 * 10), 256, 0, st>>>(s->in2, tin2, s->left2, 10, rows2);
      DSB_CHECK_LAUNCH();
    }
    if (int e = conv2d_bn_htanh_f32(s->in2, S, 32, 81, tin2, m->convs[1], s->lens, s->x0, T2, true, st)) return e;
    prof_end(ST_CONV, st);

    // ---- BatchRNNStream stack (model.py:219-237): hidden state carried across chunks ----
    fill_i32_kernel<<<cdiv(S, 256), 256, 0, st>>>(s->lens, S, T2);
    DSB_CHECK_LAUNCH();
    x = s->x0;
    float* ybuf[2] = {s->ya, s->yb};
    int cur = 0;
    for (size_t l = 0; l < m->rnns.size(); ++l) {
      const RnnLayer& R = m->rnns[l];
      prof_begin(ST_PROJ, st);
      if (int e = gemm_bias_f32(x, R.w_ih, R.b_ih, s->gates, (int64_t)T2 * S, R.gates * R.H, R.in_size, st)) return e;
      prof_end(ST_PROJ, st);
      prof_begin(ST_RNN, st);
      if (int e = rnn_layer_f32_state(R, s->gates, s->lens, S, T2, ybuf[cur], s->hs, s->cs, s->h + (int64_t)l * S * H,
                                      s->c + (int64_t)l * S * H, s->h_init, st))
        return e;
      prof_end(ST_RNN, st);
      x = ybuf[cur];
      cur ^= 1;
    }
  }
  s->h_init = !is_last;   // previous_hidden is dropped on is_last only

  // ---- LookaheadStream (model.py:255-279) ----
  ProfScope tail(ST_TAIL, st);
  const int64_t frame = (int64_t)S * H;
  if (!s->look_init || is_first) {
    DSB_CUDA(cudaMemcpyAsync(s->look, x, sizeof(float) * frame * T2, cudaMemcpyDeviceToDevice, st));
    s->look_len = T2;
    s->look_init = true;
    return 0;   // still buffering: the reference returns None
  }
  const int L = s->look_len + T2;
  DSB_CUDA(cudaMemcpyAsync(s->cat, s->look, sizeof(float) * frame * s->look_len, cudaMemcpyDeviceToDevice, st));
  DSB_CUDA(cudaMemcpyAsync(s->cat + frame * s->look_len, x, sizeof(float) * frame * T2, cudaMemcpyDeviceToDevice, st));
  const int keep = ctx - 1 < T2 ? ctx - 1 : T2;   // x[-(context-1):]
  DSB_CUDA(cudaMemcpyAsync(s->look, x + frame * (T2 - keep), sizeof(float) * frame * keep, cudaMemcpyDeviceToDevice, st));
  s->look_len = keep;
  // depthwise conv over the concatenation; the kernel zero-pads on the right, which is exactly the
  // is_last padding; otherwise only the first L-(context-1) outputs have their full right context
  const int n_out = is_last ? L : L - (ctx - 1);
  if (is_last) s->look_init = false;
  if (n_out <= 0) return 0;
  if (int e = lookahead_htanh_f32(s->cat, m->lookahead_w, s->lo, L, S, H, ctx, st)) return e;
  if (int e = fc_softmax_argmax_f32(s->lo, m->fc_w, m->fc_b, probs, nullptr, s->logits, n_out, S, C, H, st)) return e;
  *k_out = n_out;
  return 0;
}
