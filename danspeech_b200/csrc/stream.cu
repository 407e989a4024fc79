// Chunked streaming forward for S lock-step streams.
//
// Replaces DeepSpeech.streaming_forward (danspeech/deepspeech/model.py:517-537) and the stateful
// modules MaskConvStream (:169-201), BatchRNNStream (:219-237) and LookaheadStream (:255-279).  The
// per-module-instance state of the reference (left conv contexts, GRU hidden state, lookahead tail)
// becomes an explicit device-resident dsb_stream_state for S streams, so one model serves many
// concurrent streams.  Quirks kept on purpose: the convs still apply their own symmetric
// time padding on top of the carried context (overlapping frames are re-emitted), the first call only
// fills the lookahead buffer and emits nothing, the hidden state is reset on is_last only.
// This path runs on the exact-fp32 kernels in every precision mode.
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
  SA(left1, (int64_t)S * kFreqBins * 10);
  SA(left2, (int64_t)S * 32 * 81 * 10);
  SA(h, (int64_t)layers * S * H);
  SA(c, (int64_t)layers * S * H);
  SA(look, (int64_t)s->look_cap * S * H);
  SA(in1, (int64_t)S * kFreqBins * tin1);
  SA(c1, (int64_t)S * 32 * 81 * t1);
  SA(in2, (int64_t)S * 32 * 81 * tin2);
  SA(x0, (int64_t)tin2 * S * m->rnn_input);
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

  // ---- MaskConvStream (model.py:169-201) ----
  prof_begin(ST_CONV, st);
  const int64_t rows1 = (int64_t)S * kFreqBins, rows2 = (int64_t)S * 32 * 81;
  assemble_kernel<<<sgrid(rows1 * tin1), 256, 0, st>>>(chunk, k, s->left1, is_first ? 0 : 10, is_first ? 5 : 0, s->in1,
                                                     tin1, rows1);
  DSB_CHECK_LAUNCH();
  if (!is_last) {
    save_tail_kernel<<<sgrid(rows1 * 10), 256, 0, st>>>(s->in1, tin1, s->left1, 10, rows1);
    DSB_CHECK_LAUNCH();
  }
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
  const float* x = s->x0;
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
  if (int e = gemm_bias_f32(s->lo, m->fc_w, m->fc_b, s->logits, (int64_t)n_out * S, C, H, st)) return e;
  if (int e = softmax_argmax_f32(s->logits, probs, nullptr, n_out, S, C, st)) return e;
  *k_out = n_out;
  return 0;
}
