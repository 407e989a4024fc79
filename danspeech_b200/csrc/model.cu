// Acoustic-model handle: tensor intake, BatchNorm folding / layout packing, forward orchestration.
//
// Replaces danspeech/deepspeech/model.py: DeepSpeech.__init__ (:293-425, shape bookkeeping),
// DeepSpeech.forward (:496-515) and get_seq_lens (:540-551).  State-dict names follow SURVEY A.6.
#include "model_types.cuh"
#include <math.h>
#include <string.h>

namespace dsb {

// ------------------------------------------------------------------------------------------
// weight preparation kernels (run once in dsb_model_finalize)
// ------------------------------------------------------------------------------------------

// y = clamp(conv(x, W*s, (b-mu)*s+beta)): s = gamma/sqrt(var+eps)        
__global__ void fold_conv_kernel(const float* __restrict__ w, const float* __restrict__ b,
                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                 const float* __restrict__ mean, const float* __restrict__ var, int cout, int cin,
                                 int kh, float* __restrict__ w_out /*[kh][cin][11][cout]*/,
                                 float* __restrict__ b_out) {
  const int64_t total = (int64_t)cout * cin * kh * kConvKW;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int kw = (int)(i % kConvKW);
    int64_t r = i / kConvKW;
    int k = (int)(r % kh);
    r /= kh;
    int ci = (int)(r % cin);
    int co = (int)(r / cin);
    float s = gamma[co] / sqrtf(var[co] + 1e-5f);
    w_out[(((int64_t)k * cin + ci) * kConvKW + kw) * cout + co] = w[i] * s;
  }
  for (int co = blockIdx.x * blockDim.x + threadIdx.x; co < cout; co += gridDim.x * blockDim.x) {
    float s = gamma[co] / sqrtf(var[co] + 1e-5f);
    b_out[co] = (b[co] - mean[co]) * s + beta[co];
  }
}

// rows of W (one warp per row): W'[r][k] = W[r][k]*a[k];  b'[r] = b[r] + sum_k W[r][k]*shift[k]
// a = gamma/sqrt(var+eps), shift = beta - mu*a   (SURVEY A.4).  bn pointers may be NULL (no BN).
__global__ void fold_linear_kernel(const float* __restrict__ w, const float* __restrict__ b,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   const float* __restrict__ mean, const float* __restrict__ var, int rows, int K,
                                   float* __restrict__ w_out, float* __restrict__ b_out) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  double acc = 0.0;
  for (int k = lane; k < K; k += 32) {
    float wv = w[(int64_t)row * K + k];
    float a = 1.0f, sh = 0.0f;
    if (gamma) {
      a = gamma[k] / sqrtf(var[k] + 1e-5f);
      sh = beta[k] - mean[k] * a;
    }
    w_out[(int64_t)row * K + k] = wv * a;
    acc += (double)wv * (double)sh;
  }
  acc = warp_sum(acc);
  if (lane == 0 && b_out) b_out[row] = (float)((b ? (double)b[row] : 0.0) + acc);
}

static int grid_for(int64_t n) { return (int)(cdiv64(n, 256) < 4096 ? cdiv64(n, 256) : 4096); }

struct Lookup {
  dsb_model* m;
  int err = 0;
  const float* get(const std::string& name, int64_t numel) {
    auto it = m->tensors.find(name);
    if (it == m->tensors.end()) {
      err = set_error(DSB_ERR_STATE, "dsb_model_finalize: missing tensor '%s'", name.c_str());
      return nullptr;
    }
    if (it->second.numel != numel) {
      err = set_error(DSB_ERR_INVALID, "dsb_model_finalize: tensor '%s' has %lld elements, expected %lld",
                      name.c_str(), (long long)it->second.numel, (long long)numel);
      return nullptr;
    }
    return it->second.data;
  }
};

template <typename T>
static int dev_alloc(dsb_model* m, T** p, int64_t n) {
  void* q = nullptr;
  DSB_CUDA(cudaMalloc(&q, sizeof(T) * (size_t)(n > 0 ? n : 1)));
  m->owned.push_back(q);
  *p = reinterpret_cast<T*>(q);
  return 0;
}

static const char* kDirSuffix[2] = {"", "_reverse"};

}  // namespace dsb

using namespace dsb;

extern "C" int dsb_model_create(const dsb_model_desc* d, dsb_model** out) {
  DSB_REQUIRE(d && out, "dsb_model_create: null argument");
  DSB_REQUIRE(d->conv_layers >= 1 && d->conv_layers <= 3, "dsb_model_create: conv_layers %d not in 1..3",
              d->conv_layers);
  DSB_REQUIRE(d->rnn_layers >= 1 && d->rnn_hidden_size >= 1, "dsb_model_create: bad rnn shape");
  DSB_REQUIRE(d->rnn_type >= 0 && d->rnn_type <= 2, "dsb_model_create: bad rnn_type %d", d->rnn_type);
  DSB_REQUIRE(d->num_classes >= 2, "dsb_model_create: num_classes %d", d->num_classes);
  DSB_REQUIRE(d->bidirectional || d->context >= 1, "dsb_model_create: context must be >= 1 for uni-directional");
  dsb_model* m = new dsb_model();
  m->desc = *d;
  // conv geometry (model.py:357-396)
  const int chans[4] = {1, 32, 32, 96};
  const int khs[3] = {41, 21, 21};
  const int pds[3] = {20, 10, 10};
  int din = kFreqBins;
  for (int i = 0; i < d->conv_layers; ++i) {
    ConvLayer L;
    L.cin = chans[i];
    L.cout = chans[i + 1];
    L.kh = khs[i];
    L.pd = pds[i];
    L.sd = 2;
    L.st = i == 0 ? 2 : 1;
    L.din = din;
    L.dout = (din + 2 * L.pd - L.kh) / 2 + 1;
    din = L.dout;
    m->convs.push_back(L);
  }
  // quirk kept from the reference: streaming models always size the first RNN for the 2-conv stack
  // (model.py:477-484); only 2-conv streaming models are self-consistent.
  m->rnn_input = m->convs.back().cout * m->convs.back().dout;
  const int gates = d->rnn_type == DSB_RNN_GRU ? 3 : d->rnn_type == DSB_RNN_LSTM ? 4 : 1;
  for (int l = 0; l < d->rnn_layers; ++l) {
    RnnLayer R;
    R.in_size = l == 0 ? m->rnn_input : d->rnn_hidden_size;
    R.H = d->rnn_hidden_size;
    R.gates = gates;
    R.dirs = (d->bidirectional && !d->streaming) ? 2 : 1;
    m->rnns.push_back(R);
  }
  *out = m;
  return 0;
}

extern "C" void dsb_model_destroy(dsb_model* m) {
  if (!m) return;
  for (void* p : m->owned) cudaFree(p);
  if (m->d_abort) cudaFree(m->d_abort);
  if (m->h_abort) cudaFreeHost(m->h_abort);
  delete m;
}

extern "C" int dsb_forward_status(const dsb_model* m) {
  DSB_REQUIRE(m, "dsb_forward_status: null model");
  if (m->h_abort && *(volatile int*)m->h_abort)
    return set_error(DSB_ERR_CUDA, "dsb_forward: the persistent recurrence's step barrier timed out (results of that "
                                   "call are invalid; re-create the model)");
  return 0;
}

extern "C" int dsb_model_set_tensor(dsb_model* m, const char* name, const float* data, int64_t numel) {
  DSB_REQUIRE(m && name && data && numel > 0, "dsb_model_set_tensor: null argument");
  m->tensors[name] = HostTensor{data, numel};
  return 0;
}

extern "C" int dsb_model_precision(const dsb_model* m) { return m ? m->precision : -1; }

extern "C" int dsb_model_out_frames(const dsb_model* m, int T) {
  (void)m;
  return T <= 0 ? 0 : (T - 1) / 2 + 1;   // only the first conv strides time (stride 2, k 11, pad 5)
}

extern "C" int dsb_model_finalize(dsb_model* m, int precision, void* stream) {
  DSB_REQUIRE(m, "dsb_model_finalize: null model");
  DSB_REQUIRE(precision == DSB_PREC_FP32 || precision == DSB_PREC_BF16, "dsb_model_finalize: bad precision %d",
              precision);
  cudaStream_t st = (cudaStream_t)stream;
  for (void* p : m->owned) cudaFree(p);
  m->owned.clear();
  if (!m->d_abort) {
    DSB_CUDA(cudaMalloc(reinterpret_cast<void**>(&m->d_abort), sizeof(int)));
    DSB_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&m->h_abort), sizeof(int), cudaHostAllocDefault));
  }
  DSB_CUDA(cudaMemsetAsync(m->d_abort, 0, sizeof(int), st));
  *m->h_abort = 0;
  Lookup lk{m};
  const dsb_model_desc& d = m->desc;
  char nm[128];

  for (int i = 0; i < (int)m->convs.size(); ++i) {
    ConvLayer& L = m->convs[i];
    const int64_t wn = (int64_t)L.cout * L.cin * L.kh * kConvKW;
    snprintf(nm, sizeof nm, "conv.seq_module.%d.weight", 3 * i);
    const float* w = lk.get(nm, wn);
    snprintf(nm, sizeof nm, "conv.seq_module.%d.bias", 3 * i);
    const float* b = lk.get(nm, L.cout);
    snprintf(nm, sizeof nm, "conv.seq_module.%d.weight", 3 * i + 1);
    const float* g = lk.get(nm, L.cout);
    snprintf(nm, sizeof nm, "conv.seq_module.%d.bias", 3 * i + 1);
    const float* be = lk.get(nm, L.cout);
    snprintf(nm, sizeof nm, "conv.seq_module.%d.running_mean", 3 * i + 1);
    const float* mu = lk.get(nm, L.cout);
    snprintf(nm, sizeof nm, "conv.seq_module.%d.running_var", 3 * i + 1);
    const float* var = lk.get(nm, L.cout);
    if (lk.err) return lk.err;
    if (int e = dev_alloc(m, &L.w, wn)) return e;
    if (int e = dev_alloc(m, &L.bias, L.cout)) return e;
    fold_conv_kernel<<<grid_for(wn), 256, 0, st>>>(w, b, g, be, mu, var, L.cout, L.cin, L.kh, L.w, L.bias);
    DSB_CHECK_LAUNCH();
  }

  for (int l = 0; l < (int)m->rnns.size(); ++l) {
    RnnLayer& R = m->rnns[l];
    const int GH = R.gates * R.H;
    if (int e = dev_alloc(m, &R.w_ih, (int64_t)R.dirs * GH * R.in_size)) return e;
    if (int e = dev_alloc(m, &R.b_ih, (int64_t)R.dirs * GH)) return e;
    if (int e = dev_alloc(m, &R.w_hh, (int64_t)R.dirs * GH * R.H)) return e;
    if (int e = dev_alloc(m, &R.b_hh, (int64_t)R.dirs * GH)) return e;
    const float *g = nullptr, *be = nullptr, *mu = nullptr, *var = nullptr;
    if (l > 0) {   // layer 0 has no BatchNorm (model.py:399-400)
      snprintf(nm, sizeof nm, "rnns.%d.batch_norm.module.weight", l);
      g = lk.get(nm, R.in_size);
      snprintf(nm, sizeof nm, "rnns.%d.batch_norm.module.bias", l);
      be = lk.get(nm, R.in_size);
      snprintf(nm, sizeof nm, "rnns.%d.batch_norm.module.running_mean", l);
      mu = lk.get(nm, R.in_size);
      snprintf(nm, sizeof nm, "rnns.%d.batch_norm.module.running_var", l);
      var = lk.get(nm, R.in_size);
    }
    for (int dir = 0; dir < R.dirs; ++dir) {
      snprintf(nm, sizeof nm, "rnns.%d.rnn.weight_ih_l0%s", l, kDirSuffix[dir]);
      const float* wih = lk.get(nm, (int64_t)GH * R.in_size);
      snprintf(nm, sizeof nm, "rnns.%d.rnn.bias_ih_l0%s", l, kDirSuffix[dir]);
      const float* bih = lk.get(nm, GH);
      snprintf(nm, sizeof nm, "rnns.%d.rnn.weight_hh_l0%s", l, kDirSuffix[dir]);
      const float* whh = lk.get(nm, (int64_t)GH * R.H);
      snprintf(nm, sizeof nm, "rnns.%d.rnn.bias_hh_l0%s", l, kDirSuffix[dir]);
      const float* bhh = lk.get(nm, GH);
      if (lk.err) return lk.err;
      fold_linear_kernel<<<cdiv(GH, 8), 256, 0, st>>>(wih, bih, g, be, mu, var, GH, R.in_size,
                                                     R.w_ih + (int64_t)dir * GH * R.in_size,
                                                     R.b_ih + (int64_t)dir * GH);
      DSB_CHECK_LAUNCH();
      DSB_CUDA(cudaMemcpyAsync(R.w_hh + (int64_t)dir * GH * R.H, whh, sizeof(float) * (size_t)GH * R.H,
                               cudaMemcpyDeviceToDevice, st));
      DSB_CUDA(cudaMemcpyAsync(R.b_hh + (int64_t)dir * GH, bhh, sizeof(float) * GH, cudaMemcpyDeviceToDevice, st));
    }
  }

  const int H = d.rnn_hidden_size, C = d.num_classes;
  if (!d.bidirectional || d.streaming) {
    // offline uni-directional: "lookahead.0.conv.weight"; streaming: "lookahead.conv.weight" 
    const char* key = d.streaming ? "lookahead.conv.weight" : "lookahead.0.conv.weight";
    const float* lw = lk.get(key, (int64_t)H * d.context);
    if (lk.err) return lk.err;
    if (int e = dev_alloc(m, &m->lookahead_w, (int64_t)H * d.context)) return e;
    DSB_CUDA(cudaMemcpyAsync(m->lookahead_w, lw, sizeof(float) * (size_t)H * d.context, cudaMemcpyDeviceToDevice, st));
  }
  {
    const float* g = lk.get("fc.0.module.0.weight", H);
    const float* be = lk.get("fc.0.module.0.bias", H);
    const float* mu = lk.get("fc.0.module.0.running_mean", H);
    const float* var = lk.get("fc.0.module.0.running_var", H);
    const float* w = lk.get("fc.0.module.1.weight", (int64_t)C * H);
    if (lk.err) return lk.err;
    if (int e = dev_alloc(m, &m->fc_w, (int64_t)C * H)) return e;
    if (int e = dev_alloc(m, &m->fc_b, C)) return e;
    fold_linear_kernel<<<cdiv(C, 8), 256, 0, st>>>(w, nullptr, g, be, mu, var, C, H, m->fc_w, m->fc_b);
    DSB_CHECK_LAUNCH();
  }
  if (precision == DSB_PREC_BF16) {
    if (int e = finalize_tc(m, st)) return e;
  }
  DSB_CUDA(cudaStreamSynchronize(st));
  m->tensors.clear();
  m->precision = precision;
  m->finalized = true;
  return 0;
}

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
namespace dsb {

struct Workspace {
  int32_t* d_len;
  float* act[2];     // conv ping-pong / RNN layer in-out
  float* gates;      // [T'*B][dirs*G*H]
  float* hstate;     // 2 x [dirs][B][H]
  float* cstate;     // [dirs][B][H]
  float* logits;     // [T'*B][C]
  size_t total;
};

static Workspace carve(const dsb_model* m, int B, int T, void* base) {
  const dsb_model_desc& d = m->desc;
  const int Tp = dsb_model_out_frames(m, T);
  const int dirs = m->rnns[0].dirs, G = m->rnns[0].gates, H = d.rnn_hidden_size;
  size_t act_elems = 0;
  for (const ConvLayer& L : m->convs) act_elems = max(act_elems, (size_t)B * L.cout * L.dout * Tp);
  act_elems = max(act_elems, (size_t)Tp * B * H);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += align_up(bytes, 256);
    return o;
  };
  Workspace w{};
  char* p = reinterpret_cast<char*>(base);
  size_t o_len = take(sizeof(int32_t) * B);
  size_t o_a0 = take(sizeof(float) * act_elems);
  size_t o_a1 = take(sizeof(float) * act_elems);
  size_t o_g = take(sizeof(float) * (size_t)Tp * B * dirs * G * H);
  size_t o_h = take(sizeof(float) * 2 * (size_t)dirs * B * H);
  size_t o_c = take(sizeof(float) * (size_t)dirs * B * H);
  size_t o_l = take(sizeof(float) * (size_t)Tp * B * d.num_classes);
  w.total = off;
  if (p) {
    w.d_len = reinterpret_cast<int32_t*>(p + o_len);
    w.act[0] = reinterpret_cast<float*>(p + o_a0);
    w.act[1] = reinterpret_cast<float*>(p + o_a1);
    w.gates = reinterpret_cast<float*>(p + o_g);
    w.hstate = reinterpret_cast<float*>(p + o_h);
    w.cstate = reinterpret_cast<float*>(p + o_c);
    w.logits = reinterpret_cast<float*>(p + o_l);
  }
  return w;
}

}  // namespace dsb

extern "C" size_t dsb_forward_workspace_bytes(const dsb_model* m, int B, int T) {
  if (!m || B <= 0 || T <= 0) return 0;
  size_t fp32 = carve(m, B, T, nullptr).total;
  size_t tc = m->precision == DSB_PREC_BF16 ? forward_tc_workspace_bytes(m, B, T) : 0;
  return fp32 > tc ? fp32 : tc;
}

extern "C" int dsb_forward(dsb_model* m, const float* spect, const int32_t* lengths, int B, int T, float* probs,
                           int32_t* out_lengths, int32_t* argmax, void* workspace, size_t workspace_bytes,
                           void* stream) {
  DSB_REQUIRE(m && spect && lengths && probs && out_lengths && workspace, "dsb_forward: null argument");
  if (!m->finalized) return set_error(DSB_ERR_STATE, "dsb_forward: model not finalized");
  if (m->desc.streaming) return set_error(DSB_ERR_STATE, "dsb_forward: streaming model, use dsb_streaming_forward");
  DSB_REQUIRE(B > 0 && T > 0, "dsb_forward: bad shape B=%d T=%d", B, T);
  for (int b = 0; b < B; ++b) {
    DSB_REQUIRE(lengths[b] >= 1 && lengths[b] <= T, "dsb_forward: lengths[%d]=%d outside [1,%d]", b, lengths[b], T);
    // pack_padded_sequence(enforce_sorted=True) contract (model.py:117)
    DSB_REQUIRE(b == 0 || lengths[b] <= lengths[b - 1],
                "dsb_forward: lengths must be sorted in decreasing order (lengths[%d]=%d > lengths[%d]=%d)", b,
                lengths[b], b - 1, lengths[b - 1]);
    out_lengths[b] = dsb_model_out_frames(m, lengths[b]);
  }
  if (workspace_bytes < dsb_forward_workspace_bytes(m, B, T))
    return set_error(DSB_ERR_WORKSPACE, "dsb_forward: workspace %zu < required %zu", workspace_bytes,
                     dsb_forward_workspace_bytes(m, B, T));
  cudaStream_t st = (cudaStream_t)stream;
  if (int e = dsb_forward_status(m)) return e;   // an earlier call's recurrence aborted: the flag is read lazily
  if (m->precision == DSB_PREC_BF16)
    return forward_tc(m, spect, out_lengths, B, T, probs, argmax, workspace, st);

  Workspace ws = carve(m, B, T, workspace);
  const dsb_model_desc& d = m->desc;
  const int Tp = dsb_model_out_frames(m, T);
  const int H = d.rnn_hidden_size, C = d.num_classes;
  DSB_CUDA(cudaMemcpyAsync(ws.d_len, out_lengths, sizeof(int32_t) * B, cudaMemcpyHostToDevice, st));

  // MaskConv stack; the last block writes the [T', B, C*D] layout the RNN consumes (model.py:501-503)
  const float* x = spect;
  int cur = 0, cin = 1, din = kFreqBins, tin = T;
  prof_begin(ST_CONV, st);
  for (size_t i = 0; i < m->convs.size(); ++i) {
    const bool last = i + 1 == m->convs.size();
    if (int e = conv2d_bn_htanh_f32(x, B, cin, din, tin, m->convs[i], ws.d_len, ws.act[cur], Tp, last, st)) return e;
    x = ws.act[cur];
    cur ^= 1;
    cin = m->convs[i].cout;
    din = m->convs[i].dout;
    tin = Tp;
  }
  prof_end(ST_CONV, st);
  const int Tmax = out_lengths[0];
  for (size_t l = 0; l < m->rnns.size(); ++l) {
    const RnnLayer& R = m->rnns[l];
    const int N = R.dirs * R.gates * R.H;
    prof_begin(ST_PROJ, st);
    if (int e = gemm_bias_f32(x, R.w_ih, R.b_ih, ws.gates, (int64_t)Tp * B, N, R.in_size, st)) return e;
    prof_end(ST_PROJ, st);
    prof_begin(ST_RNN, st);
    // y rows t in [Tmax, Tp) must be zero as well: rnn_layer_f32 zeroes Tp rows
    if (int e = rnn_layer_f32(m, R, ws.gates, ws.d_len, B, Tmax, Tp, ws.act[cur], ws.hstate, ws.cstate, st))
      return e;
    prof_end(ST_RNN, st);
    x = ws.act[cur];
    cur ^= 1;
  }
  ProfScope tail_scope(ST_TAIL, st);
  if (!d.bidirectional) {
    if (int e = lookahead_htanh_f32(x, m->lookahead_w, ws.act[cur], Tp, B, H, d.context, st)) return e;
    x = ws.act[cur];
    cur ^= 1;
  }
  return fc_softmax_argmax_f32(x, m->fc_w, m->fc_b, probs, argmax, ws.logits, Tp, B, C, H, st);
}
