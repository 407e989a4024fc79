// bf16 tensor-core forward path (DSB_PREC_BF16): DeepSpeech.forward (model.py:496-515) with
//   * MaskConv blocks             -> conv_tc.cu (implicit GEMM, tcgen05) or the fp32 direct kernels
//   * BatchRNN input projections  -> gemm_tc.cu (tcgen05 GEMM, BatchNorm1d folded into W_ih/b_ih)
//   * BatchRNN recurrences        -> rnn_tc.cu (persistent, W_hh resident in shared memory)
//   * lookahead / fc / softmax    -> fp32 tail kernels
// Activations between layers are bf16 [T'*B, H]; gate pre-activations, per-direction outputs, the
// recurrent state (registers) and everything after the last RNN layer are fp32.
#include "model_types.cuh"

namespace dsb {

__global__ void f32_to_bf16_ld_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int64_t rows,
                                      int cols, int ld) {
  const int64_t total = rows * cols;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols;
    const int c = (int)(i - r * cols);
    y[r * ld + c] = __float2bfloat16_rn(x[i]);
  }
}

// b_ih_tc = b_ih + b_hh for all gates except the GRU n-gate; b_hn = b_hh of the GRU n-gate
__global__ void fold_rnn_bias_kernel(const float* __restrict__ b_ih, const float* __restrict__ b_hh, int dirs, int G,
                                     int H, float* __restrict__ b_out, float* __restrict__ b_hn) {
  const int total = dirs * G * H;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int j = i % H, g = (i / H) % G, d = i / (G * H);
    const bool n_gate = (G == 3 && g == 2);
    b_out[i] = b_ih[i] + (n_gate ? 0.f : b_hh[i]);
    if (n_gate) b_hn[d * H + j] = b_hh[i];
  }
}

int f32_to_bf16_ld(const float* x, __nv_bfloat16* y, int64_t rows, int cols, int ld, cudaStream_t st) {
  const int64_t n = rows * cols;
  f32_to_bf16_ld_kernel<<<(int)(cdiv64(n, 256) < 4096 ? cdiv64(n, 256) : 4096), 256, 0, st>>>(x, y, rows, cols, ld);
  DSB_CHECK_LAUNCH();
  return 0;
}

template <typename T>
static int dev_alloc_tc(dsb_model* m, T** p, int64_t n) {
  void* q = nullptr;
  DSB_CUDA(cudaMalloc(&q, sizeof(T) * (size_t)(n > 0 ? n : 1)));
  m->owned.push_back(q);
  *p = reinterpret_cast<T*>(q);
  return 0;
}

static int conv_grid(int64_t n) { return (int)(cdiv64(n, 256) < 4096 ? cdiv64(n, 256) : 4096); }

// The tensor-core conv stack hands the RNN channels-last features (index d*C + c) instead of the
// reference's channel-major order (c*D + d, model.py:502); permuting the columns of the first layer's
// W_ih once makes the projection identical.
__global__ void permute_w_ih0_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int rows, int C,
                                     int D, int ld) {
  const int64_t total = (int64_t)rows * C * D;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int col = (int)(i % (C * D));
    const int64_t r = i / (C * D);
    const int d = col / C, c = col - d * C;
    out[r * ld + col] = __float2bfloat16_rn(w[r * (int64_t)(C * D) + (int64_t)c * D + d]);
  }
}

int finalize_tc(dsb_model* m, cudaStream_t st) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  for (size_t i = 0; i < m->convs.size(); ++i) {
    ConvLayer& L = m->convs[i];
    const bool first = i == 0;
    if (int e = dev_alloc_tc(m, &L.w_tc, (int64_t)conv_w_tc_elems(L, first))) return e;
    if (int e = pack_conv_w_tc(L, first, L.w_tc, st)) return e;
  }
  bool first_rnn = true;
  for (RnnLayer& R : m->rnns) {
    const int GH = R.gates * R.H, rows = R.dirs * GH;
    R.in_ld = (R.in_size + 7) / 8 * 8;
    if (int e = dev_alloc_tc(m, &R.w_ih_tc, (int64_t)rows * R.in_ld)) return e;
    DSB_CUDA(cudaMemsetAsync(R.w_ih_tc, 0, sizeof(__nv_bfloat16) * (size_t)rows * R.in_ld, st));
    if (first_rnn) {
      const ConvLayer& LC = m->convs.back();
      permute_w_ih0_kernel<<<conv_grid((int64_t)rows * R.in_size), 256, 0, st>>>(R.w_ih, R.w_ih_tc, rows, LC.cout,
                                                                                 LC.dout, R.in_ld);
    } else {
      f32_to_bf16_ld_kernel<<<conv_grid((int64_t)rows * R.in_size), 256, 0, st>>>(R.w_ih, R.w_ih_tc, rows, R.in_size,
                                                                                 R.in_ld);
    }
    DSB_CHECK_LAUNCH();
    first_rnn = false;
    if (int e = dev_alloc_tc(m, &R.b_ih_tc, rows)) return e;
    if (int e = dev_alloc_tc(m, &R.b_hn, (int64_t)R.dirs * R.H)) return e;
    fold_rnn_bias_kernel<<<cdiv(rows, 256), 256, 0, st>>>(R.b_ih, R.b_hh, R.dirs, R.gates, R.H, R.b_ih_tc, R.b_hn);
    DSB_CHECK_LAUNCH();
    int cpd = 0, launches = 0;
    R.tc_recurrence = rnn_tc_supported(R, 64, sms, &cpd, &launches);
    if (R.tc_recurrence) {
      if (int e = dev_alloc_tc(m, &R.w_hh_pack, (int64_t)rnn_tc_pack_elems(R))) return e;
      if (int e = pack_whh_tc(R, R.w_hh_pack, st)) return e;
    }
    // the K-split CTA-pair kernel shares the exchange buffers and counters of the one-CTA kernel (groups of 64 rows)
    R.ks_recurrence = R.tc_recurrence && rnn_ks_supported(R, sms, nullptr, nullptr);
    if (R.ks_recurrence) {
      if (int e = dev_alloc_tc(m, &R.w_hh_pack_ks, (int64_t)rnn_ks_pack_elems(R))) return e;
      if (int e = pack_whh_ks(R, R.w_hh_pack_ks, st)) return e;
    }
  }
  return 0;
}

namespace {
struct TcWorkspace {
  int32_t* d_len;
  float* act[2];
  __nv_bfloat16* cb[2];   // channels-last bf16 conv activations (ping-pong)
  __nv_bfloat16* xb;
  float* gates;
  float* ydir;
  float* xf;
  __nv_bfloat16* hbuf;
  unsigned int* sync_words;
  float* hstate;
  float* cstate;
  float* logits;
  int ld;
  size_t total;
};

TcWorkspace carve_tc(const dsb_model* m, int B, int T, void* base) {
  const dsb_model_desc& d = m->desc;
  const int Tp = dsb_model_out_frames(m, T);
  const int dirs = m->rnns[0].dirs, G = m->rnns[0].gates, H = d.rnn_hidden_size;
  const size_t act_elems = (size_t)Tp * B * H;   // fp32 scratch for the lookahead output
  size_t cb0 = conv1_tiles_elems(B, Tp), cb1 = 0;
  for (size_t i = 0; i < m->convs.size(); ++i) {
    const ConvLayer& L = m->convs[i];
    const size_t e = (size_t)B * L.dout * Tp * L.cout;
    if (i % 2 == 0) cb1 = max(cb1, e); else cb0 = max(cb0, e);
  }
  const int ld = max((m->rnn_input + 7) / 8 * 8, (H + 7) / 8 * 8);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += align_up(bytes, 1024);
    return o;
  };
  TcWorkspace w{};
  char* p = reinterpret_cast<char*>(base);
  const size_t o_len = take(sizeof(int32_t) * B);
  const size_t o_a0 = take(sizeof(float) * act_elems);
  const size_t o_a1 = take(sizeof(float) * act_elems);
  const size_t o_cb0 = take(sizeof(__nv_bfloat16) * cb0);
  const size_t o_cb1 = take(sizeof(__nv_bfloat16) * cb1);
  const size_t o_xb = take(sizeof(__nv_bfloat16) * (size_t)Tp * B * ld);
  const size_t o_g = take(sizeof(float) * (size_t)Tp * B * dirs * G * H);
  const size_t o_y = take(sizeof(float) * (size_t)dirs * Tp * B * H);
  const size_t o_xf = take(sizeof(float) * (size_t)Tp * B * H);
  const size_t o_hb = take(sizeof(__nv_bfloat16) * rnn_tc_hbuf_elems(m->rnns[0], B));
  const size_t o_sw = take(sizeof(unsigned int) * (kRnnSyncCounters + 1));
  const size_t o_h = take(sizeof(float) * 2 * (size_t)dirs * B * H);
  const size_t o_c = take(sizeof(float) * (size_t)dirs * B * H);
  const size_t o_l = take(sizeof(float) * (size_t)Tp * B * d.num_classes);
  w.total = off;
  w.ld = ld;
  if (p) {
    w.d_len = reinterpret_cast<int32_t*>(p + o_len);
    w.act[0] = reinterpret_cast<float*>(p + o_a0);
    w.act[1] = reinterpret_cast<float*>(p + o_a1);
    w.cb[0] = reinterpret_cast<__nv_bfloat16*>(p + o_cb0);
    w.cb[1] = reinterpret_cast<__nv_bfloat16*>(p + o_cb1);
    w.xb = reinterpret_cast<__nv_bfloat16*>(p + o_xb);
    w.gates = reinterpret_cast<float*>(p + o_g);
    w.ydir = reinterpret_cast<float*>(p + o_y);
    w.xf = reinterpret_cast<float*>(p + o_xf);
    w.hbuf = reinterpret_cast<__nv_bfloat16*>(p + o_hb);
    w.sync_words = reinterpret_cast<unsigned int*>(p + o_sw);
    w.hstate = reinterpret_cast<float*>(p + o_h);
    w.cstate = reinterpret_cast<float*>(p + o_c);
    w.logits = reinterpret_cast<float*>(p + o_l);
  }
  return w;
}
}  // namespace

size_t forward_tc_workspace_bytes(const dsb_model* m, int B, int T) { return carve_tc(m, B, T, nullptr).total; }

int forward_tc(dsb_model* m, const float* spect, const int32_t* h_out_len, int B, int T, float* probs,
               int32_t* argmax, void* workspace, cudaStream_t st) {
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0)
    return set_error(DSB_ERR_INVALID, "dsb_forward: workspace must be 256-byte aligned");
  TcWorkspace ws = carve_tc(m, B, T, workspace);
  const dsb_model_desc& d = m->desc;
  const int Tp = dsb_model_out_frames(m, T);
  const int H = d.rnn_hidden_size, C = d.num_classes;
  const int64_t M = (int64_t)Tp * B;
  DSB_CUDA(cudaMemcpyAsync(ws.d_len, h_out_len, sizeof(int32_t) * B, cudaMemcpyHostToDevice, st));
  DSB_CUDA(cudaMemsetAsync(ws.sync_words, 0, sizeof(unsigned int) * kRnnSyncCounters, st));   // step counters

  prof_begin(ST_CONV, st);
  if (int e = im2col_time_tc(spect, ws.cb[0], B, T, Tp, st)) return e;
  {
    const __nv_bfloat16* x = ws.cb[0];
    for (size_t i = 0; i < m->convs.size(); ++i) {
      const ConvLayer& L = m->convs[i];
      const bool last = i + 1 == m->convs.size();
      __nv_bfloat16* y = last ? ws.xb : ws.cb[(i + 1) & 1];
      if (int e = conv_block_tc(x, L, i == 0, ws.d_len, B, Tp, y, last, m->rnns[0].in_ld, st)) return e;
      x = y;
    }
  }
  prof_end(ST_CONV, st);

  const int Tmax = h_out_len[0];
  bool used_tc_rnn = false;
  int dev_id = 0, dev_sms = 148;
  cudaGetDevice(&dev_id);
  cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, dev_id);
  for (size_t l = 0; l < m->rnns.size(); ++l) {
    const RnnLayer& R = m->rnns[l];
    const bool last = l + 1 == m->rnns.size();
    const int N = R.dirs * R.gates * R.H;
    // the persistent kernel's shared-memory plan depends on the batch-group size (64 or 128 rows): wide layers that
    // fit with 64 rows may not fit with 128, in which case this batch takes the per-step fp32 recurrence
    const bool tc_rnn = R.tc_recurrence && rnn_tc_supported(R, B, dev_sms, nullptr, nullptr);
    // the CTA-pair recurrence reads batch-minor pre-activations and writes batch-minor outputs (coalesced per warp)
    const bool bminor = tc_rnn && rnn_batch_minor(R, B);
    prof_begin(ST_PROJ, st);
    if (bminor) {
      if (int e = gemm_bias_rows_tc(ws.xb, R.in_ld, R.w_ih_tc, R.in_ld, R.b_ih_tc, ws.gates, M, (int)M, N, R.in_size, st))
        return e;
    } else if (int e = gemm_bias_tc(ws.xb, R.in_ld, R.w_ih_tc, R.in_ld, tc_rnn ? R.b_ih_tc : R.b_ih, ws.gates, N, (int)M,
                                    N, R.in_size, st)) {
      return e;
    }
    prof_end(ST_PROJ, st);
    prof_begin(ST_RNN, st);
    const int next_ld = (H + 7) / 8 * 8;
    if (tc_rnn) {
      used_tc_rnn = true;
      if (int e = rnn_layer_tc(R, ws.gates, ws.d_len, B, Tp, Tmax, ws.ydir, ws.hbuf, ws.sync_words, m->d_abort, st,
                               nullptr, nullptr, nullptr, nullptr, bminor))
        return e;
      prof_end(ST_RNN, st);
      prof_begin(ST_COMBINE, st);
      if (bminor) {
        if (int e = combine_dirs_t_tc(ws.ydir, R.dirs, Tp, B, H, ws.d_len, last ? nullptr : ws.xb, next_ld,
                                      last ? ws.xf : nullptr, st))
          return e;
      } else if (int e = combine_dirs_tc(ws.ydir, R.dirs, Tp, B, H, ws.d_len, last ? nullptr : ws.xb, next_ld,
                                         last ? ws.xf : nullptr, st)) {
        return e;
      }
      prof_end(ST_COMBINE, st);   // (the prof_end(ST_RNN) below is then a no-op)
    } else {
      if (int e = rnn_layer_f32(m, R, ws.gates, ws.d_len, B, Tmax, Tp, ws.xf, ws.hstate, ws.cstate, st)) return e;
      if (!last) {
        f32_to_bf16_ld_kernel<<<conv_grid(M * H), 256, 0, st>>>(ws.xf, ws.xb, M, H, next_ld);
        DSB_CHECK_LAUNCH();
      }
    }
    prof_end(ST_RNN, st);
  }

  prof_begin(ST_TAIL, st);
  const float* xt = ws.xf;
  if (!d.bidirectional) {
    if (int e = lookahead_htanh_f32(ws.xf, m->lookahead_w, ws.act[0], Tp, B, H, d.context, st)) return e;
    xt = ws.act[0];
  }
  if (int e = fc_softmax_argmax_f32(xt, m->fc_w, m->fc_b, probs, argmax, ws.logits, Tp, B, C, H, st)) return e;
  prof_end(ST_TAIL, st);

  if (used_tc_rnn) {
    // The persistent recurrence never spins forever: a stuck step barrier raises the (sticky) abort flag.  Its value
    // travels to pinned host memory behind the kernels of this call; nobody waits for it here -- dsb_forward_status()
    // and the next dsb_forward() read the mirror (the caller synchronises anyway before it touches the results).
    DSB_CUDA(cudaMemcpyAsync(m->h_abort, m->d_abort, sizeof(int), cudaMemcpyDeviceToHost, st));
  }
  return 0;
}

}  // namespace dsb
