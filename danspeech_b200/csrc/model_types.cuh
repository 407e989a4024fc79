// Internal model representation shared by the model-path translation units.
#pragma once
#include "common.cuh"
#include <map>
#include <string>
#include <vector>

namespace dsb {

constexpr int kFreqBins = 161;
constexpr int kConvKW = 11;   // every DanSpeech conv is kH x 11 with time padding 5
constexpr int kConvPT = 5;

struct ConvLayer {
  int cin = 0, cout = 0, kh = 0, sd = 2, st = 1, pd = 0;
  int din = 0, dout = 0;
  float* w = nullptr;      // BN-folded weights [cout][cin][kh][11] fp32
  float* bias = nullptr;   // BN-folded bias [cout]
  // bf16 tensor-core path
  __nv_bfloat16* w_tc = nullptr;   // [r][kw][pack*cout][cin_pad] (conv2/3) or [r][pack*cout][16] (conv1, kw padded to 16), conv_tc.cu
};

struct RnnLayer {
  int in_size = 0, H = 0, gates = 3, dirs = 1;
  float* w_ih = nullptr;   // [dirs*gates*H, in_size], BatchNorm1d folded in (layers >= 1)
  float* b_ih = nullptr;   // [dirs*gates*H]
  float* w_hh = nullptr;   // [dirs][gates*H][H]
  float* b_hh = nullptr;   // [dirs][gates*H]
  // bf16 tensor-core path
  __nv_bfloat16* w_ih_tc = nullptr;   // [dirs*gates*H, in_ld] bf16 copy of w_ih
  int in_ld = 0;                      // row stride of w_ih_tc / of the activation operand (multiple of 8)
  float* b_ih_tc = nullptr;           // b_ih + b_hh for every gate except the GRU n-gate
  float* b_hn = nullptr;              // [dirs][H] GRU n-gate hidden bias
  __nv_bfloat16* w_hh_pack = nullptr; // per-CTA W_hh slices for the persistent recurrence (rnn_tc.cu)
  bool tc_recurrence = false;
  __nv_bfloat16* w_hh_pack_ks = nullptr;   // per-CTA-pair W_hh slices for the K-split recurrence (rnn_ks.cu)
  bool ks_recurrence = false;
};

struct HostTensor {
  const float* data;
  int64_t numel;
};

}  // namespace dsb

struct dsb_model {
  dsb_model_desc desc{};
  int precision = -1;
  bool finalized = false;
  std::map<std::string, dsb::HostTensor> tensors;
  std::vector<dsb::ConvLayer> convs;
  std::vector<dsb::RnnLayer> rnns;
  int rnn_input = 0;                 // C*D fed to the first recurrent layer
  float* lookahead_w = nullptr;      // [H][context]
  float* fc_w = nullptr;             // BN-folded [C][H]
  float* fc_b = nullptr;             // BN-folded [C]
  std::vector<void*> owned;          // device allocations to free
  // Abort flag of the persistent recurrence (a stuck step barrier raises it instead of hanging the GPU): a device
  // word that is sticky for the life of the model, mirrored into pinned host memory by an async copy at the end of
  // every forward -- no host/device synchronisation per batch (dsb_forward_status reads the mirror).
  int* d_abort = nullptr;
  int* h_abort = nullptr;
};

namespace dsb {

// ---- kernels implemented across the translation units (fp32 CUDA-core path) ----
int conv2d_bn_htanh_f32(const float* x, int B, int cin, int din, int tin, const ConvLayer& L, const int32_t* d_len,
                        float* y, int tout, bool rnn_layout, cudaStream_t st);
int gemm_bias_f32(const float* A, const float* W, const float* bias, float* C, int64_t M, int N, int K,
                  cudaStream_t st);
int rnn_layer_f32(const dsb_model* m, const RnnLayer& L, const float* gates_x, const int32_t* d_len, int B, int Tmax,
                  int Trows, float* y, float* h_state, float* c_state, cudaStream_t st);
int lookahead_htanh_f32(const float* x, const float* w, float* y, int T, int B, int H, int context, cudaStream_t st);
int softmax_argmax_f32(const float* logits, float* probs, int32_t* argmax, int T, int B, int C, cudaStream_t st);
int fc_softmax_argmax_f32(const float* x, const float* W, const float* bias, float* probs, int32_t* argmax,
                          float* logits_scratch, int T, int B, int C, int H, cudaStream_t st);

// ---- bf16 tensor-core path (tcgen05 / TMEM / TMA) ----
int gemm_bias_tc(const __nv_bfloat16* A, int64_t lda, const __nv_bfloat16* W, int64_t ldw, const float* bias,
                 float* C, int64_t ldc, int M, int N, int K, cudaStream_t st);
size_t conv1_tiles_elems(int B, int Tp);   // bf16 elements of the block-1 operand tiles
int im2col_time_tc(const float* spect, __nv_bfloat16* x1, int B, int T, int Tp, cudaStream_t st);
size_t conv_w_tc_elems(const ConvLayer& L, bool first);   // bf16 elements of the packed weights
int pack_conv_w_tc(const ConvLayer& L, bool first, __nv_bfloat16* out, cudaStream_t st);
int conv_block_tc(const __nv_bfloat16* x, const ConvLayer& L, bool first, const int32_t* d_len, int B, int Tp,
                  __nv_bfloat16* out, bool rnn_layout, int64_t out_ld, cudaStream_t st, const int* seg = nullptr);
bool rnn_tc_supported(const RnnLayer& L, int B, int sms, int* cpd_out, int* launches_out);
size_t rnn_tc_pack_elems(const RnnLayer& L);
int pack_whh_tc(const RnnLayer& L, __nv_bfloat16* out, cudaStream_t st);
int combine_dirs_tc(const float* y, int dirs, int T, int B, int H, const int32_t* d_len, __nv_bfloat16* xb, int ldx,
                    float* xf, cudaStream_t st);
// sync words of the persistent recurrence: kRnnMaxCounters step counters (one per direction x CTA set x group in
// flight), each on its own 128-byte line (the L2 atomic unit serialises per address), then the abort flag
constexpr int kRnnCounterStride = 32;
constexpr int kRnnMaxCounters = 48;
constexpr int kRnnSyncCounters = kRnnMaxCounters * kRnnCounterStride;
int rnn_tc_max_in_flight();
size_t rnn_tc_hbuf_elems(const RnnLayer& L, int B);
int rnn_layer_tc(const RnnLayer& L, const float* gx, const int32_t* d_len, int B, int T, int Tmax, float* y,
                 __nv_bfloat16* hbuf, unsigned int* sync_words, int* abort_flag, cudaStream_t st,
                 const float* h0 = nullptr, const float* c0 = nullptr, float* hT = nullptr, float* cT = nullptr,
                 bool batch_minor = false);
bool rnn_batch_minor(const RnnLayer& L, int B);
int rnn_tc_init_hbuf(const float* h0, __nv_bfloat16* hbuf, int dirs, int B, int H, int HP, int BP, int n_bgroups,
                     cudaStream_t st);
// CTA-pair recurrence on tcgen05.mma.cta_group::2 (rnn_pair.cu): same contract and W_hh slices as rnn_layer_tc
bool rnn_tc_narrow(const RnnLayer& L, int B);
bool rnn_pair_supported(const RnnLayer& L, int B, int sms);
// batch_minor: gx is [dirs*G*H][T*B] (gemm_bias_rows_tc) and y is written as [dirs][H][T*B] (combine_dirs_t_tc)
int rnn_layer_pair(const RnnLayer& L, const float* gx, const int32_t* d_len, int B, int T, int Tmax, float* y,
                   __nv_bfloat16* hbuf, unsigned int* sync_words, int* abort_flag, cudaStream_t st,
                   const float* h0 = nullptr, const float* c0 = nullptr, float* hT = nullptr, float* cT = nullptr,
                   bool batch_minor = false);
int gemm_bias_rows_tc(const __nv_bfloat16* A, int64_t lda, const __nv_bfloat16* W, int64_t ldw, const float* bias,
                      float* Ct, int64_t ldc, int M, int N, int K, cudaStream_t st);
int combine_dirs_t_tc(const float* yt, int dirs, int T, int B, int H, const int32_t* d_len, __nv_bfloat16* xb, int ldx,
                      float* xf, cudaStream_t st);
// K-split CTA-pair recurrence (rnn_ks.cu): same contract as rnn_layer_tc, batch groups of 64 rows
bool rnn_ks_supported(const RnnLayer& L, int sms, int* pairs_out, int* launches_out);
size_t rnn_ks_pack_elems(const RnnLayer& L);
int pack_whh_ks(const RnnLayer& L, __nv_bfloat16* out, cudaStream_t st);
int rnn_layer_ks(const RnnLayer& L, const float* gx, const int32_t* d_len, int B, int T, int Tmax, float* y,
                 __nv_bfloat16* hbuf, unsigned int* sync_words, int* abort_flag, cudaStream_t st,
                 const float* h0 = nullptr, const float* c0 = nullptr, float* hT = nullptr, float* cT = nullptr);
int f32_to_bf16_ld(const float* x, __nv_bfloat16* y, int64_t rows, int cols, int ld, cudaStream_t st);
int finalize_tc(dsb_model* m, cudaStream_t st);
size_t forward_tc_workspace_bytes(const dsb_model* m, int B, int T);
int forward_tc(dsb_model* m, const float* spect, const int32_t* h_out_len, int B, int T, float* probs,
               int32_t* argmax, void* workspace, cudaStream_t st);

}  // namespace dsb
