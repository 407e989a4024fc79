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
// W_hh[:, k-chunk] and h_{t-1}[:, k-chunk] through shared memory.
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

int rnn_layer_f32(const dsb_model* m, const RnnLayer& L, const float* gates_x, const int32_t* d_len, int B, int Tmax,
                  int Trows, float* y, float* h_state, float* c_state, cudaStream_t st) {
  const int H = L.H, dirs = L.dirs;
  const size_t hbytes = sizeof(float) * (size_t)dirs * B * H;
  DSB_CUDA(cudaMemsetAsync(h_state, 0, 2 * hbytes, st));
  if (L.gates == 4) DSB_CUDA(cudaMemsetAsync(c_state, 0, hbytes, st));
  DSB_CUDA(cudaMemsetAsync(y, 0, sizeof(float) * (size_t)Trows * B * H, st));
  dim3 grid(cdiv(H, RJ), dirs, cdiv(B, RB));
  float* hbuf[2] = {h_state, h_state + (size_t)dirs * B * H};
  for (int s = 0; s < Tmax; ++s) {
    const float* hp = hbuf[s & 1];
    float* hn = hbuf[(s + 1) & 1];
    if (L.gates == 3)
      rnn_step_f32_kernel<3><<<grid, 256, 0, st>>>(gates_x, L.w_hh, L.b_hh, hp, hn, c_state, y, d_len, s, B, H, dirs);
    else if (L.gates == 4)
      rnn_step_f32_kernel<4><<<grid, 256, 0, st>>>(gates_x, L.w_hh, L.b_hh, hp, hn, c_state, y, d_len, s, B, H, dirs);
    else
      rnn_step_f32_kernel<1><<<grid, 256, 0, st>>>(gates_x, L.w_hh, L.b_hh, hp, hn, c_state, y, d_len, s, B, H, dirs);
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
  dim3 grid(cdiv(H, RJ), 1, cdiv(B, RB));
  float* hbuf[2] = {h_scratch, h_scratch + (size_t)B * H};
  for (int s = 0; s < T; ++s) {
    const float* hp = hbuf[s & 1];
    float* hn = hbuf[(s + 1) & 1];
    if (L.gates == 3)
      rnn_step_f32_kernel<3><<<grid, 256, 0, st>>>(gates_x, L.w_hh, L.b_hh, hp, hn, c_io, y, d_len, s, B, H, 1);
    else if (L.gates == 4)
      rnn_step_f32_kernel<4><<<grid, 256, 0, st>>>(gates_x, L.w_hh, L.b_hh, hp, hn, c_io, y, d_len, s, B, H, 1);
    else
      rnn_step_f32_kernel<1><<<grid, 256, 0, st>>>(gates_x, L.w_hh, L.b_hh, hp, hn, c_io, y, d_len, s, B, H, 1);
    count_launch();
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(DSB_ERR_CUDA, "rnn step launch failed: %s", cudaGetErrorString(e));
  DSB_CUDA(cudaMemcpyAsync(h_io, hbuf[T & 1], hbytes, cudaMemcpyDeviceToDevice, st));
  return 0;
}

}  // namespace dsb
