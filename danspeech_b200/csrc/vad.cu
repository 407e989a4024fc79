// Energy voice-activity detection for S concurrent 16-bit PCM streams ("next" row SURVEY 8f-4).
//
// Replaces, per stream, the phrase state machine of Recognizer.listen_stream
// (danspeech/Recognizer.py:218-324): energy = audioop.rms(buffer, 2) compared with energy_threshold; a phrase
// starts with the first loud buffer, ends after more than pause_buffer_count quiet buffers in a row and is
// kept only if it holds at least phrase_buffer_count buffers before that pause.  The reference runs this in a
// Python generator on one microphone; here one warp per stream handles a buffer of every stream per call and
// the host only sees an event code per stream (it owns the audio buffers, e.g. the non-speaking pre-roll).
// HBM-bound: 2 * chunk_samples bytes read per stream per call.
#include "common.cuh"

struct dsb_vad_state {
  int S = 0, pause_buffers = 0, phrase_buffers = 0;
  int32_t* mode = nullptr;     // [S] 0 = waiting for speech, 1 = inside a phrase
  int32_t* pause = nullptr;    // [S] consecutive quiet buffers inside the phrase
  int32_t* phrase = nullptr;   // [S] buffers read since the phrase started
};

namespace dsb {

__global__ void vad_push_kernel(const int16_t* __restrict__ chunks, int64_t stride, int n, const int32_t* __restrict__ thr,
                                int32_t* __restrict__ mode, int32_t* __restrict__ pause, int32_t* __restrict__ phrase,
                                int pause_buffers, int phrase_buffers, int32_t* __restrict__ energy_out,
                                int32_t* __restrict__ event_out, int S) {
  const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (s >= S) return;
  const int16_t* x = chunks + (int64_t)s * stride;
  unsigned long long acc = 0;   // exact: n * 2^30 fits easily
  if ((stride & 7) == 0 && (reinterpret_cast<uintptr_t>(chunks) & 15) == 0) {
    const int n8 = n >> 3;
    const uint4* x8 = reinterpret_cast<const uint4*>(x);
    for (int i = lane; i < n8; i += 32) {
      const uint4 v = __ldg(x8 + i);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int a = (int)(short)(w[j] & 0xFFFFu), b = (int)(short)(w[j] >> 16);
        acc += (unsigned long long)(a * a) + (unsigned long long)(b * b);
      }
    }
    for (int i = (n8 << 3) + lane; i < n; i += 32) {
      const int a = x[i];
      acc += (unsigned long long)(a * a);
    }
  } else {
    for (int i = lane; i < n; i += 32) {
      const int a = x[i];
      acc += (unsigned long long)(a * a);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane != 0) return;
  // audioop.rms: (unsigned int) sqrt(sum_squares / n) in double precision
  const int energy = n > 0 ? (int)(unsigned int)sqrt((double)acc / (double)n) : 0;
  const bool loud = energy > thr[s];
  int ev;
  if (mode[s] == 0) {
    if (loud) {
      mode[s] = 1;
      pause[s] = 0;
      phrase[s] = 0;
      ev = DSB_VAD_PHRASE_START;
    } else {
      ev = DSB_VAD_SILENCE;
    }
  } else {
    const int ph = phrase[s] + 1;
    const int pa = loud ? 0 : pause[s] + 1;
    if (pa > pause_buffers) {
      mode[s] = 0;
      ev = (ph - pa >= phrase_buffers) ? DSB_VAD_PHRASE_END : DSB_VAD_PHRASE_DROPPED;
    } else {
      ev = DSB_VAD_SPEECH;
    }
    phrase[s] = ph;
    pause[s] = pa;
  }
  energy_out[s] = energy;
  event_out[s] = ev;
}

}  // namespace dsb

using namespace dsb;

extern "C" int dsb_vad_state_create(int n_streams, int pause_buffers, int phrase_buffers, dsb_vad_state** out) {
  DSB_REQUIRE(out && n_streams > 0 && pause_buffers >= 0 && phrase_buffers >= 0, "dsb_vad_state_create: bad argument");
  dsb_vad_state* v = new dsb_vad_state();
  v->S = n_streams;
  v->pause_buffers = pause_buffers;
  v->phrase_buffers = phrase_buffers;
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, sizeof(int32_t) * 3 * (size_t)n_streams);
  if (e != cudaSuccess) {
    delete v;
    return set_error(DSB_ERR_CUDA, "dsb_vad_state_create: %s", cudaGetErrorString(e));
  }
  cudaMemset(p, 0, sizeof(int32_t) * 3 * (size_t)n_streams);
  v->mode = reinterpret_cast<int32_t*>(p);
  v->pause = v->mode + n_streams;
  v->phrase = v->pause + n_streams;
  *out = v;
  return 0;
}

extern "C" void dsb_vad_state_destroy(dsb_vad_state* v) {
  if (!v) return;
  cudaFree(v->mode);
  delete v;
}

extern "C" int dsb_vad_reset(dsb_vad_state* v, void* stream) {
  DSB_REQUIRE(v, "dsb_vad_reset: null state");
  DSB_CUDA(cudaMemsetAsync(v->mode, 0, sizeof(int32_t) * 3 * (size_t)v->S, (cudaStream_t)stream));
  return 0;
}

extern "C" int dsb_vad_push_s16(dsb_vad_state* v, const int16_t* chunks, int64_t chunk_stride, int chunk_samples,
                                const int32_t* energy_threshold, int32_t* energy_out, int32_t* event_out, void* stream) {
  DSB_REQUIRE(v && chunks && energy_threshold && energy_out && event_out, "dsb_vad_push_s16: null argument");
  DSB_REQUIRE(chunk_samples > 0 && chunk_stride >= chunk_samples, "dsb_vad_push_s16: bad chunk size %d (stride %lld)",
              chunk_samples, (long long)chunk_stride);
  const int warps = 8;
  vad_push_kernel<<<cdiv(v->S, warps), warps * 32, 0, (cudaStream_t)stream>>>(
      chunks, chunk_stride, chunk_samples, energy_threshold, v->mode, v->pause, v->phrase, v->pause_buffers,
      v->phrase_buffers, energy_out, event_out, v->S);
  DSB_CHECK_LAUNCH();
  return 0;
}
