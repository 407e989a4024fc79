// Fused spectrogram front-end for sm_100a.
//
// Replaces danspeech/audio/parsers.py:50-72 (SpectrogramAudioParser.parse_audio) and the
// STFT/log1p half of parsers.py:101-163 (InferenceSpectrogramAudioParser.parse_audio):
//   reflect-pad(160) -> frames of 320 every 160 -> symmetric Hamming -> rfft-320 -> |.| -> log1p
//   -> (S - mean) / unbiased-std over the whole utterance.
//
// One CTA = 32 consecutive frames of one utterance.  The 5 280 samples the tile touches are
// staged once in shared memory (frames overlap by 50 %), each warp then transforms 4 frames:
// the real 320-point FFT is a 160-point complex FFT (radix 5 in registers x radix-2^5 across
// the lanes with warp shuffles) plus the usual even/odd untangling.  log1p|X| is transposed
// through shared memory so that global stores are 128-byte rows of the [161, T] output.
//
// Algorithmic bytes = 4*n (samples) + 4*161*T (output) per utterance (the HBM roofline the kernel is
// reported against).  The transform runs in fp64 (see SpecTables) and is instruction-bound, so the FFT
// is computed exactly once (MODE_RAW writes log1p|X| plus per-tile sum / sum of squares in fp64) and the
// per-utterance normalisation is a separate elementwise pass over data that is still L2-resident.
#include "common.cuh"
#include <math.h>
#include <mutex>

namespace dsb {

constexpr int kNfft = 320;
constexpr int kHop = 160;
constexpr int kBins = 161;
constexpr int kFT = 32;                             // frames per CTA
constexpr int kTileSamples = kHop * (kFT - 1) + kNfft;  // 5280
constexpr int kWarps = 8;

enum { MODE_STATS = 0, MODE_NORM = 1, MODE_RAW = 2 };

// The reference STFT runs in float64 (librosa on a float64 signal) and int16-scale audio has ~1e6:1
// dynamic range between strong and weak bins, so an fp32 FFT leaves ~1e-3 errors in the weak bins
// after log1p.  The transform therefore runs in fp64 (B200 FP64 is half the FP32 rate; the kernel
// stays far from the FP64 roof); only log1p and the normalisation are fp32, as upstream.
// A second instantiation runs the transform in fp32 ("fast" FFT): 2-4x fewer issue slots, errors up to ~2e-3
// in weak bins, i.e. inside the 2e-2 bar of the bf16 mode whose convolutions round the input to bf16 anyway.
template <typename T> struct V2;
template <> struct V2<double> { typedef double2 type; };
template <> struct V2<float> { typedef float2 type; };
template <typename T>
struct SpecTablesT {
  T window[4][kNfft];                    // symmetric hamming, hann, blackman, bartlett (scipy.signal, parsers.py:9-10)
  typename V2<T>::type tw160[5][32];     // W160^(lane*k1)
  typename V2<T>::type tw32[4][32];      // radix-2 DIF stage twiddles, halves 16, 8, 4, 2
  typename V2<T>::type tw320[kBins + 3]; // W320^k, k = 0..160
};
typedef SpecTablesT<double> SpecTables;
__device__ SpecTablesT<double> g_tab;
__device__ SpecTablesT<float> g_tabf;
template <typename T> __device__ __forceinline__ const SpecTablesT<T>& tables();
template <> __device__ __forceinline__ const SpecTablesT<double>& tables<double>() { return g_tab; }
template <> __device__ __forceinline__ const SpecTablesT<float>& tables<float>() { return g_tabf; }

// The tables live in __device__ symbols, i.e. once per device: uploaded on first use of every device.
static std::mutex g_tab_mu;
static bool g_tab_done[64] = {false};

static cudaError_t init_tables() {
  static SpecTables h;
  cudaError_t g_tab_err = cudaSuccess;
  const double PI = 3.14159265358979323846;
  for (int k = 0; k < kNfft; ++k) {
    // scipy.signal.{hamming,hann,blackman,bartlett}(320), sym=True (librosa calls the window function with win_length)
    const double x = 2.0 * PI * k / (kNfft - 1);
    h.window[0][k] = 0.54 - 0.46 * cos(x);
    h.window[1][k] = 0.5 - 0.5 * cos(x);
    h.window[2][k] = 0.42 - 0.5 * cos(x) + 0.08 * cos(2.0 * x);
    h.window[3][k] = 1.0 - fabs(2.0 * k / (kNfft - 1) - 1.0);
  }
  for (int k1 = 0; k1 < 5; ++k1)
    for (int l = 0; l < 32; ++l) {
      double a = -2.0 * PI * (double)(l * k1) / 160.0;
      h.tw160[k1][l] = make_double2(cos(a), sin(a));
    }
  const int halves[4] = {16, 8, 4, 2};
  for (int s = 0; s < 4; ++s)
    for (int l = 0; l < 32; ++l) {
      int hh = halves[s];
      double a = -2.0 * PI * (double)(l & (hh - 1)) / (double)(2 * hh);
      h.tw32[s][l] = make_double2(cos(a), sin(a));
    }
  for (int k = 0; k < kBins + 3; ++k) {
    double a = -2.0 * PI * (double)k / 320.0;
    h.tw320[k] = make_double2(cos(a), sin(a));
  }
  g_tab_err = cudaMemcpyToSymbol(g_tab, &h, sizeof(h));
  static SpecTablesT<float> hf;
  for (int w = 0; w < 4; ++w)
    for (int k = 0; k < kNfft; ++k) hf.window[w][k] = (float)h.window[w][k];
  for (int a = 0; a < 5; ++a)
    for (int l = 0; l < 32; ++l) hf.tw160[a][l] = make_float2((float)h.tw160[a][l].x, (float)h.tw160[a][l].y);
  for (int a = 0; a < 4; ++a)
    for (int l = 0; l < 32; ++l) hf.tw32[a][l] = make_float2((float)h.tw32[a][l].x, (float)h.tw32[a][l].y);
  for (int k = 0; k < kBins + 3; ++k) hf.tw320[k] = make_float2((float)h.tw320[k].x, (float)h.tw320[k].y);
  if (g_tab_err == cudaSuccess) g_tab_err = cudaMemcpyToSymbol(g_tabf, &hf, sizeof(hf));
  return g_tab_err;
}

template <typename C, typename T>
__device__ __forceinline__ C mk2(T x, T y) {
  C r;
  r.x = x;
  r.y = y;
  return r;
}
template <typename C>
__device__ __forceinline__ C cmul(C a, C b) {
  C r;
  r.x = a.x * b.x - a.y * b.y;
  r.y = a.x * b.y + a.y * b.x;
  return r;
}
template <typename C>
__device__ __forceinline__ C cadd(C a, C b) {
  C r;
  r.x = a.x + b.x;
  r.y = a.y + b.y;
  return r;
}
template <typename C>
__device__ __forceinline__ C csub(C a, C b) {
  C r;
  r.x = a.x - b.x;
  r.y = a.y - b.y;
  return r;
}

// numpy 'reflect' padding index (periodic extension with period 2(n-1), no edge repeat).
__device__ __forceinline__ int reflect_index(int j, int n) {
  if (j >= 0 && j < n) return j;
  if (n == 1) return 0;
  int period = 2 * (n - 1);
  j %= period;
  if (j < 0) j += period;
  return j < n ? j : period - j;
}

template <typename T>
struct SpecSmemT {
  alignas(16) float samples[kTileSamples];   // read as float2 / written as float4
  alignas(16) T window[kNfft];               // read as pairs
  alignas(16) typename V2<T>::type z[kWarps][160];
  alignas(16) typename V2<T>::type tw320[kBins + 3];
  alignas(16) double red[kWarps][2];
  alignas(16) float tile[kBins][kFT + 1];
  float mean_std[2];
};

// sample j of an utterance: fp32 mono, or interleaved s16 PCM mixed down as clip(sum of channels) -- the
// audioop.tomono(buf, width, 1, 1) mix of the reference's load_audio (resources.py:302-303, quirk Q1)
template <bool S16>
__device__ __forceinline__ float load_sample(const void* __restrict__ base, int64_t j, int channels) {
  if (!S16) return __ldg(reinterpret_cast<const float*>(base) + j);
  const int16_t* p = reinterpret_cast<const int16_t*>(base) + j * channels;
  int acc = 0;
  for (int c = 0; c < channels; ++c) acc += (int)__ldg(p + c);
  return (float)max(-32768, min(32767, acc));
}

template <int MODE, typename T, bool S16>
__global__ void __launch_bounds__(kWarps * 32)
spectrogram_kernel(const void* __restrict__ audio_v, int channels, int64_t audio_stride, const int32_t* __restrict__ n_samples,
                   float* __restrict__ out, int64_t out_stride, float* __restrict__ mean_std_out,
                   double* __restrict__ partials, int n_partials, int center, int normalize, int window) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  typedef typename V2<T>::type T2;
  typedef SpecSmemT<T> SpecSmem;
  SpecSmem& sm = *reinterpret_cast<SpecSmem*>(smem_raw);
  const SpecTablesT<T>& g_tab = tables<T>();

  const int b = blockIdx.y;
  const int tile_idx = blockIdx.x;
  const int t0 = tile_idx * kFT;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = min(n_samples[b], (int)audio_stride);
  const int n_frames = center ? 1 + n / kHop : (n >= kNfft ? 1 + (n - kNfft) / kHop : 0);
  const float* audio = reinterpret_cast<const float*>(audio_v);
  const float* y = audio + (int64_t)b * audio_stride;                       // fp32 input
  const void* yb = S16 ? static_cast<const void*>(reinterpret_cast<const int16_t*>(audio_v) + (int64_t)b * audio_stride * channels)
                       : static_cast<const void*>(y);

  // mean / std from the MODE_STATS partials (fixed summation order -> deterministic).
  if (MODE == MODE_NORM) {
    if (warp == 0) {
      double s = 0.0, q = 0.0;
      const int used = normalize ? (n_frames + kFT - 1) / kFT : 0;
      for (int i = lane; i < used; i += 32) {
        s += partials[((int64_t)b * n_partials + i) * 2 + 0];
        q += partials[((int64_t)b * n_partials + i) * 2 + 1];
      }
      s = warp_sum(s);
      q = warp_sum(q);
      if (lane == 0) {
        double N = (double)n_frames * kBins;
        double mean = s / N;
        double var = (q - s * s / N) / (N - 1.0);   // torch.std: unbiased
        float m = normalize ? (float)mean : 0.0f;
        float sd = normalize ? (float)sqrt(var > 0.0 ? var : 0.0) : 1.0f;
        sm.mean_std[0] = m;
        sm.mean_std[1] = sd;
        if (tile_idx == 0 && mean_std_out) {
          mean_std_out[b * 2 + 0] = m;
          mean_std_out[b * 2 + 1] = sd;
        }
      }
    }
  }

  const int frames_here = min(kFT, n_frames - t0);   // may be <= 0 (pure zero fill)
  if (frames_here > 0) {
    // ---- stage samples (reflect padding at both utterance ends) ----
    const int g0 = t0 * kHop - (center ? kNfft / 2 : 0);
    const int need = kHop * (frames_here - 1) + kNfft;
    const bool interior = !S16 && (g0 >= 0) && (g0 + need <= n) && ((audio_stride & 3) == 0) &&
                          ((reinterpret_cast<uintptr_t>(audio) & 15) == 0);
    if (interior) {
      const float4* src = reinterpret_cast<const float4*>(y + g0);   // g0 % 4 == 0
      float4* dst = reinterpret_cast<float4*>(sm.samples);
      for (int i = tid; i < need / 4; i += kWarps * 32) dst[i] = __ldg(src + i);
    } else {
      for (int i = tid; i < need; i += kWarps * 32) {
        int j = g0 + i;
        float v = 0.0f;
        if (center) v = load_sample<S16>(yb, reflect_index(j, n), channels);
        else if (j < n) v = load_sample<S16>(yb, j, channels);
        sm.samples[i] = v;
      }
    }
    for (int i = tid; i < kNfft; i += kWarps * 32) sm.window[i] = g_tab.window[window][i];
    for (int i = tid; i < kBins + 3; i += kWarps * 32) sm.tw320[i] = g_tab.tw320[i];
  }
  // per-lane twiddles (registers)
  T2 tw160[5], tw32[4];
#pragma unroll
  for (int k1 = 0; k1 < 5; ++k1) tw160[k1] = g_tab.tw160[k1][lane];
#pragma unroll
  for (int s = 0; s < 4; ++s) tw32[s] = g_tab.tw32[s][lane];
  __syncthreads();

  double dsum = 0.0, dsq = 0.0;
  const T C1 = (T)0.30901699437494742, C2 = (T)-0.80901699437494742;
  const T S1 = (T)0.95105651629515357, S2 = (T)0.58778525229247313;
  const int rlane = __brev((unsigned)lane) >> 27;

  for (int f = warp; f < frames_here; f += kWarps) {
    const float* s = sm.samples + f * kHop;
    T2 z[5];
#pragma unroll
    for (int n1 = 0; n1 < 5; ++n1) {
      int idx = 2 * (32 * n1 + lane);
      float2 v = *reinterpret_cast<const float2*>(s + idx);
      T2 w = *reinterpret_cast<const T2*>(sm.window + idx);
      z[n1] = mk2<T2, T>((T)v.x * w.x, (T)v.y * w.y);
    }
    // radix-5 over n1
    T2 a1 = cadd(z[1], z[4]), a2 = cadd(z[2], z[3]);
    T2 b1 = csub(z[1], z[4]), b2 = csub(z[2], z[3]);
    T2 Y[5];
    Y[0] = mk2<T2, T>(z[0].x + a1.x + a2.x, z[0].y + a1.y + a2.y);
    T2 p1 = mk2<T2, T>(z[0].x + C1 * a1.x + C2 * a2.x, z[0].y + C1 * a1.y + C2 * a2.y);
    T2 p2 = mk2<T2, T>(z[0].x + C2 * a1.x + C1 * a2.x, z[0].y + C2 * a1.y + C1 * a2.y);
    T2 q1 = mk2<T2, T>(S1 * b1.x + S2 * b2.x, S1 * b1.y + S2 * b2.y);
    T2 q2 = mk2<T2, T>(S2 * b1.x - S1 * b2.x, S2 * b1.y - S1 * b2.y);
    // -i*q = (q.y, -q.x)
    Y[1] = mk2<T2, T>(p1.x + q1.y, p1.y - q1.x);
    Y[4] = mk2<T2, T>(p1.x - q1.y, p1.y + q1.x);
    Y[2] = mk2<T2, T>(p2.x + q2.y, p2.y - q2.x);
    Y[3] = mk2<T2, T>(p2.x - q2.y, p2.y + q2.x);
#pragma unroll
    for (int k1 = 1; k1 < 5; ++k1) Y[k1] = cmul(Y[k1], tw160[k1]);
    // five interleaved 32-point DIF FFTs across the lanes
#pragma unroll
    for (int st = 0; st < 5; ++st) {
      const int half = 16 >> st;
      const bool upper = (lane & half) != 0;
#pragma unroll
      for (int k1 = 0; k1 < 5; ++k1) {
        T px = __shfl_xor_sync(0xffffffffu, Y[k1].x, half);
        T py = __shfl_xor_sync(0xffffffffu, Y[k1].y, half);
        if (!upper) {
          Y[k1] = mk2<T2, T>(Y[k1].x + px, Y[k1].y + py);
        } else {
          T2 d = mk2<T2, T>(px - Y[k1].x, py - Y[k1].y);
          Y[k1] = (st < 4) ? cmul(d, tw32[st < 4 ? st : 0]) : d;
        }
      }
    }
    // lane holds k2 = bitrev5(lane): Z[k1 + 5*k2]
    __syncwarp();
#pragma unroll
    for (int k1 = 0; k1 < 5; ++k1) sm.z[warp][k1 + 5 * rlane] = Y[k1];
    __syncwarp();
    // untangle: X[k] = (Zk + conj(ZN-k))/2 - i/2 * W320^k * (Zk - conj(ZN-k))
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      int k = lane + 32 * j;
      if (k <= 160) {
        T2 zk = sm.z[warp][k == 160 ? 0 : k];
        T2 zn = sm.z[warp][k == 0 || k == 160 ? 0 : 160 - k];
        zn.y = -zn.y;
        T2 e = cadd(zk, zn);
        T2 o = csub(zk, zn);
        T2 wo = cmul(sm.tw320[k], o);
        // -i*wo = (wo.y, -wo.x); complex64 storage of D upstream -> round the parts to fp32
        float re = (float)((T)0.5 * (e.x + wo.y));
        float im = (float)((T)0.5 * (e.y - wo.x));
        float mag = sizeof(T) == 8 ? (float)sqrt((double)re * (double)re + (double)im * (double)im)
                                   : sqrtf(re * re + im * im);
        float v = log1pf(mag);
        if (MODE != MODE_NORM) {
          dsum += (double)v;
          dsq += (double)v * (double)v;
        }
        if (MODE != MODE_STATS) sm.tile[k][f] = v;
      }
    }
    __syncwarp();
  }

  if (MODE != MODE_NORM) {
    dsum = warp_sum(dsum);
    dsq = warp_sum(dsq);
    if (lane == 0) {
      sm.red[warp][0] = dsum;
      sm.red[warp][1] = dsq;
    }
  }
  __syncthreads();
  if (MODE != MODE_NORM) {
    if (tid == 0 && tile_idx < n_partials) {
      double s = 0.0, q = 0.0;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) {
        s += sm.red[w][0];
        q += sm.red[w][1];
      }
      partials[((int64_t)b * n_partials + tile_idx) * 2 + 0] = s;
      partials[((int64_t)b * n_partials + tile_idx) * 2 + 1] = q;
    }
  }
  if (MODE != MODE_STATS) {
    float mean = 0.0f, sd = 1.0f;
    if (MODE == MODE_NORM) {
      mean = sm.mean_std[0];
      sd = sm.mean_std[1];
    }
    const int t = t0 + lane;
    if (t < out_stride) {
      float* o = out + (int64_t)b * kBins * out_stride + t;
      const bool valid = lane < frames_here;
      for (int k = warp; k < kBins; k += kWarps) {
        float v = 0.0f;
        if (valid) {
          v = sm.tile[k][lane];
          if (MODE == MODE_NORM) v = (v - mean) / sd;
        }
        o[(int64_t)k * out_stride] = v;
      }
    }
  }
}

// mean and biased std per stream from the MODE_RAW partials (numpy np.mean / np.std, parsers.py:148-149)
// center = 0: streaming chunk (np.mean / biased np.std, parsers.py:148-149) -> stats (f64)
// center = 1: offline utterance (torch mean / UNBIASED std, parsers.py:66-70) -> f32 (mean, std) stored in the
//             first 8 bytes of the utterance's partials row (and in mean_std_out when given)
__global__ void stream_stats_kernel(double* __restrict__ partials, int n_partials,
                                    const int32_t* __restrict__ n_samples, double* __restrict__ stats, int S,
                                    int center, int normalize, float* __restrict__ mean_std_out, int64_t audio_stride) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  int n = min(n_samples[s], (int)audio_stride);
  int n_frames = center ? 1 + n / kHop : (n >= kNfft ? 1 + (n - kNfft) / kHop : 0);
  int used = (n_frames + kFT - 1) / kFT;
  double a = 0.0, q = 0.0;
  for (int i = 0; i < used; ++i) {
    a += partials[((int64_t)s * n_partials + i) * 2 + 0];
    q += partials[((int64_t)s * n_partials + i) * 2 + 1];
  }
  double N = (double)n_frames * kBins;
  double mean = N > 0 ? a / N : 0.0;
  if (!center) {
    double var = N > 0 ? q / N - mean * mean : 0.0;
    stats[s * 2 + 0] = mean;
    stats[s * 2 + 1] = sqrt(var > 0.0 ? var : 0.0);
  } else {
    double var = N > 1 ? (q - a * a / N) / (N - 1.0) : 0.0;
    float m = normalize ? (float)mean : 0.0f;
    float sd = normalize ? (float)sqrt(var > 0.0 ? var : 0.0) : 1.0f;
    float* slot = reinterpret_cast<float*>(partials + (int64_t)s * n_partials * 2);
    slot[0] = m;
    slot[1] = sd;
    if (mean_std_out) {
      mean_std_out[s * 2 + 0] = m;
      mean_std_out[s * 2 + 1] = sd;
    }
  }
}

// ms_stride: floats between consecutive (mean, std) pairs; frames_from_samples: n_frames holds sample counts
__global__ void stream_normalize_kernel(float* __restrict__ spect, int64_t out_stride,
                                        const int32_t* __restrict__ n_frames, const float* __restrict__ mean_std,
                                        int64_t ms_stride, int frames_from_samples, int64_t audio_stride) {
  int s = blockIdx.y;
  int nf = n_frames[s];
  if (frames_from_samples) nf = 1 + min(nf, (int)audio_stride) / kHop;
  float mean = mean_std[s * ms_stride + 0], sd = mean_std[s * ms_stride + 1];
  int64_t total = (int64_t)kBins * out_stride;
  float* p = spect + (int64_t)s * total;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int t = (int)(i % out_stride);
    if (t < nf) p[i] = (p[i] - mean) / sd;
  }
}

static int ensure_tables() {
  int dev = 0;
  DSB_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_tab_mu);
  if (dev < 0 || dev >= 64) return set_error(DSB_ERR_UNSUPPORTED, "spectrogram: device ordinal %d", dev);
  if (!g_tab_done[dev]) {
    const cudaError_t e = init_tables();
    if (e != cudaSuccess)
      return set_error(DSB_ERR_CUDA, "spectrogram table upload failed: %s", cudaGetErrorString(e));
    g_tab_done[dev] = true;
  }
  return 0;
}

template <int MODE, typename T, bool S16 = false>
static int launch_spec(const void* audio, int channels, int64_t audio_stride, const int32_t* d_n, int B, float* out,
                       int64_t out_stride, float* mean_std, double* partials, int n_partials, int tiles, int center,
                       int normalize, int window, cudaStream_t st) {
  // per device and cheap (a host-side table update): set on every launch rather than once per process
  DSB_CUDA(cudaFuncSetAttribute(spectrogram_kernel<MODE, T, S16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)sizeof(SpecSmemT<T>)));
  dim3 grid(tiles, B);
  spectrogram_kernel<MODE, T, S16><<<grid, kWarps * 32, sizeof(SpecSmemT<T>), st>>>(audio, channels, audio_stride, d_n, out, out_stride,
                                                                      mean_std, partials, n_partials, center,
                                                                      normalize, window);
  DSB_CHECK_LAUNCH();
  return 0;
}

}  // namespace dsb

using namespace dsb;

extern "C" int dsb_spectrogram_num_frames(int n_samples) { return 1 + n_samples / kHop; }
extern "C" int dsb_spectrogram_partials(int max_frames) { return cdiv(max_frames > 0 ? max_frames : 1, kFT); }


static int spectrogram_offline(const void* audio, bool s16, int channels, int64_t audio_stride,
                               const int32_t* n_samples, int B, int max_samples, float* out, int64_t out_stride,
                               float* mean_std, double* partials, int flags, void* stream) {
  const int normalize = flags & DSB_SPECT_NORMALIZE;
  const bool fast = (flags & DSB_SPECT_FAST_FFT) != 0;
  const int window = (flags & DSB_SPECT_WINDOW_MASK) >> DSB_SPECT_WINDOW_SHIFT;
  DSB_REQUIRE(audio && n_samples && out && partials && B > 0, "dsb_spectrogram: null argument or B <= 0");
  DSB_REQUIRE(max_samples >= 1 && max_samples <= audio_stride, "dsb_spectrogram: max_samples %d out of range",
              max_samples);
  DSB_REQUIRE(channels >= 1 && channels <= 8, "dsb_spectrogram: channels %d", channels);
  if (int e = ensure_tables()) return e;
  cudaStream_t st = (cudaStream_t)stream;
  const int max_frames = 1 + max_samples / kHop;
  DSB_REQUIRE(out_stride >= max_frames, "dsb_spectrogram: out_stride %lld < frames %d", (long long)out_stride,
              max_frames);
  const int n_partials = dsb_spectrogram_partials((int)out_stride);
  const int tiles_all = cdiv((int)out_stride, kFT);
  ProfScope scope(ST_SPECT, st);
  // one FFT pass writes log1p|X| and per-tile (sum, sum of squares); the normalisation is a cheap
  // elementwise pass over data that is still in L2 (the FFT is instruction-bound, not HBM-bound, so
  // recomputing it for the second pass would double the kernel time)
  int e;
  if (s16)
    e = fast ? launch_spec<MODE_RAW, float, true>(audio, channels, audio_stride, n_samples, B, out, out_stride, nullptr,
                                                  partials, n_partials, tiles_all, 1, 0, window, st)
             : launch_spec<MODE_RAW, double, true>(audio, channels, audio_stride, n_samples, B, out, out_stride, nullptr,
                                                   partials, n_partials, tiles_all, 1, 0, window, st);
  else
    e = fast ? launch_spec<MODE_RAW, float, false>(audio, 1, audio_stride, n_samples, B, out, out_stride, nullptr,
                                                   partials, n_partials, tiles_all, 1, 0, window, st)
             : launch_spec<MODE_RAW, double, false>(audio, 1, audio_stride, n_samples, B, out, out_stride, nullptr,
                                                    partials, n_partials, tiles_all, 1, 0, window, st);
  if (e) return e;
  if (!normalize) return 0;
  stream_stats_kernel<<<cdiv(B, 128), 128, 0, st>>>(partials, n_partials, n_samples, nullptr, B, 1, normalize,
                                                   mean_std, audio_stride);
  DSB_CHECK_LAUNCH();
  int64_t total = (int64_t)kBins * out_stride;
  int64_t gx = cdiv64(total, 1024);
  if (gx > 256) gx = 256;
  stream_normalize_kernel<<<dim3((unsigned)gx, B), 256, 0, st>>>(out, out_stride, n_samples,
                                                                reinterpret_cast<const float*>(partials),
                                                                (int64_t)n_partials * 4, 1, audio_stride);
  DSB_CHECK_LAUNCH();
  return 0;
}

extern "C" int dsb_spectrogram_f32(const float* audio, int64_t audio_stride, const int32_t* n_samples, int B,
                                   int max_samples, float* out, int64_t out_stride, float* mean_std,
                                   double* partials, int flags, void* stream) {
  return spectrogram_offline(audio, false, 1, audio_stride, n_samples, B, max_samples, out, out_stride, mean_std,
                             partials, flags, stream);
}

extern "C" int dsb_spectrogram_s16(const int16_t* audio, int channels, int64_t audio_stride, const int32_t* n_samples,
                                   int B, int max_samples, float* out, int64_t out_stride, float* mean_std,
                                   double* partials, int flags, void* stream) {
  return spectrogram_offline(audio, true, channels, audio_stride, n_samples, B, max_samples, out, out_stride, mean_std,
                             partials, flags, stream);
}

extern "C" int dsb_spectrogram_stream_f32(const float* audio, int64_t audio_stride, const int32_t* n_samples, int S,
                                          int max_samples, float* out, int64_t out_stride, double* stats,
                                          double* partials, int flags, void* stream) {
  DSB_REQUIRE(audio && n_samples && out && partials && stats && S > 0, "dsb_spectrogram_stream_f32: null argument");
  DSB_REQUIRE(max_samples >= kNfft && max_samples <= audio_stride,
              "dsb_spectrogram_stream_f32: max_samples %d out of range", max_samples);
  if (int e = ensure_tables()) return e;
  cudaStream_t st = (cudaStream_t)stream;
  const int max_frames = 1 + (max_samples - kNfft) / kHop;
  DSB_REQUIRE(out_stride >= max_frames, "dsb_spectrogram_stream_f32: out_stride too small");
  const int n_partials = dsb_spectrogram_partials((int)out_stride);
  if (int e = launch_spec<MODE_RAW, double, false>(audio, 1, audio_stride, n_samples, S, out, out_stride, nullptr, partials,
                                    n_partials, cdiv((int)out_stride, kFT), 0, 0,
                                    (flags & DSB_SPECT_WINDOW_MASK) >> DSB_SPECT_WINDOW_SHIFT, st))
    return e;
  stream_stats_kernel<<<cdiv(S, 128), 128, 0, st>>>(partials, n_partials, n_samples, stats, S, 0, 1, nullptr,
                                                   audio_stride);
  DSB_CHECK_LAUNCH();
  return 0;
}

// Running statistics of InferenceSpectrogramAudioParser.parse_audio (parsers.py:146-157), one thread per stream:
// alpha += inc; input_mean = (input_mean + mean)/2; input_std = (input_std + std)/2; blended with the dataset
// constants while alpha < 1 (all in float64, like the Python floats of the reference).
__global__ void stream_running_stats_kernel(double* __restrict__ run, const double* __restrict__ stats,
                                            float* __restrict__ mean_std, int S, double dataset_mean,
                                            double dataset_std, double alpha_inc) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  const double alpha = run[s * 3 + 2] + alpha_inc;
  const double im = (run[s * 3 + 0] + stats[s * 2 + 0]) / 2;
  const double is = (run[s * 3 + 1] + stats[s * 2 + 1]) / 2;
  run[s * 3 + 0] = im;
  run[s * 3 + 1] = is;
  run[s * 3 + 2] = alpha;
  double mean = im, sd = is;
  if (alpha < 1.0) {
    mean = im * alpha + (1 - alpha) * dataset_mean;
    sd = is * alpha + (1 - alpha) * dataset_std;
  }
  mean_std[s * 2 + 0] = (float)mean;
  mean_std[s * 2 + 1] = (float)sd;
}

extern "C" int dsb_spectrogram_stream_running_stats(double* run, const double* stats, float* mean_std, int S,
                                                    double dataset_mean, double dataset_std, double alpha_increment,
                                                    void* stream) {
  DSB_REQUIRE(run && stats && mean_std && S > 0, "dsb_spectrogram_stream_running_stats: null argument");
  stream_running_stats_kernel<<<cdiv(S, 128), 128, 0, (cudaStream_t)stream>>>(run, stats, mean_std, S, dataset_mean,
                                                                             dataset_std, alpha_increment);
  DSB_CHECK_LAUNCH();
  return 0;
}

extern "C" int dsb_spectrogram_stream_normalize(float* spect, int64_t out_stride, const int32_t* n_frames, int S,
                                                const float* mean_std, void* stream) {
  DSB_REQUIRE(spect && n_frames && mean_std && S > 0, "dsb_spectrogram_stream_normalize: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t total = (int64_t)kBins * out_stride;
  int64_t gx = cdiv64(total, 256); if (gx > 64) gx = 64;
  dim3 grid((unsigned)gx, S);
  stream_normalize_kernel<<<grid, 256, 0, st>>>(spect, out_stride, n_frames, mean_std, 2, 0, 0);
  DSB_CHECK_LAUNCH();
  return 0;
}
