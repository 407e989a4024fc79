/*
 * danspeech_b200 -- C ABI of the B200-native DanSpeech inference hot path.
 *
 * Boundary: plain pointers and sizes only (no torch / C++ types).  Every
 * pointer documented as "device" must be a CUDA device pointer on the current
 * device; "host" pointers are ordinary host memory.  `stream` is a cudaStream_t
 * passed as void* (NULL = legacy default stream).  All functions return 0 on
 * success and a negative dsb_status on failure; dsb_last_error() returns a
 * thread-local message for the last failure.  Nothing here allocates behind the
 * caller's back on the hot path: outputs and workspaces are caller-provided.
 *
 * Each entry point names the reference interface it replaces
 * (paths relative to the reference checkout, danspeech/danspeech).
 */
#ifndef DANSPEECH_B200_H
#define DANSPEECH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  DSB_OK = 0,
  DSB_ERR_INVALID = -1,   /* bad argument */
  DSB_ERR_CUDA = -2,      /* CUDA runtime error (message has the cudaError string) */
  DSB_ERR_STATE = -3,     /* call order violated (e.g. forward before finalize) */
  DSB_ERR_WORKSPACE = -4, /* caller workspace too small */
  DSB_ERR_UNSUPPORTED = -5,
  DSB_ERR_IO = -6         /* LM file unreadable / malformed */
} dsb_status;

const char* dsb_last_error(void);
/* ABI version of this header; bumped on any signature change. */
int dsb_abi_version(void);
/* Number of kernels launched by this library since process start (bench.py "gpu_launches"). */
uint64_t dsb_kernel_launch_count(void);

/* Stage timers: when enabled, the library brackets each stage of the path with CUDA events on the
 * caller's stream.  Stages: 0 spectrogram, 1 conv stack, 2 RNN input projections, 3 RNN recurrence,
 * 4 lookahead/fc/softmax, 5 greedy decode, 6 beam decode, 7 direction sum + layout change between recurrent layers
 * (bf16 mode).  dsb_profile_read synchronises on the
 * recorded events and returns the summed milliseconds and the number of spans since the last reset. */
void dsb_profile_enable(int on);
void dsb_profile_reset(void);
int dsb_profile_read(int stage, double* total_ms, int* spans);

/* Diagnostic tuning knobs (process-wide integers; defaults are the production settings, nothing needs to be set).
 * They exist for A/B measurements and for tests that force a code path on small shapes; the reference has no
 * equivalent (its only knobs are constructor kwargs, danspeech/DanSpeechRecognizer.py:15-17).
 *   "rnn_in_flight"  1..3  batch groups of 64 sequences in flight per CTA in the persistent recurrence (default 3;
 *                          1 = one group of up to 128 rows at a time)
 *   "rnn_max_slots"  >= 0  cap on the independent CTA sets of the recurrence (0 = as many as fit, default)
 *   "rnn_ksplit"     0/1   1 = recurrence on CTA pairs that split K (csrc/rnn_ks.cu; parity-green, measured slower
 *                          than the default one-CTA-per-W_hh-slice kernel: DESIGN.md section 5), default 0
 *   "rnn_ring_gsz"   0..4  K chunks of 64 per slot of the recurrence's h ring (0 = default: two slots of four)
 *   "rnn_producers"  1..2  TMA producer warps of the recurrence (default 1; 2 = alternate ring slots, measured equal)
 * dsb_tune_set returns DSB_ERR_INVALID for an unknown key or an out-of-range value; dsb_tune_get returns the value
 * or -1 for an unknown key. */
int dsb_tune_set(const char* key, int value);
int dsb_tune_get(const char* key);

/* ------------------------------------------------------------------------- *
 * Spectrogram (replaces danspeech/audio/parsers.py:50-72,
 * SpectrogramAudioParser.parse_audio: librosa.stft(n_fft=320, hop=160, symmetric
 * Hamming, center/reflect) -> |.| -> log1p -> (S-mean)/unbiased-std per utterance).
 *
 *   audio          device f32 [B, audio_stride]; utterance b occupies samples [0, n_samples[b])
 *   n_samples      device i32 [B];  max_samples = max over b (host scalar, sizes the grid)
 *   out            device f32 [B, n_freq=161, out_stride]; frames t >= 1+n/160 are written as 0
 *   mean_std       device f32 [B, 2] (mean, std actually applied)  -- may be NULL
 *   partials       device f64 [B, dsb_spectrogram_partials(max_frames), 2] scratch
 *   flags          DSB_SPECT_NORMALIZE: audio_conf["normalize"] (parsers.py:25);
 *                  DSB_SPECT_FAST_FFT: run the transform in fp32 instead of fp64 (errors up to ~2e-3 in weak
 *                  bins: inside the 2e-2 bar of the bf16 mode, outside the 1e-4 bar of the fp32 mode);
 *                  DSB_SPECT_WINDOW_*: audio_conf["window"], one of the four scipy.signal windows the reference
 *                  offers (parsers.py:9-10), symmetric, 320 points; 0 = hamming (the default of deepspeech/utils.py:4)
 * ------------------------------------------------------------------------- */
#define DSB_SPECT_NORMALIZE 1
#define DSB_SPECT_FAST_FFT 2
#define DSB_SPECT_WINDOW_SHIFT 4
#define DSB_SPECT_WINDOW_MASK (3 << DSB_SPECT_WINDOW_SHIFT)
#define DSB_SPECT_WINDOW_HAMMING (0 << DSB_SPECT_WINDOW_SHIFT)
#define DSB_SPECT_WINDOW_HANN (1 << DSB_SPECT_WINDOW_SHIFT)
#define DSB_SPECT_WINDOW_BLACKMAN (2 << DSB_SPECT_WINDOW_SHIFT)
#define DSB_SPECT_WINDOW_BARTLETT (3 << DSB_SPECT_WINDOW_SHIFT)
int dsb_spectrogram_num_frames(int n_samples);                 /* 1 + n/160 */
int dsb_spectrogram_partials(int max_frames);                  /* scratch rows per utterance */
int dsb_spectrogram_f32(const float* audio, int64_t audio_stride, const int32_t* n_samples, int B,
                        int max_samples, float* out, int64_t out_stride, float* mean_std, double* partials,
                        int flags, void* stream);

/* Same, fed with interleaved 16-bit PCM as it comes out of a WAV file ("next" row SURVEY 8f-3: replaces
 * AudioData.get_array_data / _wav2array / audioop.tomono, danspeech/audio/resources.py:142-171,302-303,630-640):
 *   audio  device s16 [B, audio_stride frames, channels]; channels are mixed down as clip(sum), the
 *          audioop.tomono(buf, width, 1, 1) semantics of load_audio (quirk Q1).  Halves the host->device bytes. */
int dsb_spectrogram_s16(const int16_t* audio, int channels, int64_t audio_stride, const int32_t* n_samples, int B,
                        int max_samples, float* out, int64_t out_stride, float* mean_std, double* partials,
                        int flags, void* stream);

/* Streaming spectrogram (replaces parsers.py:101-163, InferenceSpectrogramAudioParser:
 * center=False STFT of an already assembled chunk, log1p, BIASED mean/std of the chunk
 * returned to the host, which owns the running-statistics recurrence; then
 * dsb_spectrogram_stream_normalize applies (S-mean)/std with the blended values).
 *   audio     device f32 [S, audio_stride] chunk per stream (carry-over already prepended)
 *   n_samples device i32 [S] (multiple of 160, >= 320); max_samples host scalar
 *   out       device f32 [S, 161, out_stride]   un-normalised log1p|STFT|
 *   stats     device f64 [S, 2]  (mean, biased std) of each chunk
 */
int dsb_spectrogram_stream_f32(const float* audio, int64_t audio_stride, const int32_t* n_samples, int S,
                               int max_samples, float* out, int64_t out_stride, double* stats, double* partials,
                               int flags /* DSB_SPECT_WINDOW_* only */, void* stream);
int dsb_spectrogram_stream_normalize(float* spect, int64_t out_stride, const int32_t* n_frames /* device */, int S,
                                     const float* mean_std /* device f32 [S,2] */, void* stream);
/* The running-statistics recurrence of parsers.py:146-157 kept on the device for S lock-step streams (the
 * single-stream parser keeps it on the host like the reference):
 *   run      device f64 [S,3] = (input_mean, input_std, alpha), updated in place; all zero == reset() (:92-99)
 *   stats    device f64 [S,2]   chunk statistics from dsb_spectrogram_stream_f32
 *   mean_std device f32 [S,2]   out: the values dsb_spectrogram_stream_normalize applies
 *   dataset_mean/std, alpha_increment: parsers.py:89-91 (5.492418704733003, 1.7552755216970917, 0.1) */
int dsb_spectrogram_stream_running_stats(double* run, const double* stats, float* mean_std, int S,
                                         double dataset_mean, double dataset_std, double alpha_increment,
                                         void* stream);

/* ------------------------------------------------------------------------- *
 * Acoustic model (replaces danspeech/deepspeech/model.py: DeepSpeech.__init__ :293-425,
 * forward :496-515, get_seq_lens :540-551, MaskConv :65-81, BatchRNN :114-122,
 * Lookahead :143-148, SequenceWise fc :414-420, InferenceBatchSoftmax :89-93,
 * and the streaming twins :156-284, :517-537).
 * ------------------------------------------------------------------------- */
typedef struct dsb_model dsb_model;

typedef enum { DSB_RNN_GRU = 0, DSB_RNN_LSTM = 1, DSB_RNN_TANH = 2 } dsb_rnn_type;
typedef enum { DSB_PREC_FP32 = 0, DSB_PREC_BF16 = 1 } dsb_precision;

typedef struct {
  int32_t conv_layers;     /* 1..3            (model.py:344-348 raises ConvError otherwise) */
  int32_t rnn_layers;
  int32_t rnn_hidden_size;
  int32_t rnn_type;        /* dsb_rnn_type    (supported_rnns, model.py:14-19) */
  int32_t bidirectional;
  int32_t context;         /* lookahead context, uni-directional models only */
  int32_t num_classes;     /* len(labels) */
  int32_t streaming;       /* streaming_inference_model (model.py:350-351) */
} dsb_model_desc;

int dsb_model_create(const dsb_model_desc* desc, dsb_model** out);
void dsb_model_destroy(dsb_model* m);
/* Hand one state-dict tensor (reference names, e.g. "rnns.3.rnn.weight_hh_l0_reverse",
 * SURVEY A.6) to the model.  `data` is a device f32 pointer that must stay valid until
 * dsb_model_finalize returns; it is not retained afterwards. */
int dsb_model_set_tensor(dsb_model* m, const char* name, const float* data, int64_t numel);
/* Folds eval-mode BatchNorm into the adjacent weights, packs kernel layouts. */
int dsb_model_finalize(dsb_model* m, int precision, void* stream);
int dsb_model_precision(const dsb_model* m);

/* get_seq_lens (model.py:540-551): T' = (T-1)/2 + 1 for every supported conv stack. */
int dsb_model_out_frames(const dsb_model* m, int T);
size_t dsb_forward_workspace_bytes(const dsb_model* m, int B, int T);
/* DeepSpeech.forward.
 *   spect       device f32 [B, 161, T]  zero-padded after normalisation
 *   lengths     host   i32 [B]          spectrogram frames per utterance; sorted descending
 *                                       (pack_padded_sequence contract, model.py:117) -> DSB_ERR_INVALID otherwise
 *   probs       device f32 [B, T', C]   softmax rows (rows t >= T'_b hold a valid softmax row, as upstream)
 *   out_lengths host   i32 [B]
 *   argmax      device i32 [B, T'] or NULL: fused argmax of each row (torch.max(probs,2), decoder.py:195)
 */
int dsb_forward(dsb_model* m, const float* spect, const int32_t* lengths, int B, int T,
                float* probs, int32_t* out_lengths, int32_t* argmax,
                void* workspace, size_t workspace_bytes, void* stream);
/* dsb_forward only enqueues work: it does not synchronise with the device.  The one failure that can happen on the
 * device -- a step barrier of the persistent recurrence timing out (it raises a flag instead of hanging the GPU) -- is
 * mirrored into pinned host memory behind the kernels of the call.  dsb_forward_status returns DSB_ERR_CUDA if any
 * forward of this model that has COMPLETED so far aborted (call it after synchronising with the stream, e.g. after
 * reading the transcripts back); the next dsb_forward checks the same flag.  The reference has no equivalent: its
 * forward is synchronous Python (danspeech/deepspeech/model.py:496-515). */
int dsb_forward_status(const dsb_model* m);

/* Tensor-core GEMM building block of the model path (also exported for diagnostics and tests):
 * C[M,N] (f32, row stride ldc) = A[M,K] (bf16, row stride lda) * W[N,K]^T (bf16, row stride ldw) + bias[N].
 * lda/ldw must be multiples of 8 elements and the operands 16-byte aligned (TMA requirements).
 * This is the BatchRNN input projection (the x_t*W_ih^T half of nn.GRU/LSTM/RNN, model.py:107-108). */
int dsb_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, float* C,
                  int64_t ldc, int M, int N, int K, void* stream);

/* Streaming (model.py:517-537 and the *Stream modules): S lock-step streams. */
typedef struct dsb_stream_state dsb_stream_state;
int dsb_stream_state_create(dsb_model* m, int n_streams, int max_chunk_frames, dsb_stream_state** out);
void dsb_stream_state_destroy(dsb_stream_state* s);
/* chunk device f32 [S,161,k]; probs device f32 [S, k_out_max, C]; *k_out receives the
 * number of frames emitted (0 when the lookahead layer is still buffering, i.e. the
 * reference returns None, model.py:529-530). */
int dsb_stream_max_out_frames(const dsb_stream_state* s, int k);
int dsb_streaming_forward(dsb_model* m, dsb_stream_state* s, const float* chunk, int k,
                          int is_first, int is_last, float* probs, int32_t* k_out, void* stream);

/* ------------------------------------------------------------------------- *
 * Greedy CTC (replaces danspeech/deepspeech/decoder.py:151-198, GreedyDecoder:
 * argmax -> drop blanks -> drop a symbol equal to the previous frame's symbol).
 *   probs    device f32 [B,T,C] (or NULL when `argmax` device i32 [B,T] is given)
 *   sizes    device i32 [B] or NULL (= T for every row, decoder.py:156)
 *   tokens   device i32 [B,T]  label indices kept;  offsets device i32 [B,T] their frame index
 *   out_len  device i32 [B]
 * ------------------------------------------------------------------------- */
int dsb_greedy_decode(const float* probs, const int32_t* argmax, const int32_t* sizes, int B, int T, int C,
                      int blank, int32_t* tokens, int32_t* offsets, int32_t* out_len, void* stream);

/* ------------------------------------------------------------------------- *
 * Beam CTC decoder (replaces the ctcdecode.CTCBeamDecoder object built at
 * decoder.py:99-100 and called at decoder.py:140; argument order as upstream).
 *   labels_utf8  n_labels NUL-separated UTF-8 strings, one per class
 *   lm_path      ARPA text LM path or NULL (no LM)
 * dsb_beam_decode: probs device f32 [B,T,C]; seq_lens host i32 [B];
 *   out_tokens/out_timesteps device i32 [B,beam,T]; out_scores device f32 [B,beam]
 *   (positive: -log p, lower is better); out_lens device i32 [B,beam].
 * ------------------------------------------------------------------------- */
typedef struct dsb_beam dsb_beam;
int dsb_beam_create(const char* labels_utf8, int n_labels, const char* lm_path, float alpha, float beta,
                    int cutoff_top_n, float cutoff_prob, int beam_width, int blank_id, int log_probs_input,
                    dsb_beam** out);
void dsb_beam_destroy(dsb_beam* d);
size_t dsb_beam_workspace_bytes(const dsb_beam* d, int B, int T);
int dsb_beam_decode(dsb_beam* d, const float* probs, const int32_t* seq_lens, int B, int T, int C,
                    int32_t* out_tokens, int32_t* out_timesteps, float* out_scores, int32_t* out_lens,
                    void* workspace, size_t workspace_bytes, void* stream);
/* LM introspection used by the host wrapper and tests. */
/* Host-only inspection of a language-model file through the same loaders dsb_beam_create uses (no CUDA call): the
 * LM order, the number of n-grams and vocabulary words, and an order-independent 64-bit digest of the loaded model.
 * An ARPA file and the KenLM probing binary built from it load into the same model and have the same digest. */
int dsb_lm_inspect(const char* path, int* order, int64_t* n_ngrams, int64_t* n_words, uint64_t* digest);
int dsb_beam_lm_order(const dsb_beam* d);
int dsb_beam_lm_is_char_based(const dsb_beam* d);
int64_t dsb_beam_lm_num_ngrams(const dsb_beam* d);

/* ------------------------------------------------------------------------- *
 * Energy voice-activity detection for S concurrent 16-bit streams ("next" row SURVEY 8f-4; replaces, per
 * stream, the phrase state machine of Recognizer.listen_stream, danspeech/Recognizer.py:218-324, and its
 * audioop.rms calls :275,:304).  One buffer of every stream per call; the host owns the audio (pre-roll of
 * non-speaking buffers, hand-over to the streaming model) and only reads one event code per stream.
 * ------------------------------------------------------------------------- */
typedef struct dsb_vad_state dsb_vad_state;
typedef enum {
  DSB_VAD_SILENCE = 0,        /* still waiting for speech (Recognizer.py:262-277): keep the buffer in the pre-roll */
  DSB_VAD_PHRASE_START = 1,   /* first loud buffer: the generator yields (False, pre-roll frames) (:283) */
  DSB_VAD_SPEECH = 2,         /* inside the phrase: yields (False, buffer) (:313) */
  DSB_VAD_PHRASE_END = 3,     /* pause longer than pause_buffers after a long-enough phrase: yields (True, buffer) (:321-324) */
  DSB_VAD_PHRASE_DROPPED = 4  /* phrase shorter than phrase_buffers: back to waiting (:316-318) */
} dsb_vad_event;
/* pause_buffers = ceil(pause_threshold / seconds_per_buffer), phrase_buffers = ceil(phrase_threshold / ...) (:244-247) */
int dsb_vad_state_create(int n_streams, int pause_buffers, int phrase_buffers, dsb_vad_state** out);
void dsb_vad_state_destroy(dsb_vad_state* v);
int dsb_vad_reset(dsb_vad_state* v, void* stream);
/*   chunks            device s16 [S, chunk_stride], chunk_samples valid samples per stream
 *   energy_threshold  device i32 [S] (Recognizer.energy_threshold, :44)
 *   energy_out        device i32 [S]  audioop.rms of the buffer = floor(sqrt(mean(x^2)))
 *   event_out         device i32 [S]  dsb_vad_event */
int dsb_vad_push_s16(dsb_vad_state* v, const int16_t* chunks, int64_t chunk_stride, int chunk_samples,
                     const int32_t* energy_threshold, int32_t* energy_out, int32_t* event_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DANSPEECH_B200_H */
