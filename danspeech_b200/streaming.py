"""Many concurrent real-time streams on one GPU (BASELINE config 4).

The reference serves ONE microphone: ``Recognizer.listen_stream`` (Recognizer.py:218-324) cuts phrases with an
energy VAD and ``DanSpeechRecognizer.streaming_transcribe`` (DanSpeechRecognizer.py:144-216) feeds the chunks to
the streaming model, stitches the greedy transcripts and, at the end of a phrase, optionally re-decodes the whole
phrase with a bidirectional "secondary" model.  Its streaming state lives in module attributes, so N streams
need N recognisers.  Here the same per-stream behaviour runs for S lock-step streams with the state on the device:

* ``StreamVAD``             -- the phrase state machine of ``listen_stream`` for S 16-bit streams (csrc/vad.cu);
* ``MultiStreamRecognizer`` -- ``streaming_transcribe`` for S streams: batched streaming spectrogram with the
  running-statistics recurrence on the device, one ``dsb_streaming_forward`` per chunk, device greedy decode,
  the reference's transcript stitching per stream, and the secondary-model / LM final pass as ONE batched
  forward over all streams when the phrase ends.
"""
import ctypes
import math

import torch

from . import _native as N
from .deepspeech.decoder import GreedyDecoder
from .utils.stitch import stitch_transcript

SILENCE, PHRASE_START, SPEECH, PHRASE_END, PHRASE_DROPPED = 0, 1, 2, 3, 4


class StreamVAD:
    """Energy VAD with the parameters of ``Recognizer.__init__`` (Recognizer.py:42-56)."""

    def __init__(self, n_streams, energy_threshold=1000, pause_threshold=0.8, phrase_threshold=0.3,
                 non_speaking_duration=0.35, chunk=1024, sampling_rate=16000, device="cuda"):
        N.require_cuda()
        assert pause_threshold >= non_speaking_duration >= 0          # Recognizer.py:241
        self.n_streams = int(n_streams)
        self.chunk = int(chunk)
        seconds_per_buffer = float(chunk) / sampling_rate
        self.pause_buffer_count = int(math.ceil(pause_threshold / seconds_per_buffer))
        self.phrase_buffer_count = int(math.ceil(phrase_threshold / seconds_per_buffer))
        self.non_speaking_buffer_count = int(math.ceil(non_speaking_duration / seconds_per_buffer))
        self.device = torch.device(device)
        thr = torch.as_tensor(energy_threshold, dtype=torch.int32)
        self.energy_threshold = thr.expand(self.n_streams).contiguous().to(self.device)
        self._state = ctypes.c_void_p()
        N.check(N.lib().dsb_vad_state_create(self.n_streams, self.pause_buffer_count, self.phrase_buffer_count,
                                             ctypes.byref(self._state)), "dsb_vad_state_create")

    def __del__(self):
        st = getattr(self, "_state", None)
        if st:
            try:
                N.lib().dsb_vad_state_destroy(st)
            except Exception:
                pass
            self._state = None

    def reset(self):
        N.check(N.lib().dsb_vad_reset(self._state, N.current_stream()), "dsb_vad_reset")

    def push(self, chunks):
        """chunks: int16 tensor [S, n] (device, or host -> copied).  Returns (events, energy) int32 device tensors."""
        chunks = torch.as_tensor(chunks)
        if chunks.dtype != torch.int16 or chunks.dim() != 2 or chunks.size(0) != self.n_streams:
            raise ValueError("chunks must be int16 [n_streams, n]")
        chunks = chunks.to(self.device, non_blocking=True).contiguous()
        events = torch.empty(self.n_streams, dtype=torch.int32, device=self.device)
        energy = torch.empty(self.n_streams, dtype=torch.int32, device=self.device)
        N.check(N.lib().dsb_vad_push_s16(self._state, N.ptr(chunks), chunks.stride(0), chunks.size(1),
                                         N.ptr(self.energy_threshold), N.ptr(energy), N.ptr(events),
                                         N.current_stream()), "dsb_vad_push_s16")
        return events, energy


class BatchedStreamingParser:
    """``InferenceSpectrogramAudioParser`` (parsers.py:75-170) for S lock-step streams: the carry-over buffer
    is a device tensor and the running statistics never leave the GPU."""

    n_fft, hop = 320, 160
    dataset_mean, dataset_std, alpha_increment = 5.492418704733003, 1.7552755216970917, 0.1

    def __init__(self, n_streams, device="cuda", audio_config=None):
        N.require_cuda()
        from .audio.parsers import _WINDOW_FLAGS
        self.S = int(n_streams)
        self.device = torch.device(device)
        self._window_flag = _WINDOW_FLAGS[(audio_config or {}).get("window", "hamming")]
        self.reset()

    def reset(self):
        self.buffer = None
        self.run = torch.zeros((self.S, 3), dtype=torch.float64, device=self.device)

    def parse_audio(self, parts, is_last=False):
        """parts: float tensor/array [S, n] at int16 scale.  Returns a device tensor [S, 161, k] or None when the
        last part is too short for a frame (the reference returns [] and resets, parsers.py:112-114)."""
        parts = torch.as_tensor(parts)
        if parts.dim() != 2 or parts.size(0) != self.S:
            raise ValueError("parts must be [n_streams, n]")
        if is_last and parts.size(1) < self.n_fft:
            self.reset()
            return None
        x = parts.to(self.device, dtype=torch.float32, non_blocking=True)
        if self.buffer is not None:
            x = torch.cat((self.buffer, x), dim=1)
        n_all = x.size(1)
        extra = n_all % self.hop
        n = n_all - extra
        # carry: the last hop of the framed part plus the samples that did not fill a hop (parsers.py:123-133)
        self.buffer = x[:, n - self.hop:].clone()
        L = N.lib()
        audio = x[:, :n].contiguous()   # n is a multiple of the hop (160): rows stay 16-byte aligned
        frames = 1 + (n - self.n_fft) // self.hop
        n_dev = torch.full((self.S,), n, dtype=torch.int32, device=self.device)
        out = torch.empty((self.S, 161, frames), dtype=torch.float32, device=self.device)
        stats = torch.empty((self.S, 2), dtype=torch.float64, device=self.device)
        partials = torch.empty((self.S, L.dsb_spectrogram_partials(frames), 2), dtype=torch.float64, device=self.device)
        st = N.current_stream()
        N.check(L.dsb_spectrogram_stream_f32(N.ptr(audio), audio.stride(0), N.ptr(n_dev), self.S, n, N.ptr(out), frames,
                                             N.ptr(stats), N.ptr(partials), self._window_flag, st),
                "dsb_spectrogram_stream_f32")
        ms = torch.empty((self.S, 2), dtype=torch.float32, device=self.device)
        N.check(L.dsb_spectrogram_stream_running_stats(N.ptr(self.run), N.ptr(stats), N.ptr(ms), self.S,
                                                       self.dataset_mean, self.dataset_std, self.alpha_increment, st),
                "dsb_spectrogram_stream_running_stats")
        nf = torch.full((self.S,), frames, dtype=torch.int32, device=self.device)
        N.check(L.dsb_spectrogram_stream_normalize(N.ptr(out), frames, N.ptr(nf), self.S, N.ptr(ms), st),
                "dsb_spectrogram_stream_normalize")
        return out


class MultiStreamRecognizer:
    """``DanSpeechRecognizer.streaming_transcribe`` (DanSpeechRecognizer.py:144-216) for S lock-step streams.

    ``push(parts, is_first, is_last)`` takes one chunk per stream and returns a list of S strings with exactly the
    per-stream meaning of the reference call (string parts or the iterating transcript; on the last chunk the
    final transcript, re-decoded by the secondary model or the LM decoder when given)."""

    def __init__(self, streaming_model, n_streams, secondary_model=None, decoder=None, string_parts=True,
                 device="cuda"):
        N.require_cuda()
        self.S = int(n_streams)
        self.device = torch.device(device)
        self.model = streaming_model.to(self.device).eval()
        self.secondary_model = secondary_model.to(self.device).eval() if secondary_model is not None else None
        self.labels = self.model.labels
        self.greedy_decoder = GreedyDecoder(labels=self.labels, blank_index=self.labels.index("_"))
        self.decoder = decoder            # final-pass decoder (None or a GreedyDecoder == the reference's lm "greedy")
        self.string_parts = bool(string_parts)
        self.audio_parser = BatchedStreamingParser(self.S, device=self.device,
                                                   audio_config=getattr(self.model, "audio_conf", None))
        self.reset_streaming_params()

    def reset_streaming_params(self):
        self.iterating_transcript = [""] * self.S
        self.full_output = []
        self.spectrograms = []

    def push(self, parts, is_first, is_last):
        spect = self.audio_parser.parse_audio(parts, is_last)
        out = [""] * self.S
        if spect is not None:
            if self.secondary_model is not None:
                self.spectrograms.append(spect)
            probs = self.model(spect.view(self.S, 1, 161, spect.size(2)), is_first, is_last)
            if is_first:
                return out
            if self.decoder is not None and not isinstance(self.decoder, GreedyDecoder):
                self.full_output.append(probs)
            decoded = self.greedy_decoder.decode_strings(probs)
            for s in range(self.S):
                it, part = stitch_transcript(self.iterating_transcript[s], decoded[s])
                self.iterating_transcript[s] = it
                out[s] = part if self.string_parts else it
        if is_last:
            heard = [len(it) > 1 for it in self.iterating_transcript]
            final = list(self.iterating_transcript)
            if any(heard):
                if self.secondary_model is not None:
                    full = torch.cat(self.spectrograms, dim=2)
                    sizes = torch.IntTensor([full.size(2)] * self.S)
                    probs2, out_sizes = self.secondary_model(full.view(self.S, 1, 161, full.size(2)), sizes)
                    dec = self.decoder if self.decoder is not None else self.greedy_decoder
                    final = [d[0] for d in dec.decode(probs2)[0]]
                elif self.decoder is not None and not isinstance(self.decoder, GreedyDecoder):
                    final = [d[0] for d in self.decoder.decode(torch.cat(self.full_output, dim=1))[0]]
            out = [f if h else "" for f, h in zip(final, heard)]
            # like DanSpeechRecognizer.py:181-214 the state is reset only where something was heard: a stream whose
            # phrase decoded to fewer than two characters keeps its iterating transcript for the next phrase (the
            # phrase-level buffers are shared by the lock-step streams and are dropped as soon as any stream was heard)
            self.iterating_transcript = ["" if h else it for it, h in zip(self.iterating_transcript, heard)]
            if any(heard):
                self.full_output = []
                self.spectrograms = []
        return out
