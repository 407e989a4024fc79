"""Audio -> normalised log-magnitude spectrogram on the GPU.

Mirrors the interface of danspeech/audio/parsers.py (AudioParser :13-34,
SpectrogramAudioParser :37-72, InferenceSpectrogramAudioParser :75-170); the arithmetic runs in
csrc/spectrogram.cu through the C ABI (dsb_spectrogram_f32 / dsb_spectrogram_stream_f32).

Differences that are deliberate and documented in INTEGRATION.md:
  * tensors are returned on the CUDA device (the reference returns CPU tensors that the engine
    immediately moves with ``.to(device)``, DanSpeechRecognizer.py:220-221);
  * ``parse_batch`` is an addition: the reference engine is batch-1 only.
"""
import threading
from abc import ABC, abstractmethod

import numpy as np
import torch

from .. import _native as N

# audio_conf["window"] -> DSB_SPECT_WINDOW_* (the four scipy.signal windows of parsers.py:9-10)
_WINDOW_FLAGS = {"hamming": 0 << 4, "hann": 1 << 4, "blackman": 2 << 4, "bartlett": 3 << 4}

_convert_pool = None


def _pool():
    """Helper threads for the float64 -> float32 staging copies (numpy releases the GIL inside copyto)."""
    global _convert_pool
    if _convert_pool is None:
        import concurrent.futures
        import os
        _convert_pool = concurrent.futures.ThreadPoolExecutor(max_workers=max(1, min(8, os.cpu_count() or 1)),
                                                              thread_name_prefix="dsb-stage")
    return _convert_pool


class AudioParser(ABC):
    """Audio-config holder (reference: parsers.py:13-34)."""

    def __init__(self, audio_config=None):
        self.audio_config = audio_config or {}
        self.normalize = self.audio_config.get("normalize", True)
        self.sampling_rate = self.audio_config.get("sampling_rate", 16000)
        self.window = self.audio_config.get("window", "hamming")
        self.window_stride = self.audio_config.get("window_stride", 0.01)
        self.window_size = self.audio_config.get("window_size", 0.02)
        self.n_fft = int(self.sampling_rate * self.window_size)
        self.hop_length = int(self.sampling_rate * self.window_stride)
        if self.window not in _WINDOW_FLAGS:
            raise KeyError(self.window)          # the reference's windows.get() gives None and dies inside librosa
        if self.n_fft != 320 or self.hop_length != 160:
            raise NotImplementedError(
                "danspeech_b200 spectrogram kernel is specialised for the DanSpeech framing "
                "(16 kHz, 20 ms window / 10 ms stride); got %r" % (self.audio_config,))
        self._window_flag = _WINDOW_FLAGS[self.window]

    @abstractmethod
    def parse_audio(self, recording):
        pass


def _as_f32(recording):
    a = np.ascontiguousarray(np.asarray(recording), dtype=np.float32).reshape(-1)
    return a


class SpectrogramAudioParser(AudioParser):
    """Offline parser (reference: parsers.py:37-72)."""

    def __init__(self, audio_config=None, device=None, fast_fft=False):
        super().__init__(audio_config)
        self.device = device
        # fp32 FFT (the bf16 model mode's 2e-2 bar) instead of the fp64 transform that matches the reference to 1e-4
        self.fast_fft = bool(fast_fft)
        self._staging, self._staging_busy, self._staging_lock = {}, {}, threading.Lock()

    def _dev(self):
        N.require_cuda()
        return torch.device(self.device or "cuda")

    def parse_device(self, audio, n_samples, max_samples, out=None, out_stride=None):
        """audio: cuda f32 [B, stride]; n_samples: cuda i32 [B].  Returns ([B,161,out_stride], mean_std[B,2])."""
        L = N.lib()
        B, stride = audio.shape
        frames = L.dsb_spectrogram_num_frames(int(max_samples))
        out_stride = out_stride or frames
        if out is None:
            out = torch.empty((B, 161, out_stride), dtype=torch.float32, device=audio.device)
        mean_std = torch.empty((B, 2), dtype=torch.float32, device=audio.device)
        partials = torch.empty((B, L.dsb_spectrogram_partials(out_stride), 2), dtype=torch.float64, device=audio.device)
        N.check(L.dsb_spectrogram_f32(N.ptr(audio), stride, N.ptr(n_samples), B, int(max_samples), N.ptr(out),
                                      out_stride, N.ptr(mean_std), N.ptr(partials),
                                      (1 if self.normalize else 0) | (2 if self.fast_fft else 0) | self._window_flag,
                                      N.current_stream()), "dsb_spectrogram_f32")
        return out, mean_std

    def stage_batch(self, recordings, slot=0):
        """Host half of ``parse_batch``: converts the recordings to float32 straight into a cached pinned staging
        buffer (one per ``slot``, grow-only; cudaHostAlloc per call would cost more than the GPU work).  Returns
        (pinned f32 view [B, stride], n_samples list).  Rows are only read up to their own length on the device."""
        ns = [int(np.asarray(r).size) for r in recordings]
        if not ns or min(ns) == 0:
            # the reference dies in np.pad(mode="reflect") inside librosa.stft (parsers.py:59)
            raise ValueError("can't extend empty axis 0 using modes other than 'constant' or 'empty'")
        stride = (max(ns) + 3) // 4 * 4
        need = len(ns) * stride
        with self._staging_lock:   # transcribe_batches stages on a helper thread
            busy = self._staging_busy.pop(slot, None)
            buf = self._staging.get(slot)
        if busy is not None:
            busy.synchronize()   # the previous host->device copy out of this buffer must have finished
        if buf is None or buf.numel() < need:
            buf = torch.empty((max(need, 1),), dtype=torch.float32, pin_memory=torch.cuda.is_available())
            with self._staging_lock:
                self._staging[slot] = buf
        host = buf[:need].view(len(ns), stride)
        host_np = host.numpy()

        def convert(i):
            np.copyto(host_np[i, : ns[i]], np.asarray(recordings[i]).reshape(-1), casting="unsafe")

        if len(ns) >= 8 and sum(ns) >= (1 << 20):
            list(_pool().map(convert, range(len(ns))))      # 64 x 15 s float64: 19 ms on one core, 5 ms on eight
        else:
            for i in range(len(ns)):
                convert(i)
        return host, ns

    def parse_batch(self, recordings):
        """List of 1-D arrays -> (spect cuda f32 [B,1,161,Tmax] zero-padded, lengths IntTensor[B] (CPU))."""
        self._dev()
        host, ns = self.stage_batch(recordings)
        return self.parse_packed(host, ns)

    def mark_staging_busy(self, host_audio, event):
        """``event`` completes when the host->device copy out of ``host_audio`` has finished; if that tensor lives in one
        of the staging buffers, stage_batch waits for the event before it reuses the buffer."""
        with self._staging_lock:
            for slot, buf in self._staging.items():
                if buf.untyped_storage().data_ptr() == host_audio.untyped_storage().data_ptr():
                    self._staging_busy[slot] = event

    def parse_packed(self, host_audio, n_samples):
        """host_audio: (pinned) CPU f32 tensor [B, stride] already sorted by length descending."""
        dev = self._dev()
        if host_audio.dtype != torch.float32 or host_audio.dim() != 2:
            raise ValueError("parse_packed expects a 2-D float32 tensor")
        ns = [int(v) for v in n_samples]
        if not ns or min(ns) <= 0:
            raise ValueError("can't extend empty axis 0 using modes other than 'constant' or 'empty'")
        audio = host_audio.to(dev, non_blocking=True)
        evt = torch.cuda.Event()
        evt.record()
        self.mark_staging_busy(host_audio, evt)
        n_dev = torch.tensor(ns, dtype=torch.int32).to(dev, non_blocking=True)
        out, _ = self.parse_device(audio, n_dev, max(ns))
        lengths = torch.IntTensor([1 + n // self.hop_length for n in ns])
        return out.view(len(ns), 1, 161, out.shape[2]), lengths

    def parse_pcm16(self, host_pcm, n_samples):
        """Interleaved 16-bit PCM straight from WAV files: host int16 tensor [B, frames] or [B, frames, channels]
        (ideally pinned), sorted by length descending.  Stereo is mixed down on the GPU as clip(L+R)."""
        dev = self._dev()
        if host_pcm.dtype != torch.int16 or host_pcm.dim() not in (2, 3):
            raise ValueError("parse_pcm16 expects an int16 tensor [B, frames] or [B, frames, channels]")
        ns = [int(v) for v in n_samples]
        B, stride = host_pcm.shape[0], host_pcm.shape[1]
        ch = host_pcm.shape[2] if host_pcm.dim() == 3 else 1
        pcm = host_pcm.contiguous().to(dev, non_blocking=True)
        n_dev = torch.tensor(ns, dtype=torch.int32).to(dev, non_blocking=True)
        L = N.lib()
        frames = L.dsb_spectrogram_num_frames(max(ns))
        out = torch.empty((B, 161, frames), dtype=torch.float32, device=dev)
        mean_std = torch.empty((B, 2), dtype=torch.float32, device=dev)
        partials = torch.empty((B, L.dsb_spectrogram_partials(frames), 2), dtype=torch.float64, device=dev)
        N.check(L.dsb_spectrogram_s16(N.ptr(pcm), ch, stride, N.ptr(n_dev), B, max(ns), N.ptr(out), frames,
                                      N.ptr(mean_std), N.ptr(partials),
                                      (1 if self.normalize else 0) | (2 if self.fast_fft else 0) | self._window_flag,
                                      N.current_stream()), "dsb_spectrogram_s16")
        lengths = torch.IntTensor([1 + n // self.hop_length for n in ns])
        return out.view(B, 1, 161, frames), lengths

    def parse_audio(self, recording):
        """1-D numpy array at raw int16 sample scale -> FloatTensor[161, 1 + n//160] (on the CUDA device)."""
        spect, _ = self.parse_batch([recording])
        return spect[0, 0]


class InferenceSpectrogramAudioParser(AudioParser):
    """Streaming parser with adaptive normalisation (reference: parsers.py:75-170).

    The sample carry-over buffer and the running-statistics recurrence are host state exactly as in
    the reference; the STFT/log1p/statistics/normalisation run on the GPU.
    """

    def __init__(self, audio_config=None, device=None):
        super().__init__(audio_config)
        self.device = device
        self.dataset_mean = 5.492418704733003
        self.dataset_std = 1.7552755216970917
        self.alpha_increment = 0.1
        self.reset()

    def reset(self):
        self.buffer = None
        self.has_buffer = False
        self.input_mean = 0
        self.input_std = 0
        self.alpha = 0

    def parse_audio(self, part_of_recording, is_last=False):
        if is_last and len(part_of_recording) < self.n_fft:
            self.reset()
            return []
        N.require_cuda()
        dev = torch.device(self.device or "cuda")
        part = np.asarray(part_of_recording, dtype=np.float64).reshape(-1)
        if self.has_buffer:
            part = np.concatenate((self.buffer, part), axis=None)
        extra = len(part) % self.hop_length
        if extra != 0:
            extra_arr = part[-extra:]
            part = part[:-extra]
        self.buffer = part[-self.hop_length:]
        if extra != 0:
            self.buffer = np.concatenate((self.buffer, extra_arr), axis=None)
        self.has_buffer = True

        L = N.lib()
        n = len(part)
        stride = (n + 3) // 4 * 4
        host = torch.zeros((1, stride), dtype=torch.float32)
        host[0, :n] = torch.from_numpy(part.astype(np.float32))
        audio = host.to(dev)
        n_dev = torch.tensor([n], dtype=torch.int32, device=dev)
        frames = 1 + (n - self.n_fft) // self.hop_length
        out = torch.empty((1, 161, frames), dtype=torch.float32, device=dev)
        stats = torch.empty((1, 2), dtype=torch.float64, device=dev)
        partials = torch.empty((1, L.dsb_spectrogram_partials(frames), 2), dtype=torch.float64, device=dev)
        N.check(L.dsb_spectrogram_stream_f32(N.ptr(audio), stride, N.ptr(n_dev), 1, n, N.ptr(out), frames,
                                             N.ptr(stats), N.ptr(partials), self._window_flag, N.current_stream()),
                "dsb_spectrogram_stream_f32")
        chunk_mean, chunk_std = stats[0].tolist()

        # running statistics, as parsers.py:146-157
        self.alpha += self.alpha_increment
        self.input_mean = (self.input_mean + chunk_mean) / 2
        self.input_std = (self.input_std + chunk_std) / 2
        if self.alpha < 1.0:
            mean = self.input_mean * self.alpha + (1 - self.alpha) * self.dataset_mean
            std = self.input_std * self.alpha + (1 - self.alpha) * self.dataset_std
        else:
            mean, std = self.input_mean, self.input_std
        ms = torch.tensor([[mean, std]], dtype=torch.float32, device=dev)
        nf = torch.tensor([frames], dtype=torch.int32, device=dev)
        N.check(L.dsb_spectrogram_stream_normalize(N.ptr(out), frames, N.ptr(nf), 1, N.ptr(ms), N.current_stream()),
                "dsb_spectrogram_stream_normalize")
        return out[0]
