"""Oracle (test infrastructure): CPU restatement of the reference spectrogram parsers.

Follows /root/reference/danspeech/audio/parsers.py:
  * SpectrogramAudioParser.parse_audio            parsers.py:50-72
  * InferenceSpectrogramAudioParser.parse_audio   parsers.py:101-163, reset :165-170

The STFT itself lives in librosa (third party, un-vendored, unpinned:
requirements.txt:7) -- "parity unpinned".  ``stft`` / ``magphase`` below restate
the librosa 0.7-era algorithm the reference was written against:
``center=True`` reflect padding of n_fft//2, a *callable* window evaluated as
``window(win_length)`` (scipy's symmetric Hamming), frames of n_fft every
hop_length, float64 FFT stored as complex64.
"""
import numpy as np
import torch


def hamming_sym(n):
    """scipy.signal.hamming(n) (sym=True): 0.54 - 0.46 cos(2 pi k / (n-1))."""
    k = np.arange(n, dtype=np.float64)
    return 0.54 - 0.46 * np.cos(2.0 * np.pi * k / (n - 1))


def _cos_window(coeffs):
    def w(n):
        x = 2.0 * np.pi * np.arange(n, dtype=np.float64) / (n - 1)
        return sum(((-1) ** i) * a * np.cos(i * x) for i, a in enumerate(coeffs))
    return w


def bartlett_sym(n):
    """scipy.signal.bartlett(n): 1 - |2k/(n-1) - 1|."""
    k = np.arange(n, dtype=np.float64)
    return 1.0 - np.abs(2.0 * k / (n - 1) - 1.0)


# the four windows the reference offers through audio_conf["window"] (parsers.py:9-10), all symmetric
_WINDOWS = {"hamming": hamming_sym, "hann": _cos_window((0.5, 0.5)), "blackman": _cos_window((0.42, 0.5, 0.08)),
            "bartlett": bartlett_sym}


def stft(y, n_fft=320, hop_length=160, win_length=320, window=hamming_sym, center=True):
    """librosa.stft restatement (call sites parsers.py:59-60 and :138-139)."""
    y = np.asarray(y, dtype=np.float64)
    w = window(win_length) if callable(window) else np.asarray(window, dtype=np.float64)
    if center:
        y = np.pad(y, n_fft // 2, mode="reflect")
    n_frames = 1 + (len(y) - n_fft) // hop_length
    idx = np.arange(n_fft)[:, None] + hop_length * np.arange(n_frames)[None, :]
    frames = y[idx] * w[:, None]                      # [n_fft, n_frames], float64
    D = np.fft.fft(frames, axis=0)[: 1 + n_fft // 2]  # float64 math
    return D.astype(np.complex64)                     # librosa stores complex64


def magphase(D):
    """librosa.magphase restatement: magnitude (float32) and unit phase."""
    mag = np.abs(D)
    with np.errstate(divide="ignore", invalid="ignore"):
        phase = np.where(mag == 0, 1.0, D / np.where(mag == 0, 1, mag))
    return mag, phase


class SpectrogramOracle:
    """parsers.py:37-72."""

    def __init__(self, audio_config=None):
        cfg = audio_config or {}
        self.normalize = cfg.get("normalize", True)
        self.sampling_rate = cfg.get("sampling_rate", 16000)
        self.window = _WINDOWS[cfg.get("window", "hamming")]
        self.n_fft = int(self.sampling_rate * cfg.get("window_size", 0.02))
        self.hop_length = int(self.sampling_rate * cfg.get("window_stride", 0.01))

    def parse_audio(self, recording):
        D = stft(recording, self.n_fft, self.hop_length, self.n_fft, self.window)
        spect = np.log1p(np.abs(D))
        spect = torch.FloatTensor(spect)
        if self.normalize:
            mean = spect.mean()
            std = spect.std()            # unbiased
            spect.add_(-mean)
            spect.div_(std)
        return spect


class StreamingSpectrogramOracle:
    """parsers.py:75-170 (stateful, one instance per stream)."""

    dataset_mean = 5.492418704733003
    dataset_std = 1.7552755216970917

    def __init__(self, audio_config=None):
        cfg = audio_config or {}
        self.sampling_rate = cfg.get("sampling_rate", 16000)
        self.window = _WINDOWS[cfg.get("window", "hamming")]
        self.n_fft = int(self.sampling_rate * cfg.get("window_size", 0.02))
        self.hop_length = int(self.sampling_rate * cfg.get("window_stride", 0.01))
        self.reset()

    def reset(self):
        self.buffer = None
        self.has_buffer = False
        self.input_mean = 0
        self.input_std = 0
        self.alpha = 0

    def parse_audio(self, part, is_last=False):
        if is_last and len(part) < self.n_fft:
            self.reset()
            return []
        part = np.asarray(part, dtype=np.float64)
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

        D = stft(part, self.n_fft, self.hop_length, self.n_fft, self.window, center=False)
        spect = np.log1p(np.abs(D))
        self.alpha += 0.1
        self.input_mean = (self.input_mean + np.mean(spect)) / 2
        self.input_std = (self.input_std + np.std(spect)) / 2   # biased (numpy)
        if self.alpha < 1.0:
            mean = self.input_mean * self.alpha + (1 - self.alpha) * self.dataset_mean
            std = self.input_std * self.alpha + (1 - self.alpha) * self.dataset_std
        else:
            mean, std = self.input_mean, self.input_std
        spect -= mean
        spect /= std
        return torch.FloatTensor(spect)
