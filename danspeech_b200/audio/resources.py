"""Minimal audio-file ingest (WAV PCM) feeding the hot path.

Mirrors the *results* of danspeech/audio/resources.py ``load_audio`` (:22-61, via SpeechFile /
AudioData.get_array_data :630-640) and ``load_audio_wavPCM`` (:64-82) for PCM WAV input:
``load_audio`` mixes stereo down as clip(L+R) (audioop.tomono(buf, width, 1, 1), resources.py:302-303,
quirk Q1), ``load_audio_wavPCM`` averages the channels.  Microphone / FLAC / AIFF handling is out of
scope.
"""
import wave

import numpy as np

_DTYPES = {1: np.uint8, 2: np.int16, 4: np.int32}


def _read_wav(path):
    with wave.open(path, "rb") as w:
        nch, width, rate, nframes = w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()
        raw = w.readframes(nframes)
    if width not in _DTYPES:
        raise ValueError("unsupported sample width %d in %s" % (width, path))
    data = np.frombuffer(raw, dtype=_DTYPES[width]).reshape(-1, nch)
    if width == 1:   # unsigned 8-bit PCM -> signed, as audioop.bias(-128)
        data = data.astype(np.int16) - 128
    return data, width, rate


def load_audio(path, duration=None, offset=None):
    """PCM WAV -> float64 numpy array at raw integer sample scale; stereo -> clip(L+R)."""
    data, width, rate = _read_wav(path)
    if offset:
        data = data[int(offset * rate):]
    if duration:
        data = data[: int(duration * rate)]
    if data.shape[1] == 1:
        return data[:, 0].astype(float)
    lim = 1 << (8 * width - 1)
    mixed = np.clip(data.astype(np.int64).sum(axis=1), -lim, lim - 1)
    return mixed.astype(float)


def load_audio_wavPCM(path):
    """PCM WAV -> float64 array; multiple channels are averaged."""
    data, _, _ = _read_wav(path)
    if data.shape[1] == 1:
        return data[:, 0].astype(float)
    return data.mean(axis=1).astype(float)
