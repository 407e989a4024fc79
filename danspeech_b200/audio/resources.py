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


# ----------------------------------------------------------------------------------------------------------------
# Stream sources for the listening API (Recognizer.listen / listen_stream / streaming / real_time_streaming).
# The protocol is the reference's (danspeech/audio/resources.py:174-178, :324-493): a context manager exposing
# ``stream.read(n_samples) -> bytes`` (b"" at the end of the stream), ``chunk``, ``sampling_rate`` and
# ``sampling_width`` (bytes per sample).
class SpeechSource(object):
    """Base class of the audio sources the listening API accepts."""
    stream = None
    chunk = 1024
    sampling_rate = 16000
    sampling_width = 2


class _BytesStream(object):
    def __init__(self, pcm_bytes, width, seconds_per_sample=0.0):
        self._data, self._width, self._pos = pcm_bytes, width, 0
        self._pace, self._due = seconds_per_sample, None

    def read(self, n_samples=-1):
        """The next ``n_samples`` samples as bytes; a short (or empty) string at the end, like a file.  A paced
        stream blocks until the samples "have been recorded", like a microphone."""
        end = len(self._data) if n_samples is None or n_samples < 0 else self._pos + n_samples * self._width
        out = self._data[self._pos:end]
        self._pos = min(end, len(self._data))
        if self._pace and len(out):
            import time
            now = time.monotonic()
            self._due = (now if self._due is None else self._due) + self._pace * (len(out) // self._width)
            if self._due > now:
                time.sleep(self._due - now)
        return out

    def tell(self):
        """Samples handed out so far."""
        return self._pos // self._width

    def close(self):
        self._pos = len(self._data)


class ArraySource(SpeechSource):
    """In-memory mono PCM as a stream source (an addition: serving from buffers, simulations and the tests).
    ``samples`` is an int16 array (or anything castable to it without loss); ``realtime=x`` paces the stream at x times
    real time, the way a microphone delivers it."""

    def __init__(self, samples, sampling_rate=16000, chunk_size=1024, realtime=None):
        a = np.asarray(samples)
        if a.ndim != 1:
            raise ValueError("ArraySource takes mono audio [n]")
        self._pcm = np.ascontiguousarray(a.astype("<i2")).tobytes()
        self.sampling_rate, self.chunk, self.sampling_width = int(sampling_rate), int(chunk_size), 2
        self.realtime = realtime     # None: as fast as it is read; x: delivered at x times real time
        self.stream = None

    def __enter__(self):
        pace = 1.0 / (self.sampling_rate * float(self.realtime)) if self.realtime else 0.0
        self.stream = _BytesStream(self._pcm, self.sampling_width, pace)
        return self

    def __exit__(self, exc_type, exc_value, traceback):
        self.stream = None
        return False


class SpeechFile(ArraySource):
    """A PCM WAV file as a stream source; channels are mixed down as ``load_audio`` does (clip(L+R))."""

    def __init__(self, filepath, chunk_size=4096):
        self.filepath = filepath
        data, width, rate = _read_wav(filepath)
        if width != 2:
            raise ValueError("SpeechFile streams 16-bit PCM; %s has %d-byte samples" % (filepath, width))
        mono = data[:, 0] if data.shape[1] == 1 else np.clip(data.astype(np.int64).sum(axis=1), -32768, 32767)
        ArraySource.__init__(self, mono, sampling_rate=rate, chunk_size=chunk_size)
        self.duration = len(mono) / float(rate)


class Microphone(SpeechSource):
    """Live capture needs PyAudio, which is not part of this package's environment: constructing a Microphone
    without it fails loudly.  With PyAudio present it opens a 16-bit mono input stream with the reference's defaults
    (danspeech/audio/resources.py:385-423)."""

    @staticmethod
    def _pyaudio_module():
        try:
            import pyaudio
        except ImportError as e:
            raise ImportError("danspeech_b200.audio.Microphone needs the PyAudio package (not installed); "
                              "use ArraySource / SpeechFile, or push chunks into streaming.MultiStreamRecognizer") from e
        return pyaudio

    @staticmethod
    def list_microphone_names():
        """Names of the audio devices, indexed as ``device_index`` expects them."""
        audio = Microphone._pyaudio_module().PyAudio()
        try:
            return [audio.get_device_info_by_index(i).get("name") for i in range(audio.get_device_count())]
        finally:
            audio.terminate()

    def __init__(self, device_index=None, sampling_rate=16000, chunk_size=1024):
        self._pyaudio = self._pyaudio_module()
        self.device_index, self.sampling_rate, self.chunk, self.sampling_width = device_index, sampling_rate, chunk_size, 2
        self.stream = self._audio = None

    def __enter__(self):
        self._audio = self._pyaudio.PyAudio()
        raw = self._audio.open(input_device_index=self.device_index, channels=1, format=self._pyaudio.paInt16,
                               rate=self.sampling_rate, frames_per_buffer=self.chunk, input=True)
        mic = self

        class _Live(object):
            def read(self, n_samples):
                return raw.read(n_samples, exception_on_overflow=False)

            def close(self):
                raw.stop_stream()
                raw.close()
        self.stream = _Live()
        return mic

    def __exit__(self, exc_type, exc_value, traceback):
        try:
            self.stream.close()
        finally:
            self.stream = None
            self._audio.terminate()
        return False


class AudioData(object):
    """Mono PCM bytes plus their format (the subset of danspeech/audio/resources.py:495-640 the path uses)."""

    def __init__(self, frame_data, sample_rate, sample_width):
        if sample_rate <= 0 or int(sample_width) != sample_width or not 1 <= sample_width <= 4:
            raise AssertionError("AudioData needs a positive sample rate and a sample width of 1..4 bytes")
        self.frame_data, self.sample_rate, self.sample_width = frame_data, sample_rate, int(sample_width)

    def get_raw_data(self, convert_rate=None, convert_width=None):
        if (convert_rate not in (None, self.sample_rate)) or (convert_width not in (None, self.sample_width)):
            raise NotImplementedError("rate / width conversion is not part of the B200 path (16 kHz PCM in, as the "
                                      "models require)")
        return self.frame_data

    def get_segment(self, start_ms=None, end_ms=None):
        """The samples between two times (milliseconds; None = the respective end) as a new AudioData."""
        if start_ms is not None and start_ms < 0:
            raise AssertionError("start_ms must not be negative")
        if end_ms is not None and end_ms < (start_ms or 0):
            raise AssertionError("end_ms must not lie before start_ms")
        to_byte = lambda ms: int((ms * self.sample_rate * self.sample_width) // 1000)   # noqa: E731
        lo = 0 if start_ms is None else to_byte(start_ms)
        hi = len(self.frame_data) if end_ms is None else to_byte(end_ms)
        return AudioData(self.frame_data[lo:hi], self.sample_rate, self.sample_width)

    def get_wav_data(self, convert_rate=None, convert_width=None):
        """The audio as the bytes of a mono PCM WAV file."""
        import io
        raw = self.get_raw_data(convert_rate, convert_width)
        with io.BytesIO() as buf:
            with wave.open(buf, "wb") as w:
                w.setnchannels(1)
                w.setsampwidth(self.sample_width)
                w.setframerate(self.sample_rate)
                w.writeframes(raw)
            return buf.getvalue()

    def get_array_data(self, convert_rate=None, convert_width=None):
        """float64 array at raw integer sample scale -- what ``Recognizer.recognize`` takes."""
        raw = self.get_raw_data(convert_rate, convert_width)
        if self.sample_width == 3:
            raise ValueError("24-bit PCM is not supported")
        if len(raw) % self.sample_width:
            raise ValueError("The length of data is not a multiple of the sample width")
        if self.sample_width == 1:
            # the reference re-centres unsigned 8-bit PCM with audioop.bias(-128) and then reads the bytes back as
            # UNSIGNED (resources.py:565-567, :169-171): negative samples wrap to 128..255.  Kept as is.
            return ((np.frombuffer(raw, dtype=np.uint8).astype(np.int64) - 128) % 256).astype(float)
        return np.frombuffer(raw, dtype="<i%d" % self.sample_width).astype(float)
