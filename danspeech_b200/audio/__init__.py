from .resources import (ArraySource, AudioData, Microphone, SpeechFile, SpeechSource, load_audio,  # noqa: F401
                        load_audio_wavPCM)


def __getattr__(name):   # the parsers need the native library's bindings: imported on first use
    if name in ("SpectrogramAudioParser", "InferenceSpectrogramAudioParser", "AudioParser"):
        from . import parsers
        return getattr(parsers, name)
    raise AttributeError("module %r has no attribute %r" % (__name__, name))
