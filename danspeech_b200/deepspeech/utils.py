"""Default audio configuration (values of danspeech/deepspeech/utils.py:1-8)."""


def get_default_audio_config():
    return dict(normalize=True, sampling_rate=16000, window="hamming", window_stride=0.01, window_size=0.02)
