"""danspeech_b200: B200-native implementation of DanSpeech's inference hot path
(audio -> log-spectrogram -> DeepSpeech2 -> CTC decode), drop-in for the reference API
(danspeech/__init__.py:5-22 exports Recognizer and DanSpeechRecognizer)."""
from .DanSpeechRecognizer import DanSpeechRecognizer  # noqa: F401
from .Recognizer import Recognizer  # noqa: F401

__version__ = "0.1.0"
