"""Exception types raised through the recognizer API.

The class names are part of the drop-in surface (user code catches them by name; the reference keeps them in
danspeech/errors/recognizer_errors.py).  Here they share one base class so that callers can also catch
``RecognizerError`` for anything this package raises on purpose; every one is still an ``Exception`` subclass.
"""


class RecognizerError(Exception):
    """Base of the exceptions below (an addition; the reference has no common base)."""


def _error(name, doc):
    return type(name, (RecognizerError,), {"__doc__": doc, "__module__": __name__})


ModelNotInitialized = _error(
    "ModelNotInitialized",
    "A language model or a transcription was requested before an acoustic model was set "
    "(DanSpeechRecognizer.py:39-40, Recognizer.py:72-75 of the reference).")
NoDataInBuffer = _error("NoDataInBuffer", "The streaming buffer was read while empty.")
WrongUsageOfListen = _error("WrongUsageOfListen", "listen()/listen_stream() was called outside a source context.")
WaitTimeoutError = _error("WaitTimeoutError", "Listening timed out before a phrase started.")
UnknownValueError = _error("UnknownValueError", "The audio could not be transcribed.")
RequestError = _error("RequestError", "A recognition request could not be carried out.")
ArgumentMissingForOption = _error("ArgumentMissingForOption", "An option was selected without the argument it needs.")

__all__ = ["RecognizerError", "ModelNotInitialized", "NoDataInBuffer", "WrongUsageOfListen", "WaitTimeoutError",
           "UnknownValueError", "RequestError", "ArgumentMissingForOption"]
