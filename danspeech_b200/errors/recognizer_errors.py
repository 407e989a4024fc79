"""Exception types of the recognizer API (names kept from danspeech/errors/recognizer_errors.py:1-21)."""


class WaitTimeoutError(Exception):
    pass


class RequestError(Exception):
    pass


class UnknownValueError(Exception):
    pass


class ModelNotInitialized(Exception):
    """LM given without an acoustic model (reference: DanSpeechRecognizer.py:39-40, Recognizer.py:72-75)."""


class WrongUsageOfListen(Exception):
    pass


class NoDataInBuffer(Exception):
    pass


class ArgumentMissingForOption(Exception):
    pass
