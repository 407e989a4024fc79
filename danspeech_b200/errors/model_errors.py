"""Exception types raised on the model path (names kept from danspeech/errors/model_errors.py:1-10)."""


class ConvError(Exception):
    """Unsupported number of convolutional layers (reference: deepspeech/model.py:344-348)."""


class ModelDoesNotExistError(Exception):
    pass


class FreezingMoreLayersThanExist(Exception):
    pass
