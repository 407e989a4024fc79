"""Exception types of the model path.  The names are part of the drop-in surface (the reference keeps them in
danspeech/errors/model_errors.py); they share the ``ModelError`` base, an addition."""


class ModelError(Exception):
    """Base of the model-path exceptions."""


class ConvError(ModelError):
    """conv_layers outside 1..3 (raised where the reference raises it: deepspeech/model.py:344-348)."""


class FreezingMoreLayersThanExist(ModelError):
    """Kept for name compatibility: a training-side error, never raised on the inference path."""


class ModelDoesNotExistError(ModelError):
    """A pretrained-model name that `get_model_from_string` does not know."""


__all__ = ["ModelError", "ConvError", "FreezingMoreLayersThanExist", "ModelDoesNotExistError"]
