"""Model factories with the reference's names (danspeech/pretrained_models/__init__.py:1-30).

The reference factories download a ``.pth`` release artefact into ``~/.danspeech/models`` and load it
(e.g. danspeech_primary.py:22-24, utils/data_utils.py:43-88).  There is no downloading here: a factory loads the
package the reference would have cached -- ``<cache_dir or ~/.danspeech/models>/<file>.pth`` -- or an explicit
``model_path``, and raises ``FileNotFoundError`` when neither exists, so that ``Recognizer(model=DanSpeechPrimary())``
can never silently transcribe with random weights.  Seeded random weights of the *named architecture* (SURVEY A.6;
what tests and bench.py use, since the artefacts cannot be fetched offline) are an explicit opt-in:
``build_model(name, seed=...)`` or ``DanSpeechPrimary(synthetic=True)`` / ``DanSpeechPrimary(seed=0)``.
``CustomModel(path)`` loads a real reference package.
"""
import os

import torch.nn as nn

from ..deepspeech.model import DeepSpeech, supported_rnns
from ..utils.synthetic import LABELS, MODEL_SHAPES, make_state_dict


def build_model(name, model_path=None, seed=0, rnn_type="gru", **overrides):
    if model_path:
        return DeepSpeech.load_model(model_path)
    cfg = dict(MODEL_SHAPES[name])
    cfg.update(overrides)
    model = DeepSpeech(model_name=name, rnn_type=supported_rnns[rnn_type], labels=LABELS,
                       rnn_hidden_size=cfg["rnn_hidden_size"], rnn_layers=cfg["rnn_layers"],
                       bidirectional=cfg["bidirectional"], context=cfg.get("context", 20),
                       conv_layers=cfg["conv_layers"],
                       streaming_inference_model=cfg.get("streaming_inference_model", False))
    model.load_state_dict(make_state_dict(rnn_type=rnn_type, seed=seed, **cfg))
    return model


# factory name -> cached file name, as published by the reference (pretrained_models/*.py)
_PUBLISHED = {"EnglishLibrispeech": "Librispeech.pth"}


def _factory(name):
    def make(cache_dir=None, model_path=None, seed=None, synthetic=False):
        if model_path:
            return DeepSpeech.load_model(model_path)
        if synthetic or seed is not None:
            return build_model(name, seed=seed or 0)
        root = cache_dir if cache_dir is not None else os.path.join(os.path.expanduser("~"), ".danspeech", "models")
        path = os.path.join(root, _PUBLISHED.get(name, name + ".pth"))
        if os.path.isfile(path):
            return DeepSpeech.load_model(path)
        raise FileNotFoundError(
            "%s not found; this package does not download models -- place the reference's release artefact there, "
            "pass model_path=..., or ask for seeded random weights of this architecture explicitly with "
            "%s(synthetic=True)" % (path, name))
    make.__name__ = make.__qualname__ = name
    make.__doc__ = ("%s: loads <cache_dir or ~/.danspeech/models>/%s (or ``model_path``); ``synthetic=True`` / ``seed=`` "
                    "builds the architecture with seeded random weights instead." % (name, _PUBLISHED.get(name, name + ".pth")))
    return make


DanSpeechPrimary = _factory("DanSpeechPrimary")
TestModel = _factory("TestModel")
Baseline = _factory("Baseline")
CPUStreamingRNN = _factory("CPUStreamingRNN")
GPUStreamingRNN = _factory("GPUStreamingRNN")
Folketinget = _factory("Folketinget")
TransferLearned = _factory("TransferLearned")
EnglishLibrispeech = _factory("EnglishLibrispeech")


def CustomModel(model_path):
    """danspeech/pretrained_models/custom_model.py:4-13."""
    return DeepSpeech.load_model(model_path)


def get_model_from_string(model_name):
    """danspeech/pretrained_models/__init__.py:12-30 (quirk Q5 kept: 'GPUStreamingRNN' -> CPUStreamingRNN)."""
    table = dict(DanSpeechPrimary=DanSpeechPrimary, TestModel=TestModel, Baseline=Baseline,
                 CPUStreamingRNN=CPUStreamingRNN, GPUStreamingRNN=CPUStreamingRNN, Folketinget=Folketinget,
                 TransferLearned=TransferLearned, EnglishLibrispeech=EnglishLibrispeech)
    f = table.get(model_name)
    return f() if f else None
