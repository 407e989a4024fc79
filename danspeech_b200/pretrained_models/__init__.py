"""Model factories with the reference's names (danspeech/pretrained_models/__init__.py:1-30).

The reference factories download a ``.pth`` release artefact (e.g. danspeech_primary.py:22-24) which is
impossible offline, so each factory here builds the *named architecture*  and, unless a
``model_path`` to a reference package is given, fills it with seeded random weights
(``utils.synthetic.make_state_dict``).  ``CustomModel(path)`` loads a real reference package.
"""
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


def _factory(name):
    def make(cache_dir=None, model_path=None, seed=0):
        return build_model(name, model_path=model_path, seed=seed)
    make.__name__ = name
    make.__doc__ = "%s-shaped DeepSpeech model (random-init unless model_path is given)." % name
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
