import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_outputs.npz"))


def rel_err(a, b):
    """max|a-b| / max|b|  -- the error metric of SURVEY A.1 for the 1e-4 / 2e-2 bars."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def logit_rel_err(p, ref, floor=1e-20):
    """Relative error in the logit domain: log-probabilities are the logits up to one additive constant
    per row, so max|log p - log ref| / (spread of the reference log-probabilities) is the "logits within
    1e-4 relative" bar of the north star.  Entries whose reference probability underflows are skipped."""
    p = np.asarray(p, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    m = (ref > floor) & (p > 0)
    assert m.mean() > 0.3
    lp, lr = np.log(p[m]), np.log(ref[m])
    return float(np.abs(lp - lr).max() / np.abs(lr).max())


BATCH_CASES = [
    ("gru_bi_c2", "TestModel", dict(rnn_type="gru", rnn_hidden_size=96, rnn_layers=3)),
    ("gru_bi_c3", "DanSpeechPrimary", dict(rnn_type="gru", rnn_hidden_size=80, rnn_layers=2)),
    ("gru_bi_c1", "TestModel", dict(rnn_type="gru", rnn_hidden_size=64, rnn_layers=2, conv_layers=1)),
    ("lstm_bi_c2", "TestModel", dict(rnn_type="lstm", rnn_hidden_size=72, rnn_layers=2, ih_scale=2.5)),
    ("rnn_bi_c2", "TestModel", dict(rnn_type="rnn", rnn_hidden_size=72, rnn_layers=2)),
    ("gru_uni_c2", "TestModel", dict(rnn_type="gru", rnn_hidden_size=96, rnn_layers=2, bidirectional=False,
                                     context=20)),
]
BATCH_LENS = [16000, 11111, 4000]


def batch_inputs():
    """The ragged 3-utterance batch used by gen_golden.py (oracle spectrograms, zero padded)."""
    import torch
    from oracle.spectrogram import SpectrogramOracle
    from danspeech_b200.utils import synthetic as syn
    p = SpectrogramOracle()
    auds = [syn.synthetic_audio(n, seed=7 + i) for i, n in enumerate(BATCH_LENS)]
    specs = [p.parse_audio(a) for a in auds]
    x = torch.zeros(3, 1, 161, specs[0].size(1))
    for i, s in enumerate(specs):
        x[i, 0, :, : s.size(1)] = s
    return auds, x, torch.IntTensor([s.size(1) for s in specs])


def case_config(name, kw):
    from danspeech_b200.utils import synthetic as syn
    cfg = dict(syn.MODEL_SHAPES[name])
    cfg.update(kw)
    return cfg
