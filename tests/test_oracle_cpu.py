"""CPU: the oracle restatements reproduce the golden outputs of the unmodified reference."""
import os

import numpy as np
import pytest
import torch

from conftest import BATCH_CASES, batch_inputs, case_config, rel_err
from oracle import greedy as og
from oracle import model as om
from oracle import refharness
from oracle import spectrogram as osp
from danspeech_b200.utils import synthetic as syn


def test_spectrogram_oracle_matches_reference_parser(golden):
    p = osp.SpectrogramOracle()
    for name in ("u0013002", "u0042018"):
        s = p.parse_audio(golden["wav_" + name].astype(np.float64)).numpy()
        assert s.shape == golden["spect_" + name].shape
        assert np.array_equal(s, golden["spect_" + name])
    for i, n in enumerate((161, 1000, 16000, 40001)):
        s = p.parse_audio(syn.synthetic_audio(n, seed=100 + i)).numpy()
        assert s.shape == (161, 1 + n // 160)
        assert np.array_equal(s, golden["spect_syn%d" % n])


def test_spectrogram_shape_and_normalisation(golden):
    s = golden["spect_u0013002"]
    assert s.shape == (161, 419)            # SURVEY A.1 probe: 66 944 samples -> 419 frames
    assert abs(float(s.mean())) < 1e-5
    assert abs(float(s.std(ddof=1)) - 1.0) < 1e-5


def test_streaming_spectrogram_oracle(golden):
    a = golden["wav_u0013002"].astype(np.float64)
    chunks = [a[:8640]] + [a[8640 + 6240 * i: 8640 + 6240 * (i + 1)] for i in range(12)]
    chunks = [c for c in chunks if len(c) > 0]
    assert len(chunks) == int(golden["stream_n_chunks"])
    sp = osp.StreamingSpectrogramOracle()
    for i, c in enumerate(chunks):
        s = sp.parse_audio(c, is_last=(i == len(chunks) - 1))
        ref = golden["stream_spect_%d" % i]
        if len(s) == 0:
            assert ref.size == 0
        else:
            assert s.shape[0] == 161
            assert rel_err(s.numpy(), ref) < 1e-6
    assert golden["stream_spect_0"].shape == (161, 53)    # 8640 samples -> 53 frames
    assert golden["stream_spect_1"].shape == (161, 39)    # 6240 (+160 carry) -> 39 frames


def test_model_oracle_config1(golden):
    cfg = case_config("TestModel", {})
    sd = syn.make_state_dict(seed=0, **cfg)
    sp = torch.from_numpy(golden["spect_u0013002"])
    probs, sizes = om.forward(sd, sp.view(1, 1, 161, -1), torch.IntTensor([sp.size(1)]), cfg["conv_layers"],
                              cfg["rnn_layers"])
    assert probs.shape == (1, 210, 33)      # T' = 210
    assert rel_err(probs.numpy(), golden["cfg1_probs"]) < 1e-6
    assert np.array_equal(sizes.numpy(), golden["cfg1_sizes"])
    text = og.greedy_decode(probs.numpy(), sizes.numpy())[0][0][0]
    assert text == str(golden["cfg1_text"])


@pytest.mark.parametrize("tag,name,kw", BATCH_CASES, ids=[c[0] for c in BATCH_CASES])
def test_model_oracle_ragged_batch(golden, tag, name, kw):
    kw = dict(kw)
    rt = kw.pop("rnn_type")
    cfg = case_config(name, kw)
    sd = syn.make_state_dict(rnn_type=rt, seed=3, **cfg)
    _, x, xl = batch_inputs()
    probs, sizes = om.forward(sd, x, xl, cfg["conv_layers"], cfg["rnn_layers"], bidirectional=cfg["bidirectional"],
                              rnn_type=rt, context=cfg.get("context", 20))
    assert rel_err(probs.numpy(), golden["batch_%s_probs" % tag]) < 1e-6
    assert np.array_equal(sizes.numpy(), golden["batch_%s_sizes" % tag])
    strings, offs = og.greedy_decode(probs.numpy(), sizes.numpy())
    assert [s[0] for s in strings] == [str(s) for s in golden["batch_%s_text" % tag]]
    assert np.array_equal(np.concatenate([o[0] for o in offs]), golden["batch_%s_offs" % tag])


def test_streaming_model_oracle(golden):
    cfg = case_config("CPUStreamingRNN", dict(rnn_hidden_size=128, rnn_layers=3))
    sd = syn.make_state_dict(seed=5, **cfg)
    a = golden["smodel_audio"].astype(np.float64)
    chunks = [a[:8640]] + [a[8640 + 6240 * i: 8640 + 6240 * (i + 1)] for i in range(5)]
    sp = osp.StreamingSpectrogramOracle()
    m = om.StreamingOracle(sd, cfg["rnn_layers"], context=cfg["context"])
    frames = []
    for i, c in enumerate(chunks):
        last = i == len(chunks) - 1
        s = sp.parse_audio(c, is_last=last)
        o = m.forward(s.view(1, 1, 161, -1), i == 0, last)
        ref = golden["smodel_probs_%d" % i]
        if o is None:
            assert ref.size == 0
            frames.append(0)
        else:
            assert o.shape == ref.shape
            assert rel_err(o.numpy(), ref) < 1e-5
            frames.append(o.shape[1])
    assert frames[0] == 0 and frames[1] == 50 and frames[2] == 35     # SURVEY A.5 frame bookkeeping


def test_greedy_oracle_edge_cases():
    labels = "_ab "
    probs = np.zeros((2, 6, 4), dtype=np.float32)
    seq = [[1, 1, 0, 1, 3, 3], [0, 0, 0, 0, 0, 0]]
    for b in range(2):
        for t, s in enumerate(seq[b]):
            probs[b, t, s] = 1.0
    strings, offs = og.greedy_decode(probs, [6, 6], labels)
    assert strings == [["aa "], [""]]
    assert offs[0][0].tolist() == [0, 3, 4] and offs[1][0].tolist() == []
    strings, _ = og.greedy_decode(probs, [2, 0], labels)
    assert strings == [["a"], [""]]


@pytest.mark.skipif(not refharness.reference_available(), reason="reference tree only exists in the build container")
def test_repin_against_live_reference(golden):
    """When /root/reference is mounted, re-run the reference itself and compare with the oracle."""
    refharness.import_reference()
    from danspeech.deepspeech.model import DeepSpeech
    cfg = case_config("TestModel", dict(rnn_hidden_size=64, rnn_layers=2))
    sd = syn.make_state_dict(seed=11, **cfg)
    m = DeepSpeech(model_name="t", labels=syn.LABELS, rnn_hidden_size=64, rnn_layers=2, conv_layers=2).eval()
    m.load_state_dict(sd)
    _, x, xl = batch_inputs()
    with torch.no_grad():
        ref, rs = m(x, xl)
    probs, sizes = om.forward(sd, x, xl, 2, 2)
    assert rel_err(probs.numpy(), ref.numpy()) < 1e-6 and np.array_equal(sizes.numpy(), rs.numpy())
