"""CPU: the oracle restatements reproduce the golden outputs of the unmodified reference."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import BATCH_CASES, ROOT, batch_inputs, case_config, rel_err
from oracle import greedy as og
from oracle import model as om
from oracle import refharness
from oracle import spectrogram as osp
from oracle import vad as ov
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


# ------------------------------------------------------------------ energy VAD
def _vad_rows(events, non_speaking):
    """(is_last, buffers in the yield, source position) rows the reference generator produces for these events."""
    rows, pre = [], 0
    for i, ev in enumerate(events):
        if ev in (ov.SILENCE, ov.PHRASE_START):
            pre = min(pre + 1, non_speaking)
            if ev == ov.PHRASE_START:
                rows.append((0, pre, i + 1))
                pre = 0
        elif ev == ov.SPEECH:
            rows.append((0, 1, i + 1))
        elif ev == ov.PHRASE_END:
            rows.append((1, 1, i + 1))
            pre = 0
        else:
            pre = 0
    return rows


def _vad_fixture_events():
    pcm = ov.fixture_pcm()
    pause, phrase, non_speaking = ov.buffer_counts()
    o = ov.ListenStreamOracle(1000, pause, phrase)
    res = [o.push(pcm[i * 1024:(i + 1) * 1024]) for i in range(len(pcm) // 1024)]
    return pcm, [e for _, e in res], [en for en, _ in res], non_speaking


def test_vad_oracle_matches_reference_generator_golden():
    """oracle/vad.py against the yields of the unmodified Recognizer.listen_stream (tests/golden/vad_reference.npz)."""
    ref = np.load(os.path.join(os.path.dirname(__file__), "golden", "vad_reference.npz"))["yields"]
    pcm, events, _, non_speaking = _vad_fixture_events()
    n_buffers = len(pcm) // 1024
    want = [tuple(r) for r in ref.tolist() if r[2] <= n_buffers]   # the tail is the source running dry
    assert _vad_rows(events, non_speaking) == want
    assert events.count(ov.PHRASE_END) == 2 and events.count(ov.PHRASE_DROPPED) == 1


def test_vad_oracle_rms_is_audioop_rms():
    audioop = pytest.importorskip("audioop")
    rng = np.random.default_rng(3)
    for n, scale in ((1024, 3000), (1024, 50), (7, 30000), (4096, 12000)):
        x = np.clip(rng.normal(0, scale, n), -32768, 32767).astype(np.int16)
        assert ov.rms(x) == audioop.rms(x.tobytes(), 2)
    assert ov.rms(np.full(16, -32768, np.int16)) == audioop.rms(np.full(16, -32768, np.int16).tobytes(), 2) == 32768


@pytest.mark.skipif(not refharness.reference_available(), reason="reference tree only exists in the build container")
def test_vad_repin_against_live_reference():
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import gen_vad_golden
    rows = gen_vad_golden.reference_yields(ov.fixture_pcm())
    ref = np.load(os.path.join(os.path.dirname(__file__), "golden", "vad_reference.npz"))["yields"]
    assert np.array_equal(rows, ref)


# ------------------------------------------------------------------ more real audio (tests/golden/reference_wavs.npz)
WAV_NAMES = ("u0042008", "u0042012", "u0042017", "u0042019")


def test_oracle_on_more_example_wavs():
    """Oracle spectrogram + model + greedy against the unmodified reference on four further example WAVs."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_wavs.npz"))
    cfg = case_config("TestModel", {})
    sd = syn.make_state_dict(seed=0, **cfg)
    p = osp.SpectrogramOracle()
    for name in WAV_NAMES[:2]:   # two of the four keep the CPU suite short; the GPU suite checks all of them
        a = g["wav_" + name].astype(np.float64)
        spect = p.parse_audio(a)
        assert tuple(spect.shape) == tuple(g["spect_shape_" + name])
        assert rel_err(spect[[3, 80]].numpy(), g["spect_rows_" + name]) < 1e-5
        probs, sizes = om.forward(sd, spect.view(1, 1, 161, -1), torch.IntTensor([spect.size(1)]), cfg["conv_layers"],
                                  cfg["rnn_layers"])
        assert rel_err(probs[0, [0, -1]].numpy(), g["probs_ends_" + name]) < 1e-5
        text = og.greedy_decode(probs.numpy(), sizes.tolist())[0][0][0]
        assert text == str(g["text_" + name])


def test_stft_restatement_against_independent_implementations():
    """librosa is absent (the STFT is "parity unpinned"), so the restatement of its documented definition -- centred
    frames with reflect padding, scipy's symmetric Hamming window, one-sided FFT -- is cross-checked against two
    independent implementations of the same transform: torch.stft and scipy.signal.stft (undoing scipy's window-sum
    scaling).  Bounds the restatement error; the fixtures above pin everything the reference does around it."""
    import scipy.signal
    import scipy.signal.windows
    for i, n in enumerate((161, 4000, 66944)):
        a = syn.synthetic_audio(n, seed=300 + i)
        D = osp.stft(a)
        assert D.dtype == np.complex64 and D.shape == (161, 1 + n // 160)
        w = scipy.signal.windows.hamming(320, sym=True)
        assert np.array_equal(w, osp.hamming_sym(320)) or np.abs(w - osp.hamming_sym(320)).max() < 1e-15
        Dt = torch.stft(torch.from_numpy(a), n_fft=320, hop_length=160, win_length=320, window=torch.from_numpy(w),
                        center=True, pad_mode="reflect", return_complex=True).numpy()
        assert Dt.shape == D.shape
        assert np.abs(Dt - D).max() / np.abs(Dt).max() < 1e-6          # complex64 storage of a float64 transform
        ap = np.pad(a, 160, mode="reflect")
        _, _, Ds = scipy.signal.stft(ap, window=w, nperseg=320, noverlap=160, nfft=320, boundary=None, padded=False,
                                     return_onesided=True, detrend=False)
        Ds = Ds * w.sum()
        assert Ds.shape == D.shape
        assert np.abs(Ds - D).max() / np.abs(Ds).max() < 1e-6


@pytest.mark.skipif(not refharness.reference_available(), reason="reference tree only exists in the build container")
def test_oracle_equals_live_reference_recognize_on_all_13_example_wavs():
    """end-to-end ``Recognizer.recognize`` of the UNMODIFIED reference on every WAV of its
    example_files against the oracle pipeline (spectrogram -> model -> greedy), and this package's loader against the
    reference's ``load_audio`` (stereo s16 -> clip(L+R))."""
    import glob
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import gen_golden
    ref = refharness.import_reference()
    from danspeech import Recognizer
    from danspeech.audio.resources import load_audio as ref_load_audio
    from danspeech_b200.audio.resources import load_audio
    wavs = sorted(glob.glob(os.path.join(refharness.REFERENCE_ROOT, "example_files", "*.wav")))
    assert len(wavs) == 13
    cfg = case_config("TestModel", {})
    sd = syn.make_state_dict(seed=0, **cfg)
    p = osp.SpectrogramOracle()
    with torch.no_grad():
        r = Recognizer(model=gen_golden.ref_model(ref, "TestModel", seed=0))
        for path in wavs:
            a = ref_load_audio(path)
            assert np.array_equal(load_audio(path), a)
            spect = p.parse_audio(a)
            probs, sizes = om.forward(sd, spect.view(1, 1, 161, -1), torch.IntTensor([spect.size(1)]), cfg["conv_layers"],
                                      cfg["rnn_layers"])
            text = og.greedy_decode(probs.numpy(), sizes.tolist())[0][0][0]
            assert text == r.recognize(a), os.path.basename(path)
            assert len(text) > 0


@pytest.mark.parametrize("window", ["hann", "blackman", "bartlett"])
def test_window_oracles_match_reference_parsers(window):
    """audio_conf["window"] (parsers.py:9-10): oracle windows against goldens written by the unmodified reference
    parsers (tests/golden/gen_window_golden.py), offline and streaming."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "reference_windows.npz"))
    a = g["audio"].astype(np.float64)
    conf = dict(window=window)
    s = osp.SpectrogramOracle(conf).parse_audio(a).numpy()
    assert s.shape == g["spect_" + window].shape
    assert np.abs(s - g["spect_" + window]).max() < 1e-5
    sp = osp.StreamingSpectrogramOracle(conf)
    for i, c in enumerate([a[:8640], a[8640:8640 + 6240], a[8640 + 6240:]]):
        o = sp.parse_audio(c, is_last=(i == 2)).numpy()
        assert np.abs(o - g["stream_%s_%d" % (window, i)]).max() < 1e-5
    # the windows themselves against scipy's definitions
    import scipy.signal.windows as sw
    assert np.allclose(osp._WINDOWS[window](320), getattr(sw, window)(320), atol=1e-15)
