"""GPU: many-stream serving pieces -- the energy VAD kernel against its oracle and the
multi-stream recogniser against S runs of the single-stream engine (itself checked against the oracle pipeline
in test_gpu_parity.py::test_streaming_transcribe_engine_matches_oracle)."""
import numpy as np
import pytest
import torch

from oracle import vad as ov
from danspeech_b200.utils import synthetic as syn

pytestmark = pytest.mark.gpu


def test_vad_kernel_matches_oracle_per_stream():
    from danspeech_b200.streaming import StreamVAD
    S, chunk = 6, 1024
    pcms = [ov.fixture_pcm(seed=s, n_buffers=160) for s in range(S)]
    pcms[1] = np.roll(pcms[1], 17 * chunk)                 # phrases at other places
    pcms[2] = (pcms[2].astype(np.int32) // 4).astype(np.int16)   # quieter: fewer buffers above the threshold
    pcms[3] = np.full_like(pcms[3], -32768)                # extreme amplitude, always loud
    thr = [1000, 1000, 500, 1000, 3000, 200]
    vad = StreamVAD(S, energy_threshold=thr, chunk=chunk)
    pause, phrase, _ = ov.buffer_counts(chunk)
    assert (vad.pause_buffer_count, vad.phrase_buffer_count, vad.non_speaking_buffer_count) == ov.buffer_counts(chunk)
    oracles = [ov.ListenStreamOracle(thr[s], pause, phrase) for s in range(S)]
    seen = set()
    for i in range(160):
        block = np.stack([p[i * chunk:(i + 1) * chunk] for p in pcms])
        ev, en = vad.push(torch.from_numpy(block))
        want = [o.push(block[s]) for s, o in enumerate(oracles)]
        assert en.tolist() == [w[0] for w in want], i
        assert ev.tolist() == [w[1] for w in want], i
        seen.update(ev.tolist())
    assert seen == {ov.SILENCE, ov.PHRASE_START, ov.SPEECH, ov.PHRASE_END, ov.PHRASE_DROPPED}
    # ragged chunk length and reset
    vad.reset()
    ev, en = vad.push(torch.from_numpy(np.stack([p[:999] for p in pcms])))
    assert en.tolist() == [ov.rms(p[:999]) for p in pcms]


def _chunks(a):
    # engine chunk schedule: first 8640 samples, then 6240 (Recognizer.py:602-611)
    out = [a[:8640]] + [a[8640 + 6240 * i: 8640 + 6240 * (i + 1)] for i in range(64)]
    return [c for c in out if len(c) > 0]


@pytest.mark.parametrize("secondary", [False, True])
def test_multistream_recognizer_equals_single_stream_engine(secondary):
    from danspeech_b200 import Recognizer
    from danspeech_b200.pretrained_models import build_model
    from danspeech_b200.streaming import MultiStreamRecognizer
    S = 4
    kw = dict(rnn_hidden_size=128, rnn_layers=3)
    n = 8640 + 6240 * 4 + 3000
    auds = [syn.synthetic_audio(n, seed=500 + s) for s in range(S)]
    auds[3] = np.zeros(n)                       # a silent stream: nothing heard -> "" at the end
    mk = lambda: build_model("CPUStreamingRNN", seed=5, **kw).set_precision("fp32")          # noqa: E731
    mk2 = lambda: build_model("TestModel", seed=8, rnn_hidden_size=96, rnn_layers=2).set_precision("fp32")   # noqa: E731
    solo = []
    for s in range(S):
        r = Recognizer()
        r.enable_real_time_streaming(mk(), secondary_model=mk2() if secondary else None, string_parts=True)
        cs = _chunks(auds[s])
        solo.append([r.streaming_transcribe(c, is_last=(i == len(cs) - 1), is_first=(i == 0)) for i, c in enumerate(cs)])
    eng = MultiStreamRecognizer(mk(), S, secondary_model=mk2() if secondary else None, string_parts=True)
    cs = [_chunks(a) for a in auds]
    for i in range(len(cs[0])):
        parts = np.stack([cs[s][i] for s in range(S)])
        got = eng.push(parts, is_first=(i == 0), is_last=(i == len(cs[0]) - 1))
        assert got == [solo[s][i] for s in range(S)], i
    assert any(len(t) > 1 for t in got)


def test_multistream_recognizer_bf16_many_streams_runs():
    from danspeech_b200.pretrained_models import build_model
    from danspeech_b200.streaming import MultiStreamRecognizer
    S = 130
    eng = MultiStreamRecognizer(build_model("CPUStreamingRNN", seed=5, rnn_hidden_size=160, rnn_layers=2).set_precision("bf16"), S)
    a = np.stack([syn.synthetic_audio(8640 + 6240 * 2, seed=600 + (s % 3)) for s in range(S)])
    outs = [eng.push(a[:, :8640], True, False), eng.push(a[:, 8640:8640 + 6240], False, False),
            eng.push(a[:, 8640 + 6240:], False, True)]
    assert all(len(o) == S for o in outs) and outs[0] == [""] * S
    assert outs[2][0] == outs[2][3] and isinstance(outs[2][129], str)   # same audio, same batch group -> same transcript
