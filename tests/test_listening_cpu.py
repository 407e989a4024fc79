"""CPU: the listening API (danspeech_b200/listening.py, the single-stream caller of the hot path) against the
UNMODIFIED reference generators (danspeech/Recognizer.py:133-336, :717-818) and the committed golden yields."""
import os
import time

import numpy as np
import pytest

from oracle import refharness
from oracle import vad as ov
from danspeech_b200.audio.resources import ArraySource, AudioData
from danspeech_b200.errors.recognizer_errors import NoDataInBuffer, WaitTimeoutError, WrongUsageOfListen
from danspeech_b200.listening import PhraseListener, pcm_rms

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
HAVE_REF = refharness.reference_available()


def make_source(pcm, chunk=1024, base=object, short_tail=False):
    """The fake source of tests/golden/gen_vad_golden.py: int16 samples, b"" once fewer than a buffer remain
    (or, with ``short_tail``, a final short buffer like a file)."""
    class Fake(base):
        def __init__(self):
            self.pos, self.chunk, self.sampling_rate, self.sampling_width = 0, chunk, 16000, 2
            src = self

            class Stream:
                def read(self, n):
                    b = pcm[src.pos:src.pos + n]
                    src.pos += n
                    if len(b) == n or (short_tail and len(b)):
                        return b.tobytes()
                    return b""
            self.stream = Stream()
    return Fake()


def listener(**params):
    p = PhraseListener()
    p._init_listening()
    p.stream = True
    for k, v in params.items():
        setattr(p, k, v)
    return p


def reference_listener(**params):
    refharness.import_reference()
    from danspeech import Recognizer
    r = Recognizer.__new__(Recognizer)      # the listening loops only read these attributes (Recognizer.py:42-66)
    r.energy_threshold, r.pause_threshold, r.phrase_threshold, r.non_speaking_duration = 1000, 0.8, 0.3, 0.35
    r.dynamic_energy_threshold, r.dynamic_energy_adjustment_damping, r.dynamic_energy_ratio = True, 0.15, 1.5
    r.mininum_required_speaking_seconds = 0.7
    r.stream = True
    for k, v in params.items():
        setattr(r, k, v)
    return r


def reference_source_base():
    refharness.import_reference()
    from danspeech.audio.resources import SpeechSource
    return SpeechSource


def drain_phrases(lst, src, n_samples, **kw):
    """One generator per phrase, as the capture thread uses it; every yield as (is_last, bytes, source position)."""
    rows = []
    while src.pos < n_samples:
        for is_last, data in lst.listen_stream(src, **kw):
            blob = b"".join(data) if isinstance(data, list) else bytes(data)
            rows.append((bool(is_last), blob, src.pos))
            if is_last:
                break
    return rows


def test_pcm_rms_is_audioop_rms():
    audioop = pytest.importorskip("audioop")
    rng = np.random.default_rng(5)
    for width, dt, hi in ((2, "<i2", 32767), (4, "<i4", 2 ** 31 - 1), (1, "i1", 127)):
        for n, scale in ((1024, 0.1), (1024, 0.9), (7, 0.5), (1, 1.0)):
            x = np.clip(rng.normal(0, scale * hi / 3, n), -hi - 1, hi).astype(dt)
            assert pcm_rms(x.tobytes(), width) == audioop.rms(x.tobytes(), width)
    assert pcm_rms(b"", 2) == 0
    with pytest.raises(ValueError):
        pcm_rms(b"\0\0\0", 3)


def test_listen_stream_reproduces_the_golden_yields_of_the_reference():
    """tests/golden/vad_reference.npz holds (is_last, buffers in the yield, source position) of every yield of the
    unmodified ``Recognizer.listen_stream`` over oracle.vad.fixture_pcm()."""
    ref = np.load(os.path.join(GOLDEN, "vad_reference.npz"))["yields"]
    pcm = ov.fixture_pcm()
    rows = drain_phrases(listener(), make_source(pcm), len(pcm))
    got = [(int(last), len(blob) // 2048, pos // 1024) for last, blob, pos in rows]
    assert got == [tuple(r) for r in ref.tolist()]
    assert sum(1 for r in got if r[0]) >= 2


def test_listen_stream_ends_with_wrong_usage_and_honours_a_stopped_stream():
    pcm = ov.fixture_pcm()
    lst = listener()
    gen = lst.listen_stream(make_source(pcm))
    for is_last, _ in gen:
        if is_last:
            break
    with pytest.raises(WrongUsageOfListen):
        next(gen)
    lst.stream = False                                   # a stopped recognizer: one (True, []) and out
    assert list(zip(range(1), lst.listen_stream(make_source(pcm)))) == [(0, (True, []))]
    with pytest.raises(WaitTimeoutError):
        next(listener().listen_stream(make_source(np.zeros(160 * 1024, np.int16)), timeout=1.0))
    with pytest.raises(AssertionError):
        next(listener().listen_stream(object()))


CASES = [dict(seed=0, chunk=1024), dict(seed=1, chunk=512), dict(seed=2, chunk=1024, energy_threshold=300),
         dict(seed=3, chunk=2048, pause_threshold=0.5, phrase_threshold=0.6),
         dict(seed=4, chunk=1024, phrase_time_limit=0.9), dict(seed=5, chunk=1024, short_tail=True, cut=40 * 1024 + 300)]


@pytest.mark.skipif(not HAVE_REF, reason="reference tree only exists in the build container")
@pytest.mark.parametrize("case", CASES)
def test_listen_stream_and_listen_equal_the_live_reference(case):
    case = dict(case)
    chunk, seed = case.pop("chunk"), case.pop("seed")
    limit, short_tail, cut = case.pop("phrase_time_limit", None), case.pop("short_tail", False), case.pop("cut", None)
    pcm = ov.fixture_pcm(seed=seed, n_buffers=160 * 1024 // chunk, chunk=chunk)[:cut]
    base = reference_source_base()
    mine = drain_phrases(listener(**case), make_source(pcm, chunk, short_tail=short_tail), len(pcm), phrase_time_limit=limit)
    ref = drain_phrases(reference_listener(**case), make_source(pcm, chunk, base, short_tail), len(pcm), phrase_time_limit=limit)
    assert mine == ref and len(mine) > 5

    # listen(): same phrases byte for byte, same drift of the dynamic energy threshold
    a, b = listener(**case), reference_listener(**case)
    sa, sb = make_source(pcm, chunk, short_tail=short_tail), make_source(pcm, chunk, base, short_tail)
    for _ in range(3):
        x = a.listen(sa, phrase_time_limit=limit)
        y = b.listen(sb, phrase_time_limit=limit)
        assert x.frame_data == y.frame_data and (x.sample_rate, x.sample_width) == (y.sample_rate, y.sample_width)
        assert sa.pos == sb.pos and a.energy_threshold == b.energy_threshold
        assert np.array_equal(x.get_array_data(), y.get_array_data())
    with pytest.raises(WaitTimeoutError):
        listener().listen(make_source(np.zeros(64 * 1024, np.int16)), timeout=0.5)


@pytest.mark.skipif(not HAVE_REF, reason="reference tree only exists in the build container")
def test_threshold_adjustment_and_array_conversion_equal_the_live_reference():
    base = reference_source_base()
    pcm = ov.fixture_pcm(seed=7)
    for method, duration in (("adjust_for_speech", 4), ("adjust_for_speech", 0.5), ("adjust_for_ambient_noise", 2),
                             ("adjust_for_ambient_noise", 1.3)):
        a, b = listener(), reference_listener()
        sa, sb = make_source(pcm[20 * 1024:]), make_source(pcm[20 * 1024:], base=base)
        getattr(a, method)(sa, duration=duration)
        getattr(b, method)(sb, duration=duration)
        assert a.energy_threshold == b.energy_threshold and sa.pos == sb.pos
    quiet = np.zeros(8 * 1024, np.int16) + 3                 # mean energy below 80: no subtraction
    a, b = listener(), reference_listener()
    a.adjust_for_speech(make_source(quiet), duration=0.3)
    b.adjust_for_speech(make_source(quiet, base=base), duration=0.3)
    assert a.energy_threshold == b.energy_threshold == 3
    a.update_stream_parameters(energy_threshold=250, pause_threshold=0, non_speaing_duration=0.2)
    b.update_stream_parameters(energy_threshold=250, pause_threshold=0, non_speaing_duration=0.2)
    assert (a.energy_threshold, a.pause_threshold, a.phrase_threshold, a.non_speaking_duration) == \
           (b.energy_threshold, b.pause_threshold, b.phrase_threshold, b.non_speaking_duration) == (250, 0.8, 0.3, 0.2)

    from danspeech.audio.resources import AudioData as RefAudioData
    rng = np.random.default_rng(1)
    for width, dt in ((2, "<i2"), (4, "<i4"), (1, "u1")):
        raw = rng.integers(0, 255, 64 * width, dtype=np.uint8).tobytes()
        assert np.array_equal(AudioData(raw, 16000, width).get_array_data(), RefAudioData(raw, 16000, width).get_array_data())
        np.frombuffer(raw, dtype=dt)
    mine, ref = AudioData(pcm[:5000].tobytes(), 16000, 2), RefAudioData(pcm[:5000].tobytes(), 16000, 2)
    assert mine.get_wav_data() == ref.get_wav_data()
    for lo, hi in ((None, None), (10, None), (None, 100), (12.5, 250), (0, 0)):
        assert mine.get_segment(lo, hi).frame_data == ref.get_segment(lo, hi).frame_data
    with pytest.raises(AssertionError):
        mine.get_segment(50, 10)
    frames = [pcm[:1024].tobytes(), pcm[1024:1500].tobytes()]
    src = make_source(pcm)
    assert np.array_equal(PhraseListener.get_audio_data(frames, src), reference_listener().get_audio_data(frames, src))
    assert PhraseListener.get_audio_data([], src).shape == (0,)


class FakeEngine:
    """Stands in for DanSpeechRecognizer: records what the generators hand to the model."""

    def __init__(self, context=20):
        self.model = type("M", (), {"context": context})()
        self.calls = []

    def streaming_transcribe(self, samples, is_last, is_first):
        self.calls.append((len(samples), bool(is_last), bool(is_first), np.array(samples)))
        return "" if is_first else "<%d>" % len(self.calls)


def _next_with_deadline(gen, seconds=20.0):
    t0 = time.time()
    out = next(gen)
    assert time.time() - t0 < seconds
    return out


@pytest.mark.timeout(120)
def test_streaming_generator_recognizes_one_clip_per_phrase():
    pcm = ov.fixture_pcm()
    lst = listener(stream=False)
    heard = []
    lst.recognize = lambda clip: heard.append(np.array(clip)) or "phrase %d" % len(heard)
    lst.enable_streaming()
    gen = lst.streaming(ArraySource(pcm, chunk_size=1024))
    assert _next_with_deadline(gen) == "phrase 1" and _next_with_deadline(gen) == "phrase 2"
    assert list(gen) == []                                 # the array has been consumed: the generator ends by itself
    lst.disable_streaming()
    assert lst.stream is False
    # the clips are the phrases of the fixture: pre-roll + speech + the pause that ended them, straight from the PCM
    rows = drain_phrases(listener(), make_source(pcm, short_tail=True), len(pcm))
    phrases, cur = [], b""
    for last, blob, _ in rows:
        cur += blob
        if last:
            phrases.append(np.frombuffer(cur, "<i2").astype(float))
            cur = b""
    assert len(heard) == 2 and all(np.array_equal(h, p) for h, p in zip(heard, phrases))
    assert all(len(h) > 0.7 * 16000 for h in heard)


@pytest.mark.timeout(120)
def test_listen_in_background_queue_and_end_of_a_finite_source():
    pcm = ov.fixture_pcm()
    lst = listener()
    stopper, get_data = lst.listen_in_background(ArraySource(pcm))
    items, deadline = [], time.time() + 20
    while time.time() < deadline and sum(1 for last, _ in items if last) < 3:
        try:
            items.append(get_data())
        except NoDataInBuffer:
            time.sleep(0.01)
    stopper(wait_for_stop=True)                            # returns: the capture thread ended with the source
    total = sum(len(a) for _, a in items)
    assert sum(1 for last, _ in items if last) >= 2 and 0 < total <= len(pcm)
    assert all(a.dtype == np.float64 for _, a in items)
    with pytest.raises(NoDataInBuffer):
        while True:
            get_data()


@pytest.mark.timeout(180)
def test_real_time_streaming_feeds_the_model_in_lookahead_sized_passes():
    pcm = ov.fixture_pcm()
    lst = listener()
    lst.danspeech_recognizer = FakeEngine(context=20)
    gen = lst.real_time_streaming(ArraySource(pcm, chunk_size=1024, realtime=1.0))   # paced like a microphone: 10 s
    outs = list(gen)                                       # ends by itself once the recording has been consumed
    lst.stream_thread_stopper(wait_for_stop=True)
    calls = lst.danspeech_recognizer.calls
    need_later = 160 * 2 + 160 * 37                         # (context - 1) * 2 = 38 frames of look-ahead
    need_first = need_later + 160 * 15
    firsts = [c for c in calls if c[2]]
    lasts = [c for c in calls if c[1]]
    assert len(firsts) >= 2 and len(lasts) >= 2
    assert all(c[0] >= need_first and not c[1] for c in firsts)               # first pass: enough context, never last
    assert all(c[0] >= need_later for c in calls if not c[1] and not c[2])    # later passes: the look-ahead
    assert calls[0][2] and [c for c in calls if c[1] or c[2]][1][1]           # a phrase is first ... last, in order
    assert all(text and text.startswith("<") for _, text in outs)             # the first pass ("") is never yielded
    assert [last for last, _ in outs].count(True) == 2
    # what reached the model is the captured audio, in order
    rows = drain_phrases(listener(), make_source(pcm, short_tail=True), len(pcm))
    captured = np.frombuffer(b"".join(blob for _, blob, _ in rows), "<i2").astype(float)
    fed = np.concatenate([c[3] for c in calls])
    assert np.array_equal(fed, captured[:len(fed)])


def _scripted_run(make, script, context=20, empty=NoDataInBuffer):
    """Drives a ``real_time_streaming`` generator with a scripted capture queue (no threads, no sleeping): every
    ``get_data`` call consumes one script entry -- None = queue empty, (is_last, n) = n samples.  When the script is
    used up the stream flag drops, which ends the generator after the consumer has handled what it holds."""
    obj = make()
    obj.stream = True
    eng = FakeEngine(context=context)
    obj.danspeech_recognizer = eng
    todo = list(script)
    counter = [0]

    def get_data():
        if not todo:
            obj.stream = False
            raise empty
        item = todo.pop(0)
        if item is None:
            raise empty
        is_last, n = item
        counter[0] += 1
        return is_last, np.full(n, float(counter[0]))

    obj.listen_in_background = lambda source: (lambda wait_for_stop=True: None, get_data)
    source = type("S", (), {"sampling_rate": 16000, "chunk": 1024, "sampling_width": 2})()
    yields = [(bool(last), text) for last, text in obj.real_time_streaming(source)]
    calls = [(n, last, first, float(a.sum())) for n, last, first, a in eng.calls]
    return yields, calls


@pytest.mark.skipif(not HAVE_REF, reason="reference tree only exists in the build container")
def test_real_time_streaming_equals_the_live_reference_on_scripted_queues(monkeypatch):
    """Differential test of the consumer state machine (Recognizer.py:560-715): same scripted queue, same fake
    engine -> the same ``streaming_transcribe`` calls (sizes, flags, content) and the same yields, including the
    reference's quirk that a phrase ending before its first pass stays queued in front of the next one."""
    refharness.import_reference()
    import importlib
    ref_module = importlib.import_module("danspeech.Recognizer")     # (the package attribute of that name is the class)
    from danspeech.errors.recognizer_errors import NoDataInBuffer as RefNoData
    mine_module = importlib.import_module("danspeech_b200.listening")
    monkeypatch.setattr(ref_module.time, "sleep", lambda s: None)     # the same `time` module object: both are patched
    assert mine_module.time.sleep(0) is None

    make_ref = reference_listener

    rng = np.random.default_rng(42)
    n_checked = 0
    for trial in range(300):
        script = []
        for _ in range(int(rng.integers(3, 40))):
            u = rng.random()
            if u < 0.35:
                script.append(None)
            else:
                script.append((bool(rng.random() < 0.15), int(rng.integers(0, 9)) * 1024))
        script.append((bool(rng.random() < 0.5), 1024))                # end on data: the reference's poll loop only
        want = _scripted_run(make_ref, script, empty=RefNoData)        # leaves on data-then-empty or a phrase end
        got = _scripted_run(listener, script)
        assert got == want, (trial, script)
        n_checked += 1
    assert n_checked == 300


@pytest.mark.skipif(not HAVE_REF, reason="reference tree only exists in the build container")
def test_streaming_generator_equals_the_live_reference_on_scripted_queues(monkeypatch):
    """``streaming`` (Recognizer.py:433-497): phrases are collected up to their last part and recognised when longer
    than ``mininum_required_speaking_seconds``; same scripted queue -> same clips in the same order."""
    import importlib
    refharness.import_reference()
    ref_module = importlib.import_module("danspeech.Recognizer")
    from danspeech.errors.recognizer_errors import NoDataInBuffer as RefNoData
    monkeypatch.setattr(ref_module.time, "sleep", lambda s: None)

    def run(make, script, empty):
        obj = make()
        obj.stream = True
        obj.recognize = lambda clip: (len(clip), float(np.sum(clip)))
        todo, counter = list(script), [0]

        def get_data():
            item = todo.pop(0)
            if not todo:
                obj.stream = False            # the script ends on a phrase end: the generator finishes after it
            if item is None:
                raise empty
            counter[0] += 1
            return item[0], np.full(item[1], float(counter[0]))

        obj.listen_in_background = lambda source: (lambda wait_for_stop=True: None, get_data)
        source = type("S", (), {"sampling_rate": 16000, "chunk": 1024, "sampling_width": 2})()
        return list(obj.streaming(source))

    rng = np.random.default_rng(7)
    for trial in range(200):
        script = []
        for _ in range(int(rng.integers(1, 30))):
            script.append(None if rng.random() < 0.3 else (bool(rng.random() < 0.25), int(rng.integers(0, 8)) * 1024))
        script.append((True, int(rng.integers(0, 16)) * 1024))
        assert run(listener, script, NoDataInBuffer) == run(reference_listener, script, RefNoData), (trial, script)


def test_segment_audio_and_recognize_long_batch_the_phrases_of_a_recording():
    pcm = ov.fixture_pcm()
    lst = listener(dynamic_energy_threshold=False)
    spans = lst.segment_audio(pcm)
    # the same phrases `listen` returns from a live source, located in the recording
    ref, src = listener(dynamic_energy_threshold=False), make_source(pcm, short_tail=True)
    want = []
    while src.pos < len(pcm):
        want.append(np.frombuffer(ref.listen(src).frame_data, "<i2"))
    want = [w for w in want if len(w) and np.sqrt(np.mean(w.astype(float) ** 2)) > 300]
    assert len(spans) == 2 and all(0 <= lo < hi <= len(pcm) for lo, hi in spans) and spans[0][1] <= spans[1][0]
    assert all(np.array_equal(pcm[lo:hi], w) for (lo, hi), w in zip(spans, want[:2]))
    # speech of the fixture: buffers 20-45 (with a short internal pause) and 100-130; pre-roll and tail are kept
    assert spans[0][0] <= 20 * 1024 and spans[0][1] >= 45 * 1024 and spans[1][0] <= 100 * 1024 and spans[1][1] >= 130 * 1024

    batches_seen = []

    def recognize_batches(batches, show_all=False):
        batches_seen.append([[len(c) for c in b] for b in batches])
        return [["clip of %d" % len(c) for c in b] for b in batches]

    lst.recognize_batches = recognize_batches
    out = lst.recognize_long(pcm, max_batch=1)
    assert [(round(a * 16000), round(b * 16000)) for a, b, _ in out] == spans
    assert [t for _, _, t in out] == ["clip of %d" % (hi - lo) for lo, hi in spans]
    assert batches_seen == [[[max(hi - lo for lo, hi in spans)], [min(hi - lo for lo, hi in spans)]]]   # longest first
    assert lst.recognize_long(np.zeros(40 * 1024, np.int16)) == []
    # a recording that stops a few samples into the pause after a phrase: spans still index the samples exactly
    cut = pcm[:58 * 1024 + 300]
    (lo, hi), = listener(dynamic_energy_threshold=False).segment_audio(cut)
    src = make_source(cut, short_tail=True)
    assert np.array_equal(cut[lo:hi], np.frombuffer(listener(dynamic_energy_threshold=False).listen(src).frame_data, "<i2"))
