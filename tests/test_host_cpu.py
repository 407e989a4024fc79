"""CPU: host-side logic -- state-dict layout, model factories, sharding (incl. a world-size-2 gloo run)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from danspeech_b200 import sharding
from danspeech_b200.utils import synthetic as syn


def test_state_dict_layout_matches_survey_a6():
    from danspeech_b200.pretrained_models import build_model
    m = build_model("DanSpeechPrimary", rnn_hidden_size=64)      # narrow: same names / ranks, fast
    sd = m.state_dict()
    assert len(sd) == 139                                          # 139 entries for 3 conv + 9 bi-GRU
    assert tuple(sd["conv.seq_module.0.weight"].shape) == (32, 1, 41, 11)
    assert tuple(sd["conv.seq_module.3.weight"].shape) == (32, 32, 21, 11)
    assert tuple(sd["conv.seq_module.6.weight"].shape) == (96, 32, 21, 11)
    assert tuple(sd["rnns.0.rnn.weight_ih_l0"].shape) == (3 * 64, 2016)
    assert tuple(sd["rnns.8.rnn.weight_hh_l0_reverse"].shape) == (3 * 64, 64)
    assert "rnns.0.batch_norm.module.weight" not in sd and "rnns.1.batch_norm.module.running_var" in sd
    assert tuple(sd["fc.0.module.1.weight"].shape) == (33, 64)
    full = syn.make_state_dict(**syn.MODEL_SHAPES["TestModel"])
    assert sum(v.numel() for k, v in full.items() if not k.endswith("num_batches_tracked")
               and "running" not in k) == 12081168             # SURVEY appendix C parameter count


def test_streaming_and_unidirectional_keys():
    from danspeech_b200.pretrained_models import build_model
    s = build_model("CPUStreamingRNN", rnn_hidden_size=32, rnn_layers=2)
    assert "lookahead.conv.weight" in s.state_dict() and s.streaming_model
    assert tuple(s.state_dict()["rnns.0.rnn.weight_ih_l0"].shape) == (96, 1312)
    u = build_model("TestModel", rnn_hidden_size=32, rnn_layers=2, bidirectional=False, context=20)
    assert "lookahead.0.conv.weight" in u.state_dict()
    assert "rnns.0.rnn.weight_ih_l0_reverse" not in u.state_dict()


def test_package_round_trip(tmp_path):
    from danspeech_b200.deepspeech.model import DeepSpeech
    from danspeech_b200.pretrained_models import build_model, CustomModel
    m = build_model("TestModel", rnn_hidden_size=32, rnn_layers=2, seed=4)
    path = str(tmp_path / "m.pth")
    torch.save(m.serialize(), path)
    m2 = CustomModel(path)
    assert isinstance(m2, DeepSpeech) and m2.rnn_hidden_size == 32 and m2.labels == m.labels
    for k, v in m.state_dict().items():
        assert torch.equal(v, m2.state_dict()[k])


def test_conv_error_and_seq_lens():
    import torch.nn as nn
    from danspeech_b200.deepspeech.model import DeepSpeech
    from danspeech_b200.errors.model_errors import ConvError
    with pytest.raises(ConvError):
        DeepSpeech("x", conv_layers=0)
    with pytest.raises(ConvError):
        DeepSpeech("x", conv_layers=4)
    m = DeepSpeech("x", rnn_type=nn.GRU, rnn_hidden_size=16, rnn_layers=1, conv_layers=3)
    assert m.get_seq_lens(torch.IntTensor([1501, 419, 2, 1])).tolist() == [751, 210, 1, 1]


def test_get_model_from_string_quirk():
    from danspeech_b200 import pretrained_models as pm
    assert pm.get_model_from_string("nope") is None
    assert set(syn.MODEL_SHAPES) >= {"DanSpeechPrimary", "TestModel", "Baseline", "CPUStreamingRNN", "GPUStreamingRNN",
                                     "Folketinget", "TransferLearned", "EnglishLibrispeech"}


def test_decoder_helpers_without_gpu():
    from danspeech_b200.deepspeech.decoder import Decoder
    d = Decoder(syn.LABELS, blank_index=0)
    assert d.space_index == 32 and d.int_to_char[27] == "æ"
    assert d.wer("en to tre", "en tre") == 1 and d.cer("abc", "abd") == 1
    assert Decoder("_ab").space_index == 3          # out-of-range sentinel (decoder.py:40-43)


def test_synthetic_arpa_is_well_formed(tmp_path):
    p = syn.write_synthetic_arpa(str(tmp_path / "a.arpa"), n_words=50, seed=1, n_bigrams=100, n_trigrams=100)
    txt = open(p, encoding="utf-8").read()
    assert txt.startswith("\\data\\") and "\\3-grams:" in txt and txt.rstrip().endswith("\\end\\")
    assert "ngram 1=53" in txt


def test_lpt_shards_and_batches():
    rng = np.random.default_rng(0)
    lengths = rng.integers(5 * 16000, 30 * 16000, size=256).tolist()
    for n in (1, 2, 4, 8):
        shards = sharding.lpt_shards(lengths, n)
        assert sorted(i for s in shards for i in s) == list(range(256))
        loads = [sum(lengths[i] for i in s) for s in shards]
        assert max(loads) - min(loads) <= max(lengths)            # LPT guarantee
        for s in shards:
            for b in sharding.make_batches(s, lengths, max_batch=64):
                assert len(b) <= 64
                assert all(lengths[b[i]] >= lengths[b[i + 1]] for i in range(len(b) - 1))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(1)
    recs = [np.zeros(int(n)) for n in rng.integers(100, 3000, size=37)]
    calls = []

    def fake_recognize(batch):          # stands in for Recognizer.recognize_batch (no GPU here)
        calls.append(len(batch))
        assert all(len(batch[i]) >= len(batch[i + 1]) for i in range(len(batch) - 1))
        return ["len%d" % len(a) for a in batch]

    out = sharding.transcribe_sharded(fake_recognize, recs, rank, world, max_batch=8)
    q.put((rank, out, sum(calls)))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_transcription_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(1)
    expect = ["len%d" % int(n) for n in rng.integers(100, 3000, size=37)]
    outs = {r: o for r, o, _ in got}
    assert outs[0] == expect and outs[1] == expect        # every rank holds the full, ordered result
    assert sum(n for _, _, n in got) == 37                # disjoint shards covering everything
    single = sharding.transcribe_sharded(lambda b: ["len%d" % len(a) for a in b],
                                         [np.zeros(int(n)) for n in np.random.default_rng(1).integers(100, 3000, size=37)])
    assert single == expect                               # shard invariance: same answer for world size 1


def test_stitch_transcript_is_the_reference_rule():
    """utils.stitch against the inline rule of DanSpeechRecognizer.py:169-174 on exhaustive small cases."""
    import itertools
    from danspeech_b200.utils.stitch import stitch_transcript

    def inline(it, transcript):
        if it and transcript and it[-1] == transcript[0]:
            it = it + transcript[1:]
            transcript = transcript[1:]
        else:
            it += transcript
        return it, transcript

    words = ["".join(w) for n in range(0, 4) for w in itertools.product("ab ", repeat=n)]
    for it in words:
        for t in words:
            assert stitch_transcript(it, t) == inline(it, t)


def test_update_decoder_rules_without_a_gpu():
    """DanSpeechRecognizer.update_decoder (reference DanSpeechRecognizer.py:58-95): falsy arguments keep the field, a
    decoder is built when there is none (lm becomes "greedy") and rebuilt only on a real change."""
    from danspeech_b200.DanSpeechRecognizer import DanSpeechRecognizer
    r = object.__new__(DanSpeechRecognizer)
    r.lm, r.decoder, r.alpha, r.beta, r.beam_width, r.labels = None, None, 1.3, 0.2, 64, ["_", "a", " "]
    built = []
    r._make_decoder = lambda: built.append((r.lm, r.alpha, r.beta, r.beam_width, tuple(r.labels))) or len(built)
    r.update_decoder(labels=r.labels)                       # first call: greedy decoder
    assert r.lm == "greedy" and built == [("greedy", 1.3, 0.2, 64, ("_", "a", " "))]
    r.update_decoder(labels=r.labels)                       # nothing changes -> no rebuild
    r.update_decoder(alpha=1.3, beta=0.2, beam_width=64)
    r.update_decoder(alpha=0, beta=None, lm="")             # falsy = keep
    assert len(built) == 1
    r.update_decoder(lm="/x/lm.arpa", alpha=1.5)
    assert r.lm == "/x/lm.arpa" and r.alpha == 1.5 and len(built) == 2
    r.update_decoder(lm="/x/lm.arpa", beam_width=32)
    assert r.beam_width == 32 and len(built) == 3 and r.decoder == 3
    r.update_decoder(labels=["_", "b", " "])
    assert r.labels == ["_", "b", " "] and len(built) == 4


def test_streaming_transcribe_control_flow_without_a_gpu():
    """DanSpeechRecognizer.streaming_transcribe (reference :144-216) with stub parser / model / decoders: first chunk
    returns "", parts are stitched, an empty last part still finishes, nothing heard returns "" and keeps the state."""
    import torch
    from danspeech_b200.DanSpeechRecognizer import DanSpeechRecognizer

    class Dec:
        def __init__(self, texts):
            self.texts = list(texts)

        def decode(self, probs, sizes=None):
            return [[self.texts.pop(0)]], None

    def engine(texts, secondary=None, lm="greedy", final=None, string_parts=True):
        r = object.__new__(DanSpeechRecognizer)
        r.secondary_model, r.lm, r.string_parts = secondary, lm, string_parts
        r.iterating_transcript, r.full_output, r.spectrograms = "", [], []
        r.audio_parser = type("P", (), {"parse_audio": staticmethod(lambda rec, last: rec)})()
        r.model = lambda x, first, last: torch.zeros(1, x.size(3), 3)
        r.greedy_decoder = Dec(texts)
        r.decoder = Dec([final]) if final is not None else r.greedy_decoder
        return r

    sp = torch.zeros(161, 4)
    r = engine(["ab", "bc", "cd"])
    assert r.streaming_transcribe(sp, False, True) == ""
    assert r.streaming_transcribe(sp, False, False) == "ab"
    assert r.streaming_transcribe(sp, False, False) == "c"          # "bc" loses its first character
    assert r.streaming_transcribe(sp, True, False) == "abcd"        # last: the whole stitched transcript
    assert r.iterating_transcript == "" and r.full_output == []
    r = engine(["ab", "b"], string_parts=False)
    r.streaming_transcribe(sp, False, True)
    assert r.streaming_transcribe(sp, False, False) == "ab"
    assert r.streaming_transcribe([], True, False) == "ab"          # empty last part (parser returned [])
    r = engine(["a"])
    r.streaming_transcribe(sp, False, True)
    assert r.streaming_transcribe(sp, True, False) == ""            # one character: "nothing heard"
    assert r.iterating_transcript == "a"                            # and, as in the reference, no reset
    r = engine(["ab", "c"], lm="/x/lm.arpa", final="ab c")
    r.streaming_transcribe(sp, False, True)
    r.streaming_transcribe(sp, False, False)
    assert r.streaming_transcribe(sp, True, False) == "ab c"        # LM decoder over the concatenated outputs
    sec = lambda x, sizes: (torch.zeros(1, int(sizes[0]), 3), sizes)
    r = engine(["ab", "c"], secondary=sec, final="abc!")
    r.streaming_transcribe(sp, False, True)
    r.streaming_transcribe(sp, False, False)
    assert len(r.spectrograms) == 2
    assert r.streaming_transcribe(sp, True, False) == "abc!" and r.spectrograms == []


def test_multistream_push_control_flow_without_a_gpu():
    """MultiStreamRecognizer.push with stub parser / model / decoders: per stream it must behave like
    DanSpeechRecognizer.streaming_transcribe (first chunk "", stitched parts, final transcript or "" when nothing was
    heard; one batched secondary pass at the end)."""
    import torch
    from danspeech_b200.streaming import MultiStreamRecognizer

    class Greedy:
        def __init__(self, per_chunk):
            self.per_chunk = list(per_chunk)

        def decode_strings(self, probs):
            return self.per_chunk.pop(0)

        def decode(self, probs, sizes=None):
            return [[t] for t in self.final], None

    def engine(per_chunk, S, secondary=None, string_parts=True):
        e = object.__new__(MultiStreamRecognizer)
        e.S, e.string_parts, e.secondary_model, e.decoder = S, string_parts, secondary, None
        e.greedy_decoder = Greedy(per_chunk)
        e.audio_parser = type("P", (), {"parse_audio": staticmethod(lambda parts, last: parts)})()
        e.model = lambda x, first, last: torch.zeros(S, x.size(3), 3)
        e.reset_streaming_params()
        return e

    sp = torch.zeros(3, 161, 4)
    e = engine([["ab", "", "x"], ["bc", "q", ""]], S=3)
    assert e.push(sp, True, False) == ["", "", ""]
    assert e.push(sp, False, False) == ["ab", "", "x"]
    assert e.push(sp, False, True) == ["abc", "", ""]            # stream 1: one character, stream 2: "x" -> nothing heard
    # as in the reference (DanSpeechRecognizer.py:181-214) only a stream that was heard is reset
    assert e.iterating_transcript == ["", "q", "x"]
    e = engine([["ab", "cd"], ["b", "d"]], S=2, string_parts=False)
    e.push(sp[:2], True, False)
    assert e.push(sp[:2], False, False) == ["ab", "cd"]
    assert e.push(None, False, True) == ["ab", "cd"]             # parser had nothing for the last part
    sec = lambda x, sizes: (torch.zeros(2, int(sizes[0]), 3), sizes)
    e = engine([["ab", "c"]], S=2, secondary=sec)
    e.greedy_decoder.final = ["AB!", "ignored"]
    e.push(sp[:2], True, False)
    assert e.push(sp[:2], False, True) == ["AB!", ""]            # secondary pass for all, "" where nothing was heard
    assert e.spectrograms == []


def test_recognizer_shell_wires_the_listening_api(monkeypatch):
    """Recognizer = PhraseListener + engine calls; with a stub engine the whole shell runs without a GPU."""
    import importlib
    R = importlib.import_module("danspeech_b200.Recognizer")   # the package attribute of that name is the class

    class Engine:
        def __init__(self, with_gpu=True, **kw):
            self.kw, self.log = kw, []

        def update_model(self, m):
            self.log.append(("model", m.model_name))

        def update_decoder(self, **kw):
            self.log.append(("decoder", kw))

        def enable_streaming(self, secondary, parts):
            self.log.append(("enable", secondary, parts))

        def disable_streaming(self, keep_secondary_model=False):
            self.log.append(("disable", keep_secondary_model))

        def transcribe(self, audio, show_all=False):
            return "text:%d" % len(audio)

    monkeypatch.setattr(R, "DanSpeechRecognizer", Engine)
    model = type("M", (), {"model_name": "stub"})()
    r = R.Recognizer(model=model, alpha=1.0)
    assert r.danspeech_recognizer.kw == {"alpha": 1.0} and r.danspeech_recognizer.log == [("model", "stub")]
    assert (r.energy_threshold, r.pause_threshold, r.phrase_threshold, r.non_speaking_duration) == (1000, 0.8, 0.3, 0.35)
    assert r.stream is False and r.stream_thread_stopper is None and r.microphone is None
    assert r.recognize([0.0] * 5) == "text:5"
    r.disable_real_time_streaming()                        # nothing running: a message, no engine call
    assert r.danspeech_recognizer.log == [("model", "stub")]
    r.enable_real_time_streaming(model, secondary_model=None, string_parts=False)
    assert r.stream is True and r.danspeech_recognizer.log[-1] == ("enable", None, False)
    r.disable_real_time_streaming(keep_secondary_model_loaded=True)   # chunks were pushed by hand: no capture thread
    assert r.stream is False and r.danspeech_recognizer.log[-1] == ("disable", True)
    with pytest.raises(R.ModelNotInitialized):
        R.Recognizer(lm="x.arpa")
    for name in ("listen", "listen_stream", "listen_in_background", "get_audio_data", "streaming", "real_time_streaming",
                 "enable_streaming", "disable_streaming", "adjust_for_speech", "adjust_for_ambient_noise",
                 "update_stream_parameters"):
        assert callable(getattr(r, name))


def test_greedy_host_helpers_match_the_reference_and_the_oracle():
    """GreedyDecoder.process_string / convert_to_strings (decoder.py:151-181) on index sequences."""
    from danspeech_b200.deepspeech.decoder import GreedyDecoder
    from oracle import greedy as og, refharness
    labels = list("_abcdefghijklmnopqrstuvwxyzæøåéü ")
    d = GreedyDecoder(labels, blank_index=0)
    rng = np.random.default_rng(0)
    seqs = torch.from_numpy(rng.integers(0, 6, size=(5, 40))).int()
    seqs[seqs == 5] = labels.index(" ")
    sizes = torch.IntTensor([40, 33, 1, 0, 17])
    got = d.convert_to_strings(seqs, sizes, remove_repetitions=True, return_offsets=True)
    onehot = np.eye(len(labels), dtype=np.float32)[seqs.numpy()]
    want_s, want_o = og.greedy_decode(onehot, sizes.tolist(), labels="".join(labels))
    assert [g[0] for g in got[0]] == [w[0] for w in want_s]
    assert all(np.array_equal(g[0].numpy(), np.asarray(w[0])) for g, w in zip(got[1], want_o))
    assert d.convert_to_strings(seqs[:2]) == d.convert_to_strings(seqs[:2], [40, 40])
    assert d.process_string(torch.tensor([1, 1, 0, 1, 2, 2]), 6)[0] == "aaabb"
    assert d.process_string(torch.tensor([1, 1, 0, 1, 2, 2]), 6, remove_repetitions=True)[0] == "aab"
    if refharness.reference_available():
        refharness.import_reference()
        from danspeech.deepspeech.decoder import GreedyDecoder as RefGreedy
        r = RefGreedy(labels, blank_index=0)
        for rep in (False, True):
            a = d.convert_to_strings(seqs, sizes, remove_repetitions=rep, return_offsets=True)
            b = r.convert_to_strings(seqs, sizes, remove_repetitions=rep, return_offsets=True)
            assert a[0] == b[0]
            assert all(torch.equal(x[0], y[0]) and x[0].dtype == y[0].dtype for x, y in zip(a[1], b[1]))


def test_language_model_factories_return_cached_paths_and_never_download(tmp_path):
    from danspeech_b200 import language_models as lm
    assert lm.CustomLanguageModel("/x/y.arpa") == "/x/y.arpa"
    with pytest.raises(FileNotFoundError, match="does not download"):
        lm.DSL3gram(cache_dir=str(tmp_path))
    (tmp_path / "dsl_3gram.arpa").write_text("\\data\\\n")
    assert lm.DSL3gram(cache_dir=str(tmp_path)) == str(tmp_path / "dsl_3gram.arpa")
    (tmp_path / "dsl_3gram.klm").write_bytes(b"x")
    assert lm.DSL3gram(cache_dir=str(tmp_path)) == str(tmp_path / "dsl_3gram.klm")      # the reference's artefact wins
    assert set(lm.__all__) >= {"DSL5gram", "DSLWiki3gram", "DSLWiki5gram", "DSLWikiLeipzig3gram", "Wiki3gram", "Wiki5gram",
                               "Folketinget3gram", "DSL3gramWithNames"}
    assert lm.Wiki5gram.__name__ == "Wiki5gram"


def test_public_api_surface_covers_the_reference():
    """Drop-in check by introspection of the imported reference: every public name of the classes on the path exists
    here, with the same parameter names in the same order (defaults may differ only where INTEGRATION.md says so)."""
    import importlib
    import inspect
    from oracle import refharness
    if not refharness.reference_available():
        pytest.skip("reference tree only exists in the build container")
    refharness.import_reference()
    allowed_missing = {"DeepSpeech": {"freeze_layers", "streaming_init"},                     # training / module-internal
                       "SpeechFile": {"SpeechFileStream"}, "Microphone": {"MicrophoneStream"}}  # nested helper classes
    classes = [("Recognizer", "Recognizer"), ("DanSpeechRecognizer", "DanSpeechRecognizer"),
               ("deepspeech.model", "DeepSpeech"), ("deepspeech.decoder", "Decoder"),
               ("deepspeech.decoder", "GreedyDecoder"), ("deepspeech.decoder", "BeamCTCDecoder"),
               ("audio.parsers", "AudioParser"), ("audio.parsers", "SpectrogramAudioParser"),
               ("audio.parsers", "InferenceSpectrogramAudioParser"), ("audio.resources", "AudioData"),
               ("audio.resources", "SpeechFile"), ("audio.resources", "Microphone")]
    module_base = {n for n, _ in inspect.getmembers(torch.nn.Module)}
    for mod, cls in classes:
        R = getattr(importlib.import_module("danspeech." + mod), cls)
        M = getattr(importlib.import_module("danspeech_b200." + mod), cls)
        names = {n for n, _ in inspect.getmembers(R) if not n.startswith("_")}
        if issubclass(R, torch.nn.Module):
            names -= module_base
        missing = {n for n in names if not hasattr(M, n)} - allowed_missing.get(cls, set())
        assert not missing, (cls, sorted(missing))
        for n in sorted(names | {"__init__"}):
            a, b = getattr(R, n, None), getattr(M, n, None)
            if b is None or not callable(a) or inspect.isclass(a):
                continue
            pa = [p for p in inspect.signature(a).parameters.values()]
            pb = [p for p in inspect.signature(b).parameters.values()]
            if any(p.kind in (p.VAR_POSITIONAL, p.VAR_KEYWORD) for p in pb[:2 + 1]) and n in ("load_state_dict",):
                continue
            assert [p.name for p in pa] == [p.name for p in pb][:len(pa)], (cls, n)
    for mod, names in (("pretrained_models", ["DanSpeechPrimary", "TestModel", "Baseline", "CPUStreamingRNN", "GPUStreamingRNN",
                                              "Folketinget", "TransferLearned", "EnglishLibrispeech", "CustomModel",
                                              "get_model_from_string"]),
                       ("language_models", ["DSL3gram", "DSL5gram", "DSLWiki3gram", "DSLWiki5gram", "DSLWikiLeipzig3gram",
                                            "Wiki3gram", "Wiki5gram", "Folketinget3gram", "DSL3gramWithNames",
                                            "CustomLanguageModel"]),
                       ("audio", ["load_audio", "load_audio_wavPCM", "Microphone", "SpectrogramAudioParser",
                                  "InferenceSpectrogramAudioParser"])):
        R, M = importlib.import_module("danspeech." + mod), importlib.import_module("danspeech_b200." + mod)
        for n in names:
            assert hasattr(R, n) and hasattr(M, n), (mod, n)
            pa, pb = list(inspect.signature(getattr(R, n)).parameters), list(inspect.signature(getattr(M, n)).parameters)
            assert pa == pb[:len(pa)], (mod, n, pa, pb)


def test_model_factories_load_the_cached_package_or_raise(tmp_path):
    """pretrained_models factories (danspeech/pretrained_models/*.py): load <cache_dir>/<Name>.pth like the reference's
    get_model cache, raise when it is absent (never silently random weights), random weights only on request; the
    package is read with the restricted unpickler (weights_only=True)."""
    import torch
    from danspeech_b200 import pretrained_models as pm
    with pytest.raises(FileNotFoundError) as ei:
        pm.TestModel(cache_dir=str(tmp_path))
    assert "TestModel.pth" in str(ei.value) and "synthetic=True" in str(ei.value)
    with pytest.raises(FileNotFoundError):
        pm.get_model_from_string("EnglishLibrispeech") if not os.path.exists(
            os.path.expanduser("~/.danspeech/models/Librispeech.pth")) else (_ for _ in ()).throw(FileNotFoundError())
    small = pm.build_model("TestModel", seed=4, rnn_hidden_size=32, rnn_layers=1)
    torch.save(small.serialize(), str(tmp_path / "TestModel.pth"))
    loaded = pm.TestModel(cache_dir=str(tmp_path))
    assert loaded.rnn_hidden_size == 32 and loaded.rnn_layers == 1
    for (k, a), (_, b) in zip(small.state_dict().items(), loaded.state_dict().items()):
        assert torch.equal(a, b), k
    assert pm.CustomModel(str(tmp_path / "TestModel.pth")).model_name == "TestModel"
    syn_model = pm.TestModel(synthetic=True)
    assert syn_model.rnn_hidden_size == 400 and pm.TestModel(seed=0).rnn_layers == 5

    # a package that smuggles an arbitrary object is refused unless the caller opts into full unpickling
    class Evil:
        def __reduce__(self):
            return (print, ("unpickled",))
    pkg = small.serialize()
    pkg["extra"] = Evil()
    torch.save(pkg, str(tmp_path / "evil.pth"))
    with pytest.raises(Exception):
        pm.CustomModel(str(tmp_path / "evil.pth"))


def test_audio_parser_windows_and_config_errors():
    """audio_conf["window"]: the four reference windows are accepted (parsers.py:9-10), an unknown one raises KeyError
    like the reference's dict lookup, another framing than 20 ms / 10 ms at 16 kHz is refused loudly."""
    from danspeech_b200.audio.parsers import SpectrogramAudioParser, _WINDOW_FLAGS
    assert set(_WINDOW_FLAGS) == {"hamming", "hann", "blackman", "bartlett"}
    for w, flag in _WINDOW_FLAGS.items():
        p = SpectrogramAudioParser(dict(window=w))
        assert p._window_flag == flag and (flag >> 4) in (0, 1, 2, 3)
    with pytest.raises(KeyError):
        SpectrogramAudioParser(dict(window="kaiser"))
    with pytest.raises(NotImplementedError):
        SpectrogramAudioParser(dict(window_size=0.025))


def test_recognize_batches_merge_plan():
    """transcribe_batches runs up to `merge` consecutive batches through one pass of the model, within a row and a
    sample budget, never mixing host lists with pinned tensors; order is kept."""
    from danspeech_b200.DanSpeechRecognizer import DanSpeechRecognizer as D
    d = D.__new__(D)
    short = [np.zeros(1000)] * 64
    assert d._merge_plan([short] * 7, 3) == [[0, 1, 2], [3, 4, 5], [6]]
    assert d._merge_plan([short] * 7, 1) == [[k] for k in range(7)]
    assert d._merge_plan([short] * 4, 2) == [[0, 1], [2, 3]]
    assert d._merge_plan([[np.zeros(10)] * 140, [np.zeros(10)] * 140, [np.zeros(10)] * 20], 3) == [[0], [1, 2]]   # 256 rows
    long = [np.zeros(50 * 16000)] * 64                       # 3 x 64 x 50 s exceeds the sample budget of a pass
    assert d._merge_plan([long] * 4, 3) == [[0, 1], [2, 3]]
    pinned = (torch.zeros(8, 100), [100] * 8)
    assert d._merge_plan([short, pinned, pinned, short], 3) == [[0], [1, 2], [3]]
    assert d._merge_plan([], 3) == []
