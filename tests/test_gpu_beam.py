"""GPU: BeamCTCDecoder (device prefix beam search + device n-gram LM) against the CPU oracle.

Bars (BASELINE.json north star): beam top-1 transcripts identical on >= 99.5 % of utterances, beam
scores within 1e-3 (relative to the score magnitude).  Asserted here: every hypothesis of every beam, with its
character time steps, equal to the oracle's, scores to 1e-5."""
import numpy as np
import pytest
import torch

from danspeech_b200.utils import synthetic as syn
from oracle.beam import CTCBeamDecoderOracle

pytestmark = pytest.mark.gpu


def _spelled_probs(rng, vocab, B, T, noise=0.25):
    """Probabilities that roughly spell sequences of vocabulary words, with confusable alternatives."""
    C = len(syn.LABELS)
    probs = rng.dirichlet(np.ones(C) * noise, size=(B, T)).astype(np.float64) * 0.35
    lens = []
    for b in range(B):
        t = 0
        while t < T - 8:
            w = vocab[int(rng.integers(len(vocab)))] + " "
            for ch in w:
                k = syn.LABELS.index(ch)
                for _ in range(int(rng.integers(1, 3))):
                    if t < T:
                        probs[b, t, k] += 0.65
                        t += 1
                if rng.random() < 0.5 and t < T:
                    probs[b, t, 0] += 0.65
                    t += 1
        lens.append(int(rng.integers(T // 2, T + 1)))
    probs /= probs.sum(-1, keepdims=True)
    lens[0] = T
    return probs.astype(np.float32), sorted(lens, reverse=True)


def _compare(gpu, ref, probs, lens):
    """Top beam: labels, score and character time steps equal to the oracle's on EVERY utterance (the GPU arena keeps
    ctcdecode's PathTrie node identity and lifetime), and the whole beam: the same set of (labels, time steps) for all W
    hypotheses and the same scores.  (Measured on the B200: identical on every utterance of every case below.)"""
    out, scores, ts, out_len = [x.cpu().numpy() for x in gpu.decode_device(torch.from_numpy(probs).cuda(),
                                                                            torch.IntTensor(lens))]
    r_out, r_scores, r_ts, r_len = ref.decode(probs, lens)
    B, W = probs.shape[0], out.shape[1]
    whole = 0
    for b in range(B):
        n, rn = out_len[b, 0], r_len[b, 0]
        assert n == rn and np.array_equal(out[b, 0, :n], r_out[b, 0, :rn]), "top beam of utterance %d differs" % b
        assert abs(scores[b, 0] - r_scores[b, 0]) <= 1e-5 * max(1.0, abs(r_scores[b, 0])), b
        assert np.array_equal(ts[b, 0, :n], r_ts[b, 0, :rn]), "character time steps of utterance %d differ" % b
        mine = {(tuple(out[b, j, :out_len[b, j]].tolist()), tuple(ts[b, j, :out_len[b, j]].tolist())) for j in range(W)}
        theirs = {(tuple(r_out[b, j, :r_len[b, j]].tolist()), tuple(r_ts[b, j, :r_len[b, j]].tolist())) for j in range(W)}
        whole += mine == theirs
    print("whole beams (labels + time steps of all %d hypotheses) identical on %d / %d utterances" % (W, whole, B))
    assert whole == B
    k = min(8, W)
    assert np.allclose(np.sort(scores[:, :k], axis=1), np.sort(r_scores[:, :k], axis=1), rtol=1e-5, atol=1e-5)
    return whole


@pytest.mark.parametrize("beam", [1, 16, 64, 100])
def test_beam_without_lm_matches_oracle(beam):
    from danspeech_b200.deepspeech.decoder import BeamCTCDecoder
    rng = np.random.default_rng(10 + beam)
    vocab = syn.synthetic_vocab(200, 0)
    probs, lens = _spelled_probs(rng, vocab, B=16, T=120)
    gpu = BeamCTCDecoder(syn.LABELS, None, 0, 0, 40, 1.0, beam, 4, 0)
    ref = CTCBeamDecoderOracle(syn.LABELS, None, 0, 0, 40, 1.0, beam, 4, 0)
    _compare(gpu, ref, probs, lens)


def test_beam_word_lm_matches_oracle(tmp_path):
    from danspeech_b200.deepspeech.decoder import BeamCTCDecoder
    arpa = syn.write_synthetic_arpa(str(tmp_path / "w.arpa"), n_words=2000, seed=0)
    vocab = syn.synthetic_vocab(2000, 0)
    rng = np.random.default_rng(21)
    probs, lens = _spelled_probs(rng, vocab, B=32, T=160)
    gpu = BeamCTCDecoder(syn.LABELS, arpa, 1.3, 0.2, 40, 1.0, 64, 6, 0)
    ref = CTCBeamDecoderOracle(syn.LABELS, arpa, 1.3, 0.2, 40, 1.0, 64, 6, 0)
    _compare(gpu, ref, probs, lens)
    strings, _ = gpu.decode(torch.from_numpy(probs).cuda(), torch.IntTensor(lens))
    r_strings, _ = ref.decode_strings(probs, lens)
    assert [s[0] for s in strings] == [s[0] for s in r_strings]
    assert len(strings[0]) == 64


def test_beam_char_lm_matches_oracle(tmp_path):
    from danspeech_b200.deepspeech.decoder import BeamCTCDecoder
    arpa = syn.write_synthetic_arpa(str(tmp_path / "c.arpa"), char_based=True, seed=2, n_bigrams=400, n_trigrams=800)
    rng = np.random.default_rng(22)
    probs, lens = _spelled_probs(rng, syn.synthetic_vocab(200, 0), B=8, T=80)
    gpu = BeamCTCDecoder(syn.LABELS, arpa, 0.8, 0.1, 40, 1.0, 32, 4, 0)
    ref = CTCBeamDecoderOracle(syn.LABELS, arpa, 0.8, 0.1, 40, 1.0, 32, 4, 0)
    _compare(gpu, ref, probs, lens)


def test_beam_top_n_pruning_and_edge_lengths(tmp_path):
    from danspeech_b200.deepspeech.decoder import BeamCTCDecoder
    rng = np.random.default_rng(23)
    probs, _ = _spelled_probs(rng, syn.synthetic_vocab(200, 0), B=4, T=40)
    lens = [40, 17, 1, 0]
    gpu = BeamCTCDecoder(syn.LABELS, None, 0, 0, 10, 1.0, 16, 4, 0)      # cutoff_top_n < C
    ref = CTCBeamDecoderOracle(syn.LABELS, None, 0, 0, 10, 1.0, 16, 4, 0)
    _compare(gpu, ref, probs, lens)


def test_recognizer_with_lm_end_to_end(tmp_path, golden):
    """BASELINE config 3 shape in miniature: model + BeamCTCDecoder(beam 64, synthetic 3-gram ARPA)."""
    from danspeech_b200 import Recognizer
    from danspeech_b200.pretrained_models import build_model
    arpa = syn.write_synthetic_arpa(str(tmp_path / "w.arpa"), n_words=2000, seed=0)
    model = build_model("TestModel", seed=0).set_precision("fp32")
    r = Recognizer(model=model, lm=arpa, alpha=1.3, beta=0.2, beam_width=64)
    audio = golden["wav_u0013002"].astype(np.float64)
    beams = r.recognize(audio, show_all=True)
    assert isinstance(beams, list) and len(beams) == 64
    ref = CTCBeamDecoderOracle(syn.LABELS, arpa, 1.3, 0.2, 40, 1.0, 64, 6, 0)
    r_strings, _ = ref.decode_strings(golden["cfg1_probs"], golden["cfg1_sizes"])
    assert r.recognize(audio) == r_strings[0][0]


@pytest.mark.parametrize("char_based", [False, True])
def test_beam_klm_equals_arpa(tmp_path, char_based):
    """the reference hands the decoder a KenLM binary (language_models/dsl_3gram.py:16-20).  The same
    model as ARPA text and as a probing .klm (written by oracle/klm_writer.py) must decode identically: every beam,
    every score, bit for bit -- both files load into the same device table."""
    from danspeech_b200.deepspeech.decoder import BeamCTCDecoder
    from oracle import klm_writer as kw
    arpa = syn.write_synthetic_arpa(str(tmp_path / "m.arpa"), n_words=2000, seed=0, char_based=char_based)
    klm = str(tmp_path / "m.klm")
    kw.write_klm(arpa, klm)
    rng = np.random.default_rng(31)
    probs, lens = _spelled_probs(rng, syn.synthetic_vocab(2000, 0), B=16, T=140)
    p, ln = torch.from_numpy(probs).cuda(), torch.IntTensor(lens)
    a = BeamCTCDecoder(syn.LABELS, arpa, 1.3, 0.2, 40, 1.0, 64, 6, 0)
    k = BeamCTCDecoder(syn.LABELS, klm, 1.3, 0.2, 40, 1.0, 64, 6, 0)
    L = a._handle and __import__("danspeech_b200")._native.lib()
    assert L.dsb_beam_lm_order(k._handle) == L.dsb_beam_lm_order(a._handle) == 3
    assert L.dsb_beam_lm_num_ngrams(k._handle) == L.dsb_beam_lm_num_ngrams(a._handle)
    assert L.dsb_beam_lm_is_char_based(k._handle) == L.dsb_beam_lm_is_char_based(a._handle) == int(char_based)
    for x, y in zip(a.decode_device(p, ln), k.decode_device(p, ln)):
        assert torch.equal(x, y)
    # and the ARPA path still agrees with the CPU oracle (which keys its LM by word tuples, not by hashes)
    ref = CTCBeamDecoderOracle(syn.LABELS, arpa, 1.3, 0.2, 40, 1.0, 64, 6, 0)
    _compare(k, ref, probs, lens)


def test_recognizer_accepts_a_klm_language_model(tmp_path, golden):
    """Recognizer(model, lm="x.klm") -- the reference's own call shape (Recognizer.py:39-80 with a language_models path)."""
    from danspeech_b200 import Recognizer
    from danspeech_b200.pretrained_models import build_model
    from oracle import klm_writer as kw
    arpa = syn.write_synthetic_arpa(str(tmp_path / "dsl_3gram.arpa"), n_words=2000, seed=0)
    kw.write_klm(arpa, str(tmp_path / "dsl_3gram.klm"))
    from danspeech_b200.language_models import DSL3gram
    lm = DSL3gram(cache_dir=str(tmp_path))
    assert lm.endswith(".klm")
    model = build_model("TestModel", seed=0).set_precision("fp32")
    audio = golden["wav_u0013002"].astype(np.float64)
    got = Recognizer(model=model, lm=lm, alpha=1.3, beta=0.2, beam_width=64).recognize(audio, show_all=True)
    want = Recognizer(model=model, lm=arpa, alpha=1.3, beta=0.2, beam_width=64).recognize(audio, show_all=True)
    assert got == want and len(got) == 64
