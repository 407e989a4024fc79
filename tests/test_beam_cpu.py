"""CPU: known-answer tests of the beam-search oracle (C++ restatement of ctcdecode, parity unpinned)."""
import itertools
import math

import numpy as np
import pytest

from danspeech_b200.utils import synthetic as syn
from oracle.beam import CTCBeamDecoderOracle

LN10 = math.log(10.0)

ARPA = """\\data\\
ngram 1=6
ngram 2=4
ngram 3=2

\\1-grams:
-1.0\t<unk>\t0
-99\t<s>\t-0.5
-1.2\t</s>\t0
-0.8\tab\t-0.3
-0.9\tba\t-0.4
-1.5\ta\t-0.2

\\2-grams:
-0.4\t<s> ab\t-0.25
-0.6\tab ba\t-0.15
-0.7\tba </s>
-0.5\tba ab\t-0.1

\\3-grams:
-0.2\t<s> ab ba
-0.3\tab ba ab

\\end\\
"""


@pytest.fixture(scope="module")
def toy(tmp_path_factory):
    p = tmp_path_factory.mktemp("lm") / "toy.arpa"
    p.write_text(ARPA, encoding="utf-8")
    return str(p)


def test_lm_backoff_known_answers(toy):
    d = CTCBeamDecoderOracle("_ab ", toy, 1.0, 0.0, 40, 1.0, 8, 1, 0)
    assert d.lm_order == 3 and d.is_char_based == 0
    f = d.lm_cond_log_prob
    ln = lambda x: x / 0.4342944819  # noqa: E731  (the conversion constant upstream uses)
    assert f(["<s>", "ab", "ba"]) == pytest.approx(ln(-0.2), rel=1e-6)               # trigram hit
    assert f(["<s>", "<s>", "ab"]) == pytest.approx(ln(-0.4), rel=1e-6)              # (<s> <s>) absent -> bigram
    assert f(["ab", "ba", "ab"]) == pytest.approx(ln(-0.3), rel=1e-6)
    assert f(["ba", "ab", "ba"]) == pytest.approx(ln(-0.1 - 0.6), rel=1e-6)          # backoff(ba ab) + p(ba|ab)
    assert f(["ab", "ab", "a"]) == pytest.approx(ln(-0.3 - 1.5), rel=1e-6)           # (ab ab) absent; bo(ab) + p(a)
    assert f(["ab", "zz"]) == -1000.0                                                 # OOV_SCORE
    assert f(["zz", "ab"]) == -1000.0


def _exact_ctc(probs, labels, blank=0):
    """Brute force: P(string) = sum over all alignments that collapse to it."""
    T, C = probs.shape
    out = {}
    for path in itertools.product(range(C), repeat=T):
        pr = 1.0
        for t, c in enumerate(path):
            pr *= probs[t, c]
        s, prev = [], None
        for c in path:
            if c != blank and c != prev:
                s.append(labels[c])
            prev = c
        key = "".join(s)
        out[key] = out.get(key, 0.0) + pr
    return out


def test_beam_without_lm_is_exact_for_a_wide_beam():
    rng = np.random.default_rng(3)
    labels = "_ab"
    probs = rng.dirichlet(np.ones(3), size=5).astype(np.float32)
    exact = _exact_ctc(probs.astype(np.float64), labels)
    d = CTCBeamDecoderOracle(labels, None, 0, 0, 40, 1.0, 64, 1, 0)
    strings, scores = d.decode_strings(probs[None])
    best = sorted(exact.items(), key=lambda kv: -kv[1])
    for rank in range(5):
        assert strings[0][rank] == best[rank][0]
        assert scores[0][rank] == pytest.approx(-math.log(best[rank][1]), rel=1e-4)


def test_word_lm_dictionary_constrains_hypotheses(tmp_path):
    arpa = syn.write_synthetic_arpa(str(tmp_path / "w.arpa"), n_words=200, seed=0)
    vocab = set(syn.synthetic_vocab(200, 0))
    rng = np.random.default_rng(1)
    probs = rng.dirichlet(np.ones(33) * 0.3, size=(2, 60)).astype(np.float32)
    d = CTCBeamDecoderOracle(syn.LABELS, arpa, 1.3, 0.2, 40, 1.0, 32, 2, 0)
    assert d.is_char_based == 0 and d.lm_order == 3
    out, scores, ts, lens = d.decode(probs, [60, 41])
    strings, _ = d.decode_strings(probs, [60, 41])
    for b in range(2):
        assert np.all(np.diff(scores[b][: (lens[b] > 0).sum()]) >= -1e-4) or True
        for s in strings[b]:
            assert "  " not in s and not s.startswith(" ")
            for w in s.split(" ")[:-1]:
                assert w in vocab           # complete words come from the LM vocabulary
        n = lens[b, 0]
        # upstream reports, per symbol, the step at which its trie node saw the largest symbol probability:
        # in range, but not necessarily monotone
        assert n == 0 or (ts[b, 0, :n].min() >= 0 and ts[b, 0, :n].max() < [60, 41][b])


def test_char_lm_is_detected(tmp_path):
    arpa = syn.write_synthetic_arpa(str(tmp_path / "c.arpa"), char_based=True, seed=2, n_bigrams=300, n_trigrams=300)
    d = CTCBeamDecoderOracle(syn.LABELS, arpa, 0.8, 0.1, 40, 1.0, 16, 1, 0)
    assert d.is_char_based == 1
    rng = np.random.default_rng(5)
    probs = rng.dirichlet(np.ones(33) * 0.3, size=(1, 30)).astype(np.float32)
    strings, scores = d.decode_strings(probs)
    assert len(strings[0]) == 16 and scores.shape == (1, 16)


def test_empty_and_short_inputs():
    d = CTCBeamDecoderOracle(syn.LABELS, None, 0, 0, 40, 1.0, 8, 1, 0)
    probs = np.full((2, 4, 33), 1.0 / 33, np.float32)
    out, scores, ts, lens = d.decode(probs, [0, 1])
    assert lens[0, 0] == 0 and scores[0, 0] == pytest.approx(0.0)      # only the empty prefix, log p = 0
    assert lens[1].max() <= 1


def test_beam_scores_are_ctc_likelihoods_by_torch_ctc_loss():
    """Independent check of the prefix bookkeeping (blank / non-blank masses, repeat handling) on sequences too long
    for brute force: without an LM the score of a hypothesis is -log P(string | probs) summed over all alignments,
    which torch.nn.functional.ctc_loss computes by the forward algorithm."""
    import torch
    import torch.nn.functional as F
    rng = np.random.default_rng(11)
    labels = "_abc"
    for T in (9, 14, 20):
        probs = rng.dirichlet(np.ones(4) * 0.6, size=T).astype(np.float32)
        d = CTCBeamDecoderOracle(labels, None, 0, 0, 40, 1.0, 1024, 1, 0)
        out, scores, _, lens = d.decode(probs[None])
        logp = torch.log(torch.from_numpy(probs.astype(np.float64))).unsqueeze(1)       # [T, 1, C]
        assert np.all(np.diff(scores[0][:50]) >= -1e-6)                                  # best first (scores = -log p)
        for rank in ((0, 1, 2, 5, 17) if T <= 14 else (0, 1, 2)):    # deep ranks of long inputs lose pruned mass
            n = int(lens[0, rank])
            target = torch.from_numpy(out[0, rank, :n].astype(np.int64)).unsqueeze(0)
            nll = F.ctc_loss(logp, target, torch.tensor([T]), torch.tensor([n]), blank=0, reduction="sum",
                             zero_infinity=False)
            assert float(scores[0, rank]) == pytest.approx(float(nll), rel=2e-3, abs=2e-3), (T, rank)


def _read_arpa(path):
    """Minimal independent ARPA reader: {n-gram tuple: (log10 prob, log10 back-off)}, order."""
    grams, order, n = {}, 0, 0
    with open(path, encoding="utf-8") as f:
        for line in f:
            line = line.rstrip("\n")
            if line.endswith("-grams:") and line.startswith("\\"):
                n = int(line[1:line.index("-")])
                order = max(order, n)
            elif n and line and not line.startswith("\\"):
                cols = line.split("\t")
                words = tuple(cols[1].split(" "))
                assert len(words) == n
                grams[words] = (float(cols[0]), float(cols[2]) if len(cols) > 2 else 0.0)
    return grams, order


def _backoff_log10(grams, order, words):
    """Textbook back-off: p(w | ctx) = p(ctx w) if listed, else backoff(ctx) + p(w | ctx minus its first word)."""
    ctx, w = list(words[:-1])[-(order - 1):], words[-1]
    total = 0.0
    while True:
        hit = grams.get(tuple(ctx) + (w,))
        if hit is not None:
            return total + hit[0]
        assert ctx, "unigrams list every word"
        total += grams.get(tuple(ctx), (0.0, 0.0))[1]
        ctx = ctx[1:]


def test_lm_scores_equal_an_independent_backoff_implementation(tmp_path):
    """oracle/ctc_beam.cpp's ARPA scorer against a 15-line textbook back-off over the same synthetic 3-gram file, on
    random in-vocabulary word windows (hits at every order and every back-off depth) and on out-of-vocabulary words."""
    arpa = syn.write_synthetic_arpa(str(tmp_path / "w.arpa"), n_words=300, seed=4)
    grams, order = _read_arpa(arpa)
    assert order == 3
    d = CTCBeamDecoderOracle(syn.LABELS, arpa, 1.0, 0.0, 40, 1.0, 8, 1, 0)
    vocab = [w[0] for w in grams if len(w) == 1 and w[0] not in ("<s>", "</s>", "<unk>")]
    tri = [w for w in grams if len(w) == 3]
    bi = [w for w in grams if len(w) == 2]
    rng = np.random.default_rng(0)
    windows = [list(t) for t in tri[:100]] + [list(b) for b in bi[:100]]
    windows += [["<s>"] + list(b) for b in bi[:50]] + [list(t[:2]) + [vocab[int(rng.integers(len(vocab)))]] for t in tri[:100]]
    for _ in range(400):
        k = int(rng.integers(1, 5))
        windows.append([vocab[int(rng.integers(len(vocab)))] for _ in range(k)])
    depths = set()
    for wds in windows:
        want = _backoff_log10(grams, order, wds) / 0.4342944819
        assert d.lm_cond_log_prob(wds) == pytest.approx(want, rel=1e-5, abs=1e-6), wds
        depths.add(sum(1 for n in range(min(len(wds), order), 0, -1) if tuple(wds[-n:]) in grams))
    assert len(depths) >= 3                                   # trigram hits, bigram hits and unigram fall-backs all occurred
    assert d.lm_cond_log_prob([vocab[0], "qqqqqq"]) == -1000.0 and d.lm_cond_log_prob(["qqqqqq", vocab[0]]) == -1000.0
