"""CPU: language-model files -> host model (danspeech_b200/csrc/lm_load.cu through dsb_lm_inspect, no CUDA call).

SURVEY 8f-1: every LM the reference hands the decoder is a KenLM binary (language_models/dsl_3gram.py:16-20).  KenLM
itself is absent, so the reader is pinned against a fixture generated here: the committed script writes an ARPA
(utils.synthetic.write_synthetic_arpa) and the independent numpy writer oracle/klm_writer.py lays the same model out as
a probing ``.klm``; both must load into the SAME model.
"""
import ctypes
import os
import struct

import pytest

from danspeech_b200 import _native as N
from danspeech_b200.utils import synthetic as syn
from oracle import klm_writer as kw


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    return N.lib()


def inspect(lib, path):
    o, n, w, d = ctypes.c_int(), ctypes.c_int64(), ctypes.c_int64(), ctypes.c_uint64()
    rc = lib.dsb_lm_inspect(str(path).encode(), o, n, w, d)
    return rc, lib.dsb_last_error().decode() if rc else "", (o.value, n.value, w.value, d.value)


@pytest.mark.parametrize("char_based,mult", [(False, 1.5), (True, 1.5), (False, 2.25)])
def test_klm_loads_into_the_same_model_as_its_arpa(lib, tmp_path, char_based, mult):
    arpa, klm = tmp_path / "m.arpa", tmp_path / "m.klm"
    syn.write_synthetic_arpa(str(arpa), n_words=500, seed=3, char_based=char_based, n_bigrams=1500, n_trigrams=1500)
    info = kw.write_klm(str(arpa), str(klm), probing_multiplier=mult)
    rc_a, msg_a, a = inspect(lib, arpa)
    rc_k, msg_k, k = inspect(lib, klm)
    assert rc_a == 0, msg_a
    assert rc_k == 0, msg_k
    assert a == k                                   # order, n-grams, words, digest of every (key, order, prob, backoff)
    assert a[0] == 3 and a[1] == sum(info["counts"]) and a[2] == info["words"]
    other = tmp_path / "other.arpa"
    syn.write_synthetic_arpa(str(other), n_words=500, seed=4, char_based=char_based, n_bigrams=1500, n_trigrams=1500)
    assert inspect(lib, other)[2][3] != a[3]        # the digest does tell models apart


def test_klm_hashes_restated_twice_agree(lib, tmp_path):
    """MurmurHash64A (vocabulary) and CombineWordHash (n-gram keys): the C++ reader finds every word of the numpy
    writer's table where its own hash says, with non-ASCII words and all tail lengths 0..7."""
    arpa, klm = tmp_path / "w.arpa", tmp_path / "w.klm"
    words = ["a", "ab", "abc", "abcd", "abcde", "abcdef", "abcdefg", "abcdefgh", "æøå", "smørrebrød", "x" * 23]
    with open(arpa, "w", encoding="utf-8") as f:
        f.write("\\data\\\nngram 1=%d\nngram 2=%d\n\n\\1-grams:\n" % (len(words) + 3, len(words)))
        for w in ["<unk>", "<s>", "</s>"] + words:
            f.write("-1.5\t%s\t-0.25\n" % w)
        f.write("\n\\2-grams:\n")
        for w in words:
            f.write("-0.75\t<s> %s\n" % w)
        f.write("\n\\end\\\n")
    kw.write_klm(str(arpa), str(klm))
    rc, msg, k = inspect(lib, klm)
    assert rc == 0, msg
    assert k == inspect(lib, arpa)[2] and k[2] == len(words) + 3
    assert kw.murmur_hash64a(b"") == 0
    assert kw.chain_hash([5]) == 5 and kw.chain_hash([1, 2, 3]) == kw.combine_word_hash(kw.combine_word_hash(3, 2), 1)


def test_klm_variants_this_reader_does_not_understand_are_refused_by_name(lib, tmp_path):
    arpa = tmp_path / "m.arpa"
    syn.write_synthetic_arpa(str(arpa), n_words=50, seed=1, n_bigrams=80, n_trigrams=80)
    for model_type, name in ((1, "rest-probing"), (2, "trie"), (3, "quantised trie"), (5, "quantised array trie")):
        p = tmp_path / ("t%d.klm" % model_type)
        kw.write_klm(str(arpa), str(p), model_type=model_type)
        rc, msg, _ = inspect(lib, p)
        assert rc == -5 and name in msg and "probing" in msg and "order 3" in msg, msg
    p = tmp_path / "novocab.klm"
    kw.write_klm(str(arpa), str(p), include_vocab=False)
    rc, msg, _ = inspect(lib, p)
    assert rc != 0 and ("vocabulary" in msg or "layout" in msg), msg
    good = tmp_path / "good.klm"
    kw.write_klm(str(arpa), str(good))
    raw = good.read_bytes()
    # another format version, a truncated file, a corrupted vocabulary slot, a corrupted n-gram count
    v4 = tmp_path / "v4.klm"
    v4.write_bytes(raw.replace(b"version 5", b"version 4", 1))
    rc, msg, _ = inspect(lib, v4)
    assert rc == -5 and "version '4'" in msg, msg
    cut = tmp_path / "cut.klm"
    cut.write_bytes(raw[: len(raw) // 2])
    rc, msg, _ = inspect(lib, cut)
    assert rc != 0 and "layout not understood" in msg, msg
    header = (108 + 8 * 3 + 7) // 8 * 8
    bad = bytearray(raw)
    for off in range(header + 8, header + 8 + 12 * 200, 12):          # first occupied vocabulary slot: flip its key
        if struct.unpack_from("<Q", bad, off)[0]:
            bad[off] ^= 0xFF
            break
    bv = tmp_path / "badvocab.klm"
    bv.write_bytes(bytes(bad))
    rc, msg, _ = inspect(lib, bv)
    assert rc == -5 and "layout not understood" in msg, msg
    bad = bytearray(raw)
    struct.pack_into("<Q", bad, 108 + 8, struct.unpack_from("<Q", bad, 108 + 8)[0] + 1)   # one more bigram than stored
    bc = tmp_path / "badcount.klm"
    bc.write_bytes(bytes(bad))
    rc, msg, _ = inspect(lib, bc)
    assert rc != 0 and "layout not understood" in msg, msg


def test_malformed_arpa_lines_are_rejected_not_skipped(lib, tmp_path):
    good = "\\data\\\nngram 1=3\nngram 2=1\n\n\\1-grams:\n-1.0\t<unk>\t-0.5\n-1.0\t<s>\t-0.5\n-2.0\ta\t-0.5\n\n" \
           "\\2-grams:\n-0.5\t<s> a\n\n\\end\\\n"
    p = tmp_path / "g.arpa"
    p.write_text(good)
    rc, msg, info = inspect(lib, p)
    assert rc == 0 and info[:3] == (2, 4, 3), msg
    cases = {
        "field count": good.replace("-0.5\t<s> a\n", "-0.5\t<s> a b c\n"),
        "probability": good.replace("-2.0\ta", "minus-two\ta"),
        "positive probability": good.replace("-2.0\ta", "2.0\ta"),
        "back-off": good.replace("-2.0\ta\t-0.5", "-2.0\ta\tzero"),
        "declares": good.replace("ngram 2=1", "ngram 2=2"),
        "without \\end\\": good.replace("\\end\\\n", ""),
        "unknown ARPA section": good.replace("\\2-grams:", "\\two-grams"),
    }
    for what, text in cases.items():
        q = tmp_path / "bad.arpa"
        q.write_text(text)
        rc, msg, _ = inspect(lib, q)
        assert rc == -6, (what, msg)
    rc, msg, _ = inspect(lib, tmp_path / "missing.arpa")
    assert rc == -6 and "cannot open" in msg
