"""Oracle / test infrastructure: writes an ARPA model in KenLM's probing binary layout ("format version 5").

KenLM (the library behind ctcdecode's Scorer, which the reference hands a ``.klm`` path to:
danspeech/language_models/dsl_3gram.py:16-20, DanSpeechRecognizer.py:89-92, decoder.py:99-100) is an absent third-party
dependency, so no real ``build_binary`` output exists here.  This writer restates what ``build_binary probing`` stores,
independently of the C++ reader in danspeech_b200/csrc/lm_load.cu (numpy, table construction by insertion):
lm/binary_format.cc (Sanity, FixedWidthParameters, counts), lm/vocab.cc (ProbingVocabulary: MurmurHash64A keys, ids in
unigram order, <unk> = 0), lm/search_hashed.{hh,cc} (unigram array, probing tables per order, CombineWordHash keys
chained from the last word backwards, prob sign bit = "extends left", backoff -0.0 = "does not extend right"),
util/probing_hash_table.hh (slot = key % buckets, linear probing, key 0 = empty).  PARITY UNPINNED against KenLM itself.
"""
import struct

import numpy as np

MAGIC = b"mmap lm http://kheafield.com/code format version 5\n\0"
M64 = (1 << 64) - 1


def murmur_hash64a(data, seed=0):
    m, r = 0xc6a4a7935bd1e995, 47
    h = (seed ^ (len(data) * m)) & M64
    n8 = len(data) // 8
    for i in range(n8):
        k = struct.unpack_from("<Q", data, 8 * i)[0]
        k = (k * m) & M64
        k ^= k >> r
        k = (k * m) & M64
        h ^= k
        h = (h * m) & M64
    tail = data[8 * n8:]
    if tail:
        for i in reversed(range(len(tail))):
            h ^= tail[i] << (8 * i)
        h = (h * m) & M64
    h ^= h >> r
    h = (h * m) & M64
    h ^= h >> r
    return h


def combine_word_hash(current, nxt):
    return ((current * 8978948897894561157) & M64) ^ (((1 + nxt) * 17894857484156487943) & M64)


def chain_hash(ids):
    h = ids[-1]
    for w in reversed(ids[:-1]):
        h = combine_word_hash(h, w)
    return h


def parse_arpa(path):
    """-> (order, words (id order, <unk> first), grams[n] = list of (ids tuple, prob, backoff or None))"""
    words, vocab, grams, cur = ["<unk>"], {"<unk>": 0, "<UNK>": 0}, {}, 0
    for line in open(path, encoding="utf-8"):
        line = line.strip()
        if not line:
            continue
        if line.startswith("\\"):
            if line.endswith("-grams:"):
                cur = int(line[1:line.index("-")])
                grams[cur] = []
            elif line == "\\end\\":
                break
            continue
        if cur == 0:
            continue
        tok = line.split()
        prob = float(tok[0])
        backoff = float(tok[cur + 1]) if len(tok) > cur + 1 else None
        ids = []
        for w in tok[1:1 + cur]:
            if cur == 1 and w not in vocab:
                vocab[w] = len(words)
                words.append(w)
            ids.append(vocab.get(w, 0))
        grams[cur].append((tuple(ids), prob, backoff))
    return max(grams), words, grams


def _buckets(entries, mult):
    return max(entries + 1, int(np.float32(mult) * np.float32(entries)))


def _probe_insert(keys, key):
    n = len(keys)
    slot = key % n
    while keys[slot] != 0:
        slot = (slot + 1) % n
    keys[slot] = key
    return slot


def write_klm(arpa_path, out_path, probing_multiplier=1.5, model_type=0, include_vocab=True):
    order, words, grams = parse_arpa(arpa_path)
    counts = [len(grams[n]) for n in range(1, order + 1)]
    # n-grams that are extended to the left by a longer one lose the sign bit of their probability (lm/search_hashed.cc:
    # MarkExtends); contexts without an explicit back-off get -0.0 (kNoExtensionBackoff)
    extended = set()
    for n in range(2, order + 1):
        for ids, _, _ in grams[n]:
            extended.add(ids[1:])

    def stored_prob(ids, p):
        return abs(p) if ids in extended else -abs(p)

    def stored_backoff(b):
        return -0.0 if b is None else (b if b != 0 else 0.0)

    head = MAGIC.ljust(56, b"\0") + struct.pack("<fffII4xQ", 0.0, 1.0, -0.5, 1, 0xFFFFFFFF, 1)
    assert len(head) == 88
    head += struct.pack("<B3xfiB3xI", order, probing_multiplier, model_type, 1 if include_vocab else 0, 0)
    head += struct.pack("<%dQ" % order, *counts)
    head = head.ljust((len(head) + 7) // 8 * 8, b"\0")

    vb = _buckets(counts[0], probing_multiplier)
    vkeys, vvals = [0] * vb, [0] * vb
    for i, w in enumerate(words):
        if i == 0:
            continue
        s = _probe_insert(vkeys, murmur_hash64a(w.encode("utf-8")))
        vvals[s] = i
    vocab = struct.pack("<II", 0, len(words)) + b"".join(struct.pack("<QI", k, v) for k, v in zip(vkeys, vvals))

    uni = np.zeros((counts[0] + 1, 2), np.float32)
    uni[0] = (-100.0, 0.0)                       # unknown_missing_logprob until the ARPA says otherwise
    for ids, p, b in grams[1]:
        uni[ids[0]] = (stored_prob(ids, p), stored_backoff(b))
    search = uni.tobytes()
    for n in range(2, order + 1):
        nb = _buckets(counts[n - 1], probing_multiplier)
        keys, vals = [0] * nb, [None] * nb
        for ids, p, b in grams[n]:
            s = _probe_insert(keys, chain_hash(list(ids)))
            vals[s] = (stored_prob(ids, p), stored_backoff(b))
        if n < order:
            search += b"".join(struct.pack("<Qff", k, *(v or (0.0, 0.0))) for k, v in zip(keys, vals))
        else:
            search += b"".join(struct.pack("<Qf", k, (v or (0.0, 0.0))[0]) for k, v in zip(keys, vals))
    strings = b"".join(w.encode("utf-8") + b"\0" for w in words) if include_vocab else b""
    with open(out_path, "wb") as f:
        f.write(head + vocab + search + strings)
    return {"order": order, "counts": counts, "words": len(words)}
