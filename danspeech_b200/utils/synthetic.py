"""Deterministic synthetic weights, audio and language models for the named model shapes.

Pretrained DanSpeech weights and KenLM files are GitHub-release downloads
(danspeech/pretrained_models/*.py, danspeech/language_models/*.py) that cannot be fetched offline;
benchmarks and tests therefore use random-init weights of the named architectures, seeded synthetic
16 kHz audio at int16 scale and a small synthetic ARPA LM (BASELINE.json north_star).

Everything here is generated with numpy's PCG64 so the same seed gives the same bits on any machine
(both for this package and for the reference modules the oracle loads the state dict into).
"""
from collections import OrderedDict

import numpy as np
import torch

LABELS = "_abcdefghijklmnopqrstuvwxyzæøåéü "

# name -> constructor kwargs (shapes from danspeech/pretrained_models/*.py)
MODEL_SHAPES = {
    "TestModel": dict(conv_layers=2, rnn_layers=5, rnn_hidden_size=400, bidirectional=True),
    "Baseline": dict(conv_layers=2, rnn_layers=5, rnn_hidden_size=800, bidirectional=True),
    "TransferLearned": dict(conv_layers=2, rnn_layers=5, rnn_hidden_size=800, bidirectional=True),
    "EnglishLibrispeech": dict(conv_layers=2, rnn_layers=5, rnn_hidden_size=800, bidirectional=True),
    "DanSpeechPrimary": dict(conv_layers=3, rnn_layers=9, rnn_hidden_size=1200, bidirectional=True),
    "Folketinget": dict(conv_layers=3, rnn_layers=9, rnn_hidden_size=1200, bidirectional=True),
    "CPUStreamingRNN": dict(conv_layers=2, rnn_layers=5, rnn_hidden_size=800, bidirectional=False, context=20,
                            streaming_inference_model=True),
    "GPUStreamingRNN": dict(conv_layers=2, rnn_layers=5, rnn_hidden_size=2000, bidirectional=False, context=20,
                            streaming_inference_model=True),
}

_CONV = [(1, 32, 41), (32, 32, 21), (32, 96, 21)]
_GATES = {"gru": 3, "lstm": 4, "rnn": 1}


def _bn(rng, n, prefix, sd):
    # randomised statistics: identity BatchNorm would hide folding bugs
    sd[prefix + ".weight"] = rng.uniform(0.5, 1.5, n)
    sd[prefix + ".bias"] = rng.normal(0.0, 0.2, n)
    sd[prefix + ".running_mean"] = rng.normal(0.0, 0.5, n)
    sd[prefix + ".running_var"] = rng.uniform(0.5, 2.0, n)
    sd[prefix + ".num_batches_tracked"] = np.array(0, dtype=np.int64)


def make_state_dict(conv_layers=2, rnn_layers=5, rnn_hidden_size=400, bidirectional=True, rnn_type="gru", context=20,
                    streaming_inference_model=False, num_classes=len(LABELS), seed=0, fc_scale=10.0,
                    ih_scale=None):
    """State dict with the reference's names and shapes, float32 torch tensors."""
    if ih_scale is None:
        # deeper random stacks amplify perturbations more: with x4 a 9-layer stack already turns plain bf16
        # rounding of weights and layer inputs (emulated on the CPU) into ~3e-2 logit error
        ih_scale = 4.0 if rnn_layers <= 5 else 2.5
    rng = np.random.default_rng(seed)
    sd = OrderedDict()
    H = rnn_hidden_size
    freq = 161
    for i, (cin, cout, kh) in enumerate(_CONV[:conv_layers]):
        bound = 1.0 / np.sqrt(cin * kh * 11)
        sd["conv.seq_module.%d.weight" % (3 * i)] = rng.uniform(-bound, bound, (cout, cin, kh, 11)) * 3.0
        sd["conv.seq_module.%d.bias" % (3 * i)] = rng.uniform(-bound, bound, cout)
        _bn(rng, cout, "conv.seq_module.%d" % (3 * i + 1), sd)
    n_for_size = 2 if streaming_inference_model else conv_layers
    for i, (cin, cout, kh) in enumerate(_CONV[:n_for_size]):
        freq = (freq + 2 * (kh // 2) - kh) // 2 + 1
    in_size = freq * _CONV[n_for_size - 1][1]
    G = _GATES[rnn_type]
    dirs = 2 if (bidirectional and not streaming_inference_model) else 1
    bound = 1.0 / np.sqrt(H)
    for l in range(rnn_layers):
        isz = in_size if l == 0 else H
        if l > 0:
            _bn(rng, isz, "rnns.%d.batch_norm.module" % l, sd)
        for d in range(dirs):
            sfx = "_reverse" if d == 1 else ""
            # input drive x4: a random GRU stack otherwise forgets its input and the argmax path degenerates
            # to one symbol (nothing to decode); much larger gains make the random network chaotic, so that
            # bf16 rounding alone (emulated on the CPU) already exceeds the 2e-2 bar
            sd["rnns.%d.rnn.weight_ih_l0%s" % (l, sfx)] = rng.uniform(-bound, bound, (G * H, isz)) * ih_scale
            sd["rnns.%d.rnn.weight_hh_l0%s" % (l, sfx)] = rng.uniform(-bound, bound, (G * H, H))
            sd["rnns.%d.rnn.bias_ih_l0%s" % (l, sfx)] = rng.uniform(-bound, bound, G * H)
            sd["rnns.%d.rnn.bias_hh_l0%s" % (l, sfx)] = rng.uniform(-bound, bound, G * H)
    if streaming_inference_model:
        sd["lookahead.conv.weight"] = rng.uniform(-0.3, 0.3, (H, 1, context))
    elif not bidirectional:
        sd["lookahead.0.conv.weight"] = rng.uniform(-0.3, 0.3, (H, 1, context))
    _bn(rng, H, "fc.0.module.0", sd)
    sd["fc.0.module.0.bias"] = sd["fc.0.module.0.bias"] * 0.1
    sd["fc.0.module.0.running_mean"] = sd["fc.0.module.0.running_mean"] * 0.1
    # peaky FC: large argmax margins so greedy bit-exactness does not hinge on 1e-7 near-ties
    sd["fc.0.module.1.weight"] = rng.uniform(-bound, bound, (num_classes, H)) * fc_scale
    out = OrderedDict()
    for k, v in sd.items():
        t = torch.from_numpy(np.asarray(v))
        out[k] = t if t.dtype == torch.int64 else t.to(torch.float32)
    return out


def synthetic_audio(n_samples, seed):
    """Speech-like band-limited noise x slow envelope at int16 scale, float64, integer valued, clipped."""
    rng = np.random.default_rng(1234 + seed)
    x = rng.standard_normal(n_samples + 64)
    k = np.hanning(33)
    x = np.convolve(x, k / k.sum(), mode="same")[:n_samples]
    t = np.arange(n_samples) / 16000.0
    env = 0.55 + 0.45 * np.sin(2 * np.pi * (1.3 + 0.4 * (seed % 5)) * t + seed)
    y = x * env
    y = y / (y.std() + 1e-9) * 3000.0
    return np.clip(np.rint(y), -32768, 32767).astype(np.float64)


def synthetic_vocab(n_words=2000, seed=0, char_based=False):
    rng = np.random.default_rng(seed)
    letters = [c for c in LABELS if c not in "_ "]
    if char_based:
        return letters
    words = set()
    while len(words) < n_words:
        n = int(rng.integers(1, 7))
        words.add("".join(rng.choice(letters, n)))
    return sorted(words)


def write_synthetic_arpa(path, n_words=2000, seed=0, char_based=False, n_bigrams=6000, n_trigrams=6000):
    """Seeded 3-gram ARPA text LM (log10 probs in [-5,-0.1], back-offs in [-1,0]) -- SURVEY 8d."""
    rng = np.random.default_rng(seed + 7)
    vocab = synthetic_vocab(n_words, seed, char_based)
    toks = ["<unk>", "<s>", "</s>"] + vocab
    uni = []
    for w in toks:
        p = -99.0 if w == "<s>" else float(np.round(rng.uniform(-5.0, -0.1), 4))
        bo = 0.0 if w == "</s>" else float(np.round(rng.uniform(-1.0, 0.0), 4))
        uni.append((p, w, bo))
    ctx_words = ["<s>"] + vocab
    nxt_words = vocab + ["</s>"]
    bigrams = {}
    target_bi = min(n_bigrams, len(ctx_words) * len(nxt_words))
    while len(bigrams) < target_bi:
        a = ctx_words[int(rng.integers(len(ctx_words)))]
        b = nxt_words[int(rng.integers(len(nxt_words)))]
        bigrams[(a, b)] = (float(np.round(rng.uniform(-5.0, -0.1), 4)), float(np.round(rng.uniform(-1.0, 0.0), 4)))
    bkeys = sorted(bigrams)
    trigrams = {}
    cont = [k for k in bkeys if k[1] != "</s>"]
    target_tri = min(n_trigrams, len(cont) * 8)
    while len(trigrams) < target_tri and cont:
        a, b = cont[int(rng.integers(len(cont)))]
        c = nxt_words[int(rng.integers(len(nxt_words)))]
        trigrams[(a, b, c)] = float(np.round(rng.uniform(-5.0, -0.1), 4))
    with open(path, "w", encoding="utf-8") as f:
        f.write("\\data\\\nngram 1=%d\nngram 2=%d\nngram 3=%d\n\n" % (len(uni), len(bigrams), len(trigrams)))
        f.write("\\1-grams:\n")
        for p, w, bo in uni:
            f.write("%.4f\t%s\t%.4f\n" % (p, w, bo))
        f.write("\n\\2-grams:\n")
        for (a, b) in bkeys:
            p, bo = bigrams[(a, b)]
            if b == "</s>":
                f.write("%.4f\t%s %s\n" % (p, a, b))
            else:
                f.write("%.4f\t%s %s\t%.4f\n" % (p, a, b, bo))
        f.write("\n\\3-grams:\n")
        for (a, b, c) in sorted(trigrams):
            f.write("%.4f\t%s %s %s\n" % (trigrams[(a, b, c)], a, b, c))
        f.write("\n\\end\\\n")
    return path
