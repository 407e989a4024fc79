"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):   python tests/golden/gen_golden.py
The reference is imported through oracle/refharness.py (stubs for the absent third-party modules).
Weights are regenerated at test time from the same seeds (danspeech_b200.utils.synthetic), so only
inputs that cannot be regenerated (WAV-derived audio) and the reference OUTPUTS are stored.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import refharness  # noqa: E402
from danspeech_b200.utils import synthetic as syn  # noqa: E402

torch.set_num_threads(8)


def ref_model(ref, name, rnn_type="gru", seed=0, **over):
    from danspeech.deepspeech.model import DeepSpeech, supported_rnns
    cfg = dict(syn.MODEL_SHAPES[name])
    cfg.update(over)
    m = DeepSpeech(model_name=name, rnn_type=supported_rnns[rnn_type], labels=syn.LABELS,
                   rnn_hidden_size=cfg["rnn_hidden_size"], rnn_layers=cfg["rnn_layers"],
                   bidirectional=cfg["bidirectional"], context=cfg.get("context", 20), conv_layers=cfg["conv_layers"],
                   streaming_inference_model=cfg.get("streaming_inference_model", False))
    m.load_state_dict(syn.make_state_dict(rnn_type=rnn_type, seed=seed, **cfg))
    return m.eval()


def main():
    ref = refharness.import_reference()
    from danspeech.audio.resources import load_audio
    from danspeech.audio.parsers import SpectrogramAudioParser, InferenceSpectrogramAudioParser
    from danspeech.deepspeech.decoder import GreedyDecoder
    from danspeech import Recognizer

    out = {}
    # ---- audio fixtures derived from the reference's example WAVs through its own loader (clip(L+R)) ----
    wavs = {}
    for name in ("u0013002", "u0042018"):
        a = load_audio(os.path.join(refharness.REFERENCE_ROOT, "example_files", name + ".wav"))
        assert np.all(a == np.rint(a)) and np.abs(a).max() <= 32768
        wavs[name] = a
        out["wav_" + name] = a.astype(np.int16)

    # ---- spectrograms through the reference parser (STFT via the librosa restatement: parity unpinned) ----
    parser = SpectrogramAudioParser(syn_audio_conf())
    for name, a in wavs.items():
        out["spect_" + name] = parser.parse_audio(a).numpy()
    for i, n in enumerate((161, 1000, 16000, 40001)):
        a = syn.synthetic_audio(n, seed=100 + i)
        out["spect_syn%d" % n] = parser.parse_audio(a).numpy()

    # ---- streaming parser over the engine's chunk schedule (Recognizer.py:602-611) ----
    sp = InferenceSpectrogramAudioParser(syn_audio_conf())
    a = wavs["u0013002"]
    chunks = [a[:8640]] + [a[8640 + 6240 * i: 8640 + 6240 * (i + 1)] for i in range(12)]
    chunks = [c for c in chunks if len(c) > 0]
    for i, c in enumerate(chunks):
        s = sp.parse_audio(c, is_last=(i == len(chunks) - 1))
        out["stream_spect_%d" % i] = s.numpy() if len(s) else np.zeros((0,), np.float32)
    out["stream_n_chunks"] = np.array(len(chunks))

    # ---- config 1: TestModel-shaped, u0013002.wav, greedy, through Recognizer.recognize ----
    with torch.no_grad():
        m = ref_model(ref, "TestModel", seed=0)
        r = Recognizer(model=m)
        text = r.recognize(wavs["u0013002"])
        spect = parser.parse_audio(wavs["u0013002"])
        probs, sizes = m(spect.view(1, 1, 161, -1), torch.IntTensor([spect.size(1)]))
        out["cfg1_probs"] = probs.numpy()
        out["cfg1_sizes"] = sizes.numpy()
        out["cfg1_text"] = np.array(text)

        # ---- ragged batch (3 utterances), every rnn type, uni + bi, 1/2/3 conv: small shapes ----
        lens = [16000, 11111, 4000]
        auds = [syn.synthetic_audio(n, seed=7 + i) for i, n in enumerate(lens)]
        specs = [parser.parse_audio(a) for a in auds]
        Tm = specs[0].size(1)
        x = torch.zeros(3, 1, 161, Tm)
        for i, s in enumerate(specs):
            x[i, 0, :, : s.size(1)] = s
        xl = torch.IntTensor([s.size(1) for s in specs])
        gd = GreedyDecoder(syn.LABELS, blank_index=0)
        for tag, name, kw in [
            ("gru_bi_c2", "TestModel", dict(rnn_type="gru", rnn_hidden_size=96, rnn_layers=3)),
            ("gru_bi_c3", "DanSpeechPrimary", dict(rnn_type="gru", rnn_hidden_size=80, rnn_layers=2)),
            ("gru_bi_c1", "TestModel", dict(rnn_type="gru", rnn_hidden_size=64, rnn_layers=2, conv_layers=1)),
            ("lstm_bi_c2", "TestModel", dict(rnn_type="lstm", rnn_hidden_size=72, rnn_layers=2, ih_scale=2.5)),
            ("rnn_bi_c2", "TestModel", dict(rnn_type="rnn", rnn_hidden_size=72, rnn_layers=2)),
            ("gru_uni_c2", "TestModel", dict(rnn_type="gru", rnn_hidden_size=96, rnn_layers=2, bidirectional=False,
                                             context=20)),
        ]:
            rt = kw.pop("rnn_type")
            m = ref_model(ref, name, rnn_type=rt, seed=3, **kw)
            probs, sizes = m(x, xl)
            strings, offs = gd.decode(probs, sizes)
            out["batch_%s_probs" % tag] = probs.numpy()
            out["batch_%s_sizes" % tag] = sizes.numpy()
            out["batch_%s_text" % tag] = np.array([s[0] for s in strings])
            out["batch_%s_offs" % tag] = np.concatenate([o[0].numpy() for o in offs]).astype(np.int32)

        # ---- streaming model, CPUStreamingRNN-shaped but narrow (H=128), engine chunk schedule ----
        m = ref_model(ref, "CPUStreamingRNN", seed=5, rnn_hidden_size=128, rnn_layers=3)
        sp = InferenceSpectrogramAudioParser(syn_audio_conf())
        a = syn.synthetic_audio(8640 + 6240 * 4 + 3000, seed=42)
        chunks = [a[:8640]] + [a[8640 + 6240 * i: 8640 + 6240 * (i + 1)] for i in range(5)]
        for i, c in enumerate(chunks):
            last = i == len(chunks) - 1
            s = sp.parse_audio(c, is_last=last)
            o = m(s.view(1, 1, 161, -1), i == 0, last)
            out["smodel_probs_%d" % i] = o.numpy() if o is not None else np.zeros((0,), np.float32)
        out["smodel_audio"] = a.astype(np.int16)

    np.savez_compressed(os.path.join(HERE, "reference_outputs.npz"), **out)
    print("wrote", os.path.join(HERE, "reference_outputs.npz"), "with", len(out), "arrays;",
          "cfg1 text = %r" % text)


def syn_audio_conf():
    return dict(normalize=True, sampling_rate=16000, window="hamming", window_stride=0.01, window_size=0.02)


if __name__ == "__main__":
    main()
