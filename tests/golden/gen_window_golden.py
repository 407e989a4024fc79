"""Golden vectors for the non-default windows of audio_conf["window"] (hann, blackman, bartlett): the UNMODIFIED
reference parsers (danspeech/audio/parsers.py:9-10, :37-72, :75-170, imported through oracle/refharness.py, STFT via the
librosa restatement) on seeded synthetic audio.  Writes tests/golden/reference_windows.npz.

    python tests/golden/gen_window_golden.py        # build container only (/root/reference)
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from oracle import refharness  # noqa: E402
from danspeech_b200.utils import synthetic as syn  # noqa: E402


def main():
    refharness.import_reference()
    from danspeech.audio.parsers import InferenceSpectrogramAudioParser, SpectrogramAudioParser
    out = {}
    a = syn.synthetic_audio(8640 + 6240 * 2 + 700, seed=321)
    out["audio"] = a.astype(np.int16)
    for w in ("hann", "blackman", "bartlett"):
        conf = dict(normalize=True, sampling_rate=16000, window=w, window_stride=0.01, window_size=0.02)
        out["spect_" + w] = SpectrogramAudioParser(conf).parse_audio(a).numpy()
        sp = InferenceSpectrogramAudioParser(conf)
        chunks = [a[:8640], a[8640:8640 + 6240], a[8640 + 6240:]]
        for i, c in enumerate(chunks):
            out["stream_%s_%d" % (w, i)] = sp.parse_audio(c, is_last=(i == 2)).numpy()
    np.savez_compressed(os.path.join(HERE, "reference_windows.npz"), **out)
    print("wrote reference_windows.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
