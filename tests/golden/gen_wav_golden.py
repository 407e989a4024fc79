"""More real-audio golden vectors: four further example WAVs of the reference through its own loader
(``danspeech.audio.load_audio``: clip(L+R), resources.py:630-640) and through ``Recognizer.recognize`` of the UNMODIFIED
reference with a TestModel-shaped random-weight model (seed 0), greedy decoding.  Stored: the mono int16 audio, the
transcript, the spectrogram shape and two of its rows, and the first/last softmax rows.

Run in the build container only:  python tests/golden/gen_wav_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
from oracle import refharness  # noqa: E402
import gen_golden  # noqa: E402

NAMES = ("u0042008", "u0042012", "u0042017", "u0042019")


def main():
    ref = refharness.import_reference()
    from danspeech.audio.resources import load_audio
    from danspeech.audio.parsers import SpectrogramAudioParser
    from danspeech import Recognizer
    torch.set_num_threads(8)
    parser = SpectrogramAudioParser(gen_golden.syn_audio_conf())
    out = {}
    with torch.no_grad():
        m = gen_golden.ref_model(ref, "TestModel", seed=0)
        r = Recognizer(model=m)
        for name in NAMES:
            a = load_audio(os.path.join(refharness.REFERENCE_ROOT, "example_files", name + ".wav"))
            assert np.all(a == np.rint(a)) and np.abs(a).max() <= 32768
            spect = parser.parse_audio(a)
            probs, sizes = m(spect.view(1, 1, 161, -1), torch.IntTensor([spect.size(1)]))
            out["wav_" + name] = a.astype(np.int16)
            out["text_" + name] = np.array(r.recognize(a))
            out["spect_shape_" + name] = np.array(spect.shape)
            out["spect_rows_" + name] = spect[[3, 80]].numpy()
            out["probs_ends_" + name] = probs[0, [0, -1]].numpy()
            print(name, len(a), tuple(spect.shape), repr(str(out["text_" + name]))[:60])
    np.savez_compressed(os.path.join(HERE, "reference_wavs.npz"), **out)
    print("wrote reference_wavs.npz", os.path.getsize(os.path.join(HERE, "reference_wavs.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
