"""Golden vectors for the energy VAD: runs the UNMODIFIED reference generator ``Recognizer.listen_stream``
(danspeech/Recognizer.py:218-324) over oracle.vad.fixture_pcm() through a fake ``SpeechSource`` and stores
every yield as (is_last, buffers in the yield, source position in buffers after the yield).

Run in the build container only:  python tests/golden/gen_vad_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import refharness, vad  # noqa: E402


def reference_yields(pcm, chunk=1024, energy_threshold=1000):
    refharness.import_reference()
    from danspeech import Recognizer
    from danspeech.audio.resources import SpeechSource

    class FakeSource(SpeechSource):
        def __init__(self):
            self.pos, self.chunk, self.sampling_rate, self.sampling_width = 0, chunk, 16000, 2
            src = self

            class Stream:
                def read(self, n):
                    b = pcm[src.pos:src.pos + n]
                    src.pos += n
                    return b.tobytes() if len(b) == n else b""
            self.stream = Stream()

    r = Recognizer.__new__(Recognizer)   # listen_stream only reads these attributes (Recognizer.py:42-56)
    r.energy_threshold, r.pause_threshold, r.phrase_threshold, r.non_speaking_duration = energy_threshold, 0.8, 0.3, 0.35
    r.stream = True
    src = FakeSource()
    rows = []
    while src.pos < len(pcm):
        for is_last, data in r.listen_stream(src):   # one generator per phrase, as the engine uses it
            n = len(data) if isinstance(data, list) else (len(data) // (2 * chunk))
            rows.append((int(is_last), n, src.pos // chunk))
            if is_last:
                break
    return np.array(rows, dtype=np.int32)


def main():
    pcm = vad.fixture_pcm()
    rows = reference_yields(pcm)
    np.savez_compressed(os.path.join(HERE, "vad_reference.npz"), yields=rows)
    print("wrote vad_reference.npz:", rows.shape, rows[:4].tolist(), "...", rows[-2:].tolist())


if __name__ == "__main__":
    main()
