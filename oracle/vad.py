"""Oracle (TEST INFRASTRUCTURE ONLY): CPU restatement of the phrase state machine of
``Recognizer.listen_stream`` (danspeech/Recognizer.py:218-324) as an event code per buffer.

Pinned: tests/test_oracle_cpu.py drives the UNMODIFIED reference generator with a fake ``SpeechSource``
(when /root/reference is mounted) and checks that its yields are exactly what these events imply; the
events of the same fixture are stored in tests/golden/reference_outputs.npz for the GPU box.
The energy is ``audioop.rms`` (Recognizer.py:275,304) restated in numpy: floor(sqrt(mean(x^2))).
"""
import math

import numpy as np

SILENCE, PHRASE_START, SPEECH, PHRASE_END, PHRASE_DROPPED = 0, 1, 2, 3, 4


def rms(buf_int16):
    """audioop.rms(fragment, 2): (unsigned int) sqrt(sum(x*x) / n) with a double accumulator."""
    x = np.asarray(buf_int16, dtype=np.int64)
    if x.size == 0:
        return 0
    return int(math.sqrt(float(np.sum(x * x)) / float(x.size)))


def buffer_counts(chunk=1024, sampling_rate=16000, pause_threshold=0.8, phrase_threshold=0.3,
                  non_speaking_duration=0.35):
    """Recognizer.py:244-250."""
    spb = float(chunk) / sampling_rate
    return (int(math.ceil(pause_threshold / spb)), int(math.ceil(phrase_threshold / spb)),
            int(math.ceil(non_speaking_duration / spb)))


class ListenStreamOracle:
    """One stream.  ``push(buffer)`` returns (energy, event) for one ``source.chunk`` of samples."""

    def __init__(self, energy_threshold=1000, pause_buffers=13, phrase_buffers=5):
        self.energy_threshold = energy_threshold
        self.pause_buffers = pause_buffers
        self.phrase_buffers = phrase_buffers
        self.in_phrase = False
        self.pause_count = 0
        self.phrase_count = 0

    def push(self, buf_int16):
        energy = rms(buf_int16)
        loud = energy > self.energy_threshold
        if not self.in_phrase:                      # Recognizer.py:262-277
            if loud:
                self.in_phrase = True
                self.pause_count = self.phrase_count = 0   # :286
                return energy, PHRASE_START
            return energy, SILENCE
        self.phrase_count += 1                      # :300
        self.pause_count = 0 if loud else self.pause_count + 1   # :306-309
        if self.pause_count > self.pause_buffers:   # :311
            self.in_phrase = False
            if self.phrase_count - self.pause_count >= self.phrase_buffers:   # :316-318
                return energy, PHRASE_END
            return energy, PHRASE_DROPPED
        return energy, SPEECH


def replay_yields(events, buffers, non_speaking_buffers):
    """What the reference generator yields for these events, up to and including the first (True, ...):
    list of (is_last, [buffers]) -- used to compare with the live reference."""
    out, pre = [], []
    for ev, b in zip(events, buffers):
        if ev == SILENCE:
            pre.append(b)
            if len(pre) > non_speaking_buffers:
                pre.pop(0)
        elif ev == PHRASE_START:
            pre.append(b)
            if len(pre) > non_speaking_buffers:
                pre.pop(0)
            out.append((False, list(pre)))
            pre = []
        elif ev == SPEECH:
            out.append((False, [b]))
        elif ev == PHRASE_END:
            out.append((True, [b]))
            break
        else:   # PHRASE_DROPPED: the generator starts over with an empty pre-roll (:258-260)
            pre = []
    return out


def fixture_pcm(seed=0, n_buffers=160, chunk=1024):
    """Deterministic int16 test signal: background noise, a long phrase with a short internal pause, a click
    that is too short to count as a phrase, and a second long phrase."""
    rng = np.random.default_rng(seed)
    n = n_buffers * chunk
    env = np.zeros(n)
    for a, b in ((20, 45), (50, 52), (75, 77), (100, 130)):
        env[a * chunk:b * chunk] = 1.0
    x = rng.normal(0, 3000, n) * env + rng.normal(0, 100, n)
    return np.clip(np.rint(x), -32768, 32767).astype(np.int16)
