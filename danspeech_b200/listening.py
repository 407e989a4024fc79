"""Host side of the listening API: energy VAD over a stream source, background capture, and the two generators that
feed the recognizer from it.

This is the single-stream caller of the hot path -- danspeech/Recognizer.py ``listen`` (:133-216), ``listen_stream``
(:218-324), ``get_audio_data`` (:326-336), ``listen_in_background`` (:338-397), ``enable_streaming`` /
``disable_streaming`` (:399-431), ``streaming`` (:433-497), ``real_time_streaming`` (:560-715), ``adjust_for_speech``
(:717-757), ``adjust_for_ambient_noise`` (:759-797), ``update_stream_parameters`` (:800-818) -- with the same
parameters, yields and quirks, so that scripts written against the reference run unchanged.  It is host logic (I/O
pacing and a few integer operations per 64 ms buffer); the many-stream, on-device form of the same state machine is
``streaming.StreamVAD`` / ``MultiStreamRecognizer``.  ``PhraseListener`` has no GPU dependency of its own and is
tested on CPU against the unmodified reference generators (tests/test_listening_cpu.py).
"""
import collections
import math
import threading
import time

import numpy as np

from .audio.resources import AudioData
from .errors.recognizer_errors import NoDataInBuffer, WaitTimeoutError, WrongUsageOfListen


def pcm_rms(fragment, width):
    """``audioop.rms``: floor(sqrt(mean(x^2))) of little-endian signed PCM (audioop leaves the stdlib in 3.13)."""
    if width not in (1, 2, 4):
        raise ValueError("unsupported sample width %r" % (width,))
    x = np.frombuffer(fragment, dtype="<i%d" % width).astype(np.int64)   # audioop treats every width as signed
    if x.size == 0:
        return 0
    if width == 4:      # squares exceed 2^53: add them one by one in double precision, as audioop does
        total = 0.0
        for v in x.tolist():
            total += float(v) * float(v)
    else:
        total = float(np.sum(x * x))                                      # exact in int64, exact as a double
    return int(math.sqrt(total / float(x.size)))


def _check_source(source, what):
    ok = all(hasattr(source, a) for a in ("stream", "chunk", "sampling_rate", "sampling_width"))
    assert ok, "Source must be an audio source"
    assert source.stream is not None, ("Audio source must be entered before %s; are you using ``source`` outside of "
                                       "a ``with`` statement?" % what)


class PhraseListener(object):
    """Phrase detection parameters and the loops built on them (mixed into ``Recognizer``)."""

    def _init_listening(self):
        self.energy_threshold = 1000                   # minimum RMS energy of a buffer that counts as speech
        self.pause_threshold = 0.8                     # seconds of quiet that end a phrase
        self.phrase_threshold = 0.3                    # minimum seconds of speech for a phrase (filters clicks)
        self.non_speaking_duration = 0.35              # seconds of quiet kept on both sides of a phrase
        self.mininum_required_speaking_seconds = 0.7   # (sic) `streaming` ignores shorter clips
        self.dynamic_energy_threshold = True
        self.dynamic_energy_adjustment_damping = 0.15
        self.dynamic_energy_ratio = 1.5
        self.stream = False
        self.stream_thread_stopper = None

    # ------------------------------------------------------------------ parameters
    def update_stream_parameters(self, energy_threshold=None, pause_threshold=None, phrase_threshold=None,
                                 non_speaing_duration=None):
        """Falsy arguments keep the current value; the last keyword keeps the reference's spelling."""
        for name, value in (("energy_threshold", energy_threshold), ("pause_threshold", pause_threshold),
                            ("phrase_threshold", phrase_threshold), ("non_speaking_duration", non_speaing_duration)):
            if value:
                setattr(self, name, value)

    def _buffer_counts(self, source):
        assert self.pause_threshold >= self.non_speaking_duration >= 0
        spb = float(source.chunk) / source.sampling_rate
        n = lambda seconds: int(math.ceil(seconds / spb))   # noqa: E731
        return spb, n(self.pause_threshold), n(self.phrase_threshold), n(self.non_speaking_duration)

    def _energies(self, source, duration):
        """Energy of every buffer read during ``duration`` seconds of stream time."""
        spb = (source.chunk + 0.0) / source.sampling_rate
        elapsed = spb
        while elapsed <= duration:
            yield pcm_rms(source.stream.read(source.chunk), source.sampling_width), spb
            elapsed += spb

    def adjust_for_speech(self, source, duration=4):
        """Threshold = mean buffer energy while somebody talks, minus 80 when that leaves something."""
        _check_source(source, "adjusting")
        assert self.pause_threshold >= self.non_speaking_duration >= 0
        levels = [e for e, _ in self._energies(source, duration)]
        mean = sum(levels) / len(levels)
        self.energy_threshold = mean - 80 if mean > 80 else mean

    def adjust_for_ambient_noise(self, source, duration=2):
        """Threshold follows 1.5 x the background energy with the damped average ``listen`` also uses."""
        _check_source(source, "adjusting")
        assert self.pause_threshold >= self.non_speaking_duration >= 0
        for energy, spb in self._energies(source, duration):
            self._follow_energy(energy, spb)

    def _follow_energy(self, energy, spb):
        keep = self.dynamic_energy_adjustment_damping ** spb
        self.energy_threshold = self.energy_threshold * keep + energy * self.dynamic_energy_ratio * (1 - keep)

    # ------------------------------------------------------------------ one phrase, blocking
    def listen(self, source, timeout=None, phrase_time_limit=None):
        """Blocks until one phrase has been heard and returns it as ``AudioData`` (with up to
        ``non_speaking_duration`` of quiet on both sides)."""
        kept, _ = self._listen_buffers(source, timeout, phrase_time_limit)
        return AudioData(b"".join(kept), source.sampling_rate, source.sampling_width)

    def _listen_buffers(self, source, timeout=None, phrase_time_limit=None):
        """The loop of ``listen``: (buffers of the phrase, bytes of trailing quiet that were read but not kept)."""
        _check_source(source, "listening")
        spb, pause_n, phrase_n, keep_n = self._buffer_counts(source)
        clock = 0.0
        while True:
            kept = collections.deque()
            chunk = b""
            while True:                                   # quiet: keep a short pre-roll, follow the noise floor
                clock += spb
                if timeout and clock > timeout:
                    raise WaitTimeoutError("listening timed out while waiting for phrase to start")
                chunk = source.stream.read(source.chunk)
                if not len(chunk):
                    break
                kept.append(chunk)
                if len(kept) > keep_n:
                    kept.popleft()
                energy = pcm_rms(chunk, source.sampling_width)
                if energy > self.energy_threshold:
                    break
                if self.dynamic_energy_threshold:
                    self._follow_energy(energy, spb)
            quiet = heard = 0
            started = clock
            while True:                                   # phrase: until the pause is long enough
                clock += spb
                if phrase_time_limit and clock - started > phrase_time_limit:
                    break
                chunk = source.stream.read(source.chunk)
                if not len(chunk):
                    break
                kept.append(chunk)
                heard += 1
                quiet = 0 if pcm_rms(chunk, source.sampling_width) > self.energy_threshold else quiet + 1
                if quiet > pause_n:
                    break
            if heard - quiet >= phrase_n or not len(chunk):
                break                                     # long enough, or the stream ended
        dropped = 0                                       # bytes of trailing quiet beyond what is kept
        for _ in range(quiet - keep_n):
            dropped += len(kept.pop())
        return kept, dropped

    # ------------------------------------------------------------------ one phrase, buffer by buffer
    def listen_stream(self, source, timeout=None, phrase_time_limit=None):
        """Generator over one phrase: yields ``(False, [pre-roll buffers])`` when speech starts, ``(False, buffer)``
        for every buffer of the phrase, and finally ``(True, buffer)`` (``(True, [])`` at the end of the source).  A
        too-short burst restarts the search without an ``is_last``.  Advancing it once more raises
        ``WrongUsageOfListen``: a finished listen must be replaced by a new generator."""
        _check_source(source, "listening")
        spb, pause_n, phrase_n, keep_n = self._buffer_counts(source)
        clock = 0.0
        chunk = []
        while self.stream:
            pre = []
            while self.stream:
                clock += spb
                if timeout and clock > timeout:
                    raise WaitTimeoutError("listening timed out while waiting for phrase to start")
                chunk = source.stream.read(source.chunk)
                if not len(chunk):
                    break
                pre.append(chunk)
                if len(pre) > keep_n:
                    del pre[0]
                if pcm_rms(chunk, source.sampling_width) > self.energy_threshold:
                    break
            if not self.stream:                           # stopped while searching: let the consumer thread run out
                yield False, []
            yield False, pre
            quiet = heard = 0
            started = clock
            while True:
                chunk = source.stream.read(source.chunk)
                if not len(chunk):
                    break
                clock += spb
                if phrase_time_limit and clock - started > phrase_time_limit:
                    break
                heard += 1
                quiet = 0 if pcm_rms(chunk, source.sampling_width) > self.energy_threshold else quiet + 1
                if quiet > pause_n:
                    break
                yield False, chunk
            if heard - quiet >= phrase_n or not len(chunk):
                break
        yield True, (chunk if len(chunk) else [])
        raise WrongUsageOfListen("Wrong usage of stream. Overwrite the listen generator with a new generator instance"
                                 "since this instance has completed a full listen.")

    # ------------------------------------------------------------------ long recordings (an addition)
    def segment_audio(self, samples, sampling_rate=16000, chunk_size=1024, phrase_time_limit=None):
        """Phrases of an in-memory recording as ``[(first_sample, end_sample)]``, found by the same ``listen`` loop a
        live source goes through (including its drifting energy threshold).  The reference's long-audio example
        (example_scripts/video_transcribe_simulation.py:93-143) re-implements this loop around ``recognize``."""
        from .audio.resources import ArraySource
        spans = []
        with ArraySource(samples, sampling_rate=sampling_rate, chunk_size=chunk_size) as src:
            total = len(np.asarray(samples))
            while src.stream.tell() < total:
                kept, dropped = self._listen_buffers(src, phrase_time_limit=phrase_time_limit)
                n = sum(len(b) for b in kept) // src.sampling_width
                end = min(src.stream.tell(), total) - dropped // src.sampling_width
                if n and self._is_phrase(kept, src):
                    spans.append((end - n, end))
        return spans

    def _is_phrase(self, kept, source):
        """``listen`` also returns when the source ends; such a tail counts only if it holds speech."""
        return any(pcm_rms(b, source.sampling_width) > self.energy_threshold for b in kept)

    def recognize_long(self, samples, sampling_rate=16000, max_batch=64, phrase_time_limit=None, show_all=False):
        """Long recording -> ``[(start_seconds, end_seconds, transcript)]``: segmented by ``segment_audio`` and then
        recognised as length-sorted batches (``recognize_batches``) instead of one utterance at a time."""
        a = np.asarray(samples)
        spans = self.segment_audio(a, sampling_rate=sampling_rate, phrase_time_limit=phrase_time_limit)
        order = sorted(range(len(spans)), key=lambda i: spans[i][0] - spans[i][1])          # longest first
        batches = [order[i:i + max_batch] for i in range(0, len(order), max_batch)]
        texts = [None] * len(spans)
        clips = [[a[spans[i][0]:spans[i][1]].astype(float) for i in batch] for batch in batches]
        for batch, out in zip(batches, self.recognize_batches(clips, show_all=show_all) if clips else []):
            for i, t in zip(batch, out):
                texts[i] = t
        return [(lo / float(sampling_rate), hi / float(sampling_rate), t) for (lo, hi), t in zip(spans, texts)]

    @staticmethod
    def get_audio_data(frames, source):
        """Byte buffers of a stream -> the float array the models take."""
        return AudioData(b"".join(frames), source.sampling_rate, source.sampling_width).get_array_data()

    # ------------------------------------------------------------------ background capture
    def listen_in_background(self, source):
        """Starts a daemon thread that keeps running ``listen_stream`` generators over ``source`` and queues
        ``(is_last, samples)``.  Returns ``(stopper, get_data)``; ``get_data`` raises ``NoDataInBuffer`` when the
        queue is empty."""
        assert all(hasattr(source, a) for a in ("chunk", "sampling_rate", "sampling_width")), \
            "Source must be an audio source"
        running = [True]
        queue = collections.deque()

        def capture():
            with source as s:
                while running[0]:
                    heard_anything = False
                    try:
                        for is_last, part in self.listen_stream(s):
                            part = part if isinstance(part, list) else [part]
                            heard_anything = heard_anything or len(part) > 0
                            queue.append((is_last, self.get_audio_data(part, source)))
                            if is_last:
                                break
                    except WaitTimeoutError:
                        pass
                    if self.stream and not heard_anything:
                        break      # a finite source (file, array) has run dry: end the capture instead of spinning

        worker = threading.Thread(target=capture, daemon=True)

        def stopper(wait_for_stop=True):
            running[0] = False
            if wait_for_stop:
                worker.join()

        def get_data():
            try:
                return queue.popleft()
            except IndexError:
                raise NoDataInBuffer

        get_data.capture_ended = lambda: not worker.is_alive()   # finite sources only: a microphone never ends
        worker.start()
        return stopper, get_data

    def enable_streaming(self):
        if self.stream:
            print("Streaming already enabled...")
        else:
            self.stream = True

    def disable_streaming(self):
        if self.stream:
            self.stream = False
            self.stream_thread_stopper(wait_for_stop=False)
        else:
            self.stream = True     # as the reference (Recognizer.py:430-431)

    # ------------------------------------------------------------------ consumers
    def streaming(self, source):
        """Generator of transcripts, one per phrase heard on ``source`` (after ``enable_streaming()``): the phrase
        is collected until the capture thread reports its end and then goes through ``recognize`` if it is longer
        than ``mininum_required_speaking_seconds``."""
        self.stream_thread_stopper, get_data = self.listen_in_background(source)
        capture_ended = getattr(get_data, "capture_ended", lambda: False)
        parts = []
        while self.stream:
            done = capture_ended()                        # read BEFORE polling: then an empty queue is final
            try:
                is_last, samples = get_data()
            except NoDataInBuffer:
                if done:
                    return                                # a finite source has been consumed (an addition)
                time.sleep(0.2)
                continue
            parts.append(samples)
            if not is_last:
                continue
            clip = np.concatenate(parts) if len(parts) > 1 else parts[0]
            parts = []
            if len(clip) > self.mininum_required_speaking_seconds * source.sampling_rate:
                yield self.recognize(clip)

    def real_time_streaming(self, source):
        """Generator of ``(is_last, text)`` while a phrase is being spoken (after ``enable_real_time_streaming``).

        Audio is handed to ``streaming_transcribe`` as soon as enough of it has arrived: the first pass of a phrase
        needs the look-ahead of the model plus the context of the first convolutions, later passes the look-ahead
        alone, and the pass that carries the end of the phrase goes through whatever its length.  A phrase that ends
        before its first pass is not transcribed.  Sample requirements as Recognizer.py:600-610 (10 ms hop, 20 ms
        window)."""
        per_10ms = int(source.sampling_rate / 100)
        lookahead_frames = (self.danspeech_recognizer.model.context - 1) * 2
        need_later = per_10ms * 2 + per_10ms * (lookahead_frames - 1)
        need_first = need_later + per_10ms * 15

        self.stream_thread_stopper, get_data = self.listen_in_background(source)
        capture_ended = getattr(get_data, "capture_ended", lambda: False)
        time.sleep(0.2)                                   # let the capture thread start
        pending, first_pass, ended = [], True, False
        got_some, misses = False, 0
        while self.stream:
            # drain the queue; stop draining at a phrase end, or once a run of data is followed by an empty queue
            while not ended:
                done = capture_ended()                    # read BEFORE polling: then an empty queue is final
                try:
                    ended, samples = get_data()
                    pending.append(samples)
                    got_some = True
                except NoDataInBuffer:
                    if got_some:
                        got_some, misses = False, 0
                        break
                    if done:
                        return                            # a finite source has been consumed (an addition)
                    if not pending:
                        time.sleep(0.4)
                    else:
                        misses += 1
                    if misses == 2:
                        misses = 0
                        time.sleep(0.3)
            have = sum(len(p) for p in pending)
            text = None
            if first_pass:
                if not ended and have >= need_first:
                    text = self._transcribe_pending(pending, is_last=False, is_first=True)
                    first_pass = False
            elif ended or have >= need_later:
                text = self._transcribe_pending(pending, is_last=ended, is_first=False)
            if text:
                yield ended, text
            if ended:
                # As in the reference, a phrase that ended before its first pass is not transcribed but its samples
                # stay queued in front of the next phrase (Recognizer.py:667-669 resets nothing); kept for parity.
                first_pass, ended = True, False

    def _transcribe_pending(self, pending, is_last, is_first):
        samples = np.concatenate(pending) if len(pending) > 1 else pending[0]
        del pending[:]
        return self.danspeech_recognizer.streaming_transcribe(samples, is_last=is_last, is_first=is_first)
