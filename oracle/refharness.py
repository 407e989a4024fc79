"""Oracle (test infrastructure): import the UNMODIFIED reference from /root/reference.

Only usable in the build container (the GPU box has no /root/reference).  Used by
tests/golden/gen_golden.py to generate golden vectors and by the ``not gpu``
tests that re-pin the oracle when the reference tree is present.

The reference's third-party imports that are absent from this image are stubbed
*before* import:
  * ``Levenshtein``  (decoder.py:19; only used by wer/cer, off the path)
  * ``wget``         (utils/data_utils.py:4; downloads, off the path)
  * ``librosa``      (parsers.py:7) -> oracle.spectrogram.stft / magphase
  * ``scipy.signal.{hamming,hann,blackman,bartlett}`` (parsers.py:9-10) moved to
    scipy.signal.windows in modern scipy.
"""
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "danspeech"))


def import_reference():
    """Returns the reference ``danspeech`` package (imported once)."""
    if "danspeech" in sys.modules and getattr(sys.modules["danspeech"], "__file__", "").startswith(REFERENCE_ROOT):
        return sys.modules["danspeech"]
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    from . import spectrogram as _sp
    import scipy.signal
    import scipy.signal.windows as _w

    for name in ("hamming", "hann", "blackman", "bartlett"):
        if not hasattr(scipy.signal, name):
            setattr(scipy.signal, name, getattr(_w, name))

    lev = types.ModuleType("Levenshtein")
    lev.distance = lambda a, b: 0
    sys.modules.setdefault("Levenshtein", lev)
    sys.modules.setdefault("wget", types.ModuleType("wget"))
    if "librosa" not in sys.modules:
        lib = types.ModuleType("librosa")
        lib.stft = lambda y, n_fft=2048, hop_length=None, win_length=None, window="hann", center=True: _sp.stft(
            y, n_fft=n_fft, hop_length=hop_length, win_length=win_length, window=window, center=center)
        lib.magphase = _sp.magphase
        sys.modules["librosa"] = lib
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import danspeech  # noqa: E402  (the reference)
    return danspeech
