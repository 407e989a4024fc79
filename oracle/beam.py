"""Oracle (test infrastructure): ctypes wrapper of oracle/ctc_beam.cpp with the call shape of
``ctcdecode.CTCBeamDecoder`` (constructed at decoder.py:99-100, called at decoder.py:140).
PARITY UNPINNED: ctcdecode/KenLM are absent third-party dependencies (see ctc_beam.cpp header)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libbeam_oracle.so")
_lib = None


def _load():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "ctc_beam.cpp")
        if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
            subprocess.run(["make", "-s", "-C", _HERE], check=True)
        L = ctypes.CDLL(_LIB)
        L.oracle_beam_create.restype = ctypes.c_void_p
        L.oracle_beam_create.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_double,
                                         ctypes.c_double, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int]
        L.oracle_beam_destroy.argtypes = [ctypes.c_void_p]
        L.oracle_beam_is_char_based.argtypes = [ctypes.c_void_p]
        L.oracle_beam_lm_order.argtypes = [ctypes.c_void_p]
        L.oracle_lm_cond_log_prob.restype = ctypes.c_double
        L.oracle_lm_cond_log_prob.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int]
        L.oracle_beam_decode.argtypes = [ctypes.c_void_p] + [ctypes.c_void_p] * 2 + [ctypes.c_int] * 3 + \
            [ctypes.c_void_p] * 4 + [ctypes.c_int]
        _lib = L
    return _lib


class CTCBeamDecoderOracle(object):
    def __init__(self, labels, model_path=None, alpha=0, beta=0, cutoff_top_n=40, cutoff_prob=1.0, beam_width=100,
                 num_processes=4, blank_id=0):
        L = _load()
        self.labels = list(labels)
        self.beam_width = beam_width
        self.num_processes = num_processes
        blob = b"\0".join(c.encode("utf-8") for c in self.labels) + b"\0"
        self._h = L.oracle_beam_create(blob, len(self.labels), model_path.encode() if model_path else None,
                                       float(alpha), float(beta), int(cutoff_top_n), float(cutoff_prob),
                                       int(beam_width), int(blank_id))
        if not self._h:
            raise IOError("could not load language model %r" % model_path)

    def __del__(self):
        if getattr(self, "_h", None):
            _load().oracle_beam_destroy(self._h)
            self._h = None

    @property
    def is_char_based(self):
        return _load().oracle_beam_is_char_based(self._h)

    @property
    def lm_order(self):
        return _load().oracle_beam_lm_order(self._h)

    def lm_cond_log_prob(self, words):
        blob = b"\0".join(w.encode("utf-8") for w in words) + b"\0"
        return _load().oracle_lm_cond_log_prob(self._h, blob, len(words))

    def decode(self, probs, seq_lens=None):
        """probs [B,T,C] -> (output [B,beam,T] int32, scores [B,beam] f32, timesteps [B,beam,T], out_seq_len [B,beam])."""
        probs = np.ascontiguousarray(np.asarray(probs, dtype=np.float32))
        B, T, C = probs.shape
        lens = np.ascontiguousarray(np.asarray(seq_lens if seq_lens is not None else [T] * B, dtype=np.int32))
        W = self.beam_width
        out = np.zeros((B, W, T), np.int32)
        ts = np.zeros((B, W, T), np.int32)
        scores = np.zeros((B, W), np.float32)
        out_len = np.zeros((B, W), np.int32)
        _load().oracle_beam_decode(self._h, probs.ctypes.data, lens.ctypes.data, B, T, C, out.ctypes.data,
                                   ts.ctypes.data, scores.ctypes.data, out_len.ctypes.data, int(self.num_processes))
        return out, scores, ts, out_len

    def decode_strings(self, probs, seq_lens=None):
        out, scores, ts, out_len = self.decode(probs, seq_lens)
        strings = [["".join(self.labels[i] for i in out[b, p, : out_len[b, p]]) for p in range(out.shape[1])]
                   for b in range(out.shape[0])]
        return strings, scores
