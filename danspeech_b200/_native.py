"""ctypes binding of libdanspeech_b200.so (the C ABI in include/danspeech_b200.h).

The product path has NO fallback: if the library is missing or a call fails, this raises.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_uint64, c_void_p

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libdanspeech_b200.so")
_lib = None


class NativeError(RuntimeError):
    """A danspeech_b200 C-ABI call failed."""


class ModelDesc(ctypes.Structure):
    _fields_ = [(n, c_int32) for n in ("conv_layers", "rnn_layers", "rnn_hidden_size", "rnn_type", "bidirectional",
                                       "context", "num_classes", "streaming")]


RNN_TYPES = {"gru": 0, "lstm": 1, "rnn": 2}
PRECISIONS = {"fp32": 0, "bf16": 1}

_SIGS = {
    "dsb_last_error": (c_char_p, []),
    "dsb_abi_version": (c_int, []),
    "dsb_kernel_launch_count": (c_uint64, []),
    "dsb_profile_enable": (None, [c_int]),
    "dsb_profile_reset": (None, []),
    "dsb_profile_read": (c_int, [c_int, POINTER(c_double), POINTER(c_int)]),
    "dsb_tune_set": (c_int, [c_char_p, c_int]),
    "dsb_tune_get": (c_int, [c_char_p]),
    "dsb_spectrogram_num_frames": (c_int, [c_int]),
    "dsb_spectrogram_partials": (c_int, [c_int]),
    "dsb_spectrogram_f32": (c_int, [c_void_p, c_int64, c_void_p, c_int, c_int, c_void_p, c_int64, c_void_p, c_void_p,
                                    c_int, c_void_p]),
    "dsb_spectrogram_s16": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_int, c_int, c_void_p, c_int64, c_void_p,
                                    c_void_p, c_int, c_void_p]),
    "dsb_spectrogram_stream_f32": (c_int, [c_void_p, c_int64, c_void_p, c_int, c_int, c_void_p, c_int64, c_void_p,
                                           c_void_p, c_int, c_void_p]),
    "dsb_spectrogram_stream_normalize": (c_int, [c_void_p, c_int64, c_void_p, c_int, c_void_p, c_void_p]),
    "dsb_spectrogram_stream_running_stats": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_double, c_double, c_double,
                                                     c_void_p]),
    "dsb_model_create": (c_int, [POINTER(ModelDesc), POINTER(c_void_p)]),
    "dsb_model_destroy": (None, [c_void_p]),
    "dsb_model_set_tensor": (c_int, [c_void_p, c_char_p, c_void_p, c_int64]),
    "dsb_model_finalize": (c_int, [c_void_p, c_int, c_void_p]),
    "dsb_model_precision": (c_int, [c_void_p]),
    "dsb_model_out_frames": (c_int, [c_void_p, c_int]),
    "dsb_forward_workspace_bytes": (c_size_t, [c_void_p, c_int, c_int]),
    "dsb_forward": (c_int, [c_void_p, c_void_p, POINTER(c_int32), c_int, c_int, c_void_p, POINTER(c_int32), c_void_p,
                            c_void_p, c_size_t, c_void_p]),
    "dsb_forward_status": (c_int, [c_void_p]),
    "dsb_gemm_bf16": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int, c_int, c_int,
                              c_void_p]),
    "dsb_stream_state_create": (c_int, [c_void_p, c_int, c_int, POINTER(c_void_p)]),
    "dsb_stream_state_destroy": (None, [c_void_p]),
    "dsb_stream_max_out_frames": (c_int, [c_void_p, c_int]),
    "dsb_streaming_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, POINTER(c_int32),
                                      c_void_p]),
    "dsb_greedy_decode": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                  c_void_p, c_void_p]),
    "dsb_beam_create": (c_int, [c_char_p, c_int, c_char_p, c_float, c_float, c_int, c_float, c_int, c_int, c_int,
                                POINTER(c_void_p)]),
    "dsb_beam_destroy": (None, [c_void_p]),
    "dsb_beam_workspace_bytes": (c_size_t, [c_void_p, c_int, c_int]),
    "dsb_beam_decode": (c_int, [c_void_p, c_void_p, POINTER(c_int32), c_int, c_int, c_int, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "dsb_lm_inspect": (c_int, [c_char_p, POINTER(c_int), POINTER(c_int64), POINTER(c_int64), POINTER(c_uint64)]),
    "dsb_beam_lm_order": (c_int, [c_void_p]),
    "dsb_beam_lm_is_char_based": (c_int, [c_void_p]),
    "dsb_beam_lm_num_ngrams": (c_int64, [c_void_p]),
    "dsb_vad_state_create": (c_int, [c_int, c_int, c_int, POINTER(c_void_p)]),
    "dsb_vad_state_destroy": (None, [c_void_p]),
    "dsb_vad_reset": (c_int, [c_void_p, c_void_p]),
    "dsb_vad_push_s16": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGS)


def lib_path():
    return _LIB_PATH


def lib():
    """Loads the native library once; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise NativeError(
                "danspeech_b200 native library not found at %s -- build it with "
                "`python -c 'import __graft_entry__ as g; g.build()'` (there is no CPU fallback)" % _LIB_PATH)
        L = ctypes.CDLL(_LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(code, what=""):
    if code != 0:
        msg = lib().dsb_last_error()
        raise NativeError("%s failed (%d): %s" % (what or "native call", code, msg.decode("utf-8", "replace") if msg else ""))


def launch_count():
    return int(lib().dsb_kernel_launch_count())


STAGES = ("spectrogram", "conv", "rnn_input_proj", "rnn_recurrence", "tail", "greedy", "beam", "rnn_combine")


def profile_read():
    """{stage: (total_ms, spans)} since the last dsb_profile_reset()."""
    out = {}
    for i, name in enumerate(STAGES):
        ms, n = c_double(0), c_int(0)
        check(lib().dsb_profile_read(i, ms, n), "dsb_profile_read")
        out[name] = (ms.value, n.value)
    return out


def tune(**knobs):
    """Sets diagnostic tuning knobs (dsb_tune_set), e.g. tune(rnn_in_flight=1); returns the previous values."""
    prev = {}
    for k, v in knobs.items():
        prev[k] = int(lib().dsb_tune_get(k.encode()))
        check(lib().dsb_tune_set(k.encode(), int(v)), "dsb_tune_set(%s)" % k)
    return prev


def ptr(t):
    """Raw device/host pointer of a torch tensor (or None)."""
    return None if t is None else c_void_p(t.data_ptr())


def i32_array(values):
    arr = (c_int32 * len(values))(*[int(v) for v in values])
    return arr


def current_stream():
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise NativeError("danspeech_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
