"""CPU: the C-ABI library loads, exports every symbol include/danspeech_b200.h declares, and its
argument validation works without a GPU (no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from danspeech_b200 import _native as N
    return N.lib()


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "danspeech_b200.h"), encoding="utf-8").read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dsb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(lib):
    from danspeech_b200 import _native as N
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), "library does not export %s" % name
    assert set(declared) == set(N.EXPORTED_SYMBOLS), "ctypes table and header disagree"


def test_abi_version_and_errors(lib):
    from danspeech_b200 import _native as N
    assert lib.dsb_abi_version() == 1
    assert lib.dsb_spectrogram_num_frames(240000) == 1501
    assert lib.dsb_spectrogram_num_frames(66944) == 419
    rc = lib.dsb_greedy_decode(None, None, None, 1, 1, 33, 0, None, None, None, None)
    assert rc == -1 and b"null" in lib.dsb_last_error()
    h = ctypes.c_void_p()
    bad = N.ModelDesc(conv_layers=4, rnn_layers=5, rnn_hidden_size=400, rnn_type=0, bidirectional=1, context=20,
                      num_classes=33, streaming=0)
    assert lib.dsb_model_create(bad, h) == -1
    ok = N.ModelDesc(conv_layers=3, rnn_layers=9, rnn_hidden_size=1200, rnn_type=0, bidirectional=1, context=20,
                     num_classes=33, streaming=0)
    assert lib.dsb_model_create(ok, h) == 0
    assert lib.dsb_model_out_frames(h, 1501) == 751 and lib.dsb_model_out_frames(h, 419) == 210
    assert lib.dsb_forward_workspace_bytes(h, 64, 1501) > 0
    lib.dsb_model_destroy(h)


def test_product_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from danspeech_b200 import _native as N
    from danspeech_b200 import Recognizer
    with pytest.raises(N.NativeError):
        Recognizer()
    with pytest.raises(N.NativeError):
        Recognizer(with_gpu=False)
