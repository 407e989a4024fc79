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


def test_kenlm_binary_is_recognised_and_refused(lib, tmp_path):
    """A .klm (KenLM binary, lm/binary_format.cc header) must fail with what it is, not with an ARPA parse error."""
    import struct
    from danspeech_b200.utils import synthetic as syn
    magic = b"mmap lm http://kheafield.com/code format version 5\n\0"
    sanity = magic.ljust(56, b"\0") + struct.pack("<fffIIQ", 0.0, 1.0, -0.5, 1, 0xFFFFFFFF, 1) + b"\0" * 4
    assert len(sanity) == 88
    params = struct.pack("<B3xfiB3xI", 3, 1.5, 2, 1, 1)          # order 3, trie, with vocabulary
    p = tmp_path / "dsl_3gram.klm"
    p.write_bytes(sanity + params + struct.pack("<QQQ", 12345, 99, 7) + b"\0" * 64)
    labels = "\0".join(syn.LABELS).encode("utf-8") + b"\0"
    h = ctypes.c_void_p()
    rc = lib.dsb_beam_create(labels, len(syn.LABELS), str(p).encode(), 1.3, 0.2, 40, 1.0, 64, 0, 0, ctypes.byref(h))
    msg = lib.dsb_last_error().decode()
    assert rc != 0 and "KenLM binary" in msg and "trie" in msg and "order 3" in msg and "12345 unigrams" in msg


def test_public_header_is_plain_c_and_cxx(tmp_path):
    """include/danspeech_b200.h is what a foreign-function binding reads: it must compile on its own as C99 and as
    C++11 (no torch / CUDA types in the signatures)."""
    import shutil
    import subprocess
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    src = tmp_path / "t.c"
    src.write_text('#include "danspeech_b200.h"\nint main(void) { return 0; }\n')
    if shutil.which("gcc"):
        subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", inc, "-fsyntax-only", str(src)],
                       check=True)
    if shutil.which("g++"):
        subprocess.run(["g++", "-std=c++11", "-Wall", "-Wextra", "-Werror", "-I", inc, "-fsyntax-only", "-x", "c++", str(src)],
                       check=True)


def test_integration_doc_names_only_symbols_the_header_declares():
    """INTEGRATION.md is the binding recipe: every ``dsb_*`` function it mentions must be declared in the header,
    and its Python snippets must at least parse."""
    import ast
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    doc = open(os.path.join(root, "INTEGRATION.md"), encoding="utf-8").read()
    header = open(os.path.join(root, "include", "danspeech_b200.h"), encoding="utf-8").read()
    declared = set(re.findall(r"\b(dsb_[a-z0-9_]+)\s*\(", header))
    types_and_consts = set(re.findall(r"\b(dsb_[a-z0-9_]+)\b", header)) - declared
    mentioned = set(re.findall(r"\b(dsb_[a-z0-9_]+)\b", doc))
    wildcard = {m for m in mentioned if m.endswith("_")}          # e.g. "dsb_profile_enable/reset/read" prefixes
    unknown = {m for m in mentioned - wildcard if m not in declared and m not in types_and_consts}
    assert not unknown, sorted(unknown)
    assert len(mentioned & declared) >= 15
    for block in re.findall(r"```python\n(.*?)```", doc, flags=re.S):
        if "# reference" in block and "# danspeech_b200" in block:
            continue                                               # the two-column import comparison is not code
        ast.parse(block)
