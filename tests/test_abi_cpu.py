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
    assert lib.dsb_abi_version() == 2
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
    """A .klm variant the reader does not handle (here: a trie model) must fail with what it is, not with an ARPA parse
    error (the probing model is read: tests/test_lm_cpu.py)."""
    import struct
    from danspeech_b200.utils import synthetic as syn
    magic = b"mmap lm http://kheafield.com/code format version 5\n\0"
    sanity = magic.ljust(56, b"\0") + struct.pack("<fffII4xQ", 0.0, 1.0, -0.5, 1, 0xFFFFFFFF, 1)
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


def test_reference_citations_resolve():
    """Every ``file.py:line(-line)`` citation in the header, the docs and the sources points into an existing file of
    the reference checkout (build container only) with that many lines."""
    import glob
    ref = "/root/reference"
    if not os.path.isdir(os.path.join(ref, "danspeech")):
        pytest.skip("reference tree only exists in the build container")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lines = {}
    for d, _, files in os.walk(ref):
        for f in files:
            if f.endswith((".py", ".rst", ".txt", ".json")):
                with open(os.path.join(d, f), encoding="utf-8", errors="ignore") as fh:
                    lines.setdefault(f, []).append((os.path.join(d, f), sum(1 for _ in fh)))
    ours = ["include/danspeech_b200.h", "DESIGN.md", "INTEGRATION.md", "README.md", "bench.py"]
    for pat in ("danspeech_b200/**/*.py", "danspeech_b200/csrc/*", "oracle/*.py", "oracle/*.cpp", "tests/*.py", "tests/golden/*.py"):
        ours += [os.path.relpath(p, root) for p in glob.glob(os.path.join(root, pat), recursive=True)]
    cite = re.compile(r"([A-Za-z_][\w/\.]*\.(?:py|rst|txt|json)):(\d+)(?:-(\d+))?")
    checked, bad = 0, []
    for rel in ours:
        path = os.path.join(root, rel)
        if not os.path.isfile(path):
            continue
        text = open(path, encoding="utf-8", errors="ignore").read()
        for m in cite.finditer(text):
            name, a, b = m.group(1), int(m.group(2)), int(m.group(3) or m.group(2))
            cands = lines.get(os.path.basename(name))
            if not cands:
                if not (os.path.exists(os.path.join(root, name)) or os.path.exists(os.path.join(root, "danspeech_b200", name))
                        or os.path.exists(os.path.join(root, "tests", os.path.basename(name)))):
                    bad.append((rel, m.group(0), "no such file"))
                continue
            cands = [c for c in cands if c[0].endswith(name)] or cands
            checked += 1
            if not any(a <= b <= n for _, n in cands):
                bad.append((rel, m.group(0), "line range"))
    assert not bad, bad[:10]
    assert checked > 150


def test_tuning_knobs_are_validated(lib):
    """dsb_tune_set / dsb_tune_get: known keys only, values inside their ranges, defaults = production settings."""
    assert lib.dsb_tune_get(b"rnn_in_flight") == 3 and lib.dsb_tune_get(b"rnn_ksplit") == 0
    assert lib.dsb_tune_get(b"rnn_max_slots") == 0 and lib.dsb_tune_get(b"rnn_producers") == 1
    # the CTA-pair recurrence (rnn_pair.cu) with batch-minor pre-activations is the production path for >= 2 groups
    assert lib.dsb_tune_get(b"rnn_pair") == 1 and lib.dsb_tune_get(b"rnn_batch_minor") == 1
    assert lib.dsb_tune_get(b"rnn_pair_in_flight") == 2
    assert lib.dsb_tune_set(b"rnn_pair_in_flight", 4) != 0 and b"outside" in lib.dsb_last_error()
    assert lib.dsb_tune_get(b"no_such_knob") == -1
    assert lib.dsb_tune_set(b"no_such_knob", 1) != 0 and b"unknown key" in lib.dsb_last_error()
    assert lib.dsb_tune_set(b"rnn_in_flight", 4) != 0 and b"outside" in lib.dsb_last_error()
    assert lib.dsb_tune_set(b"rnn_in_flight", 2) == 0 and lib.dsb_tune_get(b"rnn_in_flight") == 2
    assert lib.dsb_tune_set(b"rnn_in_flight", 3) == 0
    from danspeech_b200 import _native as N
    prev = N.tune(rnn_max_slots=1)
    assert prev == {"rnn_max_slots": 0} and lib.dsb_tune_get(b"rnn_max_slots") == 1
    N.tune(**prev)
    assert lib.dsb_tune_get(b"rnn_max_slots") == 0


def test_forward_status_needs_a_model(lib):
    assert lib.dsb_forward_status(None) != 0
