"""CPU oracle for the DanSpeech inference hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and only as the checker.  The product package
(``danspeech_b200``) never imports this package and fails loudly when its CUDA
library is missing.

Pinning status ("Oracle"):
  * model arithmetic (conv/BN/GRU/LSTM/RNN/lookahead/fc/softmax), greedy CTC,
    get_seq_lens, streaming forward: PINNED -- checked against the reference's
    own modules imported unmodified in the build container
    (tests/golden/gen_golden.py writes the fixtures under tests/golden/).
  * spectrogram: librosa is a third-party dependency that is absent from
    /root/reference and from this image (unpinned in requirements.txt:7).  The
    restatement in ``oracle/spectrogram.py`` follows librosa 0.7 semantics and
    is driven through the reference's own ``parsers.py`` call sites, but the
    STFT itself is "parity unpinned".
  * beam search: ``ctcdecode`` (parlance/ctcdecode, git master, unpinned;
    docs_source/installation.rst:25-30) + KenLM are absent.  The restatement in
    ``oracle/ctc_beam.cpp`` follows the published algorithm; "parity unpinned".
"""
