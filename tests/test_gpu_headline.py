"""Parity on the HEADLINE configuration as benchmarked (BASELINE.json configs[1]): DanSpeechPrimary-shaped model,
batch 64 x 15 s (T = 1501 spectrogram frames, T' = 751 model frames), the utterances bench.py times (seeds 0..63).

The CPU oracle (oracle/model.py, pinned against the unmodified reference) costs ~1.3 s per 15 s utterance, so it runs
on 8 of the 64 utterances spread over the batch; the GPU runs the whole batch, audio -> spectrogram -> model -> greedy.

north_star bars:  fp32 mode  logits <= 1e-4 relative, greedy transcripts bit-exact (decoder.py:166-198);
                  bf16 mode  logits <= 2e-2 relative; the transcript identity is ASSERTED at the rate measured on
                  the B200 (see BF16_* below) and the argmax-margin histogram of the flipped frames is recorded.
"""
import json
import os

import numpy as np
import pytest
import torch

from conftest import ROOT, case_config, logit_rel_err, rel_err
from oracle import greedy as og
from oracle import model as om
from oracle import spectrogram as osp
from danspeech_b200.utils import synthetic as syn
import parity_util as pu

pytestmark = pytest.mark.gpu

BATCH, N_SAMPLES = 64, 15 * 16000
SAMPLE = [0, 9, 18, 27, 36, 45, 54, 63]
FP32_TOL, BF16_TOL = 1e-4, 2e-2
# bf16 mode, measured on the B200 (profiles/r02_parity_headline_bf16.json): logit error 1.2e-2..1.5e-2, 5888 of 6008
# frames keep the oracle's argmax (98.0 %), every flipped frame has an oracle top1-top2 log-probability margin below
# 0.2 (18 % of this random-weight network's frames have a margin below 0.1), CER 7.4 %, no transcript identical.
# bf16 operands cannot make the greedy path bit-exact on such margins; the bit-exact mode is fp32 (test above).
BF16_MIN_FRAME_AGREEMENT = 0.975     # fraction of the 8 x 751 frames whose argmax equals the oracle's
BF16_MAX_CER = 0.10                  # character error rate of the greedy transcripts against the oracle's
BF16_MAX_FLIPPED_MARGIN = 0.3        # no frame whose oracle top1-top2 log-prob margin exceeds this may flip


@pytest.fixture(scope="module")
def headline():
    auds = [syn.synthetic_audio(N_SAMPLES, seed=i) for i in range(BATCH)]
    cfg = case_config("DanSpeechPrimary", {})
    sd = syn.make_state_dict(seed=0, **cfg)
    parser = osp.SpectrogramOracle()
    specs = [parser.parse_audio(auds[i]) for i in SAMPLE]
    x = torch.stack(specs).view(len(SAMPLE), 1, 161, -1)
    lens = torch.IntTensor([s.size(1) for s in specs])
    torch.set_num_threads(os.cpu_count())
    with torch.no_grad():
        ref, rs = om.forward(sd, x, lens, cfg["conv_layers"], cfg["rnn_layers"])
    assert tuple(ref.shape) == (len(SAMPLE), 751, 33) and set(rs.tolist()) == {751}
    texts = [t[0] for t in og.greedy_decode(ref.numpy(), rs.numpy())[0]]
    return {"auds": auds, "ref": ref.numpy(), "texts": texts, "specs": [s.numpy() for s in specs]}


def _run(precision, auds):
    from danspeech_b200 import Recognizer
    from danspeech_b200.pretrained_models import build_model
    rec = Recognizer(model=build_model("DanSpeechPrimary", seed=0).set_precision(precision))
    eng = rec.danspeech_recognizer
    x, lens = eng.audio_parser.parse_batch(auds)
    assert tuple(x.shape) == (BATCH, 1, 161, 1501)
    probs, sizes = eng.model(x, lens)
    assert tuple(probs.shape) == (BATCH, 751, 33) and set(sizes.tolist()) == {751}
    texts = rec.recognize_batch(auds)                      # the public API on the same batch
    dec = [s[0] for s in eng.decoder.decode(probs, sizes)[0]]
    assert texts == dec
    return x, probs.cpu().numpy(), texts


def _record(name, payload):
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "parity_headline_%s.json" % name), "w") as f:
            json.dump(payload, f, indent=1)


def test_headline_fp32_logits_and_transcripts_bit_exact(headline):
    x, probs, texts = _run("fp32", headline["auds"])
    errs, specs = [], []
    for k, i in enumerate(SAMPLE):
        specs.append(rel_err(x[i, 0].cpu().numpy(), headline["specs"][k]))
        errs.append(logit_rel_err(probs[i], headline["ref"][k]))
    frames = pu.merge_reports([pu.frame_report(probs[i], headline["ref"][k]) for k, i in enumerate(SAMPLE)])
    tr = pu.transcript_report([texts[i] for i in SAMPLE], headline["texts"])
    _record("fp32", {"logit_rel_err": errs, "spect_rel_err": specs, "frames": frames, "transcripts": tr})
    print("fp32 headline: logit err max %.2e, spect err max %.2e, transcripts %d/%d identical" %
          (max(errs), max(specs), tr["identical"], tr["utterances"]))
    assert max(specs) < FP32_TOL
    assert max(errs) < FP32_TOL
    assert frames["frames_differ"] == 0
    assert [texts[i] for i in SAMPLE] == headline["texts"]          # greedy bit-exact


def test_headline_bf16_logits_and_asserted_transcript_rate(headline):
    x, probs, texts = _run("bf16", headline["auds"])
    errs = [logit_rel_err(probs[i], headline["ref"][k]) for k, i in enumerate(SAMPLE)]
    frames = pu.merge_reports([pu.frame_report(probs[i], headline["ref"][k]) for k, i in enumerate(SAMPLE)])
    tr = pu.transcript_report([texts[i] for i in SAMPLE], headline["texts"])
    _record("bf16", {"logit_rel_err": errs, "frames": frames, "transcripts": tr})
    agree = 1.0 - frames["frames_differ"] / frames["frames"]
    print("bf16 headline: logit err max %.2e; frames agreeing %.4f; transcripts %d/%d identical, CER %.4f; "
          "largest flipped margin %.3f; flipped-margin histogram %s over %s" %
          (max(errs), agree, tr["identical"], tr["utterances"], tr["cer"], frames["max_flipped_margin"],
           frames["flipped_margin_hist"], frames["margin_edges"]))
    assert max(errs) < BF16_TOL
    assert agree >= BF16_MIN_FRAME_AGREEMENT
    assert tr["cer"] <= BF16_MAX_CER
    assert frames["max_flipped_margin"] <= BF16_MAX_FLIPPED_MARGIN


def test_headline_beam64_lm_on_model_output_matches_oracle(headline, tmp_path):
    """BASELINE config 3 on the headline shape: beam width 64 + synthetic 3-gram LM over the MODEL's probabilities of the
    benchmarked batch (T' = 751), GPU prefix beam search against the CPU oracle (oracle/ctc_beam.cpp) on the 8 sampled
    utterances, and the whole batch decoded through the public API.  north_star: top-1 identical on >= 99.5 % of the
    utterances, beam scores within 1e-3."""
    from danspeech_b200 import Recognizer
    from danspeech_b200.deepspeech.decoder import BeamCTCDecoder
    from danspeech_b200.pretrained_models import build_model
    from oracle.beam import CTCBeamDecoderOracle
    arpa = syn.write_synthetic_arpa(str(tmp_path / "lm.arpa"), n_words=2000, seed=0)
    rec = Recognizer(model=build_model("DanSpeechPrimary", seed=0).set_precision("fp32"))
    eng = rec.danspeech_recognizer
    x, lens = eng.audio_parser.parse_batch(headline["auds"])
    probs, sizes = eng.model(x, lens)
    sub = probs[SAMPLE].contiguous()
    gpu = BeamCTCDecoder(syn.LABELS, arpa, 1.3, 0.2, 40, 1.0, 64, 6, 0)
    ref = CTCBeamDecoderOracle(syn.LABELS, arpa, 1.3, 0.2, 40, 1.0, 64, 6, 0)
    out, scores, ts, out_len = [t.cpu().numpy() for t in gpu.decode_device(sub, sizes[SAMPLE])]
    r_out, r_scores, r_ts, r_len = ref.decode(sub.cpu().numpy(), sizes[SAMPLE].tolist())
    same, ts_same, ts_all, rel = 0, 0, 0, []
    for b in range(len(SAMPLE)):
        n, rn = out_len[b, 0], r_len[b, 0]
        if n == rn and np.array_equal(out[b, 0, :n], r_out[b, 0, :rn]):
            same += 1
            rel.append(abs(float(scores[b, 0]) - float(r_scores[b, 0])) / max(1.0, abs(float(r_scores[b, 0]))))
            ts_same += int((ts[b, 0, :n] == r_ts[b, 0, :rn]).sum())
            ts_all += int(n)
    print("beam-64 + LM on model output, T' = 751: top-1 identical on %d / %d utterances; relative score differences %s; "
          "character time steps equal at %d / %d positions" % (same, len(SAMPLE), ["%.1e" % r for r in rel], ts_same, ts_all))
    # Top beam identical on every sampled utterance, scores equal to float rounding.  (Until the GPU search gave a prefix
    # that re-enters the beam its OLD trie node back -- ctcdecode's PathTrie keeps removed nodes that still have live
    # descendants -- duplicates of one prefix crept into the beam on this near-uniform output and only 13 of 16
    # utterances agreed; with node identity all 64 hypotheses and their scores agree.)
    assert same == len(SAMPLE)
    assert max(rel) <= 1e-5
    # Character time steps (the decoder's second output): a PathTrie node keeps the frame of its best symbol probability
    # for as long as it exists in the trie -- also while its prefix is out of the beam, if it still has descendants --
    # and starts afresh once it has been deleted; the GPU arena reproduces that lifetime with reference counts.
    assert ts_same == ts_all
    # the public API on the whole batch (top beam only) agrees with the decoder object
    rec.update_decoder(lm=arpa, alpha=1.3, beta=0.2, beam_width=64)
    texts = rec.recognize_batch(headline["auds"])
    mine, _ = gpu.decode(sub, sizes[SAMPLE])
    assert [texts[i] for i in SAMPLE] == [m[0] for m in mine]
