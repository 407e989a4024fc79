"""Diagnostic: beam-64 + LM on the model output of the headline shape (T' = 751), GPU search vs CPU oracle, with the
rank and score of every GPU top-1 hypothesis inside the oracle's beam."""
import os
import sys
import tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import __graft_entry__ as g  # noqa: E402
g.build()
from danspeech_b200 import Recognizer  # noqa: E402
from danspeech_b200.deepspeech.decoder import BeamCTCDecoder  # noqa: E402
from danspeech_b200.pretrained_models import build_model  # noqa: E402
from danspeech_b200.utils import synthetic as syn  # noqa: E402
from oracle.beam import CTCBeamDecoderOracle  # noqa: E402

N = int(os.environ.get("UTTS", "16"))
auds = [syn.synthetic_audio(15 * 16000, seed=i) for i in range(N)]
arpa = syn.write_synthetic_arpa(os.path.join(tempfile.mkdtemp(), "lm.arpa"), n_words=2000, seed=0)
rec = Recognizer(model=build_model("DanSpeechPrimary", seed=0).set_precision("fp32"))
eng = rec.danspeech_recognizer
x, lens = eng.audio_parser.parse_batch(auds)
probs, sizes = eng.model(x, lens)
gpu = BeamCTCDecoder(syn.LABELS, arpa, 1.3, 0.2, 40, 1.0, 64, 6, 0)
ref = CTCBeamDecoderOracle(syn.LABELS, arpa, 1.3, 0.2, 40, 1.0, 64, 6, 0)
out, scores, ts, out_len = [t.cpu().numpy() for t in gpu.decode_device(probs, sizes)]
r_out, r_scores, r_ts, r_len = ref.decode(probs.cpu().numpy(), sizes.tolist())
for b in range(N):
    mine = tuple(out[b, 0, :out_len[b, 0]].tolist())
    rank = -1
    for j in range(64):
        if tuple(r_out[b, j, :r_len[b, j]].tolist()) == mine:
            rank = j
            break
    # how many of the GPU's 64 hypotheses are in the oracle's beam
    rset = {tuple(r_out[b, j, :r_len[b, j]].tolist()) for j in range(64)}
    common = sum(tuple(out[b, j, :out_len[b, j]].tolist()) in rset for j in range(64))
    print("utt %2d: gpu top-1 is oracle rank %2d | scores gpu %.4f oracle top-1 %.4f top-2 %.4f%s | %d/64 hypotheses in common"
          % (b, rank, scores[b, 0], r_scores[b, 0], r_scores[b, 1],
             (" oracle[rank] %.4f" % r_scores[b, rank]) if rank > 0 else "", common))
