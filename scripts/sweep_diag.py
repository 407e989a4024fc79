"""Diagnostic: the config-5 sweep (256 ragged utterances, one GPU) under different recurrence / merge settings, with the
library's stage timers."""
import json
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import __graft_entry__ as g  # noqa: E402
g.build()
from danspeech_b200 import Recognizer, sharding, _native as N  # noqa: E402
from danspeech_b200.pretrained_models import build_model  # noqa: E402
from danspeech_b200.utils import synthetic as syn  # noqa: E402

lens = np.random.default_rng(1234).integers(5 * 16000, 30 * 16000, size=256)
recs = {i: syn.synthetic_audio(int(lens[i]), seed=1234 + i) for i in range(256)}
batches = sharding.make_batches(list(range(256)), lens.tolist(), max_batch=64)
rec = Recognizer(model=build_model("DanSpeechPrimary", seed=0).set_precision("bf16"))
L = N.lib()
for merge, pair, bm in ((3, 0, 0), (4, 0, 0), (4, 1, 0), (4, 1, 1), (2, 1, 1)):
    N.tune(rnn_pair=pair, rnn_batch_minor=bm)
    bl = [[recs[i] for i in b] for b in batches]
    rec.recognize_batches(bl, merge=merge)
    torch.cuda.synchronize()
    L.dsb_profile_reset()
    L.dsb_profile_enable(1)
    t0 = time.perf_counter()
    out = rec.recognize_batches(bl, merge=merge)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    prof = N.profile_read()
    L.dsb_profile_enable(0)
    print(json.dumps({"merge": merge, "pair": pair, "batch_minor": bm, "seconds": round(dt, 4),
                      "stages_ms": {k: round(v[0], 2) for k, v in prof.items()}}), flush=True)
