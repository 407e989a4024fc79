"""Short single-step run of the hot path used as the ncu target (never a bench number)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import __graft_entry__ as _g  # noqa: E402
_g.build()

from danspeech_b200 import Recognizer  # noqa: E402
from danspeech_b200.pretrained_models import build_model  # noqa: E402
from danspeech_b200.utils import synthetic as syn  # noqa: E402

precision = os.environ.get("PRECISION", "fp32")
name = os.environ.get("NCU_MODEL", "DanSpeechPrimary")
B = int(os.environ.get("NCU_BATCH", "64"))
secs = float(os.environ.get("NCU_SECONDS", "15"))
model = build_model(name, seed=0).set_precision(precision)
rec = Recognizer(model=model)
auds = [syn.synthetic_audio(int(secs * 16000), seed=i) for i in range(B)]
for _ in range(int(os.environ.get("NCU_STEPS", "1"))):
    out = rec.recognize_batch(auds)
torch.cuda.synchronize()
print("ok", len(out), repr(out[0][:40]))
