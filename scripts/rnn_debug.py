"""One Primary-shaped bf16 forward with DSB_RNN_DEBUG=1: prints the per-phase cycle breakdown of the
persistent recurrence kernel (diagnostic, not a bench)."""
import os
import sys
os.environ["DSB_RNN_DEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import __graft_entry__ as _g  # noqa: E402
_g.build()
from danspeech_b200 import Recognizer  # noqa: E402
from danspeech_b200.pretrained_models import build_model  # noqa: E402
from danspeech_b200.utils import synthetic as syn  # noqa: E402

layers = int(os.environ.get("LAYERS", "2"))
model = build_model("DanSpeechPrimary", seed=0, rnn_layers=layers).set_precision("bf16")
rec = Recognizer(model=model)
auds = [syn.synthetic_audio(15 * 16000, seed=i) for i in range(64)]
for _ in range(2):
    out = rec.recognize_batch(auds)
torch.cuda.synchronize()
print("ok")
