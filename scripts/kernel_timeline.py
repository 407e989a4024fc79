"""Per-kernel launch list of one pass of the hot path WITHOUT replay: CUPTI activity records through torch.profiler
(ncu cannot replay the cooperative cluster launch of the CTA-pair recurrence, and serialises kernels; this records the
real, concurrent durations).  Prints a CSV (kernel, launches, total_ms, share) and a JSON summary line.

  PRECISION=bf16 BATCHES=4 python scripts/kernel_timeline.py > gpurun_out/kernels.csv
"""
import json
import os
import sys
from collections import OrderedDict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import __graft_entry__ as _g  # noqa: E402
_g.build()
from torch.profiler import ProfilerActivity, profile  # noqa: E402
from danspeech_b200 import Recognizer  # noqa: E402
from danspeech_b200.pretrained_models import build_model  # noqa: E402
from danspeech_b200.utils import synthetic as syn  # noqa: E402

precision = os.environ.get("PRECISION", "bf16")
n_batches = int(os.environ.get("BATCHES", "4"))
secs = float(os.environ.get("SECONDS", "15"))
rec = Recognizer(model=build_model("DanSpeechPrimary", seed=0).set_precision(precision))
auds = [syn.synthetic_audio(int(secs * 16000), seed=i) for i in range(64)]
batches = [auds] * n_batches
rec.recognize_batches(batches)
rec.recognize_batches(batches)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    rec.recognize_batches(batches)
    torch.cuda.synchronize()
agg = OrderedDict()
for ev in prof.events():
    if ev.device_type.name != "CUDA" or ev.device_time_total <= 0:
        continue
    name = ev.name.split("(")[0].replace("void ", "").replace("dsb::", "")
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ev.device_time_total / 1e3
total = sum(a[1] for a in agg.values())
print("kernel,launches,total_ms,share_pct")
for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('"%s",%d,%.4f,%.2f' % (name[:90], n, ms, 100 * ms / total))
print("# " + json.dumps({"precision": precision, "batches_per_pass": n_batches, "seconds": secs, "sum_kernel_ms": round(total, 3),
                         "ms_per_batch_of_64": round(total / n_batches, 3)}))
