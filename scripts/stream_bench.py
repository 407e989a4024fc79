"""BASELINE config 4: CPUStreamingRNN-shaped uni-GRU + lookahead, chunked streaming over S lock-step streams.
Prints one JSON line: audio-seconds per wall-second over the chunk schedule of the reference engine."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import __graft_entry__ as g  # noqa: E402
g.build()
from danspeech_b200 import _native as N  # noqa: E402
from danspeech_b200.pretrained_models import build_model  # noqa: E402

S = int(os.environ.get("STREAMS", "1024"))
CHUNKS = int(os.environ.get("CHUNKS", "8"))
PRECISION = os.environ.get("PRECISION", "bf16")
model = build_model("CPUStreamingRNN", seed=0).cuda().eval().set_precision(PRECISION)
gen = torch.Generator(device="cuda").manual_seed(0)
first = torch.randn((S, 1, 161, 53), generator=gen, device="cuda")
mid = torch.randn((S, 1, 161, 39), generator=gen, device="cuda")


def run():
    frames = 0
    for i in range(CHUNKS):
        o = model(first if i == 0 else mid, i == 0, i == CHUNKS - 1)
        frames += 0 if o is None else o.shape[1]
    return frames


run()
torch.cuda.synchronize()
N.lib().dsb_profile_reset()
N.lib().dsb_profile_enable(1)
t0 = time.perf_counter()
frames = run()
torch.cuda.synchronize()
dt = time.perf_counter() - t0
prof = N.profile_read()
audio_s = S * (8640 + 6240 * (CHUNKS - 1)) / 16000.0
print(json.dumps({"workload": "CPUStreamingRNN-shaped (2 conv, 5 x 800 uni-GRU, lookahead 20), %d lock-step streams, "
                              "%d chunks (8640 then 6240 samples)" % (S, CHUNKS),
                  "precision": PRECISION, "rtfx": audio_s / dt, "ms_per_chunk_step": 1e3 * dt / CHUNKS, "frames_out_per_stream": frames,
                  "streams_real_time": audio_s / dt, "stages_ms": {k: round(v[0], 2) for k, v in prof.items() if v[0] > 0}}))
