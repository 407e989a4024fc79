"""BASELINE config 4: CPUStreamingRNN-shaped uni-GRU + lookahead, chunked streaming over S lock-step streams.
Prints one JSON line.  `rtfx` = audio-seconds per second of the model alone (spectrogram chunks resident in HBM,
CUDA-event timed); `e2e_rtfx` = the same through MultiStreamRecognizer.push with pinned host audio (H2D copy,
streaming spectrogram, model, greedy decode and transcript stitching inside the timed region, wall clock)."""
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
from danspeech_b200.streaming import MultiStreamRecognizer  # noqa: E402
from danspeech_b200.utils import synthetic as syn  # noqa: E402

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
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
frames = run()
e1.record()
torch.cuda.synchronize()
dt = e0.elapsed_time(e1) / 1e3
prof = N.profile_read()
N.lib().dsb_profile_enable(0)
audio_s = S * (8640 + 6240 * (CHUNKS - 1)) / 16000.0

# ---- end to end: pinned host audio -> transcripts ----
n_total = 8640 + 6240 * (CHUNKS - 1)
base = [syn.synthetic_audio(n_total, seed=900 + i) for i in range(8)]
host = torch.stack([torch.from_numpy(base[s % 8].astype("float32")) for s in range(S)]).pin_memory()
eng = MultiStreamRecognizer(model, S)


def run_e2e():
    outs = None
    for i in range(CHUNKS):
        a = 0 if i == 0 else 8640 + 6240 * (i - 1)
        b = 8640 if i == 0 else a + 6240
        outs = eng.push(host[:, a:b], i == 0, i == CHUNKS - 1)
    return outs


run_e2e()
torch.cuda.synchronize()
t0 = time.perf_counter()
texts = run_e2e()
torch.cuda.synchronize()
dt2 = time.perf_counter() - t0
print(json.dumps({"workload": "CPUStreamingRNN-shaped (2 conv, 5 x 800 uni-GRU, lookahead 20), %d lock-step streams, "
                              "%d chunks (8640 then 6240 samples)" % (S, CHUNKS),
                  "precision": PRECISION, "rtfx": audio_s / dt, "ms_per_chunk_step": 1e3 * dt / CHUNKS,
                  "frames_out_per_stream": frames,
                  "stages_ms": {k: round(v[0], 2) for k, v in prof.items() if v[0] > 0},
                  "e2e_rtfx": audio_s / dt2, "e2e_ms_per_chunk_step": 1e3 * dt2 / CHUNKS,
                  "e2e_h2d_bytes_per_chunk": int(S * 6240 * 4), "e2e_transcripts": len(texts),
                  "e2e_sample": texts[0][:40]}))
