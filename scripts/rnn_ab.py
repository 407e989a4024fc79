"""A/B of the persistent recurrence schedules on the headline shape: one Primary-shaped bf16 forward at B = 64, 128,
192 x SECONDS s with 1, 2, 3 batch groups in flight (dsb_tune_set), stage timers from the library.  Diagnostic."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import __graft_entry__ as _g  # noqa: E402
_g.build()
from danspeech_b200 import _native as N  # noqa: E402
from danspeech_b200.pretrained_models import build_model  # noqa: E402

SECONDS = float(os.environ.get("SECONDS", "15"))
LAYERS = int(os.environ.get("LAYERS", "9"))
BS = [int(b) for b in os.environ.get("BATCHES", "64,128,192").split(",")]
NIFS = [int(b) for b in os.environ.get("NIFS", "1,2,3").split(",")]
REPS = int(os.environ.get("REPS", "3"))
KSPLITS = [int(b) for b in os.environ.get("KSPLITS", "0,1").split(",")]   # 0 one CTA, 1 K-split pairs, 2 cta_group::2 pairs
T = 1 + int(SECONDS * 16000) // 160
BIDIR = os.environ.get("BIDIR", "1") == "1"
H = int(os.environ.get("HIDDEN", "1200"))
model = build_model("DanSpeechPrimary", seed=0, rnn_layers=LAYERS, bidirectional=BIDIR,
                    rnn_hidden_size=H).cuda().eval().set_precision("bf16")
L = N.lib()
gen = torch.Generator(device="cuda").manual_seed(0)
for B in BS:
    x = torch.randn((B, 1, 161, T), generator=gen, device="cuda")
    lens = torch.IntTensor([T] * B)
    ref = None
    for nif, ks in [(n, k) for n in NIFS for k in KSPLITS]:
        if (nif > 1 and B <= 64) or (nif == 1 and ks == 1 and B > 64) or (ks == 2 and B <= 64):
            continue
        if ks == 2:     # cta_group::2 pairs: nif = PAIR ITEMS (two groups each) in flight
            if B < 128 * nif and nif > 1 and B <= 128:
                continue
            N.tune(rnn_in_flight=3, rnn_ksplit=0, rnn_pair=1, rnn_pair_in_flight=nif,
                   rnn_ring_gsz=int(os.environ.get("RING_GSZ", "0")))
        else:
            N.tune(rnn_in_flight=max(nif, 2) if ks else nif, rnn_ksplit=ks, rnn_pair=0,
                   rnn_ring_gsz=int(os.environ.get("RING_GSZ", "0")), rnn_producers=int(os.environ.get("PRODUCERS", "1")))
        probs, _ = model(x, lens)
        torch.cuda.synchronize()
        L.dsb_profile_reset()
        L.dsb_profile_enable(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(REPS):
            probs, _ = model(x, lens)
        e1.record()
        torch.cuda.synchronize()
        prof = N.profile_read()
        L.dsb_profile_enable(0)
        ms = e0.elapsed_time(e1) / REPS
        if ref is None:
            ref = probs.clone()
        diff = float((probs - ref).abs().max())
        print(json.dumps({"B": B, "in_flight": nif, "ksplit": ks, "ms_per_forward": round(ms, 3), "ms_per_64": round(ms * 64 / B, 3),
                          "rnn_ms_per_64": round(prof["rnn_recurrence"][0] / REPS * 64 / B, 3),
                          "proj_ms_per_64": round(prof["rnn_input_proj"][0] / REPS * 64 / B, 3),
                          "conv_ms_per_64": round(prof["conv"][0] / REPS * 64 / B, 3),
                          "max_prob_diff_vs_first": diff}), flush=True)
N.tune(rnn_in_flight=3, rnn_ksplit=0, rnn_pair=1, rnn_pair_in_flight=2)
