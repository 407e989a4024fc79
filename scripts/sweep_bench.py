"""BASELINE config 5: utterance-sharded sweep -- 256 utterances of 5-30 s split across the ranks (LPT bin packing
on length, length-sorted batches of <= 64 per rank, no data-path collective), greedy and beam-64 + 3-gram LM.
Run alone (1 GPU) or under torchrun (one rank per GPU); rank 0 prints one JSON line.  Wall-clock, end to end from
host float audio to transcripts; max over ranks.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/sweep_bench.py
"""
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import __graft_entry__ as g  # noqa: E402

rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
if rank == 0:
    g.build()
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dist.barrier()
from danspeech_b200 import Recognizer, sharding  # noqa: E402
from danspeech_b200.pretrained_models import build_model  # noqa: E402
from danspeech_b200.utils import synthetic as syn  # noqa: E402

N_UTT = int(os.environ.get("UTTERANCES", "256"))
PRECISION = os.environ.get("PRECISION", "bf16")
MAX_BATCH = int(os.environ.get("MAX_BATCH", "64"))
rng = np.random.default_rng(1234)
lens = rng.integers(5 * 16000, 30 * 16000, size=N_UTT)
# every rank synthesises only what it owns (deterministic by utterance id)
mine = sharding.lpt_shards(lens.tolist(), world)[rank]
recs = {i: syn.synthetic_audio(int(lens[i]), seed=1234 + i) for i in mine}
audio_s_total = float(lens.sum()) / 16000.0


def run(rec):
    out = {}
    batches = sharding.make_batches(mine, lens.tolist(), max_batch=MAX_BATCH)
    for batch, texts in zip(batches, rec.recognize_batches([[recs[i] for i in b] for b in batches])):
        for i, t in zip(batch, texts):
            out[i] = t
    return out


def timed(rec):
    run(rec)                                   # warm-up (allocations, autotuned nothing: just caches)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    out = run(rec)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        merged = sharding.gather_transcripts(out, world)
    else:
        merged = out
    return float(dt.item()), merged


model = build_model("DanSpeechPrimary", seed=0).set_precision(PRECISION)
res = {"workload": "%d utterances of 5-30 s (%.0f audio-s), DanSpeechPrimary-shaped, %s mode, LPT shards over %d GPU(s), "
                   "batches <= %d" % (N_UTT, audio_s_total, PRECISION, world, MAX_BATCH), "n_gpus": world}
rec = Recognizer(model=model)
dt, texts = timed(rec)
res["greedy"] = {"seconds": dt, "rtfx": audio_s_total / dt, "utt_per_s": N_UTT / dt, "transcripts": len(texts)}
with tempfile.TemporaryDirectory() as td:
    arpa = os.path.join(td, "lm_%d.arpa" % rank)
    syn.write_synthetic_arpa(arpa, n_words=2000, seed=7)
    rec.update_decoder(lm=arpa, alpha=1.3, beta=0.2, beam_width=64)
    dt, texts_b = timed(rec)
res["beam64_lm"] = {"seconds": dt, "rtfx": audio_s_total / dt, "utt_per_s": N_UTT / dt, "transcripts": len(texts_b)}
if rank == 0:
    import hashlib
    res["greedy_digest"] = hashlib.sha1("\n".join(texts[i] for i in sorted(texts)).encode()).hexdigest()[:12]
    print(json.dumps(res))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
