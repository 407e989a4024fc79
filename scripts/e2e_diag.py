"""Diagnostic: where does the end-to-end leg of bench.py lose time at N > 1?  Run under torchrun; every rank prints
its H2D bandwidth from pinned memory (alone and with all ranks copying at once) and its recognize_batches step time."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
import __graft_entry__ as g  # noqa: E402
if rank == 0:
    g.build()
if world > 1:
    dist.barrier()
from danspeech_b200 import Recognizer  # noqa: E402
from danspeech_b200.pretrained_models import build_model  # noqa: E402

B, n = 64, 240000
host = torch.randn((B, n), dtype=torch.float32).mul_(3000).pin_memory()
d = torch.empty_like(host, device=dev)


def bar():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def h2d(reps=10):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        d.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    return host.numel() * 4 * reps / (time.perf_counter() - t0) / 1e9


out = {"rank": rank, "cpus": len(os.sched_getaffinity(0)), "omp": os.environ.get("OMP_NUM_THREADS")}
h2d(2)
for r in range(world):          # one rank at a time
    bar()
    if r == rank:
        out["h2d_alone_GBs"] = round(h2d(), 1)
bar()
out["h2d_together_GBs"] = round(h2d(), 1)

rec = Recognizer(model=build_model("DanSpeechPrimary", seed=0).set_precision("bf16"), device=dev)
batch = (host, [n] * B)
rec.recognize_batches([batch] * 3)
for steps in (5, 20):
    bar()
    t0 = time.perf_counter()
    rec.recognize_batches([batch] * steps)
    torch.cuda.synchronize()
    out["e2e_ms_per_step_%d" % steps] = round(1e3 * (time.perf_counter() - t0) / steps, 2)
for r in range(world):          # one rank at a time: is it contention between the ranks?
    bar()
    if r == rank:
        t0 = time.perf_counter()
        rec.recognize_batches([batch] * 5)
        torch.cuda.synchronize()
        out["e2e_ms_per_step_alone"] = round(1e3 * (time.perf_counter() - t0) / 5, 2)
bar()
# device-only steps at the same time on all ranks, wall clock
eng = rec.danspeech_recognizer
audio_dev = host.to(dev)
n_dev = torch.full((B,), n, dtype=torch.int32, device=dev)
lengths = torch.IntTensor([1 + n // 160] * B)
t0 = time.perf_counter()
for _ in range(5):
    spect, _ = eng.audio_parser.parse_device(audio_dev, n_dev, n)
    probs, sizes = eng.model(spect.view(B, 1, 161, -1), lengths)
    eng.decoder.decode_device(probs, sizes)
torch.cuda.synchronize()
out["device_ms_per_step_wall"] = round(1e3 * (time.perf_counter() - t0) / 5, 2)
print(json.dumps(out), flush=True)
bar()
if world > 1:
    dist.destroy_process_group()
