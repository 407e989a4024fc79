#!/usr/bin/env python
"""Benchmark of the DanSpeech inference hot path (audio -> spectrogram -> DeepSpeech2 -> CTC decode).

Contract: ``python bench.py --gpus N --steps K --warmup W`` prints ONE JSON line (rank 0).  Under
torchrun (N > 1) every rank processes its own batch of 64 utterances (utterance sharding, no
collective on the data path; "scaling": "weak"); without a launcher environment, ``--gpus N`` with N > 1
starts the N ranks itself the same way (``torch.distributed.run``, rendezvous on 127.0.0.1).

Workload (BASELINE.json configs[1]): DanSpeechPrimary-shaped bi-GRU DeepSpeech2 (3 conv + 9 x 1200
bi-GRU, random-init), batch 64 x 15 s synthetic 16 kHz audio, greedy decode.  Metric: audio-seconds
per wall-second (RTFx).
  value : device-timed (CUDA events), audio already resident in HBM; K steps = K batches of 64, up to three
          consecutive batches sharing one pass of the model (`single_batch` is the one-batch-per-pass figure)
  e2e   : through Recognizer.recognize_batches from HOST lists of numpy float64 arrays (the API's input type): every
          step's float64->float32 staging, H2D copy, kernels, D2H of the token/offset tensors and transcript string
          building inside the timed region (staging + copy of pass k+1 overlap the kernels of pass k)
  plus the driver-visible secondary workloads: `streaming` (config 4), `sweep` (config 5), `fp32` mode, `library_bar`
  (stock torch.nn / cuDNN forward on the same GPU), `parity` (benchmarked mode vs the oracle), `beam` (config 3)
``--impl reference`` times the oracle's CPU restatement of the reference path (torch CPU, all host
threads) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

BATCH = 64
SECONDS = 15
SR = 16000
MODEL = "DanSpeechPrimary"
WORKLOAD = "DanSpeechPrimary-shaped bi-GRU DeepSpeech2 (3 conv + 9x1200 bi-GRU, random init), batch 64 x 15 s synthetic 16 kHz audio, greedy decode"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("DANSPEECH_B200_PRECISION", "bf16"),
                    choices=["fp32", "bf16"])
    ap.add_argument("--cpu-sample", type=int, default=1, help="utterances in the cpu_baseline sample")
    ap.add_argument("--ref-batch", type=int, default=0,
                    help="utterances per reference-arm step (0 = as many as keep the whole run near three minutes, <= 8)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-beam", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the secondary workloads (streaming, sweep, fp32 mode, library bar)")
    ap.add_argument("--merge", type=int, default=4,
                    help="batches of 64 that share one pass of the model (1..4; 1 = one batch per pass)")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------- helpers
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


def make_audio(n_utts, seed0):
    from danspeech_b200.utils import synthetic as syn
    return [syn.synthetic_audio(SECONDS * SR, seed=seed0 + i) for i in range(n_utts)]


def model_flops(B, Tp, conv_layers=3, layers=9, H=1200):
    conv = 2.0 * Tp * (32 * 81 * 451 + 32 * 41 * 7392 + 96 * 21 * 7392) * B
    proj = 2.0 * Tp * 2 * 3 * H * (2016 + (layers - 1) * H) * B
    rec = 2.0 * Tp * 2 * layers * H * 3 * H * B
    return conv, proj, rec


# ----------------------------------------------------------------------------------------- CPU arm
def cpu_model_name():
    try:
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.lower().startswith("model name"):
                    return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_reference_rtfx(n_utts, reps, warmup, threads=None, want_outputs=False):
    """Oracle CPU restatement of the reference path (spectrogram -> DeepSpeech.forward -> greedy)."""
    from danspeech_b200.utils import synthetic as syn
    from oracle import greedy as og
    from oracle import model as om
    from oracle import spectrogram as osp
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    cfg = dict(syn.MODEL_SHAPES[MODEL])
    sd = syn.make_state_dict(seed=0, **cfg)
    auds = make_audio(n_utts, 0)
    parser = osp.SpectrogramOracle()

    def step():
        specs = [parser.parse_audio(a) for a in auds]
        x = torch.zeros(len(specs), 1, 161, specs[0].size(1))
        for i, s in enumerate(specs):
            x[i, 0] = s
        lens = torch.IntTensor([s.size(1) for s in specs])
        with torch.no_grad():
            probs, sizes = om.forward(sd, x, lens, cfg["conv_layers"], cfg["rnn_layers"])
        return og.greedy_decode(probs.numpy(), sizes.numpy()), probs.numpy()

    for _ in range(warmup):
        step()
    times, out = [], None
    for _ in range(reps):
        t0 = time.perf_counter()
        out = step()
        times.append(time.perf_counter() - t0)
    sec = float(np.mean(times))
    if want_outputs:
        return n_utts * SECONDS / sec, sec, threads, [t[0] for t in out[0][0]], out[1]
    return n_utts * SECONDS / sec, sec, threads


def reference_batch(args):
    # bounded sample: a 15 s utterance costs the host ~1.3 s, so size the step for ~150 s over steps + warm-up
    return args.ref_batch or max(1, min(8, int(150.0 / (max(1, args.steps) + 1) / 1.3)))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nb = reference_batch(args)
    steps = max(1, args.steps)
    rtfx, sec, threads = cpu_reference_rtfx(nb, steps, max(1, min(args.warmup, 1)))
    line = {
        "impl": "reference", "metric": "audio-sec/sec (RTFx)", "value": rtfx, "unit": "audio-s/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": max(1, min(args.warmup, 1)), "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "reference_step": "%d of the 64 utterances per step" % nb},
        "cpu_baseline": {"value": rtfx, "unit": "audio-s/s", "cores": threads, "kind": "port", "cpu_model": cpu_model_name(),
                         "sample": "%d x 15 s utterances per step (oracle restatement of parsers.py:50-72 + "
                                   "model.py:496-515 + decoder.py:183-198 on torch CPU fp32)" % nb},
        "e2e": {"value": rtfx, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------- GPU arm
def all_max(x, dev, world):
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def passes_for(k, merge):
    """K steps (batches) as passes of at most `merge` batches: exactly K batches are processed."""
    out = [merge] * (k // merge)
    if k % merge:
        out.append(k % merge)
    return out


def bench_streaming(dev, rank, world, precision):
    """BASELINE config 4: CPUStreamingRNN-shaped uni-GRU + lookahead, chunked streaming over 1024 lock-step streams
    (1024 / world per GPU).  `value`: model alone on spectrogram chunks resident in HBM (CUDA events); `e2e`: pinned
    host audio -> MultiStreamRecognizer.push (H2D copy, streaming spectrogram, model, greedy decode, stitching)."""
    from danspeech_b200.pretrained_models import build_model
    from danspeech_b200.streaming import MultiStreamRecognizer
    from danspeech_b200.utils import synthetic as syn
    S, chunks = 1024 // world, 8
    model = build_model("CPUStreamingRNN", seed=0).to(dev).eval().set_precision(precision)
    gen = torch.Generator(device=dev).manual_seed(rank)
    first = torch.randn((S, 1, 161, 53), generator=gen, device=dev)
    mid = torch.randn((S, 1, 161, 39), generator=gen, device=dev)

    def run():
        for i in range(chunks):
            model(first if i == 0 else mid, i == 0, i == chunks - 1)

    run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run()
    e1.record()
    torch.cuda.synchronize()
    dt = all_max(e0.elapsed_time(e1) / 1e3, dev, world)
    n_total = 8640 + 6240 * (chunks - 1)
    audio_s = S * world * n_total / 16000.0
    base = [syn.synthetic_audio(n_total, seed=900 + i) for i in range(8)]
    host = torch.stack([torch.from_numpy(base[s % 8].astype("float32")) for s in range(S)]).pin_memory()
    eng = MultiStreamRecognizer(model, S)

    def run_e2e():
        outs = None
        for i in range(chunks):
            a = 0 if i == 0 else 8640 + 6240 * (i - 1)
            b = 8640 if i == 0 else a + 6240
            outs = eng.push(host[:, a:b], i == 0, i == chunks - 1)
        return outs

    run_e2e()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    texts = run_e2e()
    torch.cuda.synchronize()
    dt2 = all_max(time.perf_counter() - t0, dev, world)
    del eng, model
    return {"workload": "CPUStreamingRNN-shaped (2 conv, 5 x 800 uni-GRU, lookahead 20), %d lock-step streams (%d per GPU), "
                        "%d chunks (8640 then 6240 samples), %s mode" % (S * world, S, chunks, precision),
            "value": audio_s / dt, "unit": "audio-s/s", "ms_per_chunk_step": 1e3 * dt / chunks,
            "e2e": {"value": audio_s / dt2, "unit": "audio-s/s", "h2d_bytes_per_chunk": int(S * 6240 * 4),
                    "transcripts": len(texts) * world}}


def bench_sweep(rec, dev, rank, world, precision):
    """BASELINE config 5: 256 utterances of 5-30 s split across the ranks (LPT bin packing on length, length-sorted
    batches of <= 64 per rank, no data-path collective), greedy and beam-64 + 3-gram LM; wall clock from host
    float64 audio to transcripts, max over ranks: STRONG scaling (the job is fixed, the ranks share it)."""
    import hashlib
    import tempfile
    from danspeech_b200 import sharding
    from danspeech_b200.utils import synthetic as syn
    n_utt = 256
    lens = np.random.default_rng(1234).integers(5 * SR, 30 * SR, size=n_utt)
    mine = sharding.lpt_shards(lens.tolist(), world)[rank]
    recs = {i: syn.synthetic_audio(int(lens[i]), seed=1234 + i) for i in mine}
    audio_s = float(lens.sum()) / SR
    batches = sharding.make_batches(mine, lens.tolist(), max_batch=BATCH)

    def run():
        out = {}
        for batch, texts in zip(batches, rec.recognize_batches([[recs[i] for i in b] for b in batches])):
            for i, t in zip(batch, texts):
                out[i] = t
        return out

    def timed():
        import torch.distributed as dist
        run()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        best = None
        for _ in range(2):      # best of two timed runs (a one-off host hiccup -- allocator, GC -- is not the workload)
            t0 = time.perf_counter()
            out = run()
            torch.cuda.synchronize()
            dt = all_max(time.perf_counter() - t0, dev, world)
            best = dt if best is None else min(best, dt)
            if world > 1:
                dist.barrier()
        return best, sharding.gather_transcripts(out, world)

    res = {"workload": "256 utterances of 5-30 s (%.0f audio-s), DanSpeechPrimary-shaped, %s mode, LPT shards over %d "
                       "GPU(s), batches <= %d, host float64 audio -> transcripts, best of 2 timed runs" % (audio_s, precision, world, BATCH),
           "scaling": "strong"}
    dt, texts = timed()
    res["greedy"] = {"value": audio_s / dt, "unit": "audio-s/s", "seconds": dt, "utt_per_s": n_utt / dt}
    res["greedy_digest"] = hashlib.sha1("\n".join(texts[i] for i in sorted(texts)).encode()).hexdigest()[:12]
    with tempfile.TemporaryDirectory() as td:
        arpa = os.path.join(td, "lm_%d.arpa" % rank)
        syn.write_synthetic_arpa(arpa, n_words=2000, seed=7)
        rec.update_decoder(lm=arpa, alpha=1.3, beta=0.2, beam_width=64)
        dt, texts_b = timed()
        rec.update_decoder(lm="greedy")
    res["beam64_lm"] = {"value": audio_s / dt, "unit": "audio-s/s", "seconds": dt, "utt_per_s": n_utt / dt}
    res["beam_digest"] = hashlib.sha1("\n".join(texts_b[i] for i in sorted(texts_b)).encode()).hexdigest()[:12]
    return res


def bench_library_bar(dev):
    """SURVEY 8(d) "library kernels" bar: the same network written with stock torch.nn modules (cuDNN convolutions and
    GRU, cuBLAS) on this GPU, model forward + argmax only.  Uses neither oracle/ nor the reference nor this package."""
    import torch.nn as nn
    H, layers, C = 1200, 9, 33
    T = 1 + SECONDS * SR // 160

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv = nn.Sequential(
                nn.Conv2d(1, 32, (41, 11), stride=(2, 2), padding=(20, 5)), nn.BatchNorm2d(32), nn.Hardtanh(0, 20),
                nn.Conv2d(32, 32, (21, 11), stride=(2, 1), padding=(10, 5)), nn.BatchNorm2d(32), nn.Hardtanh(0, 20),
                nn.Conv2d(32, 96, (21, 11), stride=(2, 1), padding=(10, 5)), nn.BatchNorm2d(96), nn.Hardtanh(0, 20))
            self.norms = nn.ModuleList([nn.BatchNorm1d(H) for _ in range(layers - 1)])
            self.rnns = nn.ModuleList([nn.GRU(21 * 96 if i == 0 else H, H, bidirectional=True) for i in range(layers)])
            self.fc = nn.Sequential(nn.BatchNorm1d(H), nn.Linear(H, C, bias=False))

        def forward(self, x):
            x = self.conv(x)                                   # equal lengths: the time mask is the identity
            x = x.view(x.size(0), -1, x.size(3)).permute(2, 0, 1).contiguous()
            for i, rnn in enumerate(self.rnns):
                if i:
                    t, b, h = x.shape
                    x = self.norms[i - 1](x.view(t * b, h)).view(t, b, h)
                x, _ = rnn(x)
                x = x[..., :H] + x[..., H:]
            t, b, h = x.shape
            x = self.fc(x.view(t * b, h)).view(t, b, C).transpose(0, 1)
            return torch.softmax(x, dim=-1).argmax(dim=-1)

    def timed(fn, warm=2, reps=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    torch.manual_seed(0)
    net = Net().to(dev).eval()
    x = torch.randn(BATCH, 1, 161, T, device=dev)
    out = {"workload": "stock torch.nn DeepSpeech2 of the DanSpeechPrimary shape (cuDNN %s convolutions + GRU, cuBLAS), "
                       "batch %d x %d s, model forward + argmax only" % (torch.backends.cudnn.version(), BATCH, SECONDS),
           "unit": "audio-s/s"}
    with torch.no_grad():
        ms = timed(lambda: net(x))
        out["fp32_tf32"] = {"value": BATCH * SECONDS / (ms / 1e3), "ms_per_batch": ms}
        with torch.autocast("cuda", dtype=torch.bfloat16):
            ms = timed(lambda: net(x))
        out["bf16_autocast"] = {"value": BATCH * SECONDS / (ms / 1e3), "ms_per_batch": ms}
    del net, x
    torch.cuda.empty_cache()
    return out


def run_ours(args):
    import torch.distributed as dist
    from danspeech_b200 import Recognizer, _native as N
    from danspeech_b200.pretrained_models import build_model

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    import __graft_entry__ as g
    if rank == 0:
        g.build()
    if world > 1:
        dist.barrier()
    L = N.lib()
    merge = max(1, min(4, args.merge))

    model = build_model(MODEL, seed=0).set_precision(args.precision)
    rec = Recognizer(model=model, device=dev)
    eng = rec.danspeech_recognizer
    parser, decoder = eng.audio_parser, eng.decoder

    # each rank owns its own batch (utterance sharding); seeds differ per rank
    auds = make_audio(BATCH, seed0=rank * BATCH)
    n = SECONDS * SR
    host = torch.empty((BATCH, n), dtype=torch.float32).pin_memory()
    for i, a in enumerate(auds):
        host[i] = torch.from_numpy(a.astype(np.float32))
    # up to `merge` batches go through one pass of the model (Recognizer.recognize_batches does the same): the
    # batches in flight are copies of this rank's batch -- the kernels are data-independent
    audio_dev = host.to(dev).repeat(merge, 1)
    n_dev = torch.full((BATCH * merge,), n, dtype=torch.int32, device=dev)
    T = 1 + n // 160
    spect_buf = torch.empty((BATCH * merge, 161, T), dtype=torch.float32, device=dev)

    def step_device(m):
        rows = BATCH * m
        spect, _ = parser.parse_device(audio_dev[:rows], n_dev[:rows], n, out=spect_buf[:rows])
        probs, sizes = eng.model(spect.view(rows, 1, 161, -1), torch.IntTensor([T] * rows))
        return decoder.decode_device(probs, sizes)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_device(m_merge):
        for m in sorted(set(passes_for(args.steps, m_merge))) * max(args.warmup, 3):
            step_device(m)
        torch.cuda.synchronize()
        L.dsb_profile_reset()
        L.dsb_profile_enable(1)
        sampler = ClockSampler(local_rank)
        barrier()
        sampler.start()
        launches0 = N.launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for m in passes_for(args.steps, m_merge):
            step_device(m)
        ev1.record()
        barrier()
        launches = N.launch_count() - launches0
        clocks = sampler.stop()
        prof = N.profile_read()
        L.dsb_profile_enable(0)
        return all_max(ev0.elapsed_time(ev1), dev, world), prof, launches, clocks

    # ---- device-timed region: exactly K steps (batches of 64), CUDA events, max over ranks ----
    audio_s = BATCH * SECONDS * world
    ms_max, prof, launches, clocks = timed_device(merge)
    value = audio_s * args.steps / (ms_max / 1e3)
    single = None
    if merge > 1:
        ms1, prof1, _, _ = timed_device(1)
        single = {"value": audio_s * args.steps / (ms1 / 1e3), "unit": "audio-s/s", "ms_per_step": ms1 / args.steps,
                  "note": "one batch of 64 per pass (nothing in flight): the latency of a single batch",
                  "rnn_recurrence_ms_per_step": prof1["rnn_recurrence"][0] / args.steps}

    # ---- end to end through the public API, from HOST buffers of the API's own input type: lists of numpy float64
    # arrays (Recognizer.py:82-95).  Every step's float64 -> float32 staging, H2D copy, kernels, D2H of the
    # token/offset tensors and transcript string building are inside the timed region; recognize_batches overlaps the
    # staging + copy of pass k+1 (helper thread, side stream) with the kernels of pass k. ----
    def timed_e2e(batches):
        rec.recognize_batches(batches[:merge] * max(1, min(args.warmup, 2)))
        barrier()
        t0 = time.perf_counter()
        texts = rec.recognize_batches(batches)
        torch.cuda.synchronize()
        return all_max(time.perf_counter() - t0, dev, world), texts

    e2e_s, texts_all = timed_e2e([auds] * args.steps)
    texts = texts_all[-1]
    e2e_pinned_s, _ = timed_e2e([(host, [n] * BATCH)] * args.steps)
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, texts)     # final host gather of transcripts (off the data path)
        n_texts = sum(len(x) for x in gathered)
    else:
        n_texts = len(texts)
    e2e_value = audio_s * args.steps / e2e_s

    # ---- secondary: BASELINE config 3, beam-64 decode with a synthetic 3-gram ARPA LM (utterances/s) ----
    beam = None
    if not args.no_beam:
        import tempfile
        from danspeech_b200.deepspeech.decoder import BeamCTCDecoder
        from danspeech_b200.utils import synthetic as syn
        arpa = os.path.join(tempfile.mkdtemp(prefix="dsb_lm_%d_" % rank), "synthetic3gram.arpa")
        syn.write_synthetic_arpa(arpa, n_words=2000, seed=0)
        bdec = BeamCTCDecoder(labels=eng.labels, lm_path=arpa, alpha=1.3, beta=0.2, beam_width=64, num_processes=6,
                              cutoff_prob=1.0, cutoff_top_n=40, blank_index=eng.labels.index("_"))
        spect, _ = parser.parse_device(audio_dev[:BATCH], n_dev[:BATCH], n, out=spect_buf[:BATCH])
        probs, sizes = eng.model(spect.view(BATCH, 1, 161, -1), torch.IntTensor([T] * BATCH))
        bdec.decode_device(probs, sizes)
        barrier()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record()
        for _ in range(args.steps):
            bdec.decode_device(probs, sizes)
        b1.record()
        barrier()
        beam_ms = all_max(b0.elapsed_time(b1), dev, world) / args.steps
        del bdec
        # the whole config-3 pipeline through the public API: host float64 audio -> beam transcripts, the beam search
        # of pass j (decode stream) under the forward of pass j+1
        rec.update_decoder(lm=arpa, alpha=1.3, beta=0.2, beam_width=64)
        beam_e2e_s, beam_texts = timed_e2e([auds] * args.steps)
        rec.update_decoder(lm="greedy")
        beam = {"utt_per_s": BATCH * world / (beam_ms / 1e3), "ms_per_batch": beam_ms, "beam_width": 64,
                "lm": "synthetic 3-gram ARPA, 2000 words, alpha 1.3, beta 0.2",
                "rtfx_forward_plus_beam": audio_s / ((ms_max / args.steps + beam_ms) / 1e3),
                "rtfx_forward_plus_beam_note": "device-timed forward and beam kernels back to back (no overlap)",
                "e2e": {"value": audio_s * args.steps / beam_e2e_s, "unit": "audio-s/s",
                        "utt_per_s": BATCH * world * args.steps / beam_e2e_s,
                        "note": "Recognizer.recognize_batches with the beam-64 + LM decoder from host float64 lists: "
                                "beam search of pass j on the decode stream under the forward of pass j+1"}}
    Tp = (T - 1) // 2 + 1
    d2h = BATCH * (1 + 2 * Tp) * 4

    # ---- driver-visible secondary workloads (BASELINE configs 4 and 5, fp32 mode, library bar) ----
    extras = {}
    if not args.no_extras:
        probs_bench = None
        if rank == 0 and not args.no_cpu_baseline and world == 1:
            spect, _ = parser.parse_device(audio_dev[:BATCH], n_dev[:BATCH], n, out=spect_buf[:BATCH])
            probs_bench = eng.model(spect.view(BATCH, 1, 161, -1), torch.IntTensor([T] * BATCH))[0].cpu().numpy()
        extras["probs_bench"] = probs_bench
        extras["streaming"] = bench_streaming(dev, rank, world, args.precision)
        extras["sweep"] = bench_sweep(rec, dev, rank, world, args.precision)
        # fp32 mode (exact accumulation, greedy transcripts bit-exact with the reference): two timed steps
        eng.model.set_precision("fp32")
        eng.update_model(eng.model)
        parser, decoder = eng.audio_parser, eng.decoder
        step_device(1)
        barrier()
        L.dsb_profile_reset()
        L.dsb_profile_enable(1)
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(2):
            step_device(1)
        f1.record()
        barrier()
        fms = all_max(f0.elapsed_time(f1), dev, world) / 2
        fprof = N.profile_read()
        L.dsb_profile_enable(0)
        extras["fp32"] = {"value": audio_s / (fms / 1e3), "unit": "audio-s/s", "ms_per_step": fms, "steps": 2,
                          "stages_ms": {k: round(v[0] / 2, 3) for k, v in fprof.items() if v[0] > 0},
                          "note": "fp32 mode: CUDA-core kernels with exact fp32 accumulation, logits within 1e-4 and greedy "
                                  "transcripts bit-exact against the oracle (tests/test_gpu_headline.py)"}
        eng.model.set_precision(args.precision)
        eng.update_model(eng.model)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel family, from the live stage timers ----
    peaks, peak_src = measured_peaks()
    conv_f, proj_f, rec_f = model_flops(BATCH, Tp)
    stage_ms = {k: v[0] / max(args.steps, 1) for k, v in prof.items()}
    dominant = max(("conv", "rnn_input_proj", "rnn_recurrence"), key=lambda k: stage_ms[k])
    flops = {"conv": conv_f, "rnn_input_proj": proj_f, "rnn_recurrence": rec_f}[dominant]
    achieved = flops / (stage_ms[dominant] / 1e3) / 1e12 if stage_ms[dominant] > 0 else 0.0
    peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1400.0)))
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if os.path.exists(tp) and args.precision == "bf16":
        tj = json.load(open(tp)).get(dominant, {})
        traffic, traffic_src = tj.get("bytes_per_launch"), "profiles/r02_traffic.json (one ncu --set full capture: " \
            "dram__bytes_read.sum + dram__bytes_write.sum of this kernel, %s)" % tj.get("capture", "")
    n_launch = prof[dominant][1] / max(args.steps, 1)
    roofline = {"bound": "tensor", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": peak_src + ", sustained bf16 (kernel timed inside a long step)",
                "ms_per_step": stage_ms[dominant], "launches_per_step": n_launch,
                "algorithmic_flop_per_launch": flops * args.steps / max(prof[dominant][1], 1)}
    spect_bytes = BATCH * (4 * n + 4 * 161 * T)
    stages = {k: {"ms_per_step": round(v, 4)} for k, v in stage_ms.items() if v > 0}
    if stage_ms["spectrogram"] > 0:
        gbs = spect_bytes / (stage_ms["spectrogram"] / 1e3) / 1e9
        stages["spectrogram"].update({"GB/s": round(gbs, 1), "frac_hbm": round(gbs / float(peaks["hbm_gbs"]), 4)})
    for k, f in (("conv", conv_f), ("rnn_input_proj", proj_f), ("rnn_recurrence", rec_f)):
        if stage_ms[k] > 0:
            tf = f / (stage_ms[k] / 1e3) / 1e12
            stages[k].update({"TFLOP/s": round(tf, 2), "frac_tensor": round(tf / peak, 4)})

    cpu, cpu_batch, parity = None, None, None
    if not args.no_cpu_baseline and world == 1:      # rank 0 at N = 1 only (under torchrun the other ranks spin in a barrier
                                                     # on the same host cores and OMP_NUM_THREADS is forced to 1)
        v, sec, threads, ref_texts, ref_probs = cpu_reference_rtfx(args.cpu_sample, 2, 1, want_outputs=True)
        cpu = {"value": v, "unit": "audio-s/s", "cores": threads, "kind": "port", "cpu_model": cpu_model_name(),
               "sample": "batch %d (the reference engine's own batch size, DanSpeechRecognizer.py:220-223): %d x 15 s "
                         "utterance(s) per repetition, 1 warm-up + 2 timed, oracle CPU restatement (torch CPU fp32, %d "
                         "threads)" % (args.cpu_sample, args.cpu_sample, threads)}
        nb = reference_batch(args)
        vb, secb, _, ref_texts, ref_probs = cpu_reference_rtfx(nb, 1, 0, want_outputs=True)
        cpu_batch = {"value": vb, "unit": "audio-s/s", "cores": threads, "kind": "port",
                     "sample": "batch %d (what `--impl reference` times per step): %d x 15 s utterances, 1 timed "
                               "repetition" % (nb, nb)}
        # parity of the benchmarked mode on the benchmarked utterances: the oracle's greedy transcripts / frames of
        # utterances 0..nb-1 against what the timed path produced for them
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import parity_util as pu
        parity = {"mode": args.precision, "against": "oracle (CPU restatement of the reference path) on utterances 0..%d "
                  "of the benchmarked batch" % (nb - 1)}
        parity.update(pu.transcript_report(texts[:nb], ref_texts))
        pb = extras.get("probs_bench")
        if pb is not None:
            fr = pu.merge_reports([pu.frame_report(pb[i], ref_probs[i]) for i in range(nb)])
            parity.update({"frames": fr["frames"], "frames_differ": fr["frames_differ"],
                           "frame_agreement": 1.0 - fr["frames_differ"] / fr["frames"],
                           "max_flipped_margin": fr["max_flipped_margin"]})
    extras.pop("probs_bench", None)
    if not args.no_extras:
        extras["library_bar"] = bench_library_bar(dev)

    line = {
        "metric": "audio-sec/sec (RTFx)", "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "precision_mode": args.precision, "utterances_per_gpu": BATCH,
                   "batches_in_flight": merge,
                   "step": "one batch of 64 x 15 s; up to %d consecutive batches share one pass of the model "
                           "(the CTA-pair recurrence multiplies two groups of 64 sequences per MMA, two such items in flight), K steps = %s passes"
                           % (merge, "+".join(str(m) for m in passes_for(args.steps, merge))),
                   "l2": "working set per step (weights >= 325 MB + activations > 1 GB) exceeds the 126 MB L2",
                   "sharding": "independent utterance batches per rank, no collective on the data path"},
        "e2e": {"value": e2e_value, "unit": "audio-s/s", "h2d_bytes_per_step": BATCH * n * 4,
                "d2h_bytes_per_step": d2h, "transcripts": n_texts,
                "input": "lists of numpy float64 arrays (the API's input type); float64->float32 staging inside the timed region",
                "pinned_f32_value": audio_s * args.steps / e2e_pinned_s},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "stages": stages,
        "single_batch": single, "cpu_baseline": cpu, "cpu_baseline_batch": cpu_batch, "parity": parity, "beam": beam,
    }
    line.update(extras)
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def relaunch_under_torchrun(args):
    """`python bench.py --gpus N` with N > 1 and no launcher environment: start the N ranks ourselves, the way the
    driver does (one process per GPU, rendezvous on 127.0.0.1)."""
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
    os.execv(sys.executable, cmd)


def main():
    args = parse_args()
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        relaunch_under_torchrun(args)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
