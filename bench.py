#!/usr/bin/env python
"""Benchmark of the DanSpeech inference hot path (audio -> spectrogram -> DeepSpeech2 -> CTC decode).

Contract: ``python bench.py --gpus N --steps K --warmup W`` prints ONE JSON line (rank 0).  Under
torchrun (N > 1) every rank processes its own batch of 64 utterances (utterance sharding, no
collective on the data path; "scaling": "weak"); without a launcher environment, ``--gpus N`` with N > 1
starts the N ranks itself the same way (``torch.distributed.run``, rendezvous on 127.0.0.1).

Workload (BASELINE.json configs[1]): DanSpeechPrimary-shaped bi-GRU DeepSpeech2 (3 conv + 9 x 1200
bi-GRU, random-init), batch 64 x 15 s synthetic 16 kHz audio, greedy decode.  Metric: audio-seconds
per wall-second (RTFx).
  value : device-timed (CUDA events), audio already resident in HBM
  e2e   : through Recognizer.recognize_batches with HOST (pinned) audio: every step's H2D copy, kernels, D2H of
          the token/offset tensors and transcript string building inside the timed region (the copy of step k+1
          runs on a side stream under the kernels of step k)
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


def cpu_reference_rtfx(n_utts, reps, warmup, threads=None):
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
        probs, sizes = om.forward(sd, x, lens, cfg["conv_layers"], cfg["rnn_layers"])
        return og.greedy_decode(probs.numpy(), sizes.numpy())

    for _ in range(warmup):
        step()
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    sec = float(np.mean(times))
    return n_utts * SECONDS / sec, sec, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # bounded sample: a 15 s utterance costs the host ~1.3 s, so size the step for ~150 s over steps + warm-up
    nb = args.ref_batch or max(1, min(8, int(150.0 / (max(1, args.steps) + 1) / 1.3)))
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
    audio_dev = host.to(dev)
    n_dev = torch.full((BATCH,), n, dtype=torch.int32, device=dev)
    lengths = torch.IntTensor([1 + n // 160] * BATCH)
    spect_buf = torch.empty((BATCH, 161, 1 + n // 160), dtype=torch.float32, device=dev)

    def step_device():
        spect, _ = parser.parse_device(audio_dev, n_dev, n, out=spect_buf)
        probs, sizes = eng.model(spect.view(BATCH, 1, 161, -1), lengths)
        return decoder.decode_device(probs, sizes)

    def step_e2e():
        return rec.recognize_batch((host, [n] * BATCH))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        out = step_device()
    torch.cuda.synchronize()

    # ---- device-timed region: exactly K steps, CUDA events, max over ranks ----
    L.dsb_profile_reset()
    L.dsb_profile_enable(1)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    launches0 = N.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        out = step_device()
    ev1.record()
    barrier()
    launches = N.launch_count() - launches0
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    prof = N.profile_read()
    L.dsb_profile_enable(0)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    audio_s = BATCH * SECONDS * world
    value = audio_s * args.steps / (ms_max / 1e3)

    # ---- end-to-end through the public API with host buffers ----
    # every step copies its batch from pinned host memory and reads its transcripts back; Recognizer.recognize_batches
    # runs the steps back to back with the copy of step k+1 (side stream) overlapping the kernels of step k
    texts = step_e2e()
    rec.recognize_batches([(host, [n] * BATCH)] * max(args.warmup, 3))
    barrier()
    t0 = time.perf_counter()
    texts = rec.recognize_batches([(host, [n] * BATCH)] * args.steps)[-1]
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        gathered = [None] * world
        dist.all_gather_object(gathered, texts)     # final host gather of transcripts (off the data path)
        n_texts = sum(len(x) for x in gathered)
    else:
        n_texts = len(texts)
    e2e_value = audio_s * args.steps / float(te.item())

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
        spect, _ = parser.parse_device(audio_dev, n_dev, n, out=spect_buf)
        probs, sizes = eng.model(spect.view(BATCH, 1, 161, -1), lengths)
        bdec.decode_device(probs, sizes)
        barrier()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record()
        for _ in range(args.steps):
            bdec.decode_device(probs, sizes)
        b1.record()
        barrier()
        tb = torch.tensor([b0.elapsed_time(b1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tb, op=dist.ReduceOp.MAX)
        beam_ms = float(tb.item()) / args.steps
        beam = {"utt_per_s": BATCH * world / (beam_ms / 1e3), "ms_per_batch": beam_ms, "beam_width": 64,
                "lm": "synthetic 3-gram ARPA, 2000 words, alpha 1.3, beta 0.2",
                "rtfx_forward_plus_beam": audio_s / ((ms_max / args.steps + beam_ms) / 1e3)}
    Tp = (1 + n // 160 - 1) // 2 + 1
    d2h = BATCH * (1 + 2 * Tp) * 4

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel family (RNN recurrence), from the live stage timers ----
    peaks, peak_src = measured_peaks()
    conv_f, proj_f, rec_f = model_flops(BATCH, Tp)
    stage_ms = {k: v[0] / max(args.steps, 1) for k, v in prof.items()}
    dominant = max(("conv", "rnn_input_proj", "rnn_recurrence"), key=lambda k: stage_ms[k])
    flops = {"conv": conv_f, "rnn_input_proj": proj_f, "rnn_recurrence": rec_f}[dominant]
    achieved = flops / (stage_ms[dominant] / 1e3) / 1e12 if stage_ms[dominant] > 0 else 0.0
    peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1400.0)))
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tp) and args.precision == "bf16":
        traffic = json.load(open(tp)).get(dominant, {}).get("bytes_per_launch")
    roofline = {"bound": "tensor", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src + ", sustained bf16 (kernel timed inside a long step)",
                "ms_per_step": stage_ms[dominant], "launches_per_step": prof[dominant][1] / max(args.steps, 1)}
    spect_bytes = BATCH * (4 * n + 4 * 161 * (1 + n // 160))
    stages = {k: {"ms_per_step": round(v, 4)} for k, v in stage_ms.items() if v > 0}
    if stage_ms["spectrogram"] > 0:
        gbs = spect_bytes / (stage_ms["spectrogram"] / 1e3) / 1e9
        stages["spectrogram"].update({"GB/s": round(gbs, 1), "frac_hbm": round(gbs / float(peaks["hbm_gbs"]), 4)})
    for k, f in (("conv", conv_f), ("rnn_input_proj", proj_f), ("rnn_recurrence", rec_f)):
        if stage_ms[k] > 0:
            tf = f / (stage_ms[k] / 1e3) / 1e12
            stages[k].update({"TFLOP/s": round(tf, 2), "frac_tensor": round(tf / peak, 4)})

    cpu = None
    if not args.no_cpu_baseline:
        v, sec, threads = cpu_reference_rtfx(args.cpu_sample, 2, 1)
        cpu = {"value": v, "unit": "audio-s/s", "cores": threads, "kind": "port", "cpu_model": cpu_model_name(),
               "sample": "%d x 15 s utterance(s) per repetition, 1 warm-up + 2 timed, oracle CPU restatement "
                         "(torch CPU fp32, %d threads)" % (args.cpu_sample, threads)}

    line = {
        "metric": "audio-sec/sec (RTFx)", "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "precision_mode": args.precision, "utterances_per_gpu": BATCH,
                   "l2": "working set per step (weights >= 325 MB + activations > 1 GB) exceeds the 126 MB L2",
                   "sharding": "independent utterance batches per rank, no collective on the data path"},
        "e2e": {"value": e2e_value, "unit": "audio-s/s", "h2d_bytes_per_step": BATCH * n * 4,
                "d2h_bytes_per_step": d2h, "transcripts": n_texts},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "stages": stages,
        "cpu_baseline": cpu, "beam": beam,
    }
    print(json.dumps(line))
    if world > 1:
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
