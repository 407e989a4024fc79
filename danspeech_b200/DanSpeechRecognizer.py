"""Inference engine: wires parser, acoustic model and decoder together.

Mirrors danspeech/DanSpeechRecognizer.py (class DanSpeechRecognizer :12-231): same constructor,
attributes, ``update_model`` / ``update_decoder`` / ``enable_streaming`` / ``streaming_transcribe`` /
``transcribe`` semantics.  Every stage runs on the GPU; ``with_gpu=False`` raises because this
framework has no CPU path.  ``transcribe_batch`` is an addition (the reference engine is batch-1).
"""
import warnings

import torch

from . import _native as N
from .audio.parsers import InferenceSpectrogramAudioParser, SpectrogramAudioParser
from .deepspeech.decoder import BeamCTCDecoder, GreedyDecoder
from .errors.recognizer_errors import ModelNotInitialized
from .utils.stitch import stitch_transcript


class NoLmInstantiatedWarning(Warning):
    pass


class DanSpeechRecognizer(object):

    def __init__(self, model_name=None, lm_name=None, alpha=1.3, beta=0.2, with_gpu=True, beam_width=64,
                 device=None):
        if not with_gpu:
            raise N.NativeError("danspeech_b200 runs on a B200 GPU only (with_gpu=False has no implementation)")
        N.require_cuda()
        self.device = torch.device(device or "cuda")
        print("Using device: {0}".format(self.device))

        self.model = None
        self.model_name = None
        self.labels = None
        self.audio_config = None
        self.audio_parser = None
        self.lm = None
        self.decoder = None
        self.alpha = alpha
        self.beta = beta
        self.beam_width = beam_width
        self.secondary_model = None
        self._copy_streams = {}     # device -> side stream of transcribe_batches
        self._decode_streams = {}   # device -> decode stream of transcribe_batches
        if model_name:
            self.update_model(model_name)
        if lm_name:
            if not self.model:
                raise ModelNotInitialized("Trying to initialize LM without also choosing a DanSpeech model.")
            self.update_decoder(lm_name)

    def update_model(self, model):
        self.audio_config = model.audio_conf
        self.model = model.to(self.device)
        self.model.eval()
        self.audio_parser = SpectrogramAudioParser(self.audio_config, device=self.device,
                                                   fast_fft=getattr(self.model, "precision", "fp32") == "bf16")
        self.labels = self.model.labels
        # as DanSpeechRecognizer.py:54-56 (quirk Q4: labels are assigned first, so an existing decoder is
        # only rebuilt when lm/alpha/beta/beam_width change)
        self.update_decoder(labels=self.labels)

    def update_decoder(self, lm=None, alpha=None, beta=None, labels=None, beam_width=None):
        """Update rules of DanSpeechRecognizer.py:58-95: a falsy argument means "keep"; the decoder is rebuilt only when
        one of lm / alpha / beta / labels / beam_width really changes, or when there is none yet (then lm = "greedy")."""
        rebuild = not self.lm and not self.decoder
        if rebuild:
            self.lm = "greedy"
        for field, value in (("lm", lm), ("alpha", alpha), ("beta", beta), ("labels", labels),
                             ("beam_width", beam_width)):
            if value and getattr(self, field) != value:
                setattr(self, field, value)
                rebuild = True
        if rebuild:
            self.decoder = self._make_decoder()

    def _make_decoder(self):
        blank = self.labels.index("_")
        if self.lm == "greedy":
            return GreedyDecoder(labels=self.labels, blank_index=blank)
        # constructor arguments as DanSpeechRecognizer.py:89-92
        return BeamCTCDecoder(labels=self.labels, lm_path=self.lm, alpha=self.alpha, beta=self.beta,
                              beam_width=self.beam_width, num_processes=6, cutoff_prob=1.0, cutoff_top_n=40,
                              blank_index=blank)

    # ------------------------------------------------------------------ streaming (DanSpeechRecognizer.py:98-216)
    def enable_streaming(self, secondary_model=None, return_string_parts=True):
        self.full_output = []
        self.iterating_transcript = ""
        if secondary_model:
            self.secondary_model = secondary_model.to(self.device)
            self.secondary_model.eval()
        else:
            self.secondary_model = None
        self.spectrograms = []
        self.greedy_decoder = GreedyDecoder(labels=self.labels, blank_index=self.labels.index("_"))
        self.audio_parser = InferenceSpectrogramAudioParser(audio_config=self.audio_config, device=self.device)
        self.string_parts = bool(return_string_parts)

    def disable_streaming(self, keep_secondary_model=False):
        self.audio_parser = SpectrogramAudioParser(self.audio_config, device=self.device,
                                                   fast_fft=getattr(self.model, "precision", "fp32") == "bf16")
        self.greedy_decoder = None
        self.reset_streaming_params()
        self.string_parts = False
        if not keep_secondary_model:
            self.secondary_model = None

    def reset_streaming_params(self):
        self.iterating_transcript = ""
        self.full_output = []
        self.spectrograms = []

    def streaming_transcribe(self, recording, is_last, is_first):
        """One chunk of one stream (DanSpeechRecognizer.py:144-216): returns the new string part (or the iterating
        transcript when string parts are off), "" for the first chunk, and on the last chunk the final transcript --
        re-decoded by the secondary model or the LM decoder when there is one."""
        spect = self.audio_parser.parse_audio(recording, is_last)
        out = ""
        if len(spect) != 0:
            if self.secondary_model:
                self.spectrograms.append(spect)
            probs = self.model(spect.view(1, 1, spect.size(0), spect.size(1)), is_first, is_last)
            if is_first:
                return ""
            self.full_output.append(probs)
            decoded, _ = self.greedy_decoder.decode(probs)
            self.iterating_transcript, part = stitch_transcript(self.iterating_transcript, decoded[0][0])
            out = part if self.string_parts else self.iterating_transcript
        return self._finish_stream() if is_last else out

    def _finish_stream(self):
        heard = len(self.iterating_transcript) > 1
        final = ""
        if heard and self.secondary_model:
            full = torch.cat(self.spectrograms, dim=1)
            probs, _ = self.secondary_model(full.view(1, 1, full.size(0), full.size(1)), torch.IntTensor([full.size(1)]))
            final = self.decoder.decode(probs)[0][0][0]
        elif heard and self.lm != "greedy":
            final = self.decoder.decode(torch.cat(self.full_output, dim=1))[0][0][0]
        elif heard:
            final = self.iterating_transcript
        if heard:
            self.reset_streaming_params()
        return final

    # ------------------------------------------------------------------ offline (DanSpeechRecognizer.py:218-231)
    def transcribe(self, recording, show_all=False):
        if hasattr(self.audio_parser, "parse_batch"):
            spect, input_sizes = self.audio_parser.parse_batch([recording])
        else:
            # real-time streaming is enabled: like the reference (DanSpeechRecognizer.py:218-222) the recording goes
            # through whatever parser is installed, here the streaming one with its adaptive normalisation
            s = self.audio_parser.parse_audio(recording)
            spect, input_sizes = s.view(1, 1, s.size(0), s.size(1)), torch.IntTensor([s.size(1)])
        out, output_sizes = self.model(spect, input_sizes)
        decoded_output, _ = self.decoder.decode(out, output_sizes)
        if show_all:
            if self.lm == "greedy":
                warnings.warn("You are trying to get all beams but no LM has been instantiated.",
                              NoLmInstantiatedWarning)
            return decoded_output[0]
        return decoded_output[0][0]

    def transcribe_batch(self, recordings, show_all=False):
        """Batched ``transcribe``: sorts by length (pack_padded_sequence contract), restores input order."""
        if isinstance(recordings, tuple):
            # (host float32 tensor [B, stride] (ideally pinned), n_samples) already sorted by length descending
            host_audio, n_samples = recordings
            order = list(range(len(n_samples)))
        else:
            order = sorted(range(len(recordings)), key=lambda i: -len(recordings[i]))
            host_audio, n_samples = self.audio_parser.stage_batch([recordings[i] for i in order])
        return self._transcribe_staged(host_audio, n_samples, order, show_all)

    def _transcribe_staged(self, host_audio, n_samples, order, show_all=False):
        spect, input_sizes = self.audio_parser.parse_packed(host_audio, n_samples)
        out, output_sizes = self.model(spect, input_sizes)
        decoded_output, _ = self.decoder.decode_finish(self.decoder.decode_device(out, output_sizes), out,
                                                       top_only=not show_all)
        results = [None] * len(order)
        for pos, i in enumerate(order):
            results[i] = decoded_output[pos] if show_all else decoded_output[pos][0]
        return results

    # Batches that transcribe_batches runs through ONE pass of the model: the CTA-pair recurrence (csrc/rnn_pair.cu)
    # multiplies two groups of 64 sequences per MMA and keeps two such items in flight per CTA pair, so four batches of
    # 64 cost far less than four passes (the one-CTA kernel, csrc/rnn_tc.cu, takes a single batch).
    batches_in_flight = 4
    max_merged_rows = 256
    max_merged_samples = 256 * 30 * 16000     # bound on rows x longest recording of a merged pass (workspace size)

    def _merge_plan(self, batches, merge):
        """Consecutive batches -> passes of at most `merge` batches within the row / sample budgets."""
        plan, cur, rows, longest = [], [], 0, 0
        for k, b in enumerate(batches):
            if isinstance(b, tuple):
                n, ln = len(b[1]), max(int(v) for v in b[1])
            else:
                n, ln = len(b), max(len(r) for r in b)
            fits = cur and len(cur) < merge and rows + n <= self.max_merged_rows and \
                (rows + n) * max(longest, ln) <= self.max_merged_samples and \
                isinstance(b, tuple) == isinstance(batches[cur[0]], tuple)
            if not fits and cur:
                plan.append(cur)
                cur, rows, longest = [], 0, 0
            cur.append(k)
            rows += n
            longest = max(longest, ln)
        if cur:
            plan.append(cur)
        return plan

    def transcribe_batches(self, batches, show_all=False, merge=None):
        """Several batches back to back.  Up to ``merge`` (default ``batches_in_flight``) consecutive batches go through
        the model as ONE pass (their sequences sorted by length together), and a helper thread converts the next pass
        into one of two cached pinned buffers and starts its host->device copy on a side stream while the GPU works on
        the current one, so neither the float64->float32 conversion nor the PCIe copy sits between two passes.  A batch
        is a list of recordings or a (pinned f32 tensor [B, stride], n_samples) tuple sorted by length descending.
        Returns one result list per batch, in input order."""
        import concurrent.futures

        dev = torch.device(self.device if isinstance(self.device, (str, torch.device)) else "cuda")
        if dev.type != "cuda":
            dev = torch.device("cuda")
        # one side stream per recognizer and device: the caching allocator keeps a pool per stream, so a fresh stream
        # per call would pay a cudaMalloc (a device-wide synchronisation) for the first batches of every call
        copy_stream = self._copy_streams.get(str(dev))
        if copy_stream is None:
            copy_stream = self._copy_streams[str(dev)] = torch.cuda.Stream(device=dev)
        plan = self._merge_plan(batches, max(1, int(merge or self.batches_in_flight)))

        def stage(j):
            members = plan[j]
            sizes = [len(batches[k][1]) if isinstance(batches[k], tuple) else len(batches[k]) for k in members]
            if isinstance(batches[members[0]], tuple):
                # pinned tensors, each already sorted: copy them side by side, then order the rows on the device
                ns = [int(v) for k in members for v in batches[k][1]]
                stride = max(batches[k][0].shape[1] for k in members)
                order = sorted(range(len(ns)), key=lambda i: -ns[i])          # stable: equal lengths keep their order
                with torch.cuda.device(dev), torch.cuda.stream(copy_stream):
                    if len(members) == 1:
                        audio = batches[members[0]][0].to(dev, non_blocking=True)
                    else:
                        audio = torch.empty((len(ns), stride), dtype=torch.float32, device=dev)
                        row = 0
                        for k in members:
                            h = batches[k][0]
                            audio[row:row + h.shape[0], : h.shape[1]].copy_(h, non_blocking=True)
                            row += h.shape[0]
                    if order != list(range(len(ns))):
                        audio = audio.index_select(0, torch.tensor(order, dtype=torch.int64).to(dev, non_blocking=True))
                    ns = [ns[i] for i in order]
                    n_dev = torch.tensor(ns, dtype=torch.int32).to(dev, non_blocking=True)
                    done = torch.cuda.Event()
                    done.record(copy_stream)
                for k in members:
                    self.audio_parser.mark_staging_busy(batches[k][0], done)
            else:
                recs = [r for k in members for r in batches[k]]
                order = sorted(range(len(recs)), key=lambda i: -len(recs[i]))
                host, ns = self.audio_parser.stage_batch([recs[i] for i in order], slot=j & 1)
                with torch.cuda.device(dev), torch.cuda.stream(copy_stream):
                    audio = host.to(dev, non_blocking=True)
                    n_dev = torch.tensor(ns, dtype=torch.int32).to(dev, non_blocking=True)
                    done = torch.cuda.Event()
                    done.record(copy_stream)
                self.audio_parser.mark_staging_busy(host, done)
            return audio, n_dev, ns, order, sizes, done

        # Decode runs on its own stream: the beam search (or the greedy kernel) of pass j and the host's string building
        # overlap the forward pass j+1 -- dsb_forward does not synchronise, so the model of the next pass is enqueued
        # before the results of this one are read back.
        decode_stream = self._decode_streams.get(str(dev))
        if decode_stream is None:
            decode_stream = self._decode_streams[str(dev)] = torch.cuda.Stream(device=dev)
        results = []

        def finish(pending):
            dev_out, probs, order, sizes = pending
            with torch.cuda.device(dev), torch.cuda.stream(decode_stream):
                decoded_output, _ = self.decoder.decode_finish(dev_out, probs, top_only=not show_all)
            flat = [None] * len(order)
            for pos, i in enumerate(order):
                flat[i] = decoded_output[pos] if show_all else decoded_output[pos][0]
            row = 0
            for n in sizes:
                results.append(flat[row:row + n])
                row += n

        with concurrent.futures.ThreadPoolExecutor(max_workers=1) as ex:
            nxt = ex.submit(stage, 0) if plan else None
            pending = None
            for j in range(len(plan)):
                audio, n_dev, ns, order, sizes, done = nxt.result()
                nxt = ex.submit(stage, j + 1) if j + 1 < len(plan) else None
                main = torch.cuda.current_stream(dev)
                main.wait_event(done)
                audio.record_stream(main)
                n_dev.record_stream(main)
                spect, _ = self.audio_parser.parse_device(audio, n_dev, max(ns))
                input_sizes = torch.IntTensor([1 + n // self.audio_parser.hop_length for n in ns])
                out, output_sizes = self.model(spect.view(len(ns), 1, 161, spect.shape[2]), input_sizes)
                fwd_done = torch.cuda.Event()
                fwd_done.record(main)
                with torch.cuda.device(dev), torch.cuda.stream(decode_stream):
                    decode_stream.wait_event(fwd_done)
                    out.record_stream(decode_stream)
                    argmax = getattr(out, "_dsb_argmax", None)
                    if argmax is not None:
                        argmax.record_stream(decode_stream)
                    dev_out = self.decoder.decode_device(out, output_sizes)
                if pending is not None:
                    finish(pending)            # pass j-1: read back and build strings while pass j runs
                pending = (dev_out, out, order, sizes)
            if pending is not None:
                finish(pending)
        return results
