"""CTC decoders backed by the danspeech_b200 CUDA library.

Interfaces mirror danspeech/deepspeech/decoder.py: Decoder (:24-88), BeamCTCDecoder (:91-144),
GreedyDecoder (:147-198).  ``wer``/``cer`` (decoder.py:45-74) are scoring utilities outside the
inference path and are implemented here without the Levenshtein C dependency.
"""
import numpy as np
import torch

from .. import _native as N


def _edit_distance(a, b):
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return prev[-1]


class Decoder(object):
    """Base class (reference: decoder.py:24-88)."""

    def __init__(self, labels, blank_index=0):
        self.labels = labels
        self.int_to_char = dict([(i, c) for (i, c) in enumerate(labels)])
        self.blank_index = blank_index
        space_index = len(labels)   # out-of-range sentinel when the label set has no space (decoder.py:40-43)
        if " " in labels:
            space_index = labels.index(" ")
        self.space_index = space_index

    def wer(self, s1, s2):
        return _edit_distance(s1.split(), s2.split())

    def cer(self, s1, s2):
        return _edit_distance(s1.replace(" ", ""), s2.replace(" ", ""))

    def decode(self, probs, sizes=None):
        raise NotImplementedError


def _after_sync(probs):
    """Called once the host has synchronised for a decode result: the forward that produced `probs` has completed, so
    its device-side status can be read without waiting (DeepSpeech.check_status)."""
    chk = getattr(probs, "_dsb_status", None)
    if chk is not None:
        chk()


def _sizes_to_device(sizes, B, dev):
    if sizes is None:
        return None
    if not torch.is_tensor(sizes):
        sizes = torch.tensor([int(s) for s in sizes])
    return sizes.to(device=dev, dtype=torch.int32).contiguous().view(B)


class GreedyDecoder(Decoder):
    """argmax -> collapse repeats -> drop blanks, on the GPU (reference: decoder.py:147-198)."""

    def __init__(self, labels, blank_index=0):
        super().__init__(labels, blank_index)

    # ---- host-side helpers of the reference's interface, for callers that already hold index sequences
    def process_string(self, sequence, size, remove_repetitions=False):
        """Index sequence -> (string, IntTensor of the frame of every emitted character): blanks are dropped and,
        with ``remove_repetitions``, a symbol equal to the previous FRAME's symbol (decoder.py:166-181)."""
        ids = sequence.detach().cpu().numpy() if torch.is_tensor(sequence) else np.asarray(sequence)
        ids = ids.reshape(-1)[:int(size)].astype(np.int64)
        keep = ids != self.blank_index
        if remove_repetitions and ids.size > 1:
            keep[1:] &= ids[1:] != ids[:-1]
        frames = np.nonzero(keep)[0]
        text = "".join(self.int_to_char[i] for i in ids[frames].tolist())
        return text, torch.from_numpy(frames.astype(np.int32))

    def convert_to_strings(self, sequences, sizes=None, remove_repetitions=False, return_offsets=False):
        """Batch form of ``process_string``: List[B] of [string] (one path), plus List[B] of [offsets] on request
        (decoder.py:151-164)."""
        strings, offsets = [], []
        for b in range(len(sequences)):
            n = sizes[b] if sizes is not None else len(sequences[b])
            text, frames = self.process_string(sequences[b], n, remove_repetitions)
            strings.append([text])
            offsets.append([frames])
        return (strings, offsets) if return_offsets else strings

    def decode_device(self, probs, sizes=None):
        """Returns device tensors (tokens[B,T], offsets[B,T], out_len[B]) without synchronising."""
        N.require_cuda()
        if not probs.is_cuda:
            probs = probs.cuda()
        argmax = getattr(probs, "_dsb_argmax", None)
        B, T, C = probs.shape
        dev = probs.device
        probs_c = probs if probs.is_contiguous() else None
        if argmax is not None:
            argmax = argmax.contiguous()
        elif probs_c is None:
            probs_c = probs.contiguous()
        with torch.cuda.device(dev):
            tokens = torch.empty((B, max(T, 1)), dtype=torch.int32, device=dev)
            offsets = torch.empty((B, max(T, 1)), dtype=torch.int32, device=dev)
            out_len = torch.empty((B,), dtype=torch.int32, device=dev)
            d_sizes = _sizes_to_device(sizes, B, dev)
            N.check(N.lib().dsb_greedy_decode(N.ptr(probs_c) if argmax is None else None, N.ptr(argmax),
                                              N.ptr(d_sizes), B, T, C, int(self.blank_index), N.ptr(tokens),
                                              N.ptr(offsets), N.ptr(out_len), N.current_stream()),
                    "dsb_greedy_decode")
        return tokens, offsets, out_len

    def decode_strings(self, probs, sizes=None):
        """Transcripts only (List[B] of str): the serving loops need no offsets, and building B offset tensors
        costs more host time than the kernel."""
        tokens, _, out_len = self.decode_device(probs, sizes)
        packed = torch.cat([out_len.view(-1, 1), tokens], dim=1).cpu().numpy()
        _after_sync(probs)
        if not hasattr(self, "_char_arr"):
            self._char_arr = np.array([self.int_to_char[i] for i in range(len(self.int_to_char))])
        chars = self._char_arr[np.clip(packed[:, 1:], 0, len(self._char_arr) - 1)]   # entries past out_len are unwritten
        return ["".join(chars[b, :n]) for b, n in enumerate(packed[:, 0].tolist())]

    def decode(self, probs, sizes=None):
        """Returns (strings: List[B][1] str, offsets: List[B][1] IntTensor) -- decoder.py:183-198."""
        return self.decode_finish(self.decode_device(probs, sizes), probs)

    def decode_finish(self, dev_out, probs=None, top_only=False):
        """Host half of ``decode``: one D2H copy of what ``decode_device`` produced, then the strings."""
        tokens, offsets, out_len = dev_out
        packed = torch.cat([out_len.view(-1, 1), tokens, offsets], dim=1).cpu().numpy()   # one D2H copy
        _after_sync(probs)
        B, T = tokens.shape
        chars = self.int_to_char
        lens = packed[:, 0].tolist()
        offs_all = torch.from_numpy(packed[:, 1 + T:].copy())
        strings, offs = [], []
        for b, n in enumerate(lens):
            strings.append(["".join([chars[i] for i in packed[b, 1:1 + n].tolist()])])
            offs.append([offs_all[b, :n]])
        return strings, offs


class BeamCTCDecoder(Decoder):
    """Prefix beam search with n-gram LM scoring on the GPU.

    Constructor signature as decoder.py:92-93; replaces the ctcdecode.CTCBeamDecoder object built at
    decoder.py:99-100.  ``lm_path`` is an ARPA text file (KenLM binaries are a "next" row).
    """

    def __init__(self, labels, lm_path=None, alpha=0, beta=0, cutoff_top_n=40, cutoff_prob=1.0, beam_width=100,
                 num_processes=4, blank_index=0):
        super().__init__(labels)   # reference quirk Q7: the python-side blank_index stays 0
        self._beam_width = int(beam_width)
        self._blank_id = int(blank_index)
        self._num_processes = num_processes   # accepted for compatibility; the GPU decoder has no host pool
        L = N.lib()
        blob = b"\0".join(c.encode("utf-8") for c in labels) + b"\0"
        h = N.c_void_p()
        N.check(L.dsb_beam_create(blob, len(labels), lm_path.encode() if lm_path else None, float(alpha), float(beta),
                                  int(cutoff_top_n), float(cutoff_prob), int(beam_width), int(blank_index), 0, h),
                "dsb_beam_create")
        self._handle = h
        self._ws = None
        self.last_scores = None

    def __del__(self):
        try:
            if getattr(self, "_handle", None):
                N.lib().dsb_beam_destroy(self._handle)
                self._handle = None
        except Exception:
            pass

    def decode_device(self, probs, sizes=None):
        N.require_cuda()
        if not probs.is_cuda:
            probs = probs.cuda()
        probs = probs.float().contiguous()
        B, T, C = probs.shape
        dev = probs.device
        if sizes is None:
            lens = [T] * B
        else:
            lens = [int(v) for v in (sizes.tolist() if torch.is_tensor(sizes) else sizes)]
        L = N.lib()
        W = self._beam_width
        with torch.cuda.device(dev):
            out = torch.zeros((B, W, T), dtype=torch.int32, device=dev)
            ts = torch.zeros((B, W, T), dtype=torch.int32, device=dev)
            scores = torch.zeros((B, W), dtype=torch.float32, device=dev)
            out_len = torch.zeros((B, W), dtype=torch.int32, device=dev)
            need = L.dsb_beam_workspace_bytes(self._handle, B, T)
            if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
                self._ws = torch.empty(int(need), dtype=torch.uint8, device=dev)
            N.check(L.dsb_beam_decode(self._handle, N.ptr(probs), N.i32_array(lens), B, T, C, N.ptr(out), N.ptr(ts),
                                      N.ptr(scores), N.ptr(out_len), N.ptr(self._ws), self._ws.numel(),
                                      N.current_stream()), "dsb_beam_decode")
        return out, scores, ts, out_len

    def convert_to_strings(self, out, seq_len):
        results = []
        for b in range(out.shape[0]):
            utterances = []
            for p in range(out.shape[1]):
                size = int(seq_len[b][p])
                utterances.append("".join(self.int_to_char[i] for i in out[b, p, :size].tolist()) if size > 0 else "")
            results.append(utterances)
        return results

    def convert_tensor(self, offsets, sizes):
        results = []
        for b in range(offsets.shape[0]):
            utterances = []
            for p in range(offsets.shape[1]):
                size = int(sizes[b][p])
                utterances.append(offsets[b, p, :size] if size > 0 else torch.tensor([], dtype=torch.int))
            results.append(utterances)
        return results

    def decode(self, probs, sizes=None):
        """Returns (strings: List[B][beam] str, offsets: List[B][beam] IntTensor) -- decoder.py:129-144."""
        return self.decode_finish(self.decode_device(probs, sizes), probs)

    def decode_finish(self, dev_out, probs=None, top_only=False):
        """Host half of ``decode``: D2H copies of what ``decode_device`` produced, then strings and offsets.
        ``top_only`` reads back and converts the best beam only (what ``transcribe`` without ``show_all`` keeps,
        DanSpeechRecognizer.py:225-231): the full [B, beam, T] token and timestep tensors are 2 x 12 MB per batch of
        64 x 751 frames and 4096 Python string conversions."""
        out, scores, ts, out_len = dev_out
        if top_only:
            out, ts, out_len = out[:, :1].contiguous(), ts[:, :1].contiguous(), out_len[:, :1].contiguous()
        out, scores, ts, out_len = out.cpu(), scores.cpu(), ts.cpu(), out_len.cpu()
        _after_sync(probs)
        self.last_scores = scores
        return self.convert_to_strings(out, out_len), self.convert_tensor(ts, out_len)
