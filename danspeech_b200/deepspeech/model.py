"""DeepSpeech2 acoustic model whose forward runs in the danspeech_b200 CUDA library.

Interface and state-dict layout mirror danspeech/deepspeech/model.py (class DeepSpeech :287-666;
parameter names/shapes as listed in SURVEY.md A.6) so that a reference ``.pth`` package loads
unchanged.  The torch modules below are *containers for parameters only*: ``forward`` never calls
them -- it hands raw device pointers to dsb_forward / dsb_streaming_forward (C ABI), which run the
hand-written sm_100a kernels.  There is no CPU or eager-PyTorch fallback.
"""
import os
from collections import OrderedDict

import torch
import torch.nn as nn

from .. import _native as N
from ..errors.model_errors import ConvError
from .utils import get_default_audio_config

supported_rnns = {"lstm": nn.LSTM, "rnn": nn.RNN, "gru": nn.GRU}
supported_rnns_inv = dict((v, k) for k, v in supported_rnns.items())

_CONV_SPECS = [  # (cin, cout, (kh, kw), stride, padding)   reference: model.py:357-392
    (1, 32, (41, 11), (2, 2), (20, 5)),
    (32, 32, (21, 11), (2, 1), (10, 5)),
    (32, 96, (21, 11), (2, 1), (10, 5)),
]


class _Holder(nn.Module):
    """Named parameter container; calling it is a bug (the CUDA library does the arithmetic)."""

    def forward(self, *a, **k):
        raise RuntimeError("danspeech_b200 parameter container modules are not callable")


def _wrap(name, module):
    h = _Holder()
    h.add_module(name, module)
    return h


def _default_labels():
    """The default symbol inventory of the reference models (its deepspeech/labels.json, read at model.py:321-323):
    CTC blank, a-z, the Danish letters and two accented ones, space.  Index = column of the fc output."""
    import string
    return "_" + string.ascii_lowercase + "\u00e6\u00f8\u00e5\u00e9\u00fc" + " "


class DeepSpeech(nn.Module):
    """Drop-in for ``danspeech.deepspeech.model.DeepSpeech`` (inference only)."""

    def __init__(self, model_name, rnn_type=nn.GRU, labels=None, rnn_hidden_size=768, rnn_layers=5, audio_conf=None,
                 bidirectional=True, context=20, conv_layers=2, streaming_inference_model=False):
        super().__init__()
        if not labels:
            labels = _default_labels()
        if audio_conf is None:
            audio_conf = get_default_audio_config()
        self.model_name = model_name
        self.rnn_hidden_size = rnn_hidden_size
        self.rnn_layers = rnn_layers
        self.rnn_type = rnn_type
        self.audio_conf = audio_conf or {}
        self.labels = labels
        self.bidirectional = bidirectional
        self.conv_layers = conv_layers
        self.streaming_model = streaming_inference_model
        self.context = context
        self.precision = os.environ.get("DANSPEECH_B200_PRECISION", "fp32")

        if conv_layers == 0:
            raise ConvError("0 convolutional layers configuration not supported by DanSpeech")
        if conv_layers > 3:
            raise ConvError("Maximum amount of convolutional layers supported by DanSpeech is 3")
        if rnn_type not in supported_rnns_inv:
            raise ValueError("rnn_type must be one of nn.GRU, nn.LSTM, nn.RNN")

        # conv.seq_module.{0,3,6} Conv2d, {1,4,7} BatchNorm2d, {2,5,8} Hardtanh
        mods = []
        for cin, cout, k, s, p in _CONV_SPECS[:conv_layers]:
            mods += [nn.Conv2d(cin, cout, kernel_size=k, stride=s, padding=p), nn.BatchNorm2d(cout),
                     nn.Hardtanh(0, 20, inplace=True)]
        self.conv = _wrap("seq_module", nn.Sequential(*mods))

        freq = 161
        n_for_size = 2 if self.streaming_model else conv_layers   # reference quirk: model.py:477-484
        for cin, cout, k, s, p in _CONV_SPECS[:n_for_size]:
            freq = (freq + 2 * p[0] - k[0]) // s[0] + 1
        rnn_input_size = freq * _CONV_SPECS[n_for_size - 1][1]

        uni = self.streaming_model or not bidirectional
        rnns = []
        for i in range(rnn_layers):
            layer = _Holder()
            in_size = rnn_input_size if i == 0 else rnn_hidden_size
            if i > 0:
                layer.add_module("batch_norm", _wrap("module", nn.BatchNorm1d(in_size)))
            layer.add_module("rnn", rnn_type(input_size=in_size, hidden_size=rnn_hidden_size,
                                             bidirectional=not uni, bias=True))
            rnns.append((str(i), layer))
        self.rnns = nn.Sequential(OrderedDict(rnns))

        if self.streaming_model:
            # streaming key: lookahead.conv.weight
            self.lookahead = _wrap("conv", nn.Conv1d(rnn_hidden_size, rnn_hidden_size, kernel_size=context, stride=1,
                                                     groups=rnn_hidden_size, padding=0, bias=False))
        elif not bidirectional:
            # offline key: lookahead.0.conv.weight
            self.lookahead = nn.Sequential(
                _wrap("conv", nn.Conv1d(rnn_hidden_size, rnn_hidden_size, kernel_size=context, stride=1,
                                        groups=rnn_hidden_size, padding=0, bias=False)),
                nn.Hardtanh(0, 20, inplace=True))
        else:
            self.lookahead = None

        self.fc = nn.Sequential(_wrap("module", nn.Sequential(
            nn.BatchNorm1d(rnn_hidden_size), nn.Linear(rnn_hidden_size, len(self.labels), bias=False))))

        self._handle = None
        self._handle_key = None
        self._workspace = None
        self._stream_state = None
        self._stream_key = None
        for p in self.parameters():
            p.requires_grad_(False)
        if self.streaming_model:
            self.forward = self.streaming_forward

    # ------------------------------------------------------------------ native handle management
    def _invalidate(self):
        L = N.lib() if (self._handle or self._stream_state) else None
        if self._stream_state:
            L.dsb_stream_state_destroy(self._stream_state)
        if self._handle:
            L.dsb_model_destroy(self._handle)
        self._handle = None
        self._handle_key = None
        self._stream_state = None
        self._stream_key = None

    def __del__(self):
        try:
            self._invalidate()
        except Exception:
            pass

    def _apply(self, fn, *a, **k):
        self._invalidate()
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._invalidate()
        return super().load_state_dict(*a, **k)

    def set_precision(self, precision):
        """'fp32' (CUDA-core kernels, 1e-4 parity) or 'bf16' (tcgen05 tensor-core kernels)."""
        if precision not in N.PRECISIONS:
            raise ValueError("precision must be one of %s" % (tuple(N.PRECISIONS),))
        if precision != self.precision:
            self.precision = precision
            self._invalidate()
        return self

    def _native(self):
        N.require_cuda()
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise N.NativeError("DeepSpeech weights are on %s; move the model to a CUDA device "
                                "(danspeech_b200 has no CPU path)" % dev)
        key = (dev.index, self.precision)
        if self._handle is not None and self._handle_key == key:
            return self._handle
        self._invalidate()
        L = N.lib()
        desc = N.ModelDesc(conv_layers=self.conv_layers, rnn_layers=self.rnn_layers,
                           rnn_hidden_size=self.rnn_hidden_size, rnn_type=N.RNN_TYPES[supported_rnns_inv[self.rnn_type]],
                           bidirectional=1 if self.bidirectional else 0, context=self.context,
                           num_classes=len(self.labels), streaming=1 if self.streaming_model else 0)
        h = N.c_void_p()
        N.check(L.dsb_model_create(desc, h), "dsb_model_create")
        keep = []
        try:
            with torch.cuda.device(dev):
                for name, t in self.state_dict().items():
                    if name.endswith("num_batches_tracked"):
                        continue
                    t = t.detach().to(device=dev, dtype=torch.float32).contiguous()
                    keep.append(t)
                    N.check(L.dsb_model_set_tensor(h, name.encode(), N.ptr(t), t.numel()), "dsb_model_set_tensor")
                N.check(L.dsb_model_finalize(h, N.PRECISIONS[self.precision], N.current_stream()), "dsb_model_finalize")
        except Exception:
            L.dsb_model_destroy(h)
            raise
        self._handle, self._handle_key = h, key
        return h

    def _get_workspace(self, nbytes, dev):
        if self._workspace is None or self._workspace.numel() < nbytes or self._workspace.device != dev:
            self._workspace = None
            self._workspace = torch.empty(int(nbytes), dtype=torch.uint8, device=dev)
        return self._workspace

    # ------------------------------------------------------------------ reference API
    def get_seq_lens(self, input_length):
        """model.py:540-551: only the first conv strides time (k=11, pad 5, stride 2)."""
        seq_len = input_length
        for (_, _, k, s, p) in _CONV_SPECS[: self.conv_layers]:
            seq_len = (seq_len + 2 * p[1] - (k[1] - 1) - 1) // s[1] + 1
        return seq_len.int()

    def forward(self, x, lengths):
        """x: cuda f32 [B,1,161,T] (zero padded after normalisation); lengths: int tensor [B], sorted descending.

        Returns (probs cuda f32 [B, T'max, C], output_lengths IntTensor[B] on CPU) -- model.py:496-515.
        """
        h = self._native()
        L = N.lib()
        if x.dim() != 4 or x.size(1) != 1 or x.size(2) != 161:
            raise ValueError("expected input of shape [B,1,161,T], got %s" % (tuple(x.shape),))
        dev = next(self.parameters()).device
        x = x.to(device=dev, dtype=torch.float32).contiguous()
        B, T = x.size(0), x.size(3)
        lens = [int(v) for v in lengths.cpu().int().tolist()]
        if any(lens[i] < lens[i + 1] for i in range(len(lens) - 1)):
            # the reference fails inside pack_padded_sequence (model.py:117)
            raise RuntimeError("`lengths` array must be sorted in decreasing order")
        Tp = L.dsb_model_out_frames(h, T)
        C = len(self.labels)
        with torch.cuda.device(dev):
            probs = torch.empty((B, Tp, C), dtype=torch.float32, device=dev)
            argmax = torch.empty((B, Tp), dtype=torch.int32, device=dev)
            ws = self._get_workspace(L.dsb_forward_workspace_bytes(h, B, T), dev)
            c_len = N.i32_array(lens)
            c_out = N.i32_array([0] * B)
            N.check(L.dsb_forward(h, N.ptr(x), c_len, B, T, N.ptr(probs), c_out, N.ptr(argmax), N.ptr(ws), ws.numel(),
                                  N.current_stream()), "dsb_forward")
        output_lengths = torch.IntTensor(list(c_out))
        t_max = int(output_lengths.max())
        probs = probs[:, :t_max]
        probs._dsb_argmax = argmax[:, :t_max]   # fused torch.max(probs, 2) for GreedyDecoder
        probs._dsb_status = self.check_status    # the decoders call it once they have synchronised for the results
        return probs, output_lengths

    def check_status(self):
        """dsb_forward does not synchronise; raises if a completed forward of this model aborted on the device."""
        if self._handle is not None:
            N.check(N.lib().dsb_forward_status(self._handle), "dsb_forward")

    def streaming_forward(self, x, is_first, is_last):
        """Chunked streaming forward (model.py:517-537); returns probs [S,k,C] or None while buffering."""
        h = self._native()
        L = N.lib()
        dev = next(self.parameters()).device
        x = x.to(device=dev, dtype=torch.float32).contiguous()
        S, k = x.size(0), x.size(3)
        with torch.cuda.device(dev):
            if self._stream_state is not None and self._stream_key[0] == S and k > self._stream_key[1] and not is_first:
                raise N.NativeError("streaming chunk of %d frames exceeds the %d frames this stream's state was sized "
                                    "for; start the stream (is_first) with the largest chunk or call "
                                    "reserve_stream_frames()" % (k, self._stream_key[1]))
            if self._stream_state is None or self._stream_key[0] != S or k > self._stream_key[1]:
                # (re)size the per-stream state: the reference accepts any chunk length (model.py:517-537)
                if self._stream_state is not None:
                    L.dsb_stream_state_destroy(self._stream_state)
                    self._stream_state = None
                max_k = max(512, getattr(self, "_stream_reserve", 0), k)
                st = N.c_void_p()
                N.check(L.dsb_stream_state_create(h, S, max_k, st), "dsb_stream_state_create")
                self._stream_state, self._stream_key = st, (S, max_k)
            k_max = L.dsb_stream_max_out_frames(self._stream_state, k)
            flat = torch.empty((S * max(k_max, 1) * len(self.labels),), dtype=torch.float32, device=dev)
            k_out = N.c_int32(0)
            N.check(L.dsb_streaming_forward(h, self._stream_state, N.ptr(x), k, 1 if is_first else 0,
                                            1 if is_last else 0, N.ptr(flat), k_out, N.current_stream()),
                    "dsb_streaming_forward")
        if k_out.value == 0:
            return None
        C = len(self.labels)
        return flat[: S * k_out.value * C].view(S, k_out.value, C)

    def reserve_stream_frames(self, max_chunk_frames):
        """Sizes the streaming state for chunks of up to ``max_chunk_frames`` spectrogram frames (default 512 = 5.1 s)."""
        self._stream_reserve = int(max_chunk_frames)
        return self

    # ------------------------------------------------------------------ (de)serialisation
    @classmethod
    def load_model(cls, path, trust_pickle=False):
        """model.py:599-624.  Reference packages are plain pickles of tensors, str, int, bool, list and (ordered)
        dict, which the restricted unpickler of ``weights_only=True`` accepts; ``trust_pickle=True`` opts into full
        unpickling for a package from a trusted source that carries other objects."""
        package = torch.load(path, map_location=lambda storage, loc: storage, weights_only=not trust_pickle)
        return cls.load_model_package(package)

    @classmethod
    def load_model_package(cls, package):
        """model.py:626-650."""
        model = cls(model_name=package["model_name"], rnn_hidden_size=package["rnn_hidden_size"],
                    rnn_layers=package["rnn_layers"], labels=package["labels"], audio_conf=package["audio_conf"],
                    rnn_type=supported_rnns[package["rnn_type"]], bidirectional=package["bidirectional"],
                    conv_layers=package["conv_layers"], context=package["context"],
                    streaming_inference_model=package["streaming_model"])
        model.load_state_dict(package["state_dict"])
        return model

    def serialize(self):
        """Package dict with the keys load_model expects (model.py:608-619)."""
        return dict(model_name=self.model_name, rnn_hidden_size=self.rnn_hidden_size, rnn_layers=self.rnn_layers,
                    labels=self.labels, audio_conf=self.audio_conf, rnn_type=supported_rnns_inv[self.rnn_type],
                    bidirectional=self.bidirectional, conv_layers=self.conv_layers, context=self.context,
                    streaming_model=self.streaming_model, state_dict=self.state_dict())

    @staticmethod
    def get_param_size(model):
        return sum(p.numel() for p in model.parameters())
