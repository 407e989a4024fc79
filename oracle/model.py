"""Oracle (test infrastructure): CPU restatement of the reference DeepSpeech forward passes.

Functional torch-CPU fp32 (optionally fp64) restatement driven directly by a reference-layout
state dict.  Follows /root/reference/danspeech/deepspeech/model.py:
  * get_seq_lens                       model.py:540-551
  * MaskConv.forward                   model.py:65-81   (conv -> BN -> Hardtanh, masked after every module)
  * DeepSpeech.forward                 model.py:496-515
  * BatchRNN.forward                   model.py:114-122 (BN1d, pack, rnn, unpack, sum directions)
  * Lookahead.forward + Hardtanh       model.py:143-148, :407-411
  * fc (BN1d + Linear, no bias)        model.py:414-420 ; InferenceBatchSoftmax :89-93
  * MaskConvStream / BatchRNNStream / LookaheadStream / streaming_forward   model.py:156-284, :517-537
PINNED: tests/golden/gen_golden.py checks this file against the unmodified reference modules.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

_CONV = [((41, 11), (2, 2), (20, 5)), ((21, 11), (2, 1), (10, 5)), ((21, 11), (2, 1), (10, 5))]
_RNN = {"gru": nn.GRU, "lstm": nn.LSTM, "rnn": nn.RNN}


def get_seq_lens(lengths, conv_layers):
    seq = lengths
    for (k, s, p) in _CONV[:conv_layers]:
        seq = (seq + 2 * p[1] - (k[1] - 1) - 1) // s[1] + 1
    return seq.int()


def _bn(x, sd, prefix):
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"], sd[prefix + ".weight"],
                        sd[prefix + ".bias"], training=False, eps=1e-5)


def _mask_time(x, lengths):
    # zero x[b, ..., t >= lengths[b]]   (model.py:73-80)
    T = x.size(-1)
    m = torch.arange(T).view(1, 1, 1, T) >= lengths.view(-1, 1, 1, 1)
    return x.masked_fill(m, 0)


def _make_rnn(sd, layer, rnn_type, bidirectional, dtype):
    w = sd["rnns.%d.rnn.weight_ih_l0" % layer]
    H = sd["rnns.%d.rnn.weight_hh_l0" % layer].shape[1]
    rnn = _RNN[rnn_type](input_size=w.shape[1], hidden_size=H, bidirectional=bidirectional, bias=True).to(dtype)
    rnn.load_state_dict({k[len("rnns.%d.rnn." % layer):]: v.to(dtype) for k, v in sd.items()
                         if k.startswith("rnns.%d.rnn." % layer)})
    return rnn.eval()


def conv_stack(sd, x, out_lengths, conv_layers, mask=True):
    for i, (k, s, p) in enumerate(_CONV[:conv_layers]):
        x = F.conv2d(x, sd["conv.seq_module.%d.weight" % (3 * i)], sd["conv.seq_module.%d.bias" % (3 * i)],
                     stride=s, padding=p)
        if mask:
            x = _mask_time(x, out_lengths)
        x = _bn(x, sd, "conv.seq_module.%d" % (3 * i + 1))
        if mask:
            x = _mask_time(x, out_lengths)
        x = F.hardtanh(x, 0, 20)
        if mask:
            x = _mask_time(x, out_lengths)
    return x


@torch.no_grad()
def forward(sd, x, lengths, conv_layers, rnn_layers, bidirectional=True, rnn_type="gru", context=20,
            dtype=torch.float32, return_intermediates=False):
    """x [B,1,161,T], lengths [B] sorted descending -> (probs [B,T',C], out_lengths)."""
    sd = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}
    x = x.to(dtype)
    inter = {}
    lengths = lengths.cpu().int()
    out_lengths = get_seq_lens(lengths, conv_layers)
    x = conv_stack(sd, x, out_lengths, conv_layers)
    B, C, D, T = x.shape
    x = x.view(B, C * D, T).transpose(1, 2).transpose(0, 1).contiguous()   # T x B x (C*D)
    inter["conv"] = x
    for l in range(rnn_layers):
        if l > 0:
            t, n = x.size(0), x.size(1)
            x = _bn(x.view(t * n, -1), sd, "rnns.%d.batch_norm.module" % l).view(t, n, -1)
        rnn = _make_rnn(sd, l, rnn_type, bidirectional, dtype)
        packed = nn.utils.rnn.pack_padded_sequence(x, out_lengths)
        y, _ = rnn(packed)
        x, _ = nn.utils.rnn.pad_packed_sequence(y)
        if bidirectional:
            x = x.view(x.size(0), x.size(1), 2, -1).sum(2).view(x.size(0), x.size(1), -1)
        inter["rnn%d" % l] = x
    if not bidirectional:
        w = sd["lookahead.0.conv.weight"]
        h = x.transpose(0, 1).transpose(1, 2)
        h = F.pad(h, pad=(0, context - 1), value=0)
        h = F.conv1d(h, w, groups=w.shape[0])
        x = F.hardtanh(h.transpose(1, 2).transpose(0, 1).contiguous(), 0, 20)
    t, n = x.size(0), x.size(1)
    h = _bn(x.view(t * n, -1), sd, "fc.0.module.0")
    logits = F.linear(h, sd["fc.0.module.1.weight"]).view(t, n, -1).transpose(0, 1)
    inter["logits"] = logits
    probs = F.softmax(logits, dim=-1)
    if return_intermediates:
        return probs, out_lengths, inter
    return probs, out_lengths


class StreamingOracle:
    """Stateful restatement of the streaming model (2-conv, uni-directional)."""

    def __init__(self, sd, rnn_layers, rnn_type="gru", context=20, dtype=torch.float32):
        self.sd = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}
        self.rnn_layers = rnn_layers
        self.context = context
        self.dtype = dtype
        self.rnns = [_make_rnn(self.sd, l, rnn_type, False, dtype) for l in range(rnn_layers)]
        self.left = [None, None]
        self.hidden = [None] * rnn_layers
        self.look = None

    @torch.no_grad()
    def forward(self, x, is_first, is_last):
        sd = self.sd
        x = x.to(self.dtype)
        for i, (k, s, p) in enumerate(_CONV[:2]):
            # MaskConvStream.forward (model.py:169-201)
            if is_first:
                x = F.pad(x, pad=(5, 0), value=0)
            elif is_last:
                x = F.pad(x, pad=(0, 5), value=0)
            if not is_first:
                x = torch.cat([self.left[i], x], dim=3)
            if not is_last:
                self.left[i] = x[:, :, :, -10:]
            x = F.conv2d(x, sd["conv.seq_module.%d.weight" % (3 * i)], sd["conv.seq_module.%d.bias" % (3 * i)],
                         stride=s, padding=p)
            x = _bn(x, sd, "conv.seq_module.%d" % (3 * i + 1))
            x = F.hardtanh(x, 0, 20)
        B, C, D, T = x.shape
        x = x.view(B, C * D, T).transpose(1, 2).transpose(0, 1).contiguous()
        for l in range(self.rnn_layers):   # BatchRNNStream.forward (model.py:219-237)
            if l > 0:
                t, n = x.size(0), x.size(1)
                x = _bn(x.view(t * n, -1), sd, "rnns.%d.batch_norm.module" % l).view(t, n, -1)
            x, h = self.rnns[l](x) if self.hidden[l] is None else self.rnns[l](x, self.hidden[l])
            self.hidden[l] = None if is_last else h
        # LookaheadStream.forward (model.py:255-279)
        if self.look is None or is_first:
            self.look = x
            return None
        out = torch.cat([self.look, x], dim=0)
        self.look = x[-(self.context - 1):, :, :]
        out = out.transpose(0, 1).transpose(1, 2)
        if is_last:
            out = F.pad(out, pad=(0, self.context - 1), value=0)
        w = sd["lookahead.conv.weight"]
        out = F.conv1d(out, w, groups=w.shape[0])
        out = F.hardtanh(out.transpose(1, 2).transpose(0, 1).contiguous(), 0, 20)
        if is_last:
            self.look = None
        t, n = out.size(0), out.size(1)
        h = _bn(out.view(t * n, -1), sd, "fc.0.module.0")
        logits = F.linear(h, sd["fc.0.module.1.weight"]).view(t, n, -1).transpose(0, 1)
        return F.softmax(logits, dim=-1)
