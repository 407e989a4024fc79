"""Oracle (test infrastructure): CPU restatement of the reference GreedyDecoder.

Follows /root/reference/danspeech/deepspeech/decoder.py: decode :183-198 (torch.max over classes),
convert_to_strings :151-164, process_string :166-181 (skip blank; skip a symbol equal to the previous
frame's symbol; space symbol emits ' ').  PINNED by tests/golden/gen_golden.py against the reference.
"""
import numpy as np


def greedy_decode(probs, sizes=None, labels="_abcdefghijklmnopqrstuvwxyzæøåéü ", blank_index=0):
    """probs [B,T,C] array-like -> (strings List[B][1], offsets List[B][1] int32 arrays)."""
    probs = np.asarray(probs)
    B, T, _ = probs.shape
    best = probs.argmax(axis=2)          # first index on ties, like torch.max
    strings, offsets = [], []
    for b in range(B):
        n = int(sizes[b]) if sizes is not None else T
        out, offs = [], []
        for i in range(n):
            s = int(best[b, i])
            if s == blank_index:
                continue
            if i != 0 and s == int(best[b, i - 1]):
                continue
            out.append(labels[s])
            offs.append(i)
        strings.append(["".join(out)])
        offsets.append([np.asarray(offs, dtype=np.int32)])
    return strings, offsets
