"""Prints the parity margins of every model fixture in both precision modes (run on the GPU box)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import __graft_entry__ as g  # noqa: E402
g.build()
from conftest import BATCH_CASES, batch_inputs, logit_rel_err  # noqa: E402
from danspeech_b200.deepspeech.decoder import GreedyDecoder  # noqa: E402
from danspeech_b200.pretrained_models import build_model  # noqa: E402
from danspeech_b200.utils import synthetic as syn  # noqa: E402

golden = np.load(os.path.join(ROOT, "tests", "golden", "reference_outputs.npz"))
dec = GreedyDecoder(syn.LABELS, blank_index=0)
_, x, xl = batch_inputs()
print("%-14s %-5s %-28s %s" % ("case", "mode", "logit rel err per utterance", "greedy == reference"))
for tag, name, kw in BATCH_CASES:
    kw = dict(kw)
    rt = kw.pop("rnn_type")
    for mode in ("fp32", "bf16"):
        m = build_model(name, seed=3, rnn_type=rt, **kw).cuda().eval().set_precision(mode)
        probs, sizes = m(x.cuda(), xl)
        ref = golden["batch_%s_probs" % tag]
        errs = [logit_rel_err(probs[b, :L].cpu().numpy(), ref[b, :L]) for b, L in enumerate(sizes.tolist())]
        texts = [s[0] for s in dec.decode(probs, sizes)[0]]
        same = [a == str(b) for a, b in zip(texts, golden["batch_%s_text" % tag])]
        print("%-14s %-5s %-28s %s" % (tag, mode, " ".join("%.1e" % e for e in errs), same))
for mode in ("fp32", "bf16"):
    m = build_model("TestModel", seed=0).cuda().eval().set_precision(mode)
    sp = torch.from_numpy(golden["spect_u0013002"]).cuda()
    probs, sizes = m(sp.view(1, 1, 161, -1), torch.IntTensor([sp.size(1)]))
    text = dec.decode(probs, sizes)[0][0][0]
    print("%-14s %-5s %-28s %s" % ("config1", mode, "%.1e" % logit_rel_err(probs.cpu().numpy(), golden["cfg1_probs"]),
                                    [text == str(golden["cfg1_text"])]))
