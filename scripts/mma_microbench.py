"""tcgen05.mma issue-cost micro-benchmark (diagnostic).  Prints cycles per MMA for several shapes/variants."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import __graft_entry__ as g  # noqa: E402
g.build()
from danspeech_b200 import _native as N  # noqa: E402

L = ctypes.CDLL(N.lib_path())
torch.zeros(1).cuda()
out = (ctypes.c_longlong * 2)()
n = 1024
print("%-10s %-6s %-34s %10s %12s" % ("shape", "group", "variant", "issue/mma", "complete/mma"))
names = {0: "lane0: mma only", 7: "lane0: +commit+fence+try_wait", 8: "elect: mma only", 15: "elect: +commit+fence+try_wait"}
for (M, Nn) in ((64, 64), (128, 64), (128, 32), (128, 96), (128, 128), (128, 256)):
    for group in (1, 4, 16):
        for variant in (0, 7, 8, 15):
            rc = L.dsb_debug_mma_bench(M, Nn, n, group, variant, out)
            assert rc == 0, rc
            print("%-10s %-6d %-34s %10.1f %12.1f" % ("%dx%dx16" % (M, Nn), group, names[variant], out[0] / n, out[1] / n))
