"""tcgen05.mma issue-cost micro-benchmark (diagnostic).  Prints cycles per MMA for several shapes/variants."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "microbench"))
import build as _mb  # noqa: E402

L = ctypes.CDLL(_mb.build())
torch.zeros(1).cuda()
out = (ctypes.c_longlong * 2)()
n = 1024
print("%-10s %-6s %-44s %10s %12s" % ("shape", "group", "variant", "issue/mma", "complete/mma"))
names = {0: "lane0: mma only", 7: "lane0: +commit+fence+try_wait", 8: "elect: mma only", 15: "elect: +commit+fence+try_wait",
         24: "elect: A from TMEM, mma only", 31: "elect: A from TMEM, +commit+fence+try_wait"}
TS_ONLY = os.environ.get("TS_ONLY", "0") == "1"
for (M, Nn) in ((128, 64), (128, 32), (128, 96), (128, 128), (128, 256), (64, 64)):
    for group in ((16,) if TS_ONLY else (1, 4, 16)):
        for variant in ((8, 15, 24, 31) if TS_ONLY else (0, 7, 8, 15, 24, 31)):
            rc = L.dsb_debug_mma_bench(M, Nn, n, group, variant, out)
            if rc != 0:
                print('%dx%dx16 group %d variant %d: rc=%d' % (M, Nn, group, variant, rc), flush=True)
                continue
            print("%-10s %-6d %-44s %10.1f %12.1f" % ("%dx%dx16" % (M, Nn), group, names[variant], out[0] / n, out[1] / n))
