"""TMA intake micro-benchmark (diagnostic): cycles per box and bytes/clk/SM for the recurrence's h stream."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "microbench"))
import torch  # noqa: E402
import build as _mb  # noqa: E402

L = ctypes.CDLL(_mb.build())
torch.zeros(1).cuda()
out = (ctypes.c_longlong * 2)()
names = {0: "3-D tensor boxes", 1: "1-D bulk copies (pre-tiled)", 2: "tensor boxes, cluster-2 multicast", 3: "tensor boxes, two issuing warps"}
n_box = 400
print("%-36s %4s %5s %5s %10s %12s" % ("mode", "gsz", "depth", "grid", "cyc/box", "B/clk/SM"))
for grid in (1, 120):
    for mode in (0, 1, 2, 3):
        for gsz, depth in ((1, 2), (1, 8), (2, 2), (2, 4), (4, 1), (4, 2), (4, 4), (8, 1), (8, 2)):
            if mode == 3 and depth % 2:
                continue        # two issuing warps take alternate slots
            g = grid if mode != 2 else max(2, grid // 2 * 2)
            rc = L.dsb_debug_tma_bench(mode, gsz, depth, g, n_box, out)
            if rc != 0:
                print("%-36s %4d %5d %5d rc=%d" % (names[mode], gsz, depth, g, rc))
                continue
            per_box = out[1] / n_box
            # multicast: every CTA receives all boxes but issues half of them
            print("%-36s %4d %5d %5d %10.0f %12.1f" % (names[mode], gsz, depth, g, per_box, gsz * 8192 / per_box), flush=True)
