"""TMEM accumulator layout probe (diagnostic): prints which (row m, column n) of D each TMEM lane / column holds
for a few tcgen05.mma shapes, including the cta_group::2 M=128 case (64 rows per CTA)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "microbench"))
import build as _mb  # noqa: E402


def rng(a):
    a = [int(x) for x in a]
    if not a:
        return "none"
    out, s0, prev = [], a[0], a[0]
    for x in a[1:]:
        if x != prev + 1:
            out.append("%d-%d" % (s0, prev))
            s0 = x
        prev = x
    out.append("%d-%d" % (s0, prev))
    return ",".join(out)


L = ctypes.CDLL(_mb.build())
torch.zeros(1).cuda()
for cg, M, Nn in ((1, 64, 64), (2, 128, 64), (2, 128, 128), (2, 256, 64)):
    buf = np.zeros((cg, 128, 128), np.float32)
    rc = L.dsb_debug_layout_probe(cg, M, Nn, buf.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
    print("== cta_group::%d M=%d N=%d rc=%d" % (cg, M, Nn, rc))
    if rc:
        continue
    for rank in range(cg):
        v = buf[rank]
        w = v >= 0
        m = (v.astype(np.int64) % 256) - 1
        n = (v.astype(np.int64) // 256) - 1
        print(" rank %d: written lanes %s | columns %s" % (rank, rng(np.where(w.any(axis=1))[0]), rng(np.where(w.any(axis=0))[0])))
        for lane in (0, 1, 15, 16, 31, 32, 47, 48, 63, 64, 79, 96, 127):
            if w[lane].any():
                c = np.where(w[lane])[0]
                print("   lane %3d: cols %s -> m=%s n=%d..%d (n step %d)" % (
                    lane, rng(c), sorted(set(m[lane][c].tolist())), n[lane][c[0]], n[lane][c[-1]],
                    (n[lane][c[1]] - n[lane][c[0]]) if len(c) > 1 else 0))
