"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`):  python scripts/launch_summary.py gpurun_out/launches.csv"""
import csv
import sys
from collections import OrderedDict


def main(path):
    rows = [r for r in csv.reader(open(path)) if r and not r[0].startswith("==")]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0].replace("void ", "").replace("dsb::", "")
        v = float(r[vi].replace(",", ""))
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v * scale
    total = sum(a[1] for a in agg.values())
    print("| kernel | launches | total ms | share |")
    print("|---|---|---|---|")
    for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
        print("| `%s` | %d | %.3f | %.1f %% |" % (name[:70], n, ms, 100 * ms / total))
    print("\nTotal %.2f ms over %d launches" % (total, sum(a[0] for a in agg.values())))


if __name__ == "__main__":
    main(sys.argv[1])
