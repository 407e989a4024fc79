"""Per-kernel SASS evidence for profiles/: counts of the Blackwell-native mnemonics (UTC*MMA = tcgen05.mma, LDTM/STTM =
tcgen05.ld/st, UTMALDG/UTMASTG/UBLKCP = TMA / bulk copies, SYNCS = mbarrier, UTCBAR = tcgen05.commit, legacy HMMA must
be 0) plus registers / spills / static shared memory from `cuobjdump -res-usage`.  Runs without a GPU:
    python scripts/sass_summary.py > profiles/rNN_sass_summary.md"""
import collections
import os
import re
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SO = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "danspeech_b200", "lib", "libdanspeech_b200.so")
PAT = [("UTC*MMA", re.compile(r"\bUTC[A-Z]*MMA")), ("UTCBAR", re.compile(r"\bUTCBAR")), ("LDTM", re.compile(r"\bLDTM")),
       ("STTM", re.compile(r"\bSTTM")), ("UTMALDG", re.compile(r"\bUTMALDG")), ("UTMASTG", re.compile(r"\bUTMASTG")),
       ("UBLKCP", re.compile(r"\bUBLKCP")), ("SYNCS", re.compile(r"\bSYNCS")), ("HMMA", re.compile(r"\bHMMA")),
       ("STG.256", re.compile(r"\bSTG\.E\.ENL2\.256")), ("MUFU", re.compile(r"\bMUFU")), ("DFMA", re.compile(r"\bDFMA")),
       ("RED/ATOM", re.compile(r"\b(REDG?|ATOMG?)\b|\bRED\.|\bATOMG\."))]


def demangle(names):
    out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def short(n):
    n = re.sub(r"^void ", "", n)
    depth = 0
    for i, ch in enumerate(n):          # cut the parameter list: the first '(' outside the template brackets
        if ch == "<":
            depth += 1
        elif ch == ">":
            depth -= 1
        elif ch == "(" and depth == 0:
            n = n[:i]
            break
    n = n.replace("(int)", "").replace("(bool)", "").replace("(dsb::SpectMode)", "")
    n = n.replace("(anonymous namespace)::", "")
    return n


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    counts, total, cur = collections.OrderedDict(), {}, None
    for ln in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            total[cur] = 0
            continue
        if cur is None or not re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln):
            continue
        total[cur] += 1
        for k, p in PAT:
            if p.search(ln):
                counts[cur][k] += 1
    res = subprocess.run(["cuobjdump", "-res-usage", SO], capture_output=True, text=True).stdout
    usage, cur = {}, None
    for ln in res.splitlines():
        m = re.match(r"\s*Function (\S+):", ln)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"REG:(\d+).*?SHARED:(\d+).*?LOCAL:(\d+)", ln)
        if m and cur:
            usage[cur] = (int(m.group(1)), int(m.group(2)), int(m.group(3)))
    dm = demangle(list(counts))
    print("# SASS summary of libdanspeech_b200.so (sm_100a), from `cuobjdump -sass` / `-res-usage`\n")
    print("`UTC*MMA` = tcgen05.mma, `UTCBAR` = tcgen05.commit, `LDTM`/`STTM` = tcgen05.ld/st, `UTMALDG` = TMA tensor load, "
          "`UBLKCP` = cp.async.bulk, `SYNCS` = mbarrier ops, `STG.256` = 256-bit global stores; legacy `HMMA` must be 0. "
          "LOCAL = bytes of local memory per thread (stack/spills); static SMEM includes the 1 KB the toolchain reserves per CTA "
          "on sm_100 (the tensor-core kernels use dynamic shared memory on top).\n")
    keys = [k for k, _ in PAT]
    print("| kernel | SASS instr | REG | static SMEM | LOCAL | " + " | ".join(keys) + " |")
    print("|---|---:|---:|---:|---:|" + "---:|" * len(keys))
    rows = sorted(counts, key=lambda f: short(dm.get(f, f)))
    for f in rows:
        u = usage.get(f, ("?", "?", "?"))
        c = counts[f]
        print("| `%s` | %d | %s | %s | %s | " % (short(dm.get(f, f)), total[f], u[0], u[1], u[2]) +
              " | ".join(str(c[k]) if c[k] else "" for k in keys) + " |")
    tc = [f for f in rows if counts[f]["UTC*MMA"]]
    print("\n%d kernels, %d of them issue tcgen05.mma, %d use TMA / bulk copies, %d legacy HMMA instructions in the library." % (
        len(rows), len(tc), sum(1 for f in rows if counts[f]["UTMALDG"] or counts[f]["UBLKCP"]),
        sum(counts[f]["HMMA"] for f in rows)))


if __name__ == "__main__":
    main()
