"""Per-kernel SASS opcode summary of libsam3b.so (cuobjdump -sass): counts of the Blackwell-specific instructions
(UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA load / store, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit,
UTCCP, SYNCS = mbarrier), MUFU, and the total instruction count.  Written to profiles/ as evidence that the hot kernels
run on tcgen05 + TMA.    python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
KEYS = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "MUFU", "HMMA", "LDGSTS", "ATOMG", "RED"]


def main():
    so = ROOT / "sam3_lora_b200" / "libsam3b.so"
    out = subprocess.run(["cuobjdump", "-sass", str(so)], capture_output=True, text=True, check=True).stdout
    demangle = {}
    kernels = collections.OrderedDict()
    cur = None
    for ln in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
        if m:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            for k in KEYS:
                if op.startswith(k):
                    kernels[cur][k] += 1
    names = list(kernels)
    try:
        dm = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True, check=True).stdout.splitlines()
        demangle = dict(zip(names, dm))
    except (OSError, subprocess.CalledProcessError):
        demangle = {n: n for n in names}
    print(f"# {so.name}: {len(kernels)} kernels (sm_100a).  Columns: " + " ".join(KEYS) + " | total instructions")
    tot = collections.Counter()
    for n, c in kernels.items():
        short = demangle[n].replace("(anonymous namespace)::", "").replace("sam3b::", "")
        short = re.sub(r"^void ", "", re.sub(r"\(.*", "", short))
        cols = " ".join(f"{c.get(k, 0):5d}" for k in KEYS)
        print(f"{short[:70]:70s} {cols} | {c['_total']:6d}")
        tot.update(c)
    print(f"{'TOTAL':70s} " + " ".join(f"{tot.get(k, 0):5d}" for k in KEYS) + f" | {tot['_total']:6d}")


if __name__ == "__main__":
    sys.exit(main())
