"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: total time and share per kernel.

    python tools/launch_shares.py gpurun_out/launches.csv [--skip N] > profiles/rNN_launch_shares.txt
"""
import csv
import re
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = re.sub(r"^void |sam3b::|\(anonymous namespace\)::|<unnamed>::", "", name)
        rows.append((name, us))
    tot = sum(u for _, u in rows)
    agg = defaultdict(lambda: [0.0, 0])
    for n, u in rows:
        agg[n][0] += u
        agg[n][1] += 1
    print(f"# total {tot / 1000:.1f} ms over {len(rows)} launches (per-launch times under ncu are cold-cache + serialised: compare SHARES)")
    for n, (u, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{u / 1000:9.2f} ms {100 * u / tot:5.1f}%  n={c:4d}  avg {u / c:8.1f} us  {n}")


if __name__ == "__main__":
    main()
