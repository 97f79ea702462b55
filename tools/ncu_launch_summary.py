"""Per-kernel shares of an ncu launch list (`--metrics gpu__time_duration.sum --csv`).  usage: ncu_launch_summary.py launches.csv "header line" > summary.txt"""
import csv
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    iK, iN, iV, iU = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[1:]:
        if r[iN] != "gpu__time_duration.sum":
            continue
        v = float(r[iV].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iU].replace("second", "s").replace("nsecond", "ns"), 1e-6)
        k = r[iK].split("(")[0]
        tot[k] += v; cnt[k] += 1
    total = sum(tot.values())
    if len(sys.argv) > 2:
        print(sys.argv[2])
    print("per-launch times are cold-cache and serialised (under ncu the parts do not overlap): compare SHARES, not absolutes")
    for k in sorted(tot, key=lambda k: -tot[k]):
        print(f"{k[:60]:60s} launches {cnt[k]:4d} total {tot[k]:9.3f} ms share {100 * tot[k] / total:5.1f}%")


if __name__ == "__main__":
    main()
