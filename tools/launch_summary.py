"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.

    python tools/launch_summary.py launches.csv [top] [regex of kernel names to keep]
"""
import collections
import csv
import re
import sys


def main(path, top=25, only=None):
    rows = list(csv.reader(open(path, errors="replace")))
    hdr = None
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            try:
                v = float(d["Metric Value"].replace(",", ""))
            except ValueError:
                continue
            if d.get("Metric Unit", "ns") in ("us", "usecond"):
                v *= 1e3
            elif d.get("Metric Unit", "ns") in ("ms", "msecond"):
                v *= 1e6
            k = d["Kernel Name"].split("(")[0][:70]
            if only and not re.search(only, k):
                continue
            agg[k][0] += 1
            agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"total {tot / 1e6:.3f} ms over {sum(v[0] for v in agg.values())} launches")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
        print(f"{k:70s} n={v[0]:5d} total={v[1] / 1e6:9.3f} ms avg={v[1] / v[0] / 1e3:9.1f} us share={v[1] / tot:.3f}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25, sys.argv[3] if len(sys.argv) > 3 else None)
