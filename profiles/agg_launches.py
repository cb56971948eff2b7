"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.

    python profiles/agg_launches.py gpurun_out/launches.csv
"""
import collections
import csv
import sys


def main(path, width=70):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg, tot = collections.defaultdict(lambda: [0, 0.0]), 0.0
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (ValueError, KeyError):
            continue
        unit = row["Metric Unit"]
        v = v / 1000 if unit in ("ns", "nsecond") else (v * 1000 if unit in ("ms", "msecond") else v)
        name = row["Kernel Name"][:width]
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
    print(f"{'total us':>10} {'n':>5} {'avg us':>8} {'share':>6}  kernel")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t:10.1f} {n:5d} {t / n:8.2f} {100 * t / tot:5.1f}%  {k}")
    print(f"{tot:10.1f} total")


if __name__ == "__main__":
    main(sys.argv[1])
