"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: count, total, average, share.
Usage: python tools/ncu_summary.py gpurun_out/launches.csv [> profiles/rNN_launches_summary.txt]"""
import collections
import csv
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = row["Kernel Name"].split("(")[0][:70]
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"{'kernel':72s} {'n':>6s} {'total us':>12s} {'avg us':>10s} {'share':>7s}")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:72s} {n:6d} {t:12.1f} {t / n:10.1f} {100 * t / tot:6.1f}%")
    print(f"{'TOTAL':72s} {sum(v[0] for v in agg.values()):6d} {tot:12.1f}")


if __name__ == "__main__":
    main(sys.argv[1])
