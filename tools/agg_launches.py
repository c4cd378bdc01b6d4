"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (us)."""
import collections
import csv
import sys


def aggregate(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = row["Kernel Name"].split("(")[0].replace("void ", "")
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
        a = agg.setdefault(k, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += v
        a[2] = max(a[2], v)
    return agg


if __name__ == "__main__":
    agg = aggregate(sys.argv[1])
    tot = sum(a[1] for a in agg.values())
    print(f"{'kernel':44s} {'n':>5s} {'total_us':>10s} {'avg_us':>9s} {'max_us':>9s} {'share':>6s}")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:44]:44s} {a[0]:5d} {a[1]:10.1f} {a[1] / a[0]:9.1f} {a[2]:9.1f} {a[1] / tot:6.3f}")
    print(f"{'TOTAL':44s} {sum(a[0] for a in agg.values()):5d} {tot:10.1f}")
