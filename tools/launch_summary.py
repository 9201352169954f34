"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, share."""
import collections
import csv
import re
import sys


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        try:
            t = float(row["Metric Value"].replace(",", ""))
        except (ValueError, KeyError):
            continue
        unit = row.get("Metric Unit", "us")
        t = t / 1000 if unit == "ns" else (t * 1000 if unit == "ms" else t)
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        name = re.sub(r"^void |<unnamed>::", "", name)[:64]
        agg[name][0] += 1
        agg[name][1] += t
    tot = sum(v[1] for v in agg.values())
    print(f"{'kernel':66s} {'n':>5s} {'total_us':>10s} {'avg_us':>8s} {'share':>6s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:66s} {v[0]:5d} {v[1]:10.1f} {v[1] / v[0]:8.1f} {100 * v[1] / tot:5.1f}%")
    print(f"{'TOTAL':66s} {sum(v[0] for v in agg.values()):5d} {tot:10.1f}")


if __name__ == "__main__":
    main(sys.argv[1])
