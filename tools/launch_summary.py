"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
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
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = {"ns": v, "us": v * 1e3, "usecond": v * 1e3, "ms": v * 1e6}.get(unit, v)
        agg.setdefault(row["Kernel Name"].split("(")[0][-48:], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f"{'kernel':50s} {'n':>5s} {'mean_us':>10s} {'min_us':>10s} {'max_us':>10s} {'share':>7s}")
    for k, v in agg.items():
        print(f"{k:50s} {len(v):5d} {sum(v)/len(v)/1e3:10.2f} {min(v)/1e3:10.2f} {max(v)/1e3:10.2f} {sum(v)/tot:7.1%}")


if __name__ == "__main__":
    main(sys.argv[1])
