"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel name the
launch count, total and mean device time and the share of the captured window.
  python tools/ncu_summary.py gpurun_out/launches.csv [skip_first_n]
"""
import csv
import re
import sys
from collections import OrderedDict


def main(path, skip=0):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = val * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)
        rows.append((int(r["ID"]), r["Kernel Name"], ns))
    rows = rows[skip:]
    agg = OrderedDict()
    for _, name, ns in rows:
        short = re.sub(r"\(.*", "", name)
        a = agg.setdefault(short, [0, 0.0])
        a[0] += 1
        a[1] += ns
    total = sum(a[1] for a in agg.values())
    print("%d launches, %.1f us total device time (cold-cache, serialised by ncu)" % (len(rows), total / 1e3))
    print("%-70s %6s %10s %9s %7s" % ("kernel", "count", "total_us", "mean_us", "share"))
    for k, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-70s %6d %10.1f %9.2f %6.1f%%" % (k[:70], c, ns / 1e3, ns / c / 1e3, 100 * ns / total))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
