"""Per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv)."""
import collections
import csv
import re
import sys

for path in sys.argv[1:]:
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print("%s: %d launches, %.1f us in total" % (path, sum(v[0] for v in agg.values()), tot))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("  %-60s n=%5d avg=%9.2f us share=%.3f" % (k[:60], v[0], v[1] / v[0], v[1] / tot))
