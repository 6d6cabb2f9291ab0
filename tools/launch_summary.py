"""Developer tool: aggregate an ncu launch list (gpu__time_duration.sum, --csv) per kernel."""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    agg.setdefault(row["Kernel Name"][:48], []).append(float(row["Metric Value"].replace(",", "")))
tot = sum(sum(v) for v in agg.values())
for k, v in agg.items():
    print(f"{k:48s} n={len(v):3d} mean={sum(v)/len(v)/1e3:8.1f} us  share={sum(v)/tot*100:5.1f}%")
