"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` log: per kernel name count / total / min / median / max (us)."""
import csv, sys, collections, statistics
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
r = csv.DictReader(lines)
d = collections.defaultdict(list)
for row in r:
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    u = row.get("Metric Unit", "ns")
    us = v / 1000.0 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1000.0)
    name = row["Kernel Name"].split("(")[0][:48]
    d[name].append(us)
tot = sum(sum(v) for v in d.values())
print("total %.1f us over %d launches" % (tot, sum(len(v) for v in d.values())))
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    print("%-48s n=%4d total %9.1f us  min %7.2f  med %7.2f  max %8.2f" % (k, len(v), sum(v), min(v), statistics.median(v), max(v)))
