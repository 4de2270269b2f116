"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / total / share."""
import csv, sys, collections, re
path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
tot = collections.OrderedDict()
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    t = tot.setdefault(name, [0, 0.0])
    t[0] += 1; t[1] += us
total = sum(t[1] for t in tot.values())
print("total us %.1f" % total)
for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-90s n=%5d  total %9.1f us  avg %8.1f us  share %5.1f%%" % (k[:90], n, us, us / n, 100 * us / total))
