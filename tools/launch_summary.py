"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel (functor) name."""
import collections
import csv
import re
import sys

for f in sys.argv[1:]:
    with open(f) as fh:
        lines = [l for l in fh if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        n = row["Kernel Name"]
        m = re.search(r"spt::(\w+Kernel)", n)
        key = m.group(1) if m else re.sub(r"\(.*", "", n)[-40:]
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v *= {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}[u]
        agg[key][0] += 1
        agg[key][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f, "total ms %.3f" % tot)
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("  %-28s n=%5d  %10.3f ms  %5.1f%%  avg %.4f ms" % (k, v[0], v[1], 100 * v[1] / tot, v[1] / v[0]))
