#!/usr/bin/env python
"""Per-kernel totals of one A3C iteration from an ncu launch list (gpu__time_duration.sum, --csv):
    python profiles/launch_summary.py <launches.csv> [iteration index]"""
import collections
import csv
import re
import sys

rows = []
with open(sys.argv[1]) as f:
    for x in csv.DictReader(l for l in f if l.startswith('"')):
        rows.append((x["Kernel Name"], float(x["Metric Value"])))
marks = [i for i, (n, _) in enumerate(rows) if "rmsprop" in n]
it = int(sys.argv[2]) if len(sys.argv) > 2 else 0
seg = rows[marks[it] + 1:marks[it + 1] + 1]
agg = collections.defaultdict(lambda: [0, 0.0])
for n, v in seg:
    n = re.sub(r"\(.*", "", n)[:100]
    agg[n][0] += 1
    agg[n][1] += v
tot = sum(v for _, v in agg.values())
print("launches %d, total %.1f us" % (len(seg), tot / 1e3))
for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%8.1f us %5.1f%% x%3d  %s" % (v / 1e3, 100 * v / tot, c, n))
if "-v" in sys.argv:
    for n, v in seg:
        print("%8.1f  %s" % (v / 1e3, re.sub(r"\(.*", "", n)[:100]))
