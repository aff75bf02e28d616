#!/usr/bin/env python
"""Per-kernel shares from an `ncu --metrics gpu__time_duration.sum --csv --log-file X` launch list."""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ki])
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1.0)
    tot[name] += v
    cnt[name] += 1
total = sum(tot.values())
print("# kernel, launches, total_us, avg_us, share")
for k, v in tot.most_common():
    print("%-90s %5d %12.1f %10.2f %6.1f%%" % (k[:90], cnt[k], v, v / cnt[k], 100 * v / total))
