#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total and share per ww:: kernel."""
import collections
import csv
import re
import sys

path = sys.argv[1]
rows = [r for r in csv.reader(open(path)) if len(r) > 14]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    name = r[ki]
    if "ww::" not in name and not re.match(r"(void )?(mfcc|cnn|ctc|cmvn|tdm|augment)", name):
        continue
    name = re.sub(r"\(.*", "", name.replace("void ", ""))
    if not name.startswith("ww::"):
        name = "ww::" + name
    t = float(r[vi].replace(",", ""))
    t_us = {"ns": t / 1e3, "us": t, "ms": t * 1e3, "s": t * 1e6}.get(r[ui], t / 1e3)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += t_us
tot = sum(a[1] for a in agg.values())
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-40s launches %4d  total %9.3f ms  share %5.1f%%  avg %9.1f us" % (name, n, t / 1e3, 100 * t / tot, t / n))
