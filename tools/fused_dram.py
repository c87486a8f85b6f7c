#!/usr/bin/env python
"""DRAM bytes per clip of the fused scorer from an ncu launch list (no GPU needed):

  ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
      -k regex:'mfcc_kernel|cnn_' --csv --log-file gpurun_out/fused_dram.csv python tools/prof_driver.py <clips> fused tensor

`--cache-control none` matters: the default flushes L2 before every kernel, which is exactly the hand-off being measured.
Usage: python tools/fused_dram.py gpurun_out/fused_dram.csv <clips> <iterations of prof_driver (3)>"""
import csv
import sys
from collections import defaultdict

path, clips, iters = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 3
rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
hdr = rows[0]
iname, imetric, ivalue, iunit = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1, "usecond": 1e-6,
         "nsecond": 1e-9, "msecond": 1e-3, "second": 1}
agg = defaultdict(lambda: defaultdict(float))
count = defaultdict(int)
for r in rows[1:]:
    k = r[iname].split("(")[0].split("<")[0].replace("void ", "").replace("ww::", "")
    v = float(r[ivalue].replace(",", "")) * scale.get(r[iunit], 1)
    agg[k][r[imetric]] += v
    if r[imetric] == "gpu__time_duration.sum":
        count[k] += 1
tot_r = tot_w = 0.0
for k in sorted(agg):
    m = agg[k]
    rd, wr = m.get("dram__bytes_read.sum", 0), m.get("dram__bytes_write.sum", 0)
    tot_r += rd
    tot_w += wr
    print(f"{k:28s} launches {count[k]:5d}  time {m.get('gpu__time_duration.sum', 0) * 1e3:9.3f} ms  "
          f"DRAM read {rd / 1e9:8.3f} GB  write {wr / 1e9:8.3f} GB")
n = clips * iters
print(f"total over {iters} x {clips} clips: read {tot_r / n:9.1f} B/clip, write {tot_w / n:8.1f} B/clip, "
      f"sum {(tot_r + tot_w) / n:9.1f} B/clip  (algorithmic 32 005 B/clip: ratio {(tot_r + tot_w) / n / 32005:.3f})")
