#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: executed warp instructions by opcode and stall samples."""
import csv
import collections
import sys

path = sys.argv[1]
rows = list(csv.reader(open(path)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter()
stall = collections.Counter()
samples = collections.Counter()
total = 0
recs = []
for r in rows[2:]:
    if r and r[0] == "Address":
        continue
    if len(r) < len(hdr):
        continue
    src = r[ix["Source"]].strip()
    n = int(r[ix["Instructions Executed"]] or 0)
    s = int(r[ix["# Samples"]] or 0)
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
    base = op.split(".")[0]
    key = base
    if base in ("LDS", "STS", "LDG", "STG", "LDSM"):
        key = ".".join(op.split(".")[:3]) if "." in op else op
        key = base + "." + (op.split(".")[-1] if op.split(".")[-1] in ("64", "128", "U16", "S16", "U8") else "32")
    ops[key] += n
    samples[key] += s
    total += n
    recs.append((n, s, src))
    for h in hdr:
        if h.startswith("stall_") and "Not Issued" not in h:
            stall[h] += int(r[ix[h]] or 0)
print(f"total warp instructions executed: {total}")
for k, v in ops.most_common(30):
    print(f"  {k:14s} {v:12d} {100.0 * v / total:6.2f}%   samples {samples[k]}")
print("stall samples:")
tot_s = sum(stall.values())
for k, v in stall.most_common(12):
    print(f"  {k:28s} {v:8d} {100.0 * v / max(tot_s, 1):6.2f}%")
if len(sys.argv) > 2:
    print("hottest instructions by samples:")
    for n, s, src in sorted(recs, key=lambda t: -t[1])[: int(sys.argv[2])]:
        print(f"  {s:6d} {n:10d}  {src}")
