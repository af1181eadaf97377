#!/usr/bin/env python
"""Per-source-line instruction / stall-sample shares from `ncu --page source --print-source cuda,sass --csv`."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
fname = ""
data = []
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ci, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr and r[0].isdigit() and len(r) > ci and r[ci].replace(".", "").isdigit():
        data.append((float(r[ci]), float(r[si]) if r[si].replace(".", "").isdigit() else 0.0, fname, r[0], r[1].strip()[:100]))
tot = sum(d[0] for d in data) or 1
tots = sum(d[1] for d in data) or 1
print("total warp-instructions %.4g, samples %d" % (tot, tots))
data.sort(reverse=True)
for d in data[:top]:
    print("%5.1f%% inst %5.1f%% stall  %s:%s  %s" % (100 * d[0] / tot, 100 * d[1] / tots, d[2], d[3], d[4]))
