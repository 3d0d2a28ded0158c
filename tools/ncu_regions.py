#!/usr/bin/env python
"""Executed-instruction and stall-sample shares of consecutive SASS regions of one kernel in an .ncu-rep
(regions = runs of instructions with the same execution count):  python tools/ncu_regions.py rep.ncu-rep kernel_name"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--kernel-name", sys.argv[2]], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
start = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[start]
ie, isrc, isamp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
data = []
for r in rows[start + 1:]:
    if len(r) <= ie or not r[ie].isdigit():
        if r and r[0] == "Kernel Name" and data:
            break   # the first captured launch only
        continue
    data.append((len(data), r[isrc].strip(), int(r[ie]), int(r[isamp])))
tot, ts = sum(d[2] for d in data), sum(d[3] for d in data)
print("instructions", tot, "samples", ts)
groups = []
for d in data:
    if groups and groups[-1][2] == d[2]:
        groups[-1][1] += 1; groups[-1][3] += d[2]; groups[-1][4] += d[3]
    else:
        groups.append([d[0], 1, d[2], d[2], d[3], d[1]])
for g in groups:
    if g[3] > tot * 0.003 or g[4] > ts * 0.01:
        print(f"idx {g[0]:5d} n={g[1]:4d} exec/inst={g[2]:9d} inst={100*g[3]/tot:5.1f}% samples={100*g[4]/ts:5.1f}%  first: {g[5][:60]}")
if len(sys.argv) > 3:
    top = sorted(data, key=lambda d: -d[3])[: int(sys.argv[3])]
    for d in sorted(top):
        print(d[0], f"{100*d[3]/ts:.1f}%", d[1][:70])
