#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: the launches of the LAST
step (from the last launch whose name contains <first-kernel>) with their share of the step.
usage: launch_summary.py <launches.csv> <first-kernel-substring> [occurrence, default -1]
The step ends at the next launch of the first kernel (or at the FFMA peak probe / end of list)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
i = [k for k, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[i]
kn, mv, gs, bs = (hdr.index(x) for x in ("Kernel Name", "Metric Value", "Grid Size", "Block Size"))
data = rows[i + 1:]
occ = int(sys.argv[3]) if len(sys.argv) > 3 else -1
first = [k for k, r in enumerate(data) if sys.argv[2] in r[kn]][occ]
step = []
for n, r in enumerate(data[first:]):
    if "k_ffma_peak" in r[kn] or (n > 0 and sys.argv[2] in r[kn]):
        break
    step.append((float(r[mv].replace(",", "")) / 1e3, r[gs], r[bs], r[kn]))
tot = sum(t for t, *_ in step)
for t, g, b, k in step:
    print(f"{t:10.1f} us {100 * t / tot:5.1f}%  grid {g:14s} block {b:12s} {k[:100]}")
print(f"{tot:10.1f} us total (cold-cache, serialised: compare shares, not absolutes)")
