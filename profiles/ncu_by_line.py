#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` SASS dump by CUDA source line.

usage: ncu_by_line.py <ncu_sass.csv> <cubin> <mangled-kernel-substring> [top] [outer | in:<file>]
The SASS instruction order of the report is aligned with `nvdisasm -gi` line markers.  With
`outer`, an instruction is charged to the OUTERMOST line of its inline chain (the line of the
kernel body that the inlined helper was called from) instead of the innermost one."""
import csv
import re
import subprocess
import sys
from collections import defaultdict

sass_csv, cubin, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
outer = len(sys.argv) > 5 and sys.argv[5] == "outer"
# "in:<file>": charge an instruction to the outermost line of its inline chain that lies in <file>
in_file = sys.argv[5][3:] if len(sys.argv) > 5 and sys.argv[5].startswith("in:") else None
dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout
# isolate the function
m = re.search(r"\.text\.[^\n]*" + re.escape(kern) + r"[^\n]*:\n", dis)
start = m.end()
nxt = re.search(r"\n\s*\.section|\n\.text\.", dis[start:])
body = dis[start:start + nxt.start()] if nxt else dis[start:]
lines = []  # (file, line) per instruction
cur = ("?", 0)
chain_open = False
for ln in body.splitlines():
    mm = re.search(r'//## File "([^"]+)", line (\d+)( inlined at)?', ln)
    if mm:
        # a chain is printed innermost first; each entry but the last says "inlined at"
        here = (mm.group(1).split("/")[-1], int(mm.group(2)))
        if in_file:
            if not chain_open:
                cur = here  # innermost entry: the default when the chain never enters in_file
            if here[0] == in_file:
                cur = here
        elif outer or not chain_open:
            cur = here
        chain_open = bool(mm.group(3))
        continue
    if re.search(r"/\*[0-9a-f]{4,}\*/\s+", ln) and ";" in ln:
        lines.append(cur)
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
n = min(len(data), len(lines))
if len(data) != len(lines):
    print(f"# warning: {len(data)} report instructions vs {len(lines)} disassembled", file=sys.stderr)
agg = defaultdict(lambda: [0, 0, 0, 0])
for (f, l), r in zip(lines[:n], data[:n]):
    a = agg[(f, l)]
    a[0] += int(r[ix["# Samples"]] or 0)
    a[1] += int(r[ix["Instructions Executed"]] or 0)
    a[2] += int(r[ix["Thread Instructions Executed"]] or 0)
    a[3] += 1
tot_s = sum(a[0] for a in agg.values()) or 1
tot_i = sum(a[1] for a in agg.values()) or 1
print(f"{'file:line':28s} {'samples%':>8s} {'inst%':>7s} {'avg_thr':>7s} {'sass':>5s}")
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{f + ':' + str(l):28s} {100 * a[0] / tot_s:8.2f} {100 * a[1] / tot_i:7.2f} "
          f"{(a[2] / a[1]) if a[1] else 0:7.1f} {a[3]:5d}")
