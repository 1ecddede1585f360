#!/bin/bash
# launch list of one SAP call (profile_pass cfg4)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r_launches_cfg4.csv python tools/profile_pass.py cfg4 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r_launches_cfg4.csv')) if len(r)>5 and r[0].isdigit()]
tot=0
for r in rows[:40]:
    print('%9.1f us  %s'%(float(r[-1])/1e3, r[4][:70]))
PY
