#!/bin/bash
# "young tier" experiment: tier 0 = G=2 with small caps, tier 1 = the 40/24 group kernel resuming it
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for name in "$@"; do
  export CB2_LIB_PATH=$PWD/collision_rs_b200/variants/libcb2_$name.so
  echo "== $name"
  timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
  timeout 300 python bench.py --steps 5 --warmup 3 --no-extra 2>/dev/null | \
      python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$name', round(d['value']/1e6,1), 'Mpairs/s', {k: round(v,2) for k,v in d['roofline']['stage_ms'].items()}, 'non_ok', d['run']['non_ok_status'])"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_epa --csv --log-file gpurun_out/m_$name.csv python tools/profile_pass.py cfg3 > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/m_$name.csv')) if len(r)>5 and r[0].isdigit()]
for r in rows: print('   ', r[4][:60], r[-1], r[-2])
PY
done
